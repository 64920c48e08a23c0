#!/bin/bash
# End-of-round-2 validation at HEAD (trimmed to the GPU minutes left): GPU parity suite, smoke, c2 / reference / c3 bench lines,
# launch lists of one eager c2 / c3 step, ncu --set full of the c2 top and tower kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  |pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
timeout 500 python bench.py --config c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json
for n in ("c2", "reference", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(n, "ms/step %.4f value %.3e e2e %s roofline %s %s" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("ms_per_step"),
              r.get("frac"), (r.get("kernel") or "")[:40]), "| gather", (d.get("roofline_gather") or {}).get("frac"),
              "| fp32", (d.get("value_fp32") or {}).get("ms_per_step"), "| keras_dense", (d.get("value_keras_dense") or {}).get("ms_per_step"))
    except Exception as e:
        print(n, "no line:", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fm_tile_kernel|gather_fm_fwd_lean" -s 6 -c 2 \
    -o gpurun_out/prof_c2_top -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full1.log 2>&1
echo "ncu full c2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_skinny_bwd|deepfm_tail" -s 6 -c 4 \
    -o gpurun_out/prof_c2_tower -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full4.log 2>&1
echo "ncu full tower exit $?"
