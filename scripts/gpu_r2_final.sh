#!/bin/bash
# Final round-2 validation at HEAD: GPU parity suite, smoke, bench lines of every config, launch lists, ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  |pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"
tail -c 1200 gpurun_out/bench_c2.json | head -c 600; echo; tail -3 gpurun_out/bench_c2.err
timeout 300 python bench.py --dist uniform --no-cpu-baseline --no-extras > gpurun_out/bench_c2_uniform.json 2> gpurun_out/bench_c2_uniform.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
for c in c3 c4 c1; do
  timeout 500 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
  tail -3 gpurun_out/bench_$c.err
done
ETR_C4_MODEL=fwfm timeout 300 python bench.py --config c4 --no-cpu-baseline > gpurun_out/bench_c4_fwfm.json 2> gpurun_out/bench_c4_fwfm.err
python - <<'PY'
import json
for n in ("c2", "c2_uniform", "reference", "c3", "c4", "c4_fwfm", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(n, "ms/step %.4f value %.3e e2e %s roofline %s %s" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("ms_per_step"),
              r.get("frac"), (r.get("kernel") or "")[:40]), "| gather", (d.get("roofline_gather") or {}).get("frac"),
              "| fp32", (d.get("value_fp32") or {}).get("ms_per_step"), "| keras_dense", (d.get("value_keras_dense") or {}).get("ms_per_step"))
    except Exception as e:
        print(n, "no line:", e)
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 launches exit $?"
timeout 300 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_c3_eager.json 2> gpurun_out/bench_c3_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launches exit $?"
timeout 300 python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_c4_eager.json 2> gpurun_out/bench_c4_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fm_tile_kernel|gather_fm_fwd_lean" -s 6 -c 2 \
    -o gpurun_out/prof_c2_top python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full1.log 2>&1
echo "ncu full c2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_pair_fwd|field_pair_bwd" -s 4 -c 2 \
    -o gpurun_out/prof_c4_pair python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full2.log 2>&1
echo "ncu full c4 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_persist -s 0 -c 1 \
    -o gpurun_out/prof_c3_persist2 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full3.log 2>&1
echo "ncu full c3 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mlp_skinny_bwd|deepfm_tail" -s 6 -c 4 \
    -o gpurun_out/prof_c2_tower python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full4.log 2>&1
echo "ncu full tower exit $?"
