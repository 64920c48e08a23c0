#!/bin/bash
# Validation round at HEAD: full GPU parity suite, smoke, bench lines for c2/c3/c4/c1, launch lists, ncu full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  |pytest exit" gpurun_out/pytest_gpu.log | tail -15
echo "pytest took $(( $(date +%s) - S )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
S=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $? took $(( $(date +%s) - S )) s"
tail -c 4000 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 300 python bench.py --dist uniform --no-cpu-baseline --no-extras > gpurun_out/bench_c2_uniform.json 2> gpurun_out/bench_c2_uniform.err
tail -c 1500 gpurun_out/bench_c2_uniform.json
for c in c3 c4 c1; do
  S=$(date +%s)
  timeout 500 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $? took $(( $(date +%s) - S )) s"
  tail -c 2500 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err
done
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 12 -c 3 \
    -o gpurun_out/prof_c3_gemm python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full3.log 2>&1
echo "ncu full c3 exit $?"
ls -la gpurun_out | tail -40
