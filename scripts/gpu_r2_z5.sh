#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tcgen05 exit $?"
grep -E "passed|failed|FAILED|^E  |Error" gpurun_out/pytest_tc.log | head -12
for m in 0 1 2; do
  ETR_GEMM_PERSIST=$m timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p$m.json 2> gpurun_out/bench_c3_p$m.err; echo "bench c3 mode $m exit $?"
  tail -2 gpurun_out/bench_c3_p$m.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_p$m.json').read().strip().splitlines()[-1]); print('c3 mode $m ms/step', d['ms_per_step'], 'value', d['value']); r=d['roofline']; print({k:r[k] for k in r if k in ('kernel','achieved','frac','kernel_ms')})"
done
timeout 300 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_c3_eager.json 2> gpurun_out/bench_c3_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_persist -s 0 -c 1 \
    -o gpurun_out/prof_c3_persist2 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full_p2.log 2>&1
echo "ncu persist2 exit $?"
timeout 300 python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_c4_eager.json 2> gpurun_out/bench_c4_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_pair_fwd|field_pair_bwd" -s 4 -c 2 \
    -o gpurun_out/prof_c4_pair python bench.py --config c4 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full2.log 2>&1
echo "ncu full c4 exit $?"
