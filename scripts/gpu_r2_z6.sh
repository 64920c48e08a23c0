#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tcgen05 exit $?"
grep -E "passed|failed|FAILED|^E  |Error" gpurun_out/pytest_tc.log | head -12
ETR_GEMM_EPI_PIPE=1 timeout 300 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 120 -k persistent > gpurun_out/pytest_tc_pipe.log 2>&1; echo "pytest tcgen05 (pipe) exit $?"
for cfg in "0 0" "1 0" "2 0" "1 1" "2 1"; do
  set -- $cfg
  ETR_GEMM_PERSIST=$1 ETR_GEMM_EPI_PIPE=$2 timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p$1$2.json 2> gpurun_out/bench_c3_p$1$2.err; echo "bench c3 mode $1 pipe $2 exit $?"
  tail -2 gpurun_out/bench_c3_p$1$2.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_p$1$2.json').read().strip().splitlines()[-1]); print('c3 mode $1 pipe $2 ms/step', d['ms_per_step']); r=d['roofline']; print({k:r[k] for k in r if k in ('achieved','frac','kernel_ms')})"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_persist -s 0 -c 1 \
    -o gpurun_out/prof_c3_persist2 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full_p2.log 2>&1
echo "ncu persist2 exit $?"
