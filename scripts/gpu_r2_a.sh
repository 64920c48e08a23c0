#!/bin/bash
# Round 2, call A: new parity tests + apply-variant sweep + short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_r2.log 2>&1; echo "pytest r2 exit $?" >> gpurun_out/pytest_r2.log
tail -25 gpurun_out/pytest_r2.log
: > gpurun_out/mb_apply_r2.txt
timeout 120 python scripts/mb_apply_r2.py plain zipf uniform >> gpurun_out/mb_apply_r2.txt 2>&1
for v in D2S0 D3S0 D4S0 D2S1 D3S1 D4S1; do
  ETR_FUSED_REC=$v timeout 120 python scripts/mb_apply_r2.py record zipf uniform >> gpurun_out/mb_apply_r2.txt 2>&1
done
ETR_FUSED_REC=off timeout 120 python scripts/mb_apply_r2.py record zipf >> gpurun_out/mb_apply_r2.txt 2>&1
cat gpurun_out/mb_apply_r2.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --ignore tests/test_gpu_parity_r2.py -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
