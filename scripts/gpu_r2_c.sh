#!/bin/bash
# Round 2, call C: occurrence-parallel fused apply -- parity, timing vs the row-parallel kernels, ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_fm.py -m gpu -q --timeout 300 > gpurun_out/pytest_r2.log 2>&1; echo "pytest r2 exit $?" >> gpurun_out/pytest_r2.log
tail -25 gpurun_out/pytest_r2.log
: > gpurun_out/mb_apply_r2c.txt
ETR_FUSED_APPLY=flat timeout 120 python scripts/mb_apply_r2.py record zipf uniform >> gpurun_out/mb_apply_r2c.txt 2>&1
ETR_FUSED_APPLY=rows ETR_FUSED_REC=off timeout 120 python scripts/mb_apply_r2.py record zipf uniform >> gpurun_out/mb_apply_r2c.txt 2>&1
cat gpurun_out/mb_apply_r2c.txt
ETR_MB_ITERS=4 ETR_FUSED_APPLY=flat timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_fused_flat_kernel -s 2 -c 1 \
    -o gpurun_out/r02_prof_apply_flat python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_c1.log 2>&1
echo "ncu flat exit $?"
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 4500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
