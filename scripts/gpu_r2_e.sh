#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r2.log 2>&1; echo "pytest r2 exit $?" >> gpurun_out/pytest_r2.log
tail -30 gpurun_out/pytest_r2.log
: > gpurun_out/mb_apply_r2e.txt
ETR_FUSED_APPLY=flat timeout 120 python scripts/mb_apply_r2.py record zipf uniform >> gpurun_out/mb_apply_r2e.txt 2>&1
ETR_FUSED_APPLY=rows timeout 120 python scripts/mb_apply_r2.py record zipf uniform >> gpurun_out/mb_apply_r2e.txt 2>&1
cat gpurun_out/mb_apply_r2e.txt
ETR_MB_ITERS=4 ETR_FUSED_APPLY=flat timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_apply_flat.csv python scripts/mb_apply_r2.py record zipf > /dev/null 2>&1
grep -E "fm_fused" gpurun_out/launches_apply_flat.csv | tail -4 | cut -d, -f5,15-
ETR_MB_ITERS=4 ETR_FUSED_APPLY=flat timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_fused_flat_kernel -s 2 -c 1 \
    -o gpurun_out/r02_prof_apply_flat2 python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_e1.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
