#!/bin/bash
# Round 2, call B: ncu of the apply kernels (old 4-lane kernel on the record layout vs the copy-engine pipeline)
mkdir -p gpurun_out
export ETR_MB_ITERS=4
ETR_FUSED_REC=off timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_fused_short -s 2 -c 1 \
    -o gpurun_out/r02_prof_apply_old_record python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_b1.log 2>&1
echo "ncu old exit $?"
ETR_FUSED_REC=D3S0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_fused_record -s 2 -c 1 \
    -o gpurun_out/r02_prof_apply_rec_d3s0 python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_b2.log 2>&1
echo "ncu rec exit $?"
tail -3 gpurun_out/ncu_b1.log gpurun_out/ncu_b2.log
