#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_apply_tile.py > gpurun_out/check_tile.txt 2>&1; echo "check exit $?" >> gpurun_out/check_tile.txt
tail -4 gpurun_out/check_tile.txt
ETR_TILE_T=16 ETR_TILE_ITEM=64 timeout 300 python scripts/check_apply_tile.py zipf > gpurun_out/check_tile2.txt 2>&1; echo "check exit $?" >> gpurun_out/check_tile2.txt
tail -3 gpurun_out/check_tile2.txt
: > gpurun_out/mb_apply_r2t.txt
run() { env "$@" ETR_MB_ITERS=10 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" | cut -c1-200 >> gpurun_out/mb_apply_r2t.txt; }
run ETR_TILE_OCC=7
run ETR_TILE_OCC=6
run ETR_TILE_OCC=5
run ETR_TILE_OCC=7 ETR_TILE_T=16
run ETR_TILE_OCC=7 ETR_TILE_T=64
run ETR_TILE_OCC=7 ETR_MB_CLEAN=1
cat gpurun_out/mb_apply_r2t.txt
export ETR_MB_ITERS=4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_tile_kernel -s 2 -c 1 -o gpurun_out/r02_prof_apply_tile5 python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_tile5.log 2>&1
echo "ncu tile5 exit $?"
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_fm.py tests/test_gpu_tcgen05.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 2>&1 | tail -3
