#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_apply_tile.py > gpurun_out/check_tile.txt 2>&1; echo "check exit $?" >> gpurun_out/check_tile.txt
tail -5 gpurun_out/check_tile.txt
: > gpurun_out/mb_apply_r2n.txt
run() { env "$@" ETR_MB_ITERS=10 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" >> gpurun_out/mb_apply_r2n.txt; }
run ETR_FUSED_APPLY=tile
run ETR_FUSED_APPLY=tile ETR_TILE_T=8
run ETR_FUSED_APPLY=tile ETR_TILE_T=32
run ETR_FUSED_APPLY=tile ETR_TILE_T=64
run ETR_FUSED_APPLY=tile ETR_TILE_RB=4 ETR_TILE_OCC=3
run ETR_FUSED_APPLY=tile ETR_TILE_GRID=4 ETR_TILE_ICTA=1
run ETR_FUSED_APPLY=tile ETR_TILE_GRID=2 ETR_TILE_ICTA=2
run ETR_FUSED_APPLY=tile ETR_TILE_GRID=3 ETR_TILE_ICTA=2
run ETR_FUSED_APPLY=tile ETR_TILE_ITEM=128
run ETR_FUSED_APPLY=tile ETR_TILE_ITEM=512
run ETR_FUSED_APPLY=tile ETR_TILE_OCC=5 ETR_TILE_GRID=4
run ETR_FUSED_APPLY=flat
cat gpurun_out/mb_apply_r2n.txt
ETR_FUSED_APPLY=tile timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_fm.py tests/test_gpu_tcgen05.py -m gpu -q --timeout 300 2>&1 | tail -5
