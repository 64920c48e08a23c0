#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/mb_apply_r2x.txt
run() { env "$@" ETR_MB_ITERS=12 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" | cut -c1-160 >> gpurun_out/mb_apply_r2x.txt; }
for t in 24 32 48 64; do run ETR_TILE_T=$t; done
for it in 128 512 1024; do run ETR_TILE_ITEM=$it; done
run ETR_TILE_T=48 ETR_TILE_ITEM=512
run ETR_TILE_OCC=6
run ETR_TILE_OCC=6 ETR_TILE_T=48
cat gpurun_out/mb_apply_r2x.txt
