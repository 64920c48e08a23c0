#!/bin/bash
mkdir -p gpurun_out
ETR_TILE_T=32 timeout 300 python scripts/check_apply_tile.py > gpurun_out/check_tile.txt 2>&1; echo "check exit $?" >> gpurun_out/check_tile.txt
tail -4 gpurun_out/check_tile.txt
ETR_TILE_T=16 ETR_TILE_ITEM=64 timeout 300 python scripts/check_apply_tile.py zipf > gpurun_out/check_tile2.txt 2>&1; echo "check exit $?" >> gpurun_out/check_tile2.txt
tail -3 gpurun_out/check_tile2.txt
: > gpurun_out/mb_apply_r2p.txt
run() { env "$@" ETR_MB_ITERS=10 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" | cut -c1-200 >> gpurun_out/mb_apply_r2p.txt; }
for occ in 5 6 7 8; do run ETR_TILE_T=32 ETR_TILE_OCC2=$occ; done
run ETR_TILE_T=16 ETR_TILE_OCC2=7
run ETR_TILE_T=64 ETR_TILE_OCC2=7
run ETR_TILE_T=32 ETR_TILE_OCC2=7 ETR_TILE_ITEM=128
run ETR_TILE_T=32 ETR_TILE_OCC2=7 ETR_MB_CLEAN=1
run ETR_TILE_T=32 ETR_TILE_V2=0
cat gpurun_out/mb_apply_r2p.txt
export ETR_TILE_T=32 ETR_MB_ITERS=4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_apply_tile2.csv python scripts/mb_apply_r2.py record zipf > /dev/null 2>&1
grep "fm_tile" gpurun_out/launches_apply_tile2.csv | cut -d, -f5,13- | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_tile2_kernel -s 2 -c 1 -o gpurun_out/r02_prof_apply_tile2 python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_tile2.log 2>&1
echo "ncu tile2 exit $?"
