#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/mb_apply_r2o.txt
run() { env "$@" ETR_MB_ITERS=10 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" >> gpurun_out/mb_apply_r2o.txt; }
run ETR_FUSED_APPLY=tile ETR_TILE_T=32 ETR_MB_CLEAN=1
run ETR_FUSED_APPLY=tile ETR_TILE_T=32 ETR_TILE_GRID=2 ETR_TILE_ICTA=2
run ETR_FUSED_APPLY=tile ETR_TILE_T=32 ETR_TILE_GRID=2 ETR_TILE_ICTA=2 ETR_MB_CLEAN=1
run ETR_FUSED_APPLY=rows ETR_MB_CLEAN=1
cat gpurun_out/mb_apply_r2o.txt
export ETR_FUSED_APPLY=tile ETR_TILE_T=32 ETR_MB_ITERS=4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_apply_tile.csv python scripts/mb_apply_r2.py record zipf > /dev/null 2>&1
grep "fm_tile" gpurun_out/launches_apply_tile.csv | cut -d, -f5,13- | tail -9
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_tile_kernel -s 2 -c 1 -o gpurun_out/r02_prof_apply_tile python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_tile.log 2>&1
echo "ncu tile exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_tile_item_kernel -s 2 -c 1 -o gpurun_out/r02_prof_apply_item python scripts/mb_apply_r2.py record zipf > gpurun_out/ncu_item.log 2>&1
echo "ncu item exit $?"
