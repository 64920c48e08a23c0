ETR_TILE_RPG=4 ETR_TILE_OCC=6 timeout 600 python -m pytest tests/test_gpu_fm.py tests/test_gpu_parity_r2.py tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
for cfg in "2 7" "4 5" "4 6" "4 7"; do set -- $cfg; ETR_TILE_RPG=$1 ETR_TILE_OCC=$2 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/RPG=$1 OCC=$2 /"; done > gpurun_out/mb_tile_rpg.txt 2>&1
cat gpurun_out/mb_tile_rpg.txt
