timeout 300 python -m pytest tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
for r in 0 1; do for u in 1 2; do for b in 0 16; do ETR_OWNER_REC=$r ETR_OWNER_U=$u ETR_OWNER_BPS=$b timeout 120 python scripts/mb_owner.py 2 2>&1 | grep "layout=record" | sed "s/^/rec=$r /"; done; done; done > gpurun_out/mb_owner.txt 2>&1
for r in 0 1; do ETR_OWNER_REC=$r timeout 120 python scripts/mb_owner.py 8 2>&1 | grep "layout=record" | sed "s/^/rec=$r /" >> gpurun_out/mb_owner.txt; done
cat gpurun_out/mb_owner.txt
