timeout 300 python -m pytest tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
for f in 0 1; do for w in 2 8; do ETR_OWNER_LPR5=$f timeout 120 python scripts/mb_owner.py $w 2>&1 | grep "layout=record" | sed "s/^/LPR5=$f /"; done; done > gpurun_out/mb_owner5.txt 2>&1
cat gpurun_out/mb_owner5.txt
