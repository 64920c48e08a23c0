timeout 300 python -m pytest tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
bash scripts/gpu_r2_w.sh 2
