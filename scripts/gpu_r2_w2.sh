timeout 300 python -m pytest tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
ETR_SERVE_FLAT=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 scripts/mgpu_phases.py 2>&1 | grep " us " | grep "serve" | sed 's/^/FLAT=0 /'
bash scripts/gpu_r2_w.sh 2
