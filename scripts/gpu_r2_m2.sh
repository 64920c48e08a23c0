#!/bin/bash
# 2-GPU round: sharded correctness (all forms, all families) + the sharded bench with its in-line sharded_check
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_sharded_check.py > gpurun_out/mgpu_check_2.log 2>&1; echo "mgpu check exit $?"
tail -4 gpurun_out/mgpu_check_2.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
