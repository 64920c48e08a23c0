#!/bin/bash
# N-GPU gpurun call: sharded-table check under torchrun + bench at N ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mgpu_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/mgpu_sharded_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "mgpu check exit $?" >> gpurun_out/mgpu_check_$N.log
tail -5 gpurun_out/mgpu_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
