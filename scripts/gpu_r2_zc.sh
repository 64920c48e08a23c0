#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/mgpu_sharded_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "mgpu check exit $?" >> gpurun_out/mgpu_check_$N.log
grep "MGPU_OK\|Error\|exit\|err=" gpurun_out/mgpu_check_$N.log | tail -30 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --config c3 --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "bench c3 N=$N exit $?"
tail -2 gpurun_out/bench_c3_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_n$N.json').read().strip().splitlines()[-1]); print('c3 N=$N ms/step', d['ms_per_step'], 'value', d['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err; echo "bench N=$N exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n${N}_r2.json').read().strip().splitlines()[-1]); print('c2 N=$N ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['sharded_check']['ok'])"
