#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tests/mgpu_sharded_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "mgpu check exit $?" >> gpurun_out/mgpu_check_$N.log
grep "MGPU_OK\|Error\|error\|exit\|assert" gpurun_out/mgpu_check_$N.log | tail -8
for flag in "" "--no-plan-ahead"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 20 --warmup 5 $flag > gpurun_out/bench_n${N}_r2$flag.json 2> gpurun_out/bench_n${N}_r2$flag.err; echo "bench N=$N $flag exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_r2$flag.json").read().strip().splitlines()[-1])
    print("N=$N $flag value %.1f M samples/s  %.3f ms/step  e2e %.1f M (%.3f ms)  check %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], str(d.get("sharded_check"))[:200]))
except Exception as e:
    print("N=$N: no bench line:", e)
PY
tail -3 gpurun_out/bench_n${N}_r2$flag.err
done
