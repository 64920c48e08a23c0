#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err; echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_r2.json").read().strip().splitlines()[-1])
    print("N=$N value %.1f M samples/s  %.3f ms/step  e2e %.1f M (%.3f ms)  check %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], str(d.get("sharded_check"))[:200]))
except Exception as e:
    print("N=$N: no bench line:", e)
PY
tail -3 gpurun_out/bench_n${N}_r2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      scripts/mgpu_phases.py 2>&1 | grep " us " > gpurun_out/mgpu_phases_$N.txt; cat gpurun_out/mgpu_phases_$N.txt
