#!/bin/bash
# N-GPU gpurun call: sharded-table checks under torchrun + bench at the given rank counts.
#   bash scripts/gpu_round_mgpu.sh "2 4"      (on a box with >= max N GPUs)
NS=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mgpu_gpus.txt 2>&1
for N in $NS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tests/mgpu_sharded_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "mgpu check exit $?" >> gpurun_out/mgpu_check_$N.log
  grep "MGPU_OK\|Error\|exit" gpurun_out/mgpu_check_$N.log | tail -4
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.1f M samples/s  %.3f ms/step  e2e %.1f M  K1 %.1f us" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["roofline"]["kernel_ms"] * 1e3))
except Exception as e:
    print("N=$N: no bench line:", e)
PY
  tail -3 gpurun_out/bench_n$N.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      scripts/mgpu_phases.py 2>&1 | grep " us " > gpurun_out/mgpu_phases_$N.txt; cat gpurun_out/mgpu_phases_$N.txt
done
