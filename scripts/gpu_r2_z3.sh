#!/bin/bash
mkdir -p gpurun_out
ETR_GEMM_PERSIST=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_persist -s 0 -c 1 \
    -o gpurun_out/prof_c3_persist2 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_full_p2.log 2>&1
echo "ncu persist2 exit $?"
for m in 0 2; do
  ETR_GEMM_PERSIST=$m timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p$m.json 2> gpurun_out/bench_c3_p$m.err; echo "bench c3 mode $m exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_p$m.json').read().strip().splitlines()[-1]); print('c3 mode $m ms/step', d['ms_per_step'], 'value', d['value']); r=d['roofline']; print({k:r[k] for k in r if k in ('kernel','achieved','frac','kernel_ms')})"
done
