#!/bin/bash
mkdir -p gpurun_out
for kern in flat rows; do
ETR_FUSED_APPLY=$kern timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_$kern.json 2> gpurun_out/bench_$kern.err; echo "bench $kern exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_$kern.json')); print('$kern', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'apply', (d.get('roofline_apply') or d['roofline'])['kernel_ms'])"
done
ETR_FUSED_APPLY=flat timeout 120 python scripts/mb_apply_r2.py record zipf uniform
