#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/mb_apply_r2k.txt
for lean in 0 4 5 6; do
ETR_FUSED_LEAN=$lean ETR_FUSED_APPLY=rows timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | sed "s/^/lean=$lean /" >> gpurun_out/mb_apply_r2k.txt
done
cat gpurun_out/mb_apply_r2k.txt
for lean in 5 6; do
ETR_FUSED_LEAN=$lean ETR_FUSED_APPLY=rows timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_lean$lean.json 2> gpurun_out/bench_lean$lean.err
python -c "
import json; d=json.load(open('gpurun_out/bench_lean$lean.json')); print('rows lean$lean', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'apply', (d.get('roofline_apply') or d['roofline'])['kernel_ms'])"
done
ETR_FUSED_LEAN=5 ETR_FUSED_APPLY=rows timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_fm.py -m gpu -q --timeout 300 2>&1 | tail -3
