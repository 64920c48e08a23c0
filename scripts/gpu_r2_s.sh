#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s.log
grep -E "passed|failed|FAILED|^E  |pytest exit" gpurun_out/pytest_s.log | tail -12
: > gpurun_out/mb_apply_r2s.txt
run() { env "$@" ETR_MB_ITERS=10 timeout 120 python scripts/mb_apply_r2.py record zipf uniform 2>&1 | grep "fused apply" | sed "s/^/$* /" | cut -c1-200 >> gpurun_out/mb_apply_r2s.txt; }
run ETR_TILE_OCC=7
run ETR_TILE_OCC=6
run ETR_TILE_OCC=7 ETR_MB_CLEAN=1
cat gpurun_out/mb_apply_r2s.txt
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tile.json')); print('tile', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'apply', (d.get('roofline_apply') or d['roofline'])['kernel_ms'], 'launches', d['gpu_launches_per_step'])"
ETR_FUSED_APPLY=flat timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_flat2.json 2> gpurun_out/bench_flat2.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_flat2.json')); print('flat', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'apply', (d.get('roofline_apply') or d['roofline'])['kernel_ms'], 'launches', d['gpu_launches_per_step'])"
