#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"
tail -2 gpurun_out/pytest_gpu.log
for v in ""; do
timeout 600 python bench.py --no-cpu-baseline --no-extras $v > gpurun_out/bench_c2_pa.json 2> gpurun_out/bench_c2_pa.err; echo "bench c2 [$v] exit $?"
tail -3 gpurun_out/bench_c2_pa.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_pa.json').read().strip().splitlines()[-1]); print('c2 [$v] ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'loss', d['e2e']['last_loss'], d['step_ms_min_median_max'])"
done
