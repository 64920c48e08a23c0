#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_interactions.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_y.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_y.log
grep -E "passed|failed|FAILED|^E  |pytest exit" gpurun_out/pytest_y.log | tail -15
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_r2.json 2> gpurun_out/bench_c3_r2.err; echo "bench c3 exit $?"
tail -2 gpurun_out/bench_c3_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_r2.json').read().strip().splitlines()[-1]); print('c3 ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d.get('e2e',{}).get('ms_per_step')); r=d['roofline']; print({k:r[k] for k in r if k in ('kernel','achieved','frac','kernel_ms')})"
