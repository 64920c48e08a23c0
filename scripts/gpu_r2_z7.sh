#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config c4 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 exit $?"
tail -2 gpurun_out/bench_c4.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4.json').read().strip().splitlines()[-1]); print('c4 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); r=d['roofline']; print({k:r[k] for k in r if k in ('achieved','frac','kernel_ms')})"
ETR_C4_MODEL=fwfm timeout 300 python bench.py --config c4 --no-cpu-baseline > gpurun_out/bench_c4_fwfm.json 2> gpurun_out/bench_c4_fwfm.err; echo "bench c4 fwfm exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_fwfm.json').read().strip().splitlines()[-1]); print('c4 fwfm ms/step', d['ms_per_step']); r=d['roofline']; print({k:r[k] for k in r if k in ('achieved','frac','kernel_ms')})"
