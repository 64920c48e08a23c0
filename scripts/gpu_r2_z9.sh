#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2b.json 2> gpurun_out/bench_c2b.err; echo "bench c2 exit $?"
tail -3 gpurun_out/bench_c2b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2b.json').read().strip().splitlines()[-1]); print('c2 ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['e2e']['repetitions'])"
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 launches exit $?"
