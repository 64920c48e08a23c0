#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"
grep -E "passed|failed|FAILED|^E  |Error" gpurun_out/pytest_gpu.log | head -12
for o in 1; do
ETR_CROSS_BWD_ONEPASS=$o timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_o$o.json 2> gpurun_out/bench_c3_o$o.err; echo "bench c3 onepass $o exit $?"
tail -2 gpurun_out/bench_c3_o$o.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_o$o.json').read().strip().splitlines()[-1]); print('c3 onepass $o ms/step', d['ms_per_step']); r=d['roofline']; print({k:r[k] for k in r if k in ('achieved','frac','kernel_ms')})"
done
timeout 300 python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/bench_c3_eager.json 2> gpurun_out/bench_c3_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launches exit $?"
