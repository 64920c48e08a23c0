#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fm.py tests/test_gpu_fullsize.py tests/test_gpu_parity_r2.py tests/test_gpu_golden.py tests/test_golden.py tests/test_gpu_sharded_single.py tests/test_gpu_f1_f3.py -m gpu -q --timeout 300 > gpurun_out/pytest_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_i.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_i.log | tail -12
for which in lean generic; do
  if [ $which = generic ]; then export ETR_GATHER=generic; else unset ETR_GATHER; fi
  timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_k1_$which.json 2> gpurun_out/bench_k1_$which.err; echo "bench $which exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_k1_$which.json')); g=d.get('roofline_gather') or d['roofline']; print('$which', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'K1 ms', g['kernel_ms'], 'frac', g['frac'])"
done
unset ETR_GATHER
timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 20 --dist uniform > gpurun_out/bench_k1_uniform.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_k1_uniform.json')); g=d.get('roofline_gather') or d['roofline']; print('uniform', 'ms/step', d['ms_per_step'], 'K1 ms', g['kernel_ms'], 'frac', g['frac'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_fm_fwd_lean -s 30 -c 1 -o gpurun_out/r02_prof_gather_lean python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_i.log 2>&1
echo "ncu exit $?"
