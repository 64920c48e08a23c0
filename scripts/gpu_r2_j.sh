#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fm.py tests/test_gpu_fullsize.py tests/test_gpu_parity_r2.py tests/test_golden.py tests/test_gpu_sharded_single.py tests/test_gpu_f1_f3.py -m gpu -q --timeout 300 > gpurun_out/pytest_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_i.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_i.log | tail -12
for occ in 3 4; do
  ETR_K1_OCC=$occ timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 > gpurun_out/bench_k1_occ$occ.json 2> gpurun_out/bench_k1_occ$occ.err; echo "bench occ$occ exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_k1_occ$occ.json')); g=d.get('roofline_gather') or d['roofline']; print('occ$occ', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'K1 ms', g['kernel_ms'], 'frac', g['frac'])"
done
