#!/bin/bash
# persistent GEMM bring-up: mode 1 (one-CTA persistent) then mode 2 (CTA pairs), each under its own timeout
mkdir -p gpurun_out
for k in 1 2; do
  ETR_TEST_PERSIST_MODES=$k timeout 240 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 120 -k persistent > gpurun_out/pytest_persist_$k.log 2>&1; echo "pytest persistent mode $k exit $?"
  grep -E "passed|failed|FAILED|^E  |Error" gpurun_out/pytest_persist_$k.log | head -12
done
for m in 0 1 2; do
  ETR_GEMM_PERSIST=$m timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p$m.json 2> gpurun_out/bench_c3_p$m.err; echo "bench c3 mode $m exit $?"
  tail -2 gpurun_out/bench_c3_p$m.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_p$m.json').read().strip().splitlines()[-1]); print('c3 mode $m ms/step', d['ms_per_step'], 'value', d['value']); r=d['roofline']; print({k:r[k] for k in r if k in ('kernel','achieved','frac','kernel_ms')})"
done
timeout 300 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tcgen05 exit $?"
tail -3 gpurun_out/pytest_tc.log
