#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q --timeout 600 -k "c3_shape or c4_shape" --durations=6 > gpurun_out/pytest_shape.log 2>&1; echo "pytest shape exit $?"
grep -E "passed|failed|FAILED|^E  |Error|s call" gpurun_out/pytest_shape.log | head -30
cat gpurun_out/parity_r2.json | head -60
