#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (graph + eager), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --ignore tests/test_gpu_tcgen05.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 240 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q --timeout 60 > gpurun_out/pytest_tcgen05.log 2>&1; echo "pytest tcgen05 exit $?" >> gpurun_out/pytest_tcgen05.log
tail -15 gpurun_out/pytest_tcgen05.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --dist uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err
tail -c 1500 gpurun_out/bench_uniform.json
timeout 400 python bench.py --config c3 --steps 6 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"
tail -c 1600 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
# full ncu capture of the two headline kernels (after the same command exited 0 above)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_fm_fwd_tile -s 4 -c 2 \
    -o gpurun_out/prof_gather_fwd python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "ncu full gather exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fm_fused_short -s 3 -c 1 \
    -o gpurun_out/prof_fused_short python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "ncu full fused exit $?"
timeout 300 python bench.py --config c3 --steps 1 --warmup 3 --no-graph > gpurun_out/bench_c3_eager.json 2> gpurun_out/bench_c3_eager.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 0 -c 2 \
    -o gpurun_out/prof_cross python bench.py --config c3 --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_full3.log 2>&1
echo "ncu full cross exit $?"
