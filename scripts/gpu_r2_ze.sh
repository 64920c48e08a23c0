timeout 600 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity_r2.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_c2_quick.json 2> gpurun_out/bench_c2_quick.err; tail -c 1500 gpurun_out/bench_c2_quick.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_c2.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_c2.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
last = rows[-24:]
for r in last: print(f"{float(r[vi])/1e3:8.1f} us  {r[ki][:70]}")
PY
