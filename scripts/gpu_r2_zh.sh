timeout 600 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity_r2.py tests/test_gpu_interactions.py tests/test_gpu_sharded_single.py -x -q 2>&1 | tail -3
for f in 0 1; do ETR_FLAT_SEGRED=$f timeout 300 python bench.py --config c3 --steps 10 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/bench_c3_f$f.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('FLAT_SEGRED=$f', d['ms_per_step'], d['value'], d['roofline']['frac'] if isinstance(d['roofline'], dict) else d['roofline'])"; done
cp gpurun_out/bench_c3_f1.err gpurun_out/bench_c3_f1.err.keep 2>/dev/null; tail -2 gpurun_out/bench_c3_f1.err
