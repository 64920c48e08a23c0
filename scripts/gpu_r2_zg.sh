for r in 2 4; do ETR_TILE_RPG=$r ETR_TILE_OCC=$((9 - r * 3 / 4 - 1)) timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/bench_rpg$r.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('RPG=$r', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
