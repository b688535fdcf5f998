# interleaved A/B of prologue variants: QP_LIB_SUFFIX libraries, 8B decode step, 3 repetitions
out=${1:-r2_ab_prologue}; shift
for rep in 1 2 3; do for sfx in "$@"; do
QP_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('variant[$sfx] rep $rep', d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'])" >> gpurun_out/${out}.log 2>&1
done; done
