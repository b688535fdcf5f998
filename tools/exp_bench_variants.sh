#!/bin/bash
# usage: tools/exp_bench_variants.sh out_prefix suffix...   8B decode step with variant libraries (FAST_BUILD ones are enough)
out=$1; shift
for rep in 1 2; do
for sfx in "$@"; do
  [ "$sfx" = "default" ] && sfx=""
  echo "=== variant '$sfx'" >> gpurun_out/${out}.log
  QP_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'], 'down us', d['roofline']['us_per_launch'], {k: v['us_per_launch'] for k, v in d['roofline_detail'].items() if isinstance(v, dict)})" >> gpurun_out/${out}.log 2>&1
done
done
