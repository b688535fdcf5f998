#!/bin/bash
# usage: tools/exp_ab.sh out_prefix suffix...  -- isolated GEMVs and the 8B decode step for each variant library, interleaved twice
out=$1; shift
for rep in 1 2; do
for sfx in "$@"; do
  echo "=== variant '$sfx' rep $rep" >> gpurun_out/${out}.log
  for c in tcq:4096:14336:6,7 tcq:4096:4096:6,7 tcq:28672:4096:6,7 tcq:6144:4096:6,7; do
    QP_LIB_SUFFIX=$sfx timeout 100 python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1 >> gpurun_out/${out}.log
  done
  QP_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'], 'down us', d['roofline']['us_per_launch'])" >> gpurun_out/${out}.log 2>&1
done
done
