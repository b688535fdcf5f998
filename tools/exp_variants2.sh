#!/bin/bash
# usage: tools/exp_variants2.sh out_prefix suffix...   (FAST_BUILD variant libraries: tcomb_6_7 only)
out=$1; shift
for sfx in "$@"; do
  echo "=== variant '$sfx'" >> gpurun_out/${out}.log
  QP_LIB_SUFFIX=$sfx timeout 300 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "tcomb_6_7" 2>&1 | tail -1 >> gpurun_out/${out}.log
  for rep in ${REPS:-1}; do
  for c in tcq:4096:14336:6,7 tcq:4096:4096:6,7 tcq:28672:4096:6,7 tcq:6144:4096:6,7; do
    QP_LIB_SUFFIX=$sfx timeout 100 python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1 >> gpurun_out/${out}.log
  done
  done
done
