#!/bin/bash
# GEMV v2 experiment driver (run under gpurun): parity of the default library, then timing of variant libraries
# usage: tools/exp_v2.sh out_prefix variant_suffix...   ("default" = libqpalette.so)
out=$1; shift
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_shapes.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${out}_tests.log
for sfx in "$@"; do
  [ "$sfx" = "default" ] && sfx=""
  echo "=== variant '$sfx'" >> gpurun_out/${out}_bench.log
  cases="tcq:4096:14336:6,7 tcq:4096:4096:6,7 tcq:28672:4096:6,7 tcq:6144:4096:6,7"
  [ -z "$sfx" ] && cases="$cases tcq:4096:14336:8 tcq:4096:14336:6 tcq:4096:14336:7 tcq:4096:14336:4 lut:4096:14336:8,2 lut:4096:14336:6,2 lut:4096:14336:4,2 lut:4096:14336:4,1"
  for c in $cases; do
    QP_LIB_SUFFIX=$sfx timeout 100 python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1 >> gpurun_out/${out}_bench.log
  done
done
