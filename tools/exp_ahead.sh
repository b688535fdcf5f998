#!/bin/bash
# prologue-order experiment (run under gpurun): parity first, then isolated GEMVs under each scheduling hint, then the 8B decode
# step with the hints forced / mixed.   usage: tools/exp_ahead.sh out_prefix
out=$1
python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${out}_tests.log
for hint in 0 2 4; do
  echo "=== QP_GEMV_HINT=$hint (0 default / 2 decode ahead / 4 table late)" >> gpurun_out/${out}_gemv.log
  for c in tcq:4096:14336:6,7 tcq:4096:4096:6,7 tcq:28672:4096:6,7 tcq:6144:4096:6,7; do
    QP_GEMV_HINT=$hint timeout 100 python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1 >> gpurun_out/${out}_gemv.log
  done
done
for mode in old late ahead mixed old; do
  echo "=== QP_AHEAD_MODE=$mode" >> gpurun_out/${out}_bench.log
  QP_AHEAD_MODE=$mode timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'], 'roofline', d['roofline']['us_per_launch'], d['roofline_detail'])" >> gpurun_out/${out}_bench.log 2>&1
done
