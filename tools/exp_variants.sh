python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
for sfx in "$@"; do
  [ "$sfx" = "default" ] && sfx=""
  echo "=== variant '$sfx'"
  for c in tcq:4096:14336:6,7 tcq:4096:14336:8 tcq:4096:14336:6 tcq:4096:4096:6,7 tcq:1024:4096:6 tcq:28672:4096:6,7 lut:4096:14336:8,2 lut:4096:14336:6,2 lut:4096:14336:4,1; do
    QP_LIB_SUFFIX=$sfx python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1
  done
done
