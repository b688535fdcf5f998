for sfx in "$@"; do
  [ "$sfx" = "default" ] && sfx=""
  echo "=== variant '$sfx'"
  QP_LIB_SUFFIX=$sfx timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -1
  for c in tcq:4096:14336:6,7 tcq:4096:14336:8 tcq:4096:4096:6,7 tcq:28672:4096:6,7 lut:4096:14336:8,2 lut:4096:14336:6,2; do
    QP_LIB_SUFFIX=$sfx timeout 100 python tools/bench_gemv.py --one $c --iters 300 2>&1 | tail -1
  done
done
