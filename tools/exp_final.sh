#!/bin/bash
# round-end evidence on one box: GPU tests, default bench, mma GEMM comparison, ncu captures
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2g_tests.log
python bench.py > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
python bench.py --workload figure1d --no-cpu-baseline --no-tp-extra > gpurun_out/r2g_bench_figure1d.json 2>> gpurun_out/r2g_bench_n1.err
python tools/repro_mma.py > gpurun_out/r2g_mma.log 2>&1
bash tools/exp_ncu_final.sh > gpurun_out/r2g_ncu.log 2>&1
