#!/bin/bash
# round-2 final ncu evidence (run under gpurun, after the same commands have exited 0 without ncu)
set -x
python tools/bench_gemv.py --one tcq:4096:14336:6,7 --iters 60 > gpurun_out/r2f_gemv_plain.log 2>&1 || exit 1
# (1) full-set capture of the dominant kernel, warm (skip the first launches), eager launches
ncu --set full --clock-control none --import-source on -k regex:tcq_gemv_kernel -s 6 -c 2 -o gpurun_out/r2f_tcq_gemv_full -f \
    python tools/bench_gemv.py --one tcq:4096:14336:6,7 --iters 12 --nograph > gpurun_out/r2f_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_tcq_gemv_full.ncu-rep 40 > gpurun_out/r2f_tcq_gemv_full_summary.txt 2>&1
rm -f gpurun_out/r2f_tcq_gemv_full.ncu-rep
# (2) the same kernel as nodes of a replayed CUDA graph (what bench.py times)
ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:tcq_gemv_kernel -s 40 -c 24 --csv --log-file gpurun_out/r2f_graphnode_gemv.csv \
    python tools/bench_gemv.py --one tcq:4096:14336:6,7 --iters 60 > gpurun_out/r2f_ncu_graph.log 2>&1
# (3) launch list of the bench command
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tp-extra > gpurun_out/r2f_bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tcq_|lut_|silu_mul|rope_attention|gemv_f16|embed_kernel|argmax_kernel|fused_norm_had|step_advance" -c 420 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tp-extra > gpurun_out/r2f_ncu_launches.log 2>&1
ls -la gpurun_out | tail -8
