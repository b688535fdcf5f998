#!/bin/bash
# end-of-round evidence on the shipped library (run under gpurun): GPU tests, smoke, bench lines, reference arm,
# launch list of the bench command, the dominant kernel as nodes of a replayed graph
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gpu_tests.log 2>&1; tail -2 gpurun_out/r2g_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
python bench.py > gpurun_out/r2g_bench_8b.json 2>/dev/null
python bench.py --workload figure1d --no-cpu-baseline --no-tp-extra > gpurun_out/r2g_bench_figure1d.json 2>/dev/null
python bench.py --workload figure1c --no-cpu-baseline --no-tp-extra > gpurun_out/r2g_bench_figure1c.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2>/dev/null
for f in 8b figure1d figure1c; do tail -1 gpurun_out/r2g_bench_$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['workload'][:44], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d.get('extra',{}).get('long_context',{}).get('tok_s'), d.get('extra',{}).get('tp70b',{}).get('tok_s'))"; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tp-extra > gpurun_out/r2g_bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tcq_|lut_|silu_mul|rope_attention|gemv_f16|embed_kernel|argmax_kernel|fused_norm_had|step_advance" -c 420 --csv --log-file gpurun_out/r2g_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tp-extra > gpurun_out/r2g_ncu_launches.log 2>&1
ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:tcq_gemv_kernel -s 40 -c 24 --csv --log-file gpurun_out/r2g_graphnode_gemv.csv \
    python tools/bench_gemv.py --one tcq:4096:14336:6,7 --iters 60 > gpurun_out/r2g_ncu_graph.log 2>&1
ls -la gpurun_out | tail -6
