"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: this library's kernels only, by kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_bench_summary.txt"""
import collections
import csv
import sys

OURS = ("tcq_", "lut_", "simt_", "silu_mul", "rope_attention", "gemv_f16", "embed_kernel", "argmax_kernel", "fused_norm_had",
        "step_advance", "hadamard", "xchg", "x_to_frag", "scale_epilogue", "gemm_tc")
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
n = 0
for r in rows:
    name, val = r[4], float(r[14].replace(",", ""))
    if not any(k in name for k in OURS):
        continue
    tot[name[:70]] += val / 1e3
    cnt[name[:70]] += 1
    n += 1
total = sum(tot.values())
print("ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tp-extra` (first 420 launches of this library's\n"
      "kernels = the eager decode steps the graph capture warms up with; gpu__time_duration.sum per launch: cold caches,\n"
      "serialised, no PDL overlap, --clock-control none).  Shares, not absolute times, compare with the timed (graph-replayed) run.\n")
for k, v in tot.most_common():
    print(f"{k:70s} n={cnt[k]:4d} total {v:8.1f} us avg {v / cnt[k]:7.2f} us {v / total * 100:5.1f}%")
print(f"total {total:.1f} us over {n} launches")
