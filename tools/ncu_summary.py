"""print the key metrics of an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}]: {[r[i][:90] for r in data]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
        print(f"  {h[len('smsp__pcsamp_warps_issue_stalled_'):]}: {[r[i] for r in data]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    h = rows[1]
    ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > iex:
            body.append(r)
    tot = sum(int(r[isamp]) for r in body)
    totex = sum(int(r[iex]) for r in body)
    print("static instrs", len(body), "samples", tot, "executed", totex)
    for r in sorted(body, key=lambda r: -int(r[isamp]))[:topn]:
        print(r[isamp].rjust(5), r[iex].rjust(9), r[ia].strip()[:110])
    c = Counter()
    for r in body:
        t = r[ia].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0]] += int(r[iex])
    print([(k, v, round(v / totex * 100, 1)) for k, v in c.most_common(18)])
