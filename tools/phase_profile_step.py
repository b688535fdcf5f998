"""debug: timeline of every TCQ GEMV launch inside ONE captured decode step (QP_PROFILE_PHASES build: QP_LIB_SUFFIX=_prof).
Each GEMV CTA appends {shape, 7 %globaltimer stamps} to a device log (csrc/tcq_kernels.cu, qp_debug_plog); the records are
grouped into launches by time and printed relative to the previous launch's last CTA, so that one sees, per projection, when its
CTAs became resident, how long they waited for the dependency, how long the glue kernel in between ran, and the loop time.

    QP_LIB_SUFFIX=_prof [QP_AHEAD_MODE=old|late|ahead|mixed] python tools/phase_profile_step.py [layers]
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette._cabi import lib  # noqa: E402
from qpalette.decode import LLAMA31_8B, DecodeRunner, uniform_qdict  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
shape = LLAMA31_8B
qd, mi = uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * shape.num_hidden_layers
r = DecodeRunner(shape, qd, mi, max_seq=64, seed=0, num_layers=layers)
r.capture()
r.reset(1)
for _ in range(6):
    r.step()
torch.cuda.synchronize()
L = lib()
L.qp_debug_plog.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
cap = 40000
buf = np.zeros((cap, 9), dtype=np.uint64)
n = ctypes.c_uint(0)
L.qp_debug_plog(buf.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n))  # clear
r.step()
torch.cuda.synchronize()
L.qp_debug_plog(buf.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n))
rec = buf[: n.value].astype(np.int64)
M, K = rec[:, 0] >> 32, rec[:, 0] & 0xFFFFFFFF
ahead = (rec[:, 1] >> 20) & 3
t = rec[:, 2:9]
# launches: sort by the dependency-wait-done stamp (one release per launch), split where shape changes or the gap is large
order = np.argsort(t[:, 3], kind="stable")
rec, M, K, ahead, t = rec[order], M[order], K[order], ahead[order], t[order]
groups, start = [], 0
for i in range(1, len(rec) + 1):
    if i == len(rec) or M[i] != M[start] or K[i] != K[start] or t[i, 3] - t[i - 1, 3] > 2500:
        groups.append((start, i))
        start = i
print(f"mode {os.environ.get('QP_AHEAD_MODE', 'mixed')}: {len(rec)} CTA records, {len(groups)} launches")
names = ["start", "prefetch issued", "pre-wait work done", "wait done", "x staged", "warp0 done", "cta done"]
prev_end = None
t0 = t[:, 0].min()
rows = []
for gi, (a, b) in enumerate(groups):
    g = t[a:b]
    ref = prev_end if prev_end is not None else g[:, 0].min()
    med = lambda c: int(np.median(g[:, c] - ref))
    rows.append((int(M[a]), int(K[a]), int(ahead[a]), b - a, int(g[:, 0].min() - ref), med(0), int(g[:, 0].max() - ref), med(2), med(3),
                 int(g[:, 3].max() - ref), med(4), int(g[:, 4].max() - ref), med(6), int(g[:, 6].max() - ref),
                 int(np.median(g[:, 6] - g[:, 4]))))
    if gi >= len(groups) - 5:  # which CTAs start late, and do the late starters finish last?
        cta = rec[a:b, 1] & 0xFFFF
        o_ = np.argsort(g[:, 0])
        late = o_[-40:]
        print(f"launch {gi} {int(M[a])}x{int(K[a])}: latest 40 starters (cta:start_ns:done_ns rel. to first start): "
              + " ".join(f"{int(cta[j])}:{int(g[j, 0] - g[:, 0].min())}:{int(g[j, 6] - g[:, 0].min())}" for j in late))
        fin = np.argsort(g[:, 6])[-24:]
        z0 = g[:, 0].min()
        print("   latest 24 finishers (cta:start:wait_done:x_staged:done): "
              + " ".join(f"{int(cta[j])}:{int(g[j, 0] - z0)}:{int(g[j, 3] - z0)}:{int(g[j, 4] - z0)}:{int(g[j, 6] - z0)}" for j in fin))
        print(f"   median (start, wait_done, x_staged, done) = {int(np.median(g[:, 0]) - z0)} {int(np.median(g[:, 3]) - z0)} "
              f"{int(np.median(g[:, 4]) - z0)} {int(np.median(g[:, 6]) - z0)}")
        print(f"   median start of cta < 100: {int(np.median(g[cta < 100, 0] - g[:, 0].min()))}; corr(cta index, start) = "
              f"{np.corrcoef(cta, g[:, 0] - g[:, 0].min())[0, 1]:.2f}; corr(start, done) = {np.corrcoef(g[:, 0], g[:, 6])[0, 1]:.2f}")
    prev_end = g[:, 6].max()
print("times in ns relative to the PREVIOUS GEMV launch's last CTA end")
print(f"{'M':>6} {'K':>6} ah ctas | start min/med/max      | prework med | wait done med/max | x staged med/max | cta done med/max | loop med")
for rw in rows[-4 * min(layers, 3) - 1:]:
    print(f"{rw[0]:6d} {rw[1]:6d} {rw[2]:2d} {rw[3]:4d} | {rw[4]:6d} {rw[5]:6d} {rw[6]:6d} | {rw[7]:8d} | {rw[8]:7d} {rw[9]:7d} | {rw[10]:7d} {rw[11]:7d} | "
          f"{rw[12]:7d} {rw[13]:7d} | {rw[14]:6d}")
# sub-phases of the fused x-producer prologue: g_xphase holds the LAST fused launch of the step (the last layer's up/gate GEMV)
try:
    xb = np.zeros((256, 8), dtype=np.uint64)
    L.qp_debug_xphases.argtypes = [ctypes.c_void_p]
    L.qp_debug_xphases(xb.ctypes.data_as(ctypes.c_void_p))
    xb = xb[:148, :6].astype(np.int64)
    xmode = (rec[:, 1] >> 16) & 3
    fused = [(a, b) for (a, b) in groups if xmode[a] == 1]
    a, b = fused[-1]
    cta = rec[a:b, 1] & 0xFFFF
    wait_done = np.zeros(148, dtype=np.int64)
    wait_done[cta] = t[a:b, 3]
    staged = np.zeros(148, dtype=np.int64)
    staged[cta] = t[a:b, 4]
    relx = xb - wait_done[:, None]
    print(f"x-producer sub-phases of the last fused launch ({int(M[a])}x{int(K[a])}), ns after the CTA's dependency wait, median / max over CTAs:")
    for i, nm in enumerate(["enter", "loads issued", "inputs arrived, residual added", "normalised", "in-warp stages done, in smem", "hadamard done"]):
        print(f"    {nm:32s} {int(np.median(relx[:, i])):6d} / {int(relx[:, i].max()):6d}")
    print(f"    {'x staged (fragment order, barrier)':32s} {int(np.median(staged - wait_done)):6d} / {int((staged - wait_done).max()):6d}")
except Exception as e:  # noqa: BLE001
    print("x-phase stamps unavailable:", e)
per_layer = (t[:, 6].max() - t0) / max(1, layers)
print(f"span of all GEMV launches of the step: {(t[:, 6].max() - t0) / 1e3:.1f} us = {per_layer / 1e3:.1f} us / layer")
