"""debug: per-phase timeline of the TCQ GEMV kernel (needs the QP_PROFILE_PHASES build: QP_LIB_SUFFIX=_prof).
    QP_LIB_SUFFIX=_prof python tools/phase_profile.py M K
Launches the tcomb_6_7 GEMV over 14 distinct weight buffers back to back (as the decode graph does), reads the
%globaltimer stamps of the LAST launch (one row per CTA) and prints the per-phase distribution plus the launch period
measured with CUDA events, so that in-kernel time and launch-to-launch overhead can be separated."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops, _cabi
from qpalette._cabi import SPLIT_IN
M, K = int(sys.argv[1]), int(sys.argv[2])
kv = (6, 7)
dev = "cuda"
tl = torch.randn((512, 2), device=dev).half(); x = torch.randn((1, K), device=dev).half()
bufs = [(torch.randint(0, 256, (M * (K // 2) * kv[0] // 16,), dtype=torch.uint8, device=dev),
         torch.randint(0, 256, (M * (K // 2) * kv[1] // 16,), dtype=torch.uint8, device=dev)) for _ in range(14)]
out = torch.zeros((1, M), dtype=torch.float32, device=dev)
def run():
    for b in bufs:
        ops.tcq_gemv(b[0], x, tl, M, K, 9, kv[0], b[1], kv[1], SPLIT_IN, K // 2, out=out, accumulate=True)
run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
g.replay(); torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    g.replay()
b_.record(); torch.cuda.synchronize()
print(f"{M}x{K} tcomb_6_7: launch period {a.elapsed_time(b_) * 1e3 / 20 / len(bufs):.2f} us (graph of {len(bufs)} launches)")
h = np.zeros((256, 8), dtype=np.uint64)
_cabi.lib().qp_debug_phases.argtypes = [ctypes.c_void_p]
_cabi.lib().qp_debug_phases(h.ctypes.data_as(ctypes.c_void_p))
h = h[:148].astype(np.int64)
t0 = h[:, 0].min()
rel = h - t0
names = ["start", "prefetch+tlut issued", "table built", "pdl wait done", "x staged", "warp0 done", "cta done"]
for i, n in enumerate(names):
    print(f"{n:22s} min {rel[:, i].min():7d} ns  median {int(np.median(rel[:, i])):7d} ns  max {rel[:, i].max():7d} ns")
dur = h[:, 6] - h[:, 0]
print(f"per-CTA span start->done: min {dur.min()} median {int(np.median(dur))} max {dur.max()} ns; loop (x staged -> cta done) median "
      f"{int(np.median(h[:, 6] - h[:, 4]))} ns; grid span {rel[:, 6].max()} ns")
