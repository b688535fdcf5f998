"""debug: per-phase timeline of the TCQ GEMV kernel (needs the QP_PROFILE_PHASES build: QP_LIB_SUFFIX=_prof)."""
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
for b in bufs:
    ops.tcq_gemv(b[0], x, tl, M, K, 9, kv[0], b[1], kv[1], SPLIT_IN, K // 2, out=out, accumulate=True)
torch.cuda.synchronize()
h = np.zeros((256, 8), dtype=np.uint64)
_cabi.lib().qp_debug_phases.argtypes = [ctypes.c_void_p]
_cabi.lib().qp_debug_phases(h.ctypes.data_as(ctypes.c_void_p))
h = h[:148].astype(np.int64)
t0 = h[:, 0].min()
rel = h - t0
names = ["start", "tlut copied", "table built", "pdl wait done", "x staged", "warp0 done", "cta done"]
for i, n in enumerate(names):
    print(f"{n:14s} min {rel[:, i].min():7d} ns  median {int(np.median(rel[:, i])):7d} ns  max {rel[:, i].max():7d} ns")
