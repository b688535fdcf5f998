"""debug: phase timeline of the TCQ GEMV with the fused x-producer prologue (QP_PROFILE_PHASES build, QP_LIB_SUFFIX=_prof)."""
import ctypes, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import _cabi
from qpalette._cabi import SPLIT_IN, check, lib
M, K = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "full"   # full: residual + norm + sign + hadamard; had: sign + hadamard only
dev = "cuda"
f16 = dict(dtype=torch.float16, device=dev)
tl = torch.randn((512, 2), **f16)
bufs = [(torch.randint(0, 256, (M * (K // 2) * 6 // 16,), dtype=torch.uint8, device=dev),
         torch.randint(0, 256, (M * (K // 2) * 7 // 16,), dtype=torch.uint8, device=dev)) for _ in range(14)]
h, h2 = torch.randn(K, **f16), torch.zeros(K, **f16)
acc, ws, norm, su = torch.randn(K, device=dev), torch.rand(K, **f16), torch.ones(K, **f16), torch.ones(K, **f16)
z1, z2 = torch.zeros(4096, device=dev), torch.zeros(28672, device=dev)
if len(sys.argv) > 4 and sys.argv[4] == "nozero":
    z1 = z2 = None
out = torch.zeros((1, M), dtype=torch.float32, device=dev)
p = lambda t: t.data_ptr() if t is not None else None
if mode == "full":
    xp = _cabi.XProd(p(h), p(h2), p(acc), p(ws), 64.0, p(norm), 1e-5, p(su), 1.0 / (math.sqrt(K) * 64), None, p(z1), z1.numel() if z1 is not None else 0, p(z2), z2.numel() if z2 is not None else 0)
else:
    xp = _cabi.XProd(p(h), None, None, None, 64.0, None, 1e-5, p(su), 1.0 / (math.sqrt(K) * 64), None, p(z1), z1.numel() if z1 is not None else 0, p(z2), z2.numel() if z2 is not None else 0)
st = torch.cuda.current_stream().cuda_stream
for b in bufs:
    check(lib().qp_tcq_gemv_fused(p(out), p(b[0]), p(b[1]), ctypes.byref(xp), p(tl), M, K, 9, 6, 7, SPLIT_IN, K // 2, st))
torch.cuda.synchronize()
hst = np.zeros((256, 8), dtype=np.uint64)
lib().qp_debug_phases.argtypes = [ctypes.c_void_p]
lib().qp_debug_phases(hst.ctypes.data_as(ctypes.c_void_p))
hst = hst[:148].astype(np.int64)
rel = hst - hst[:, 0].min()
names = ["start", "prefetch issued", "table built", "pdl wait done", "x produced", "warp0 done", "cta done"]
print(f"fused GEMV {M}x{K} mode={mode}")
for i, n in enumerate(names):
    print(f"  {n:16s} min {rel[:, i].min():7d} ns  median {int(np.median(rel[:, i])):7d} ns  max {rel[:, i].max():7d} ns")
x = np.zeros((256, 8), dtype=np.uint64)
lib().qp_debug_xphases.argtypes = [ctypes.c_void_p]
lib().qp_debug_xphases(x.ctypes.data_as(ctypes.c_void_p))
x = x[:148, :6].astype(np.int64) - hst[:, [3]]
for i, n in enumerate(["enter", "loads issued", "inputs arrived", "normalised", "warp stages + smem", "hadamard done"]):
    print(f"    x-producer {n:18s} median {int(np.median(x[:, i])):6d} ns after the dependency wait")
