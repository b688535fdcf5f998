"""debug: where the roles of the tcgen05 GEMM's CTA 0 spend their cycles (needs QP_PROFILE_PHASES=1 QP_LIB_SUFFIX=_prof build)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops, _cabi
kind, bs = sys.argv[1], int(sys.argv[2])
M, K, dev = 14336, 4096, "cuda"
x = torch.randn((bs, K), device=dev).half()
rnd = lambda n: torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
fn = _cabi.lib().qp_debug_tc_prof
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
h = np.zeros(16, dtype=np.uint64)
n = 8
if kind.startswith("lut"):
    bits = int(kind[3:]); lut = torch.randn((1 << bits, 2), device=dev).half()
    bufs = [rnd(M * K * bits // 16) for _ in range(n + 2)]
    run = lambda b: ops.lut_gemm_tc(b, x, lut, M, K, bits, 2)
else:
    kv = int(kind[3:]); tl = torch.randn((512, 2), device=dev).half()
    bufs = [rnd(M * K * kv // 16) for _ in range(n + 2)]
    run = lambda b: ops.tcq_gemm_tc(b, x, tl, M, K, 9, kv)
run(bufs[0]); run(bufs[1]); torch.cuda.synchronize()
fn(h.ctypes.data_as(ctypes.c_void_p), 1)
for b in bufs[2:]:
    run(b)
torch.cuda.synchronize()
fn(h.ctypes.data_as(ctypes.c_void_p), 1)
h = h.astype(np.float64) / max(1.0, float(h[10]))
print(f"{kind} bs={bs}: per launch, CTA 0, cycles")
print(f"  kernel total {h[8]:9.0f}   prologue {h[7]:8.0f}   epilogue+teardown {h[9]:8.0f}")
print(f"  decode warps (sum of 16): loop {h[2]:9.0f}  wait payload {h[0]:9.0f}  wait free stage {h[1]:9.0f}  -> busy/warp {(h[2]-h[0]-h[1])/16:8.0f}")
print(f"  MMA warp: loop {h[4]:9.0f}  wait full {h[3]:9.0f}")
print(f"  loader:   loop {h[6]:9.0f}  wait free slot {h[5]:9.0f}")
