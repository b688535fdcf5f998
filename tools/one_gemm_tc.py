"""a few launches of the fused tcgen05 GEMM for profiling: python tools/one_gemm_tc.py {lut4|lut8|tcq8|tcomb} BS"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops
from qpalette._cabi import SPLIT_IN

kind, bs = sys.argv[1], int(sys.argv[2])
M, K, dev = 14336, 4096, "cuda"
x = torch.randn((bs, K), device=dev).half()
rnd = lambda n: torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
for it in range(4):
    if kind.startswith("lut"):
        bits = int(kind[3:])
        ops.lut_gemm_tc(rnd(M * K * bits // 16), x, torch.randn((1 << bits, 2), device=dev).half(), M, K, bits, 2)
    elif kind == "tcq8":
        ops.tcq_gemm_tc(rnd(M * K * 8 // 16), x, torch.randn((512, 2), device=dev).half(), M, K, 9, 8)
    else:
        ops.tcq_gemm_tc(rnd(M * (K // 2) * 6 // 16), x, torch.randn((512, 2), device=dev).half(), M, K, 9, 6,
                        rnd(M * (K // 2) * 7 // 16), 7, SPLIT_IN, K // 2)
torch.cuda.synchronize()
print("ok")
