"""a few launches of the batched mma GEMM (14336 x 4096, tcomb_6_7) for an ncu capture: python tools/ncu_mma.py [bs]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops
from qpalette._cabi import SPLIT_IN
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M, K = 14336, 4096
tl = torch.randn((512, 2), device="cuda").half()
bufs = [(torch.randint(0, 256, (M * (K // 2) * 6 // 16,), dtype=torch.uint8, device="cuda"),
         torch.randint(0, 256, (M * (K // 2) * 7 // 16,), dtype=torch.uint8, device="cuda")) for _ in range(12)]
x = torch.randn((bs, K), device="cuda").half()
out = torch.zeros((bs, M), dtype=torch.float32, device="cuda")
for b in bufs:
    ops.tcq_gemm_mma(b[0], x, tl, M, K, 9, 6, b[1], 7, SPLIT_IN, K // 2, out=out, accumulate=True)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
