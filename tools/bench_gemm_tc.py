"""batch sweep (BASELINE.json configs[3]): fused tcgen05 GEMM vs the reference-style path (dequantise to fp16 + cuBLAS)
and vs the bs<=8 GEMV, on a Llama-3.1-8B up_proj-shaped layer (14336 x 4096), weights rotated over > 2x L2."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops
from qpalette._cabi import SPLIT_IN

dev = "cuda"
M, K = 14336, 4096
cases = [("ldlq_2_4 (2 bpw)", "lut", (4, 2)), ("ldlq_2_6 (3 bpw)", "lut", (6, 2)), ("ldlq_2_8 (4 bpw)", "lut", (8, 2)),
         ("tcq_8 (4 bpw)", "tcq", 8), ("tcomb_6_7 (3.25 bpw)", "tcq", (6, 7))]
if len(sys.argv) > 1:
    cases = [c for c in cases if sys.argv[1] in c[0]]


def timeit(fns, iters=6):
    """all rotated launches captured in ONE CUDA graph (python / allocator overhead excluded)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns[:2]:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (iters * len(fns))


for name, kind, p in cases:
    if kind == "lut":
        bits, vec = p
        nb = M * K * bits // 8 // vec
        lut = torch.randn((1 << bits, vec), device=dev).half()
        bufs = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(max(2, int(300e6 // nb) + 1))]
    else:
        tl = torch.randn((512, 2), device=dev).half()
        if isinstance(p, tuple):
            nb = M * (K // 2) * (p[0] + p[1]) // 16
            bufs = [(torch.randint(0, 256, (M * (K // 2) * p[0] // 16,), dtype=torch.uint8, device=dev),
                     torch.randint(0, 256, (M * (K // 2) * p[1] // 16,), dtype=torch.uint8, device=dev)) for _ in range(13)]
        else:
            nb = M * K * p // 16
            bufs = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(max(2, int(300e6 // nb) + 1))]
    for bs in (1, 8, 16, 32, 64, 128):
        x = torch.randn((bs, K), device=dev).half()
        if kind == "lut":
            fused = [(lambda b=b: ops.lut_gemm_tc(b, x, lut, M, K, bits, vec)) for b in bufs]
            ref = [(lambda b=b: x @ ops.lut_dequant(b, lut, M, K, bits, vec).T) for b in bufs]
            gemv = [(lambda b=b: ops.lut_gemv(b, x, lut, M, K, bits, vec)) for b in bufs] if bs <= 8 else None
        elif isinstance(p, tuple):
            fused = [(lambda b=b: ops.tcq_gemm_tc(b[0], x, tl, M, K, 9, p[0], b[1], p[1], SPLIT_IN, K // 2)) for b in bufs]
            ref = [(lambda b=b: x @ ops.tcq_dequant(b[0], tl, M, K, 9, p[0], b[1], p[1], SPLIT_IN, K // 2).T) for b in bufs]
            gemv = [(lambda b=b: ops.tcq_gemv(b[0], x, tl, M, K, 9, p[0], b[1], p[1], SPLIT_IN, K // 2)) for b in bufs] if bs <= 8 else None
        else:
            fused = [(lambda b=b: ops.tcq_gemm_tc(b, x, tl, M, K, 9, p)) for b in bufs]
            ref = [(lambda b=b: x @ ops.tcq_dequant(b, tl, M, K, 9, p).T) for b in bufs]
            gemv = [(lambda b=b: ops.tcq_gemv(b, x, tl, M, K, 9, p)) for b in bufs] if bs <= 8 else None
        tf, tr = timeit(fused), timeit(ref)
        tg = timeit(gemv) if gemv else float("nan")
        flops = 2.0 * bs * M * K
        print(f"{name:22s} bs={bs:4d}  fused tcgen05 {tf:8.1f} us ({nb / tf / 1e3:7.1f} GB/s, {flops / tf / 1e6:7.1f} TFLOP/s)   "
              f"dequant+cuBLAS {tr:8.1f} us   gemv {tg:8.1f} us   speedup vs dequant+cuBLAS {tr / tf:5.2f}x", flush=True)
