import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200")); sys.path.insert(0, ROOT)
from qpalette import ops
from qpalette._cabi import SPLIT_IN
from oracle import qp_oracle as O
rng = np.random.default_rng(0)
d = lambda a: torch.from_numpy(a).cuda()
for M, K, bs in ((256, 512, 9), (1024, 1024, 16), (1024, 2048, 24), (512, 1024, 32), (512, 2560, 32), (4096, 14336, 20), (512, 1024, 40), (256, 512, 64), (32, 64, 12)):
    b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
    tl = (rng.standard_normal((512, 2)) * 0.9).astype(np.float16)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemm_mma(d(b1), d(x), d(tl), M, K, 9, 6, d(b2), 7, SPLIT_IN, K // 2)
    torch.cuda.synchronize()
    ref = O.gemv_ref(O.tcq_decode_combt(b1, b2, tl, M, K, 6, 7, 9), x)
    o = out.cpu().numpy()
    print(M, K, bs, "rel-L2", np.linalg.norm(o - ref) / np.linalg.norm(ref), flush=True)
# timing at 14336 x 4096 (graph over rotated buffers)
M, K = 14336, 4096
tl = torch.randn((512, 2), device="cuda").half()
bufs = [(torch.randint(0, 256, (M * (K // 2) * 6 // 16,), dtype=torch.uint8, device="cuda"),
         torch.randint(0, 256, (M * (K // 2) * 7 // 16,), dtype=torch.uint8, device="cuda")) for _ in range(12)]
for bs in (9, 16, 24, 32, 48, 64):
    x = torch.randn((bs, K), device="cuda").half()
    out = torch.zeros((bs, M), dtype=torch.float32, device="cuda")
    for name, fn in (("mma", ops.tcq_gemm_mma), ("tc", None)):
        if fn is None:
            saved = ops.MMA_GEMM_MAX_BS; ops.MMA_GEMM_MAX_BS = 0; fn = ops.tcq_gemm_tc
        run = lambda: [fn(b[0], x, tl, M, K, 9, 6, b[1], 7, SPLIT_IN, K // 2, out=out, accumulate=True) for b in bufs]
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        g.replay(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.replay()
        b_.record(); torch.cuda.synchronize()
        print(f"14336x4096 tcomb_6_7 bs={bs:3d} {name:4s} {a.elapsed_time(b_) * 1e3 / 10 / len(bufs):7.2f} us", flush=True)
        if name == "tc":
            ops.MMA_GEMM_MAX_BS = saved
# VQ sweep of BASELINE config 4 (ldlq_2_4 / 2_6 / 2_8) at the same shape
for R in (4, 6, 8):
    lut = torch.randn((1 << R, 2), device="cuda").half()
    nbytes = M * K * R // 16
    lbufs = [torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda") for _ in range(max(3, int(300e6 // nbytes)))]
    for bs in (16, 32, 64):
        x = torch.randn((bs, K), device="cuda").half()
        out = torch.zeros((bs, M), dtype=torch.float32, device="cuda")
        for name in ("mma", "tc"):
            saved = ops.MMA_GEMM_MAX_BS
            ops.MMA_GEMM_MAX_BS = 128 if name == "mma" else 0
            run = lambda: [ops.lut_gemm_tc(b, x, lut, M, K, R, 2, out=out, accumulate=True) for b in lbufs]
            run(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            ops.MMA_GEMM_MAX_BS = saved
            g.replay(); torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                g.replay()
            b_.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b_) * 1e3 / 10 / len(lbufs)
            print(f"14336x4096 ldlq_2_{R} bs={bs:3d} {name:4s} {us:7.2f} us  {nbytes / us / 1e3:7.1f} GB/s", flush=True)
