"""Micro-benchmark of the fused dequant-GEMV kernels at Llama shapes (device-resident inputs, CUDA events,
weights rotated over > 2x L2 worth of distinct buffers so every launch streams from HBM).

    python tools/bench_gemv.py [--cases tcq|lut|simt|all] [--iters 200]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette import ops  # noqa: E402
from qpalette._cabi import SPLIT_IN, SPLIT_NONE  # noqa: E402

PEAK = 6451.2
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def time_fn(fns, iters):
    """fns: list of callables (one per rotated buffer).  returns avg ms per call."""
    for f in fns[: min(len(fns), 4)]:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ncopies(nbytes):
    return max(2, int(300e6 // nbytes) + 1)


def tcq_case(M, K, kv, S=9, bs=1, graph=True):
    dev = "cuda"
    two = isinstance(kv, tuple)
    tl = torch.randn((1 << S, 2), device=dev).half()
    x = torch.randn((bs, K), device=dev).half()
    if two:
        nb = M * (K // 2) * (kv[0] + kv[1]) // 16
    else:
        nb = M * K * kv // 16
    bufs = []
    for _ in range(ncopies(nb)):
        if two:
            b1 = torch.randint(0, 256, (M * (K // 2) * kv[0] // 16,), dtype=torch.uint8, device=dev)
            b2 = torch.randint(0, 256, (M * (K // 2) * kv[1] // 16,), dtype=torch.uint8, device=dev)
            bufs.append((b1, b2))
        else:
            bufs.append((torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev), None))
    out = torch.zeros((bs, M), dtype=torch.float32, device=dev)

    def mk(b):
        if two:
            return lambda: ops.tcq_gemv(b[0], x, tl, M, K, S, kv[0], b[1], kv[1], SPLIT_IN, K // 2, out=out, accumulate=True)
        return lambda: ops.tcq_gemv(b[0], x, tl, M, K, S, kv, out=out, accumulate=True)

    fns = [mk(b) for b in bufs]
    alg = nb + 2 * bs * K + 4 * bs * M + (1 << S) * 4
    return fns, alg


def lut_case(M, K, bits, vec, bs=1, simt=False):
    dev = "cuda"
    lut = torch.randn((1 << bits, vec), device=dev).half()
    x = torch.randn((bs, K), device=dev).half()
    nb = M * K * bits // 8 // vec
    bufs = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(ncopies(nb))]
    out = torch.zeros((bs, M), dtype=torch.float32, device=dev)
    if simt:
        fns = [(lambda b=b: ops.simt_gemv(b, x, lut, M, K, bits, vec)) for b in bufs]
        alg = nb + 2 * bs * K + 2 * bs * M + (1 << bits) * vec * 2
    else:
        fns = [(lambda b=b: ops.lut_gemv(b, x, lut, M, K, bits, vec, out=out, accumulate=True)) for b in bufs]
        alg = nb + 2 * bs * K + 4 * bs * M + (1 << bits) * vec * 2
    return fns, alg


def graphed(fns):
    """capture the whole rotation (one launch per distinct buffer) in ONE CUDA graph: python/ctypes and per-graph
    launch overhead are excluded; returns a single-callable list and the number of launches per call."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    return [g.replay], len(fns)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="tcq")
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--nograph", action="store_true")
    ap.add_argument("--one", default="", help="single case, e.g. tcq:4096:14336:6,7 or lut:4096:14336:8,2")
    a = ap.parse_args()
    rows = []
    cases = []
    if a.cases in ("tcq", "all"):
        for shape in ((4096, 14336), (14336, 4096), (4096, 4096), (6144, 4096), (28672, 4096), (1024, 4096)):
            for kv in ((6, 7), 6, 7, 8):
                cases.append(("tcq", shape, kv))
        for kv in (2, 3, 4, 5, 9, 10):
            cases.append(("tcq", (4096, 14336), kv))
        for bs in (2, 4, 8):
            cases.append(("tcq", (4096, 4096), (6, 7), bs))
    if a.cases in ("lut", "all"):
        for shape in ((4096, 14336), (4096, 4096)):
            for vec, bits in ((2, 4), (2, 6), (2, 8), (2, 10), (2, 12), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)):
                cases.append(("lut", shape, (bits, vec)))
    if a.cases in ("simt", "all"):
        for shape in ((4096, 14336), (14336, 4096)):
            for vec, bits in ((2, 6), (2, 8), (2, 10), (1, 4), (1, 6), (4, 8), (4, 10), (4, 12)):
                cases.append(("simt", shape, (bits, vec)))
    if a.one:
        kind, M, K, p = a.one.split(":")
        pp = tuple(int(v) for v in p.split(","))
        cases = [(kind, (int(M), int(K)), pp if (kind != "tcq" or len(pp) > 1) else pp[0])]
    for c in cases:
        kind, (M, K), p = c[0], c[1], c[2]
        bs = c[3] if len(c) > 3 else 1
        if kind == "tcq":
            S = 9 if (max(p) if isinstance(p, tuple) else p) <= 8 else (max(p) if isinstance(p, tuple) else p) + 1
            fns, alg = tcq_case(M, K, p, S, bs)
        elif kind == "lut":
            fns, alg = lut_case(M, K, p[0], p[1], bs)
        else:
            fns, alg = lut_case(M, K, p[0], p[1], bs, simt=True)
        per = 1
        if not a.nograph:
            fns, per = graphed(fns)
        ms = time_fn(fns, max(3, a.iters // per)) / per
        gbs = alg / ms / 1e6
        rows.append((kind, M, K, p, bs, ms * 1e3, gbs, gbs / PEAK))
        print(f"{kind:5s} {M:6d}x{K:<6d} {str(p):9s} bs={bs} {ms*1e3:8.2f} us  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of measured {PEAK:.0f}",
              flush=True)
        del fns
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
