"""debug: which layer / projection of a (mixed-scheme) model makes the fused and the un-fused launch lists diverge.
    python tools/diag_fused.py [config=figure1d] [layers=3]
1. logits of fused vs un-fused runners for 1..layers layers;
2. every projection of those layers stand-alone: plain GEMV on x = qp_fused_norm_had(h) vs the fused-prologue GEMV on h."""
import ctypes, json, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "q-palette_b200")):
    sys.path.insert(0, p_)
from qpalette import _cabi
from qpalette._cabi import check, lib
from qpalette.decode import DecodeRunner, LLAMA31_8B, uniform_qdict

name = sys.argv[1] if len(sys.argv) > 1 else "figure1d"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name == "uniform":
    qd, mi = uniform_qdict(LLAMA31_8B, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]] * 32
else:
    cfg = json.load(open(os.path.join(ROOT, "configs", name + ".json")))
    qd, mi = {k: tuple(v) for k, v in cfg["qdict"].items()}, cfg["merge_info"]
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
L = lib()
p = lambda t: t.data_ptr() if t is not None else None
for nl in range(1, layers + 1):
    outs = []
    for fused in (True, False):
        r = DecodeRunner(LLAMA31_8B, qd, mi, max_seq=16, seed=11, num_layers=nl, fused=fused)
        r.reset(9)
        lg = []
        for _ in range(2):
            r.step(); torch.cuda.synchronize()
            lg.append(r.logits.clone())
        outs.append(lg)
        if fused:
            rf = r
    print(f"layers={nl}: fused vs unfused logits rel-L2 step0 {rel(outs[0][0], outs[1][0]):.2e} step1 {rel(outs[0][1], outs[1][1]):.2e}", flush=True)
# stand-alone projections of the fused runner's layers
st = torch.cuda.current_stream().cuda_stream
r = rf
for li, ly in enumerate(r.layers):
    for gname, projs, K in (("qkv", ly["qkv"], r.H), ("o", [(ly["o"], 0)], r.H), ("ug", ly["ug"], r.H), ("down", [(ly["down"], 0)], r.I)):
        h = torch.randn(K, device="cuda").half()
        su = ((torch.randn(K, device="cuda") > 0).half() * 2 - 1)
        nw = (torch.rand(K, device="cuda") + 0.5).half()
        scale = 1.0 / (math.sqrt(K) * 64.0)
        x = torch.zeros(K, dtype=torch.float16, device="cuda")
        hh = h.clone()
        check(L.qp_fused_norm_had(p(x), p(hh), 0, None, None, 0.0, p(nw), 1e-5, p(su), K, scale, 1, None, 0, st))
        for pr, off in projs:
            a = torch.zeros(pr.M, dtype=torch.float32, device="cuda")
            pr.launch(p(a), p(x), st)
            torch.cuda.synchronize()
            msg = f"layer {li} {gname:4s} {pr.qs:24s} {pr.M}x{pr.K}: plain |out| {float(a.norm()):.3e}"
            if pr.can_fuse() and K == r.H or pr.can_fuse():
                b = torch.zeros(pr.M, dtype=torch.float32, device="cuda")
                h_out = torch.zeros(K, dtype=torch.float16, device="cuda")
                xo = torch.zeros(K, dtype=torch.float16, device="cuda")
                xp = _cabi.XProd(p(h), p(h_out), None, None, 64.0, p(nw), 1e-5, p(su), scale, p(xo), None, 0, None, 0)
                try:
                    pr.launch_fused(p(b), xp, st)
                    torch.cuda.synchronize()
                    msg += f"  fused vs plain rel-L2 {rel(b, a):.2e}  x_out vs x {rel(xo.float(), x.float()):.2e}"
                except Exception as ex:
                    msg += f"  fused launch failed: {ex}"
            print(msg, flush=True)
