"""in-graph time of the decode glue kernels (32 back-to-back launches per graph replay, PDL edges as in the real step)"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-palette_b200"))
from qpalette._cabi import lib, check

dev = "cuda"
H, I, V = 4096, 14336, 128256
f16 = dict(dtype=torch.float16, device=dev)
h = torch.randn(H, **f16); x_h = torch.zeros(H, **f16); x_i = torch.zeros(I, **f16)
acc_h = torch.randn(H, device=dev); acc_ug = torch.randn(2 * I, device=dev); acc_qkv = torch.randn(6144, device=dev)
w_h = torch.rand(H, **f16); w_ug = torch.rand(2 * I, **f16); w_qkv = torch.rand(6144, **f16)
norm = torch.ones(H, **f16); su_h = torch.ones(H, **f16); su_i = torch.ones(I, **f16)
inv = torch.rand(64, device=dev); kc = torch.zeros((256, 8, 128), **f16); vc = torch.zeros_like(kc)
pos = torch.full((1,), 64, dtype=torch.int32, device=dev); tok = torch.zeros(1, dtype=torch.int32, device=dev)
attn = torch.zeros(H, **f16)
L = lib()
p = lambda t: t.data_ptr()
s_h, s_i = 1 / (math.sqrt(H) * 64), 1 / (math.sqrt(I) * 64)


def timeit(name, fn, n=32, iters=20):
    st = torch.cuda.current_stream().cuda_stream
    fn(st); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"{name:34s} {a.elapsed_time(b) * 1e3 / (iters * n):8.2f} us per launch (in graph, dependent chain)")


timeit("step_advance (launch floor)", lambda st: check(L.qp_step_advance(p(pos), None, p(tok), 0, st)))
pos.fill_(64)
timeit("fused_norm_had n=4096 full", lambda st: check(L.qp_fused_norm_had(p(x_h), p(h), 1, p(acc_h), p(w_h), 64.0, p(norm), 1e-5, p(su_h), H, s_h, 1, p(acc_ug), 2 * I, st)))
timeit("fused_norm_had n=4096 full, no zeroing", lambda st: check(L.qp_fused_norm_had(p(x_h), p(h), 1, p(acc_h), p(w_h), 64.0, p(norm), 1e-5, p(su_h), H, s_h, 1, None, 0, st)))
timeit("fused_norm_had n=4096 had only", lambda st: check(L.qp_fused_norm_had(p(x_h), p(attn), 0, None, None, 0.0, None, 0.0, p(su_h), H, s_h, 1, p(acc_h), H, st)))
timeit("fused_norm_had n=4096 norm only", lambda st: check(L.qp_fused_norm_had(p(x_h), p(h), 1, p(acc_h), p(w_h), 64.0, p(norm), 1e-5, None, H, 1.0, 0, None, 0, st)))
timeit("silu_mul_had I=14336", lambda st: check(L.qp_silu_mul_had(p(x_i), p(acc_ug), p(w_ug), 64.0, p(su_i), I, s_i, p(acc_h), H, st)))
timeit("silu_mul_had I=14336 no zeroing", lambda st: check(L.qp_silu_mul_had(p(x_i), p(acc_ug), p(w_ug), 64.0, p(su_i), I, s_i, None, 0, st)))
MAXS = 4096 + 64
kc = torch.zeros(MAXS * 8 * 128, dtype=torch.float16, device="cuda")
vc = torch.zeros(MAXS * 8 * 128, dtype=torch.float16, device="cuda")
ascr = torch.zeros(int(L.qp_rope_attention_scratch_bytes(32, 128, MAXS)), dtype=torch.uint8, device="cuda")
for ps in (16, 64, 127, 150, 512, 2048, 4096):
    pos.fill_(ps)
    timeit(f"rope_attention pos={ps}", lambda st: check(L.qp_rope_attention(p(attn), p(acc_qkv), p(w_qkv), 64.0, p(inv), p(kc), p(vc), p(pos), 32, 8, 128, MAXS, 0, None, 0, p(ascr), st)))
