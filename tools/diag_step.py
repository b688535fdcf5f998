"""stage-by-stage comparison of one decode step (1 layer, un-fused launch list) with the fp16-rounding-point restatement:
prints the rel-L2 of every intermediate buffer so that a logits mismatch can be attributed to a stage.
    python tools/diag_step.py [heads] [steps]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "q-palette_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
import _restate as R
from qpalette.decode import DecodeRunner, LlamaShape, uniform_qdict

heads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
shape = LlamaShape(hidden_size=512, intermediate_size=28 * 128, num_hidden_layers=1, num_attention_heads=heads,
                   num_key_value_heads=2, vocab_size=1024)
qd, mi = uniform_qdict(shape, "tcomb_6_7_0.5_none_0.9"), [["merge_qkv", "merge_ug"]]
r = DecodeRunner(shape, qd, mi, max_seq=80, seed=3, fused=False)
H, I, kvd = r.H, r.I, r.kvd
D, nh, nkv = shape.head_dim, shape.num_attention_heads, shape.num_key_value_heads
n16 = lambda a: a.cpu().numpy()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))
ly = r.layers[0]
Kc, Vc = [], []
tok = 5
r.reset(tok)
for step in range(steps):
    h = R.h16(n16(r.embed[tok]))
    r.step()
    torch.cuda.synchronize()
    out = {}
    z = R.incoherent_in(R.rmsnorm16(h, n16(ly["norm1"]), shape.rms_norm_eps), n16(ly["SU_qkv"]), 64.0)
    acc = np.zeros((1, H + 2 * kvd))
    for p, off in ly["qkv"]:
        acc[:, off:off + p.M] = R.matvec(R.decode_weight(p), z)
    out["acc_qkv"] = rel(n16(r.acc_qkv), acc.reshape(-1))
    qkv = R.scaled_acc(acc, n16(ly["W_qkv"]), 64.0).reshape(-1)
    q = R.rope16(qkv[:H].reshape(nh, D), step, n16(r.inv_freq))
    k = R.rope16(qkv[H:H + kvd].reshape(nkv, D), step, n16(r.inv_freq))
    v = qkv[H + kvd:].reshape(nkv, D)
    Kc.append(k); Vc.append(v)
    out["kcache"] = rel(n16(ly["kc"][step].float()), k)
    out["vcache"] = rel(n16(ly["vc"][step].float()), v)
    a = R.attend(q, np.stack(Kc), np.stack(Vc), nh // nkv)
    out["attn"] = rel(n16(r.attn.float()), a.reshape(-1))
    # continue from the GPU's attention output as well, to separate the stages
    for name, a_in in (("ref", a.reshape(-1)), ("gpu", n16(r.attn))):
        z = R.incoherent_in(a_in, n16(ly["SU_o"]), 64.0)
        acc_o = R.matvec(R.decode_weight(ly["o"]), z).reshape(-1)
        out[f"acc_o[{name}]"] = rel(n16(r.acc_o), acc_o)
    h2 = R.add16(h, R.scaled_acc(acc_o, n16(ly["W_o"]), 64.0))  # from the GPU attention output
    z = R.incoherent_in(R.rmsnorm16(h2, n16(ly["norm2"]), shape.rms_norm_eps), n16(ly["SU_ug"]), 64.0)
    out["x_h(ug in)"] = rel(n16(r.x_h.float()), z)
    acc = np.zeros((1, 2 * I))
    for p, off in ly["ug"]:
        acc[:, off:off + p.M] = R.matvec(R.decode_weight(p), z)
    out["acc_ug"] = rel(n16(r.acc_ug), acc.reshape(-1))
    ug = R.scaled_acc(n16(r.acc_ug), n16(ly["W_ug"]), 64.0).reshape(-1)  # from the GPU accumulators
    z = R.incoherent_in(R.silu_mul16(ug[:I], ug[I:]), n16(ly["SU_dp"]), 64.0)
    out["x_i(down in)"] = rel(n16(r.x_i.float()), z)
    acc_dn = R.matvec(R.decode_weight(ly["down"]), n16(r.x_i)).reshape(-1)
    out["acc_dn"] = rel(n16(r.acc_dn), acc_dn)
    h3 = R.add16(h2, R.scaled_acc(n16(r.acc_dn), n16(ly["W_dp"]), 64.0))
    out["h"] = rel(n16(r.h.float()), h3)
    xf = R.rmsnorm16(n16(r.h), n16(r.final_norm), shape.rms_norm_eps)
    out["xf"] = rel(n16(r.xf.float()), xf)
    out["logits"] = rel(n16(r.logits), R.f64(n16(r.lm_head)) @ R.f64(n16(r.xf)))
    print(f"step {step}: " + "  ".join(f"{k} {v:.1e}" for k, v in out.items()), flush=True)
    tok = int(r.token.item())
