"""Float64 restatement of the incoherent layer math WITH the fp16 rounding points of the reference graph, shared by the
GPU parity tests (test infrastructure; builds on oracle/qp_oracle.py).

Reference rounding points (everything is a fp16 tensor between ops):
  lib/linear/incoherent_linear.py:76-108   x.half() * SU -> hadamard (fp16 out) / scale -> linear(.) (fp32 acc -> fp16)
                                           * Wscale (fp16) * scale (fp16)
  lib/linear/incoherent_linear.py:324-338  act_fn(x_gate) (fp16) * x_up (fp16)
  lib/linear/incoherent_linear.py:486-506  same pattern with scale = 32
  HF LlamaRMSNorm / residual adds          weight * (x.float() * rstd).to(fp16);  h + y in fp16
  HF apply_rotary_pos_emb                  cos/sin cast to fp16
Between two rounding points the math below is float64; the kernels use fp32 -- differences are accumulation order only.
"""
import math

import numpy as np

from oracle import qp_oracle as O


def h16(a):
    return np.asarray(a, np.float16)


def f64(a):
    return np.asarray(a, np.float64)


def scaled_acc(acc, wscale, scale):
    """fp16(fp16(fp16(acc) * Wscale) * scale)   -- `linear(x).half() * Wscale * scale`"""
    return h16(h16(h16(acc).astype(np.float32) * h16(wscale).astype(np.float32)).astype(np.float32) * np.float32(scale))


def incoherent_in(x16, su, scale):
    """fp16( U^T (x * SU) / scale ) along the last dim"""
    x = f64(h16(x16)) * f64(su)
    shp = x.shape
    y = O.hadamard_ref(x.reshape(-1, shp[-1]))
    return h16(y.reshape(shp) / scale)


def rmsnorm16(h, w, eps):
    v = f64(h)
    rstd = 1.0 / np.sqrt((v * v).mean(-1, keepdims=True) + eps)
    return h16(h16(w).astype(np.float32) * h16(v * rstd).astype(np.float32))


def add16(a, b):
    return h16(h16(a).astype(np.float32) + h16(b).astype(np.float32))


def silu_mul16(up16, gate16):
    g = f64(gate16)
    act = h16(g / (1.0 + np.exp(-g)))
    return h16(act.astype(np.float32) * h16(up16).astype(np.float32))


def matvec(W, z16):
    """(bs, K) fp16 @ W(M, K).T -> float64 (bs, M), row-blocked to bound memory"""
    z = f64(z16).reshape(-1, W.shape[1])
    out = np.zeros((z.shape[0], W.shape[0]), np.float64)
    for r0 in range(0, W.shape[0], 1024):
        out[:, r0:r0 + 1024] = z @ f64(W[r0:r0 + 1024]).T
    return out


def incoherent_linear(x, W, su, wscale, scale):
    """IncoherentLinear (skip_r) / one projection of IncoherentMLP / IncoherentSdpaAttention: fp16 (bs, M)"""
    z = incoherent_in(x, su, scale)
    return scaled_acc(matvec(W, z), wscale, scale)


def rope16(x16, pos, inv_freq, fused=True):
    """x16 (..., D) fp16.  fused=True: fp16(x*cos + rot(x)*sin) evaluated in one go (the decode kernel);
    fused=False: fp16(fp16(x*cos) + fp16(rot(x)*sin)) (torch eager, HF apply_rotary_pos_emb)."""
    D = x16.shape[-1]
    ang = float(pos) * f64(np.asarray(inv_freq, np.float32))
    cos = h16(np.concatenate([np.cos(ang)] * 2).astype(np.float32))
    sin = h16(np.concatenate([np.sin(ang)] * 2).astype(np.float32))
    x = f64(h16(x16))
    rot = np.concatenate([-x[..., D // 2:], x[..., :D // 2]], -1)
    if fused:
        return h16(x * f64(cos) + rot * f64(sin))
    return add16(h16(x * f64(cos)), h16(rot * f64(sin)))


def attend(q16, Kc16, Vc16, n_rep):
    """q16 (nh, D); Kc16 / Vc16 (T, nkv, D) fp16 -> fp16 (nh, D); fp32-style softmax restated in float64"""
    nh, D = q16.shape
    out = np.zeros((nh, D), np.float64)
    for hd in range(nh):
        kv = hd // n_rep
        s = f64(Kc16[:, kv]) @ f64(q16[hd]) / math.sqrt(D)
        p = np.exp(s - s.max())
        out[hd] = (p @ f64(Vc16[:, kv])) / p.sum()
    return h16(out)


def decode_weight(p):
    """oracle-decoded fp16 weight of a qpalette.decode._Proj (cached on the object)"""
    if getattr(p, "_Wref", None) is None:
        t = lambda a: a.cpu().numpy()
        if p.kind == "tcq_ldlq":
            p._Wref = O.tcq_decode(t(p.codes1), t(p.lut), p.M, p.K, p.KV1, p.S)
        elif p.kind == "combt_ldlq":
            p._Wref = O.tcq_decode_combt(t(p.codes1), t(p.codes2), t(p.lut), p.M, p.K, p.KV1, p.KV2, p.S)
        elif p.simt:
            p._Wref = O.simt_decode(t(p.codes1), t(p.lut), p.M, p.K, p.bits, p.vec)
        else:
            p._Wref = O.lut_tc_decode(t(p.codes1), t(p.lut), p.M, p.K, p.bits, p.vec)
    return p._Wref


def decode_step_ref(r, x16, pos, caches, scale=64.0):
    """one decode step of an UNSHARDED qpalette.decode.DecodeRunner `r`, restated: x16 = fp16 embedding row of the input
    token, caches = [([k rows], [v rows]) per layer] (appended to).  Returns (h fp16, logits float64)."""
    sh = r.shape
    H, I, kvd = r.H, r.I, r.kvd
    D, nh, nkv = sh.head_dim, sh.num_attention_heads, sh.num_key_value_heads
    n16 = lambda a: a.cpu().numpy()
    h = h16(x16)
    inv = n16(r.inv_freq)

    def group(projs, z, width):
        acc = np.zeros((1, width))
        for p, off in projs:
            acc[:, off:off + p.M] = matvec(decode_weight(p), z)
        return acc

    for li, ly in enumerate(r.layers):
        z = incoherent_in(rmsnorm16(h, n16(ly["norm1"]), sh.rms_norm_eps), n16(ly["SU_qkv"]), scale)
        qkv = scaled_acc(group(ly["qkv"], z, H + 2 * kvd), n16(ly["W_qkv"]), scale).reshape(-1)
        q = rope16(qkv[:H].reshape(nh, D), pos, inv)
        ko, vo = (H + kvd, H) if ly.get("qvk") else (H, H + kvd)  # merge_qv layers keep q | v | k
        k = rope16(qkv[ko:ko + kvd].reshape(nkv, D), pos, inv)
        v = qkv[vo:vo + kvd].reshape(nkv, D)
        caches[li][0].append(k)
        caches[li][1].append(v)
        a = attend(q, np.stack(caches[li][0]), np.stack(caches[li][1]), nh // nkv)
        z = incoherent_in(a.reshape(-1), n16(ly["SU_o"]), scale)
        h = add16(h, scaled_acc(matvec(decode_weight(ly["o"]), z), n16(ly["W_o"]), scale).reshape(-1))
        z = incoherent_in(rmsnorm16(h, n16(ly["norm2"]), sh.rms_norm_eps), n16(ly["SU_ug"]), scale)
        ug = scaled_acc(group(ly["ug"], z, 2 * I), n16(ly["W_ug"]), scale).reshape(-1)
        z = incoherent_in(silu_mul16(ug[:I], ug[I:]), n16(ly["SU_dp"]), scale)
        h = add16(h, scaled_acc(matvec(decode_weight(ly["down"]), z), n16(ly["W_dp"]), scale).reshape(-1))
    xf = rmsnorm16(h, n16(r.final_norm), sh.rms_norm_eps)
    return h, f64(n16(r.lm_head)) @ f64(xf)
