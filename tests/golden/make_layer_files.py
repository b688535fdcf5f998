"""Write per-layer `.pt` files in the reference's on-disk schema (build container only; needs /root/reference).

    python tests/golden/make_layer_files.py      # -> tests/golden/layers/<quantizer_str>/0_<layer key>.pt + expected.npz

What `quantize_layer.py` leaves on disk for one projection is `IncoherentLinear.save_info(path, quant_info)`
(lib/linear/incoherent_linear.py:467-484): {in_features, out_features, hadU, hadV, dtype, scale, Wscale, rot_info,
linear_info, bias, SU, SV, quant_info} with linear_info = the `_info()` of the quantized linear class
(tcq_linear.py:47-62, comb_linear.py:204-221, vq_linear.py:35-46) under `{quant_dir}/{quantizer_str}/{layer}_{key}.pt`
(incoherent_linear.py:382-386, eval_qdict.py:41-83).  The real quantizers (LDLQ / Viterbi) need a GPU and Hessians, so the
fixture takes the quantizer's OUTPUT as given -- random tail-biting trellis walks / random VQ codes -- and produces everything
downstream of it with the reference's own CPU code:
    packed codes   bitshift_codebook.pack_trellis (bitshift.py:296-329) + the kernel swizzle of tcq_quant.py:47-60 (inline in a
                   CUDA-only function there: restated by oracle.tcq_swizzle, which tests/test_oracle_golden.py pins against the
                   reference's decode_compressed), quant_op.pack_qweight for VQ
    linear_info    `_info()` of the reference's own QTIPLinearTCQ / CombtLinearTCQ / VQLinearPackTensorCore classes
    Wscale         the left-only rule of linear_to_incoherent_for_tcq (tcq_quant.py:124-126) on the rotated weight
    expected       rows of the reference's `recons` (the weights the packed codes stand for), for the loader test
A mini-Llama layer is used (hidden 512, 8 heads of 64, 2 KV heads, intermediate 28*32) so the files stay < 1 MB in total.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import REF, _import_reference, tail_biting_states  # noqa: E402

H, I, KVD = 512, 28 * 32, 128
LAYOUT = [  # layer key, (M, K), quantizer string
    ("self_attn.q_proj", (H, H), "tcq_6_none_0.9"),
    ("self_attn.k_proj", (KVD, H), "tcq_6_none_0.9"),
    ("self_attn.v_proj", (KVD, H), "tcq_6_none_0.9"),
    ("self_attn.o_proj", (H, H), "tcomb_6_7_0.5_none_0.9"),
    ("mlp.up_proj", (I, H), "ldlq_2_8_none_1.0"),
    ("mlp.gate_proj", (I, H), "ldlq_2_8_none_1.0"),
    ("mlp.down_proj", (H, I), "tcq_7_none_0.9"),
]


def load_ref_module(name):
    spec = importlib.util.spec_from_file_location(f"qp_reference_{name}", os.path.join(REF, "lib", "linear", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out_dir = os.path.join(HERE, "layers")
    _import_reference()  # chdir(REF), stubs glog / fast_hadamard_transform, neutralises .cuda()
    from oracle import qp_oracle as O
    from lib.codebook.bitshift import bitshift_codebook
    from lib.quantizer import quant_op
    from lib.utils import matmul_had
    from lib.utils.mem_op import get_quant_info
    tcq_mod, comb_mod, vq_mod = load_ref_module("tcq_linear"), load_ref_module("comb_linear"), load_ref_module("vq_linear")

    class OnCpu:  # vq_linear.py allocates with device='cuda'
        def __getattr__(self, k):
            v = getattr(torch, k)
            if k in ("randint", "randn"):
                return lambda *a, **kw: v(*a, **{**kw, "device": "cpu"})
            return v
    vq_mod.torch = OnCpu()

    rng = np.random.default_rng(2026)
    torch.manual_seed(2026)
    tlut = torch.load(f"{REF}/assets/lut_cache/kmeans_9_2.pt").half()
    vq_lut = torch.randn(256, 2).half()   # one shared VQ codebook, as vq_codebook does per (vec, bits)
    expected = {}
    for key, (M, K), qs in LAYOUT:
        qi = get_quant_info(qs)
        W = torch.randn(M, K) * (K ** -0.5)
        SU = ((torch.randn(K) > 0.0) * 2.0 - 1.0).float()
        Wr = matmul_had.matmul_hadUt_head(W * SU, K)                     # left-only rotation, tcq_quant.py:122

        def tcq_part(m, k, KV):
            cb = bitshift_codebook(L=16, KV=KV, V=2, tlut_bits=9, decode_mode="quantlut_sym", tlut=tlut)
            states = tail_biting_states(rng, (m // 16) * (k // 16), KV)
            packed = cb.pack_trellis(torch.from_numpy(states)).view(torch.int16).numpy()
            recon = O.tcq_expected_from_states(states.astype(np.uint16), tlut.numpy(), m, k, 9)  # == cb.recons, reordered
            return torch.from_numpy(O.tcq_swizzle(packed, m, k, KV)), recon, float(cb.lut.double().square().mean().sqrt())

        if qi["quantizer"] == "tcq_ldlq":
            trellis, What, lut_rms = tcq_part(M, K, qi["KV"])
            lin = tcq_mod.QTIPLinearTCQ(K, M, 16, 16, 16, qi["KV"], 2, 9, False, torch.float16)
            lin.trellis.data.copy_(trellis.reshape(lin.trellis.shape))
            lin.tlut.data.copy_(tlut)
        elif qi["quantizer"] == "combt_ldlq":
            (t1, W1, lut_rms), (t2, W2, _) = tcq_part(M, K // 2, qi["KV"][0]), tcq_part(M, K // 2, qi["KV"][1])
            What = np.concatenate([W1, W2], axis=1)
            lin = comb_mod.CombtLinearTCQ(K, M, 16, 16, (K // 2, K // 2), 16, tuple(qi["KV"]), 2, 9, False, torch.float16)
            lin.trellis1.data.copy_(t1.reshape(lin.trellis1.shape))
            lin.trellis2.data.copy_(t2.reshape(lin.trellis2.shape))
            lin.tlut.data.copy_(tlut)
        else:
            R, vec = qi["lut_bits"], qi["vec_sz"]
            codes = torch.from_numpy(rng.integers(0, 1 << R, size=(M, K // vec), dtype=np.int64))
            lin = vq_mod.VQLinearPackTensorCore(K, M, R, vec, False, torch.float16)
            lin.qweight.data.copy_(quant_op.pack_qweight(codes, vec, R).view(torch.int32).reshape(lin.qweight.shape))
            lin.lut.data.copy_(vq_lut)
            What = vq_lut[codes].reshape(M, K).numpy()
            lut_rms = float(vq_lut.double().square().mean().sqrt())
        scale_override = float(qs.split("_")[-1])
        Wscale = (Wr.double().square().mean(-1).sqrt() / (lut_rms * scale_override)).float() / scale_override
        info = {"in_features": K, "out_features": M, "hadU": K, "hadV": M, "dtype": torch.float32, "scale": 32.0,
                "Wscale": Wscale.cpu(), "rot_info": "skip_r", "linear_info": lin._info(), "bias": None,
                "SU": (1.0 / SU).cpu(), "SV": torch.ones(M), "quant_info": {**qi, "layer_key": f"model.layers.0.{key}",
                                                                           "save_path": f"{qs}/0_{key}.pt", "rot_info": "skip_r"}}
        path = os.path.join(out_dir, qs, f"0_{key}.pt")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(info, path)
        rows = np.sort(rng.choice(M, size=16, replace=False))
        expected[f"{key}:rows"] = rows
        expected[f"{key}:W"] = np.asarray(What, np.float16)[rows]
        print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KB)")
    np.savez_compressed(os.path.join(out_dir, "expected.npz"), **expected)


if __name__ == "__main__":
    main()
