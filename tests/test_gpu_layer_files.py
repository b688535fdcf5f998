"""The non-dummy loading path: per-layer `.pt` files in the reference's `save_info` schema (written by
tests/golden/make_layer_files.py with the reference's own packers) loaded through
`Incoherent{Linear,MLP,SdpaAttention}.gen_layer_from_quantizer_str_and_key(..., dummy=False)` -- the calls
model/incoherent_llama.py / eval_qdict.py make (lib/linear/incoherent_linear.py:259-277,382-394,548-560) -- and run on the
GPU.  Decoded weights must equal the reference's `recons` of the same codes bit for bit (sampled rows, expected.npz); layer
outputs must match the restatement with the graph's fp16 rounding points within rel-L2 1e-3."""
import os

import numpy as np
import pytest
import torch

import _restate as R
from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QDIR = os.path.join(ROOT, "tests", "golden", "layers")
TOL = 1e-3
QS = {"self_attn.q_proj": "tcq_6_none_0.9", "self_attn.k_proj": "tcq_6_none_0.9", "self_attn.v_proj": "tcq_6_none_0.9",
      "self_attn.o_proj": "tcomb_6_7_0.5_none_0.9", "mlp.up_proj": "ldlq_2_8_none_1.0", "mlp.gate_proj": "ldlq_2_8_none_1.0",
      "mlp.down_proj": "tcq_7_none_0.9"}


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cfg():
    from qpalette.decode import LlamaShape
    return LlamaShape(hidden_size=512, intermediate_size=28 * 32, num_hidden_layers=1, num_attention_heads=8,
                      num_key_value_heads=2, vocab_size=64)


def load(key):
    return torch.load(os.path.join(QDIR, QS[key], f"0_{key}.pt"), weights_only=False)


def decoded(info):
    li = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in info["linear_info"].items()}
    M, K = li["out_features"], li["in_features"]
    if "trellis" in li:
        return O.tcq_decode(li["trellis"], li["tlut"], M, K, li["KV"], li["tlut_bits"])
    if "trellis1" in li:
        return O.tcq_decode_combt(li["trellis1"], li["trellis2"], li["tlut"], M, K, li["KV"][0], li["KV"][1], li["tlut_bits"])
    return O.lut_tc_decode(li["qweight"], li["lut"], M, K, li["lut_bits"], li["vec_sz"])


@pytest.mark.parametrize("key", list(QS))
def test_incoherent_linear_from_file(key):
    """IncoherentLinear.gen_layer_from_quantizer_str_and_key(dummy=False): decoded weights == reference recons, output within
    tolerance; `save_info` writes the schema back unchanged."""
    from qpalette import ops
    from qpalette.linear import IncoherentLinear
    exp = np.load(os.path.join(QDIR, "expected.npz"))
    info = load(key)
    layer = IncoherentLinear.gen_layer_from_quantizer_str_and_key(cfg(), QDIR, QS[key], f"0_{key}", merge_layers=True).cuda()
    assert layer.rot_info == "skip_r" and layer.skip_r and not layer.skip_l
    M, K = info["out_features"], info["in_features"]
    # the quantized linear's weights, dequantized on the GPU, against the reference's recons of the same codes
    lin = layer.linear
    eye = torch.eye(K, device="cuda").half()
    Wgpu = torch.cat([lin(eye[i:i + 256]) for i in range(0, K, 256)]).T.contiguous().cpu().numpy()  # bs > 8: dequant path
    rows = exp[f"{key}:rows"]
    assert np.array_equal(Wgpu[rows].view(np.uint16), exp[f"{key}:W"].view(np.uint16)), "decoded weights differ from the reference recons"
    W = decoded(info)
    assert np.array_equal(W[rows].view(np.uint16), exp[f"{key}:W"].view(np.uint16))
    x = torch.randn(3, K, device="cuda").half()
    out = layer(x).float().cpu().numpy()
    ref = R.incoherent_linear(x.cpu().numpy(), W, info["SU"].numpy(), info["Wscale"].numpy(), 32.0)
    assert rel_l2(out, ref) <= TOL
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        layer.save_info(os.path.join(d, "x.pt"), info["quant_info"])
        back = torch.load(os.path.join(d, "x.pt"), weights_only=False)
    assert set(back) == set(info)
    for k in ("in_features", "out_features", "hadU", "hadV", "rot_info", "scale"):
        assert back[k] == info[k], k
    for k in ("Wscale", "SU", "SV"):
        assert torch.allclose(back[k].float(), info[k].float(), rtol=1e-3)
    assert set(back["linear_info"]) == set(info["linear_info"])


@pytest.mark.parametrize("merge_ug", [False, True])
def test_incoherent_mlp_from_files(merge_ug):
    from qpalette.linear import IncoherentMLP
    up, gate, down = load("mlp.up_proj"), load("mlp.gate_proj"), load("mlp.down_proj")
    mlp = IncoherentMLP.gen_layer_from_quantizer_str_and_key(cfg(), QDIR, QS["mlp.up_proj"], QS["mlp.gate_proj"], QS["mlp.down_proj"],
                                                             "0_mlp.up_proj", "0_mlp.gate_proj", "0_mlp.down_proj",
                                                             merge_ug=merge_ug).cuda()
    x = torch.randn(1, 1, 512, device="cuda").half()
    xs = x.cpu().numpy().reshape(1, -1)
    n = lambda t: t.numpy()
    u = R.incoherent_linear(xs, decoded(up), n(up["SU"]), n(up["Wscale"]), 64.0)
    g = R.incoherent_linear(xs, decoded(gate), n(up["SU"]), n(gate["Wscale"]), 64.0)
    ref = R.incoherent_linear(R.silu_mul16(u, g), decoded(down), n(down["SU"]), n(down["Wscale"]), 64.0)
    assert rel_l2(mlp(x).float().cpu().numpy().reshape(1, -1), ref) <= TOL


@pytest.mark.parametrize("merge", [dict(), dict(merge_qkv=True), dict(merge_kv=True)])
def test_incoherent_attention_from_files(merge):
    from qpalette.linear import IncoherentSdpaAttention, StaticKVCache
    q, k, v, o = (load(f"self_attn.{n}_proj") for n in "qkvo")
    attn = IncoherentSdpaAttention.gen_layer_from_quantizer_str_and_key(
        cfg(), 0, QDIR, QS["self_attn.q_proj"], QS["self_attn.k_proj"], QS["self_attn.v_proj"], QS["self_attn.o_proj"],
        "0_self_attn.q_proj", "0_self_attn.k_proj", "0_self_attn.v_proj", "0_self_attn.o_proj", **merge).cuda()
    x = torch.randn(1, 1, 512, device="cuda").half()
    xs = x.cpu().numpy().reshape(1, -1)
    n = lambda t: t.numpy()
    mq, mk, mv = attn.compute_qkv(x)
    for got, inf in ((mq, q), (mk, k), (mv, v)):
        ref = R.incoherent_linear(xs, decoded(inf), n(q["SU"]), n(inf["Wscale"]), 64.0)
        assert rel_l2(got.float().cpu().numpy().reshape(1, -1), ref) <= TOL
    cache = StaticKVCache(1, 8, 2, 64)
    y, _, _ = attn(x, past_key_value=cache, cache_position=torch.tensor([0], device="cuda"))
    # one position: softmax over a single key = 1, so the attention output is v repeated over the query groups
    vr = R.incoherent_linear(xs, decoded(v), n(q["SU"]), n(v["Wscale"]), 64.0).reshape(2, 64)
    a = np.repeat(vr, 4, axis=0).reshape(1, -1)
    ref = R.incoherent_linear(a, decoded(o), n(o["SU"]), n(o["Wscale"]), 64.0)
    assert rel_l2(y.float().cpu().numpy().reshape(1, -1), ref) <= TOL
