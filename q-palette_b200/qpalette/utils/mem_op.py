"""Layer-shape tables, the quantizer-string grammar and the synthetic ("dummy") layer generator with the reference's
names (lib/utils/mem_op.py:2-307).  Shapes are generated from (hidden, kv_out, intermediate, nlayers)."""
import math

import torch

_MODELS = {
    # key: (hidden, kv_out, intermediate, nlayers)
    "2_7b": (4096, 4096, 11008, 32),
    "2_13b": (5120, 5120, 13824, 40),
    "2_70b": (8192, 1024, 28672, 80),
    "3_8b": (4096, 1024, 14336, 32),
    "3_3b": (3072, 1024, 8192, 28),
    "3_1b": (2048, 512, 8192, 16),
    "3_70b": (8192, 1024, 28672, 80),  # Llama-3.1-70B-shaped (same layer shapes as 2_70b)
}


def _layer_table(hidden, kv_out, inter, nlayers):
    io = lambda i, o: {"in_features": i, "out_features": o}
    return {
        "nlayers": nlayers,
        "self_attn.q_proj": io(hidden, hidden), "self_attn.k_proj": io(hidden, kv_out),
        "self_attn.v_proj": io(hidden, kv_out), "self_attn.o_proj": io(hidden, hidden),
        "mlp.gate_proj": io(hidden, inter), "mlp.up_proj": io(hidden, inter), "mlp.down_proj": io(inter, hidden),
    }


LAYER_INFO = {k: _layer_table(*v) for k, v in _MODELS.items()}


def get_layer_info(model_key):
    return LAYER_INFO["3_8b" if model_key == "3_8b_0" else model_key]


def get_quant_info(quantizer_str):
    """`tcq_{KV}_{hess}_{scale}`, `tcomb_{KV1}_{KV2}_{ratio}_{hess}_{scale}`, `comb_...`, `ldlq_{vec}_{bits}_{hess}_{scale}`,
    `sq_{bits}_...`, `vq2_{bits}_...` (quantize_layer.py:29-92, mem_op.py:271-307).  tlut_bits = 9 / KV+1."""
    f = quantizer_str.split("_")
    if quantizer_str.startswith("tcq"):
        kv = int(f[1])
        return {"quantizer_str": quantizer_str, "quantizer": "tcq_ldlq", "KV": kv, "V": 2,
                "tlut_bits": 9 if kv <= 8 else kv + 1}
    if quantizer_str.startswith("tcomb") or quantizer_str.startswith("comb"):
        kv1, kv2, ratio = int(f[1]), int(f[2]), float(f[3])
        return {"quantizer_str": quantizer_str,
                "quantizer": "combt_ldlq" if quantizer_str.startswith("tcomb") else "comb_ldlq",
                "KV": [kv1, kv2], "V": 2, "tlut_bits": 9 if max(kv1, kv2) <= 8 else max(kv1, kv2) + 1, "ratio": ratio}
    if quantizer_str.startswith("ldlq"):
        return {"quantizer_str": quantizer_str, "quantizer": "vq_ldlq", "vec_sz": int(f[1]), "lut_bits": int(f[2])}
    if quantizer_str.startswith("sq"):
        return {"quantizer_str": quantizer_str, "quantizer": "vq", "vec_sz": 1, "lut_bits": int(f[1])}
    if quantizer_str.startswith("vq2"):
        return {"quantizer_str": quantizer_str, "quantizer": "vq", "vec_sz": 2, "lut_bits": int(f[1])}
    if quantizer_str == "default":
        return {"quantizer_str": quantizer_str}
    raise ValueError(f"Unknown quantizer: {quantizer_str}")


def get_dummy_quant_results(model_key, layer_key, quantizer_str, device=None, generator=None, full_range=False,
                            in_features=None, out_features=None):
    """random-init synthetic layer info, same schema and value ranges as the reference's `--dummy` generator
    (mem_op.py:198-269): trellis = randint(0, 2^14) int16, qweight = randint(0, 2^30) int32, tlut/lut = randn fp16.
    `full_range=True` draws every bit uniformly instead (for bandwidth timing)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if in_features is None:
        li = get_layer_info(model_key)[layer_key]
        in_features, out_features = li["in_features"], li["out_features"]
    qi = get_quant_info(quantizer_str)
    g = generator

    def rint16(shape):
        if full_range:
            return torch.randint(-2 ** 15, 2 ** 15, shape, dtype=torch.int16, device=device, generator=g)
        return torch.randint(0, 2 ** 14, shape, dtype=torch.int16, device=device, generator=g)

    def rint32(shape):
        if full_range:
            return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int32, device=device, generator=g)
        return torch.randint(0, 2 ** 30, shape, dtype=torch.int32, device=device, generator=g)

    def rlut(shape):
        return torch.randn(shape, dtype=torch.float32, device=device, generator=g).to(torch.float16)

    info = {"quant_info": qi}
    if quantizer_str.startswith("ldlq") or quantizer_str.startswith("sq") or quantizer_str.startswith("vq2"):
        bits, vec = qi["lut_bits"], qi["vec_sz"]
        linear_info = {"in_features": in_features, "out_features": out_features, "lut_bits": bits, "dtype": torch.float16,
                       "vec_sz": vec, "qweight": rint32((out_features, bits * in_features // 32 // vec)),
                       "lut": rlut((2 ** bits, vec)), "bias": None}
    elif quantizer_str.startswith("tcq"):
        kv = qi["KV"]
        linear_info = {"in_features": in_features, "out_features": out_features, "td_x": 16, "td_y": 16, "L": 16,
                       "KV": kv, "V": 2, "tlut_bits": qi["tlut_bits"], "dtype": torch.float16,
                       "trellis": rint16(((out_features // 16) * (in_features // 16), math.ceil(256 * kv / 16 / 2))),
                       "tlut": rlut((2 ** qi["tlut_bits"], 2)), "bias": None}
    elif quantizer_str.startswith("tcomb"):
        assert qi["ratio"] == 0.5, "only ratio = 0.5 is supported (as in the reference)"
        in_part = (in_features // 2, in_features // 2)
        kv = qi["KV"]
        linear_info = {"in_features": in_features, "out_features": out_features, "td_x": 16, "td_y": 16,
                       "in_part": in_part, "L": 16, "KV": kv, "V": 2, "tlut_bits": qi["tlut_bits"],
                       "dtype": torch.float16,
                       "trellis1": rint16(((out_features // 16) * (in_part[0] // 16), math.ceil(256 * kv[0] / 16 / 2))),
                       "trellis2": rint16(((out_features // 16) * (in_part[1] // 16), math.ceil(256 * kv[1] / 16 / 2))),
                       "tlut": rlut((2 ** qi["tlut_bits"], 2)), "bias": None}
    elif quantizer_str.startswith("comb"):
        assert qi["ratio"] == 0.5
        out_part = (out_features // 2, out_features // 2)
        kv = qi["KV"]
        linear_info = {"in_features": in_features, "out_features": out_features, "td_x": 16, "td_y": 16,
                       "out_part": out_part, "L": 16, "KV": kv, "V": 2, "tlut_bits": qi["tlut_bits"],
                       "dtype": torch.float16,
                       "trellis1": rint16(((out_part[0] // 16) * (in_features // 16), math.ceil(256 * kv[0] / 16 / 2))),
                       "trellis2": rint16(((out_part[1] // 16) * (in_features // 16), math.ceil(256 * kv[1] / 16 / 2))),
                       "tlut": rlut((2 ** qi["tlut_bits"], 2)), "bias": None}
    elif quantizer_str == "default":
        linear_info = {"in_features": in_features, "out_features": out_features, "dtype": torch.float16, "bias": None}
    else:
        raise ValueError(f"Unknown quantizer: {quantizer_str}")
    info.update({"linear_info": linear_info, "in_features": in_features, "out_features": out_features,
                 "dtype": torch.float16, "bias": None})
    return info


def get_layer_mem(model_key, layer_key, quantizer_str="default"):
    """bytes of one quantized layer, codebook included (mem_op.py:309-326)."""
    li = get_layer_info(model_key)[layer_key]
    i, o = li["in_features"], li["out_features"]
    if quantizer_str == "default":
        return i * o * 2
    qi = get_quant_info(quantizer_str)
    if qi["quantizer"] in ("vq_ldlq", "vq"):
        return i * o * qi["lut_bits"] / qi["vec_sz"] / 8 + (2 ** qi["lut_bits"]) * qi["vec_sz"] * 2
    if qi["quantizer"] == "tcq_ldlq":
        return i * o * qi["KV"] / 2 / 8 + (2 ** qi["tlut_bits"]) * 4
    return i * o * (qi["KV"][0] + qi["KV"][1]) / 2 / 2 / 8 + (2 ** qi["tlut_bits"]) * 4
