"""Incoherence wrappers with the reference's interface (lib/linear/incoherent_linear.py:13-560):

    out = Q( U^T (x * SU) / s ) * Wscale * s          s = 32 (IncoherentLinear) / 64 (IncoherentMLP, IncoherentSdpaAttention)

`x * SU -> FWHT (+28x28 factor) -> / s -> fp16` is ONE kernel (qp_hadamard) and `fp16(acc) * Wscale * s` (+ SiLU*mul for the
merged up|gate) is one kernel (qp_scale_epilogue); the reference issues ~6 elementwise kernels per linear and relies on
torch.compile to fuse them.  All shipped quantizations are left-only (`rot_info="skip_r"`, quantize_layer.py:126-130); the
right-side transform of IncoherentLinear is kept for completeness.

The attention module does not depend on HF transformers: RoPE (incl. llama3 frequency scaling) and the KV cache interface
(`past_key_value.update(k, v, layer_idx, cache_kwargs)`) are provided here.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._cabi import EPI_NONE, EPI_SILU_MUL
from ..utils.matmul_had import get_hadK, matmul_hadU_head_cuda
from .comb_linear import CombLinearTCQ, CombtLinearTCQ
from .tcq_linear import QTIPLinearTCQ, _default_device
from .vq_linear import VQLinearPackSIMT, VQLinearPackTensorCore

MODEL_KEYS = {"meta-llama/Llama-3.1-8B": "3_8b", "meta-llama/Llama-3.2-1B": "3_1b", "meta-llama/Llama-3.2-3B": "3_3b",
              "meta-llama/Llama-3.1-70B": "3_70b"}


def _is_vq(qs):
    return "sq" in qs or "vq" in qs or "ldlq" in qs


def linear_class_for(quantizer_str, use_simt=False):
    """class dispatch by substring, in the reference's order (incoherent_linear.py:13-26)."""
    if "tcq" in quantizer_str:
        return QTIPLinearTCQ
    if _is_vq(quantizer_str):
        return VQLinearPackSIMT if use_simt else VQLinearPackTensorCore
    if "tcomb" in quantizer_str:
        return CombtLinearTCQ
    if "comb" in quantizer_str:
        return CombLinearTCQ
    return None


def make_linear(info, use_simt=False):
    cls = linear_class_for(info["quant_info"]["quantizer_str"], use_simt)
    if cls is None:
        return nn.Linear(info["in_features"], info["out_features"], bias=False)
    return cls.gen_layer_from_info(info["linear_info"])


def _merge(infos, use_simt):
    cls = linear_class_for(infos[0]["quant_info"]["quantizer_str"], use_simt)
    merged = infos[0]["linear_info"]
    for it in infos[1:]:
        merged = cls.merge_infos(merged, it["linear_info"])
    return cls.gen_layer_from_info(merged)


def _incoherent_in(x, SU, scale):
    """fp16( U^T (x * SU) / scale )  -- one kernel"""
    n = x.shape[-1]
    return ops.hadamard(x.half().contiguous(), SU, 1.0 / (math.sqrt(n) * scale), out_dtype=torch.float16)


def _run_linear(linear, x):
    return linear(x)


# ---------------------------------------------------------------------------------------------------------------------
class StaticKVCache:
    """minimal static KV cache with the `update()` contract the reference's attention relies on
    (model/cache_utils.py:1048 StaticCache)."""

    def __init__(self, n_layers, max_seq, n_kv_heads, head_dim, batch=1, dtype=torch.float16, device=None):
        device = device or _default_device()
        self.k = [torch.zeros((batch, n_kv_heads, max_seq, head_dim), dtype=dtype, device=device) for _ in range(n_layers)]
        self.v = [torch.zeros((batch, n_kv_heads, max_seq, head_dim), dtype=dtype, device=device) for _ in range(n_layers)]
        self.seen = [0] * n_layers

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        pos = cache_kwargs["cache_position"]
        self.k[layer_idx].index_copy_(2, pos, key_states)
        self.v[layer_idx].index_copy_(2, pos, value_states)
        self.seen[layer_idx] = int(pos.max().item()) + 1
        n = self.seen[layer_idx]
        return self.k[layer_idx][:, :, :n], self.v[layer_idx][:, :, :n]


def rope_inv_freq(config, device=None):
    """inverse RoPE frequencies incl. the llama3 scaling rule (HF `_compute_llama3_parameters`)."""
    head_dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
    inv = 1.0 / (config.rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
    rs = getattr(config, "rope_scaling", None)
    if rs and rs.get("rope_type", rs.get("type")) == "llama3":
        factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
        old = rs["original_max_position_embeddings"]
        wavelen = 2 * math.pi / inv
        smooth = ((old / wavelen) - lo) / (hi - lo)
        scaled = torch.where(wavelen > old / lo, inv / factor, inv)
        mid = (wavelen <= old / lo) & (wavelen >= old / hi)
        scaled = torch.where(mid, (1 - smooth) * inv / factor + smooth * inv, scaled)
        inv = scaled
    return inv.float().to(device or _default_device())


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


# ---------------------------------------------------------------------------------------------------------------------
class IncoherentSdpaAttention(nn.Module):
    def __init__(self, config, merge_qk=False, merge_kv=False, merge_qv=False, merge_qkv=False, layer_idx=None,
                 dtype=torch.float16):
        super().__init__()
        self.config = config
        self.attention_dropout = getattr(config, "attention_dropout", 0.0)
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = getattr(config, "head_dim", None) or self.hidden_size // self.num_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.kv_out = self.hidden_size * self.num_key_value_heads // self.num_heads
        self.is_causal = True
        self.q_proj = self.k_proj = self.v_proj = self.o_proj = None
        self.qk_proj = self.qv_proj = self.kv_proj = self.qkv_proj = None
        self.dtype, self.layer_idx = dtype, layer_idx
        dev = _default_device()
        self.register_buffer("SU_qkv", torch.ones(config.hidden_size, dtype=dtype, device=dev))
        self.register_buffer("SU_o", torch.ones(config.hidden_size, dtype=dtype, device=dev))
        _, self.hidden_K = get_hadK(config.hidden_size)  # validates the size; the factor itself lives in the kernel
        self.register_buffer("Wscale_qkv", torch.ones(config.hidden_size + 2 * self.kv_out, dtype=dtype, device=dev),
                             persistent=False)
        self.register_buffer("Wscale_o", torch.ones(config.hidden_size, dtype=dtype, device=dev), persistent=False)
        self.register_buffer("inv_freq", rope_inv_freq(config, dev), persistent=False)
        self.scale = 64.0
        self.merge_qk, self.merge_kv, self.merge_qv, self.merge_qkv = merge_qk, merge_kv, merge_qv, merge_qkv
        assert sum([merge_qk, merge_kv, merge_qv, merge_qkv]) <= 1, "Only one of merge_qk, merge_kv, merge_qv, merge_qkv can be True"

    def _scaled(self, linear, x, wscale):
        return ops.scale_epilogue(ops_acc(linear, x), wscale, self.scale)

    def compute_qkv(self, input):
        n, h, kv = len(self.SU_qkv), self.hidden_size, self.kv_out
        x = _incoherent_in(input.view(-1, n), self.SU_qkv, self.scale)
        W = self.Wscale_qkv
        if self.merge_qkv:
            q, k, v = self._scaled(self.qkv_proj, x, W).split([h, kv, kv], dim=-1)
        elif self.merge_qk:
            q, k = self._scaled(self.qk_proj, x, W[:h + kv]).split([h, kv], dim=-1)
            v = self._scaled(self.v_proj, x, W[h + kv:])
        elif self.merge_kv:
            k, v = self._scaled(self.kv_proj, x, W[h:]).split([kv, kv], dim=-1)
            q = self._scaled(self.q_proj, x, W[:h])
        elif self.merge_qv:  # Wscale_qkv is stored in q, v, k order for this mode (incoherent_linear.py:211-213)
            q, v = self._scaled(self.qv_proj, x, W[:h + kv]).split([h, kv], dim=-1)
            k = self._scaled(self.k_proj, x, W[h + kv:])
        else:
            q = self._scaled(self.q_proj, x, W[:h])
            k = self._scaled(self.k_proj, x, W[h:h + kv])
            v = self._scaled(self.v_proj, x, W[h + kv:])
        lead = input.shape[:-1]
        return q.reshape(*lead, n), k.reshape(*lead, kv), v.reshape(*lead, kv)

    def compute_o(self, input):
        n = len(self.SU_o)
        x = _incoherent_in(input.view(-1, n), self.SU_o, self.scale)
        return self._scaled(self.o_proj, x, self.Wscale_o).view(*input.shape[:-1], n)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None, **kwargs):
        bsz, q_len, _ = hidden_states.size()
        q, k, v = self.compute_qkv(hidden_states)
        q = q.view(bsz, q_len, self.num_heads, self.head_dim).transpose(1, 2)
        k = k.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        v = v.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        if position_embeddings is None:
            if position_ids is None:
                position_ids = cache_position.view(1, -1) if cache_position is not None else \
                    torch.arange(q_len, device=q.device).view(1, -1)
            freqs = position_ids[..., None].float() * self.inv_freq[None, None, :]
            emb = torch.cat((freqs, freqs), dim=-1)
            cos, sin = emb.cos().to(q.dtype), emb.sin().to(q.dtype)
        else:
            cos, sin = position_embeddings
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        if past_key_value is not None:
            k, v = past_key_value.update(k, v, self.layer_idx, {"sin": sin, "cos": cos, "cache_position": cache_position})
        k = k.repeat_interleave(self.num_key_value_groups, dim=1)
        v = v.repeat_interleave(self.num_key_value_groups, dim=1)
        mask = attention_mask[:, :, :, :k.shape[-2]] if attention_mask is not None else None
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0,
                                             is_causal=(mask is None and q_len > 1))
        out = out.transpose(1, 2).contiguous().view(bsz, q_len, -1)
        return self.compute_o(out), None, past_key_value

    @staticmethod
    def gen_layer_from_info(config, layer_idx, info_q, info_k, info_v, info_o, merge_qk=False, merge_qv=False,
                            merge_kv=False, merge_qkv=False, dummy=False, use_simt=False, use_simt_q=None,
                            use_simt_k=None, use_simt_v=None, use_simt_o=None):
        attn = IncoherentSdpaAttention(config, merge_qk=merge_qk, merge_qv=merge_qv, merge_kv=merge_kv,
                                       merge_qkv=merge_qkv, layer_idx=layer_idx)
        if not dummy:
            attn.SU_qkv.data.copy_(info_q["SU"])
            attn.SU_o.data.copy_(info_o["SU"])
            order = [info_q, info_v, info_k] if merge_qv else [info_q, info_k, info_v]
            attn.Wscale_qkv.data.copy_(torch.cat([i["Wscale"] for i in order], dim=-1))
            attn.Wscale_o.data.copy_(info_o["Wscale"])
        sq = use_simt if use_simt_q is None else use_simt_q
        sk = use_simt if use_simt_k is None else use_simt_k
        sv = use_simt if use_simt_v is None else use_simt_v
        so = use_simt if use_simt_o is None else use_simt_o
        if merge_qkv:
            attn.qkv_proj = _merge([info_q, info_k, info_v], sq)
        elif merge_qk:
            attn.qk_proj = _merge([info_q, info_k], sq)
            attn.v_proj = make_linear(info_v, sv)
        elif merge_kv:
            attn.kv_proj = _merge([info_k, info_v], sk)
            attn.q_proj = make_linear(info_q, sq)
        elif merge_qv:
            attn.qv_proj = _merge([info_q, info_v], sq)
            attn.k_proj = make_linear(info_k, sk)
        else:
            attn.q_proj, attn.k_proj, attn.v_proj = make_linear(info_q, sq), make_linear(info_k, sk), make_linear(info_v, sv)
        attn.o_proj = make_linear(info_o, so)
        return attn

    @staticmethod
    def gen_layer_from_quantizer_str_and_key(config, layer_idx, quant_dir, quantizer_str_q, quantizer_str_k,
                                             quantizer_str_v, quantizer_str_o, key_q, key_k, key_v, key_o, dummy=False,
                                             **kw):
        infos = _load_infos(config, quant_dir, dummy,
                            [(quantizer_str_q, key_q, "self_attn.q_proj"), (quantizer_str_k, key_k, "self_attn.k_proj"),
                             (quantizer_str_v, key_v, "self_attn.v_proj"), (quantizer_str_o, key_o, "self_attn.o_proj")])
        return IncoherentSdpaAttention.gen_layer_from_info(config, layer_idx, *infos, dummy=dummy, **kw)


def ops_acc(linear, x):
    """raw fp32 accumulators (bs, M) of a quantized linear, without the cast back to fp16 its forward() does."""
    if isinstance(linear, QTIPLinearTCQ):
        return ops.tcq_gemv(linear.trellis, x, linear.tlut, linear.out_features, linear.in_features, linear.tlut_bits, linear.KV)
    if isinstance(linear, (CombLinearTCQ, CombtLinearTCQ)):
        return linear._gemv(x)
    if isinstance(linear, VQLinearPackTensorCore):
        return ops.lut_gemv(linear.qweight, x, linear.lut, linear.out_features, linear.in_features, linear.lut_bits, linear.vec_sz)
    if isinstance(linear, VQLinearPackSIMT):
        return ops.simt_gemv(linear.qweight, x, linear.lut, linear.out_features, linear.in_features, linear.lut_bits,
                             linear.vec_sz, out_dtype=torch.float32)
    return linear(x.to(linear.weight.dtype)).float()


def _load_infos(config, quant_dir, dummy, items):
    out = []
    for qs, key, layer_key in items:
        if not dummy:
            out.append(torch.load(f"{quant_dir}/{qs}/{key}.pt", weights_only=False))
        else:
            from ..utils.mem_op import get_dummy_quant_results
            out.append(get_dummy_quant_results(MODEL_KEYS[config._name_or_path], layer_key, qs))
    return out


# ---------------------------------------------------------------------------------------------------------------------
class IncoherentMLP(nn.Module):
    """left-only incoherent MLP with a shared SU for up/gate (reference: incoherent_linear.py:279-394)."""

    def __init__(self, hidden_size, intermediate_size, hidden_act, merge_ug=False, bias=False, dtype=torch.float16):
        super().__init__()
        assert bias is False, "bias is not supported"
        assert hidden_act == "silu", "the fused epilogue implements SiLU (Llama)"
        self.hidden_size, self.intermediate_size, self.dtype = hidden_size, intermediate_size, dtype
        self.up_proj = self.gate_proj = self.ug_proj = self.down_proj = None
        dev = _default_device()
        self.register_buffer("SU_ug", torch.ones(hidden_size, dtype=dtype, device=dev))
        self.register_buffer("SU_dp", torch.ones(intermediate_size, dtype=dtype, device=dev))
        _, self.hidden_K = get_hadK(hidden_size)
        _, self.inter_K = get_hadK(intermediate_size)
        self.register_buffer("Wscale_ug", torch.ones(intermediate_size * 2, dtype=dtype, device=dev), persistent=False)
        self.register_buffer("Wscale_dp", torch.ones(hidden_size, dtype=dtype, device=dev), persistent=False)
        self.scale = 64.0
        self.merge_ug = merge_ug

    def forward(self, input):
        n = len(self.SU_ug)
        x = self.compute_dp(self.compute_ug(input.view(-1, n).half()))
        return x.view(*input.shape[:-1], n).to(input.dtype)

    def compute_ug(self, x):
        x = _incoherent_in(x, self.SU_ug, self.scale)
        I = self.intermediate_size
        if self.merge_ug:
            return ops.scale_epilogue(ops_acc(self.ug_proj, x), self.Wscale_ug, self.scale, EPI_SILU_MUL)
        acc = torch.cat([ops_acc(self.up_proj, x), ops_acc(self.gate_proj, x)], dim=-1)
        return ops.scale_epilogue(acc, self.Wscale_ug, self.scale, EPI_SILU_MUL)

    def compute_dp(self, x):
        x = _incoherent_in(x, self.SU_dp, self.scale)
        return ops.scale_epilogue(ops_acc(self.down_proj, x), self.Wscale_dp, self.scale, EPI_NONE)

    @staticmethod
    def gen_layer_from_info(config, info_up, info_gate, info_down, merge_ug=False, dummy=False, use_simt=False,
                            use_simt_u=None, use_simt_g=None, use_simt_d=None):
        mlp = IncoherentMLP(config.hidden_size, config.intermediate_size, config.hidden_act, merge_ug=merge_ug)
        if not dummy:
            mlp.SU_ug.data.copy_(info_up["SU"])
            mlp.SU_dp.data.copy_(info_down["SU"])
            mlp.Wscale_ug.data.copy_(torch.cat([info_up["Wscale"], info_gate["Wscale"]], dim=-1))
            mlp.Wscale_dp.data.copy_(info_down["Wscale"])
        su = use_simt if use_simt_u is None else use_simt_u
        sg = use_simt if use_simt_g is None else use_simt_g
        sd = use_simt if use_simt_d is None else use_simt_d
        if merge_ug:
            mlp.ug_proj = _merge([info_up, info_gate], su)
        else:
            mlp.up_proj, mlp.gate_proj = make_linear(info_up, su), make_linear(info_gate, sg)
        mlp.down_proj = make_linear(info_down, sd)
        return mlp

    @staticmethod
    def gen_layer_from_quantizer_str_and_key(config, quant_dir, quantizer_str_up, quantizer_str_gate, quantizer_str_down,
                                             key_up, key_gate, key_down, merge_ug=False, dummy=False, **kw):
        infos = _load_infos(config, quant_dir, dummy,
                            [(quantizer_str_up, key_up, "mlp.up_proj"), (quantizer_str_gate, key_gate, "mlp.gate_proj"),
                             (quantizer_str_down, key_down, "mlp.down_proj")])
        return IncoherentMLP.gen_layer_from_info(config, *infos, merge_ug=merge_ug, dummy=dummy, **kw)


# ---------------------------------------------------------------------------------------------------------------------
class IncoherentLinear(nn.Module):
    def __init__(self, in_features, out_features, hadU, hadV, bias=False, dtype=torch.float16, use_linear=True):
        super().__init__()
        self.in_features, self.out_features, self.dtype = in_features, out_features, dtype
        dev = _default_device()
        self.linear = nn.Linear(in_features, out_features, bias=False, dtype=dtype, device=dev) if use_linear else None
        if bias:
            self.register_buffer("bias", torch.ones(out_features, device=dev))
        else:
            self.bias = None
        self.register_buffer("SU", torch.ones(in_features, dtype=dtype, device=dev))
        self.register_buffer("SV", torch.ones(out_features, dtype=dtype, device=dev))
        self.hadU, self.hadV = hadU, hadV
        _, self.K_left = get_hadK(hadU)
        _, self.K_right = get_hadK(hadV)
        self.register_buffer("Wscale", torch.ones(out_features, dtype=dtype, device=dev), persistent=False)
        self.scale = 32.0
        self.rot_info = "all"
        self.skip_l = self.skip_r = False

    def apply_rot_info(self):
        table = {"all": (False, False), "skip_l": (True, False), "skip_r": (False, True), "skip_lr": (True, True)}
        if self.rot_info not in table:
            raise ValueError(f"Invalid rot_info: {self.rot_info}")
        self.skip_l, self.skip_r = table[self.rot_info]

    def save_info(self, path, quant_info=None):
        info = {"in_features": self.in_features, "out_features": self.out_features, "hadU": self.hadU, "hadV": self.hadV,
                "dtype": self.dtype, "scale": self.scale, "Wscale": self.Wscale.detach().cpu(), "rot_info": self.rot_info,
                "linear_info": self.linear._info(), "bias": self.bias.detach().cpu() if self.bias is not None else None,
                "SU": self.SU.detach().cpu(), "SV": self.SV.detach().cpu(), "quant_info": quant_info}
        torch.save(info, path)

    def forward(self, input):
        n, m = len(self.SU), len(self.SV)
        x = input.view(-1, n).half()
        if not self.skip_l:
            if self.hadU == n:
                x = ops.hadamard(x.contiguous(), self.SU, 1.0 / (math.sqrt(n) * self.scale), out_dtype=torch.float16)
            else:  # block-diagonal transform over heads of size hadU
                x = matmul_hadU_head_cuda(x * self.SU, None, self.K_left, self.hadU) / self.scale
        else:
            x = x / self.scale
        if self.skip_r:
            x = ops.scale_epilogue(ops_acc(self.linear, x.half()), self.Wscale, self.scale)
        else:
            x = self.linear(x.half()) * self.Wscale
            x = matmul_hadU_head_cuda(x, None, self.K_right, self.hadV)
            x = x * (self.SV * self.scale)
        x = x.view(*input.shape[:-1], m).to(input.dtype)
        if self.bias is not None:
            x = x + self.bias
        return x

    @staticmethod
    def gen_layer_from_info(info, merge_layers=False, dummy=False, use_simt=False):
        layer = IncoherentLinear(info["in_features"], info["out_features"], info.get("hadU", info["in_features"]),
                                 info.get("hadV", info["out_features"]), bias=info["bias"] is not None,
                                 dtype=info["dtype"], use_linear=False)
        if not dummy:
            if info["bias"] is not None:
                layer.bias.data.copy_(info["bias"])
            layer.SU.data.copy_(info["SU"])
            layer.SV.data.copy_(info["SV"])
            layer.Wscale.data.copy_(info["Wscale"])
        if info["quant_info"] is not None:
            layer.linear = make_linear(info, use_simt)
        if info["quant_info"] is not None and "rot_info" in info["quant_info"]:
            layer.rot_info = info["quant_info"]["rot_info"]
        elif "rot_info" in info:
            layer.rot_info = info["rot_info"]
        else:
            layer.rot_info = "all"
        if layer.rot_info is None:
            layer.rot_info = "all"
        if merge_layers:
            layer.apply_rot_info()
        return layer

    @staticmethod
    def gen_layer_from_quantizer_str_and_key(config, quant_dir, quantizer_str, key, merge_layers=False, dummy=False,
                                             use_simt=False):
        layer_id = key.split("_")[0]
        layer_key = key.replace(f"{layer_id}_", "")
        (info,) = _load_infos(config, quant_dir, dummy, [(quantizer_str, key, layer_key)])
        return IncoherentLinear.gen_layer_from_info(info, merge_layers=merge_layers, dummy=dummy, use_simt=use_simt)
