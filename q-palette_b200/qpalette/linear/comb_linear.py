"""CombLinearTCQ (two rates split on OUTPUT rows) and CombtLinearTCQ (split on INPUT columns: the `tcomb_*` TCQ-x.25/x.75
quantizers) -- reference API of lib/linear/comb_linear.py:5-320.  Unlike the reference, unequal parts also run as ONE fused
launch (the C ABI takes the split boundary), so `use_comb_kernel` is kept only as an attribute."""
import math

import torch
import torch.nn as nn

from .. import ops
from .._cabi import SPLIT_IN, SPLIT_OUT
from .tcq_linear import QTIPLinearTCQ


class _CombBase(nn.Module):
    _split = None  # SPLIT_IN / SPLIT_OUT
    _part_key = None

    def __init__(self, in_features, out_features, td_x, td_y, part, L, KV, V, tlut_bits, bias=False, dtype=torch.float16):
        super().__init__()
        assert len(part) == 2 and len(KV) == 2
        assert td_x == 16 and td_y == 16 and L == 16 and V == 2
        self.in_features, self.out_features = in_features, out_features
        self.td_x, self.td_y, self.L, self.KV, self.V, self.tlut_bits, self.dtype = td_x, td_y, L, tuple(KV), V, tlut_bits, dtype
        setattr(self, self._part_key, tuple(part))
        if self._split == SPLIT_OUT:
            assert part[0] + part[1] == out_features
            shapes = [(part[i] // td_x) * (in_features // td_y) for i in range(2)]
        else:
            assert part[0] + part[1] == in_features
            shapes = [(out_features // td_x) * (part[i] // td_y) for i in range(2)]
        for i in range(2):
            self.register_buffer(f"trellis{i + 1}",
                                 torch.zeros(shapes[i], math.ceil((td_x * td_y) * KV[i] / 16 / V), dtype=torch.int16))
        self.tlut = nn.Parameter(torch.zeros(2 ** tlut_bits, V, dtype=torch.float16), requires_grad=False)
        if bias:
            self.register_buffer("bias", torch.ones(out_features))
        else:
            self.bias = None
        self.use_comb_kernel = part[0] == part[1]

    @property
    def _part(self):
        return getattr(self, self._part_key)

    def _info(self):
        return {"in_features": self.in_features, "out_features": self.out_features, "td_x": self.td_x, "td_y": self.td_y,
                self._part_key: self._part, "L": self.L, "KV": self.KV, "V": self.V, "tlut_bits": self.tlut_bits,
                "dtype": self.dtype, "trellis1": self.trellis1.detach().cpu(), "trellis2": self.trellis2.detach().cpu(),
                "tlut": self.tlut.detach().cpu().half(),
                "bias": self.bias.detach().cpu() if self.bias is not None else None}

    def _gemv(self, x):
        m, k = self.out_features, self.in_features
        mode = "combt" if self._split == SPLIT_IN else "comb"
        if self.use_comb_kernel and self.KV[1] == self.KV[0] + 1:
            op = ops.resolve(f"decompress_gemm_tcq_{mode}_{m}_{x.shape[0]}_{k}_{self.tlut_bits}_{self.KV[0]}_{self.KV[1]}")
            return op(self.trellis1, self.trellis2, x, self.tlut)
        return ops.tcq_gemv(self.trellis1, x, self.tlut, m, k, self.tlut_bits, self.KV[0], self.trellis2, self.KV[1],
                            self._split, self._part[0])

    def get_weight(self):
        return ops.tcq_dequant(self.trellis1, self.tlut, self.out_features, self.in_features, self.tlut_bits, self.KV[0],
                               self.trellis2, self.KV[1], self._split, self._part[0])

    def forward(self, inp, **kwargs):
        x = inp.view(-1, self.in_features)
        if x.shape[0] <= 8:
            x = self._gemv(x)
        elif ops.tc_gemm_supported(self.out_features, self.in_features) and self._part[0] % 128 == 0 and \
                (self._split == SPLIT_OUT or self.KV[1] == self.KV[0] + 1):
            x = ops.tcq_gemm_tc(self.trellis1, x, self.tlut, self.out_features, self.in_features, self.tlut_bits,
                                self.KV[0], self.trellis2, self.KV[1], self._split, self._part[0])
        else:
            x = ops.batched_matmul(x, self.get_weight)
        return x.view(*inp.shape[:-1], self.out_features).to(inp.dtype)

    @classmethod
    def gen_layer_from_info(cls, info):
        layer = cls(info["in_features"], info["out_features"], info["td_x"], info["td_y"], info[cls._part_key], info["L"],
                    info["KV"], info["V"], info["tlut_bits"], info["bias"] is not None, info["dtype"])
        layer = layer.to(info["trellis1"].device)
        layer.trellis1.data.copy_(info["trellis1"])
        layer.trellis2.data.copy_(info["trellis2"])
        layer.tlut.data.copy_(info["tlut"])
        if info["bias"] is not None:
            layer.bias.data.copy_(info["bias"])
        return layer


class CombLinearTCQ(_CombBase):
    _split = SPLIT_OUT
    _part_key = "out_part"

    def __init__(self, in_features, out_features, td_x, td_y, out_part, L, KV, V, tlut_bits, bias=False, dtype=torch.float16):
        super().__init__(in_features, out_features, td_x, td_y, out_part, L, KV, V, tlut_bits, bias, dtype)

    @staticmethod
    def merge_infos(info1, info2):
        raise NotImplementedError("merging two output-split layers would interleave their parts (the reference's "
                                  "CombLinearTCQ.merge_infos has the same restriction for the fused kernel)")


class CombtLinearTCQ(_CombBase):
    _split = SPLIT_IN
    _part_key = "in_part"

    def __init__(self, in_features, out_features, td_x, td_y, in_part, L, KV, V, tlut_bits, bias=False, dtype=torch.float16):
        super().__init__(in_features, out_features, td_x, td_y, in_part, L, KV, V, tlut_bits, bias, dtype)

    @staticmethod
    def merge_infos(info1, info2):
        for key in ("in_features", "td_x", "td_y", "L", "V", "tlut_bits", "dtype"):
            assert info1[key] == info2[key], key
        assert tuple(info1["KV"]) == tuple(info2["KV"]) and tuple(info1["in_part"]) == tuple(info2["in_part"])
        assert info1["bias"] is None and info2["bias"] is None
        if not torch.allclose(info1["tlut"].float().cpu(), info2["tlut"].float().cpu(), atol=1e-4):
            print("warning: tlut is not close. it is unexpected behavior if you do not use dummy quantizers.")
        info = {k: info1[k] for k in ("in_features", "td_x", "td_y", "L", "KV", "V", "tlut_bits", "dtype", "tlut", "in_part")}
        info["out_features"] = info1["out_features"] + info2["out_features"]
        info["bias"] = None
        info["trellis1"] = torch.cat([info1["trellis1"], info2["trellis1"]], dim=0)
        info["trellis2"] = torch.cat([info1["trellis2"], info2["trellis2"]], dim=0)
        return info
