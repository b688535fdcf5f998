"""QTIPLinearTCQ -- same constructor, buffers, `_info()` schema, `gen_layer_from_info`, `merge_infos` and op names as the
reference (lib/linear/tcq_linear.py:5-122); the ops resolve onto libqpalette's shape-generic sm_100a kernels."""
import math

import torch
import torch.nn as nn

from .. import ops


def _default_device():
    return "cuda" if torch.cuda.is_available() else "cpu"


class QTIPLinearTCQ(nn.Module):
    def __init__(self, in_features, out_features, td_x, td_y, L, KV, V, tlut_bits, bias=False, dtype=torch.float16):
        super().__init__()
        assert td_x == 16 and td_y == 16 and L == 16 and V == 2, "the packed layout is defined for 16x16 tiles, L=16, V=2"
        self.in_features, self.out_features = in_features, out_features
        self.td_x, self.td_y, self.L, self.KV, self.V, self.tlut_bits, self.dtype = td_x, td_y, L, KV, V, tlut_bits, dtype
        self.register_buffer("trellis", torch.zeros((out_features // td_x) * (in_features // td_y),
                                                    math.ceil((td_x * td_y) * KV / 16 / V), dtype=torch.int16))
        self.tlut = nn.Parameter(torch.zeros(2 ** tlut_bits, V, dtype=torch.float16), requires_grad=False)
        if bias:
            self.register_buffer("bias", torch.ones(out_features))
        else:
            self.bias = None

    def _info(self):
        return {"in_features": self.in_features, "out_features": self.out_features, "td_x": self.td_x, "td_y": self.td_y,
                "L": self.L, "KV": self.KV, "V": self.V, "tlut_bits": self.tlut_bits, "dtype": self.dtype,
                "trellis": self.trellis.detach().cpu(), "tlut": self.tlut.detach().cpu().half(),
                "bias": self.bias.detach().cpu() if self.bias is not None else None}

    def forward(self, inp, **kwargs):
        x = inp.view(-1, self.in_features)
        bs, m, k = x.shape[0], self.out_features, self.in_features
        if bs <= 8:
            op = ops.resolve(f"decompress_gemm_tcq_{m}_{bs}_{k}_{self.tlut_bits}_{self.KV}")
            x = op(self.trellis, x, self.tlut)
        elif ops.tc_gemm_supported(m, k) or (bs <= ops.MMA_GEMM_MAX_BS and ops._tcq_mma_supported(self.tlut_bits, self.KV, 0)):
            # fused dequant + GEMM: mma.sync on the GEMV loop up to bs = 32, tcgen05 above (the reference: dequantise + cuBLAS)
            x = ops.tcq_gemm_tc(self.trellis, x, self.tlut, m, k, self.tlut_bits, self.KV)
        else:
            x = ops.batched_matmul(x, lambda: ops.resolve(f"decompress_tcq_{self.tlut_bits}_{self.KV}")(
                self.trellis, self.tlut, m, k))
        return x.view(*inp.shape[:-1], m).to(inp.dtype)

    @staticmethod
    def gen_layer_from_info(info):
        layer = QTIPLinearTCQ(info["in_features"], info["out_features"], info["td_x"], info["td_y"], info["L"],
                              info["KV"], info["V"], info["tlut_bits"], info["bias"] is not None, info["dtype"])
        layer = layer.to(info["trellis"].device)
        layer.trellis.data.copy_(info["trellis"])
        layer.tlut.data.copy_(info["tlut"])
        if info["bias"] is not None:
            layer.bias.data.copy_(info["bias"])
        return layer

    @staticmethod
    def merge_infos(info1, info2):
        """stack two layers along the output rows: the packed layout is strip-major, so this is a plain cat(dim=0)."""
        for key in ("in_features", "td_x", "td_y", "L", "KV", "V", "tlut_bits", "dtype"):
            assert info1[key] == info2[key], key
        assert info1["bias"] is None and info2["bias"] is None
        if not torch.allclose(info1["tlut"].float().cpu(), info2["tlut"].float().cpu(), atol=1e-4):
            print("warning: tlut is not close. it is unexpected behavior if you do not use dummy quantizers.")
        info = {k: info1[k] for k in ("in_features", "td_x", "td_y", "L", "KV", "V", "tlut_bits", "dtype", "tlut")}
        info["out_features"] = info1["out_features"] + info2["out_features"]
        info["bias"] = None
        info["trellis"] = torch.cat([info1["trellis"], info2["trellis"]], dim=0)
        return info
