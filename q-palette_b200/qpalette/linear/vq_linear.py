"""VQLinearPackTensorCore / VQLinearPackSIMT -- reference API of lib/linear/vq_linear.py:5-208 (buffers `qweight` int32
(M, lut_bits*K/32/vec_sz) and `lut` (2^lut_bits, vec_sz))."""
import torch
import torch.nn as nn

from .. import ops
from .tcq_linear import _default_device


class _VQBase(nn.Module):
    def __init__(self, in_features, out_features, lut_bits, vec_sz, bias=False, dtype=torch.half):
        super().__init__()
        self.in_features, self.out_features, self.lut_bits, self.dtype, self.vec_sz = in_features, out_features, lut_bits, dtype, vec_sz
        dev = _default_device()
        self.register_buffer("qweight", torch.randint(0, 4, (out_features, lut_bits * in_features // 32 // vec_sz),
                                                      dtype=torch.int32, device=dev))
        self.register_buffer("lut", torch.randn((2 ** lut_bits, vec_sz), dtype=self.dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.randn((out_features,), dtype=self.dtype, device=dev))
        else:
            self.bias = None

    def _info(self):
        return {"in_features": self.in_features, "out_features": self.out_features, "lut_bits": self.lut_bits,
                "dtype": self.dtype, "vec_sz": self.vec_sz, "qweight": self.qweight.detach().cpu(),
                "lut": self.lut.detach().cpu().half(),
                "bias": self.bias.detach().cpu() if self.bias is not None else None}

    @staticmethod
    def merge_infos(info1, info2):
        for key in ("in_features", "lut_bits", "vec_sz", "dtype"):
            assert info1[key] == info2[key], key
        assert info1["bias"] is None and info2["bias"] is None
        if not torch.allclose(info1["lut"].float().cpu(), info2["lut"].float().cpu(), atol=1e-4):
            print("warning: lut is not close. it is unexpected behavior if you do not use dummy quantizers.")
        info = {k: info1[k] for k in ("in_features", "lut_bits", "vec_sz", "dtype", "lut")}
        info["out_features"] = info1["out_features"] + info2["out_features"]
        info["bias"] = None
        info["qweight"] = torch.cat([info1["qweight"], info2["qweight"]], dim=0)
        return info


class VQLinearPackTensorCore(_VQBase):
    def __init__(self, in_features, out_features, lut_bits, vec_sz=2, bias=False, dtype=torch.half):
        super().__init__(in_features, out_features, lut_bits, vec_sz, bias, dtype)
        self.vq_type = f"vq{self.vec_sz}" if self.vec_sz > 1 else "sq_dup" if lut_bits <= 4 else "sq"

    def forward(self, inp, **kwargs):
        x = inp.view(-1, self.in_features)
        bs, m, k = x.shape[0], self.out_features, self.in_features
        if bs <= 8:
            x = ops.resolve(f"decompress_gemm_{m}_{bs}_{k}_{self.lut_bits}_{self.vq_type}")(self.qweight, x, self.lut)
        elif bs <= ops.MMA_GEMM_MAX_BS:  # fused dequant + GEMM on the GEMV's decode loop, every format
            x = ops.lut_gemm_mma(self.qweight, x, self.lut, m, k, self.lut_bits, self.vec_sz)
        elif ops.tc_gemm_supported(m, k) and (self.vec_sz == 2 or self.lut_bits <= 5):  # ... on tcgen05
            x = ops.lut_gemm_tc(self.qweight, x, self.lut, m, k, self.lut_bits, self.vec_sz)
        else:
            x = ops.batched_matmul(x, lambda: ops.resolve(f"decompress_{self.lut_bits}_{self.vq_type}")(
                self.qweight, self.lut, m, k))
        return x.view(*inp.shape[:-1], m).to(inp.dtype)

    @staticmethod
    def gen_layer_from_info(info):
        layer = VQLinearPackTensorCore(info["in_features"], info["out_features"], info["lut_bits"], info["vec_sz"],
                                       info["bias"] is not None, info["dtype"])
        layer.qweight.data.copy_(info["qweight"])
        layer.lut.data.copy_(info["lut"])
        if info["bias"] is not None:
            layer.bias.data.copy_(info["bias"])
        return layer


class VQLinearPackSIMT(_VQBase):
    def __init__(self, in_features, out_features, lut_bits, vec_sz=1, bias=False, dtype=torch.half):
        super().__init__(in_features, out_features, lut_bits, vec_sz, bias, dtype)

    def forward(self, inp, **kwargs):
        x = inp.view(-1, 1, self.in_features)
        bs, m, k = x.shape[0], self.out_features, self.in_features
        if bs <= 8:
            if self.vec_sz == 1:
                x = ops.resolve("sq_pack_gemm_simt")(x, self.qweight, self.lut, self.lut_bits)
            else:
                x = ops.resolve(f"vq_pack_gemm_simt_{bs}_{self.vec_sz}_{self.lut_bits}")(x, self.qweight, self.lut)
        else:
            if self.vec_sz == 1:
                get = lambda: ops.resolve("sq_pack_dequant_simt")(self.qweight, self.lut, self.lut_bits, m, k)
            else:
                get = lambda: ops.resolve(f"vq_pack_dequant_simt_{self.vec_sz}_{self.lut_bits}")(self.qweight, self.lut, m, k)
            x = ops.batched_matmul(x.view(-1, k), get)
        return x.view(*inp.shape[:-1], m).to(inp.dtype)

    @staticmethod
    def gen_layer_from_info(info):
        """`info["qweight"]` is in the tensor-core layout (as quantize_layer.py stores it) and is converted to the SIMT
        layout here, on the GPU (the reference does this with numba on the CPU: vq_linear.py:177-182)."""
        layer = VQLinearPackSIMT(info["in_features"], info["out_features"], info["lut_bits"], info["vec_sz"],
                                 info["bias"] is not None, info["dtype"])
        if info["vec_sz"] <= 2:
            qw = info["qweight"].to(layer.qweight.device).contiguous()
            layer.qweight.data.copy_(ops.convert_tc_to_simt(qw, info["out_features"], info["in_features"],
                                                            info["lut_bits"], info["vec_sz"]))
        else:
            layer.qweight.data.copy_(info["qweight"])
        layer.lut.data.copy_(info["lut"])
        if info["bias"] is not None:
            layer.bias.data.copy_(info["bias"])
        return layer
