"""The reference's UNMODIFIED module files (lib/linear/{tcq,comb,vq}_linear.py) on top of this repo's operator binding.

Those files only `import torch` and reach the kernels through `getattr(torch.ops.ours_lib, "<shape-templated name>")`
(tcq_linear.py:68-72, comb_linear.py:80-127,223-270, vq_linear.py:48-68,139-172).  `import qpalette` hooks the `ours_lib`
namespace so that such a lookup registers the op on first use (qpalette/ops.py: install_namespace_hook) -- no call to
`ops.resolve` is needed, which is what makes libqpalette.so a drop-in under the reference's own classes.

The reference sources are read from /root/reference where they lie (never copied into the repo): these tests run in the
build container and skip on the GPU box, which has no reference checkout.  Without a GPU the forward passes run on the
`meta` device, i.e. through the ops' registered fake implementations -- name resolution, schemas and output shapes are
checked; with a GPU (and a reference checkout) outputs are compared bit-for-bit with this repo's own module classes.
The hook itself is exercised on the GPU box by tests/test_gpu_kernels.py::test_reference_call_pattern_without_resolve.
"""
import importlib.util
import os

import pytest
import torch

REF = "/root/reference/lib/linear"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")

CASES = [  # quantizer string, reference file, class name, use_simt
    ("tcq_6_none_0.9", "tcq_linear", "QTIPLinearTCQ"),
    ("tcq_9_none_0.9", "tcq_linear", "QTIPLinearTCQ"),
    ("tcomb_6_7_0.5_none_0.9", "comb_linear", "CombtLinearTCQ"),
    ("comb_7_8_0.5_none_0.9", "comb_linear", "CombLinearTCQ"),
    ("ldlq_2_8_none_1.0", "vq_linear", "VQLinearPackTensorCore"),
    ("ldlq_1_4_none_1.0", "vq_linear", "VQLinearPackTensorCore"),
    ("ldlq_2_6_none_1.0", "vq_linear", "VQLinearPackSIMT"),
    ("ldlq_1_6_none_1.0", "vq_linear", "VQLinearPackSIMT"),
]


class _TorchOnCpu:
    """stands in for the `torch` global of a reference module in the GPU-less container: vq_linear.py allocates its
    buffers with a hard-coded device='cuda' (vq_linear.py:17-28); everything else is forwarded untouched"""

    def __getattr__(self, k):
        v = getattr(torch, k)
        if k in ("randint", "randn", "zeros", "ones", "empty"):
            def on_cpu(*a, **kw):
                if str(kw.get("device", "")).startswith("cuda"):
                    kw["device"] = "cpu"
                return v(*a, **kw)
            return on_cpu
        return v


def load_ref(name, cpu_only=False):
    spec = importlib.util.spec_from_file_location(f"qp_reference_{name}", os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if cpu_only:
        mod.torch = _TorchOnCpu()
    return mod


def make_info(qs, K, M, device):
    from qpalette.utils import get_dummy_quant_results
    return get_dummy_quant_results(None, None, qs, in_features=K, out_features=M, device=device)["linear_info"]


def stub_quant_op():
    """VQLinearPackSIMT.gen_layer_from_info converts the tensor-core layout with `lib.quantizer.quant_op.
    convert_tensor_core_to_simt` (vq_linear.py:178; numba on the CPU, imports glog).  That converter is not on the path under
    test: stand in the GPU conversion of this repo (identity on the meta device, where only shapes matter)."""
    import sys
    import types
    if "lib.quantizer.quant_op" in sys.modules:
        return
    def convert_tensor_core_to_simt(mat_packed, N, K, vec_sz, lut_bit, code_n, codeT_sz=32, td_x=16, td_y=16):
        if mat_packed.is_cuda:
            from qpalette import ops
            return ops.convert_tc_to_simt(mat_packed, N, K, lut_bit, vec_sz)
        return mat_packed
    for name in ("lib", "lib.quantizer", "lib.quantizer.quant_op"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lib.quantizer.quant_op"].convert_tensor_core_to_simt = convert_tensor_core_to_simt


@needs_ref
@pytest.mark.parametrize("qs,fname,cname", CASES)
@pytest.mark.parametrize("bs", [1, 8, 12])
def test_reference_modules_resolve_ops_on_meta(qs, fname, cname, bs):
    import qpalette  # noqa: F401  (installs the namespace hook)
    K, M = 512, 256
    simt = cname.endswith("SIMT")
    stub_quant_op()
    info = make_info(qs, K, M, "cpu")
    cls = getattr(load_ref(fname, cpu_only=not torch.cuda.is_available()), cname)
    layer = cls.gen_layer_from_info(info).to("meta")
    x = torch.empty((bs, 1, K) if simt else (bs, K), dtype=torch.float16, device="meta")
    y = layer(x)  # getattr(torch.ops.ours_lib, <name>) inside: no ops.resolve() has been called for this name
    assert y.shape[-1] == M and y.shape[0] == bs and y.dtype == torch.float16 and y.device.type == "meta"


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("qs,fname,cname", CASES)
@pytest.mark.parametrize("bs", [1, 5, 12])
def test_reference_modules_match_repo_modules_on_gpu(qs, fname, cname, bs):
    import qpalette.linear as L
    torch.manual_seed(0)
    K, M = 1024, 512
    simt = cname.endswith("SIMT")
    stub_quant_op()
    info = make_info(qs, K, M, "cuda")
    ref_layer = getattr(load_ref(fname), cname).gen_layer_from_info(info).cuda()
    own_layer = getattr(L, cname).gen_layer_from_info(info).cuda()
    x = torch.randn((bs, 1, K) if simt else (bs, K), device="cuda").half()
    a, b = ref_layer(x), own_layer(x)
    assert a.shape == b.shape and a.dtype == b.dtype
    # same kernels underneath; bs <= 8 sums with fp32 atomics (order varies run to run), bs > 8 takes different GEMM paths
    assert torch.allclose(a.float(), b.float(), rtol=2e-3, atol=2e-3)
