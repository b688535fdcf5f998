"""ctypes binding of libqpalette.so (include/qpalette.h).  There is no fallback: if the library is missing or a call
fails this raises, so a GPU test can never silently pass on a CPU / eager path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, f"libqpalette{os.environ.get('QP_LIB_SUFFIX', '')}.so")

QP_OK = 0
SPLIT_NONE, SPLIT_IN, SPLIT_OUT = 0, 1, 2
FLAG_ACCUMULATE = 1
EPI_NONE, EPI_SILU_MUL = 0, 1

_lib = None

_vp, _i, _u, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_float

# name -> argtypes  (restype is int unless listed in _RESTYPES); must mirror include/qpalette.h
SIGNATURES = {
    "qp_version": [],
    "qp_last_error": [],
    "qp_launch_count": [],
    "qp_device_sm_count": [],
    "qp_tcq_gemv": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _u, _vp],
    "qp_tcq_dequant": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "qp_tcq_gemm_tc": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _u, _vp],
    "qp_gemm_mma_scratch_bytes": [_i, _i],
    "qp_tcq_gemm_mma": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _u, _vp],
    "qp_lut_gemm_tc": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u, _vp],
    "qp_lut_gemm_mma": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u, _vp],
    "qp_lut_gemv": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u, _vp],
    "qp_lut_dequant": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "qp_simt_gemv": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "qp_simt_dequant": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "qp_convert_tc_to_simt": [_vp, _vp, _i, _i, _i, _i, _vp],
    "qp_hadamard": [_vp, _vp, _vp, _i, _i, _f, _i, _i, _vp],
    "qp_scale_epilogue": [_vp, _vp, _vp, _i, _i, _f, _i, _vp],
    "qp_fused_norm_had": [_vp, _vp, _i, _vp, _vp, _f, _vp, _f, _vp, _i, _f, _i, _vp, _i, _vp],
    "qp_fused_norm_had_xchg": [_vp, _vp, _i, _vp, _vp, _f, _vp, _f, _vp, _i, _f, _i, _vp, _i, _vp, _vp],
    "qp_xchg_send_ll": [_vp, _i, _i, _vp, _i, _vp, _vp],
    "qp_silu_mul_had_grid_xchg": [_vp, _vp, _vp, _f, _vp, _i, _f, _vp, _i, _vp, _vp, _vp],
    "qp_set_spin_timeout_ms": [ctypes.c_longlong],
    "qp_peer_alloc": [_vp, ctypes.c_size_t],
    "qp_peer_free": [_vp],
    "qp_peer_export": [_vp, _vp],
    "qp_peer_import": [_vp, _vp],
    "qp_peer_close": [_vp],
    "qp_silu_mul_had": [_vp, _vp, _vp, _f, _vp, _i, _f, _vp, _i, _vp],
    "qp_silu_mul_had_cluster": [_vp, _vp, _vp, _f, _vp, _i, _f, _vp, _i, _vp],
    "qp_silu_mul_had_grid": [_vp, _vp, _vp, _f, _vp, _i, _f, _vp, _i, _vp, _vp],
    "qp_rope_attention": [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "qp_rope_attention_scratch_bytes": [_i, _i, _i],
    "qp_gemv_f16": [_vp, _vp, _vp, _i, _i, _vp],
    "qp_argmax": [_vp, _vp, _i, _vp, _vp],
    "qp_embed": [_vp, _vp, _vp, _i, _vp],
    "qp_step_advance": [_vp, _vp, _vp, _i, _vp],
    "qp_tcq_gemv_fused": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "qp_lut_gemv_fused": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "qp_tcq_gemv_host": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
}
_RESTYPES = {"qp_last_error": ctypes.c_char_p, "qp_launch_count": ctypes.c_uint64, "qp_gemm_mma_scratch_bytes": ctypes.c_size_t,
             "qp_rope_attention_scratch_bytes": ctypes.c_size_t}


class XProd(ctypes.Structure):
    """mirror of `qp_xprod` (include/qpalette.h)"""
    _fields_ = [("src_f16", _vp), ("h_out_f16", _vp), ("acc", _vp), ("wscale_f16", _vp), ("acc_scale", _f),
                ("norm_w_f16", _vp), ("eps", _f), ("su_f16", _vp), ("had_scale", _f), ("x_out_f16", _vp),
                ("zero1", _vp), ("zero1_count", _i), ("zero2", _vp), ("zero2_count", _i),
                ("ll", _vp), ("ll_epoch", _vp), ("ll_kind", _i)]


class Xchg(ctypes.Structure):
    """mirror of `qp_xchg` (include/qpalette.h)"""
    _fields_ = [("peer_base", _vp), ("peer_flags", _vp), ("epoch", _vp), ("offset", ctypes.c_longlong),
                ("slice_bytes", _i), ("rank", _i), ("nranks", _i), ("site", _i)]


class QPaletteError(RuntimeError):
    pass


def lib():
    """load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QPaletteError(
                f"{LIB_PATH} not found: build it with `python q-palette_b200/build.py` (there is no CPU fallback)")
        h = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is missing
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = h
    return _lib


def check(rc):
    if rc != QP_OK:
        msg = lib().qp_last_error()
        raise QPaletteError(f"libqpalette error {rc}: {msg.decode() if msg else '?'}")


def launch_count():
    return int(lib().qp_launch_count())
