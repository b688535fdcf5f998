"""GPU parity at BASELINE.json's own shapes (VERDICT r01 "weak 2"): the Llama-3.1-8B projections quantized with
`tcomb_6_7` (TCQ-3.25) and `ldlq_2_8`, checked against the oracle -- decoded weights BIT-EXACT over the whole matrix,
GEMV outputs rel-L2 <= 1e-3 (north_star tolerance) -- plus the fused-prologue GEMV at K = 4096 / 8192.

Full-size decoded weights come from the multi-threaded C port of the oracle (oracle/qp_cref.c, itself pinned to the numpy
oracle by tests/test_oracle_cref.py); a sample of 32-row strips is additionally decoded with the numpy oracle directly, so
the comparison does not rest on the C port alone."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL_L2_TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from qpalette import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def cref():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libqp_cref.so"], check=True, capture_output=True)
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libqp_cref.so"))


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def cref_combt(cref, b1, b2, tlut, M, K, KV1, KV2, S):
    """fp16 W (M, K) of a tcomb layer decoded on the host by the C port"""
    W = np.zeros((M, K), np.float16)
    assert cref.qp_cref_tcq(vp(b1), vp(tlut), M, K // 2, KV1, S, None, 0, K, 0, 0, M, None, vp(W), K) == 0
    assert cref.qp_cref_tcq(vp(b2), vp(tlut), M, K // 2, KV2, S, None, 0, K, K // 2, 0, M, None, vp(W), K) == 0
    return W


def gemv64(W, x):
    """x (bs, K) @ W.T in float64 without materialising a float64 copy of a 58 M-element matrix per call"""
    out = np.zeros((x.shape[0], W.shape[0]), np.float64)
    x64 = x.astype(np.float64)
    for r0 in range(0, W.shape[0], 512):
        out[:, r0:r0 + 512] = x64 @ W[r0:r0 + 512].astype(np.float64).T
    return out


@pytest.mark.parametrize("M,K", [(4096, 4096), (4096, 14336), (6144, 4096), (28672, 4096)])
def test_tcomb_6_7_llama8b_shapes(ops, cref, M, K):
    from qpalette._cabi import SPLIT_IN
    KV1, KV2, S = 6, 7, 9
    rng = np.random.default_rng(M + K)
    b1 = rng.integers(0, 256, size=M * (K // 2) * KV1 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K // 2) * KV2 // 16, dtype=np.uint8)
    tlut = (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)
    d1, d2, dt = cuda(b1), cuda(b2), cuda(tlut)
    W = ops.tcq_dequant(d1, dt, M, K, S, KV1, d2, KV2, SPLIT_IN, K // 2).cpu().numpy()
    Wref = cref_combt(cref, b1, b2, tlut, M, K, KV1, KV2, S)
    assert np.array_equal(W.view(np.uint16), Wref.view(np.uint16)), "decoded weights are not bit-exact"
    # numpy oracle, straight from the format definition, on a sample of strips (a strip = contiguous bytes of each part)
    for mh in (0, M // 64 + 1, M // 32 - 1):
        for buf, KV, c0 in ((b1, KV1, 0), (b2, KV2, K // 2)):
            sb = (K // 2 // 32) * 64 * KV
            Ws = O.tcq_decode(buf[mh * sb:(mh + 1) * sb], tlut, 32, K // 2, KV, S)
            assert np.array_equal(W[32 * mh:32 * mh + 32, c0:c0 + K // 2].view(np.uint16), Ws.view(np.uint16)), (mh, KV)
    for bs in (1, 8):
        x = rng.standard_normal((bs, K)).astype(np.float16)
        out = ops.tcq_gemv(d1, cuda(x), dt, M, K, S, KV1, d2, KV2, SPLIT_IN, K // 2).cpu().numpy()
        assert rel_l2(out, gemv64(Wref, x)) <= REL_L2_TOL, bs


@pytest.mark.parametrize("M,K,vec,R", [(14336, 4096, 2, 8), (4096, 14336, 2, 6), (4096, 14336, 1, 4)])
def test_lut_llama8b_shapes(ops, M, K, vec, R):
    rng = np.random.default_rng(R + M)
    lut = rng.standard_normal((1 << R, vec)).astype(np.float16)
    buf = rng.integers(0, 256, size=M * K * R // 8 // vec, dtype=np.uint8)
    d, dl = cuda(buf), cuda(lut)
    W = ops.lut_dequant(d, dl, M, K, R, vec).cpu().numpy()
    rows_per = 1024
    for r0 in range(0, M, rows_per):  # strips are contiguous in the packed buffer: decode the oracle piecewise
        nb = rows_per * K * R // 8 // vec
        Ws = O.lut_tc_decode(buf[(r0 // rows_per) * nb:(r0 // rows_per + 1) * nb].view(np.int32), lut, rows_per, K, R, vec)
        assert np.array_equal(W[r0:r0 + rows_per].view(np.uint16), Ws.view(np.uint16)), r0
    for bs in (1, 8):
        x = rng.standard_normal((bs, K)).astype(np.float16)
        out = ops.lut_gemv(d, cuda(x), dl, M, K, R, vec).cpu().numpy()
        assert rel_l2(out, gemv64(W, x)) <= REL_L2_TOL, bs


def _h(a):
    return np.asarray(a, np.float16)


@pytest.mark.parametrize("K", [4096, 8192])
@pytest.mark.parametrize("kind", ["tcomb", "vq"])
def test_fused_prologue_gemv(ops, cref, K, kind):
    """qp_*_gemv_fused: residual add + RMSNorm + sign + Hadamard computed in the GEMV prologue, against the restatement
    with the fp16 rounding points of the reference graph (incoherent_linear.py:76-108,324-338 + fp16 residual/RMSNorm)."""
    from qpalette import _cabi
    from qpalette._cabi import SPLIT_IN, check, lib
    M, S, eps, scale = 1024, 9, 1e-5, 64.0
    rng = np.random.default_rng(K)
    h = rng.standard_normal(K).astype(np.float16)
    acc = (rng.standard_normal(K) * 40).astype(np.float32)
    ws = (rng.uniform(0.5, 1.5, K) / 64 / 40).astype(np.float16)
    nw = rng.uniform(0.5, 1.5, K).astype(np.float16)
    su = rng.choice([-1.0, 1.0], K).astype(np.float16)
    had_scale = 1.0 / (math.sqrt(K) * scale)
    # restatement (float64 between the fp16 rounding points)
    t = _h(_h(_h(acc) * ws) * np.float16(scale))
    h2 = _h(h.astype(np.float32) + t.astype(np.float32))
    v = h2.astype(np.float64)
    rstd = 1.0 / math.sqrt((v * v).mean() + eps)
    y = _h(nw.astype(np.float32) * _h(v * rstd).astype(np.float32)).astype(np.float64) * su.astype(np.float64)
    xref = _h(O.hadamard_ref(y[None, :])[0] * math.sqrt(K) * had_scale)

    if kind == "tcomb":
        b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
        b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
        tl = (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)
        W = cref_combt(cref, b1, b2, tl, M, K, 6, 7, S)
        codes = (cuda(b1), cuda(b2), cuda(tl))
    else:
        R = 8
        tl = rng.standard_normal((1 << R, 2)).astype(np.float16)
        b1 = rng.integers(0, 256, size=M * K * R // 16, dtype=np.uint8)
        W = O.lut_tc_decode(b1.view(np.int32), tl, M, K, R, 2)
        codes = (cuda(b1), None, cuda(tl))
    ref = gemv64(W, xref[None, :])

    d = dict(h=cuda(h), acc=cuda(acc), ws=cuda(ws), nw=cuda(nw), su=cuda(su))
    h_out = torch.zeros(K, dtype=torch.float16, device="cuda")
    x_out = torch.zeros(K, dtype=torch.float16, device="cuda")
    z1 = torch.ones(3000, dtype=torch.float32, device="cuda")
    z2 = torch.ones(520, dtype=torch.float32, device="cuda")
    out = torch.zeros(M, dtype=torch.float32, device="cuda")
    p = lambda t_: t_.data_ptr()
    xp = _cabi.XProd(p(d["h"]), p(h_out), p(d["acc"]), p(d["ws"]), scale, p(d["nw"]), eps, p(d["su"]), had_scale, p(x_out),
                     p(z1), z1.numel(), p(z2), z2.numel())
    st = torch.cuda.current_stream().cuda_stream
    if kind == "tcomb":
        check(lib().qp_tcq_gemv_fused(p(out), p(codes[0]), p(codes[1]), ctypes.addressof(xp), p(codes[2]), M, K, S, 6, 7,
                                      SPLIT_IN, K // 2, st))
    else:
        check(lib().qp_lut_gemv_fused(p(out), p(codes[0]), ctypes.addressof(xp), p(codes[2]), M, K, 8, 2, st))
    torch.cuda.synchronize()
    assert float(z1.abs().sum()) == 0.0 and float(z2.abs().sum()) == 0.0  # accumulator-clearing duty
    assert np.array_equal(h_out.cpu().numpy().view(np.uint16), h2.view(np.uint16))  # residual stream: bit-exact
    assert rel_l2(x_out.float().cpu().numpy(), xref) <= REL_L2_TOL
    assert rel_l2(out.cpu().numpy(), ref) <= REL_L2_TOL


def test_tcq_gemv_host_entry(ops, cref):
    """qp_tcq_gemv_host (include/qpalette.h): activations / result in HOST memory, copies on the given stream"""
    from qpalette._cabi import SPLIT_IN, check, lib
    M, K, S = 4096, 4096, 9  # BASELINE.json configs[0]
    rng = np.random.default_rng(11)
    b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
    tl = (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)
    x = torch.from_numpy(rng.standard_normal((1, K)).astype(np.float16)).pin_memory()
    out_h = torch.zeros((1, M), dtype=torch.float32).pin_memory()
    d1, d2, dt = cuda(b1), cuda(b2), cuda(tl)
    x_d = torch.empty((1, K), dtype=torch.float16, device="cuda")
    out_d = torch.empty((1, M), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    check(lib().qp_tcq_gemv_host(out_h.data_ptr(), out_d.data_ptr(), d1.data_ptr(), d2.data_ptr(), x.data_ptr(), x_d.data_ptr(),
                                 dt.data_ptr(), M, K, 1, S, 6, 7, SPLIT_IN, K // 2, st.cuda_stream))
    st.synchronize()
    ref = gemv64(cref_combt(cref, b1, b2, tl, M, K, 6, 7, S), x.numpy())
    assert rel_l2(out_h.numpy(), ref) <= REL_L2_TOL
