"""mma.sync fused dequant + GEMM for 9 <= bs <= 32 (K-slabs of x in shared memory, `qp_tcq_gemm_mma`) against the float64
oracle (rel-L2 <= 1e-3); the reference's path at these batch sizes is dequantise + cuBLAS (lib/linear/tcq_linear.py:75-84)."""
import numpy as np
import pytest
import torch

from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rand_tlut(rng, S):
    return (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)


@pytest.mark.parametrize("KV,S", [(2, 9), (3, 9), (4, 9), (5, 9), (6, 9), (7, 9), (8, 9), (9, 10), (10, 11)])
@pytest.mark.parametrize("bs", [9, 16, 17, 32])
def test_single_rate(KV, S, bs):
    from qpalette import ops
    rng = np.random.default_rng(KV * 100 + bs)
    M, K = 256, 448  # 14 super-tile columns: one narrow slab
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemm_mma(cuda(buf), cuda(x), cuda(tl), M, K, S, KV).cpu().numpy()
    assert out.shape == (bs, M)
    assert rel_l2(out, O.gemv_ref(O.tcq_decode(buf, tl, M, K, KV, S), x)) <= TOL


@pytest.mark.parametrize("mode,kv,S", [("combt", (6, 7), 9), ("combt", (3, 4), 9), ("combt", (8, 9), 10), ("combt", (9, 10), 11),
                                       ("comb", (6, 7), 9), ("comb", (2, 3), 9)])
@pytest.mark.parametrize("shape,bs", [((512, 1024), 12), ((512, 2560), 32), ((128, 4224), 20), ((32, 64), 9)])
def test_two_rate(mode, kv, S, shape, bs):
    """K = 2560 / 4224: parts of 40 / 66 super-tile columns = a full slab + a narrow one (NB = 4) or one narrow / 64 + 2 (NB = 2);
    32 x 64: fewer super-tiles than CTAs"""
    from qpalette import ops
    from qpalette._cabi import SPLIT_IN, SPLIT_OUT
    KV1, KV2 = kv
    M, K = shape
    if mode == "comb" and M < 64:
        pytest.skip("out-split needs two 32-row halves")
    rng = np.random.default_rng(KV1 * 16 + KV2 + M + bs)
    if mode == "combt":
        m1, k1, m2, k2 = M, K // 2, M, K // 2
    else:
        m1, k1, m2, k2 = M // 2, K, M // 2, K
    b1 = rng.integers(0, 256, size=m1 * k1 * KV1 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=m2 * k2 * KV2 // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    dec = O.tcq_decode_combt if mode == "combt" else O.tcq_decode_comb
    Wref = dec(b1, b2, tl, M, K, KV1, KV2, S)
    split, part1 = (SPLIT_IN, K // 2) if mode == "combt" else (SPLIT_OUT, M // 2)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemm_mma(cuda(b1), cuda(x), cuda(tl), M, K, S, KV1, cuda(b2), KV2, split, part1).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(Wref, x)) <= TOL


@pytest.mark.parametrize("M,K", [(4096, 14336), (6144, 4096)])
def test_llama_shapes_and_dispatch(M, K):
    """down_proj / merged qkv of Llama-3.1-8B at bs = 24 through the public dispatcher (tcq_gemm_tc routes bs <= 32 here) and at
    bs = 70 through the mma entry itself (three launches of <= 32 batch rows), accumulating onto an existing output"""
    from qpalette import ops
    from qpalette._cabi import SPLIT_IN
    rng = np.random.default_rng(M + K)
    b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
    tl = rand_tlut(rng, 9)
    W = O.tcq_decode_combt(b1, b2, tl, M, K, 6, 7, 9)
    args = (cuda(b1), None, cuda(tl), M, K, 9, 6, cuda(b2), 7, SPLIT_IN, K // 2)
    x = rng.standard_normal((24, K)).astype(np.float16)
    assert 24 <= ops.MMA_GEMM_MAX_BS
    out = ops.tcq_gemm_tc(args[0], cuda(x), *args[2:]).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(W, x)) <= TOL
    x = rng.standard_normal((70, K)).astype(np.float16)
    base = rng.standard_normal((70, M)).astype(np.float32)
    out = ops.tcq_gemm_mma(args[0], cuda(x), *args[2:], out=cuda(base), accumulate=True).cpu().numpy()
    assert rel_l2(out - base, O.gemv_ref(W, x)) <= TOL


def test_matches_gemv_and_tcgen05():
    """the three fused paths (GEMV bs <= 8, mma, tcgen05) agree on the same weights"""
    from qpalette import ops
    rng = np.random.default_rng(5)
    M, K, KV, S = 512, 1024, 6, 9
    buf = cuda(rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8))
    tl = cuda(rand_tlut(rng, S))
    x = cuda(rng.standard_normal((16, K)).astype(np.float16))
    a = ops.tcq_gemm_mma(buf, x, tl, M, K, S, KV)
    old, ops.MMA_GEMM_MAX_BS = ops.MMA_GEMM_MAX_BS, 0
    try:
        b = ops.tcq_gemm_tc(buf, x, tl, M, K, S, KV)
    finally:
        ops.MMA_GEMM_MAX_BS = old
    c = torch.cat([ops.tcq_gemv(buf, x[i:i + 8], tl, M, K, S, KV) for i in (0, 8)])
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= TOL
    assert rel_l2(a.cpu().numpy(), c.float().cpu().numpy()) <= TOL


@pytest.mark.parametrize("vec,R", [(2, 2), (2, 3), (2, 4), (2, 6), (2, 8), (2, 9), (2, 11), (2, 12), (1, 2), (1, 3), (1, 4), (1, 5),
                                   (1, 6), (1, 7), (1, 8)])
@pytest.mark.parametrize("bs", [9, 32])
def test_lut_gemm_mma(vec, R, bs):
    """VQ (vec_sz 2) and SQ (vec_sz 1, incl. the 6..8-bit split-lookup formats the tcgen05 kernel does not take)"""
    from qpalette import ops
    rng = np.random.default_rng(R * 2 + vec + bs)
    M, K = 384, 1312  # 41 super-tile columns: a full slab + a narrow one at NB = 4
    lut = rng.standard_normal((1 << R, vec)).astype(np.float16)
    buf = rng.integers(0, 256, size=M * K * R // 8 // vec, dtype=np.uint8)
    Wref = O.lut_tc_decode(buf.view(np.int32), lut, M, K, R, vec)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.lut_gemm_mma(cuda(buf), cuda(x), cuda(lut), M, K, R, vec).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(Wref, x)) <= TOL


def test_lut_llama_shape_through_module_dispatch():
    """ldlq_2_8 up_proj-sized layer at bs = 20 through the public dispatcher, accumulating"""
    from qpalette import ops
    rng = np.random.default_rng(3)
    M, K, R, bs = 14336, 4096, 8, 20
    lut = rng.standard_normal((1 << R, 2)).astype(np.float16)
    buf = rng.integers(0, 256, size=M * K * R // 16, dtype=np.uint8)
    Wref = O.lut_tc_decode(buf.view(np.int32), lut, M, K, R, 2)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    base = rng.standard_normal((bs, M)).astype(np.float32)
    out = ops.lut_gemm_tc(cuda(buf), cuda(x), cuda(lut), M, K, R, 2, out=cuda(base), accumulate=True).cpu().numpy()
    assert rel_l2(out - base, O.gemv_ref(Wref, x)) <= TOL
