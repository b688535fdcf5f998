"""tcgen05 fused dequant + batched GEMM (bs >= 16) against the float64 oracle (rel-L2 <= 1e-3)."""
import numpy as np
import pytest
import torch

from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(autouse=True)
def tcgen05_only(monkeypatch):
    """this file tests the tcgen05 kernel: keep small batches off the mma.sync path (tests/test_gpu_gemm_mma.py)"""
    from qpalette import ops
    monkeypatch.setattr(ops, "MMA_GEMM_MAX_BS", 0)


def rand_tlut(rng, S):
    return (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)


@pytest.mark.parametrize("KV,S", [(6, 9), (7, 9), (8, 9), (4, 9), (3, 9), (9, 10), (10, 11)])
@pytest.mark.parametrize("bs", [16, 40, 128])
def test_tcq_gemm_tc(KV, S, bs):
    from qpalette import ops
    rng = np.random.default_rng(KV * 100 + bs)
    M, K = 256, 448
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemm_tc(cuda(buf), cuda(x), cuda(tl), M, K, S, KV).cpu().numpy()
    ref = O.gemv_ref(O.tcq_decode(buf, tl, M, K, KV, S), x)
    assert out.shape == (bs, M)
    assert rel_l2(out, ref) <= 1e-3


@pytest.mark.parametrize("mode,kv", [("combt", (6, 7)), ("combt", (7, 8)), ("comb", (6, 7)), ("comb", (8, 6))])
def test_tcq_gemm_tc_two_rate(mode, kv):
    from qpalette import ops
    from qpalette._cabi import SPLIT_IN, SPLIT_OUT
    KV1, KV2 = kv
    rng = np.random.default_rng(KV1 * 16 + KV2)
    M, K, S, bs = 512, 1024, 9, 33
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
    out = ops.tcq_gemm_tc(cuda(b1), cuda(x), cuda(tl), M, K, S, KV1, cuda(b2), KV2, split, part1).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(Wref, x)) <= 1e-3


@pytest.mark.parametrize("vec,R", [(2, 2), (2, 6), (2, 8), (2, 11), (2, 12), (1, 2), (1, 4), (1, 5)])
def test_lut_gemm_tc(vec, R):
    from qpalette import ops
    rng = np.random.default_rng(R * 2 + vec)
    M, K, bs = 384, 320, 64
    lut = rng.standard_normal((1 << R, vec)).astype(np.float16)
    buf = rng.integers(0, 256, size=M * K * R // 8 // vec, dtype=np.uint8)
    Wref = O.lut_tc_decode(buf.view(np.int32), lut, M, K, R, vec)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.lut_gemm_tc(cuda(buf), cuda(x), cuda(lut), M, K, R, vec).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(Wref, x)) <= 1e-3


def test_llama_shape_large_batch():
    """a q_proj-sized layer at bs = 200 (two chunks of the 128-row limit), tcomb_6_7"""
    from qpalette import ops
    from qpalette._cabi import SPLIT_IN
    rng = np.random.default_rng(9)
    M, K, S, bs = 1024, 4096, 9, 200
    b1 = rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemm_tc(cuda(b1), cuda(x), cuda(tl), M, K, S, 6, cuda(b2), 7, SPLIT_IN, K // 2).cpu().numpy()
    ref = O.gemv_ref(O.tcq_decode_combt(b1, b2, tl, M, K, 6, 7, S), x)
    assert rel_l2(out, ref) <= 1e-3


@pytest.mark.parametrize("accumulate", [False, True])
def test_many_row_blocks_plain_store(accumulate):
    """enough 256-row blocks that K is not split: the epilogue stores (or read-modify-writes) instead of atomics"""
    from qpalette import ops
    rng = np.random.default_rng(31 + accumulate)
    M, K, KV, S, bs = 256 * 120 + 128, 128, 5, 9, 24  # the last block has 128 rows
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    ref = O.gemv_ref(O.tcq_decode(buf, tl, M, K, KV, S), x)
    if accumulate:
        base = rng.standard_normal((bs, M)).astype(np.float32)
        out = ops.tcq_gemm_tc(cuda(buf), cuda(x), cuda(tl), M, K, S, KV, out=cuda(base), accumulate=True).cpu().numpy()
        assert rel_l2(out - base, ref) <= 1e-3
    else:
        out = ops.tcq_gemm_tc(cuda(buf), cuda(x), cuda(tl), M, K, S, KV).cpu().numpy()
        assert rel_l2(out, ref) <= 1e-3


def test_split_k_accumulate():
    """few row blocks (K split across CTAs, atomics) adding onto an existing output"""
    from qpalette import ops
    rng = np.random.default_rng(77)
    M, K, R, bs = 512, 1024, 6, 100
    lut = rng.standard_normal((1 << R, 2)).astype(np.float16)
    buf = rng.integers(0, 256, size=M * K * R // 16, dtype=np.uint8)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    ref = O.gemv_ref(O.lut_tc_decode(buf.view(np.int32), lut, M, K, R, 2), x)
    base = rng.standard_normal((bs, M)).astype(np.float32)
    out = ops.lut_gemm_tc(cuda(buf), cuda(x), cuda(lut), M, K, R, 2, out=cuda(base), accumulate=True).cpu().numpy()
    assert rel_l2(out - base, ref) <= 1e-3
