"""GPU parity tests: the CUDA path (through the C ABI, libqpalette.so) against the oracle and against the committed
reference-generated golden vectors.  Decoded weights must be BIT-EXACT; GEMV outputs rel-L2 <= 1e-3 vs the fp64 oracle
(the tolerance is the one BASELINE.json's north_star states)."""
import numpy as np
import pytest
import torch

from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu

REL_L2_TOL = 1e-3


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from qpalette import ops as _ops
    return _ops


def S_of(KV):
    return 9 if KV <= 8 else KV + 1


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bits_equal(t, ref):
    return np.array_equal(t.cpu().numpy().view(np.uint16), np.asarray(ref).view(np.uint16))


def rand_tlut(rng, S):
    return (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)


# ---------------------------------------------------------------------------------------------------------------- TCQ
@pytest.mark.parametrize("KV", range(2, 11))
def test_tcq_dequant_golden_and_random(ops, golden, KV):
    S = S_of(KV)
    tlut = golden[f"tlut_{S}"]
    if KV % 2 == 0:  # reference torch decoder output, straight from the fixture
        W = ops.tcq_dequant(cuda(golden[f"tcq_buf_{KV}"]), cuda(tlut), 64, 128, S, KV)
        assert bits_equal(W, golden[f"tcq_decode_compressed_{KV}"])
    # reference pack_trellis + swizzle -> kernel decode == reference recons
    M, K = 64, 128
    packed = O.tcq_swizzle(golden[f"tcq_pack_trellis_{KV}"], M, K, KV)
    W = ops.tcq_dequant(cuda(packed), cuda(tlut), M, K, S, KV)
    assert bits_equal(W, O.tcq_expected_from_states(golden[f"tcq_states_{KV}"], tlut, M, K, S))
    # random bytes, odd shape multiples, every S that the reference registers for this KV
    rng = np.random.default_rng(100 + KV)
    for S2 in {9: (9,), 10: (9, 10), 11: (9, 10, 11)}[max(9, min(11, KV + 1 if KV >= 8 else 9))]:
        M, K = 96, 160
        buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
        tl = rand_tlut(rng, S2)
        W = ops.tcq_dequant(cuda(buf), cuda(tl), M, K, S2, KV)
        assert bits_equal(W, O.tcq_decode(buf, tl, M, K, KV, S2)), (KV, S2)


@pytest.mark.parametrize("KV", range(2, 11))
@pytest.mark.parametrize("bs", [1, 3, 8])
def test_tcq_gemv(ops, KV, bs):
    S = S_of(KV)
    rng = np.random.default_rng(7 * KV + bs)
    M, K = 160, 448
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = ops.tcq_gemv(cuda(buf), cuda(x), cuda(tl), M, K, S, KV).cpu().numpy()
    ref = O.gemv_ref(O.tcq_decode(buf, tl, M, K, KV, S), x)
    assert out.shape == (bs, M)
    assert rel_l2(out, ref) <= REL_L2_TOL


def test_tcq_gemv_llama_shape_and_accumulate(ops):
    """a q_proj-sized strip count with a K that does not divide the per-warp run; also the accumulate flag."""
    rng = np.random.default_rng(5)
    M, K, KV, S = 1024, 4096, 6, 9
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    x = rng.standard_normal((1, K)).astype(np.float16)
    ref = O.gemv_ref(O.tcq_decode(buf, tl, M, K, KV, S), x)
    d_buf, d_x, d_tl = cuda(buf), cuda(x), cuda(tl)
    out = ops.tcq_gemv(d_buf, d_x, d_tl, M, K, S, KV)
    assert rel_l2(out.cpu().numpy(), ref) <= REL_L2_TOL
    out2 = ops.tcq_gemv(d_buf, d_x, d_tl, M, K, S, KV, out=out.clone(), accumulate=True)
    assert rel_l2(out2.cpu().numpy(), 2 * ref) <= REL_L2_TOL
    # kernel all-ones known-answer test (the reference's only arithmetic KAT idea, vq-tensor-kernels/src/test.cu):
    ones = np.ones((1 << S, 2), np.float16)
    out3 = ops.tcq_gemv(d_buf, cuda(np.ones((1, K), np.float16)), cuda(ones), M, K, S, KV).cpu().numpy()
    # sign bit may flip component 0: |out| <= K and out == K - 2 * (#negated)
    assert np.all(np.abs(out3) <= K) and np.all(out3 == np.round(out3))


@pytest.mark.parametrize("mode", ["combt", "comb"])
@pytest.mark.parametrize("kv", [(6, 7), (7, 8), (5, 6), (9, 10), (8, 6)])
def test_tcq_two_rate(ops, mode, kv):
    KV1, KV2 = kv
    S = S_of(max(kv))
    rng = np.random.default_rng(KV1 * 16 + KV2)
    M, K = 128, 512
    if mode == "combt":
        m1, k1, m2, k2 = M, K // 2, M, K // 2
    else:
        m1, k1, m2, k2 = M // 2, K, M // 2, K
    b1 = rng.integers(0, 256, size=m1 * k1 * KV1 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=m2 * k2 * KV2 // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    dec = O.tcq_decode_combt if mode == "combt" else O.tcq_decode_comb
    Wref = dec(b1, b2, tl, M, K, KV1, KV2, S)
    from qpalette._cabi import SPLIT_IN, SPLIT_OUT
    split, part1 = (SPLIT_IN, K // 2) if mode == "combt" else (SPLIT_OUT, M // 2)
    W = ops.tcq_dequant(cuda(b1), cuda(tl), M, K, S, KV1, cuda(b2), KV2, split, part1)
    assert bits_equal(W, Wref)
    for bs in (1, 4):
        x = rng.standard_normal((bs, K)).astype(np.float16)
        out = ops.tcq_gemv(cuda(b1), cuda(x), cuda(tl), M, K, S, KV1, cuda(b2), KV2, split, part1).cpu().numpy()
        assert rel_l2(out, O.gemv_ref(Wref, x)) <= REL_L2_TOL


def test_tcq_unequal_in_split(ops):
    """the reference falls back to two launches when the halves differ (comb_linear.py:198-201); here it is one call."""
    from qpalette._cabi import SPLIT_IN
    rng = np.random.default_rng(77)
    M, K, k1, KV1, KV2, S = 64, 384, 128, 6, 7, 9
    b1 = rng.integers(0, 256, size=M * k1 * KV1 // 16, dtype=np.uint8)
    b2 = rng.integers(0, 256, size=M * (K - k1) * KV2 // 16, dtype=np.uint8)
    tl = rand_tlut(rng, S)
    Wref = O.tcq_decode_combt(b1, b2, tl, M, K, KV1, KV2, S, in_part=(k1, K - k1))
    x = rng.standard_normal((2, K)).astype(np.float16)
    out = ops.tcq_gemv(cuda(b1), cuda(x), cuda(tl), M, K, S, KV1, cuda(b2), KV2, SPLIT_IN, k1).cpu().numpy()
    assert rel_l2(out, O.gemv_ref(Wref, x)) <= REL_L2_TOL


def test_error_paths(ops):
    from qpalette._cabi import QPaletteError
    z = torch.zeros(64, dtype=torch.uint8, device="cuda")
    tl = torch.zeros((512, 2), dtype=torch.float16, device="cuda")
    with pytest.raises(QPaletteError):
        ops.tcq_gemv(z, torch.zeros((1, 48), dtype=torch.float16, device="cuda"), tl, 32, 48, 9, 6)  # K % 32
    with pytest.raises(QPaletteError):
        ops.tcq_gemv(z, torch.zeros((9, 64), dtype=torch.float16, device="cuda"), tl, 32, 64, 9, 6)  # bs > 8
    with pytest.raises(QPaletteError):
        ops.tcq_gemv(z, torch.zeros((1, 64), dtype=torch.float16, device="cuda"), tl, 32, 64, 12, 6)  # S
    with pytest.raises(QPaletteError):
        ops.tcq_gemv(z.cpu(), torch.zeros((1, 64), dtype=torch.float16), tl.cpu(), 32, 64, 9, 6)  # CPU tensors


# ------------------------------------------------------------------------------------------------- VQ / SQ, TC layout
LUT_CASES = [(1, r) for r in range(2, 9)] + [(2, r) for r in range(2, 13)]


@pytest.mark.parametrize("vec,R", LUT_CASES)
def test_lut_tc(ops, golden, vec, R):
    rng = np.random.default_rng(R * 2 + vec)
    lut = rng.standard_normal((1 << R, vec)).astype(np.float16)
    # golden: codes packed by the reference's pack_qweight
    M, K = 64, 128
    W = ops.lut_dequant(cuda(golden[f"lut_tc_packed_{vec}_{R}"]), cuda(lut), M, K, R, vec)
    assert bits_equal(W, lut[golden[f"lut_tc_codes_{vec}_{R}"]].reshape(M, K))
    # random bytes + GEMV
    M, K = 96, 320
    buf = rng.integers(0, 256, size=M * K * R // 8 // vec, dtype=np.uint8)
    Wref = O.lut_tc_decode(buf.view(np.int32), lut, M, K, R, vec)
    assert bits_equal(ops.lut_dequant(cuda(buf), cuda(lut), M, K, R, vec), Wref)
    for bs in (1, 5):
        x = rng.standard_normal((bs, K)).astype(np.float16)
        out = ops.lut_gemv(cuda(buf), cuda(x), cuda(lut), M, K, R, vec).cpu().numpy()
        assert rel_l2(out, O.gemv_ref(Wref, x)) <= REL_L2_TOL


# --------------------------------------------------------------------------------------------------------- SIMT layout
SIMT_CASES = [(1, r, k) for r in (2, 3, 4, 5, 6, 7, 8) for k in (1024, 1280)] + \
             [(2, r, k) for r in (2, 3, 5, 6, 8, 9, 10, 12) for k in (2048, 2560)] + \
             [(4, r, k) for r in (6, 7, 8, 9, 10, 11, 12) for k in (4096, 5120)]  # 4-wide vectors: lib/linear/__init__.py:383-420


@pytest.mark.parametrize("vec,R,K", SIMT_CASES)
def test_simt(ops, golden, vec, R, K):
    rng = np.random.default_rng(R * 5 + vec + K)
    lut = rng.standard_normal((1 << R, vec)).astype(np.float16)
    M = 40
    codes = rng.integers(0, 1 << R, size=(M, K // vec))
    packed = O.simt_pack(codes, M, K, R, vec)
    key = f"simt_packed_{vec}_{R}_{K}"
    if key in golden.files:  # reference-packed buffer
        gW = ops.simt_dequant(cuda(golden[key]), cuda(lut), 8, K, R, vec)
        assert bits_equal(gW, lut[golden[f"simt_codes_{vec}_{R}_{K}"]].reshape(8, K))
    Wref = lut[codes].reshape(M, K)
    assert bits_equal(ops.simt_dequant(cuda(packed), cuda(lut), M, K, R, vec), Wref)
    for bs in (1, 4):
        x = rng.standard_normal((bs, K)).astype(np.float16)
        out = ops.simt_gemv(cuda(packed), cuda(x), cuda(lut), M, K, R, vec).cpu().numpy()
        assert out.dtype == np.float16
        assert rel_l2(out, O.gemv_ref(Wref, x)) <= 2e-3  # fp16 output rounding on top of the 1e-3 budget


@pytest.mark.parametrize("vec,R", [(1, 4), (2, 6), (1, 7), (2, 11)])
def test_convert_tc_to_simt(ops, golden, vec, R):
    key = f"conv_tc_{vec}_{R}"
    if key in golden.files:
        got = ops.convert_tc_to_simt(cuda(golden[key]), 32, 2048, R, vec).cpu().numpy()
        assert np.array_equal(got, golden[f"conv_simt_{vec}_{R}"])
    rng = np.random.default_rng(R)
    M, K = 64, 2048 + 512 * vec
    codes = rng.integers(0, 1 << R, size=(M, K // vec))
    tc = O.lut_tc_pack(codes, M, K, R, vec)
    got = ops.convert_tc_to_simt(cuda(tc), M, K, R, vec).cpu().numpy()
    assert np.array_equal(got, O.simt_pack(codes, M, K, R, vec))


# ------------------------------------------------------------------------------------------------------------ Hadamard
@pytest.mark.parametrize("n", [1024, 4096, 14336])
def test_hadamard_golden(ops, golden, n):
    x = golden[f"had_x_{n}"]
    y = ops.hadamard(cuda(x), None, n ** -0.5).cpu().numpy()
    assert np.allclose(y, golden[f"had_Ut_{n}"], atol=2e-4)
    assert np.allclose(y, golden[f"had_cuda_T_{n}"], atol=2e-4)


@pytest.mark.parametrize("n", [64, 512, 2048, 8192, 16384, 28672, 3584])
@pytest.mark.parametrize("dt", ["f16", "f32"])
def test_hadamard_random(ops, n, dt):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(np.float16 if dt == "f16" else np.float32)
    su = np.where(rng.standard_normal(n) > 0, 1.0, -1.0).astype(np.float16)
    y = ops.hadamard(cuda(x), cuda(su), n ** -0.5 / 64.0).cpu().numpy()
    ref = O.hadamard_ref(x.astype(np.float64) * su.astype(np.float64)) / 64.0
    assert y.dtype == x.dtype
    assert rel_l2(y, ref) <= (2e-3 if dt == "f16" else 1e-5)


def test_scale_epilogue(ops):
    from qpalette._cabi import EPI_SILU_MUL
    rng = np.random.default_rng(3)
    acc = rng.standard_normal((2, 256)).astype(np.float32) * 3
    ws = (rng.random(256) * 0.02 + 0.01).astype(np.float16)
    out = ops.scale_epilogue(cuda(acc), cuda(ws), 64.0).cpu()
    ref = (torch.from_numpy(acc).half() * torch.from_numpy(ws) * 64.0)
    assert torch.equal(out, ref)
    out2 = ops.scale_epilogue(cuda(acc), cuda(ws), 64.0, EPI_SILU_MUL).cpu()
    up, gate = ref[:, :128], ref[:, 128:]
    ref2 = torch.nn.functional.silu(gate.float()).half() * up
    assert torch.allclose(out2.float(), ref2.float(), rtol=2e-3, atol=1e-4)


# ------------------------------------------------------------------------------------------ reference operator names
def test_reference_operator_names(ops):
    """the reference's shape-templated `torch.ops.ours_lib.*` names (lib/linear/__init__.py:43-420) resolve onto the C ABI:
    same schemas, same results as the direct wrappers, fake (meta) implementations for tracing, AttributeError for garbage"""
    rng = np.random.default_rng(5)
    M, K, KV, S = 256, 512, 6, 9
    buf = cuda(rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8))
    tl = cuda(rand_tlut(rng, S))
    x = cuda(rng.standard_normal((1, K)).astype(np.float16))
    op = ops.resolve(f"decompress_gemm_tcq_{M}_1_{K}_{S}_{KV}")
    assert op is getattr(torch.ops.ours_lib, f"decompress_gemm_tcq_{M}_1_{K}_{S}_{KV}")
    y = op(buf.view(torch.int16), x, tl)
    assert y.dtype == torch.float32 and tuple(y.shape) == (1, M)
    assert torch.allclose(y, ops.tcq_gemv(buf, x, tl, M, K, S, KV), rtol=1e-4, atol=1e-4)  # fp32 atomics: order varies
    W = ops.resolve(f"decompress_tcq_{S}_{KV}")(buf.view(torch.int16), tl, M, K)
    assert W.dtype == torch.float16 and tuple(W.shape) == (M, K)
    assert bits_equal(W, O.tcq_decode(buf.cpu().numpy(), tl.cpu().numpy(), M, K, KV, S))
    # vector quantizer, tensor-core layout
    R = 6
    lut = cuda(rng.standard_normal((1 << R, 2)).astype(np.float16))
    q = cuda(rng.integers(0, 256, size=M * K * R // 16, dtype=np.uint8))
    y2 = ops.resolve(f"decompress_gemm_{M}_1_{K}_{R}_vq2")(q.view(torch.int32).view(M, -1), x, lut)
    assert torch.allclose(y2, ops.lut_gemv(q, x, lut, M, K, R, 2), rtol=1e-4, atol=1e-4)
    # meta / fake implementation (what torch.compile traces)
    from torch._subclasses.fake_tensor import FakeTensorMode
    b16 = buf.view(torch.int16)
    with FakeTensorMode() as mode:
        fy = op(mode.from_tensor(b16), mode.from_tensor(x), mode.from_tensor(tl))
        assert tuple(fy.shape) == (1, M) and fy.dtype == torch.float32
    with pytest.raises(AttributeError):
        ops.resolve("decompress_gemm_tcq_not_an_op")
    # SIMT ops with 4-wide vectors exist for 6..12-bit codes, as in the reference (lib/linear/__init__.py:383-420)
    R4, K4, M4 = 8, 4096, 64
    lut4 = rng.standard_normal((1 << R4, 4)).astype(np.float16)
    codes4 = rng.integers(0, 1 << R4, size=(M4, K4 // 4))
    packed4 = cuda(O.simt_pack(codes4, M4, K4, R4, 4))
    x4 = cuda(rng.standard_normal((2, 1, K4)).astype(np.float16))
    y4 = ops.resolve(f"vq_pack_gemm_simt_2_4_{R4}")(x4, packed4, cuda(lut4))
    assert tuple(y4.shape) == (2, 1, M4) and y4.dtype == torch.float16
    W4 = ops.resolve(f"vq_pack_dequant_simt_4_{R4}")(packed4, cuda(lut4), M4, K4)
    assert bits_equal(W4, lut4[codes4].reshape(M4, K4))
    assert rel_l2(y4.cpu().numpy().reshape(2, M4), O.gemv_ref(lut4[codes4].reshape(M4, K4), x4.cpu().numpy().reshape(2, K4))) <= 2e-3
    for bad in ("vq_pack_gemm_simt_1_4_5", "vq_pack_dequant_simt_4_13", "vq_pack_dequant_simt_3_8"):
        with pytest.raises(AttributeError):
            ops.resolve(bad)
