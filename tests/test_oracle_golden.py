"""Pin oracle/qp_oracle.py against vectors produced by the reference's own python (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import qp_oracle as O

KVS = list(range(2, 11))


def S_of(KV):
    return 9 if KV <= 8 else KV + 1


@pytest.mark.parametrize("S", [9, 10, 11])
def test_quantlut_sym(golden, S):
    exp = O.quantlut_sym(golden[f"tlut_{S}"], S)
    idx = golden[f"quantlut_sym_{S}_sample_idx"]
    assert exp.dtype == np.float16
    assert np.array_equal(exp[idx].view(np.uint16), golden[f"quantlut_sym_{S}_sample"].view(np.uint16))
    s = golden[f"quantlut_sym_{S}_sum"]
    assert np.isclose(exp.astype(np.float64).sum(), s[0]) and np.isclose(np.abs(exp.astype(np.float64)).sum(), s[1])


@pytest.mark.parametrize("KV", KVS)
def test_pack_trellis_matches_reference(golden, KV):
    st = golden[f"tcq_states_{KV}"]
    assert np.array_equal(O.tcq_pack_trellis(st, KV).view(np.int16), golden[f"tcq_pack_trellis_{KV}"])


@pytest.mark.parametrize("KV", KVS)
def test_tcq_roundtrip_identity(golden, KV):
    """SURVEY 8c identity 1: decode(swizzle(pack_trellis(states))) == quantlut_sym[states] via _INV_PERMUTE,
    with pack_trellis and recons taken from the reference run."""
    M, K, S = 64, 128, S_of(KV)
    st = golden[f"tcq_states_{KV}"]
    packed = O.tcq_swizzle(golden[f"tcq_pack_trellis_{KV}"], M, K, KV)
    tlut = golden[f"tlut_{S}"]
    W = O.tcq_decode(packed, tlut, M, K, KV, S)
    # reference recons: (2, tiles, 128) -> (tiles, 128, 2) in trellis order
    rec = golden[f"tcq_recons_{KV}"].transpose(1, 2, 0).reshape(-1, 256)[:, O.INV_PERMUTE]
    Wref = rec.reshape(M // 16, K // 16, 16, 16).transpose(0, 2, 1, 3).reshape(M, K)
    assert np.array_equal(W.view(np.uint16), Wref.view(np.uint16))
    assert np.array_equal(W.view(np.uint16), O.tcq_expected_from_states(st, tlut, M, K, S).view(np.uint16))
    # states recovered exactly, both decoders
    got = O.tcq_states(packed, M, K, KV)
    assert np.array_equal(got, O.tcq_states_bitwise(packed, M, K, KV))
    assert np.array_equal(O.tcq_pack(st, M, K, KV), packed)


@pytest.mark.parametrize("KV", [2, 4, 6, 8, 10])
def test_tcq_decode_vs_reference_decoder(golden, KV):
    """even KV: the reference's torch decoder (lib/utils/kernel_decompress.py) on random bytes."""
    M, K, S = 64, 128, S_of(KV)
    W = O.tcq_decode(golden[f"tcq_buf_{KV}"], golden[f"tlut_{S}"], M, K, KV, S)
    assert np.array_equal(W.view(np.uint16), golden[f"tcq_decode_compressed_{KV}"].view(np.uint16))


@pytest.mark.parametrize("KV", KVS)
def test_tcq_random_bytes_two_decoders(KV):
    rng = np.random.default_rng(KV)
    M, K = 64, 96
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    assert np.array_equal(O.tcq_states(buf, M, K, KV), O.tcq_states_bitwise(buf, M, K, KV))


@pytest.mark.parametrize("vec,R", [(1, r) for r in range(2, 9)] + [(2, r) for r in range(2, 13)])
def test_lut_tc_layout(golden, vec, R):
    M, K = 64, 128
    Q = golden[f"lut_tc_codes_{vec}_{R}"]
    packed = golden[f"lut_tc_packed_{vec}_{R}"]
    assert np.array_equal(O.lut_tc_codes(packed, M, K, R, vec), Q)
    assert np.array_equal(O.lut_tc_pack(Q, M, K, R, vec), packed)


@pytest.mark.parametrize("vec,R,K", [(1, r, k) for r in (2, 3, 4, 5, 8) for k in (1024, 1280)] +
                         [(2, r, k) for r in (3, 6, 8, 12) for k in (2048, 2560)] +
                         [(4, r, k) for r in (6, 8, 10, 12) for k in (4096, 5120)])
def test_simt_layout(golden, vec, R, K):
    Q = golden[f"simt_codes_{vec}_{R}_{K}"]
    packed = golden[f"simt_packed_{vec}_{R}_{K}"]
    assert np.array_equal(O.simt_codes(packed, 8, K, R, vec), Q)
    assert np.array_equal(O.simt_pack(Q, 8, K, R, vec), packed)


@pytest.mark.parametrize("vec,R", [(1, 4), (2, 6)])
def test_convert_tc_to_simt(golden, vec, R):
    M, K = 32, 2048
    codes = O.lut_tc_codes(golden[f"conv_tc_{vec}_{R}"], M, K, R, vec)
    assert np.array_equal(O.simt_pack(codes, M, K, R, vec), golden[f"conv_simt_{vec}_{R}"])


def test_had28(golden):
    assert np.array_equal(O.had28(), golden["had28"])
    H = O.had28()
    assert np.array_equal(H @ H.T, 28 * np.eye(28))


@pytest.mark.parametrize("n", [1024, 4096, 14336])
def test_hadamard(golden, n):
    x = golden[f"had_x_{n}"]
    yt = O.hadamard_ref(x, transpose=True)
    for key in ("had_Ut", "had_cuda_T", "had_head_cuda_T"):
        assert np.allclose(yt, golden[f"{key}_{n}"], atol=2e-4), key
    assert np.allclose(O.hadamard_ref(x, transpose=False), golden[f"had_U_{n}"], atol=2e-4)
    assert np.allclose(O.hadamard_ref(O.hadamard_ref(x, True), False), x, atol=1e-5)  # U Ut = I


def test_quant_info(golden):
    for key in golden.files:
        if not key.startswith("qinfo_"):
            continue
        qi = O.get_quant_info(key[len("qinfo_"):])
        kv = qi.get("KV", -1)
        flat = [qi.get("tlut_bits", -1), qi.get("lut_bits", -1), qi.get("vec_sz", -1)]
        flat += list(kv) if isinstance(kv, list) else [kv, -1]
        assert flat == list(golden[key])
