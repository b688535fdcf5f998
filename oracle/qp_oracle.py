"""CPU oracle for the Q-Palette quantized-linear decode path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's *format definitions and torch decode
functions* for the hot path.  It is the checker the CUDA path is
compared against.  Nothing in the product path (`q-palette_b200/`) may import it: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the reference's own python
(`/root/reference/lib/...`, CPU) in the build container and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those vectors.
The one third-party piece, Dao-AILab `fast_hadamard_transform` (unpinned master, not under
/root/reference), is pinned through the reference's own pure-torch `matmul_hadU/matmul_hadUt`
(lib/utils/matmul_had.py:68-91) -- see `hadamard_ref`.

Reference citations (relative to /root/reference):
  TCQ codebook            lib/codebook/bitshift.py:71-79 (quantlut_sym), :198-200 (recons)
  TCQ bit-stream packer   lib/codebook/bitshift.py:296-329 (pack_trellis)
  TCQ kernel swizzle      lib/quantizer/tcq_quant.py:47-60 (== comb_quant.py:11-27)
  fragment permutation    lib/algo/ldlq.py:10-13 (_PERMUTE)
  TCQ torch decoder       lib/utils/kernel_decompress.py:5-61 (even KV only)
  VQ/SQ tensor layout     lib/quantizer/quant_op.py:101-162 (pack), :186-244 (decode)
  SIMT layouts            lib/quantizer/pack_op.py:243-335, lib/quantizer/quant_op.py:15-86
  Hadamard                lib/utils/matmul_had.py:10-65,68-147,261
  layer math              lib/linear/incoherent_linear.py:76-108,324-338,486-506
  synthetic layers        lib/utils/mem_op.py:198-307
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# shared geometry
# --------------------------------------------------------------------------------------

# lib/algo/ldlq.py:10-13 -- tile-local index permutation (trellis order p -> row-major index)
PERMUTE = np.arange(256).reshape(2, 8, 2, 4, 2).transpose(1, 3, 2, 0, 4).reshape(-1)
INV_PERMUTE = np.zeros(256, dtype=np.int64)
INV_PERMUTE[PERMUTE] = np.arange(256)


def _nibbles(buf_u8: np.ndarray) -> np.ndarray:
    """little-endian nibble stream of a byte buffer: byte b -> (b & 15), (b >> 4)."""
    buf_u8 = np.ascontiguousarray(buf_u8).reshape(-1)
    out = np.empty(buf_u8.size * 2, dtype=np.uint8)
    out[0::2] = buf_u8 & 15
    out[1::2] = buf_u8 >> 4
    return out


def _from_nibbles(nib: np.ndarray) -> np.ndarray:
    nib = np.ascontiguousarray(nib, dtype=np.uint8).reshape(-1, 2)
    return (nib[:, 0] | (nib[:, 1] << 4)).astype(np.uint8)


def _frag_to_matrix(vals: np.ndarray, M: int, K: int) -> np.ndarray:
    """vals[mh, kh, lane, kl, ml, j, e] (e = the 2 elements of a half2 register) -> (M, K).

    lane g, register j  <->  row = g//4 + 8*(j%2), cols = 2*(g%4) + 8*(j//2) + e   (m16n8k16 A fragment;
    equivalently lib/algo/ldlq.py:10-13).  Flat order [M/32][K/32][32][2][2] per tcq_quant.py:56-58.
    """
    mh, kh = M // 32, K // 32
    a = vals.reshape(mh, kh, 8, 4, 2, 2, 2, 2, 2)  # mh kh g_hi g_lo kl ml j_hi j_lo e
    a = a.transpose(0, 5, 7, 2, 1, 4, 6, 3, 8)  # mh ml j_lo g_hi | kh kl j_hi g_lo e
    return np.ascontiguousarray(a).reshape(M, K)


def _matrix_to_frag(W: np.ndarray) -> np.ndarray:
    """inverse of _frag_to_matrix: (M, K) -> [mh, kh, lane, kl, ml, j, e]."""
    M, K = W.shape
    mh, kh = M // 32, K // 32
    a = W.reshape(mh, 2, 2, 8, kh, 2, 2, 4, 2)  # mh ml j_lo g_hi kh kl j_hi g_lo e
    a = a.transpose(0, 4, 3, 7, 5, 1, 6, 2, 8)  # mh kh g_hi g_lo kl ml j_hi j_lo e
    return np.ascontiguousarray(a).reshape(mh, kh, 32, 2, 2, 4, 2)


# --------------------------------------------------------------------------------------
# TCQ  (bitshift trellis, L = 16, V = 2, KV bits per weight pair)
# --------------------------------------------------------------------------------------

def quantlut_sym(tlut: np.ndarray, S: int, L: int = 16) -> np.ndarray:
    """expanded 2^L-entry LUT; lib/codebook/bitshift.py:71-79.

    t = s*(s+1); index = (t >> (15-S)) & (2^S-1); component 0 negated when bit 15 of t is set.
    `tlut` is (2^S, 2); result has tlut's dtype and shape (2^L, 2).
    """
    assert L == 16 and tlut.shape == (1 << S, 2)
    s = np.arange(1 << L, dtype=np.int64)
    t = (s + 1) * s
    sflp = 1 - ((t >> 15) & 1) * 2
    idx = (t >> (16 - S - 1)) & ((1 << S) - 1)
    lut = tlut[idx].copy()
    lut[:, 0] = lut[:, 0] * sflp.astype(lut.dtype)
    return lut


def tcq_chunks(packed: np.ndarray, M: int, K: int, KV: int) -> np.ndarray:
    """packed int16/uint8 buffer -> per-(lane, tile) stream chunks, uint64 [mh, kh, lane, kl, ml].

    chunk = sum_j nibble[j] << 4j  (tcq_quant.py:52-59 read backwards: nibble split, flips, low nibble first).
    """
    assert M % 32 == 0 and K % 32 == 0
    buf = np.ascontiguousarray(packed).view(np.uint8).reshape(-1)
    assert buf.size * 8 == M * K * KV // 2, (buf.size, M, K, KV)
    nib = _nibbles(buf).reshape(M // 32, K // 32, 32, 2, 2, KV)
    chunk = np.zeros(nib.shape[:-1], dtype=np.uint64)
    for j in range(KV):
        chunk |= nib[..., j].astype(np.uint64) << np.uint64(4 * j)
    return chunk


def _shl64(a: np.ndarray, s: int) -> np.ndarray:
    if s >= 64 or s <= -64:
        return np.zeros_like(a)
    return a << np.uint64(s) if s >= 0 else a >> np.uint64(-s)


def tcq_states(packed: np.ndarray, M: int, K: int, KV: int) -> np.ndarray:
    """-> uint16 states [mh, kh, lane, kl, ml, j]: state p = 4*lane + j of the tile's circular stream,
    i.e. bits [p*KV, p*KV + 16) MSB-first (lib/codebook/bitshift.py:296-329)."""
    c = tcq_chunks(packed, M, K, KV)
    B = 4 * KV
    n1 = np.roll(c, -1, axis=2)  # lane+1 (tail-biting wrap 31 -> 0)
    n2 = np.roll(c, -2, axis=2)
    n3 = np.roll(c, -3, axis=2)
    X = _shl64(c, 64 - B) | _shl64(n1, 64 - 2 * B) | _shl64(n2, 64 - 3 * B) | _shl64(n3, 64 - 4 * B)
    st = np.empty(c.shape + (4,), dtype=np.uint16)
    for j in range(4):
        st[..., j] = ((X >> np.uint64(48 - j * KV)) & np.uint64(0xFFFF)).astype(np.uint16)
    return st


def tcq_states_bitwise(packed: np.ndarray, M: int, K: int, KV: int) -> np.ndarray:
    """slow, literal restatement (bit arrays) used to cross-check tcq_states."""
    c = tcq_chunks(packed, M, K, KV)  # mh kh lane kl ml
    B = 4 * KV
    c = c.transpose(0, 1, 3, 4, 2)  # mh kh kl ml lane
    bits = ((c[..., None] >> np.arange(B - 1, -1, -1, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)
    stream = bits.reshape(c.shape[:-1] + (32 * B,))  # 128*KV bits, circular
    pos = (np.arange(128)[:, None] * KV + np.arange(16)[None, :]) % (128 * KV)
    w = stream[..., pos].astype(np.uint32)  # ... 128 16
    s = (w << np.arange(15, -1, -1, dtype=np.uint32)).sum(-1).astype(np.uint16)  # ... 128
    s = s.reshape(c.shape[:-1] + (32, 4)).transpose(0, 1, 4, 2, 3, 5)  # mh kh lane kl ml j
    return np.ascontiguousarray(s)


def tcq_decode(packed: np.ndarray, tlut: np.ndarray, M: int, K: int, KV: int, S: int) -> np.ndarray:
    """packed trellis -> W (M, K) in tlut's dtype (fp16 for bit-exact parity).

    decoded pair = quantlut_sym(tlut)[state]  (bitshift.py:198-200), placed per the fragment order.
    """
    st = tcq_states(packed, M, K, KV)
    lut = quantlut_sym(np.asarray(tlut), S)
    return _frag_to_matrix(lut[st.astype(np.int64)], M, K)


def tcq_decode_combt(p1, p2, tlut, M, K, KV1, KV2, S, in_part=None):
    """tcomb: split on INPUT columns (lib/linear/comb_linear.py:178-187, 223-270)."""
    k1, k2 = in_part if in_part is not None else (K // 2, K // 2)
    return np.concatenate([tcq_decode(p1, tlut, M, k1, KV1, S), tcq_decode(p2, tlut, M, k2, KV2, S)], axis=1)


def tcq_decode_comb(p1, p2, tlut, M, K, KV1, KV2, S, out_part=None):
    """comb: split on OUTPUT rows (lib/linear/comb_linear.py:35-44, 80-127)."""
    m1, m2 = out_part if out_part is not None else (M // 2, M // 2)
    return np.concatenate([tcq_decode(p1, tlut, m1, K, KV1, S), tcq_decode(p2, tlut, m2, K, KV2, S)], axis=0)


def tcq_pack_trellis(states: np.ndarray, KV: int, L: int = 16) -> np.ndarray:
    """(B, T) tail-biting state sequences -> (B, T*KV/16) uint16 words; bitshift.py:296-329.

    The stream is state[0] (16 bits) followed by the low KV bits of each later state, truncated to T*KV bits.
    """
    Bn, T = states.shape
    st = states.astype(np.int64)
    assert ((st[:, :-1] & ((1 << (L - KV)) - 1)) == (st[:, 1:] >> KV)).all(), "states are not a trellis walk"
    nbits = T * KV + L - KV
    bf = np.zeros((Bn, nbits), dtype=np.uint8)
    bf[:, :L] = (st[:, :1] >> np.arange(L - 1, -1, -1)) & 1
    low = (st[:, 1:, None] >> np.arange(KV - 1, -1, -1)) & 1  # B, T-1, KV
    bf[:, L:] = low.reshape(Bn, -1)
    bf = bf[:, : T * KV]
    assert (T * KV) % 16 == 0
    w = bf.reshape(Bn, -1, 16).astype(np.uint32)
    return (w << np.arange(15, -1, -1, dtype=np.uint32)).sum(-1).astype(np.uint16)


def tcq_swizzle(packed_words: np.ndarray, M: int, K: int, KV: int) -> np.ndarray:
    """pack_trellis output (tiles in row-major [M/16][K/16] order) -> kernel layout; tcq_quant.py:47-60."""
    p8 = np.ascontiguousarray(packed_words).view(np.uint8).reshape(-1, 2)
    p4 = np.stack([p8 & 15, p8 >> 4], axis=-1).reshape(-1, 4)[:, ::-1]
    p4 = p4.reshape(M // 32, 2, K // 32, 2, 32, KV).transpose(0, 2, 4, 3, 1, 5)[..., ::-1]
    out8 = _from_nibbles(np.ascontiguousarray(p4).reshape(-1))
    return out8.view(np.int16).reshape(packed_words.shape)


def tcq_random_walk(rng: np.random.Generator, n_tiles: int, KV: int) -> np.ndarray:
    """random tail-biting state sequences (n_tiles, 128): windows of a random circular bit stream."""
    nb = 128 * KV
    bits = rng.integers(0, 2, size=(n_tiles, nb), dtype=np.uint8)
    pos = (np.arange(128)[:, None] * KV + np.arange(16)[None, :]) % nb
    w = bits[:, pos].astype(np.uint32)
    return (w << np.arange(15, -1, -1, dtype=np.uint32)).sum(-1).astype(np.uint16)


def tcq_pack(states_tiles: np.ndarray, M: int, K: int, KV: int) -> np.ndarray:
    """(M/16*K/16, 128) states (tile row-major, trellis order) -> int16 kernel-layout buffer of shape
    (M/16*K/16, 8*KV), exactly what quantize_layer.py stores in `trellis`."""
    return tcq_swizzle(tcq_pack_trellis(states_tiles, KV), M, K, KV)


def tcq_expected_from_states(states_tiles: np.ndarray, tlut: np.ndarray, M: int, K: int, S: int) -> np.ndarray:
    """W implied by the ORIGINAL states: quantlut_sym[states] re-ordered by _INV_PERMUTE (SURVEY 8c identity 1)."""
    lut = quantlut_sym(np.asarray(tlut), S)
    w = lut[states_tiles.astype(np.int64)].reshape(-1, 256)[:, INV_PERMUTE]
    return w.reshape(M // 16, K // 16, 16, 16).transpose(0, 2, 1, 3).reshape(M, K)


# --------------------------------------------------------------------------------------
# VQ (vec_sz 2) / SQ (vec_sz 1) -- tensor-core ("TC") layout, quant_op.py:101-162
# --------------------------------------------------------------------------------------

def lut_tc_codes(qweight: np.ndarray, M: int, K: int, R: int, vec_sz: int) -> np.ndarray:
    """int32 (M, R*K/32/vec) buffer -> code indices (M, K/vec_sz)."""
    assert vec_sz in (1, 2) and M % 32 == 0 and K % 32 == 0
    buf = np.ascontiguousarray(qweight).view(np.uint8).reshape(-1)
    ncode = 8 // vec_sz  # codes per (lane, tile): 4 half2 registers
    nbits = ncode * R
    assert buf.size * 8 == M * K * R // vec_sz
    nib = _nibbles(buf).reshape(M // 32, K // 32, 32, 2, 2, nbits // 4)
    payload = np.zeros(nib.shape[:-1], dtype=np.uint64)
    for j in range(nbits // 4):
        payload |= nib[..., j].astype(np.uint64) << np.uint64(4 * j)
    codes = np.empty(payload.shape + (ncode,), dtype=np.int64)
    for c in range(ncode):
        codes[..., c] = ((payload >> np.uint64(R * c)) & np.uint64((1 << R) - 1)).astype(np.int64)
    if vec_sz == 2:
        full = np.repeat(codes[..., None], 2, axis=-1)  # j, e share the code
        return _frag_to_matrix(full, M, K)[:, 0::2]
    full = codes.reshape(payload.shape + (4, 2))  # register j holds codes 2j, 2j+1
    return _frag_to_matrix(full, M, K)


def lut_tc_pack(codes: np.ndarray, M: int, K: int, R: int, vec_sz: int) -> np.ndarray:
    """inverse of lut_tc_codes: (M, K/vec) indices -> int32 (M, R*K/32/vec)."""
    codes = np.asarray(codes, dtype=np.int64)
    if vec_sz == 2:
        full = np.repeat(codes, 2, axis=1)
        frag = _matrix_to_frag(full)[..., 0]  # mh kh lane kl ml j
    else:
        frag = _matrix_to_frag(codes).reshape(M // 32, K // 32, 32, 2, 2, 8)
    ncode = frag.shape[-1]
    payload = np.zeros(frag.shape[:-1], dtype=np.uint64)
    for c in range(ncode):
        payload |= frag[..., c].astype(np.uint64) << np.uint64(R * c)
    nn = ncode * R // 4
    nib = np.empty(payload.shape + (nn,), dtype=np.uint8)
    for j in range(nn):
        nib[..., j] = ((payload >> np.uint64(4 * j)) & np.uint64(15)).astype(np.uint8)
    return _from_nibbles(nib.reshape(-1)).view(np.int32).reshape(M, -1)


def lut_tc_decode(qweight, lut, M, K, R, vec_sz):
    """W (M, K) = lut[code]  (lib/codebook/vq_codebook.py:34)."""
    codes = lut_tc_codes(qweight, M, K, R, vec_sz)
    lut = np.asarray(lut).reshape(1 << R, vec_sz)
    return lut[codes].reshape(M, K)


# --------------------------------------------------------------------------------------
# SIMT layouts -- pack_op.py:288-335 (vec 1), quant_op.py:15-78 (vec 2/4)
# --------------------------------------------------------------------------------------

def _simt_plan(K: int, vec_sz: int):
    """yield (chunk_col0, word0_per_bit, eff) for each K chunk; chunk = 32 threads * 32*vec weights."""
    chunk = 32 * 32 * vec_sz
    nfull, rem = divmod(K, chunk)
    plan = [(k * chunk, k * 32, 32) for k in range(nfull)]
    if rem:
        assert rem % (32 * vec_sz) == 0
        plan.append((nfull * chunk, nfull * 32, rem // (32 * vec_sz)))
    return plan


def simt_codes(qweight: np.ndarray, M: int, K: int, bits: int, vec_sz: int) -> np.ndarray:
    """uint32 (M, bits*K/32/vec) SIMT buffer -> code indices (M, K/vec_sz)."""
    q = np.ascontiguousarray(qweight).view(np.uint32).reshape(M, -1)
    assert q.shape[1] == bits * K // 32 // vec_sz
    out = np.zeros((M, K // vec_sz), dtype=np.int64)
    gw = 8 // vec_sz if vec_sz <= 8 else 1  # codes per group of 8 consecutive weights
    for col0, w0, eff in _simt_plan(K, vec_sz):
        base = w0 * bits
        words = q[:, base: base + bits * eff].reshape(M, bits, eff)  # word j of thread t at t + j*eff
        big = np.zeros((M, eff, 32), dtype=np.int64)
        # the thread's 32 codes, LSB-first over the concatenation of its `bits` words
        wbits = ((words[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8)  # M bits eff 32
        stream = wbits.transpose(0, 2, 1, 3).reshape(M, eff, bits * 32)
        cb = stream.reshape(M, eff, 32, bits).astype(np.int64)
        big = (cb << np.arange(bits)).sum(-1)  # M eff 32
        # codes cover groups g = 0..4*vec-1 of 8 consecutive weights starting at (col0/8 + t + g*eff)*8
        ng = 4 * vec_sz
        big = big.reshape(M, eff, ng, gw)
        for g in range(ng):
            for t in range(eff):
                w_start = col0 + (t + g * eff) * 8
                out[:, w_start // vec_sz: w_start // vec_sz + gw] = big[:, t, g, :]
    return out


def simt_pack(codes: np.ndarray, M: int, K: int, bits: int, vec_sz: int) -> np.ndarray:
    """inverse of simt_codes."""
    codes = np.asarray(codes, dtype=np.int64)
    gw = 8 // vec_sz
    ng = 4 * vec_sz
    q = np.zeros((M, bits * K // 32 // vec_sz), dtype=np.uint32)
    for col0, w0, eff in _simt_plan(K, vec_sz):
        big = np.zeros((M, eff, ng, gw), dtype=np.int64)
        for g in range(ng):
            for t in range(eff):
                w_start = col0 + (t + g * eff) * 8
                big[:, t, g, :] = codes[:, w_start // vec_sz: w_start // vec_sz + gw]
        big = big.reshape(M, eff, 32)
        cb = ((big[..., None] >> np.arange(bits)) & 1).astype(np.uint64)  # M eff 32 bits
        stream = cb.reshape(M, eff, bits, 32)  # word j, bit b
        words = (stream << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)  # M eff bits
        base = w0 * bits
        q[:, base: base + bits * eff] = words.transpose(0, 2, 1).reshape(M, bits * eff)
    return q.view(np.int32)


def simt_decode(qweight, lut, M, K, bits, vec_sz):
    codes = simt_codes(qweight, M, K, bits, vec_sz)
    lut = np.asarray(lut).reshape(1 << bits, vec_sz)
    return lut[codes].reshape(M, K)


# --------------------------------------------------------------------------------------
# Hadamard / incoherence
# --------------------------------------------------------------------------------------

def had28() -> np.ndarray:
    """the 28x28 Hadamard factor the reference uses (lib/utils/matmul_had.py:261): Paley type-II for q = 13,
    H = [[S+I, S-I], [S-I, -S-I]] with S the bordered Jacobsthal matrix of GF(13). Checked bit-for-bit against
    get_had28() by tests/golden (had28 fixture)."""
    q = 13
    chi = -np.ones(q, dtype=np.int64)
    for x in range(1, q):
        chi[(x * x) % q] = 1
    chi[0] = 0
    Q = np.array([[chi[(j - i) % q] for j in range(q)] for i in range(q)])
    S = np.zeros((q + 1, q + 1), dtype=np.int64)
    S[0, 1:] = 1
    S[1:, 0] = 1
    S[1:, 1:] = Q
    I = np.eye(q + 1, dtype=np.int64)
    return np.block([[S + I, S - I], [S - I, -S - I]]).astype(np.float32)


def get_hadK(n: int):
    """lib/utils/matmul_had.py:10-65 restricted to the factors this build supports (28 and 1)."""
    if n % 28 == 0 and ((n // 28) & (n // 28 - 1)) == 0:
        return had28(), 28
    assert n & (n - 1) == 0, f"unsupported Hadamard size {n}"
    return None, 1


def fwht(x: np.ndarray) -> np.ndarray:
    """unnormalised Sylvester-ordered Walsh-Hadamard transform along the last axis (== x @ H.T)."""
    x = np.array(x, dtype=np.float64, copy=True)
    n = x.shape[-1]
    h = 1
    while h < n:
        y = x.reshape(x.shape[:-1] + (n // (2 * h), 2, h))
        a = y[..., 0, :] + y[..., 1, :]
        b = y[..., 0, :] - y[..., 1, :]
        x = np.stack([a, b], axis=-2).reshape(x.shape)
        h *= 2
    return x


def hadamard_ref(x: np.ndarray, transpose: bool = True) -> np.ndarray:
    """y = x @ (hadK (x) H_{n/K})^{T or not} / sqrt(n), float64.

    transpose=True is `matmul_hadUt` (matmul_had.py:90-91), which is what every inference call site applies
    (`matmul_hadU_cuda(x, had_left_T, K)` with the pre-transposed factor, incoherent_linear.py:58,82,106,325,336,491).
    """
    n = x.shape[-1]
    hadK, K = get_hadK(n)
    y = x.astype(np.float64).reshape(-1, K, n // K)
    y = fwht(y)
    if K > 1:
        hk = hadK.T if transpose else hadK
        y = np.einsum("ij,bjn->bin", hk.astype(np.float64), y)
    return (y / math.sqrt(n)).reshape(x.shape)


def incoherent_linear_ref(x, W, SU, Wscale, scale):
    """left-only incoherent layer: z = fp16(Ut(x*SU)/s); y = (W z) * Wscale * s.

    incoherent_linear.py:486-506 (s=32) and :76-108, :324-338 (s=64). float64 math with the reference's fp16
    rounding point on z; returns float64 (M,) / (bs, M).
    """
    xs = (np.asarray(x, np.float16) * np.asarray(SU, np.float16)).astype(np.float16)
    z = (hadamard_ref(xs.astype(np.float32)) / scale).astype(np.float16)
    y = z.astype(np.float64) @ np.asarray(W, np.float64).T
    return y * np.asarray(Wscale, np.float64) * scale


# --------------------------------------------------------------------------------------
# quantizer-string grammar and synthetic layers (mem_op.py:198-307)
# --------------------------------------------------------------------------------------

def get_quant_info(qs: str) -> dict:
    if qs.startswith("tcq"):
        _, kv, _h, _s = qs.split("_")
        return dict(quantizer_str=qs, quantizer="tcq_ldlq", KV=int(kv), V=2, tlut_bits=9 if int(kv) <= 8 else int(kv) + 1)
    if qs.startswith("tcomb"):
        _, a, b, ratio, _h, _s = qs.split("_")
        return dict(quantizer_str=qs, quantizer="combt_ldlq", KV=[int(a), int(b)], V=2,
                    tlut_bits=9 if int(b) <= 8 else int(b) + 1, ratio=float(ratio))
    if qs.startswith("ldlq"):
        _, vec, bits, _h, _s = qs.split("_")
        return dict(quantizer_str=qs, quantizer="vq_ldlq", vec_sz=int(vec), lut_bits=int(bits))
    raise ValueError(qs)


def gemv_ref(W: np.ndarray, x: np.ndarray) -> np.ndarray:
    """out (bs, M) = x.float() @ W.float().T in float64 (the torch oracle of SURVEY 8c, widened)."""
    return np.asarray(x, np.float64).reshape(-1, W.shape[1]) @ np.asarray(W, np.float64).T
