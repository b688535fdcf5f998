"""The kernels' integer decode logic (csrc/tcq_bits.cuh, lut_bits.cuh), compiled for the host with the warp shuffle
emulated, against the oracle.  This is the CPU-side guard for the bit extraction every CUDA kernel runs."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import qp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libqp_emul.so"], check=True,
                   capture_output=True)
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libqp_emul.so"))


@pytest.mark.parametrize("KV", range(2, 11))
def test_tcq_state_extraction(emul, KV):
    M, K = 64, 96
    rng = np.random.default_rng(KV)
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    out = np.zeros((M // 32) * (K // 32) * 32 * 16, dtype=np.uint16)
    assert emul.qp_emul_tcq_states(buf.ctypes.data_as(ctypes.c_void_p), M, K, KV,
                                   out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(out, O.tcq_states(buf, M, K, KV).reshape(-1))
    assert emul.qp_emul_tcq_max_read(KV) <= 0  # the per-lane word loads never leave the super-tile


@pytest.mark.parametrize("vec,R", [(1, r) for r in range(2, 9)] + [(2, r) for r in range(2, 13)])
def test_lut_code_extraction(emul, vec, R):
    M, K = 64, 96
    E = R if vec == 2 else 2 * R
    rng = np.random.default_rng(3 * R + vec)
    buf = rng.integers(0, 256, size=M * K * E // 16, dtype=np.uint8)
    out = np.zeros((M // 32) * (K // 32) * 32 * 16, dtype=np.uint32)
    assert emul.qp_emul_lut_pairs(buf.ctypes.data_as(ctypes.c_void_p), M, K, E,
                                  out.ctypes.data_as(ctypes.c_void_p)) == 0
    codes = O.lut_tc_codes(buf.view(np.int32), M, K, R, vec)
    if vec == 2:
        frag = O._matrix_to_frag(np.repeat(codes, 2, axis=1))[..., 0]
    else:
        f = O._matrix_to_frag(codes)
        frag = f[..., 0] | (f[..., 1] << R)
    assert np.array_equal(out.reshape(frag.shape), frag)


@pytest.mark.parametrize("T", [0, 1, 5, 147, 148, 3551, 3552, 3553, 16384, 24576, 28672, 57344, 114688, 1000003])
def test_run_split_tiles_the_range(emul, T):
    """csrc/run_split.cuh: every (even, flipped or skewed) two-level split of T super-tiles over 148 CTAs x 24 warps tiles [0, T)
    in (CTA, warp) order, the CTAs of a class carry the same load +-1, no warp more than its CTA's share / 24 rounded up, and a
    skewed split gives the last CTAs the requested smaller share"""
    lo_, hi_, late_, wm = (ctypes.c_uint(0) for _ in range(4))
    for flip in (0, 1):
        for late, pm in ((0, 0), (7, 890), (19, 900), (10, 960), (147, 500), (1, 0), (148, 900), (7, 1000), (30, 10)):
            rc = emul.qp_emul_split_cover(ctypes.c_long(T), 148, 24, late, pm, flip, ctypes.byref(lo_), ctypes.byref(hi_),
                                          ctypes.byref(late_), ctypes.byref(wm))
            assert rc == 0, (late, pm, flip)
            assert hi_.value - lo_.value <= 1, (late, pm, flip)
            assert wm.value <= -(-max(hi_.value, late_.value) // 24), (late, pm, flip)
            if 0 < late < 148 and pm < 1000 and T >= 57344:
                assert abs(late_.value / max(hi_.value, 1) - pm / 1000) < 0.02, (late, pm, hi_.value, late_.value)
