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
