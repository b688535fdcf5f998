"""The multi-threaded C port used for the CPU baseline (oracle/qp_cref.c) against the numpy oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import qp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cref():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libqp_cref.so"], check=True, capture_output=True)
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libqp_cref.so"))


@pytest.mark.parametrize("KV", range(2, 11))
def test_cref_matches_oracle(cref, KV):
    S = 9 if KV <= 8 else KV + 1
    rng = np.random.default_rng(KV)
    M, K, bs = 96, 160, 3
    buf = rng.integers(0, 256, size=M * K * KV // 16, dtype=np.uint8)
    tlut = (rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)
    x = rng.standard_normal((bs, K)).astype(np.float16)
    out = np.zeros((bs, M), np.float32)
    W = np.zeros((M, K), np.float16)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert cref.qp_cref_tcq(vp(buf), vp(tlut), M, K, KV, S, vp(x), bs, K, 0, 0, M, vp(out), vp(W), K) == 0
    Wref = O.tcq_decode(buf, tlut, M, K, KV, S)
    assert np.array_equal(W.view(np.uint16), Wref.view(np.uint16))
    assert np.allclose(out, O.gemv_ref(Wref, x), rtol=1e-5, atol=1e-4)
