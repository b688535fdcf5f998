"""qp_rope_attention at head size 128 (the split-KV kernel: one CTA per (query head, 128 positions), partial softmaxes combined
by the last CTA of a head) against the float64 restatement of the reference's attention (IncoherentSdpaAttention.forward,
lib/linear/incoherent_linear.py:110-203: Wscale epilogue -> RoPE -> cache append -> softmax(q K^T / sqrt(D)) V), at positions on
both sides of every split boundary and deep into a long context.  The small-model tests of test_gpu_decode.py have head size 64
and run the one-CTA-per-head kernel."""
import numpy as np
import pytest
import torch

import _restate as R

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("qvk_order", [0, 1])
@pytest.mark.parametrize("max_seq,positions", [(100, [0, 1, 63, 99]), (640, [0, 127, 128, 129, 255, 256, 300, 639]),
                                               (4160, [2048, 4100])])
def test_rope_attention_d128(max_seq, positions, qvk_order):
    from qpalette._cabi import check, lib
    L = lib()
    H, Hkv, D, S = 8, 2, 128, 64.0
    rng = np.random.default_rng(max_seq + qvk_order)
    n = (H + 2 * Hkv) * D
    inv_freq = (1.0 / (500000.0 ** (np.arange(0, D, 2) / D))).astype(np.float32)
    ws = (rng.uniform(0.5, 1.5, n) / 64 / 20).astype(np.float16)
    kc = (rng.standard_normal((max_seq, Hkv, D)) * 0.5).astype(np.float16)
    vc = rng.standard_normal((max_seq, Hkv, D)).astype(np.float16)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    kc_d, vc_d, ws_d, inv_d = d(kc), d(vc), d(ws), d(inv_freq)
    nb = int(L.qp_rope_attention_scratch_bytes(H, D, max_seq))
    assert (nb > 0) == (max_seq > 128)
    scratch = torch.zeros(max(nb, 16), dtype=torch.uint8, device="cuda")
    out = torch.zeros(H * D, dtype=torch.float16, device="cuda")
    pos_d = torch.zeros(1, dtype=torch.int32, device="cuda")
    zero = torch.ones(300, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for pos in positions:
        acc = (rng.standard_normal(n) * 20).astype(np.float32)
        pos_d.fill_(pos)
        zero.fill_(1.0)
        check(L.qp_rope_attention(out.data_ptr(), d(acc).data_ptr(), ws_d.data_ptr(), S, inv_d.data_ptr(), kc_d.data_ptr(),
                                  vc_d.data_ptr(), pos_d.data_ptr(), H, Hkv, D, max_seq, qvk_order, zero.data_ptr(), 300,
                                  scratch.data_ptr() if nb else None, st))
        torch.cuda.synchronize()
        qkv = R.scaled_acc(acc, ws, S)
        q = qkv[:H * D].reshape(H, D)
        rest = qkv[H * D:].reshape(2, Hkv, D)
        k_new, v_new = (rest[1], rest[0]) if qvk_order else (rest[0], rest[1])
        qr, kr = R.rope16(q, pos, inv_freq), R.rope16(k_new, pos, inv_freq)
        kc[pos], vc[pos] = kr, R.h16(v_new)
        ref = R.f64(R.attend(qr, kc[:pos + 1], vc[:pos + 1], H // Hkv)).reshape(-1)
        got = out.float().cpu().numpy().astype(np.float64)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= TOL, pos
        # the cache rows of this position were appended (bit-exact), the accumulator-clearing duty done, tickets back at zero
        assert np.array_equal(kc_d[pos].cpu().numpy().view(np.uint16), kc[pos].view(np.uint16))
        assert np.array_equal(vc_d[pos].cpu().numpy().view(np.uint16), vc[pos].view(np.uint16))
        assert float(zero.abs().sum()) == 0.0
        if nb:
            assert int(scratch[:H * 4].view(torch.int32).abs().sum()) == 0
