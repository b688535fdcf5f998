"""The NVLink peer-exchange kernels of the row-sharded decode path (qp_fused_norm_had_xchg, qp_silu_mul_had_grid_xchg)
exercised on ONE device: two "ranks" are two exchange regions on the same GPU whose kernels run concurrently on two streams
and push into each other's region exactly as two processes would over NVLink (same flags, epochs, release/acquire
protocol).  Results are checked against the un-sharded kernels and against the float64 restatement, over several epochs,
so that the driver's single-GPU box covers the exchange code (tests/test_gpu_tp.py needs >= 2 GPUs)."""
import ctypes
import math

import numpy as np
import pytest
import torch

import _restate as R

pytestmark = pytest.mark.gpu


class _Span:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class TwoRegions:
    """two exchange regions on the current device: [buffer (buf_bytes) | flags (nsites * 2 words)]"""

    def __init__(self, buf_bytes, nsites):
        from qpalette._cabi import check, lib
        self.L, self.check = lib(), check
        self.buf_bytes = (buf_bytes + 255) & ~255
        self.nbytes = self.buf_bytes + ((nsites * 2 * 4 + 255) & ~255)
        self.bases = []
        for _ in range(2):
            p = ctypes.c_void_p()
            check(self.L.qp_peer_alloc(ctypes.byref(p), self.nbytes))
            self.bases.append(p.value)
        i64 = dict(dtype=torch.int64, device="cuda")
        self.d_bases = torch.tensor(self.bases, **i64)
        self.d_flags = torch.tensor([b + self.buf_bytes for b in self.bases], **i64)
        self.epochs = [torch.zeros(nsites, dtype=torch.int32, device="cuda") for _ in range(2)]
        check(self.L.qp_set_spin_timeout_ms(5000))  # a protocol bug must fail the test, not hang the box

    def view(self, rank, dtype, count):
        return torch.as_tensor(_Span(self.bases[rank], self.buf_bytes), device="cuda").view(dtype)[:count]

    def xchg(self, rank, site, slice_bytes):
        from qpalette._cabi import Xchg
        return Xchg(self.d_bases.data_ptr(), self.d_flags.data_ptr(), self.epochs[rank].data_ptr(), 0, slice_bytes, rank, 2, site)

    def close(self):
        torch.cuda.synchronize()
        for b in self.bases:
            self.L.qp_peer_free(b)
        self.check(self.L.qp_set_spin_timeout_ms(60000))


@pytest.mark.parametrize("n,sharded", [(4096, "h"), (8192, "acc"), (14336, "h")])
def test_fused_norm_had_self_exchange(n, sharded):
    from qpalette._cabi import check, lib
    L = lib()
    rng = np.random.default_rng(n)
    eps, S = 1e-5, 64.0
    had_scale = 1.0 / (math.sqrt(n) * S)
    su = torch.from_numpy(rng.choice([-1.0, 1.0], n).astype(np.float16)).cuda()
    nw = torch.from_numpy(rng.uniform(0.5, 1.5, n).astype(np.float16)).cuda()
    ws = torch.from_numpy((rng.uniform(0.5, 1.5, n) / 64 / 30).astype(np.float16)).cuda()
    elt = 2 if sharded == "h" else 4
    reg = TwoRegions(n * elt, nsites=4)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    p = lambda t: t.data_ptr() if t is not None else None
    try:
        for epoch in range(3):
            site = epoch % 2  # two sites, one of them used twice: the epoch counters must advance per site
            full_h = rng.standard_normal(n).astype(np.float16)
            full_acc = (rng.standard_normal(n) * 30).astype(np.float32)
            bufs, hs, xs, zeros = [], [], [], []
            for r in range(2):
                if sharded == "h":   # gathered buffer = the fp16 vector (attention output / SiLU*mul activations)
                    b = reg.view(r, torch.float16, n)
                    b.fill_(float("nan"))
                    b[r * n // 2:(r + 1) * n // 2] = torch.from_numpy(full_h[r * n // 2:(r + 1) * n // 2]).cuda()
                    hs.append(b)
                else:                # gathered buffer = the fp32 accumulators (o / down projection outputs), residual in h
                    b = reg.view(r, torch.float32, n)
                    b.fill_(float("nan"))
                    b[r * n // 2:(r + 1) * n // 2] = torch.from_numpy(full_acc[r * n // 2:(r + 1) * n // 2]).cuda()
                    hs.append(torch.from_numpy(full_h.copy()).cuda())
                bufs.append(b)
                xs.append(torch.zeros(n, dtype=torch.float16, device="cuda"))
                zeros.append(torch.ones(777, dtype=torch.float32, device="cuda"))
            torch.cuda.synchronize()
            keep = []
            for r in range(2):
                xc = reg.xchg(r, site, n * elt // 2)
                keep.append(xc)
                with torch.cuda.stream(streams[r]):
                    st = streams[r].cuda_stream
                    if sharded == "h":
                        check(L.qp_fused_norm_had_xchg(p(xs[r]), p(hs[r]), 0, None, None, 0.0, None, 0.0, p(su), n, had_scale, 1,
                                                       p(zeros[r]), 777, ctypes.byref(xc), st))
                    else:
                        check(L.qp_fused_norm_had_xchg(p(xs[r]), p(hs[r]), 1, p(bufs[r]), p(ws), S, p(nw), eps, p(su), n, had_scale,
                                                       1, p(zeros[r]), 777, ctypes.byref(xc), st))
            torch.cuda.synchronize()
            # restatement with the graph's fp16 rounding points
            if sharded == "h":
                ref = R.incoherent_in(full_h, su.cpu().numpy(), S)
            else:
                h2 = R.add16(full_h, R.scaled_acc(full_acc, ws.cpu().numpy(), S))
                ref = R.incoherent_in(R.rmsnorm16(h2, nw.cpu().numpy(), eps), su.cpu().numpy(), S)
            for r in range(2):
                got_buf = bufs[r].cpu().numpy()
                want = full_h if sharded == "h" else full_acc
                assert np.array_equal(got_buf, want), (epoch, r)              # both regions hold the complete buffer
                assert float(zeros[r].abs().sum()) == 0.0
                x = xs[r].float().cpu().numpy()
                assert np.linalg.norm(x - ref.astype(np.float64)) / np.linalg.norm(ref.astype(np.float64)) <= 1e-3, (epoch, r)
                if sharded == "acc":
                    assert np.array_equal(hs[r].cpu().numpy().view(np.uint16), h2.view(np.uint16))
            assert torch.equal(xs[0], xs[1])
            assert int(reg.epochs[0][site]) == int(reg.epochs[1][site]) == epoch // 2 + 1
    finally:
        reg.close()


@pytest.mark.parametrize("I", [28 * 512, 28 * 1024, 8192])
def test_silu_mul_had_grid_self_exchange(I):
    from qpalette._cabi import check, lib
    L = lib()
    rng = np.random.default_rng(I)
    S = 64.0
    had_scale = 1.0 / (math.sqrt(I) * S)
    Il = I // 2
    su = torch.from_numpy(rng.choice([-1.0, 1.0], I).astype(np.float16)).cuda()
    reg = TwoRegions(I * 4, nsites=2)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    syncs = [torch.zeros(4, dtype=torch.int32, device="cuda") for _ in range(2)]
    p = lambda t: t.data_ptr()
    try:
        for epoch in range(3):
            acc = (rng.standard_normal(2 * I) * 3).astype(np.float32)       # [up (I) | gate (I)], full width
            ws = (rng.uniform(0.5, 1.5, 2 * I) / 64).astype(np.float16)
            ug = R.scaled_acc(acc, ws, S)
            ref = R.incoherent_in(R.silu_mul16(ug[:I], ug[I:]), su.cpu().numpy(), S).astype(np.float64)
            xs, keep, loc = [], [], []
            for r in range(2):  # rank r holds rows [r, r+1) * I/2 of up and of gate
                sl = slice(r * Il, (r + 1) * Il)
                a = torch.from_numpy(np.concatenate([acc[:I][sl], acc[I:][sl]])).cuda()
                w = torch.from_numpy(np.concatenate([ws[:I][sl], ws[I:][sl]])).cuda()
                loc.append((a, w))
                xs.append(torch.zeros(I, dtype=torch.float16, device="cuda"))
            zero = [torch.ones(1000, dtype=torch.float32, device="cuda") for _ in range(2)]
            torch.cuda.synchronize()
            for r in range(2):
                xc = reg.xchg(r, 1, 0)
                keep.append(xc)
                with torch.cuda.stream(streams[r]):
                    check(L.qp_silu_mul_had_grid_xchg(p(xs[r]), p(loc[r][0]), p(loc[r][1]), S, p(su), I, had_scale, p(zero[r]), 1000,
                                                      p(syncs[r]), ctypes.byref(xc), streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(2):
                x = xs[r].float().cpu().numpy()
                assert np.linalg.norm(x - ref) / np.linalg.norm(ref) <= 1e-3, (epoch, r)
                assert float(zero[r].abs().sum()) == 0.0
            assert torch.equal(xs[0], xs[1])
    finally:
        reg.close()


@pytest.mark.parametrize("kind", ["acc", "src"])
@pytest.mark.parametrize("K", [4096, 8192])
def test_ll_gather_into_fused_gemv(kind, K):
    """qp_xchg_send_ll + the LL polling of the fused GEMV prologue (qp_xprod.ll): each "rank" publishes its half of the gathered
    vector (fp32 accumulators -> fp16, or the fp16 attention output) into both receive buffers, then runs the consumer GEMV on
    its own receive buffer.  Must equal the same GEMV fed with the complete vector: residual stream bit for bit, x bit for bit,
    outputs within the atomics noise.  (Senders first, then consumers: on ONE device a 148-CTA consumer would starve the other
    rank's sender of an SM; two processes on two GPUs have no such coupling.)"""
    from qpalette import _cabi
    from qpalette._cabi import SPLIT_IN, check, lib
    L = lib()
    M, S, eps, scale = 512, 9, 1e-5, 64.0
    rng = np.random.default_rng(K + (kind == "src"))
    had_scale = 1.0 / (math.sqrt(K) * scale)
    b1 = torch.from_numpy(rng.integers(0, 256, size=M * (K // 2) * 6 // 16, dtype=np.uint8)).cuda()
    b2 = torch.from_numpy(rng.integers(0, 256, size=M * (K // 2) * 7 // 16, dtype=np.uint8)).cuda()
    tl = torch.from_numpy((rng.standard_normal((1 << S, 2)) * 0.9).astype(np.float16)).cuda()
    ws = torch.from_numpy((rng.uniform(0.5, 1.5, K) / 64 / 40).astype(np.float16)).cuda()
    nw = torch.from_numpy(rng.uniform(0.5, 1.5, K).astype(np.float16)).cuda()
    su = torch.from_numpy(rng.choice([-1.0, 1.0], K).astype(np.float16)).cuda()
    reg = TwoRegions(K * 4, nsites=2)
    p = lambda t: t.data_ptr() if t is not None else None
    st = torch.cuda.current_stream().cuda_stream

    def consumer(src, acc, ll, ll_epoch, ll_kind):
        h_out = torch.zeros(K, dtype=torch.float16, device="cuda")
        x_out = torch.zeros(K, dtype=torch.float16, device="cuda")
        out = torch.zeros(M, dtype=torch.float32, device="cuda")
        if kind == "acc":
            xp = _cabi.XProd(p(src), p(h_out), p(acc), p(ws), scale, p(nw), eps, p(su), had_scale, p(x_out), None, 0, None, 0,
                             ll, ll_epoch, ll_kind)
        else:
            xp = _cabi.XProd(p(src), None, None, None, scale, None, eps, p(su), had_scale, p(x_out), None, 0, None, 0,
                             ll, ll_epoch, ll_kind)
        check(L.qp_tcq_gemv_fused(p(out), p(b1), p(b2), ctypes.addressof(xp), p(tl), M, K, S, 6, 7, SPLIT_IN, K // 2, st))
        torch.cuda.synchronize()
        return h_out, x_out, out

    try:
        for epoch in range(3):
            h = torch.from_numpy(rng.standard_normal(K).astype(np.float16)).cuda()
            acc = torch.from_numpy((rng.standard_normal(K) * 40).astype(np.float32)).cuda()
            zero = [torch.ones(555, dtype=torch.float32, device="cuda") for _ in range(2)]
            keep = []
            for r in range(2):
                xc = reg.xchg(r, 1, K * 4 // 2)
                keep.append(xc)
                sl = slice(r * K // 2, (r + 1) * K // 2)
                src = acc[sl].contiguous() if kind == "acc" else h[sl].contiguous()
                keep.append(src)
                check(L.qp_xchg_send_ll(p(src), 1 if kind == "acc" else 0, K // 2, p(zero[r]), 555, ctypes.byref(xc), st))
            torch.cuda.synchronize()
            want = consumer(h, acc if kind == "acc" else None, None, None, 0)
            for r in range(2):
                assert float(zero[r].abs().sum()) == 0.0
                assert int(reg.epochs[r][1]) == epoch + 1
                # a consumer that does not own the complete vector: poison what the LL entries replace
                src = h if kind == "acc" else torch.full_like(h, float("nan"))
                got = consumer(src, None, reg.bases[r], reg.epochs[r].data_ptr() + 4, 1 if kind == "acc" else 2)
                if kind == "acc":
                    assert torch.equal(got[0], want[0])                                   # h' = h + fp16(acc) * Wscale * s
                assert torch.equal(got[1], want[1])                                       # x
                a, b = got[2].double(), want[2].double()
                assert float((a - b).norm() / b.norm()) <= 1e-3
    finally:
        reg.close()
