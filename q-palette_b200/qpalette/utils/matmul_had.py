"""Hadamard helpers with the reference's names (lib/utils/matmul_had.py:10-147), running on libqpalette's fused
sign + FWHT (+ 28x28 factor) kernel.

Supported sizes: n = 2^k and n = 28 * 2^k (every Llama-3 size: 4096, 8192, 1024, 14336, 28672).  The reference's other
literal factors (12, 20, 36, ... 172; Llama-1/2 sizes) are outside this build's scope and raise.
The 28x28 factor is the Paley type-II matrix for q = 13, identical to the reference's literal `get_had28()` (checked
bit-for-bit against a fixture generated from the reference); it is symmetric, so hadK.T == hadK.
"""
import math

import torch

from .. import ops


def had28():
    q = 13
    chi = [-1] * q
    for x in range(1, q):
        chi[(x * x) % q] = 1
    chi[0] = 0
    S = torch.zeros((q + 1, q + 1))
    S[0, 1:] = 1
    S[1:, 0] = 1
    for i in range(q):
        for j in range(q):
            S[1 + i, 1 + j] = chi[(j - i) % q]
    eye = torch.eye(q + 1)
    return torch.cat([torch.cat([S + eye, S - eye], 1), torch.cat([S - eye, -S - eye], 1)], 0)


def is_pow2(n):
    return n > 0 and (n & (n - 1)) == 0


def get_hadK(n, transpose=False):
    """(hadK, K) as in the reference: K = 28 when n = 28 * 2^k, else K = 1 (hadK None)."""
    if n % 28 == 0 and is_pow2(n // 28):
        h = had28()
        return (h.T.contiguous() if transpose else h), 28
    if is_pow2(n):
        return None, 1
    raise NotImplementedError(f"Hadamard size {n}: only 2^k and 28*2^k are built (Llama-3 sizes)")


def _check_hadK(hadK, K):
    if K == 1:
        return
    if K != 28:
        raise NotImplementedError(f"hadK factor {K} is not built")
    # the kernel applies the built-in (symmetric) 28x28 factor; accept hadK or hadK.T of it, reject anything else
    if hadK is not None and hadK.shape == (28, 28):
        ref = had28().to(hadK.device, hadK.dtype)
        if not torch.equal(hadK, ref):
            raise ValueError("hadK differs from the built-in 28x28 Hadamard factor")


def matmul_hadU_cuda(X, hadK, K, part=1, transpose=False):
    """y = (hadK (x) H_{n/K}) X / sqrt(n) along the last dim (lib/utils/matmul_had.py:137-147)."""
    assert part == 1, "part > 1 is unused by the decode path"
    n = X.shape[-1]
    _check_hadK(hadK, K)
    return ops.hadamard(X.contiguous(), None, 1.0 / math.sqrt(n))


def matmul_hadUt_cuda(X, hadK, K):
    return matmul_hadU_cuda(X, hadK, K, transpose=True)


def matmul_hadU_head_cuda(X, hadK, K, head_dim, transpose=False):
    """block-diagonal variant: the transform acts on consecutive blocks of `head_dim` (matmul_had.py:94-106);
    computes in fp32 and returns X.dtype like the reference."""
    n = X.shape[-1]
    _check_hadK(hadK, K)
    x = X.reshape(-1, head_dim).float().contiguous()
    y = ops.hadamard(x, None, 1.0 / math.sqrt(head_dim))
    return y.reshape(X.shape).to(X.dtype)


def matmul_hadUt_head_cuda(X, hadK, K, head_dim):
    return matmul_hadU_head_cuda(X, hadK, K, head_dim, transpose=True)
