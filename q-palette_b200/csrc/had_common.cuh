// had_common.cuh -- in-shared-memory Walsh-Hadamard transform with the optional 28x28 factor, shared by had_kernels.cu
// and decode_kernels.cu.
#pragma once
#include "qp_common.cuh"

namespace qp {

// The reference's 28x28 Hadamard matrix (lib/utils/matmul_had.py:261 get_had28) is Paley type II for q = 13:
// H = [[S+I, S-I],[S-I, -S-I]], S = bordered Jacobsthal matrix of GF(13) (checked bit-for-bit against get_had28() through
// the golden fixture).  H is symmetric, so H^T = H.
constexpr int kHadThreads = 1024;

// 2^R-point butterflies held in registers, element stride h = 2^lh in shared memory (power of two: no div/mod)
template <int R>
__device__ __forceinline__ void fwht_pass(float *v, int n, int lh) {
    constexpr int P = 1 << R;
    const int h = 1 << lh;
    for (int idx = threadIdx.x; idx < (n >> R); idx += blockDim.x) {
        const int low = idx & (h - 1), hi = idx >> lh;
        float *base = v + ((hi << (R + lh)) | low);
        float r[P];
#pragma unroll
        for (int k = 0; k < P; ++k) r[k] = base[k << lh];
#pragma unroll
        for (int s = 1; s < P; s <<= 1) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
                if ((k & s) == 0) {
                    const float a = r[k], b = r[k | s];
                    r[k] = a + b;
                    r[k | s] = a - b;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < P; ++k) base[k << lh] = r[k];
    }
}

// y = S x for the 14x14 bordered Jacobsthal matrix of GF(13): S[0][0]=0, S[0][j]=S[j][0]=1, S[1+i][1+j]=chi(j-i)
__device__ __forceinline__ void jacobsthal14(const float (&x)[14], float (&y)[14]) {
    constexpr int chi[13] = {0, 1, -1, 1, 1, -1, -1, -1, -1, 1, 1, -1, 1};  // Legendre symbol mod 13 (squares 1,3,4,9,10,12)
    float s0 = 0.f;
#pragma unroll
    for (int j = 1; j < 14; ++j) s0 += x[j];
    y[0] = s0;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        float s = x[0];
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            constexpr int dummy = 0;
            (void)dummy;
            const int c = chi[(j - i + 13) % 13];
            if (c > 0) s += x[1 + j];
            else if (c < 0) s -= x[1 + j];
        }
        y[1 + i] = s;
    }
}

// in-place (hadK^T (x) H_m) on v[n] (shared memory, fp32), n = Kf*m, m = 2^k, Kf in {1, 28}; ends with a barrier.
// The 28x28 factor is applied through its Paley structure H = [[S+I, S-I],[S-I, -S-I]] (~400 flops per column instead
// of 784); it is symmetric, so H^T = H.
__device__ __forceinline__ void hadamard_smem(float *v, int n, int m, int Kf, int lh = 0) {  // lh: log2 of the first stride still to do
    const int lm = 31 - __clz(m);
    while (lh < lm) {
        if (lh + 3 <= lm) {
            fwht_pass<3>(v, n, lh);
            lh += 3;
        } else if (lh + 2 <= lm) {
            fwht_pass<2>(v, n, lh);
            lh += 2;
        } else {
            fwht_pass<1>(v, n, lh);
            lh += 1;
        }
        __syncthreads();
    }
    if (Kf == 28) {
        for (int c = threadIdx.x; c < m; c += blockDim.x) {
            float u[14], w[14], su[14], sw[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                u[j] = v[j * m + c];
                w[j] = v[(14 + j) * m + c];
            }
            jacobsthal14(u, su);
            jacobsthal14(w, sw);
#pragma unroll
            for (int j = 0; j < 14; ++j) {
                v[j * m + c] = (su[j] + u[j]) + (sw[j] - w[j]);          // (S+I)u + (S-I)w
                v[(14 + j) * m + c] = (su[j] - u[j]) - (sw[j] + w[j]);   // (S-I)u - (S+I)w
            }
        }
        __syncthreads();
    }
}

}  // namespace qp
