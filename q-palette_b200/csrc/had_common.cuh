// had_common.cuh -- in-shared-memory Walsh-Hadamard transform with the optional 28x28 factor, shared by had_kernels.cu
// and decode_kernels.cu.
#pragma once
#include "qp_common.cuh"

namespace qp {

// the reference's 28x28 Hadamard matrix (lib/utils/matmul_had.py:261 get_had28) is Paley type II for q = 13:
// H = [[S+I, S-I],[S-I, -S-I]], S = bordered Jacobsthal matrix of GF(13).  Built at compile time; row r of the table is
// the sign mask of H[r][:] (bit j set <=> H[r][j] = -1).  H is symmetric, so H^T = H.
struct Had28 {
    uint32_t neg[28];
};
constexpr int legendre13(int x) {
    x = ((x % 13) + 13) % 13;
    if (x == 0) return 0;
    for (int y = 1; y < 13; ++y)
        if ((y * y) % 13 == x) return 1;
    return -1;
}
constexpr int sval(int i, int j) {  // S, 14 x 14
    if (i == 0 && j == 0) return 0;
    if (i == 0 || j == 0) return 1;
    return legendre13((j - 1) - (i - 1));
}
constexpr int had28_entry(int r, int c) {
    const int i = r % 14, j = c % 14;
    const int s = sval(i, j), d = (i == j) ? 1 : 0;
    if (r < 14 && c < 14) return s + d;
    if (r >= 14 && c >= 14) return -s - d;
    return s - d;
}
constexpr Had28 make_had28() {
    Had28 h = {};
    for (int r = 0; r < 28; ++r) {
        uint32_t m = 0;
        for (int c = 0; c < 28; ++c)
            if (had28_entry(r, c) < 0) m |= (1u << c);
        h.neg[r] = m;
    }
    return h;
}
static __constant__ Had28 c_had28 = make_had28();

constexpr int kHadThreads = 512;

template <int R>  // 2^R-point butterfly on registers, stride h in shared memory
__device__ __forceinline__ void fwht_pass(float *v, int n, int h) {
    constexpr int P = 1 << R;
    for (int idx = threadIdx.x; idx < n / P; idx += blockDim.x) {
        const int low = idx % h, hi = idx / h;
        float *base = v + (size_t)hi * P * h + low;
        float r[P];
#pragma unroll
        for (int k = 0; k < P; ++k) r[k] = base[k * h];
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
        for (int k = 0; k < P; ++k) base[k * h] = r[k];
    }
}


// in-place (hadK^T (x) H_m) on v[n] (shared memory, fp32), n = Kf*m, m = 2^k, Kf in {1, 28}; ends with a barrier
__device__ __forceinline__ void hadamard_smem(float *v, int n, int m, int Kf) {
    int h = 1;
    while (h < m) {
        if (h * 8 <= m) {
            fwht_pass<3>(v, n, h);
            h *= 8;
        } else if (h * 4 <= m) {
            fwht_pass<2>(v, n, h);
            h *= 4;
        } else {
            fwht_pass<1>(v, n, h);
            h *= 2;
        }
        __syncthreads();
    }
    if (Kf == 28) {
        for (int c = threadIdx.x; c < m; c += blockDim.x) {
            float col[28];
#pragma unroll
            for (int j = 0; j < 28; ++j) col[j] = v[j * m + c];
#pragma unroll
            for (int i = 0; i < 28; ++i) {
                const uint32_t neg = c_had28.neg[i];  // H^T[i][j] = H[j][i] = H[i][j] (symmetric)
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 28; ++j) s += ((neg >> j) & 1u) ? -col[j] : col[j];
                v[i * m + c] = s;  // column c is private to this thread
            }
        }
        __syncthreads();
    }
}

}  // namespace qp
