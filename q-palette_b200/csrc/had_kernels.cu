// had_kernels.cu -- incoherence processing: random-sign + fast Walsh-Hadamard (+ 28x28 factor) transform, and the
// per-row scale epilogue.  Replaces hadamard::hadamard (Dao-AILab fast_hadamard_transform, called at
// lib/utils/matmul_had.py:134) together with the dense `hadK @ .` of matmul_hadU_cuda / matmul_hadU_head_cuda
// (lib/utils/matmul_had.py:94-106,137-147) and the elementwise glue around them
// (lib/linear/incoherent_linear.py:82-108,325-338,491-503).
//
//   y = ((hadK^T (x) H_m) (x * su)) * scale,   n = Kf * m, m = 2^k, Kf in {1, 28}
// One CTA per row: the row lives in shared memory as fp32, log2(m) butterfly stages run as radix-8 register passes,
// then (Kf = 28) every column gets the 28x28 +-1 factor.  The data volume is tiny (<= 112 KiB per row); the kernel is
// latency bound and exists to fuse ~6 framework kernels into one.
#include "had_common.cuh"

namespace qp {

template <bool XF32, bool YF32>
__global__ void __launch_bounds__(kHadThreads, 1)
hadamard_kernel(void *__restrict__ y, const void *__restrict__ x, const __half *__restrict__ su, int n, int m, int Kf,
                float scale) {
    extern __shared__ __align__(16) float v[];
    const size_t row = blockIdx.x;
    pdl_wait();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float f = XF32 ? reinterpret_cast<const float *>(x)[row * n + i]
                       : __half2float(reinterpret_cast<const __half *>(x)[row * n + i]);
        if (su != nullptr) {
            // the reference multiplies in the activation dtype (x * SU with SU = +-1): exact in either precision
            f *= __half2float(su[i]);
        }
        v[i] = f;
    }
    __syncthreads();
    pdl_launch_dependents();
    hadamard_smem(v, n, m, Kf);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float f = v[i] * scale;
        if (YF32) reinterpret_cast<float *>(y)[row * n + i] = f;
        else reinterpret_cast<__half *>(y)[row * n + i] = __float2half(f);
    }
}

// out = fp16(fp16(fp16(acc) * wscale) * scale)  [rounding points of `linear(x) * Wscale * scale` in fp16],
// optional merged up|gate epilogue: out[i] = silu(y[I + i]) * y[i]
__global__ void scale_epilogue_kernel(__half *__restrict__ out, const float *__restrict__ acc,
                                      const __half *__restrict__ wscale, int bs, int M, float scale, int epilogue) {
    pdl_wait();
    const __half hs = __float2half(scale);
    if (epilogue == QP_EPI_NONE) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)bs * M; i += (size_t)gridDim.x * blockDim.x) {
            const int c = (int)(i % M);
            out[i] = __hmul(__hmul(__float2half(acc[i]), wscale[c]), hs);
        }
    } else {
        const int I = M / 2;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)bs * I; i += (size_t)gridDim.x * blockDim.x) {
            const int b = (int)(i / I), c = (int)(i % I);
            const __half up = __hmul(__hmul(__float2half(acc[(size_t)b * M + c]), wscale[c]), hs);
            const __half gate = __hmul(__hmul(__float2half(acc[(size_t)b * M + I + c]), wscale[I + c]), hs);
            const float g = __half2float(gate);
            const __half act = __float2half(g / (1.f + __expf(-g)));
            out[i] = __hmul(act, up);
        }
    }
}

}  // namespace qp

using namespace qp;

extern "C" int qp_hadamard(void *y, const void *x, const void *su_f16, int rows, int n, float scale, int x_is_f32,
                           int y_is_f32, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(y && x, "NULL pointer argument");
    QP_CHECK_ARG(rows >= 1 && n >= 2, "bad shape rows=%d n=%d", rows, n);
    int Kf = 1, m = n;
    if ((n & (n - 1)) != 0) {
        QP_CHECK_ARG(n % 28 == 0 && (((n / 28) & (n / 28 - 1)) == 0),
                     "Hadamard size %d is neither 2^k nor 28*2^k (other hadK factors are not built)", n);
        Kf = 28;
        m = n / 28;
    }
    const size_t smem = (size_t)n * sizeof(float);
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem, "n = %d too large (row must fit in shared memory)", n);
    void (*kern)(void *, const void *, const __half *, int, int, int, float);
    if (x_is_f32) kern = y_is_f32 ? hadamard_kernel<true, true> : hadamard_kernel<true, false>;
    else kern = y_is_f32 ? hadamard_kernel<false, true> : hadamard_kernel<false, false>;
    static DeviceOnce configured[4];
    const int ki = (x_is_f32 ? 2 : 0) + (y_is_f32 ? 1 : 0);
    if (configured[ki].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    }
    QP_CUDA(launch_pdl(kern, dim3(rows), dim3(kHadThreads), smem, st, y, x, (const __half *)su_f16, n, m, Kf, scale));
    return check_launch("hadamard");
}

extern "C" int qp_scale_epilogue(void *out_f16, const float *acc, const void *wscale_f16, int bs, int M, float scale,
                                 int epilogue, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out_f16 && acc && wscale_f16, "NULL pointer argument");
    QP_CHECK_ARG(epilogue == QP_EPI_NONE || (epilogue == QP_EPI_SILU_MUL && M % 2 == 0), "bad epilogue");
    const int n = bs * M;
    int blocks = (n + 255) / 256;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    QP_CUDA(launch_pdl(scale_epilogue_kernel, dim3(blocks), dim3(256), 0, st, (__half *)out_f16, acc,
                       (const __half *)wscale_f16, bs, M, scale, epilogue));
    return check_launch("scale_epilogue");
}
