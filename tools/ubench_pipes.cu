// ubench_pipes.cu -- issue cost of the instruction classes of the dequant-GEMV loop on one SM (24 warps, all schedulers):
// legacy mma.sync HMMA.16816.F32, IMAD, SHF/LOP3 (alu pipe), and alternating alu/fma.  SM cycles per warp-instruction per
// scheduler (SMSP).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kThreads = 768, kIters = 4000;

__global__ void k_hmma(unsigned long long *cyc, float *sink) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 1234.5f) sink[0] = s;
}

template <int MODE>  // 0: IMAD, 1: SHF (funnel), 2: LOP3, 3: alternating SHF / IMAD, 4: IADD3
__global__ void k_alu(unsigned long long *cyc, uint32_t *sink) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * (2 * i + 3) + 1;
    uint32_t k = threadIdx.x | 1;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) v[i] = v[i] * v[i] + k;
            else if (MODE == 1) v[i] = __funnelshift_r(v[i], k, 7);
            else if (MODE == 2) v[i] = (v[i] & 0x1ff80u) | k;
            else if (MODE == 3) v[i] = (i & 1) ? v[i] * v[i] + k : __funnelshift_r(v[i], k, 7);
            else v[i] = v[i] + k + (uint32_t)it;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= v[i];
    if (s == 0xdeadbeef) sink[0] = s;
}

static double avg(unsigned long long *d, int n) {
    unsigned long long h[256];
    cudaMemcpy(h, d, n * 8, cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; ++i) s += (double)h[i];
    return s / n;
}
int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, 2048); cudaMalloc(&sink, 64);
    const double warps_per_smsp = kThreads / 32.0 / 4.0;
#define RUN(name, launch, nper)                                                                                     \
    do {                                                                                                            \
        launch; launch;                                                                                             \
        cudaError_t e = cudaDeviceSynchronize();                                                                    \
        const double c = avg(cyc, sms) / kIters;                                                                    \
        printf("%-34s %8.1f cycles / iteration -> %5.2f cycles per warp-instruction per SMSP (%s)\n", name, c,     \
               c / (warps_per_smsp * (nper)), cudaGetErrorString(e));                                               \
    } while (0)
    RUN("HMMA.16816.F32 (mma.sync)", (k_hmma<<<sms, kThreads>>>(cyc, (float *)sink)), 8);
    RUN("IMAD", (k_alu<0><<<sms, kThreads>>>(cyc, sink)), 16);
    RUN("SHF (funnel)", (k_alu<1><<<sms, kThreads>>>(cyc, sink)), 16);
    RUN("LOP3", (k_alu<2><<<sms, kThreads>>>(cyc, sink)), 16);
    RUN("SHF / IMAD alternating", (k_alu<3><<<sms, kThreads>>>(cyc, sink)), 16);
    RUN("IADD3", (k_alu<4><<<sms, kThreads>>>(cyc, sink)), 16);
    return 0;
}
