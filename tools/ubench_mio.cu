// ubench_mio.cu -- what the SM's load/store (MIO) path sustains for the access patterns of the dequant-GEMV loop:
// conflict-free lane-replicated LDS.32 gathers, 2-way conflicted gathers, SHFL.IDX, per-lane LDG.32 at a 12-byte stride,
// and the loop's own mix.  Prints SM cycles per warp-instruction per SM (all 4 schedulers busy, 24 warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_mio tools/ubench_mio.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 768, kIters = 2000;
extern __shared__ __align__(16) uint8_t smem[];

__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <int MODE>  // 0: 128-byte slots (bank = lane), 1: 64-byte slots (2-way conflicts), 2: same address (broadcast)
__global__ void k_lds(unsigned long long *cyc, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t a[16];
    uint32_t r = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        r = r * 1664525u + 1013904223u;
        const uint32_t e = (r >> 10) & 1023u;
        a[i] = base + (MODE == 0 ? (e << 7) | (lane << 2) : MODE == 1 ? (e << 6) | ((lane & 15) << 2) : (e << 7));
    }
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i;
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= lds(a[i]);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_shfl(unsigned long long *cyc, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * (i + 3);
    const int src = (lane + 1) & 31;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], src) + 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= v[i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

template <int W>  // W = 1: three LDG.32 at a 12-byte lane stride; W = 4: one LDG.128 at a 16-byte lane stride
__global__ void k_ldg(unsigned long long *cyc, uint32_t *sink, const uint32_t *buf, int words) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    const uint32_t *p = buf + (size_t)(blockIdx.x * 24 + warp) * 128 + lane * (W == 1 ? 3 : 4);
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        const uint32_t *q = p + ((it * 4096) & (words - 1));
        if (W == 1) {
            uint32_t a, b, c;
            asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(a) : "l"(q));
            asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1+4];" : "=r"(b) : "l"(q));
            asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1+8];" : "=r"(c) : "l"(q));
            acc ^= a ^ b ^ c;
        } else {
            uint32_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(q));
            acc ^= a ^ b ^ c ^ d;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0xdeadbeef) sink[0] = acc;
}

// the loop's mix per "super-tile": 16 gathers + 1 LDS.128 + 4 SHFL (+ NLDG LDG.32) and NALU dependent-free ALU/FMA pairs
template <int NLDG, int NALU>
__global__ void k_mix(unsigned long long *cyc, uint32_t *sink, const uint32_t *buf, int words) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t a[16];
    uint32_t r = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        r = r * 1664525u + 1013904223u;
        a[i] = base + ((((r >> 10) & 1023u) << 7) | (lane << 2));
    }
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i;
    __syncthreads();
    const uint32_t *p = buf + (size_t)(blockIdx.x * 24 + warp) * 128 + lane * 3;
    const int src = (lane + 1) & 31;
    uint32_t acc = 0, s0 = lane, s1 = lane * 3, s2 = lane * 5, s3 = lane * 7, u = r;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= lds(a[i]);
        uint4 xv;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(xv.x), "=r"(xv.y), "=r"(xv.z), "=r"(xv.w) : "r"(base + (lane & 3) * 16));
        acc ^= xv.x ^ xv.y ^ xv.z ^ xv.w;
        s0 = __shfl_sync(0xffffffffu, s0, src); s1 = __shfl_sync(0xffffffffu, s1, src);
        s2 = __shfl_sync(0xffffffffu, s2, src); s3 = __shfl_sync(0xffffffffu, s3, src);
        if (NLDG) {
            const uint32_t *q = p + ((it * 4096) & (words - 1));
#pragma unroll
            for (int i = 0; i < NLDG; ++i) {
                uint32_t v;
                asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(q + i));
                acc ^= v;
            }
        }
#pragma unroll
        for (int i = 0; i < NALU; ++i) {  // one fma-pipe + one alu-pipe instruction, independent across i
            u = u * 1664525u + (uint32_t)i;
            acc ^= (u >> (i & 15));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    acc ^= s0 ^ s1 ^ s2 ^ s3;
    if (acc == 0xdeadbeef) sink[0] = acc;
}

static double avg_cycles(unsigned long long *d, int n) {
    unsigned long long h[256];
    cudaMemcpy(h, d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; ++i) s += (double)h[i];
    return s / n;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *cyc;
    uint32_t *sink, *buf;
    const int words = 1 << 22;  // 16 MB: L2 resident
    cudaMalloc(&cyc, 256 * 8); cudaMalloc(&sink, 64); cudaMalloc(&buf, (size_t)words * 4 + (1 << 20)); cudaMemset(buf, 1, (size_t)words * 4 + (1 << 20));
    const int smemB = 128 * 1024 + 1024;
    const double warps = kThreads / 32.0;
#define RUN(name, kern, nper, ...)                                                                                    \
    do {                                                                                                              \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB);                               \
        kern<<<sms, kThreads, smemB>>>(__VA_ARGS__);                                                                  \
        kern<<<sms, kThreads, smemB>>>(__VA_ARGS__);                                                                  \
        cudaError_t e = cudaDeviceSynchronize();                                                                      \
        const double c = avg_cycles(cyc, sms);                                                                        \
        printf("%-44s %8.0f cycles/iter-set  -> %6.2f SM-cycles per warp-instruction (%s)\n", name,                  \
               c / kIters, c / kIters / (warps * (nper)), cudaGetErrorString(e));                                     \
    } while (0)
    RUN("LDS.32 gather, 128B slots (conflict-free)", k_lds<0>, 16, cyc, sink);
    RUN("LDS.32 gather, 64B slots (2-way conflicts)", k_lds<1>, 16, cyc, sink);
    RUN("LDS.32 per-warp broadcast address", k_lds<2>, 16, cyc, sink);
    RUN("SHFL.IDX", k_shfl, 8, cyc, sink);
    RUN("LDG.32 x3, 12B lane stride (L2 hits)", k_ldg<1>, 3, cyc, sink, buf, words);
    RUN("LDG.128, 16B lane stride (L2 hits)", k_ldg<4>, 1, cyc, sink, buf, words);
    RUN("mix 16 LDS+LDS.128+4 SHFL            (21 MIO)", (k_mix<0, 0>), 21, cyc, sink, buf, words);
    RUN("mix + 3 LDG                          (24 MIO)", (k_mix<3, 0>), 24, cyc, sink, buf, words);
    RUN("mix + 3 LDG + 16 alu/fma pairs       (24 MIO)", (k_mix<3, 16>), 24, cyc, sink, buf, words);
    RUN("mix + 3 LDG + 32 alu/fma pairs       (24 MIO)", (k_mix<3, 32>), 24, cyc, sink, buf, words);
    RUN("mix + 3 LDG + 40 alu/fma pairs       (24 MIO)", (k_mix<3, 40>), 24, cyc, sink, buf, words);
    return 0;
}
