// gemv_common.cuh -- pieces shared by the fused dequant-GEMV kernels (TCQ and LUT, tensor-core packed layout).
#pragma once
#include "qp_common.cuh"
#include "tcq_bits.cuh"

namespace qp {

// per-lane fetch of one super-tile payload (2*E bytes, lane-contiguous) with the widest load the alignment allows
template <int E>
__device__ __forceinline__ void pack_load_raw(uint32_t (&raw)[TcqGeom<E>::kRawWords], const uint32_t *p) {
    constexpr int NW = TcqGeom<E>::kRawWords;
    constexpr int LB = TcqGeom<E>::kLaneBytes;
    if constexpr (LB % 16 == 0) {
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            const uint4 v = ldg_stream_u128(p + 4 * i);
            raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
    } else if constexpr (LB % 8 == 0) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
            const uint2 v = ldg_stream_u64(p + 2 * i);
            raw[2 * i] = v.x; raw[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) raw[i] = ldg_stream_u32(p + i);
    }
}

// stage x (bs, K) fp16 into shared memory in B-fragment order: 16 bytes per (super-tile column, n, q):
// {x[n][k0+2q..+1], x[n][k0+8+2q..+1], x[n][k0+16+2q..+1], x[n][k0+24+2q..+1]},  k0 = 32*kh
__device__ __forceinline__ void stage_x(uint32_t *xs, const uint32_t *__restrict__ x32, int K, int bs) {
    const int total = (K / 32) * bs * 16;
    const int kw = K / 2;
    for (int d = threadIdx.x; d < total; d += blockDim.x) {
        const int b = d & 1, kl = (d >> 1) & 1, q = (d >> 2) & 3;
        const int r = d >> 4;
        const int n = r % bs, kh = r / bs;
        xs[d] = __ldg(x32 + (size_t)n * kw + 16 * kh + 8 * kl + 4 * b + q);
    }
}


// add one 32-row strip of partial sums to out (bs, M): acc[ml] is the C fragment of rows row0 + 16*ml + {lane/4, +8},
// batch columns 2*(lane%4), +1.  Zeroes the accumulators.
__device__ __forceinline__ void gemv_flush(float *__restrict__ out, int M, int bs, int row0, int lane,
                                           float (&acc)[2][4]) {
    const int r = row0 + (lane >> 2);
    const int c0 = 2 * (lane & 3), c1 = c0 + 1;
#pragma unroll
    for (int ml = 0; ml < 2; ++ml) {
        const int r0 = r + ml * 16;
        if (c0 < bs) {
            atomicAdd(out + (size_t)c0 * M + r0, acc[ml][0]);
            atomicAdd(out + (size_t)c0 * M + r0 + 8, acc[ml][2]);
        }
        if (c1 < bs) {
            atomicAdd(out + (size_t)c1 * M + r0, acc[ml][1]);
            atomicAdd(out + (size_t)c1 * M + r0 + 8, acc[ml][3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[ml][j] = 0.f;
    }
}


struct PackSegment {
    const uint32_t *codes;  // packed words of this part
    int strips;             // rows / 32
    int ksuper;             // cols / 32
    int row0;               // first output row
    int ksuper0;            // first super-tile column of x
};

constexpr int kGemvThreads = 512;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvDepth = 4;  // super-tiles prefetched ahead per warp (register staged)

// issue the first kGemvDepth payload loads of this warp's run [lo, hi)
template <int E>
__device__ __forceinline__ void gemv_prefetch(const PackSegment seg, long lo, long hi,
                                              uint32_t (&raw)[kGemvDepth][TcqGeom<E>::kRawWords]) {
    using G = TcqGeom<E>;
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const uint32_t *lane_base = seg.codes + word0;
#pragma unroll
    for (int d = 0; d < kGemvDepth; ++d) {
#pragma unroll
        for (int i = 0; i < G::kRawWords; ++i) raw[d][i] = 0u;
        if (lo + d < hi) pack_load_raw<E>(raw[d], lane_base + (lo + d) * (long)(G::kSuperBytes / 4));
    }
}

// stream the warp's run of super-tiles [lo, hi) of one part: decode (Dec) -> A fragments -> mma with x (B fragments from
// shared memory) -> fp32 atomics per finished 32-row strip.
//   Dec::kE                      bits per weight pair (payload geometry TcqGeom<kE>)
//   Dec::decode(raw, bitoff, lane, tab_addr_lane, frag)   16 half2 registers of the (lane, super-tile)
template <class Dec>
__device__ __forceinline__ void gemv_run_segment(const PackSegment seg, float *__restrict__ out, int M, int bs,
                                                 uint32_t xs_addr, uint32_t tab_addr_lane, long lo, long hi,
                                                 uint32_t (&raw)[kGemvDepth][TcqGeom<Dec::kE>::kRawWords]) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const uint32_t *lane_base = seg.codes + word0;
    constexpr long kSuperWords = G::kSuperBytes / 4;
    const int n = lane >> 2, q = lane & 3;
    const bool xvalid = n < bs;

    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int mh = (int)(lo / seg.ksuper);
    int kh = (int)(lo - (long)mh * seg.ksuper);

    for (long base = lo; base < hi; base += kGemvDepth) {
#pragma unroll
        for (int d = 0; d < kGemvDepth; ++d) {
            const long it = base + d;
            if (it < hi) {
                uint32_t cur[G::kRawWords];
#pragma unroll
                for (int i = 0; i < G::kRawWords; ++i) cur[i] = raw[d][i];
                if (it + kGemvDepth < hi) pack_load_raw<E>(raw[d], lane_base + (it + kGemvDepth) * kSuperWords);

                // x fragment of this super-tile column: {b0,b1 of k-tile 0, b0,b1 of k-tile 1}
                uint4 xb = make_uint4(0u, 0u, 0u, 0u);
                if (xvalid) xb = lds_u128(xs_addr + (uint32_t)((((seg.ksuper0 + kh) * bs + n) * 4 + q) * 16));

                uint32_t frag[4][4];  // [tile = kl*2+ml][register]
                Dec::decode(cur, bitoff, lane, tab_addr_lane, frag);
                mma_16816(acc[0], frag[0][0], frag[0][1], frag[0][2], frag[0][3], xb.x, xb.y);
                mma_16816(acc[1], frag[1][0], frag[1][1], frag[1][2], frag[1][3], xb.x, xb.y);
                mma_16816(acc[0], frag[2][0], frag[2][1], frag[2][2], frag[2][3], xb.z, xb.w);
                mma_16816(acc[1], frag[3][0], frag[3][1], frag[3][2], frag[3][3], xb.z, xb.w);

                if (++kh == seg.ksuper) {
                    gemv_flush(out, M, bs, seg.row0 + mh * 32, lane, acc);
                    kh = 0;
                    ++mh;
                }
            }
        }
    }
    if (kh != 0) gemv_flush(out, M, bs, seg.row0 + mh * 32, lane, acc);
}

// decode the warp's share of one part and write fp16 W (M, K) row-major
template <class Dec>
__device__ __forceinline__ void dequant_run_segment(const PackSegment seg, __half *__restrict__ W, int K,
                                                    uint32_t tab_addr_lane, int gwarp, int nwarps) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const uint32_t *lane_base = seg.codes + word0;
    const long T = (long)seg.strips * seg.ksuper;
    const long lo = T * gwarp / nwarps, hi = T * (gwarp + 1) / nwarps;
    uint32_t *W32 = reinterpret_cast<uint32_t *>(W);
    const int kw = K / 2;
    for (long it = lo; it < hi; ++it) {
        uint32_t raw[G::kRawWords];
        pack_load_raw<E>(raw, lane_base + it * (long)(G::kSuperBytes / 4));
        uint32_t frag[4][4];
        Dec::decode(raw, bitoff, lane, tab_addr_lane, frag);
        const int mh = (int)(it / seg.ksuper), kh = (int)(it - (long)mh * seg.ksuper);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int kl = t >> 1, ml = t & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = seg.row0 + 32 * mh + 16 * ml + (lane >> 2) + 8 * (j & 1);
                const int col = 32 * (seg.ksuper0 + kh) + 16 * kl + 2 * (lane & 3) + 8 * (j >> 1);
                W32[(size_t)row * kw + (col >> 1)] = frag[t][j];
            }
        }
    }
}

inline int check_align(const void *p, size_t a, const char *name) {
    if (((uintptr_t)p) % a != 0) return fail(QP_ERR_ALIGN, "%s must be %zu-byte aligned", name, a);
    return QP_OK;
}

}  // namespace qp
