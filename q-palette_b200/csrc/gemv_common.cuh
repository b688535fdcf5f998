// gemv_common.cuh -- pieces shared by the fused dequant-GEMV kernels (TCQ and LUT, tensor-core packed layout).
#pragma once
#include "qp_common.cuh"
#include "tcq_bits.cuh"

namespace qp {

// start of the dynamic shared memory: the GEMV / dequantise kernels keep their lane-replicated codebook there, so a lookup
// address is (uniform base) + ((slot offset) | (lane column)) -- one LOP3 and an LDS with a uniform-register base, no
// per-lookup pointer add.  `tab_lane` below is that lane column: (lane & lane_mask) * 4.
extern __shared__ __align__(16) uint8_t qp_dyn_smem[];


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

// cooperative global -> shared copy of n 32-bit words; every thread issues all of its loads before its stores so the
// L2 round trips overlap instead of serialising (this prologue sits on the critical path of a ~5 us kernel)
__device__ __forceinline__ void coop_copy_words(uint32_t *dst, const uint32_t *__restrict__ src, int n) {
    constexpr int U = 8;
    for (int base = threadIdx.x; base < n; base += U * blockDim.x) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x;
            v[u] = (i < n) ? __ldg(src + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x;
            if (i < n) dst[i] = v[u];
        }
    }
}

// stage x (bs, K) fp16 into shared memory in B-fragment order: 16 bytes per (super-tile column kh, batch row n, q):
// {x[n][k0+2q..+1], x[n][k0+8+2q..+1], x[n][k0+16+2q..+1], x[n][k0+24+2q..+1]},  k0 = 32*kh.
// Each thread pulls up to 4 x 16 bytes into registers with all loads in flight together (one L2 round trip per round:
// a single round for bs*K <= 24576), then scatters the words.
__device__ __forceinline__ void stage_x(uint32_t *xs, const uint32_t *__restrict__ x32, int K, int bs) {
    const int kq = K / 8;            // uint4 per batch row
    const int total = bs * kq;       // uint4 to move
    const uint4 *x4 = reinterpret_cast<const uint4 *>(x32);
    constexpr int U = 4;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x;
            v[u] = (i < total) ? __ldg(x4 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x;
            if (i < total) {
                int n = 0, j = i;    // j = uint4 index inside row n
                if (bs != 1) {
                    n = i / kq;
                    j = i - n * kq;
                }
                const int kh = j >> 2, part = j & 3;      // part = which 4-word group of the 16-word column block
                const int kl = part >> 1, b = part & 1;
                uint32_t *d = xs + ((kh * bs + n) * 16 + kl * 2 + b);
                d[0] = v[u].x;   // q = 0
                d[4] = v[u].y;   // q = 1
                d[8] = v[u].z;   // q = 2
                d[12] = v[u].w;  // q = 3
            }
        }
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


// warp index / grid-wide warp id as provably warp-uniform values (lets ptxas keep run bounds in uniform registers and
// drop the convergence barriers around the shuffles)
__device__ __forceinline__ int warp_in_cta() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

struct PackSegment {
    const uint32_t *codes;  // packed words of this part
    int strips;             // rows / 32
    int ksuper;             // cols / 32
    int row0;               // first output row
    int ksuper0;            // first super-tile column of x
};

#ifndef QP_GEMV_THREADS
#define QP_GEMV_THREADS 768
#endif
#ifndef QP_GEMV_DEPTH
#define QP_GEMV_DEPTH 3
#endif
#ifndef QP_GEMV_CTAS
#define QP_GEMV_CTAS 1
#endif
constexpr int kGemvCtasPerSM = QP_GEMV_CTAS;    // experiments: 2 CTAs of 384 threads per SM (needs a <= 64 KiB codebook)
constexpr int kGemvThreads = QP_GEMV_THREADS;   // one CTA per SM (the lane-replicated codebook takes 128 KiB)
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvDepth = QP_GEMV_DEPTH;       // super-tiles prefetched ahead per warp (register staged)

// even split of T work items over the grid's warps, computed on the host: warp w owns
// [w*base + min(w, rem), ... + base + (w < rem))
struct RunSplit {
    unsigned base, rem;
};
inline RunSplit make_split(long T, int nwarps) { return RunSplit{(unsigned)(T / nwarps), (unsigned)(T % nwarps)}; }
__device__ __forceinline__ void split_range(const RunSplit s, int w, unsigned &lo, unsigned &hi) {
    const unsigned uw = (unsigned)w;
    lo = uw * s.base + (uw < s.rem ? uw : s.rem);
    hi = lo + s.base + (uw < s.rem ? 1u : 0u);
}

// lane-replicated table fill: `rows` slots of 128 bytes, slot r = 32 copies of value(r).  8 lanes cover a slot with one
// 16-byte store each, so a warp store instruction writes 4 consecutive slots (512 contiguous bytes, conflict-free).
template <class F>
__device__ __forceinline__ void fill_replicated_128(uint32_t *tab, int rows, F value) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
    for (int r0 = warp * 4; r0 < rows; r0 += nw * 4) {
        const int r = r0 + (lane >> 3);
        if (r < rows) {
            const uint32_t v = value(r);
            t4[r * 8 + (lane & 7)] = make_uint4(v, v, v, v);
        }
    }
}

// predicated streaming load of one payload: registers keep their old value when !pred (no select, no wait)
template <int E>
__device__ __forceinline__ void pack_load_raw_pred(uint32_t (&raw)[TcqGeom<E>::kRawWords], const uint32_t *p, bool pred) {
    constexpr int NW = TcqGeom<E>::kRawWords;
    constexpr int LB = TcqGeom<E>::kLaneBytes;
    const int ip = pred ? 1 : 0;
    if constexpr (LB % 16 == 0) {
#pragma unroll
        for (int i = 0; i < NW / 4; ++i)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                "@p ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];\n\t}"
                : "+r"(raw[4 * i]), "+r"(raw[4 * i + 1]), "+r"(raw[4 * i + 2]), "+r"(raw[4 * i + 3])
                : "l"(p + 4 * i), "r"(ip));
    } else if constexpr (LB % 8 == 0) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t"
                "@p ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];\n\t}"
                : "+r"(raw[2 * i]), "+r"(raw[2 * i + 1])
                : "l"(p + 2 * i), "r"(ip));
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                "@p ld.global.nc.L1::evict_first.b32 %0, [%1];\n\t}"
                : "+r"(raw[i])
                : "l"(p + i), "r"(ip));
    }
}

// Work assignment: the part's super-tiles [0, T) are split into one contiguous range per CTA; inside the CTA the warps
// interleave (warp w takes range_lo + w, + kGemvWarps, ...).  The CTA therefore streams a single moving window of
// kGemvWarps * 64*E contiguous bytes (DRAM-page friendly, like a copy kernel) while a warp's consecutive super-tiles still
// belong to the same 32-row strip most of the time, so accumulators are flushed only when the strip changes.
struct WarpRun {
    int n;      // super-tiles this warp processes
    int mh, kh; // strip / column of the first one
    size_t first;  // its index
};
__device__ __forceinline__ WarpRun warp_run(const PackSegment seg, unsigned clo, unsigned chi, int warp) {
    WarpRun r;
    const int span = (int)(chi - clo);
    r.n = span > warp ? (span - warp + kGemvWarps - 1) / kGemvWarps : 0;
    r.first = (size_t)clo + warp;
    const unsigned it0 = clo + (unsigned)warp;
    r.mh = (int)(it0 / (unsigned)seg.ksuper);
    r.kh = (int)(it0 - (unsigned)r.mh * (unsigned)seg.ksuper);
    return r;
}

// issue the first kGemvDepth payload loads of this warp's run
template <int E>
__device__ __forceinline__ void gemv_prefetch(const PackSegment seg, const WarpRun run,
                                              uint32_t (&raw)[kGemvDepth][TcqGeom<E>::kRawWords]) {
    using G = TcqGeom<E>;
    constexpr size_t kStride = (size_t)kGemvWarps * (G::kSuperBytes / 4);
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const uint32_t *p = seg.codes + word0 + run.first * (G::kSuperBytes / 4);
#pragma unroll
    for (int d = 0; d < kGemvDepth; ++d) {
#pragma unroll
        for (int i = 0; i < G::kRawWords; ++i) raw[d][i] = 0u;
        pack_load_raw_pred<E>(raw[d], p + d * kStride, d < run.n);
    }
}

// one super-tile: consume slot `raw` (decode -> 4 mma), refill it with the super-tile kGemvDepth steps ahead
template <class Dec, bool kRefillAlways>
__device__ __forceinline__ void gemv_step(uint32_t (&raw)[TcqGeom<Dec::kE>::kRawWords], const uint32_t *pnext,
                                          bool refill, int bitoff, int lane, uint32_t tab_lane,
                                          const uint8_t *xs_lane, bool xvalid, float (&acc)[2][4]) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    uint32_t P[G::kWords];
    tcq_align<E>(raw, bitoff, P);  // the slot's registers are dead after this: the refill below can land in them
    pack_load_raw_pred<E>(raw, pnext, kRefillAlways ? true : refill);
    uint4 xb = make_uint4(0u, 0u, 0u, 0u);
    if (xvalid) xb = *reinterpret_cast<const uint4 *>(xs_lane);
    uint32_t frag[4][4];  // [tile = kl*2+ml][register]
    Dec::decode(P, lane, tab_lane, frag);
    mma_16816(acc[0], frag[0][0], frag[0][1], frag[0][2], frag[0][3], xb.x, xb.y);
    mma_16816(acc[1], frag[1][0], frag[1][1], frag[1][2], frag[1][3], xb.x, xb.y);
    mma_16816(acc[0], frag[2][0], frag[2][1], frag[2][2], frag[2][3], xb.z, xb.w);
    mma_16816(acc[1], frag[3][0], frag[3][1], frag[3][2], frag[3][3], xb.z, xb.w);
}

// stream this warp's run of one part: decode (Dec) -> A fragments -> mma with x (B fragments from shared memory) ->
// fp32 atomics per finished 32-row strip.
//   Dec::kE                              bits per weight pair (payload geometry TcqGeom<kE>)
//   Dec::decode(P, lane, tab_lane, frag) 16 half2 registers of the (lane, super-tile) from its aligned payload words
// `between` runs once after the steady state and before the drain (used to issue the next part's first loads so its DRAM
// latency hides behind this part's tail).
template <class Dec, class Between>
__device__ __forceinline__ void gemv_run_segment(const PackSegment seg, float *__restrict__ out, int M, int bs,
                                                 const uint8_t *xs, uint32_t tab_lane, const WarpRun run,
                                                 uint32_t (&raw)[kGemvDepth][TcqGeom<Dec::kE>::kRawWords],
                                                 Between between) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    constexpr size_t kStride = (size_t)kGemvWarps * (G::kSuperBytes / 4);  // words between a warp's consecutive super-tiles
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const int nq = lane >> 2, q = lane & 3;
    const bool xvalid = nq < bs;

    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int n = run.n;
    int mh = run.mh, kh = run.kh;
    const uint32_t *p = seg.codes + word0 + run.first * (G::kSuperBytes / 4) + kGemvDepth * kStride;
    const uint8_t *xs_q = xs + ((size_t)seg.ksuper0 * bs + nq) * 64 + q * 16;
    const int xstep = bs * 64;

    auto advance = [&]() {
        kh += kGemvWarps;
        if (kh >= seg.ksuper) {  // this warp's next super-tile is in a later strip: flush (rarely more than one wrap)
            gemv_flush(out, M, bs, seg.row0 + mh * 32, lane, acc);
            do {
                kh -= seg.ksuper;
                ++mh;
            } while (kh >= seg.ksuper);
        }
    };

    // steady state: every refill is in range
    while (n >= 2 * kGemvDepth) {
#pragma unroll
        for (int d = 0; d < kGemvDepth; ++d) {
            gemv_step<Dec, true>(raw[d], p + d * kStride, true, bitoff, lane, tab_lane, xs_q + kh * xstep, xvalid, acc);
            advance();
        }
        p += kGemvDepth * kStride;
        n -= kGemvDepth;
    }
    between();
    // drain: fewer than 2*kGemvDepth left; refills are predicated
    while (n > 0) {
#pragma unroll
        for (int d = 0; d < kGemvDepth; ++d) {
            if (d < n) {
                gemv_step<Dec, false>(raw[d], p + d * kStride, d + kGemvDepth < n, bitoff, lane, tab_lane,
                                      xs_q + kh * xstep, xvalid, acc);
                if (d + 1 < n || n > kGemvDepth) advance();
            }
        }
        p += kGemvDepth * kStride;
        n -= kGemvDepth;
    }
    if (run.n > 0) gemv_flush(out, M, bs, seg.row0 + mh * 32, lane, acc);
}

template <class Dec>
__device__ __forceinline__ void gemv_run_segment(const PackSegment seg, float *__restrict__ out, int M, int bs,
                                                 const uint8_t *xs, uint32_t tab_lane, const WarpRun run,
                                                 uint32_t (&raw)[kGemvDepth][TcqGeom<Dec::kE>::kRawWords]) {
    gemv_run_segment<Dec>(seg, out, M, bs, xs, tab_lane, run, raw, [] {});
}

// decode the warp's share of one part and write fp16 W (M, K) row-major
template <class Dec>
__device__ __forceinline__ void dequant_run_segment(const PackSegment seg, __half *__restrict__ W, int K,
                                                    uint32_t tab_lane, RunSplit split, int gwarp) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    const int lane = threadIdx.x & 31;
    int word0, bitoff;
    tcq_lane_addr<E>(lane, word0, bitoff);
    const uint32_t *lane_base = seg.codes + word0;
    unsigned lo, hi;
    split_range(split, gwarp, lo, hi);
    uint32_t *W32 = reinterpret_cast<uint32_t *>(W);
    const int kw = K / 2;
    for (unsigned it = lo; it < hi; ++it) {
        uint32_t raw[G::kRawWords];
        pack_load_raw<E>(raw, lane_base + (size_t)it * (G::kSuperBytes / 4));
        uint32_t P[G::kWords];
        tcq_align<E>(raw, bitoff, P);
        uint32_t frag[4][4];
        Dec::decode(P, lane, tab_lane, frag);
        const int mh = (int)(it / (unsigned)seg.ksuper), kh = (int)(it - (unsigned)mh * (unsigned)seg.ksuper);
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
