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
// kCoherent: the source is an activation written by another kernel of the same step -> ld.global.cg (L2), see stage_x
template <bool kCoherent = false>
__device__ __forceinline__ void coop_copy_words(uint32_t *dst, const uint32_t *__restrict__ src, int n) {
    constexpr int U = 8;
    for (int base = threadIdx.x; base < n; base += U * blockDim.x) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x;
            v[u] = (i < n) ? (kCoherent ? __ldcg(src + i) : __ldg(src + i)) : 0u;
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
// a single round for bs*K <= 24576), then scatters the words (4-way bank-conflicted 32-bit stores: measured cheaper than a
// conflict-free store fed by a 4-byte gather, whose 25 % sector efficiency costs more L1 request cycles than it saves).
// x is produced by the preceding kernel, which may still overlap this one's prologue (programmatic dependent launch): it is
// read at L2 (ld.global.cg), not through the non-coherent L1 path.
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
            v[u] = (i < total) ? __ldcg(x4 + i) : make_uint4(0u, 0u, 0u, 0u);
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

// warp index / grid-wide warp id as provably warp-uniform values (lets ptxas keep run bounds in uniform registers and
// drop the convergence barriers around the shuffles)
__device__ __forceinline__ int warp_in_cta() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

struct PackSegment {
    const uint32_t *codes;  // packed words of this part
    int strips;             // rows / 32
    int ksuper;             // cols / 32
    int row0;               // first output row
    int ksuper0;            // first super-tile column of x
    unsigned kmagic = 0;    // floor(2^32 / ksuper), set by the GEMV launchers (seg_with_magic): strip / column of a flat index
                            // without the ~30-instruction division sequence on the path between "x staged" and the loop
};
inline PackSegment seg_with_magic(PackSegment s) {
    s.kmagic = s.ksuper <= 1 ? 0xffffffffu : (unsigned)((1ull << 32) / (unsigned)s.ksuper);
    return s;
}

#ifndef QP_GEMV_THREADS
#define QP_GEMV_THREADS 768
#endif
#ifndef QP_GEMV_CTAS
#define QP_GEMV_CTAS 1
#endif
constexpr int kGemvCtasPerSM = QP_GEMV_CTAS;    // experiments: 2 CTAs of 384 threads per SM (needs a <= 64 KiB codebook)
constexpr int kGemvThreads = QP_GEMV_THREADS;   // one CTA per SM (the lane-replicated codebook takes 128 KiB)
constexpr int kGemvWarps = kGemvThreads / 32;

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

// =====================================================================================================================
// Streaming skeleton (round 2).  Per-lane streaming loads into a register ring, decode,
// mma.m16n8k16 against x fragments in shared memory, fp32 atomics per finished strip -- but organised for the fewest issued
// instructions per super-tile and a small instruction footprint (the round-1 loop spent ~35 of its 155 warp-instructions
// per super-tile on strip / x-address bookkeeping and was 100 KB of SASS per instantiation):
//   * every WARP owns one contiguous run of super-tiles (a CTA's 24 runs are adjacent, so the CTA still streams one
//     contiguous byte range): the column index advances by one per step, the x-fragment address by a constant, and the
//     strip changes at most once every `ksuper` steps (counted down in a uniform register);
//   * one main loop whose D unrolled steps are all valid, one predicated tail of < D steps; the flush is out of line.
// =====================================================================================================================
#ifndef QP_GEMV2_DEPTH
#define QP_GEMV2_DEPTH 3
#endif
#ifndef QP_REFILL_LATE
#define QP_REFILL_LATE 1
#endif
#ifndef QP_TRIP_CHECK
#define QP_TRIP_CHECK 0  // measured: 18.84 -> 18.56 us at 28672x4096 but 6.15 -> 6.26 at 6144x4096, decode step 498.3 vs 503.5 tok/s (profiles/r02_loop_variants.log)
#endif
constexpr int kGemv2Depth = QP_GEMV2_DEPTH;

struct WarpRun2 {
    unsigned lo;  // first super-tile (flat [strip][column] index inside the part)
    int n;        // super-tiles of this warp
    int mh, kh;   // strip / column of the first one
};
__device__ __forceinline__ WarpRun2 warp_run2(const PackSegment seg, const RunSplit split, int gwarp) {
    unsigned lo, hi;
    split_range(split, gwarp, lo, hi);
    WarpRun2 r;
    r.lo = lo;
    r.n = (int)(hi - lo);
    // lo / ksuper by multiplication: q = hi32(lo * floor(2^32 / ksuper)) is the quotient or one less (lo < 2^32)
    unsigned q = __umulhi(lo, seg.kmagic), rem = lo - q * (unsigned)seg.ksuper;
    if (rem >= (unsigned)seg.ksuper) {
        q += 1u;
        rem -= (unsigned)seg.ksuper;
    }
    r.mh = (int)q;
    r.kh = (int)rem;
    return r;
}

template <int E>
__device__ __forceinline__ const uint32_t *gemv2_lane_ptr(const PackSegment seg, const WarpRun2 run) {
    int word0, bitoff;
    tcq_lane_addr<E>(threadIdx.x & 31, word0, bitoff);
    return seg.codes + word0 + (size_t)run.lo * (TcqGeom<E>::kSuperBytes / 4);
}

// issue the first D payload loads of this warp's run
template <int E>
__device__ __forceinline__ void gemv2_prefetch(const PackSegment seg, const WarpRun2 run,
                                               uint32_t (&raw)[kGemv2Depth][TcqGeom<E>::kRawWords]) {
    using G = TcqGeom<E>;
    const uint32_t *p = gemv2_lane_ptr<E>(seg, run);
#pragma unroll
    for (int d = 0; d < kGemv2Depth; ++d) {
#pragma unroll
        for (int i = 0; i < G::kRawWords; ++i) raw[d][i] = 0u;
        pack_load_raw_pred<E>(raw[d], p + d * (G::kSuperBytes / 4), d < run.n);
    }
}

// add the C fragments of a finished 32-row strip to out (bs, M) and clear them: lane l holds rows row + l/4 + {0, 8, 16, 24}
// of batch columns 2(l%4) and 2(l%4) + 1.  Everything it needs is recomputed here from kernel parameters: the strip changes
// once every `ksuper` steps, so nothing of this may stay live in (or be rematerialised into) the streaming loop.
__device__ __forceinline__ void gemv2_flush(float *__restrict__ out, int M, int bs, int row, float (&acc)[2][4]) {
    const int lane = threadIdx.x & 31;
    const int c0 = 2 * (lane & 3);
    float *a = out + (size_t)c0 * M + row + (lane >> 2);
    if (c0 < bs) {
        atomicAdd(a, acc[0][0]);
        atomicAdd(a + 8, acc[0][2]);
        atomicAdd(a + 16, acc[1][0]);
        atomicAdd(a + 24, acc[1][2]);
    }
    if (c0 + 1 < bs) {
        a += M;
        atomicAdd(a, acc[0][1]);
        atomicAdd(a + 8, acc[0][3]);
        atomicAdd(a + 16, acc[1][1]);
        atomicAdd(a + 24, acc[1][3]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// shared-memory address of a lane's x fragment for column 0 of a part (see stage_x): a lane past the batch reads row 0's
// fragment -- its mma columns are never stored, so no zero-fill and no predicate is needed
__device__ __forceinline__ uint32_t gemv2_xbase(uint32_t xs_addr, int ksuper0, int bs) {
    const int lane = threadIdx.x & 31, nq = lane >> 2, q = lane & 3;
    return xs_addr + ((uint32_t)ksuper0 * (uint32_t)bs + (uint32_t)(nq < bs ? nq : 0)) * 64u + (uint32_t)q * 16u;
}

// one super-tile: payload registers -> decode -> 4 mma
template <class Dec>
__device__ __forceinline__ void gemv2_consume(const uint32_t (&P)[TcqGeom<Dec::kE>::kWords], uint32_t xaddr, int lane,
                                              uint32_t tab_lane, float (&acc)[2][4]) {
    const uint4 xb = lds_u128(xaddr);
    uint32_t frag[4][4];  // [tile = kl*2+ml][register]
    Dec::decode(P, lane, tab_lane, frag);
    mma_16816(acc[0], frag[0][0], frag[0][1], frag[0][2], frag[0][3], xb.x, xb.y);
    mma_16816(acc[1], frag[1][0], frag[1][1], frag[1][2], frag[1][3], xb.x, xb.y);
    mma_16816(acc[0], frag[2][0], frag[2][1], frag[2][2], frag[2][3], xb.z, xb.w);
    mma_16816(acc[1], frag[3][0], frag[3][1], frag[3][2], frag[3][3], xb.z, xb.w);
}

// stream this warp's run of one part.  xs_addr = shared-memory address of the staged x (fragment order, see stage_x);
// `between` runs once after the main loop, before the tail (used to issue the next part's first loads).
template <class Dec, class Between>
__device__ __forceinline__ void gemv2_run(const PackSegment seg, float *__restrict__ out, int M, int bs, uint32_t xs_addr,
                                          uint32_t tab_lane, const WarpRun2 run,
                                          uint32_t (&raw)[kGemv2Depth][TcqGeom<Dec::kE>::kRawWords], Between between) {
    constexpr int E = Dec::kE, D = kGemv2Depth;
    using G = TcqGeom<E>;
    constexpr int SBw = G::kSuperBytes / 4;  // words per super-tile
    const int lane = threadIdx.x & 31;
    const int bitoff = (lane * G::kLaneBytes & 3) * 8;
    const uint32_t xstep = (uint32_t)bs * 64u;
    uint32_t xa = gemv2_xbase(xs_addr, seg.ksuper0, bs) + (uint32_t)run.kh * xstep;
    int row = seg.row0 + run.mh * 32;
    int kleft = seg.ksuper - run.kh;  // steps until the strip ends (warp-uniform)

    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const uint32_t *p = gemv2_lane_ptr<E>(seg, run) + D * SBw;  // refill source of ring slot 0
    const int n = run.n;
    int i = 0;
    auto advance = [&]() {
        xa += xstep;
        if (--kleft == 0) {  // rare: once per strip
            gemv2_flush(out, M, bs, row, acc);
            row += 32;
            kleft = seg.ksuper;
            xa = gemv2_xbase(xs_addr, seg.ksuper0, bs);
        }
    };
    // main loop: D valid steps per trip; a refill is predicated on its super-tile being inside the run
    for (; i + D <= n; i += D) {
#if QP_TRIP_CHECK
        // the strip end is tested once per trip: a trip that stays inside the strip runs D steps as ONE basic block (no countdown,
        // compare and branch per step, and ptxas may overlap a step's lookups with the next step's extraction)
        if (kleft > D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                uint32_t P[G::kWords];
                tcq_align<E>(raw[d], bitoff, P);
                gemv2_consume<Dec>(P, xa, lane, tab_lane, acc);
                pack_load_raw_pred<E>(raw[d], p + d * SBw, i + d + D < n);
                xa += xstep;
            }
            kleft -= D;
            p += D * SBw;
            continue;
        }
#endif
#pragma unroll
        for (int d = 0; d < D; ++d) {
            uint32_t P[G::kWords];
            tcq_align<E>(raw[d], bitoff, P);
#if QP_REFILL_LATE
            // the refill is issued AFTER the decode has read the slot: issued before it (0), the in/out asm operands force a copy
            // of the payload words (5 IMAD.MOV per super-tile in the SASS of the KV = 6 loop; 11.58 -> 11.38 us at 4096x14336,
            // profiles/r02_loop_variants.log)
            gemv2_consume<Dec>(P, xa, lane, tab_lane, acc);
            pack_load_raw_pred<E>(raw[d], p + d * SBw, i + d + D < n);
#else
            pack_load_raw_pred<E>(raw[d], p + d * SBw, i + d + D < n);
            gemv2_consume<Dec>(P, xa, lane, tab_lane, acc);
#endif
            advance();
        }
        p += D * SBw;
    }
    between();
    // tail: n - i < D steps left, already in the ring
#pragma unroll
    for (int d = 0; d < D - 1; ++d) {
        if (i + d < n) {
            uint32_t P[G::kWords];
            tcq_align<E>(raw[d], bitoff, P);
            gemv2_consume<Dec>(P, xa, lane, tab_lane, acc);
            advance();
        }
    }
    if (n > 0 && kleft != seg.ksuper) gemv2_flush(out, M, bs, row, acc);
}

// =====================================================================================================================
// Batched form of the same loop for 9 <= bs <= 32 (NB = 2 or 4 blocks of 8 batch rows): a super-tile is decoded ONCE and fed to
// 4 * NB mma.  The reference dequantises the whole matrix to HBM and calls cuBLAS there (lib/linear/tcq_linear.py:75-84); the
// tcgen05 kernel of gemm_tc_kernels.cu is decode-latency bound (18 warps, 30 us at 14336x4096) and stays for bs > 32.
// x (bs, K) does not fit next to the 128 KiB codebook (and re-reading its fragments from L2 per super-tile is 58-117 MB of L2
// traffic at 14336x4096, which is what bounded the first version), so the work is ordered K-SLAB major: a slab is
// kMmaSlabBytes of x in fragment order = W = 128 / NB super-tile columns; the flat work index runs over
// [part][slab][strip][column in slab], every CTA takes an equal contiguous range of it (1-2 slabs), stages the slab's
// fragments in shared memory once and its 24 warps split the range evenly.  Inside a slab a strip's W super-tiles are
// contiguous in memory and the next strip is a constant jump away; the C fragments of a (strip, slab) go to `out` with fp32
// atomics (32 * bs floats per W super-tiles: < 0.3 M sector atomics per launch).
// =====================================================================================================================
constexpr int kMmaSlabBytes = 64 * 1024;
// payload ring depth of the batched loop
#ifndef QP_MMA_DEPTH4
#define QP_MMA_DEPTH4 3
#endif
template <int NB>
constexpr int kMmaDepth = NB >= 4 ? QP_MMA_DEPTH4 : kGemv2Depth;

// x (bs, K) fp16 -> fragment order xfrag[kh][j][lane] (uint4; j = batch block, lane = (n % 8) * 4 + q); rows past bs are zero
static __global__ void x_to_frag_kernel(uint4 *__restrict__ xfrag, const uint32_t *x32, int K, int bs, int NB) {
    pdl_wait();
    pdl_launch_dependents();
    const int total = (K / 32) * NB * 32;
    const int kw = K / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int lane = i & 31, j = (i >> 5) % NB, kh = (i >> 5) / NB;
        const int n = j * 8 + (lane >> 2), q = lane & 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < bs) {
            const uint32_t *src = x32 + (size_t)n * kw + kh * 16 + q;
            v = make_uint4(__ldcg(src), __ldcg(src + 4), __ldcg(src + 8), __ldcg(src + 12));
        }
        xfrag[i] = v;
    }
}

// one CTA-uniform piece of work: the part of the CTA's flat range that lies in one slab
struct MmaPiece {
    int part;        // 0 / 1
    int col0, w;     // first super-tile column of the slab inside the part, slab width
    unsigned lo, n;  // first index inside the slab ([strip][column in slab]) and count
    unsigned next;   // flat index after this piece
};
// flat order: part A slabs, then part B slabs; all slabs of a part are W wide except its last
__device__ __forceinline__ MmaPiece mma_piece(const PackSegment a, const PackSegment b, int W, unsigned lo, unsigned hi) {
    const unsigned TA = (unsigned)a.strips * (unsigned)a.ksuper;
    MmaPiece pc;
    pc.part = lo >= TA;
    const PackSegment s = pc.part ? b : a;
    const unsigned rel = lo - (pc.part ? TA : 0u), full = (unsigned)s.strips * (unsigned)W;
    const int ns = (s.ksuper + W - 1) / W;
    int j = (int)(rel / full);
    if (j > ns - 1) j = ns - 1;
    pc.col0 = j * W;
    pc.w = min(W, s.ksuper - pc.col0);
    pc.lo = rel - (unsigned)j * full;
    const unsigned slab_n = (unsigned)s.strips * (unsigned)pc.w;
    pc.n = min(hi - lo, slab_n - pc.lo);
    pc.next = lo + pc.n;
    return pc;
}

// a warp's share of a piece: payload pointer of the prefetch stream, countdowns of both streams
struct MmaRun {
    const uint32_t *p;  // next payload to fetch (lane pointer)
    int pleft;          // fetches until the prefetch stream leaves its strip
    int kleft;          // steps until the consume stream leaves its strip
    int row;            // output row of the consume stream's strip
    int n;              // super-tiles of this warp
    uint32_t xoff;      // byte offset of the consume stream's x fragments inside the slab
};

template <int E>
__device__ __forceinline__ void mma_fetch(MmaRun &r, uint32_t (&raw)[TcqGeom<E>::kRawWords], bool pred, int w, int jump_words) {
    pack_load_raw_pred<E>(raw, r.p, pred);
    r.p += TcqGeom<E>::kSuperBytes / 4;
    if (--r.pleft == 0) {
        r.p += jump_words;
        r.pleft = w;
    }
}

template <int E, int NB>
__device__ __forceinline__ MmaRun mma_begin(const PackSegment seg, const MmaPiece pc, int warps,
                                            uint32_t (&raw)[kMmaDepth<NB>][TcqGeom<E>::kRawWords]) {
    using G = TcqGeom<E>;
    const unsigned wi = (unsigned)warp_in_cta();
    const unsigned lo = pc.lo + (unsigned)(((unsigned long long)pc.n * wi) / (unsigned)warps);
    const unsigned hi = pc.lo + (unsigned)(((unsigned long long)pc.n * (wi + 1)) / (unsigned)warps);
    const int mh = (int)(lo / (unsigned)pc.w), kw = (int)(lo - (unsigned)mh * (unsigned)pc.w);
    int word0, bitoff;
    tcq_lane_addr<E>(threadIdx.x & 31, word0, bitoff);
    MmaRun r;
    r.p = seg.codes + word0 + ((size_t)mh * seg.ksuper + pc.col0 + kw) * (G::kSuperBytes / 4);
    r.pleft = r.kleft = pc.w - kw;
    r.row = seg.row0 + mh * 32;
    r.n = (int)(hi - lo);
    r.xoff = ((uint32_t)kw * (NB * 32) + (threadIdx.x & 31)) * 16u;
    const int jump = (seg.ksuper - pc.w) * (G::kSuperBytes / 4);
#pragma unroll
    for (int d = 0; d < kMmaDepth<NB>; ++d) {
#pragma unroll
        for (int i = 0; i < G::kRawWords; ++i) raw[d][i] = 0u;
        mma_fetch<E>(r, raw[d], d < r.n, pc.w, jump);
    }
    return r;
}

template <int NB>
__device__ __forceinline__ void mma_flush(float *__restrict__ out, int M, int bs, int row, float (&acc)[NB][2][4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int c0 = j * 8 + 2 * (lane & 3);
        float *a = out + (size_t)c0 * M + row + (lane >> 2);
        if (c0 < bs) {
            atomicAdd(a, acc[j][0][0]);
            atomicAdd(a + 8, acc[j][0][2]);
            atomicAdd(a + 16, acc[j][1][0]);
            atomicAdd(a + 24, acc[j][1][2]);
        }
        if (c0 + 1 < bs) {
            a += M;
            atomicAdd(a, acc[j][0][1]);
            atomicAdd(a + 8, acc[j][0][3]);
            atomicAdd(a + 16, acc[j][1][1]);
            atomicAdd(a + 24, acc[j][1][3]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j][i][k] = 0.f;
    }
}

template <class Dec, int NB>
__device__ __forceinline__ void mma_consume(const uint32_t (&P)[TcqGeom<Dec::kE>::kWords], uint32_t xaddr, int lane,
                                            uint32_t tab_lane, float (&acc)[NB][2][4]) {
    uint32_t frag[4][4];  // [tile = kl*2+ml][register]
    Dec::decode(P, lane, tab_lane, frag);
    if constexpr (NB <= 2) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const uint4 xb = lds_u128(xaddr + j * 512);
            mma_16816(acc[j][0], frag[0][0], frag[0][1], frag[0][2], frag[0][3], xb.x, xb.y);
            mma_16816(acc[j][1], frag[1][0], frag[1][1], frag[1][2], frag[1][3], xb.x, xb.y);
            mma_16816(acc[j][0], frag[2][0], frag[2][1], frag[2][2], frag[2][3], xb.z, xb.w);
            mma_16816(acc[j][1], frag[3][0], frag[3][1], frag[3][2], frag[3][3], xb.z, xb.w);
        }
    } else {  // registers to spare (512 threads): all x fragments first, the two mma of an accumulator 2 * NB instructions apart
        uint4 xb[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) xb[j] = lds_u128(xaddr + j * 512);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            mma_16816(acc[j][0], frag[0][0], frag[0][1], frag[0][2], frag[0][3], xb[j].x, xb[j].y);
            mma_16816(acc[j][1], frag[1][0], frag[1][1], frag[1][2], frag[1][3], xb[j].x, xb[j].y);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            mma_16816(acc[j][0], frag[2][0], frag[2][1], frag[2][2], frag[2][3], xb[j].z, xb[j].w);
            mma_16816(acc[j][1], frag[3][0], frag[3][1], frag[3][2], frag[3][3], xb[j].z, xb[j].w);
        }
    }
}

// stream a warp's share of one piece; xs_addr = shared-memory address of the staged slab
template <class Dec, int NB>
__device__ __forceinline__ void mma_stream(const PackSegment seg, const MmaPiece pc, MmaRun r, float *__restrict__ out, int M,
                                           int bs, uint32_t xs_addr, uint32_t tab_lane,
                                           uint32_t (&raw)[kMmaDepth<NB>][TcqGeom<Dec::kE>::kRawWords]) {
    constexpr int E = Dec::kE, D = kMmaDepth<NB>;
    using G = TcqGeom<E>;
    const int lane = threadIdx.x & 31;
    const int bitoff = (lane * G::kLaneBytes & 3) * 8;
    const int jump = (seg.ksuper - pc.w) * (G::kSuperBytes / 4);
    const uint32_t xlane = xs_addr + (uint32_t)lane * 16u;
    uint32_t xa = xs_addr + r.xoff;
    float acc[NB][2][4];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j][i][k] = 0.f;
    auto advance = [&]() {
        xa += NB * 512;
        if (--r.kleft == 0) {
            mma_flush<NB>(out, M, bs, r.row, acc);
            r.row += 32;
            r.kleft = pc.w;
            xa = xlane;
        }
    };
    const int n = r.n;
    int i = 0;
    for (; i + D <= n; i += D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            uint32_t P[G::kWords];
            tcq_align<E>(raw[d], bitoff, P);
#if QP_REFILL_LATE
            mma_consume<Dec, NB>(P, xa, lane, tab_lane, acc);
            mma_fetch<E>(r, raw[d], i + d + D < n, pc.w, jump);
#else
            mma_fetch<E>(r, raw[d], i + d + D < n, pc.w, jump);
            mma_consume<Dec, NB>(P, xa, lane, tab_lane, acc);
#endif
            advance();
        }
    }
#pragma unroll
    for (int d = 0; d < D - 1; ++d) {
        if (i + d < n) {
            uint32_t P[G::kWords];
            tcq_align<E>(raw[d], bitoff, P);
            mma_consume<Dec, NB>(P, xa, lane, tab_lane, acc);
            advance();
        }
    }
    if (n > 0 && r.kleft != pc.w) mma_flush<NB>(out, M, bs, r.row, acc);
}

inline int check_align(const void *p, size_t a, const char *name) {
    if (((uintptr_t)p) % a != 0) return fail(QP_ERR_ALIGN, "%s must be %zu-byte aligned", name, a);
    return QP_OK;
}

// ---- host side of the batched mma form: <= kMmaMaxBatch batch rows per launch, fragment-ordered x in `scratch` ----
#ifndef QP_MMA_THREADS4
#define QP_MMA_THREADS4 512
#endif
// NB = 4 holds 32 accumulators: 16 warps of <= 128 registers instead of 24 of 80, which spilled inside the loop (measured
// 14336x4096 bs = 32: 24.0 us at 512 threads, 25.0 at 640, 26.7 at 768 with a 2-deep ring).  NB = 8 (12 warps of 168
// registers) measured 33-35 us for bs = 48 / 64 against 31-32 us of the tcgen05 kernel, so the mma form stops at bs = 32.
template <int NB>
constexpr int kMmaThreads = NB >= 4 ? QP_MMA_THREADS4 : kGemvThreads;
// batch rows per launch -> blocks of 8 batch rows the kernel carries
static inline int mma_batch_blocks(int bs) { return bs <= 16 ? 2 : 4; }
constexpr int kMmaMaxBatch = 32;

static inline size_t gemm_mma_scratch_bytes(int K, int bs) { return (size_t)(K / 32) * mma_batch_blocks(bs) * 32 * 16; }
static inline size_t gemm_mma_scratch_total(int K, int bs) {
    size_t total = 0;
    for (int b0 = 0; b0 < bs; b0 += kMmaMaxBatch) total += gemm_mma_scratch_bytes(K, bs - b0 < kMmaMaxBatch ? bs - b0 : kMmaMaxBatch);
    return total;
}
// launch_one(out rows b0.., xfrag, nb) launches the GEMM kernel for one chunk of the batch
template <class F>
static int mma_gemm_batches(float *out, const void *x_f16, void *scratch, int M, int K, int bs, cudaStream_t st, F launch_one) {
    uint8_t *sc = (uint8_t *)scratch;
    for (int b0 = 0; b0 < bs; b0 += kMmaMaxBatch) {
        const int nb = bs - b0 < kMmaMaxBatch ? bs - b0 : kMmaMaxBatch, NB = mma_batch_blocks(nb);
        const int total = (K / 32) * NB * 32;
        QP_CUDA(launch_pdl(x_to_frag_kernel, dim3((total + 255) / 256), dim3(256), 0, st, (uint4 *)sc,
                           (const uint32_t *)((const __half *)x_f16 + (size_t)b0 * K), K, nb, NB));
        int rc = check_launch("x_to_frag");
        if (rc != QP_OK) return rc;
        if ((rc = launch_one(out + (size_t)b0 * M, (const uint4 *)sc, nb)) != QP_OK) return rc;
        sc += gemm_mma_scratch_bytes(K, nb);
    }
    return QP_OK;
}

}  // namespace qp
