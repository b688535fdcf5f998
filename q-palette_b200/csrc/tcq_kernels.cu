// tcq_kernels.cu -- fused trellis-decode + GEMV (bs <= 8) and dequantise kernels for Q-Palette's TCQ quantizers.
//
// What is computed (reference: kernels/tcq-kernels/src/inference.cu:408-1819, format lib/quantizer/tcq_quant.py:47-60):
//   out[n][m] (+)= sum_k decode(W)[m][k] * x[n][k]
// Design (B200): one persistent CTA per SM, 16 warps.  The packed matrix is a flat array of 32x32 super-tiles
// (64*KV contiguous bytes each, [M/32][K/32] order); every warp owns one contiguous run of super-tiles, so the whole
// grid streams the buffer exactly once with contiguous per-warp reads and an (almost) perfectly even split whatever
// M and K are.  Per super-tile a lane: loads its 16*KV payload bits straight into registers (prefetched kDepth
// super-tiles ahead), exchanges chunk tops with lane+1 by shuffle (tail-biting trellis), forms 16 states, hashes
// them (s*(s+1)), looks the fp16 pairs up in a lane-replicated shared-memory codebook (bank = lane -> conflict-free)
// and feeds them as the A fragment of mma.m16n8k16 with x as the B fragment (fp32 accumulate).  Partial sums of a
// 32-row strip are added to `out` with fp32 atomics when the run leaves the strip.
// The decode is ALU/issue-bound on B200 (HBM bytes per SM-clock are 7x an RTX 4090's), so the inner loop is written
// to minimise issued instructions per weight pair: see DESIGN.md "TCQ GEMV instruction budget".
#include <type_traits>
#include "qp_common.cuh"
#include "tcq_bits.cuh"
#include "gemv_common.cuh"
#include "xprod.cuh"

namespace qp {

#ifdef QP_PROFILE_PHASES
// debug build only: per-CTA %globaltimer stamps of the GEMV phases.  g_phase holds the last launch (qp_debug_phases); g_plog
// is an append-only log of every CTA of every launch {M<<32|K, cta | xmode<<16, 7 stamps} (qp_debug_plog), so that the launches
// of a whole decode step can be laid out on one timeline (tools/phase_profile_step.py)
constexpr unsigned kPlogCap = 40000;
__device__ unsigned long long g_phase[256][8];
__device__ unsigned long long g_plog[kPlogCap][9];
__device__ unsigned g_plog_n;
#define QP_PHASE_DECL unsigned long long ph_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define QP_PHASE(i)                                                             \
    do {                                                                        \
        if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ph_[i])); \
    } while (0)
#else
#define QP_PHASE_DECL
#define QP_PHASE(i)
#endif

constexpr int kTcqThreads = kGemvThreads;
constexpr int kTcqWarps = kGemvWarps;

// ---- lane-replicated codebook in shared memory ---------------------------------------------------------------------
// entry slot = 2^kStrideLog2 bytes = one 4-byte copy per lane (or per lane pair for S = 11), so a warp-wide gather never
// has two lanes in the same bank with different addresses.  For S = 9 the sign flip (bit 15 of the hash) is folded into
// the table (2^10 entries) which saves the XOR in the inner loop; S = 10/11 keep the XOR (table would not fit).
template <int S>
struct TcqTable {
    static_assert(S >= 9 && S <= 11, "tlut_bits must be 9, 10 or 11");
#ifndef QP_TCQ_FOLD
#define QP_TCQ_FOLD 1
#endif
    static constexpr bool kFold = (S == 9) && QP_TCQ_FOLD;
#ifndef QP_TCQ_STRIDE9
#define QP_TCQ_STRIDE9 7  // experiments: 6 = 64-byte slots for S = 9 (16 copies, 2-way conflicts, no hash pre-shift, 64 KiB)
#endif
    static constexpr int kStrideLog2 = (S == 11) ? 6 : (S == 9 ? QP_TCQ_STRIDE9 : 7);
    static constexpr int kEntryBits = S + (kFold ? 1 : 0);
    static constexpr int kEntries = 1 << kEntryBits;
    static constexpr int kBytes = kEntries << kStrideLog2;            // 128 KiB for all three
    static constexpr int kShift = kStrideLog2 - (15 - S);             // hash bit (15-S) -> address bit kStrideLog2
    static constexpr uint32_t kMask = (uint32_t)(kEntries - 1) << kStrideLog2;
    static constexpr uint32_t kLaneMask = (1u << (kStrideLog2 - 2)) - 1u;
};

// lane-replicated codebook straight from global memory, in two halves so that the kernels can put other work between
// them: tcq_table_load issues all of a thread's tlut loads (one L2 round trip), tcq_table_store writes the copies.
// A warp store covers 4 consecutive 128-byte slots (8 lanes x 16 bytes each).
template <int S, int WARPS = kGemvWarps>
struct TcqTableRegs {
    static constexpr int kRows = TcqTable<S>::kBytes / 128;  // 128-byte rows of the table
    static constexpr int kIter = (kRows + WARPS * 4 - 1) / (WARPS * 4);
    uint32_t v[kIter];
};

template <int S, int WARPS = kGemvWarps>
__device__ __forceinline__ void tcq_table_load(TcqTableRegs<S, WARPS> &t, const uint32_t *__restrict__ tlut) {
    using T = TcqTable<S>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < TcqTableRegs<S, WARPS>::kIter; ++it) {
        const int r = (it * WARPS + warp) * 4 + (lane >> 3);
        // 128-byte row r holds 128 >> kStrideLog2 entries; this lane's 16-byte chunk belongs to entry ef (incl. the fold bit)
        const int ef = (r << (7 - T::kStrideLog2)) + ((lane & 7) >> (T::kStrideLog2 - 4));
        t.v[it] = (r < TcqTableRegs<S, WARPS>::kRows) ? __ldg(tlut + (ef & ((1 << S) - 1))) : 0u;
    }
}

// the sign fold (negate component 0 for the upper half of the folded table) is applied here, at store time: nothing between
// the load issue and this point may depend on the loaded values, so that their L2 round trip overlaps the rest of the prologue
template <int S, int WARPS = kGemvWarps>
__device__ __forceinline__ void tcq_table_store(uint32_t *tab, const TcqTableRegs<S, WARPS> &t) {
    using T = TcqTable<S>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
#pragma unroll
    for (int it = 0; it < TcqTableRegs<S, WARPS>::kIter; ++it) {
        const int r = (it * WARPS + warp) * 4 + (lane >> 3);
        const int ef = (r << (7 - T::kStrideLog2)) + ((lane & 7) >> (T::kStrideLog2 - 4));
        const uint32_t v = (T::kFold && (ef >> S)) ? (t.v[it] ^ 0x8000u) : t.v[it];
        if (r < TcqTableRegs<S, WARPS>::kRows) t4[r * 8 + (lane & 7)] = make_uint4(v, v, v, v);
    }
}

template <int S>
__device__ __forceinline__ void tcq_build_table(uint32_t *tab, const uint32_t *__restrict__ tlut) {
    TcqTableRegs<S> t;
    tcq_table_load<S>(t, tlut);
    tcq_table_store<S>(tab, t);
}

template <int S>
__device__ __forceinline__ uint32_t tcq_lookup(uint32_t tab_lane, uint32_t u) {
    using T = TcqTable<S>;
    // hash t = u*(u+1) pre-shifted by kShift with two multiply-adds (fma pipe) instead of multiply + shift/add (alu pipe,
    // which the extraction shifts and the mask already saturate): t << k = u * ((u << k) + (1 << k))
#ifdef QP_HASH_ADD
    const uint32_t t0 = tcq_hash(u);
    const uint32_t ts = (T::kShift == 1) ? (t0 + t0) : (t0 << T::kShift);
#else
    const uint32_t ts = u * (u * (1u << T::kShift) + (1u << T::kShift));
#endif
    // slot offset = hash bits [15-S, 15-S+kEntryBits) moved to [kStrideLog2, ...); tab_lane already carries the lane's
    // 4-byte column, so the address is base + offset with no further arithmetic
    uint32_t w = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + ((ts & T::kMask) | tab_lane));
    if (!T::kFold) w ^= ((ts >> T::kShift) & 0x8000u);
    return w;
}

// decode the 16 fp16 pairs of one (lane, super-tile): frag[t][j] = A-fragment register j of tile t = kl*2+ml
template <int KV, int S>
struct TcqDecoder {
    static constexpr int kE = KV;
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<KV>::kWords], int lane,
                                                  uint32_t tab_addr_lane, uint32_t (&frag)[4][4]) {
        using G = TcqGeom<KV>;
        uint32_t send[4] = {tcq_send<KV, 0>(P), tcq_send<KV, 1>(P), tcq_send<KV, 2>(P), tcq_send<KV, 3>(P)};
        uint32_t n1[4], n2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            n1[t] = __shfl_sync(0xffffffffu, send[t], (lane + 1) & 31);
            n2[t] = (G::kNeighbors == 2) ? __shfl_sync(0xffffffffu, send[t], (lane + 2) & 31) : 0u;
        }
        uint32_t u[4][4];
        tcq_states<KV, 0>(P, n1[0], n2[0], u[0]);
        tcq_states<KV, 1>(P, n1[1], n2[1], u[1]);
        tcq_states<KV, 2>(P, n1[2], n2[2], u[2]);
        tcq_states<KV, 3>(P, n1[3], n2[3], u[3]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) frag[t][j] = tcq_lookup<S>(tab_addr_lane, u[t][j]);
    }
};

using TcqSegment = PackSegment;

// ---- GEMV -----------------------------------------------------------------------------------------------------------
// XMODE = 0: x is given (plain staging).  1: the x-producer prologue of xprod.cuh (separate instantiation so that the plain
// kernel's instruction footprint stays small).  2: ... with one operand polled out of the row-sharded LL receive buffer.
template <int KVA, int KVB, int S, int XMODE>
__global__ void __launch_bounds__(kTcqThreads, kGemvCtasPerSM)
tcq_gemv_kernel(TcqSegment segA, TcqSegment segB, RunSplit splitA, RunSplit splitB, float *__restrict__ out,
                const uint32_t *__restrict__ x32, const uint32_t *__restrict__ tlut, int M, int K, int bs, XProd prod) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ float red[32];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    uint32_t *xs = reinterpret_cast<uint32_t *>(smem + TcqTable<S>::kBytes);

    const int lane = threadIdx.x & 31;
    const int warp = warp_in_cta();
    const int gwarp = blockIdx.x * kTcqWarps + warp;
    QP_PHASE_DECL;
    QP_PHASE(0);
    // this warp's contiguous run of each part (a CTA's runs are adjacent: the CTA streams one contiguous byte range)
    const WarpRun2 runA = warp_run2(segA, splitA, gwarp);
    // the HBM stream starts first (it has the longest latency and depends on nothing), then the codebook loads: both round
    // trips run under the rest of the prologue
    uint32_t rawA[kGemv2Depth][TcqGeom<KVA>::kRawWords];
    gemv2_prefetch<KVA>(segA, runA, rawA);
    TcqTableRegs<S> tregs;
    tcq_table_load<S>(tregs, tlut);
    QP_PHASE(1);
    if constexpr (XMODE != 0) {
        float *xscratch = reinterpret_cast<float *>(xs + (size_t)K * bs / 2);
        // the rest of the prologue, with the x-producer's constant inputs (scales, norm weight, signs) fetched before the wait
        auto finish_prologue = [&](auto ch_tag) {
            constexpr int CH = decltype(ch_tag)::value;
            XPre<CH> pre;
            xp_preload<CH>(pre, prod, K);
            xp_zero(prod);
            tcq_table_store<S>(tab, tregs);
            QP_PHASE(2);
            pdl_wait();  // x (and out) are produced by the preceding kernel
            QP_PHASE(3);
            produce_x<CH, XMODE == 2>(xs, xscratch, red, prod, K, pre);
        };
        if (((K >> 2) + kTcqThreads - 1) / kTcqThreads <= 2) finish_prologue(std::integral_constant<int, 2>{});
        else finish_prologue(std::integral_constant<int, 5>{});  // host guarantees K <= 5 * 4 * kTcqThreads
    } else {
        tcq_table_store<S>(tab, tregs);
        QP_PHASE(2);
        pdl_wait();  // x (and out) are produced by the preceding kernel
        QP_PHASE(3);
        stage_x(xs, x32, K, bs);
    }
    __syncthreads();
    QP_PHASE(4);
    pdl_launch_dependents();

    const uint32_t tab_addr_lane = (lane & TcqTable<S>::kLaneMask) << 2;  // the table starts the dynamic shared memory
    const uint32_t xs_addr = smem_u32(xs);
    if constexpr (KVB == 0) {
        gemv2_run<TcqDecoder<KVA, S>>(segA, out, M, bs, xs_addr, tab_addr_lane, runA, rawA, [] {});
    } else {
        const WarpRun2 runB = warp_run2(segB, splitB, gwarp);
        uint32_t rawB[kGemv2Depth][TcqGeom<KVB>::kRawWords];
        // the second part's first loads are issued while the first part's tail drains
        gemv2_run<TcqDecoder<KVA, S>>(segA, out, M, bs, xs_addr, tab_addr_lane, runA, rawA,
                                      [&] { gemv2_prefetch<KVB>(segB, runB, rawB); });
        gemv2_run<TcqDecoder<KVB, S>>(segB, out, M, bs, xs_addr, tab_addr_lane, runB, rawB, [] {});
    }
#ifdef QP_PROFILE_PHASES
    QP_PHASE(5);  // thread 0's warp done
    __syncthreads();
    QP_PHASE(6);  // whole CTA done
    if (threadIdx.x == 0) {
        for (int i = 0; i < 7; ++i) g_phase[blockIdx.x][i] = ph_[i];
        const unsigned k = atomicAdd(&g_plog_n, 1u);
        if (k < kPlogCap) {
            g_plog[k][0] = ((unsigned long long)M << 32) | (unsigned)K;
            g_plog[k][1] = blockIdx.x | (XMODE << 16);
            for (int i = 0; i < 7; ++i) g_plog[k][2 + i] = ph_[i];
        }
    }
#endif
}

// ---- batched GEMM on the GEMV loop (9 <= bs <= 32 per launch) ------------------------------------------------------------
template <int KVA, int KVB, int S, int NB>
__global__ void __launch_bounds__(kMmaThreads<NB>, 1)
tcq_gemm_mma_kernel(TcqSegment segA, TcqSegment segB, RunSplit split, float *__restrict__ out, const uint4 *xfrag,
                    const uint32_t *__restrict__ tlut, int M, int bs) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    uint4 *xs = reinterpret_cast<uint4 *>(smem + TcqTable<S>::kBytes);
    constexpr int W = kMmaSlabBytes / (NB * 512), kWarps = kMmaThreads<NB> / 32;
    const int lane = threadIdx.x & 31;
    unsigned lo, hi;
    split_range(split, (int)blockIdx.x, lo, hi);  // this CTA's range of the slab-major work order
    if (lo >= hi) return;
    uint32_t rawA[kMmaDepth<NB>][TcqGeom<KVA>::kRawWords];
    uint32_t rawB[kMmaDepth<NB>][TcqGeom<KVB == 0 ? KVA : KVB>::kRawWords];
    MmaPiece pc = mma_piece(segA, segB, W, lo, hi);
    MmaRun run;
    auto begin = [&]() {
        if (KVB == 0 || pc.part == 0) run = mma_begin<KVA, NB>(segA, pc, kWarps, rawA);
        else run = mma_begin<(KVB == 0 ? KVA : KVB), NB>(segB, pc, kWarps, rawB);
    };
    begin();  // the first payload loads are in flight while the codebook is built
    TcqTableRegs<S, kWarps> tregs;
    tcq_table_load<S, kWarps>(tregs, tlut);
    tcq_table_store<S, kWarps>(tab, tregs);
    pdl_wait();  // xfrag (and out) are produced by the preceding kernels
    const uint32_t tab_addr_lane = (lane & TcqTable<S>::kLaneMask) << 2;
    const uint32_t xs_addr = smem_u32(xs);
    while (true) {
        {   // stage the slab's x fragments
            const PackSegment sg = (KVB == 0 || pc.part == 0) ? segA : segB;
            const uint4 *src = xfrag + (size_t)(sg.ksuper0 + pc.col0) * (NB * 32);
            for (int i = threadIdx.x; i < pc.w * NB * 32; i += kMmaThreads<NB>) xs[i] = __ldcg(src + i);
        }
        __syncthreads();
        if (KVB == 0 || pc.part == 0)
            mma_stream<TcqDecoder<KVA, S>, NB>(segA, pc, run, out, M, bs, xs_addr, tab_addr_lane, rawA);
        else
            mma_stream<TcqDecoder<(KVB == 0 ? KVA : KVB), S>, NB>(segB, pc, run, out, M, bs, xs_addr, tab_addr_lane, rawB);
        if (pc.next >= hi) break;
        pc = mma_piece(segA, segB, W, pc.next, hi);
        begin();
        __syncthreads();  // every warp is done with the slab before it is overwritten
    }
    pdl_launch_dependents();
}

// ---- dequantise -----------------------------------------------------------------------------------------------------
template <int KVA, int KVB, int S>
__global__ void __launch_bounds__(kTcqThreads, 1)
tcq_dequant_kernel(TcqSegment segA, TcqSegment segB, RunSplit splitA, RunSplit splitB, __half *__restrict__ W,
                   const uint32_t *__restrict__ tlut, int K) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kTcqWarps + warp_in_cta();
    tcq_build_table<S>(tab, tlut);
    __syncthreads();
    const uint32_t tab_addr_lane = (lane & TcqTable<S>::kLaneMask) << 2;  // the table starts the dynamic shared memory
    dequant_run_segment<TcqDecoder<KVA, S>>(segA, W, K, tab_addr_lane, splitA, gwarp);
    if constexpr (KVB != 0) dequant_run_segment<TcqDecoder<KVB, S>>(segB, W, K, tab_addr_lane, splitB, gwarp);
}

// ---- host dispatch --------------------------------------------------------------------------------------------------
struct TcqLaunch {
    TcqSegment a, b;
    int kva, kvb;
};

static int make_segments(TcqLaunch &L, const void *codes1, const void *codes2, int M, int K, int KV1, int KV2,
                         int split_mode, int part1) {
    QP_CHECK_ARG(M > 0 && K > 0 && M % 32 == 0 && K % 32 == 0, "TCQ needs M %% 32 == 0 and K %% 32 == 0 (got %d x %d)", M, K);
    QP_CHECK_ARG(codes1 != nullptr, "codes1 is NULL");
    QP_CHECK_ARG(KV1 >= 2 && KV1 <= 10, "KV1 = %d out of range 2..10", KV1);
    if (split_mode == QP_SPLIT_NONE) {
        L.a = TcqSegment{(const uint32_t *)codes1, M / 32, K / 32, 0, 0};
        L.b = TcqSegment{nullptr, 0, 0, 0, 0};
        L.kva = KV1;
        L.kvb = 0;
        return QP_OK;
    }
    QP_CHECK_ARG(codes2 != nullptr, "codes2 is NULL for a two-rate layer");
    QP_CHECK_ARG(KV2 >= 2 && KV2 <= 10, "KV2 = %d out of range 2..10", KV2);
    if (split_mode == QP_SPLIT_IN) {
        QP_CHECK_ARG(part1 > 0 && part1 < K && part1 % 32 == 0, "in_part boundary %d must be a multiple of 32 inside (0,%d)", part1, K);
        L.a = TcqSegment{(const uint32_t *)codes1, M / 32, part1 / 32, 0, 0};
        L.b = TcqSegment{(const uint32_t *)codes2, M / 32, (K - part1) / 32, 0, part1 / 32};
    } else if (split_mode == QP_SPLIT_OUT) {
        QP_CHECK_ARG(part1 > 0 && part1 < M && part1 % 32 == 0, "out_part boundary %d must be a multiple of 32 inside (0,%d)", part1, M);
        L.a = TcqSegment{(const uint32_t *)codes1, part1 / 32, K / 32, 0, 0};
        L.b = TcqSegment{(const uint32_t *)codes2, (M - part1) / 32, K / 32, part1, 0};
    } else {
        return fail(QP_ERR_ARG, "unknown split_mode %d", split_mode);
    }
    L.kva = KV1;
    L.kvb = KV2;
    return QP_OK;
}

template <int KVA, int KVB, int S>
static int launch_gemv(const TcqLaunch &L, float *out, const void *x, const void *tlut, int M, int K, int bs,
                       const XProd &prod, cudaStream_t st) {
    const bool fused = prod.mode != 0;
    auto kern = prod.mode == 0 ? tcq_gemv_kernel<KVA, KVB, S, 0>
                               : (prod.mode == 1 ? tcq_gemv_kernel<KVA, KVB, S, 1> : tcq_gemv_kernel<KVA, KVB, S, 2>);
    const size_t smem = (size_t)TcqTable<S>::kBytes + (size_t)K * bs * 2 + (fused ? (size_t)K * 4 : 0);
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 256, "K = %d does not fit the shared-memory budget of the %s GEMV", K,
                 fused ? "fused-prologue" : "plain");
    static DeviceOnce configured[3];
    if (configured[prod.mode].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 256));
    }
    const int nctas = sm_count() * kGemvCtasPerSM;
    const int nwarps = nctas * kTcqWarps;
    QP_CUDA(launch_pdl(kern, dim3(nctas), dim3(kTcqThreads), smem, st, seg_with_magic(L.a), seg_with_magic(L.b),
                       make_split((long)L.a.strips * L.a.ksuper, nwarps), make_split((long)L.b.strips * L.b.ksuper, nwarps),
                       out, (const uint32_t *)x, (const uint32_t *)tlut, M, K, bs, prod));
    return check_launch("tcq_gemv");
}

template <int KVA, int KVB, int S>
static int launch_dequant(const TcqLaunch &L, __half *W, const void *tlut, int K, cudaStream_t st) {
    auto kern = tcq_dequant_kernel<KVA, KVB, S>;
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    }
    const int nwarps = sm_count() * kTcqWarps;
    kern<<<sm_count(), kTcqThreads, TcqTable<S>::kBytes, st>>>(
        L.a, L.b, make_split((long)L.a.strips * L.a.ksuper, nwarps), make_split((long)L.b.strips * L.b.ksuper, nwarps), W,
        (const uint32_t *)tlut, K);
    return check_launch("tcq_dequant");
}

// (KV1, KV2, S) combinations: every single rate 2..10 with any S in {9,10,11}; two-rate pairs (a, a+1) as the reference
// registers them (lib/linear/__init__.py:166-250) plus the reversed and equal-rate pairs are reachable through the
// generic two-launch fallback below.
#define QP_TCQ_SINGLE(FN, KV, ...)                                          \
    switch (S) {                                                            \
        case 9: return FN<KV, 0, 9>(__VA_ARGS__);                           \
        case 10: return FN<KV, 0, 10>(__VA_ARGS__);                         \
        case 11: return FN<KV, 0, 11>(__VA_ARGS__);                         \
    }                                                                       \
    break;

#define QP_TCQ_PAIR(FN, KA, ...)                                            \
    switch (S) {                                                            \
        case 9: return FN<KA, KA + 1, 9>(__VA_ARGS__);                      \
        case 10: return FN<KA, KA + 1, 10>(__VA_ARGS__);                    \
        case 11: return FN<KA, KA + 1, 11>(__VA_ARGS__);                    \
    }                                                                       \
    break;

static int dispatch_gemv(const TcqLaunch &L, int S, float *out, const void *x, const void *tlut, int M, int K, int bs,
                         const XProd &prod, cudaStream_t st) {
#ifdef QP_FAST_BUILD  // experiments: only the headline instantiation (tcomb_6_7, S = 9), seconds instead of minutes to compile
    if (L.kva == 6 && L.kvb == 7 && S == 9) return launch_gemv<6, 7, 9>(L, out, x, tlut, M, K, bs, prod, st);
    return fail(QP_ERR_ARG, "QP_FAST_BUILD library: only tcomb_6_7 / S = 9");
#else
    if (L.kvb == 0) {
        switch (L.kva) {
            case 2: QP_TCQ_SINGLE(launch_gemv, 2, L, out, x, tlut, M, K, bs, prod, st)
            case 3: QP_TCQ_SINGLE(launch_gemv, 3, L, out, x, tlut, M, K, bs, prod, st)
            case 4: QP_TCQ_SINGLE(launch_gemv, 4, L, out, x, tlut, M, K, bs, prod, st)
            case 5: QP_TCQ_SINGLE(launch_gemv, 5, L, out, x, tlut, M, K, bs, prod, st)
            case 6: QP_TCQ_SINGLE(launch_gemv, 6, L, out, x, tlut, M, K, bs, prod, st)
            case 7: QP_TCQ_SINGLE(launch_gemv, 7, L, out, x, tlut, M, K, bs, prod, st)
            case 8: QP_TCQ_SINGLE(launch_gemv, 8, L, out, x, tlut, M, K, bs, prod, st)
            case 9: QP_TCQ_SINGLE(launch_gemv, 9, L, out, x, tlut, M, K, bs, prod, st)
            case 10: QP_TCQ_SINGLE(launch_gemv, 10, L, out, x, tlut, M, K, bs, prod, st)
        }
    } else if (L.kvb == L.kva + 1) {
        switch (L.kva) {
            case 2: QP_TCQ_PAIR(launch_gemv, 2, L, out, x, tlut, M, K, bs, prod, st)
            case 3: QP_TCQ_PAIR(launch_gemv, 3, L, out, x, tlut, M, K, bs, prod, st)
            case 4: QP_TCQ_PAIR(launch_gemv, 4, L, out, x, tlut, M, K, bs, prod, st)
            case 5: QP_TCQ_PAIR(launch_gemv, 5, L, out, x, tlut, M, K, bs, prod, st)
            case 6: QP_TCQ_PAIR(launch_gemv, 6, L, out, x, tlut, M, K, bs, prod, st)
            case 7: QP_TCQ_PAIR(launch_gemv, 7, L, out, x, tlut, M, K, bs, prod, st)
            case 8: QP_TCQ_PAIR(launch_gemv, 8, L, out, x, tlut, M, K, bs, prod, st)
            case 9: QP_TCQ_PAIR(launch_gemv, 9, L, out, x, tlut, M, K, bs, prod, st)
        }
    } else {
        // arbitrary pair: two single-rate launches accumulating into the same output
        TcqLaunch a = L, b = L;
        a.kvb = 0;
        b.a = L.b;
        b.kva = L.kvb;
        b.kvb = 0;
        QP_CHECK_ARG(prod.mode == 0, "the fused prologue needs a single-launch TCQ configuration");
        int rc = dispatch_gemv(a, S, out, x, tlut, M, K, bs, prod, st);
        if (rc != QP_OK) return rc;
        return dispatch_gemv(b, S, out, x, tlut, M, K, bs, prod, st);
    }
    return fail(QP_ERR_ARG, "unsupported TCQ configuration S=%d KV=(%d,%d)", S, L.kva, L.kvb);
#endif
}

template <int KVA, int KVB, int S>
static int launch_gemm_mma(const TcqLaunch &L, float *out, const uint4 *xfrag, const void *tlut, int M, int bs, cudaStream_t st) {
    const int NB = mma_batch_blocks(bs), v = NB == 4;
    auto kern = v == 0 ? tcq_gemm_mma_kernel<KVA, KVB, S, 2> : tcq_gemm_mma_kernel<KVA, KVB, S, 4>;
    const int threads = v == 0 ? kMmaThreads<2> : kMmaThreads<4>;
    static DeviceOnce configured[2];
    if (configured[v].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 256));
    }
    const int nctas = sm_count();
    const long T = (long)L.a.strips * L.a.ksuper + (long)L.b.strips * L.b.ksuper;
    QP_CUDA(launch_pdl(kern, dim3(nctas), dim3(threads), (size_t)TcqTable<S>::kBytes + kMmaSlabBytes, st, L.a, L.b,
                       make_split(T, nctas), out, xfrag, (const uint32_t *)tlut, M, bs));
    return check_launch("tcq_gemm_mma");
}

// the reference pairs tlut_bits with the rate (9 up to KV = 8, else KV + 1: lib/utils/mem_op.py get_quant_info); other
// combinations fall back to the tcgen05 / dequantise paths on the host side
static int dispatch_gemm_mma(const TcqLaunch &L, int S, float *out, const uint4 *xfrag, const void *tlut, int M, int bs,
                             cudaStream_t st) {
#define QP_MMA1(KV, SS) \
    if (L.kvb == 0 && L.kva == KV && S == SS) return launch_gemm_mma<KV, 0, SS>(L, out, xfrag, tlut, M, bs, st);
#define QP_MMA2(KA, SS) \
    if (L.kva == KA && L.kvb == KA + 1 && S == SS) return launch_gemm_mma<KA, KA + 1, SS>(L, out, xfrag, tlut, M, bs, st);
#ifdef QP_FAST_BUILD
    QP_MMA2(6, 9)
#else
    QP_MMA1(2, 9) QP_MMA1(3, 9) QP_MMA1(4, 9) QP_MMA1(5, 9) QP_MMA1(6, 9) QP_MMA1(7, 9) QP_MMA1(8, 9) QP_MMA1(9, 10) QP_MMA1(10, 11)
    QP_MMA2(2, 9) QP_MMA2(3, 9) QP_MMA2(4, 9) QP_MMA2(5, 9) QP_MMA2(6, 9) QP_MMA2(7, 9) QP_MMA2(8, 10) QP_MMA2(9, 11)
#endif
#undef QP_MMA1
#undef QP_MMA2
    return fail(QP_ERR_ARG, "no mma GEMM instantiation for S=%d KV=(%d,%d)", S, L.kva, L.kvb);
}

static int dispatch_dequant(const TcqLaunch &L, int S, __half *W, const void *tlut, int K, cudaStream_t st) {
    if (L.kvb == 0) {
        switch (L.kva) {
            case 2: QP_TCQ_SINGLE(launch_dequant, 2, L, W, tlut, K, st)
            case 3: QP_TCQ_SINGLE(launch_dequant, 3, L, W, tlut, K, st)
            case 4: QP_TCQ_SINGLE(launch_dequant, 4, L, W, tlut, K, st)
            case 5: QP_TCQ_SINGLE(launch_dequant, 5, L, W, tlut, K, st)
            case 6: QP_TCQ_SINGLE(launch_dequant, 6, L, W, tlut, K, st)
            case 7: QP_TCQ_SINGLE(launch_dequant, 7, L, W, tlut, K, st)
            case 8: QP_TCQ_SINGLE(launch_dequant, 8, L, W, tlut, K, st)
            case 9: QP_TCQ_SINGLE(launch_dequant, 9, L, W, tlut, K, st)
            case 10: QP_TCQ_SINGLE(launch_dequant, 10, L, W, tlut, K, st)
        }
    } else {
        TcqLaunch a = L, b = L;
        a.kvb = 0;
        b.a = L.b;
        b.kva = L.kvb;
        b.kvb = 0;
        int rc = dispatch_dequant(a, S, W, tlut, K, st);
        if (rc != QP_OK) return rc;
        return dispatch_dequant(b, S, W, tlut, K, st);
    }
    return fail(QP_ERR_ARG, "unsupported TCQ configuration S=%d KV=(%d,%d)", S, L.kva, L.kvb);
}

}  // namespace qp

using namespace qp;

extern "C" int qp_tcq_gemv(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16,
                           int M, int K, int bs, int S, int KV1, int KV2, int split_mode, int part1, unsigned flags,
                           void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && x_f16 && tlut_f16, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 8, "bs = %d: the fused GEMV handles 1..8 rows (use the dequant + GEMM path above)", bs);
    QP_CHECK_ARG(S >= 9 && S <= 11, "tlut_bits S = %d not in {9,10,11}", S);
    TcqLaunch L;
    int rc = make_segments(L, codes1, codes2, M, K, KV1, KV2, split_mode, part1);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(codes1, 16, "codes1")) != QP_OK) return rc;
    if (codes2 && (rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    if ((rc = check_align(x_f16, 16, "x")) != QP_OK) return rc;
    if ((rc = check_align(tlut_f16, 4, "tlut")) != QP_OK) return rc;
    // x lives in shared memory next to the 128 KiB codebook: process the batch in chunks that fit
    const size_t avail = (size_t)kMaxSmem - 256 - 128 * 1024;
    int chunk = (int)(avail / ((size_t)K * 2));
    QP_CHECK_ARG(chunk >= 1, "K = %d too large for the shared-memory x stage", K);
    if (chunk > bs) chunk = bs;
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    for (int b0 = 0; b0 < bs; b0 += chunk) {
        const int nb = (bs - b0 < chunk) ? bs - b0 : chunk;
        XProd none = {};
        rc = dispatch_gemv(L, S, out + (size_t)b0 * M, (const __half *)x_f16 + (size_t)b0 * K, tlut_f16, M, K, nb, none, st);
        if (rc != QP_OK) return rc;
    }
    return QP_OK;
}

extern "C" int qp_tcq_dequant(void *W_f16, const void *codes1, const void *codes2, const void *tlut_f16, int M, int K,
                              int S, int KV1, int KV2, int split_mode, int part1, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(W_f16 && tlut_f16, "NULL pointer argument");
    QP_CHECK_ARG(S >= 9 && S <= 11, "tlut_bits S = %d not in {9,10,11}", S);
    TcqLaunch L;
    int rc = make_segments(L, codes1, codes2, M, K, KV1, KV2, split_mode, part1);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(codes1, 16, "codes1")) != QP_OK) return rc;
    if (codes2 && (rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    return dispatch_dequant(L, S, (__half *)W_f16, tlut_f16, K, st);
}

extern "C" int qp_tcq_gemv_host(float *out_host, float *out_dev, const void *codes1, const void *codes2,
                                const void *x_host, void *x_dev, const void *tlut_f16, int M, int K, int bs, int S,
                                int KV1, int KV2, int split_mode, int part1, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out_host && out_dev && x_host && x_dev, "NULL pointer argument");
    QP_CUDA(cudaMemcpyAsync(x_dev, x_host, (size_t)bs * K * 2, cudaMemcpyHostToDevice, st));
    int rc = qp_tcq_gemv(out_dev, codes1, codes2, x_dev, tlut_f16, M, K, bs, S, KV1, KV2, split_mode, part1, 0, stream);
    if (rc != QP_OK) return rc;
    QP_CUDA(cudaMemcpyAsync(out_host, out_dev, (size_t)bs * M * 4, cudaMemcpyDeviceToHost, st));
    return QP_OK;
}

extern "C" size_t qp_gemm_mma_scratch_bytes(int K, int bs) { return gemm_mma_scratch_total(K, bs); }

extern "C" int qp_tcq_gemm_mma(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16,
                               void *scratch, int M, int K, int bs, int S, int KV1, int KV2, int split_mode, int part1,
                               unsigned flags, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && x_f16 && tlut_f16 && scratch, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 128, "bs = %d out of range 1..128", bs);
    TcqLaunch L;
    int rc = make_segments(L, codes1, codes2, M, K, KV1, KV2, split_mode, part1);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(codes1, 16, "codes1")) != QP_OK) return rc;
    if (codes2 && (rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    if ((rc = check_align(x_f16, 4, "x")) != QP_OK) return rc;
    if ((rc = check_align(scratch, 16, "scratch")) != QP_OK) return rc;
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    return mma_gemm_batches(out, x_f16, scratch, M, K, bs, st, [&](float *o, const uint4 *xfrag, int nb) {
        return dispatch_gemm_mma(L, S, o, xfrag, tlut_f16, M, nb, st);
    });
}

#ifdef QP_PROFILE_PHASES
extern "C" int qp_debug_xphases(unsigned long long *host_out /* [256][8] */) {
    QP_CUDA(cudaMemcpyFromSymbol(host_out, qp::g_xphase, sizeof(unsigned long long) * 256 * 8));
    return QP_OK;
}
extern "C" int qp_debug_phases(unsigned long long *host_out /* [256][8] */) {
    QP_CUDA(cudaMemcpyFromSymbol(host_out, qp::g_phase, sizeof(unsigned long long) * 256 * 8));
    return QP_OK;
}
// copies the log (up to `cap` records of 9 words), returns the record count in *n and clears the log
extern "C" int qp_debug_plog(unsigned long long *host_out, unsigned cap, unsigned *n) {
    unsigned cnt = 0, zero = 0;
    QP_CUDA(cudaMemcpyFromSymbol(&cnt, qp::g_plog_n, sizeof(unsigned)));
    if (cnt > qp::kPlogCap) cnt = qp::kPlogCap;
    if (cnt > cap) cnt = cap;
    if (cnt) QP_CUDA(cudaMemcpyFromSymbol(host_out, qp::g_plog, sizeof(unsigned long long) * 9 * cnt));
    QP_CUDA(cudaMemcpyToSymbol(qp::g_plog_n, &zero, sizeof(unsigned)));
    *n = cnt;
    return QP_OK;
}
#endif

// ---- fused prologue entry point -------------------------------------------------------------------------------------
int qp_make_xprod(qp::XProd &p, const qp_xprod *u, int K);

extern "C" int qp_tcq_gemv_fused(float *out, const void *codes1, const void *codes2, const qp_xprod *xp,
                                 const void *tlut_f16, int M, int K, int S, int KV1, int KV2, int split_mode, int part1,
                                 void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && xp && tlut_f16, "NULL pointer argument");
    QP_CHECK_ARG(S >= 9 && S <= 11, "tlut_bits S = %d not in {9,10,11}", S);
    TcqLaunch L;
    int rc = make_segments(L, codes1, codes2, M, K, KV1, KV2, split_mode, part1);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(codes1, 16, "codes1")) != QP_OK) return rc;
    if (codes2 && (rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    XProd p;
    if ((rc = qp_make_xprod(p, xp, K)) != QP_OK) return rc;
    return dispatch_gemv(L, S, out, nullptr, tlut_f16, M, K, 1, p, st);
}
