// gemm_tc_kernels.cu -- fused dequantise + batched GEMM (bs <= 128 per launch) on the 5th-generation tensor cores.
//
//   out[n][m] (+)= sum_k decode(W)[m][k] * x[n][k]
//
// Replaces the reference's bs > 8 path, which materialises the fp16 weight in HBM and calls cuBLAS
// (lib/linear/tcq_linear.py:75-84 `decompress_tcq_*` + `x @ dq.T`; comb_linear.py:100-125, vq_linear.py:58-66): here the
// decoded weights never leave the SM.
//
// Orientation: the batch is the MMA's M (x tile = A operand, 128 rows, rows past the batch are don't-care) and the decoded
// weights are its N (B operand, 256 weight rows).  `tcgen05.mma` costs ~128 cycles per instruction whatever N <= 256 is
// (measured: the first version, weights as the 128-row A operand, was bound by exactly that), so the weights go where one
// instruction covers the most of them.
// One CTA owns a 256-row block of W and a slice of K.  Per 64-column step its 16 decode warps turn 16 packed super-tiles
// (8 strips x 2 columns) into the fp16 tile (256 x 64) in the K-major SWIZZLE_128B shared-memory layout -- a (lane,
// register) of the packed format is one 4-byte word of one 8-row x 16-byte swizzle chunk, so a warp-wide 32-bit store is
// conflict-free and needs no shuffle.  A loader warp drives the TMA engine (packed weights: 3-D tensor map, several steps
// per copy; x tile: 2-D tensor map that swizzles and zero-fills), an MMA warp issues 4 x `tcgen05.mma.cta_group::1
// .kind::f16` (M = 128, N = 256, K = 16) per step into a 256-column TMEM accumulator.  Everything is mbarrier-pipelined.
// Split-K across CTAs fills the 148 SMs; the epilogue (tcgen05.ld -> plain stores, or fp32 atomics when K is split)
// writes `out`.
#include <cuda.h>
#include "gemv_common.cuh"
#include "lut_bits.cuh"

#ifndef QP_TC_TCQ_STRIDE
#define QP_TC_TCQ_STRIDE 6
#endif

namespace qp {

// ---- decoders shared with the GEMV kernels (declared in their translation units; re-declared here as templates) ------
template <int S>
struct GTcqTable {
    static constexpr bool kFold = (S == 9);
    // 64-byte slots (16 lane copies, 2-way bank conflicts) instead of the GEMV's 128: the table shrinks to 64 KiB
    // (S = 11: 32-byte slots, 4-way), which buys the operand stages and the payload ring
    static constexpr int kStrideLog2 = (S == 11) ? QP_TC_TCQ_STRIDE - 1 : QP_TC_TCQ_STRIDE;
    static constexpr int kEntryBits = S + (kFold ? 1 : 0);
    static constexpr int kEntries = 1 << kEntryBits;
    static constexpr int kBytes = kEntries << kStrideLog2;
    static constexpr int kShift = kStrideLog2 - (15 - S);
    static constexpr uint32_t kMask = (uint32_t)(kEntries - 1) << kStrideLog2;
    static constexpr uint32_t kLaneMask = (1u << (kStrideLog2 - 2)) - 1u;
};

template <int S>
__device__ __forceinline__ void g_build_tcq_table(uint32_t *tab, const uint32_t *__restrict__ tlut, int nwarps) {
    using T = GTcqTable<S>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
    constexpr int kRows = T::kBytes / 128;
    for (int r0 = warp * 4; r0 < kRows; r0 += nwarps * 4) {
        const int r = r0 + (lane >> 3);
        // 128-byte row r holds 128 >> kStrideLog2 entries; this lane's 16-byte chunk belongs to entry ef (incl. the fold bit)
        const int ef = (r << (7 - T::kStrideLog2)) + ((lane & 7) >> (T::kStrideLog2 - 4));
        uint32_t v = __ldg(tlut + (ef & ((1 << S) - 1)));
        if (T::kFold && (ef >> S)) v ^= 0x8000u;
        t4[r * 8 + (lane & 7)] = make_uint4(v, v, v, v);
    }
}

template <int KV, int S>
struct GTcqDecoder {
    static constexpr int kE = KV;
    // address = (uniform table base) + ((slot offset) | (lane column)): one LOP3, no per-lookup pointer add
    __device__ static __forceinline__ uint32_t lookup(const uint8_t *tab, uint32_t lane_col, uint32_t u) {
        using T = GTcqTable<S>;
        const uint32_t ts = u * (u * (1u << T::kShift) + (1u << T::kShift));
        uint32_t w = *reinterpret_cast<const uint32_t *>(tab + ((ts & T::kMask) | lane_col));
        if (!T::kFold) w ^= ((ts >> T::kShift) & 0x8000u);
        return w;
    }
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<KV>::kWords], int lane,
                                                  const uint8_t *tab, uint32_t (&frag)[4][4]) {
        using G = TcqGeom<KV>;
        const uint32_t lane_col = ((uint32_t)lane & GTcqTable<S>::kLaneMask) << 2;
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
            for (int j = 0; j < 4; ++j) frag[t][j] = lookup(tab, lane_col, u[t][j]);
    }
};

// ---- tcgen05 / mbarrier wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();  // never hang the GPU on a protocol bug
#ifdef QP_TC_BACKOFF
        if (!ok) __nanosleep(QP_TC_BACKOFF);  // leave the issue slots to the decode warps while waiting
#endif
    }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B operand tiles: row r of a tile is 128 bytes (64 halves); its 16-byte chunk c lives at
//   (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) << 4)
// (8-row x 128-byte atoms, SBO = 1024 B between atoms, LBO unused; tiles are 1024-byte aligned).  A K = 16 slice is two
// chunks, so MMA kk of a step starts 32 * kk bytes into the tile.  The un-swizzled "interleaved" layout measured ~8x
// slower operand fetch (every 128-byte core matrix maps to the same banks).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

#ifdef QP_PROFILE_PHASES
// debug build only: clock64 accounting of CTA 0's roles (read back with qp_debug_tc_prof)
//  [0] decode: cycles waiting on the payload  [1] decode: waiting on a free stage  [2] decode: loop total (per warp sums)
//  [3] MMA: waiting on full  [4] MMA: loop total  [5] loader: waiting on a free slot  [6] loader: loop total
//  [7] prologue (entry -> after table build barrier)  [8] whole kernel  [9] epilogue
__device__ unsigned long long g_tc_prof[16];
#define QP_TC_T(v) const long long v = clock64()
#define QP_TC_ADD(i, v) atomicAdd(&g_tc_prof[i], (unsigned long long)(v))
#else
#define QP_TC_T(v)
#define QP_TC_ADD(i, v)
#endif
constexpr int kTcDecodeWarps = 16;                 // one warp per super-tile of a 256 x 64 step (8 strips x 2 columns)
constexpr int kTcMmaWarp = kTcDecodeWarps;         // one lane issues tcgen05.mma
constexpr int kTcLoaderWarp = kTcDecodeWarps + 1;  // one lane drives the TMA engine: packed weights + x tiles
constexpr int kTcThreads = (kTcDecodeWarps + 2) * 32;
constexpr int kTcMaxStages = 8;                    // operand stages (weight tile + x tile)
constexpr int kTcMaxPSlots = 4;                    // payload ring slots
constexpr int kTileN = 256, kTileK = 64;           // weight rows / columns per step
constexpr int kWBytes = kTileN * kTileK * 2;       // 32 KiB
constexpr int kTmemCols = 256;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// tiled TMA loads; the tensor map carries layout, swizzle and the zero fill of out-of-range coordinates
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// steps of packed weights per payload slot (one TMA copy): 4 while the slot stays <= 32 KiB
template <int E>
struct TcPay {
    static constexpr int kSteps = (E <= 8) ? 4 : 2;
    static constexpr int kStepBytes = 2 * TcqGeom<E>::kSuperBytes;  // one strip, one step
    static constexpr int kStripBytes = kSteps * kStepBytes;
    static constexpr int kSlotBytes = 8 * kStripBytes;              // [strip 8][step][2 super-tiles]
};

// one decode warp: its super-tile (strip s, column c) of NP consecutive steps, read from the payload slots -> weight tiles
// of the steps' stages.  All loads and table lookups of the NP super-tiles come before the first store.
template <class Dec, int NP>
__device__ __forceinline__ void tc_decode_store(const uint8_t *const (&super)[NP], uint8_t *const (&w_stage)[NP], int lane,
                                                const uint8_t *tab_lane, int s, int c) {
    constexpr int E = Dec::kE;
    using G = TcqGeom<E>;
    int w0, bitoff;
    tcq_lane_addr<E>(lane, w0, bitoff);
    uint32_t raw[NP][G::kRawWords];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(super[p]) + w0;
#pragma unroll
        for (int i = 0; i < G::kRawWords; ++i) raw[p][i] = src[i];
    }
    uint32_t frag[NP][4][4];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        uint32_t P[G::kWords];
        tcq_align<E>(raw[p], bitoff, P);
        Dec::decode(P, lane, tab_lane, frag[p]);
    }
    // register (t = kl*2+ml, j) of lane l is W[row 8*rg + l/4][cols 8*kg + 2*(l%4) + {0,1}] with rg = 4s + 2ml + (j&1),
    // kg = 4c + 2kl + (j>>1): 4 bytes at chunk kg of tile row 8*rg + l/4.  A warp-wide store covers 8 rows x 16 B whose
    // swizzled chunks fall in 8 different bank groups: conflict-free.
    const int r8 = lane >> 2;
    const int lane_off = s * 4096 + r8 * 128 + (lane & 3) * 4;
    const int cx = ((4 * c) ^ r8) << 4;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        uint8_t *lane_base = w_stage[p] + lane_off;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int kl = t >> 1, ml = t & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint32_t *>(lane_base + (2 * ml + (j & 1)) * 1024 + (cx ^ ((2 * kl + (j >> 1)) << 4))) = frag[p][t][j];
        }
    }
}

// Warp-specialised pipeline, no CTA-wide barrier inside the K loop:
//   loader warp   one lane drives the TMA engine.  Packed weights: one 3-D tiled copy per payload slot (TcPay::kSteps
//                 k-steps x 8 strips; map dims = [8-byte words of a strip-step][steps][strips]) landing on pfull[slot]
//                 (byte-counted), recycled through pempty[slot]; the ring keeps tens of KiB in flight per SM, which a
//                 register-staged prefetch cannot.  x tiles: one 2-D copy per step straight into the SWIZZLE_128B
//                 operand layout, landing on full[stage].
//   decode warps  every step: read the warp's super-tile from the payload slot, decode it into the swizzled weight tile
//                 of the step's stage, arrive on full[stage].
//   MMA warp      waits full[stage], issues the 4 K=16 MMAs of the step, commits to empty[stage] (stage recycling).
template <class DecA, class DecB, class Table, bool kTwoParts>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap pmapA,
               const __grid_constant__ CUtensorMap pmapB, int stepsA, int stepsB, float *__restrict__ out,
               const void *__restrict__ lut, int lut_arg, int M, int rows, int bs, int npad, int ksplit, int row0, int accumulate,
               int nstages, int npslots_log2) {
    QP_TC_T(t_entry);
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kTcMaxStages], empty_bar[kTcMaxStages], pfull_bar[kTcMaxPSlots],
        pempty_bar[kTcMaxPSlots], done_bar;
    __shared__ uint32_t tmem_base_slot;
    using PA = TcPay<DecA::kE>;
    using PB = TcPay<DecB::kE>;
    constexpr int kSlotBytes = PA::kSlotBytes > PB::kSlotBytes ? PA::kSlotBytes : PB::kSlotBytes;
    const int npslots = 1 << npslots_log2;
    const int x_bytes = npad * kTileK * 2;
    const int stage_bytes = kWBytes + x_bytes;
    uint8_t *tab = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 1024-byte aligned (swizzle atoms)
    uint8_t *stages = tab + ((Table::kSmemBytes + 1023) & ~1023);          // nstages x (W 32 KiB + x npad*128 B)
    uint8_t *pay = stages + nstages * stage_bytes;                         // npslots x kSlotBytes
    const int warp = warp_in_cta(), lane = threadIdx.x & 31;

    // this CTA: 256-row block `mb`, k-steps [st0, st1) of the concatenated parts; local steps [0, bnd) lie in part A.
    // Payload chunks (<= kSteps steps, never across the part boundary): nA over [0, bnd), the rest over [bnd, nsteps).
    const int mb = blockIdx.x / ksplit, ks = blockIdx.x % ksplit;
    const int total_steps = stepsA + stepsB;
    const int st0 = (int)((long)total_steps * ks / ksplit), st1 = (int)((long)total_steps * (ks + 1) / ksplit);
    const int nsteps = st1 - st0;
    const int bnd = kTwoParts ? min(max(stepsA - st0, 0), nsteps) : nsteps;
    const int nA = (bnd + PA::kSteps - 1) / PA::kSteps, nchunks = nA + (nsteps - bnd + PB::kSteps - 1) / PB::kSteps;

    // payload chunk `ch` -> ring slot (one elected lane).  Steps past the part's end / strips past the last row are
    // zero-filled by the TMA engine and still counted, so every copy completes with the full box.
    auto issue_payload = [&](int ch) {
        const int slot = ch & (npslots - 1);
        uint8_t *dst = pay + slot * kSlotBytes;
        if (!kTwoParts || ch < nA) {
            mbar_expect_tx(&pfull_bar[slot], PA::kSlotBytes);
            tma_load_3d(dst, &pmapA, 0, st0 + ch * PA::kSteps, mb * 8, &pfull_bar[slot]);
        } else {
            mbar_expect_tx(&pfull_bar[slot], PB::kSlotBytes);
            tma_load_3d(dst, &pmapB, 0, st0 + bnd + (ch - nA) * PB::kSteps - stepsA, mb * 8, &pfull_bar[slot]);
        }
    };
    auto issue_x = [&](int it) {  // x tile of local step `it` -> its operand stage
        const int stage = it % nstages;
        mbar_expect_tx(&full_bar[stage], (uint32_t)x_bytes);
        tma_load_2d(stages + stage * stage_bytes + kWBytes, &xmap, (st0 + it) * kTileK, 0, &full_bar[stage]);
    };

    if (warp == kTcLoaderWarp && lane == 0) {
        for (int i = 0; i < nstages; ++i) {
            mbar_init(&full_bar[i], kTcDecodeWarps + 1);  // 16 decode warps + the x tile's expect_tx
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < npslots; ++i) {
            mbar_init(&pfull_bar[i], 1);
            mbar_init(&pempty_bar[i], kTcDecodeWarps);    // every decode warp, once per chunk
        }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async_smem();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
        // the weights do not depend on the preceding kernel: fill the ring before the table build / PDL wait
        for (int ch = 0; ch < min(nchunks, npslots); ++ch) issue_payload(ch);
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, kTmemCols);
    Table::build(reinterpret_cast<uint32_t *>(tab), lut, lut_arg, kTcThreads / 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint8_t *tab_lane = tab;  // decoders add the lane column themselves (OR into the slot offset)
    pdl_wait();  // x / out come from the preceding kernel
    pdl_launch_dependents();
    QP_TC_T(t_loop);
#ifdef QP_PROFILE_PHASES
    long long w_a = 0, w_b = 0;
    const bool rec = blockIdx.x == 0 && lane == 0;
#endif

    if (warp < kTcDecodeWarps) {
        // ---------------------------------------------------------------- weight tile producers (decode)
        const int s = warp >> 1, c = warp & 1;  // strip / super-tile column of this warp
        struct Step {
            bool inA, last;      // part; last step of its payload chunk
            int pslot, pround;   // payload ring slot and its pass
            const uint8_t *super;  // this warp's super-tile in the slot
        };
        auto step_info = [&](int it) {
            Step r;
            r.inA = !kTwoParts || it < bnd;
            const int l = r.inA ? it : it - bnd;  // step within its part's range
            const int psteps = r.inA ? PA::kSteps : PB::kSteps;
            const int ch = (r.inA ? 0 : nA) + l / psteps, j = l % psteps;  // payload chunk, step within it
            r.last = j == psteps - 1 || it == bnd - 1 || it == nsteps - 1;
            r.pslot = ch & (npslots - 1), r.pround = ch >> npslots_log2;
            const uint8_t *slot = pay + r.pslot * kSlotBytes;
            r.super = r.inA ? slot + s * PA::kStripBytes + j * PA::kStepBytes + c * TcqGeom<DecA::kE>::kSuperBytes
                            : slot + s * PB::kStripBytes + j * PB::kStepBytes + c * TcqGeom<DecB::kE>::kSuperBytes;
            return r;
        };
        int stage = 0, round = 0;
        for (int it = 0; it < nsteps; ++it) {
            const Step a0 = step_info(it);
            uint8_t *w0 = stages + stage * stage_bytes;
            QP_TC_T(c0);
            mbar_wait(&pfull_bar[a0.pslot], (uint32_t)(a0.pround & 1));
            QP_TC_T(c1);
            if (round > 0) mbar_wait(&empty_bar[stage], (uint32_t)((round - 1) & 1));
#ifdef QP_PROFILE_PHASES
            w_a += c1 - c0, w_b += clock64() - c1;
#endif
            // (decoding two steps per iteration for more independent work per warp measured slower: 39 vs 32 us)
            const uint8_t *const sup[1] = {a0.super};
            uint8_t *const wst[1] = {w0};
            if (a0.inA) tc_decode_store<DecA, 1>(sup, wst, lane, tab_lane, s, c);
            else tc_decode_store<DecB, 1>(sup, wst, lane, tab_lane, s, c);
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                if (a0.last) mbar_arrive(&pempty_bar[a0.pslot]);  // this warp is done with the chunk's slot
                mbar_arrive(&full_bar[stage]);
            }
            if (++stage == nstages) stage = 0, ++round;
        }
#ifdef QP_PROFILE_PHASES
        if (rec) { QP_TC_ADD(0, w_a); QP_TC_ADD(1, w_b); QP_TC_ADD(2, clock64() - t_loop); }
#endif
    } else if (warp == kTcLoaderWarp) {
        // ---------------------------------------------------------------- TMA driver
        if (lane == 0) {
            for (int it = 0; it < min(nsteps, nstages); ++it) issue_x(it);
            for (int it = 0; it < nsteps; ++it) {
                // x tile of step it + nstages reuses this step's stage: wait for the MMAs of step `it`
                if (it + nstages < nsteps) {
                    QP_TC_T(c0);
                    mbar_wait(&empty_bar[it % nstages], (uint32_t)((it / nstages) & 1));
#ifdef QP_PROFILE_PHASES
                    w_a += clock64() - c0;
#endif
                    issue_x(it + nstages);
                }
                // last step of payload chunk ch: once all its readers are done, refill the slot with chunk ch + npslots
                const bool inA = !kTwoParts || it < bnd;
                const int l = inA ? it : it - bnd;
                const int psteps = inA ? PA::kSteps : PB::kSteps;
                const int ch = (inA ? 0 : nA) + l / psteps;
                const bool last = (l % psteps == psteps - 1) || it == bnd - 1 || it == nsteps - 1;
                if (last && ch + npslots < nchunks) {
                    QP_TC_T(c0);
                    mbar_wait(&pempty_bar[ch & (npslots - 1)], (uint32_t)((ch >> npslots_log2) & 1));
#ifdef QP_PROFILE_PHASES
                    w_b += clock64() - c0;
#endif
                    issue_payload(ch + npslots);
                }
            }
        }
#ifdef QP_PROFILE_PHASES
        if (rec) { QP_TC_ADD(5, w_a); QP_TC_ADD(11, w_b); QP_TC_ADD(6, clock64() - t_loop); }
#endif
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- MMA issuer: D[batch][w row] += x . W^T
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int stage = 0, round = 0;
        for (int it = 0; it < nsteps; ++it) {
            QP_TC_T(c0);
            mbar_wait(&full_bar[stage], (uint32_t)(round & 1));
#ifdef QP_PROFILE_PHASES
            w_a += clock64() - c0;
#endif
            tc_fence_after();
            if (lane == 0) {
                const uint32_t w_addr = smem_u32(stages + stage * stage_bytes), x_addr = w_addr + kWBytes;
                // the x tile holds npad rows; the MMA reads 128: rows past npad are whatever follows in shared memory and
                // only reach accumulator rows >= npad, which the epilogue never reads
#pragma unroll
                for (int kk = 0; kk < kTileK / 16; ++kk)
                    umma_f16(tmem_d, make_smem_desc(x_addr + kk * 32), make_smem_desc(w_addr + kk * 32), idesc,
                             (it > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);  // frees the stage once the MMAs above have read it
                if (it == nsteps - 1) umma_commit(&done_bar);
            }
            __syncwarp();
            if (++stage == nstages) stage = 0, ++round;
        }
#ifdef QP_PROFILE_PHASES
        if (rec) { QP_TC_ADD(3, w_a); QP_TC_ADD(4, clock64() - t_loop); }
#endif
    }
    QP_TC_T(t_epi);

    // ---------------------------------------------------------------- epilogue: TMEM (lane = batch row, column = W row) -> out
    if (warp < kTcDecodeWarps && nsteps > 0) {
        const int q = warp & 3, cg = warp >> 2;  // TMEM lane quarter this warp may read; 64-column group
        const int n = 32 * q + lane;
        if (32 * q < bs) {
            mbar_wait(&done_bar, 0u);
            tc_fence_after();
            const int tile_rows = min(kTileN, rows - mb * kTileN);
            float *dst = out + (size_t)n * M + row0 + mb * kTileN;
#pragma unroll 2
            for (int c0 = cg * 64; c0 < cg * 64 + 64; c0 += 8) {
                if (c0 < tile_rows) {  // rows % 32 == 0, so a group of 8 columns is all in or all out
                    uint32_t r[8];
                    tmem_ld_x8(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, r);
                    if (n < bs) {
                        if (ksplit > 1) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0), "r"(r[0]), "r"(r[1]),
                                         "r"(r[2]), "r"(r[3])
                                         : "memory");
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4), "r"(r[4]), "r"(r[5]),
                                         "r"(r[6]), "r"(r[7])
                                         : "memory");
                        } else {
                            float4 v0 = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                            float4 v1 = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                            float4 *d4 = reinterpret_cast<float4 *>(dst + c0);
                            if (accumulate) {
                                const float4 o0 = d4[0], o1 = d4[1];
                                v0.x += o0.x, v0.y += o0.y, v0.z += o0.z, v0.w += o0.w;
                                v1.x += o1.x, v1.y += o1.y, v1.z += o1.z, v1.w += o1.w;
                            }
                            d4[0] = v0;  // sole owner of (row block, all of K): plain stores, no memset needed
                            d4[1] = v1;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kTmemCols);
#ifdef QP_PROFILE_PHASES
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        QP_TC_ADD(7, t_loop - t_entry);
        QP_TC_ADD(8, clock64() - t_entry);
        QP_TC_ADD(9, clock64() - t_epi);
        QP_TC_ADD(10, 1);
    }
#endif
}

// ---- table policies ----------------------------------------------------------------------------------------------------
template <int S>
struct TcTcqTable {
    static constexpr int kSmemBytes = GTcqTable<S>::kBytes;
    static constexpr uint32_t kLaneMask = GTcqTable<S>::kLaneMask;
    __device__ static __forceinline__ void build(uint32_t *tab, const void *lut, int, int nwarps) {
        g_build_tcq_table<S>(tab, reinterpret_cast<const uint32_t *>(lut), nwarps);
    }
};

// VQ (vec 2, E = bits) and SQ with bits <= 5 (pair table, E = 2*bits): one lookup per pair
template <int E>
struct GLutTable {
    static constexpr int kSL = (E <= 9) ? 7 : (16 - E);  // table <= 64 KiB: leaves room for the operand stages
    static constexpr int kEntries = 1 << E;
    static constexpr int kSmemBytes = kEntries << kSL;
    static constexpr uint32_t kLaneMask = (1u << (kSL - 2)) - 1u;
    // r_single = 0: lut is (2^E, 2) fp16; else (2^r_single, 1) fp16 and the entry is {lut[c0], lut[c1]}
    __device__ static __forceinline__ void build(uint32_t *tab, const void *lut, int r_single, int) {
        constexpr int copies = 1 << (kSL - 2);
        const uint32_t *l32 = reinterpret_cast<const uint32_t *>(lut);
        const uint16_t *l16 = reinterpret_cast<const uint16_t *>(lut);
        for (int i = threadIdx.x; i < kEntries * copies; i += blockDim.x) {
            const int e = i / copies;
            uint32_t v;
            if (r_single == 0) v = __ldg(l32 + e);
            else v = (uint32_t)l16[e & ((1 << r_single) - 1)] | ((uint32_t)l16[e >> r_single] << 16);
            tab[i] = v;
        }
    }
};
template <int E>
struct GLutDecoder {
    static constexpr int kE = E;
    template <int TI>
    __device__ static __forceinline__ void tile(const uint32_t (&P)[TcqGeom<E>::kWords], const uint8_t *tab, uint32_t lc,
                                                uint32_t (&f)[4]) {
        constexpr int SL = GLutTable<E>::kSL;
        f[0] = *reinterpret_cast<const uint32_t *>(tab + (lut_pair_offset<E, TI, 0, SL>(P) | lc));
        f[1] = *reinterpret_cast<const uint32_t *>(tab + (lut_pair_offset<E, TI, 1, SL>(P) | lc));
        f[2] = *reinterpret_cast<const uint32_t *>(tab + (lut_pair_offset<E, TI, 2, SL>(P) | lc));
        f[3] = *reinterpret_cast<const uint32_t *>(tab + (lut_pair_offset<E, TI, 3, SL>(P) | lc));
    }
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<E>::kWords], int lane, const uint8_t *tab,
                                                  uint32_t (&frag)[4][4]) {
        const uint32_t lc = ((uint32_t)lane & GLutTable<E>::kLaneMask) << 2;
        tile<0>(P, tab, lc, frag[0]);
        tile<1>(P, tab, lc, frag[1]);
        tile<2>(P, tab, lc, frag[2]);
        tile<3>(P, tab, lc, frag[3]);
    }
};

struct TcPart {
    const uint32_t *codes;  // packed words; rows = the launch's rows, cols = this part's columns
    int ksuper;             // part columns / 32
    int steps;              // part columns / 64
};

// ---- host ----------------------------------------------------------------------------------------------------------------
// tensor map of x (bs, K) fp16 row-major: box = 64 columns x npad rows, SWIZZLE_128B, out-of-range rows read as zero
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int get_encode(EncodeTiledFn *out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        QP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(QP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiledFn)fn;
    }
    *out = encode;
    return QP_OK;
}
static int make_x_map(CUtensorMap *map, const void *x, int bs, int K, int npad) {
    EncodeTiledFn encode = nullptr;
    int rc0 = get_encode(&encode);
    if (rc0 != QP_OK) return rc0;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)bs};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTileK, (cuuint32_t)npad};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for x (%d x %d)", (int)r, bs, K);
    return QP_OK;
}

// tensor map of one part's packed weights, viewed as [strip][k-step][8-byte words of a strip-step] (innermost last):
// box = one payload slot (8 strips x kSteps steps x all words)
template <int E>
static int make_payload_map(CUtensorMap *map, const TcPart &p, int rows) {
    const cuuint64_t dims[3] = {(cuuint64_t)16 * E, (cuuint64_t)p.steps, (cuuint64_t)(rows / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)TcPay<E>::kStepBytes, (cuuint64_t)p.ksuper * TcqGeom<E>::kSuperBytes};
    const cuuint32_t box[3] = {(cuuint32_t)16 * E, (cuuint32_t)TcPay<E>::kSteps, 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    EncodeTiledFn encode = nullptr;
    int rc = get_encode(&encode);
    if (rc != QP_OK) return rc;
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint32_t *>(p.codes), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the packed weights (E = %d)", (int)r, E);
    return QP_OK;
}

template <class DecA, class DecB, class Table>
static int launch_tc(TcPart a, TcPart b, float *out, const void *x, const void *lut, int lut_arg, int M, int K, int bs,
                     int rows, int row0, bool accumulate, cudaStream_t st) {
    constexpr bool kTwoParts = DecA::kE != DecB::kE;
    auto kern = gemm_tc_kernel<DecA, DecB, Table, kTwoParts>;
    const int npad = (bs + 15) & ~15;
    const size_t stage_bytes = kWBytes + (size_t)npad * kTileK * 2;
    const size_t fixed = 1024 + (size_t)((Table::kSmemBytes + 1023) & ~1023);  // alignment slack + table
    // payload ring: 4 slots (2 when shared memory is short); operand stages take what is left
    const size_t slot_bytes = (size_t)(TcPay<DecA::kE>::kSlotBytes > TcPay<DecB::kE>::kSlotBytes ? TcPay<DecA::kE>::kSlotBytes
                                                                                                  : TcPay<DecB::kE>::kSlotBytes);
    // the MMA reads 128 x-tile rows whatever npad is: keep (128 - npad) * 128 bytes of slack after the last stage's x tile
    // inside the allocation (the payload ring normally provides it)
    const size_t avail = (size_t)kMaxSmem - 2048 - fixed;
    auto stages_for = [&](int np) { return avail > np * slot_bytes ? (int)((avail - np * slot_bytes) / stage_bytes) : 0; };
    int np_log2 = 2;
    if (stages_for(4) < 3) np_log2 = 1;
    int nstages = stages_for(1 << np_log2);
    if (nstages > kTcMaxStages) nstages = kTcMaxStages;
    if (nstages < 2) {
        // the table leaves too little room for this batch width: run it as two narrower launches
        QP_CHECK_ARG(bs > 16, "bs = %d does not fit the shared-memory stages", bs);
        const int half = ((bs + 1) / 2 + 15) & ~15;
        const int rc = launch_tc<DecA, DecB, Table>(a, b, out, x, lut, lut_arg, M, K, half, rows, row0, accumulate, st);
        if (rc != QP_OK) return rc;
        return launch_tc<DecA, DecB, Table>(a, b, out + (size_t)half * M, (const __half *)x + (size_t)half * K, lut, lut_arg, M, K,
                                            bs - half, rows, row0, accumulate, st);
    }
    size_t smem = fixed + (size_t)nstages * stage_bytes + ((size_t)slot_bytes << np_log2);
    if (((size_t)slot_bytes << np_log2) < (size_t)(128 - npad) * 128) smem += (size_t)(128 - npad) * 128;
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048));
    }
    CUtensorMap xmap, pmapA, pmapB;
    int rc = make_x_map(&xmap, x, bs, K, npad);
    if (rc != QP_OK) return rc;
    if ((rc = make_payload_map<DecA::kE>(&pmapA, a, rows)) != QP_OK) return rc;
    if (b.steps) {
        if ((rc = make_payload_map<DecB::kE>(&pmapB, b, rows)) != QP_OK) return rc;
    } else {
        pmapB = pmapA;
    }
    const int tiles = (rows + kTileN - 1) / kTileN;
    const int total_steps = a.steps + b.steps;
    // split K as far as it fills the SMs without a second wave; every split costs the epilogue fp32 atomics
    int ksplit = sm_count() / tiles;
    if (ksplit > total_steps) ksplit = total_steps;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 1 && !accumulate)  // split-K partial tiles are added atomically: clear this launch's rows first
        QP_CUDA(cudaMemset2DAsync(out + row0, (size_t)M * sizeof(float), 0, (size_t)rows * sizeof(float), (size_t)bs, st));
    QP_CUDA(launch_pdl(kern, dim3(tiles * ksplit), dim3(kTcThreads), smem, st, xmap, pmapA, pmapB, a.steps, b.steps, out, lut,
                       lut_arg, M, rows, bs, npad, ksplit, row0, (int)accumulate, nstages, np_log2));
    return check_launch("gemm_tc");
}

template <int KVA, int KVB, int S>
static int launch_tc_tcq(TcPart a, TcPart b, float *out, const void *x, const void *tlut, int M, int K, int bs, int rows,
                         int row0, bool acc, cudaStream_t st) {
    using DA = GTcqDecoder<KVA, S>;
    using DB = GTcqDecoder<(KVB ? KVB : KVA), S>;
    return launch_tc<DA, DB, TcTcqTable<S>>(a, b, out, x, tlut, 0, M, K, bs, rows, row0, acc, st);
}

#define QP_TC_S(FN, KA, KB, ...)                              \
    switch (S) {                                              \
        case 9: return FN<KA, KB, 9>(__VA_ARGS__);            \
        case 10: return FN<KA, KB, 10>(__VA_ARGS__);          \
        case 11: return FN<KA, KB, 11>(__VA_ARGS__);          \
    }                                                         \
    break;

static int dispatch_tc_tcq(int S, int kva, int kvb, TcPart a, TcPart b, float *out, const void *x, const void *tlut, int M,
                           int K, int bs, int rows, int row0, bool acc, cudaStream_t st) {
    if (kvb == 0) {
        switch (kva) {
            case 2: QP_TC_S(launch_tc_tcq, 2, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 3: QP_TC_S(launch_tc_tcq, 3, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 4: QP_TC_S(launch_tc_tcq, 4, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 5: QP_TC_S(launch_tc_tcq, 5, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 6: QP_TC_S(launch_tc_tcq, 6, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 7: QP_TC_S(launch_tc_tcq, 7, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 8: QP_TC_S(launch_tc_tcq, 8, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 9: QP_TC_S(launch_tc_tcq, 9, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 10: QP_TC_S(launch_tc_tcq, 10, 0, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
        }
    } else if (kvb == kva + 1) {
        switch (kva) {
            case 2: QP_TC_S(launch_tc_tcq, 2, 3, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 3: QP_TC_S(launch_tc_tcq, 3, 4, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 4: QP_TC_S(launch_tc_tcq, 4, 5, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 5: QP_TC_S(launch_tc_tcq, 5, 6, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 6: QP_TC_S(launch_tc_tcq, 6, 7, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 7: QP_TC_S(launch_tc_tcq, 7, 8, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 8: QP_TC_S(launch_tc_tcq, 8, 9, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
            case 9: QP_TC_S(launch_tc_tcq, 9, 10, a, b, out, x, tlut, M, K, bs, rows, row0, acc, st)
        }
    }
    return fail(QP_ERR_ARG, "unsupported TCQ configuration S=%d KV=(%d,%d) for the tensor-core GEMM", S, kva, kvb);
}

template <int E>
static int launch_tc_lut(TcPart a, float *out, const void *x, const void *lut, int r_single, int M, int K, int bs,
                         bool acc, cudaStream_t st) {
    TcPart none{nullptr, 0, 0};
    return launch_tc<GLutDecoder<E>, GLutDecoder<E>, GLutTable<E>>(a, none, out, x, lut, r_single, M, K, bs, M, 0, acc, st);
}

}  // namespace qp

using namespace qp;

#ifdef QP_PROFILE_PHASES
extern "C" int qp_debug_tc_prof(unsigned long long *host_out /* [16] */, int reset) {
    QP_CUDA(cudaMemcpyFromSymbol(host_out, qp::g_tc_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        QP_CUDA(cudaMemcpyToSymbol(qp::g_tc_prof, z, sizeof(z)));
    }
    return QP_OK;
}
#endif

extern "C" int qp_tcq_gemm_tc(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16,
                              int M, int K, int bs, int S, int KV1, int KV2, int split_mode, int part1, unsigned flags,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && codes1 && x_f16 && tlut_f16, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 128, "bs = %d: the tensor-core GEMM handles up to 128 rows per call", bs);
    QP_CHECK_ARG(S >= 9 && S <= 11, "tlut_bits S = %d not in {9,10,11}", S);
    QP_CHECK_ARG(M % 128 == 0 && K % 64 == 0, "tensor-core GEMM needs M %% 128 == 0 and K %% 64 == 0 (got %d x %d)", M, K);
    int rc;
    if ((rc = check_align(codes1, 16, "codes1")) != QP_OK) return rc;
    if ((rc = check_align(x_f16, 16, "x")) != QP_OK) return rc;
    const bool acc = (flags & QP_FLAG_ACCUMULATE) != 0;
    if (split_mode == QP_SPLIT_NONE) {
        TcPart a{(const uint32_t *)codes1, K / 32, K / 64}, none{nullptr, 0, 0};
        return dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, M, 0, acc, st);
    }
    QP_CHECK_ARG(codes2 != nullptr, "codes2 is NULL for a two-rate layer");
    if ((rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    if (split_mode == QP_SPLIT_IN) {
        QP_CHECK_ARG(part1 > 0 && part1 < K && part1 % 64 == 0 && (K - part1) % 64 == 0, "in_part boundary must be a multiple of 64");
        TcPart a{(const uint32_t *)codes1, part1 / 32, part1 / 64}, b{(const uint32_t *)codes2, (K - part1) / 32, (K - part1) / 64};
        if (KV2 == KV1 + 1) return dispatch_tc_tcq(S, KV1, KV2, a, b, out, x_f16, tlut_f16, M, K, bs, M, 0, acc, st);
        // arbitrary rate pair: two accumulating launches over the column halves
        TcPart none{nullptr, 0, 0};
        rc = dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, M, 0, acc, st);
        if (rc != QP_OK) return rc;
        return fail(QP_ERR_ARG, "combt with KV2 != KV1 + 1 is not supported by the tensor-core GEMM");
    }
    if (split_mode == QP_SPLIT_OUT) {
        QP_CHECK_ARG(part1 > 0 && part1 < M && part1 % 128 == 0 && (M - part1) % 128 == 0, "out_part boundary must be a multiple of 128");
        TcPart a{(const uint32_t *)codes1, K / 32, K / 64}, b{(const uint32_t *)codes2, K / 32, K / 64}, none{nullptr, 0, 0};
        rc = dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, part1, 0, acc, st);
        if (rc != QP_OK) return rc;
        return dispatch_tc_tcq(S, KV2, 0, b, none, out, x_f16, tlut_f16, M, K, bs, M - part1, part1, acc, st);
    }
    return fail(QP_ERR_ARG, "unknown split_mode %d", split_mode);
}

extern "C" int qp_lut_gemm_tc(float *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs,
                              int bits, int vec_sz, unsigned flags, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && codes && x_f16 && lut_f16, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 128, "bs = %d: the tensor-core GEMM handles up to 128 rows per call", bs);
    QP_CHECK_ARG(M % 128 == 0 && K % 64 == 0, "tensor-core GEMM needs M %% 128 == 0 and K %% 64 == 0 (got %d x %d)", M, K);
    QP_CHECK_ARG((vec_sz == 2 && bits >= 2 && bits <= 12) || (vec_sz == 1 && bits >= 2 && bits <= 5),
                 "tensor-core GEMM supports vq2 (2..12 bits) and SQ up to 5 bits (got bits=%d vec_sz=%d)", bits, vec_sz);
    int rc;
    if ((rc = check_align(codes, 16, "codes")) != QP_OK) return rc;
    if ((rc = check_align(x_f16, 16, "x")) != QP_OK) return rc;
    const bool acc = (flags & QP_FLAG_ACCUMULATE) != 0;
    TcPart a{(const uint32_t *)codes, K / 32, K / 64};
    const int E = vec_sz == 2 ? bits : 2 * bits;
    const int r_single = vec_sz == 1 ? bits : 0;
    switch (E) {
#define QP_C(e) case e: return launch_tc_lut<e>(a, out, x_f16, lut_f16, r_single, M, K, bs, acc, st);
        QP_C(2) QP_C(3) QP_C(4) QP_C(5) QP_C(6) QP_C(7) QP_C(8) QP_C(9) QP_C(10) QP_C(11) QP_C(12)
#undef QP_C
    }
    return fail(QP_ERR_ARG, "unsupported configuration");
}
