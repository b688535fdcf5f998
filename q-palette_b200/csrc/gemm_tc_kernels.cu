// gemm_tc_kernels.cu -- fused dequantise + batched GEMM (16 <= bs <= 128) on the 5th-generation tensor cores.
//
//   out[n][m] += sum_k decode(W)[m][k] * x[n][k]
//
// Replaces the reference's bs > 8 path, which materialises the fp16 weight in HBM and calls cuBLAS
// (lib/linear/tcq_linear.py:75-84 `decompress_tcq_*` + `x @ dq.T`; comb_linear.py:100-125, vq_linear.py:58-66): here the
// decoded weights never leave the SM.
//
// One CTA owns a 128-row block of W and a slice of K.  Per 64-column step its 8 decode warps turn 8 packed super-tiles
// (4 strips x 2 columns) into the fp16 A tile (128 x 64) directly in the UMMA canonical K-major / no-swizzle shared-memory
// layout -- a (lane, register) of the packed format is exactly one 4-byte word of one 8x8 "core matrix", so a warp-wide
// 32-bit store fills a core matrix with no bank conflict and no shuffle.  Warp 8 copies the x tile (N x 64) into the same
// canonical layout; after a CTA barrier one thread issues 4 x `tcgen05.mma.cta_group::1.kind::f16` (M = 128, N = bs rounded
// up to 16, K = 16) accumulating in TMEM and commits to the stage's mbarrier.  Two smem stages: the decode of step i+1
// overlaps the tensor-core work of step i.  Split-K across CTAs fills the 148 SMs; the epilogue (tcgen05.ld -> fp32
// atomics) adds the partial tile to `out`.
// The op stays decode/HBM-bound up to bs in the hundreds; tensor-pipe utilisation is low by design.
#include "gemv_common.cuh"
#include "lut_bits.cuh"

namespace qp {

// ---- decoders shared with the GEMV kernels (declared in their translation units; re-declared here as templates) ------
template <int S>
struct GTcqTable {
    static constexpr bool kFold = (S == 9);
    static constexpr int kStrideLog2 = (S == 11) ? 6 : 7;
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
        int e;
        if (T::kStrideLog2 == 7) e = r & ((1 << S) - 1);
        else e = 2 * r + ((lane >> 2) & 1);
        uint32_t v = __ldg(tlut + e);
        if (T::kFold && (r >> S)) v ^= 0x8000u;
        t4[r * 8 + (lane & 7)] = make_uint4(v, v, v, v);
    }
}

template <int KV, int S>
struct GTcqDecoder {
    static constexpr int kE = KV;
    __device__ static __forceinline__ uint32_t lookup(const uint8_t *tab_lane, uint32_t u) {
        using T = GTcqTable<S>;
        const uint32_t ts = u * (u * (1u << T::kShift) + (1u << T::kShift));
        uint32_t w = *reinterpret_cast<const uint32_t *>(tab_lane + (ts & T::kMask));
        if (!T::kFold) w ^= ((ts >> T::kShift) & 0x8000u);
        return w;
    }
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<KV>::kWords], int lane,
                                                  const uint8_t *tab_lane, uint32_t (&frag)[4][4]) {
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
            for (int j = 0; j < 4; ++j) frag[t][j] = lookup(tab_lane, u[t][j]);
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

// K-major, no-swizzle canonical layout (cute UMMA: ((8,n),2):((1,SBO),LBO) in 16-byte units): core matrix = 8 rows x 16 B,
// here stored as tile[row_group][k_group] of 128-byte core matrices => LBO = 128 B (next k group), SBO = 1024 B (next 8 rows)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46);
}

constexpr int kTcDecodeWarps = 8;
constexpr int kTcThreads = (kTcDecodeWarps + 1) * 32;  // + 1 loader / MMA-issue warp
constexpr int kTcStages = 2;
constexpr int kTileM = 128, kTileK = 64;
constexpr int kABytes = kTileM * kTileK * 2;  // 16 KiB

struct TcPart {
    const uint32_t *codes;  // packed words; rows = all M rows of the layer, cols = this part's columns
    int ksuper;             // part columns / 32
    int steps;              // part columns / 64
};

// one decode warp: super-tile (strip s, column c) of the step -> 16 core matrices of the A stage
template <class Dec>
__device__ __forceinline__ void tc_decode_store(const uint32_t (&raw)[TcqGeom<Dec::kE>::kRawWords], int bitoff, int lane,
                                                const uint8_t *tab_lane, uint8_t *a_stage, int s, int c) {
    constexpr int E = Dec::kE;
    uint32_t P[TcqGeom<E>::kWords];
    tcq_align<E>(raw, bitoff, P);
    uint32_t frag[4][4];
    Dec::decode(P, lane, tab_lane, frag);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int kl = t >> 1, ml = t & 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rg = 4 * s + 2 * ml + (j & 1);   // 8-row group of the 128-row tile
            const int kg = 4 * c + 2 * kl + (j >> 1);  // 8-column group of the 64-column tile
            *reinterpret_cast<uint32_t *>(a_stage + (rg * 8 + kg) * 128 + lane * 4) = frag[t][j];
        }
    }
}

template <class DecA, class DecB, class Table>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(TcPart partA, TcPart partB, float *__restrict__ out, const __half *__restrict__ x, const void *__restrict__ lut,
               int lut_arg, int M, int K, int bs, int npad, int ksplit, int row0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[kTcStages];
    __shared__ uint32_t tmem_base_slot;
    uint8_t *tab = smem;
    uint8_t *stages = smem + Table::kSmemBytes;  // kTcStages x (A 16 KiB + B npad*128 B)
    const int b_bytes = npad * kTileK * 2;
    const int stage_bytes = kABytes + b_bytes;
    const int warp = warp_in_cta(), lane = threadIdx.x & 31;

    // this CTA: 128-row block `mb`, k-steps [st0, st1) of the concatenated parts
    const int mb = blockIdx.x / ksplit, ks = blockIdx.x % ksplit;
    const int total_steps = partA.steps + partB.steps;
    const int st0 = (int)((long)total_steps * ks / ksplit), st1 = (int)((long)total_steps * (ks + 1) / ksplit);

    if (warp == kTcDecodeWarps && lane == 0) {
        for (int i = 0; i < kTcStages; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)(npad < 32 ? 32 : (npad <= 64 ? 64 : 128)));
    Table::build(reinterpret_cast<uint32_t *>(tab), lut, lut_arg, kTcThreads / 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint8_t *tab_lane = tab + ((lane & Table::kLaneMask) << 2);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

    // decode warp w handles strip s = w / 2, column c = w % 2 of every step
    const int s = warp >> 1, c = warp & 1;
    constexpr int EA = DecA::kE, EB = DecB::kE;
    int w0A, boA, w0B, boB;
    tcq_lane_addr<EA>(lane, w0A, boA);
    tcq_lane_addr<EB>(lane, w0B, boB);
    auto payload_ptr_A = [&](int step) {
        return partA.codes + w0A + ((size_t)(mb * 4 + s) * partA.ksuper + (size_t)(2 * step + c)) * (TcqGeom<EA>::kSuperBytes / 4);
    };
    auto payload_ptr_B = [&](int step) {
        return partB.codes + w0B + ((size_t)(mb * 4 + s) * partB.ksuper + (size_t)(2 * (step - partA.steps) + c)) * (TcqGeom<EB>::kSuperBytes / 4);
    };
    uint32_t rawA[kTcStages][TcqGeom<EA>::kRawWords], rawB[kTcStages][TcqGeom<EB>::kRawWords];
#pragma unroll
    for (int d = 0; d < kTcStages; ++d) {
#pragma unroll
        for (int i = 0; i < TcqGeom<EA>::kRawWords; ++i) rawA[d][i] = 0u;
#pragma unroll
        for (int i = 0; i < TcqGeom<EB>::kRawWords; ++i) rawB[d][i] = 0u;
    }
    if (warp < kTcDecodeWarps) {
#pragma unroll
        for (int d = 0; d < kTcStages; ++d) {
            const int st = st0 + d;
            pack_load_raw_pred<EA>(rawA[d], payload_ptr_A(st), st < st1 && st < partA.steps);
            if (partB.steps) pack_load_raw_pred<EB>(rawB[d], payload_ptr_B(st), st < st1 && st >= partA.steps);
        }
    }
    pdl_wait();  // x / out come from the preceding kernel
    pdl_launch_dependents();

    for (int base = st0; base < st1; base += kTcStages) {
#pragma unroll
        for (int d = 0; d < kTcStages; ++d) {
            const int st = base + d;
            if (st < st1) {  // CTA-uniform
                const int it = st - st0;
                uint8_t *a_stage = stages + d * stage_bytes;
                uint8_t *b_stage = a_stage + kABytes;
                // the tensor core must be done reading this stage (commit of step st - kTcStages)
                if (it >= kTcStages) mbar_wait(&bar[d], (uint32_t)(((it / kTcStages) - 1) & 1));
                if (warp < kTcDecodeWarps) {
                    if (st < partA.steps) {
                        tc_decode_store<DecA>(rawA[d], boA, lane, tab_lane, a_stage, s, c);
                    } else {
                        tc_decode_store<DecB>(rawB[d], boB, lane, tab_lane, a_stage, s, c);
                    }
                    const int nx = st + kTcStages;
                    pack_load_raw_pred<EA>(rawA[d], payload_ptr_A(nx), nx < st1 && nx < partA.steps);
                    if (partB.steps) pack_load_raw_pred<EB>(rawB[d], payload_ptr_B(nx), nx < st1 && nx >= partA.steps);
                } else {
                    // x tile: rows n < npad, columns [64*st, 64*st + 64) -> canonical core matrices (16-byte chunks)
                    const int k0 = st * kTileK;
                    for (int i = lane; i < npad * 8; i += 32) {
                        const int n = i >> 3, kg = i & 7;
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (n < bs) v = __ldg(reinterpret_cast<const uint4 *>(x + (size_t)n * K + k0 + kg * 8));
                        *reinterpret_cast<uint4 *>(b_stage + ((n >> 3) * 8 + kg) * 128 + (n & 7) * 16) = v;
                    }
                }
                fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncthreads();
                if (warp == kTcDecodeWarps && lane == 0) {
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(a_stage), b_addr = smem_u32(b_stage);
#pragma unroll
                    for (int kk = 0; kk < kTileK / 16; ++kk)
                        umma_f16(tmem_d, make_smem_desc(a_addr + kk * 256), make_smem_desc(b_addr + kk * 256), idesc,
                                 (it > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&bar[d]);  // arrives when the MMAs above (and all earlier ones) have completed
                }
            }
        }
    }
    // wait for the last commit of each stage that was used, then read the accumulators
    const int nsteps = st1 - st0;
#pragma unroll
    for (int d = 0; d < kTcStages; ++d) {
        const int uses = (nsteps - d + kTcStages - 1) / kTcStages;  // commits on stage d
        if (uses > 0) mbar_wait(&bar[d], (uint32_t)((uses - 1) & 1));
    }
    tc_fence_after();
    if (warp < 4 && nsteps > 0) {
        const int row = row0 + mb * kTileM + warp * 32 + lane;
        for (int c0 = 0; c0 < npad; c0 += 8) {
            uint32_t r[8];
            tmem_ld_x8(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + i < bs) atomicAdd(out + (size_t)(c0 + i) * M + row, __uint_as_float(r[i]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)(npad < 32 ? 32 : (npad <= 64 ? 64 : 128)));
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
    static constexpr int kSL = (E <= 10) ? 7 : (17 - E);
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
    __device__ static __forceinline__ void tile(const uint32_t (&P)[TcqGeom<E>::kWords], const uint8_t *tab, uint32_t (&f)[4]) {
        constexpr int SL = GLutTable<E>::kSL;
        f[0] = *reinterpret_cast<const uint32_t *>(tab + lut_pair_offset<E, TI, 0, SL>(P));
        f[1] = *reinterpret_cast<const uint32_t *>(tab + lut_pair_offset<E, TI, 1, SL>(P));
        f[2] = *reinterpret_cast<const uint32_t *>(tab + lut_pair_offset<E, TI, 2, SL>(P));
        f[3] = *reinterpret_cast<const uint32_t *>(tab + lut_pair_offset<E, TI, 3, SL>(P));
    }
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<E>::kWords], int, const uint8_t *tab_lane,
                                                  uint32_t (&frag)[4][4]) {
        tile<0>(P, tab_lane, frag[0]);
        tile<1>(P, tab_lane, frag[1]);
        tile<2>(P, tab_lane, frag[2]);
        tile<3>(P, tab_lane, frag[3]);
    }
};

// ---- host ----------------------------------------------------------------------------------------------------------------
template <class DecA, class DecB, class Table>
static int launch_tc(TcPart a, TcPart b, float *out, const void *x, const void *lut, int lut_arg, int M, int K, int bs,
                     int rows, int row0, cudaStream_t st) {
    auto kern = gemm_tc_kernel<DecA, DecB, Table>;
    const int npad = (bs + 15) & ~15;
    const size_t smem = (size_t)Table::kSmemBytes + (size_t)kTcStages * (kABytes + (size_t)npad * kTileK * 2);
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 2048, "bs = %d does not fit the shared-memory stages", bs);
    static bool configured = false;
    if (!configured) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048));
        configured = true;
    }
    const int mblocks = rows / kTileM;
    const int total_steps = a.steps + b.steps;
    int ksplit = (2 * sm_count() + mblocks - 1) / mblocks;  // ~2 CTAs' worth of work items per SM for balance
    if (ksplit > total_steps) ksplit = total_steps;
    if (ksplit < 1) ksplit = 1;
    QP_CUDA(launch_pdl(kern, dim3(mblocks * ksplit), dim3(kTcThreads), smem, st, a, b, out, (const __half *)x, lut, lut_arg,
                       M, K, bs, npad, ksplit, row0));
    return check_launch("gemm_tc");
}

template <int KVA, int KVB, int S>
static int launch_tc_tcq(TcPart a, TcPart b, float *out, const void *x, const void *tlut, int M, int K, int bs, int rows,
                         int row0, cudaStream_t st) {
    using DA = GTcqDecoder<KVA, S>;
    using DB = GTcqDecoder<(KVB ? KVB : KVA), S>;
    return launch_tc<DA, DB, TcTcqTable<S>>(a, b, out, x, tlut, 0, M, K, bs, rows, row0, st);
}

#define QP_TC_S(FN, KA, KB, ...)                              \
    switch (S) {                                              \
        case 9: return FN<KA, KB, 9>(__VA_ARGS__);            \
        case 10: return FN<KA, KB, 10>(__VA_ARGS__);          \
        case 11: return FN<KA, KB, 11>(__VA_ARGS__);          \
    }                                                         \
    break;

static int dispatch_tc_tcq(int S, int kva, int kvb, TcPart a, TcPart b, float *out, const void *x, const void *tlut, int M,
                           int K, int bs, int rows, int row0, cudaStream_t st) {
    if (kvb == 0) {
        switch (kva) {
            case 2: QP_TC_S(launch_tc_tcq, 2, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 3: QP_TC_S(launch_tc_tcq, 3, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 4: QP_TC_S(launch_tc_tcq, 4, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 5: QP_TC_S(launch_tc_tcq, 5, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 6: QP_TC_S(launch_tc_tcq, 6, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 7: QP_TC_S(launch_tc_tcq, 7, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 8: QP_TC_S(launch_tc_tcq, 8, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 9: QP_TC_S(launch_tc_tcq, 9, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 10: QP_TC_S(launch_tc_tcq, 10, 0, a, b, out, x, tlut, M, K, bs, rows, row0, st)
        }
    } else if (kvb == kva + 1) {
        switch (kva) {
            case 2: QP_TC_S(launch_tc_tcq, 2, 3, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 3: QP_TC_S(launch_tc_tcq, 3, 4, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 4: QP_TC_S(launch_tc_tcq, 4, 5, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 5: QP_TC_S(launch_tc_tcq, 5, 6, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 6: QP_TC_S(launch_tc_tcq, 6, 7, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 7: QP_TC_S(launch_tc_tcq, 7, 8, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 8: QP_TC_S(launch_tc_tcq, 8, 9, a, b, out, x, tlut, M, K, bs, rows, row0, st)
            case 9: QP_TC_S(launch_tc_tcq, 9, 10, a, b, out, x, tlut, M, K, bs, rows, row0, st)
        }
    }
    return fail(QP_ERR_ARG, "unsupported TCQ configuration S=%d KV=(%d,%d) for the tensor-core GEMM", S, kva, kvb);
}

template <int E>
static int launch_tc_lut(TcPart a, float *out, const void *x, const void *lut, int r_single, int M, int K, int bs,
                         cudaStream_t st) {
    TcPart none{nullptr, 0, 0};
    return launch_tc<GLutDecoder<E>, GLutDecoder<E>, GLutTable<E>>(a, none, out, x, lut, r_single, M, K, bs, M, 0, st);
}

}  // namespace qp

using namespace qp;

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
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    if (split_mode == QP_SPLIT_NONE) {
        TcPart a{(const uint32_t *)codes1, K / 32, K / 64}, none{nullptr, 0, 0};
        return dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, M, 0, st);
    }
    QP_CHECK_ARG(codes2 != nullptr, "codes2 is NULL for a two-rate layer");
    if ((rc = check_align(codes2, 16, "codes2")) != QP_OK) return rc;
    if (split_mode == QP_SPLIT_IN) {
        QP_CHECK_ARG(part1 > 0 && part1 < K && part1 % 64 == 0 && (K - part1) % 64 == 0, "in_part boundary must be a multiple of 64");
        TcPart a{(const uint32_t *)codes1, part1 / 32, part1 / 64}, b{(const uint32_t *)codes2, (K - part1) / 32, (K - part1) / 64};
        if (KV2 == KV1 + 1) return dispatch_tc_tcq(S, KV1, KV2, a, b, out, x_f16, tlut_f16, M, K, bs, M, 0, st);
        // arbitrary rate pair: two accumulating launches over the column halves
        TcPart none{nullptr, 0, 0};
        rc = dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, M, 0, st);
        if (rc != QP_OK) return rc;
        return fail(QP_ERR_ARG, "combt with KV2 != KV1 + 1 is not supported by the tensor-core GEMM");
    }
    if (split_mode == QP_SPLIT_OUT) {
        QP_CHECK_ARG(part1 > 0 && part1 < M && part1 % 128 == 0 && (M - part1) % 128 == 0, "out_part boundary must be a multiple of 128");
        TcPart a{(const uint32_t *)codes1, K / 32, K / 64}, b{(const uint32_t *)codes2, K / 32, K / 64}, none{nullptr, 0, 0};
        rc = dispatch_tc_tcq(S, KV1, 0, a, none, out, x_f16, tlut_f16, M, K, bs, part1, 0, st);
        if (rc != QP_OK) return rc;
        return dispatch_tc_tcq(S, KV2, 0, b, none, out, x_f16, tlut_f16, M, K, bs, M - part1, part1, st);
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
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    TcPart a{(const uint32_t *)codes, K / 32, K / 64};
    const int E = vec_sz == 2 ? bits : 2 * bits;
    const int r_single = vec_sz == 1 ? bits : 0;
    switch (E) {
#define QP_C(e) case e: return launch_tc_lut<e>(a, out, x_f16, lut_f16, r_single, M, K, bs, st);
        QP_C(2) QP_C(3) QP_C(4) QP_C(5) QP_C(6) QP_C(7) QP_C(8) QP_C(9) QP_C(10) QP_C(11) QP_C(12)
#undef QP_C
    }
    return fail(QP_ERR_ARG, "unsupported configuration");
}
