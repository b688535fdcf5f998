// lut_kernels.cu -- fused LUT-decode + GEMV (bs <= 8) and dequantise kernels for Q-Palette's VQ (vec_sz 2) and SQ
// (vec_sz 1) quantizers in the tensor-core packed layout (`qweight`, lib/quantizer/quant_op.py:101-162).
// Replaces kernels/vq-tensor-kernels/src/inference.cu:570-1108 (vq2 / sq / sq_dup).  Same streaming skeleton as the TCQ
// kernels (gemv_common.cuh); the decode is a plain table lookup:
//   pair mode  : one lookup of the E-bit pair code in a 2^E-entry half2 table (vq2: the lut itself; SQ with R <= 5: the
//                derived table {lut[c0], lut[c1]}, the reference's "sq_dup" idea extended to R = 5)
//   split mode : SQ with R = 6..8: two lookups in the 2^R-entry table + one byte-permute.
// Tables are lane-replicated in shared memory (slot = 2^SL bytes) so gathers are bank-conflict free up to E = 10; for
// E = 11 / 12 only 16 / 8 copies fit in 128 KiB.
#include "gemv_common.cuh"
#include "lut_bits.cuh"
#include "xprod.cuh"

namespace qp {

template <int E, bool SPLIT>
struct LutTable {
    static constexpr int kR = E / 2;                                           // single-code bits (SQ)
    static constexpr int kIndexBits = SPLIT ? kR : E;
    static constexpr int kSL = (kIndexBits <= 10) ? 7 : (17 - kIndexBits);       // log2(slot bytes)
    static constexpr int kEntries = 1 << kIndexBits;
    static constexpr int kBytes = kEntries << kSL;
    static constexpr uint32_t kLaneMask = (1u << (kSL - 2)) - 1u;
};

// compact lut copy size in 32-bit words.  r_single = 0: lut is (2^E, 2) fp16 (vq2).  r_single = R: lut is (2^R, 1) fp16.
__host__ __device__ inline int lut_copy_words(int E, int r_single) { return r_single ? ((1 << r_single) / 2) : (1 << E); }
// ... rounded up so the x stage behind it stays 16-byte aligned
__host__ __device__ inline int lut_compact_words(int E, int r_single) { return (lut_copy_words(E, r_single) + 3) & ~3; }

// expand the compact lut copy (shared memory) into the lane-replicated table
template <int E, bool SPLIT>
__device__ __forceinline__ void lut_build_table(uint32_t *tab, const uint32_t *lc, int r_single) {
    using T = LutTable<E, SPLIT>;
    const uint16_t *lut16 = reinterpret_cast<const uint16_t *>(lc);
    auto value = [&](int e) -> uint32_t {
        if (SPLIT) return lut16[e];
        if (r_single == 0) return lc[e];
        const uint32_t lo = lut16[e & ((1 << r_single) - 1)], hi = lut16[e >> r_single];
        return lo | (hi << 16);
    };
    if constexpr (T::kSL == 7) {
        fill_replicated_128(tab, T::kEntries, value);
    } else {
        constexpr int copies = 1 << (T::kSL - 2);
        for (int i = threadIdx.x; i < T::kEntries * copies; i += blockDim.x) tab[i] = value(i / copies);
    }
}

template <int E, bool SPLIT>
struct LutDecoder {
    static constexpr int kE = E;
    using T = LutTable<E, SPLIT>;

    template <int TI, int J>
    __device__ static __forceinline__ uint32_t one(const uint32_t (&P)[TcqGeom<E>::kWords], uint32_t tab) {
        if constexpr (!SPLIT) {
            return *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (lut_pair_offset<E, TI, J, T::kSL>(P) | tab));
        } else {
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (lut_single_offset<E, TI, J, 0, T::kSL>(P) | tab));
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (lut_single_offset<E, TI, J, 1, T::kSL>(P) | tab));
            return __byte_perm(w0, w1, 0x5410);
        }
    }
    template <int TI>
    __device__ static __forceinline__ void tile(const uint32_t (&P)[TcqGeom<E>::kWords], uint32_t tab, uint32_t (&f)[4]) {
        f[0] = one<TI, 0>(P, tab);
        f[1] = one<TI, 1>(P, tab);
        f[2] = one<TI, 2>(P, tab);
        f[3] = one<TI, 3>(P, tab);
    }
    __device__ static __forceinline__ void decode(const uint32_t (&P)[TcqGeom<E>::kWords], int lane,
                                                  uint32_t tab_addr_lane, uint32_t (&frag)[4][4]) {
        (void)lane;
        tile<0>(P, tab_addr_lane, frag[0]);
        tile<1>(P, tab_addr_lane, frag[1]);
        tile<2>(P, tab_addr_lane, frag[2]);
        tile<3>(P, tab_addr_lane, frag[3]);
    }
};

// LL: the fused prologue polls one operand out of the row-sharded LL receive buffer (xprod.cuh); separate instantiation
template <int E, bool SPLIT, bool LL>
__global__ void __launch_bounds__(kGemvThreads, 1)
lut_gemv_kernel(PackSegment seg, RunSplit split, float *__restrict__ out, const uint32_t *__restrict__ x32,
                const void *__restrict__ lut, int r_single, int M, int K, int bs, XProd prod) {
    using T = LutTable<E, SPLIT>;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ float red[32];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    uint32_t *lc = reinterpret_cast<uint32_t *>(smem + T::kBytes);  // compact lut copy
    uint32_t *xs = lc + lut_compact_words(E, r_single);
    const int lane = threadIdx.x & 31;
    const WarpRun2 run = warp_run2(seg, split, blockIdx.x * kGemvWarps + warp_in_cta());
    uint32_t raw[kGemv2Depth][TcqGeom<E>::kRawWords];
    gemv2_prefetch<E>(seg, run, raw);
    coop_copy_words(lc, reinterpret_cast<const uint32_t *>(lut), lut_copy_words(E, r_single));
    __syncthreads();
    lut_build_table<E, SPLIT>(tab, lc, r_single);
    if (prod.mode != 0) xp_zero(prod);  // before the wait: see xprod.cuh
    pdl_wait();
    if (prod.mode == 0) stage_x(xs, x32, K, bs);
    else produce_x_dispatch<LL>(xs, reinterpret_cast<float *>(xs + (size_t)K * bs / 2), red, prod, K);
    __syncthreads();
    pdl_launch_dependents();
    const uint32_t tab_addr_lane = (lane & T::kLaneMask) << 2;  // the table starts the dynamic shared memory
    gemv2_run<LutDecoder<E, SPLIT>>(seg, out, M, bs, smem_u32(xs), tab_addr_lane, run, raw, [] {});
}

// batched form for 9 <= bs <= 32 (see gemv_common.cuh: K-slabs of x fragments in shared memory, NB blocks of 8 batch rows);
// replaces `decompress_* + x @ dq.T` of lib/linear/vq_linear.py:58-66 at these batch sizes
template <int E, bool SPLIT, int NB>
__global__ void __launch_bounds__(kMmaThreads<NB>, 1)
lut_gemm_mma_kernel(PackSegment seg, RunSplit split, float *__restrict__ out, const uint4 *xfrag, const void *__restrict__ lut,
                    int r_single, int M, int bs) {
    using T = LutTable<E, SPLIT>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    uint32_t *lc = reinterpret_cast<uint32_t *>(smem + T::kBytes);  // compact lut copy
    uint4 *xs = reinterpret_cast<uint4 *>(lc + lut_compact_words(E, r_single));
    constexpr int W = kMmaSlabBytes / (NB * 512), kWarps = kMmaThreads<NB> / 32;
    const int lane = threadIdx.x & 31;
    unsigned lo, hi;
    split_range(split, (int)blockIdx.x, lo, hi);
    if (lo >= hi) return;
    const PackSegment none{nullptr, 0, 0, 0, 0};
    uint32_t raw[kMmaDepth<NB>][TcqGeom<E>::kRawWords];
    MmaPiece pc = mma_piece(seg, none, W, lo, hi);
    MmaRun run = mma_begin<E, NB>(seg, pc, kWarps, raw);
    coop_copy_words(lc, reinterpret_cast<const uint32_t *>(lut), lut_copy_words(E, r_single));
    __syncthreads();
    lut_build_table<E, SPLIT>(tab, lc, r_single);
    pdl_wait();
    const uint32_t tab_addr_lane = (lane & T::kLaneMask) << 2;
    const uint32_t xs_addr = smem_u32(xs);
    while (true) {
        const uint4 *src = xfrag + (size_t)(seg.ksuper0 + pc.col0) * (NB * 32);
        for (int i = threadIdx.x; i < pc.w * NB * 32; i += kMmaThreads<NB>) xs[i] = __ldcg(src + i);
        __syncthreads();
        mma_stream<LutDecoder<E, SPLIT>, NB>(seg, pc, run, out, M, bs, xs_addr, tab_addr_lane, raw);
        if (pc.next >= hi) break;
        pc = mma_piece(seg, none, W, pc.next, hi);
        run = mma_begin<E, NB>(seg, pc, kWarps, raw);
        __syncthreads();
    }
    pdl_launch_dependents();
}

template <int E, bool SPLIT>
__global__ void __launch_bounds__(kGemvThreads, 1)
lut_dequant_kernel(PackSegment seg, RunSplit split, __half *__restrict__ W, const void *__restrict__ lut, int r_single,
                   int K) {
    using T = LutTable<E, SPLIT>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kGemvWarps + warp_in_cta();
    uint32_t *lc = reinterpret_cast<uint32_t *>(smem + T::kBytes);
    coop_copy_words(lc, reinterpret_cast<const uint32_t *>(lut), lut_copy_words(E, r_single));
    __syncthreads();
    lut_build_table<E, SPLIT>(tab, lc, r_single);
    __syncthreads();
    const uint32_t tab_addr_lane = (lane & T::kLaneMask) << 2;  // the table starts the dynamic shared memory
    dequant_run_segment<LutDecoder<E, SPLIT>>(seg, W, K, tab_addr_lane, split, gwarp);
}

template <int E, bool SPLIT>
static int launch_lut_gemv(PackSegment seg, float *out, const void *x, const void *lut, int r_single, int M, int K,
                           int bs, const XProd &prod, cudaStream_t st) {
    using T = LutTable<E, SPLIT>;
    const bool ll = prod.mode == 2;
    auto kern = ll ? lut_gemv_kernel<E, SPLIT, true> : lut_gemv_kernel<E, SPLIT, false>;
    static DeviceOnce configured[2];
    if (configured[ll].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 256));
    }
    const size_t smem = (size_t)T::kBytes + 4 * (size_t)lut_compact_words(E, r_single) + (size_t)K * bs * 2 +
                        (prod.mode ? (size_t)K * 4 : 0);
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 256, "K = %d does not fit the fused-prologue shared-memory budget", K);
    QP_CUDA(launch_pdl(kern, dim3(sm_count()), dim3(kGemvThreads), smem, st, seg_with_magic(seg),
                       make_split((long)seg.strips * seg.ksuper, sm_count() * kGemvWarps), out, (const uint32_t *)x, lut,
                       r_single, M, K, bs, prod));
    return check_launch("lut_gemv");
}

template <int E, bool SPLIT>
static int launch_lut_gemm_mma(PackSegment seg, float *out, const uint4 *xfrag, const void *lut, int r_single, int M, int bs,
                               cudaStream_t st) {
    using T = LutTable<E, SPLIT>;
    const int v = mma_batch_blocks(bs) == 4;
    auto kern = v == 0 ? lut_gemm_mma_kernel<E, SPLIT, 2> : lut_gemm_mma_kernel<E, SPLIT, 4>;
    static DeviceOnce configured[2];
    if (configured[v].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 256));
    }
    const int nctas = sm_count();
    const size_t smem = (size_t)T::kBytes + 4 * (size_t)lut_compact_words(E, r_single) + kMmaSlabBytes;
    QP_CUDA(launch_pdl(kern, dim3(nctas), dim3(v == 0 ? kMmaThreads<2> : kMmaThreads<4>), smem, st, seg,
                       make_split((long)seg.strips * seg.ksuper, nctas), out, xfrag, lut, r_single, M, bs));
    return check_launch("lut_gemm_mma");
}

template <int E, bool SPLIT>
static int launch_lut_dequant(PackSegment seg, __half *W, const void *lut, int r_single, int K, cudaStream_t st) {
    using T = LutTable<E, SPLIT>;
    auto kern = lut_dequant_kernel<E, SPLIT>;
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    }
    kern<<<sm_count(), kGemvThreads, T::kBytes + 4 * lut_compact_words(E, r_single), st>>>(
        seg, make_split((long)seg.strips * seg.ksuper, sm_count() * kGemvWarps), W, lut, r_single, K);
    return check_launch("lut_dequant");
}

// smem bytes of the table for (bits, vec_sz); <0 if unsupported
static int lut_table_bytes(int bits, int vec_sz) {
    if (vec_sz == 2) {
        if (bits < 2 || bits > 12) return -1;
        return (1 << bits) << (bits <= 10 ? 7 : 17 - bits);
    }
    if (vec_sz == 1) {
        if (bits < 2 || bits > 8) return -1;
        return bits <= 5 ? ((1 << (2 * bits)) << 7) : ((1 << bits) << 7);
    }
    return -1;
}

#define QP_LUT_DISPATCH(FN, ...)                                                     \
    if (vec_sz == 2) {                                                               \
        switch (bits) {                                                              \
            case 2: return FN<2, false>(__VA_ARGS__);                                \
            case 3: return FN<3, false>(__VA_ARGS__);                                \
            case 4: return FN<4, false>(__VA_ARGS__);                                \
            case 5: return FN<5, false>(__VA_ARGS__);                                \
            case 6: return FN<6, false>(__VA_ARGS__);                                \
            case 7: return FN<7, false>(__VA_ARGS__);                                \
            case 8: return FN<8, false>(__VA_ARGS__);                                \
            case 9: return FN<9, false>(__VA_ARGS__);                                \
            case 10: return FN<10, false>(__VA_ARGS__);                              \
            case 11: return FN<11, false>(__VA_ARGS__);                              \
            case 12: return FN<12, false>(__VA_ARGS__);                              \
        }                                                                            \
    } else {                                                                         \
        switch (bits) {                                                              \
            case 2: return FN<4, false>(__VA_ARGS__);                                \
            case 3: return FN<6, false>(__VA_ARGS__);                                \
            case 4: return FN<8, false>(__VA_ARGS__);                                \
            case 5: return FN<10, false>(__VA_ARGS__);                               \
            case 6: return FN<12, true>(__VA_ARGS__);                                \
            case 7: return FN<14, true>(__VA_ARGS__);                                \
            case 8: return FN<16, true>(__VA_ARGS__);                                \
        }                                                                            \
    }

static int dispatch_lut_gemv(int bits, int vec_sz, PackSegment seg, float *out, const void *x, const void *lut, int M,
                             int K, int bs, const XProd &prod, cudaStream_t st) {
    const int r_single = vec_sz == 1 ? bits : 0;
    QP_LUT_DISPATCH(launch_lut_gemv, seg, out, x, lut, r_single, M, K, bs, prod, st)
    return fail(QP_ERR_ARG, "unsupported LUT configuration bits=%d vec_sz=%d", bits, vec_sz);
}

static int dispatch_lut_gemm_mma(int bits, int vec_sz, PackSegment seg, float *out, const uint4 *xfrag, const void *lut, int M,
                                 int bs, cudaStream_t st) {
    const int r_single = vec_sz == 1 ? bits : 0;
    QP_LUT_DISPATCH(launch_lut_gemm_mma, seg, out, xfrag, lut, r_single, M, bs, st)
    return fail(QP_ERR_ARG, "unsupported LUT configuration bits=%d vec_sz=%d", bits, vec_sz);
}

static int dispatch_lut_dequant(int bits, int vec_sz, PackSegment seg, __half *W, const void *lut, int K,
                                cudaStream_t st) {
    const int r_single = vec_sz == 1 ? bits : 0;
    QP_LUT_DISPATCH(launch_lut_dequant, seg, W, lut, r_single, K, st)
    return fail(QP_ERR_ARG, "unsupported LUT configuration bits=%d vec_sz=%d", bits, vec_sz);
}

}  // namespace qp

using namespace qp;

static int lut_check(const void *codes, int M, int K, int bits, int vec_sz) {
    QP_CHECK_ARG(codes != nullptr, "codes is NULL");
    QP_CHECK_ARG(vec_sz == 1 || vec_sz == 2, "tensor-core layout supports vec_sz 1 or 2 (got %d)", vec_sz);
    QP_CHECK_ARG(lut_table_bytes(bits, vec_sz) > 0, "unsupported bits=%d for vec_sz=%d", bits, vec_sz);
    QP_CHECK_ARG(M > 0 && K > 0 && M % 32 == 0 && K % 32 == 0, "needs M %% 32 == 0 and K %% 32 == 0 (got %d x %d)", M, K);
    return check_align(codes, 16, "codes");
}

extern "C" int qp_lut_gemv(float *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs,
                           int bits, int vec_sz, unsigned flags, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && x_f16 && lut_f16, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 8, "bs = %d: the fused GEMV handles 1..8 rows", bs);
    int rc = lut_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(x_f16, 16, "x")) != QP_OK) return rc;
    const size_t avail = (size_t)kMaxSmem - 256 - (size_t)lut_table_bytes(bits, vec_sz) - 16 * 1024;
    int chunk = (int)(avail / ((size_t)K * 2));
    QP_CHECK_ARG(chunk >= 1, "K = %d too large for the shared-memory x stage", K);
    if (chunk > bs) chunk = bs;
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    PackSegment seg{(const uint32_t *)codes, M / 32, K / 32, 0, 0};
    for (int b0 = 0; b0 < bs; b0 += chunk) {
        const int nb = (bs - b0 < chunk) ? bs - b0 : chunk;
        XProd none = {};
        rc = dispatch_lut_gemv(bits, vec_sz, seg, out + (size_t)b0 * M, (const __half *)x_f16 + (size_t)b0 * K, lut_f16,
                               M, K, nb, none, st);
        if (rc != QP_OK) return rc;
    }
    return QP_OK;
}

extern "C" int qp_lut_gemm_mma(float *out, const void *codes, const void *x_f16, const void *lut_f16, void *scratch, int M, int K,
                               int bs, int bits, int vec_sz, unsigned flags, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(out && x_f16 && lut_f16 && scratch, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 128, "bs = %d out of range 1..128", bs);
    int rc = lut_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(x_f16, 4, "x")) != QP_OK) return rc;
    if ((rc = check_align(scratch, 16, "scratch")) != QP_OK) return rc;
    if (!(flags & QP_FLAG_ACCUMULATE)) QP_CUDA(cudaMemsetAsync(out, 0, (size_t)bs * M * sizeof(float), st));
    PackSegment seg{(const uint32_t *)codes, M / 32, K / 32, 0, 0};
    return mma_gemm_batches(out, x_f16, scratch, M, K, bs, st, [&](float *o, const uint4 *xfrag, int nb) {
        return dispatch_lut_gemm_mma(bits, vec_sz, seg, o, xfrag, lut_f16, M, nb, st);
    });
}

extern "C" int qp_lut_dequant(void *W_f16, const void *codes, const void *lut_f16, int M, int K, int bits, int vec_sz,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    QP_CHECK_ARG(W_f16 && lut_f16, "NULL pointer argument");
    int rc = lut_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    PackSegment seg{(const uint32_t *)codes, M / 32, K / 32, 0, 0};
    return dispatch_lut_dequant(bits, vec_sz, seg, (__half *)W_f16, lut_f16, K, st);
}

// ---- fused prologue entry point -------------------------------------------------------------------------------------
int qp_make_xprod(qp::XProd &p, const qp_xprod *u, int K) {
    QP_CHECK_ARG(u->src_f16 != nullptr, "xprod.src_f16 is NULL");
    QP_CHECK_ARG(!u->acc || u->wscale_f16, "xprod.acc given without wscale");
    QP_CHECK_ARG(!u->ll || (u->ll_epoch && (u->ll_kind == 1 || u->ll_kind == 2)), "xprod.ll needs ll_epoch and ll_kind 1 or 2");
    QP_CHECK_ARG(!u->ll || u->ll_kind != 1 || (u->wscale_f16 && !u->acc), "xprod.ll_kind 1 replaces acc and needs wscale");
    QP_CHECK_ARG(K % 128 == 0, "fused prologue needs K %% 128 == 0 (K = %d)", K);
    int Kf = 1, m = K;
    if ((K & (K - 1)) != 0) {
        QP_CHECK_ARG(K % 28 == 0 && (((K / 28) & (K / 28 - 1)) == 0), "Hadamard size %d is neither 2^k nor 28*2^k", K);
        Kf = 28;
        m = K / 28;
    }
    QP_CHECK_ARG(m >= 128, "Hadamard block %d < 128", m);
    QP_CHECK_ARG(K <= 5 * 4 * kGemvThreads, "K = %d too large for the fused prologue", K);
    p.mode = u->ll ? 2 : 1;
    p.src = (const __half *)u->src_f16;
    p.h_out = (__half *)u->h_out_f16;
    p.acc = u->acc;
    p.wscale = (const __half *)u->wscale_f16;
    p.acc_scale = u->acc_scale;
    p.norm_w = (const __half *)u->norm_w_f16;
    p.eps = u->eps;
    p.su = (const __half *)u->su_f16;
    p.had_scale = u->had_scale;
    p.x_out = (__half *)u->x_out_f16;
    p.zero1 = u->zero1;
    p.zero1_count = u->zero1_count;
    p.zero2 = u->zero2;
    p.zero2_count = u->zero2_count;
    p.m = m;
    p.Kf = Kf;
    p.ll = (const uint4 *)u->ll;
    p.ll_epoch = u->ll_epoch;
    p.ll_kind = u->ll ? u->ll_kind : 0;
    p.ll_spin_cycles = qp_spin_limit_cycles_host();
    return QP_OK;
}

extern "C" int qp_lut_gemv_fused(float *out, const void *codes, const qp_xprod *xp, const void *lut_f16, int M, int K,
                                 int bits, int vec_sz, void *stream) {
    QP_CHECK_ARG(out && xp && lut_f16, "NULL pointer argument");
    int rc = lut_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    XProd p;
    if ((rc = qp_make_xprod(p, xp, K)) != QP_OK) return rc;
    PackSegment seg{(const uint32_t *)codes, M / 32, K / 32, 0, 0};
    return dispatch_lut_gemv(bits, vec_sz, seg, out, nullptr, lut_f16, M, K, 1, p, (cudaStream_t)stream);
}
