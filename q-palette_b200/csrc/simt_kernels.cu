// simt_kernels.cu -- LUT GEMV / dequantise for Q-Palette's SIMT packed layouts (lib/quantizer/pack_op.py:288-335 for
// vec_sz 1, lib/quantizer/quant_op.py:33-78 for vec_sz 2 and 4) and the tensor-core -> SIMT format conversion
// (lib/quantizer/quant_op.py:246-257).  Replaces kernels/sq-cuda-kernels/gemm_routines.cu:474-722 and
// kernels/vq-cuda-kernels/src/gemm_routines.cu:1913-2120.
//
// Layout: row-major per output row; K in chunks of 32 threads x 32*vec weights; thread t owns 4*vec groups of 8
// consecutive weights at chunk + (t + eff*g)*8 and its 32 codes are packed LSB-first into `bits` u32 words; word j is
// stored at chunk_word_base + t + eff*j (eff = 32, or (K mod chunk)/(32*vec) for a ragged last chunk).
// One warp per output row (grid-stride), codes loaded coalesced, LUT lane-replicated in shared memory, x staged in
// shared memory; products are summed 4 at a time in fp16x2 and accumulated in fp32 (the reference accumulates in fp16).
#include "gemv_common.cuh"

namespace qp {

constexpr int kSimtThreads = 1024;
constexpr int kSimtWarps = kSimtThreads / 32;

// Scalar codes of <= 5 bits are looked up two at a time (adjacent codes are one 2*BITS-bit field of the stream; the table
// holds the fp16 pair), like the tensor-core layout's pair tables: half the lookups and no byte permutes.
template <int BITS, int VEC>
struct SimtTable {
    static constexpr bool kPair = (VEC == 1 && BITS <= 5);
    static constexpr int kFieldBits = kPair ? 2 * BITS : BITS;
    static constexpr int kSL = (kFieldBits <= 10) ? 7 : (17 - kFieldBits);
    static constexpr int kBytes = (1 << kFieldBits) << kSL;
    // VEC = 4: an entry is four fp16 = 8 bytes, read with one LDS.64; a 64-bit shared load is served per half-warp, so 16
    // copies (one 128-byte slot) are already conflict-free
    static constexpr int kEntryShift = (VEC == 4) ? 3 : 2;
    static constexpr uint32_t kLaneMask = (1u << (kSL - kEntryShift)) - 1u;
};

template <int BITS, int VEC>
__device__ __forceinline__ void simt_build_table(uint32_t *tab, const uint32_t *lc) {
    using T = SimtTable<BITS, VEC>;
    constexpr int copies = 1 << (T::kSL - T::kEntryShift);
    const uint16_t *l16 = reinterpret_cast<const uint16_t *>(lc);
    for (int i = threadIdx.x; i < (1 << T::kFieldBits) * copies; i += blockDim.x) {
        const int e = i / copies;
        if (VEC == 4) reinterpret_cast<uint2 *>(tab)[i] = reinterpret_cast<const uint2 *>(lc)[e];
        else if (VEC == 2) tab[i] = lc[e];
        else if (T::kPair) tab[i] = (uint32_t)l16[e & ((1 << BITS) - 1)] | ((uint32_t)l16[e >> BITS] << 16);
        else tab[i] = (uint32_t)l16[e];
    }
}

// field I (FB bits wide, at bit I*FB) of the thread's BITS-word little-endian stream, shifted left by SL and masked
// (table slot offset)
template <int BITS, int FB, int I, int SL>
__device__ __forceinline__ uint32_t simt_code_offset(const uint32_t (&w)[BITS]) {
    constexpr int o = I * FB;
    constexpr int wi = o / 32, s = o % 32;
    constexpr uint32_t mask = ((1u << FB) - 1u) << SL;
    uint32_t v;
    if constexpr (s == 0) v = w[wi];
    else if constexpr (s + FB <= 32) v = w[wi] >> s;
    else v = __funnelshift_r(w[wi], w[wi + 1], s);
    return (v << SL) & mask;
}

// decode the 8 weights of group G (codes G*8/VEC ...) into 4 half2 registers
template <int BITS, int VEC, int G>
__device__ __forceinline__ void simt_group(const uint32_t (&w)[BITS], uint32_t tab, uint32_t (&h)[4]) {
    using T = SimtTable<BITS, VEC>;
    constexpr int SL = T::kSL, FB = T::kFieldBits;
    if constexpr (VEC == 4) {  // one 64-bit lookup per four weights
        const uint2 a = *reinterpret_cast<const uint2 *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 2 + 0, SL>(w) | tab));
        const uint2 b = *reinterpret_cast<const uint2 *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 2 + 1, SL>(w) | tab));
        h[0] = a.x, h[1] = a.y, h[2] = b.x, h[3] = b.y;
    } else if constexpr (VEC == 2 || T::kPair) {  // one lookup per fp16 pair
        h[0] = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 4 + 0, SL>(w) | tab));
        h[1] = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 4 + 1, SL>(w) | tab));
        h[2] = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 4 + 2, SL>(w) | tab));
        h[3] = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 4 + 3, SL>(w) | tab));
    } else {
        const uint32_t a0 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 0, SL>(w) | tab));
        const uint32_t a1 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 1, SL>(w) | tab));
        const uint32_t a2 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 2, SL>(w) | tab));
        const uint32_t a3 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 3, SL>(w) | tab));
        const uint32_t a4 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 4, SL>(w) | tab));
        const uint32_t a5 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 5, SL>(w) | tab));
        const uint32_t a6 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 6, SL>(w) | tab));
        const uint32_t a7 = *reinterpret_cast<const uint32_t *>(qp_dyn_smem + (simt_code_offset<BITS, FB, G * 8 + 7, SL>(w) | tab));
        h[0] = __byte_perm(a0, a1, 0x5410);
        h[1] = __byte_perm(a2, a3, 0x5410);
        h[2] = __byte_perm(a4, a5, 0x5410);
        h[3] = __byte_perm(a6, a7, 0x5410);
    }
}

__device__ __forceinline__ float dot8(const uint32_t (&h)[4], const uint4 xv) {
    __half2 s = __hmul2(*reinterpret_cast<const __half2 *>(&h[0]), *reinterpret_cast<const __half2 *>(&xv.x));
    s = __hfma2(*reinterpret_cast<const __half2 *>(&h[1]), *reinterpret_cast<const __half2 *>(&xv.y), s);
    s = __hfma2(*reinterpret_cast<const __half2 *>(&h[2]), *reinterpret_cast<const __half2 *>(&xv.z), s);
    s = __hfma2(*reinterpret_cast<const __half2 *>(&h[3]), *reinterpret_cast<const __half2 *>(&xv.w), s);
    const float2 f = __half22float2(s);
    return f.x + f.y;
}

template <int BITS, int VEC, int G, int NG>
__device__ __forceinline__ void simt_groups_gemv(const uint32_t (&w)[BITS], uint32_t tab, uint32_t xs_addr, int K, int bs,
                                                 int col0, int t, int eff, float (&acc)[8]) {
    if constexpr (G < NG) {
        uint32_t h[4];
        simt_group<BITS, VEC, G>(w, tab, h);
        const uint32_t xa = xs_addr + (uint32_t)(col0 + (t + G * eff) * 8) * 2u;
        acc[0] += dot8(h, lds_u128(xa));
        if (bs > 1) {  // one uniform branch per group in the bs = 1 decode case instead of eight predicated iterations
#pragma unroll
            for (int n = 1; n < 8; ++n)
                if (n < bs) acc[n] += dot8(h, lds_u128(xa + (uint32_t)n * (uint32_t)K * 2u));
        }
        simt_groups_gemv<BITS, VEC, G + 1, NG>(w, tab, xs_addr, K, bs, col0, t, eff, acc);
    }
}

template <int BITS, int VEC, int G, int NG>
__device__ __forceinline__ void simt_groups_store(const uint32_t (&w)[BITS], uint32_t tab, __half *__restrict__ Wrow,
                                                  int col0, int t, int eff) {
    if constexpr (G < NG) {
        uint32_t h[4];
        simt_group<BITS, VEC, G>(w, tab, h);
        const int col = col0 + (t + G * eff) * 8;
        *reinterpret_cast<uint4 *>(Wrow + col) = make_uint4(h[0], h[1], h[2], h[3]);
        simt_groups_store<BITS, VEC, G + 1, NG>(w, tab, Wrow, col0, t, eff);
    }
}

template <int BITS, int VEC, bool GEMV>
__global__ void __launch_bounds__(kSimtThreads, 1)
simt_kernel(__half *__restrict__ out, const uint32_t *__restrict__ codes, const __half *__restrict__ x,
            const void *__restrict__ lut, int M, int K, int bs, int out_f32) {
    using T = SimtTable<BITS, VEC>;
    constexpr int NG = 4 * VEC;
    constexpr int kChunk = 32 * 32 * VEC;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kCompactWords = (1 << BITS) * VEC / 2;
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    uint32_t *lc = reinterpret_cast<uint32_t *>(smem + T::kBytes);
    uint32_t *xs = lc + (kCompactWords < 4 ? 4 : kCompactWords);
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kSimtWarps + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kSimtWarps;

    coop_copy_words(lc, reinterpret_cast<const uint32_t *>(lut), kCompactWords);
    __syncthreads();
    simt_build_table<BITS, VEC>(tab, lc);
    // the dequantise form waits too: its output may alias a stream-ordered temporary the preceding kernel still reads
    pdl_wait();
    if (GEMV) coop_copy_words<true>(xs, reinterpret_cast<const uint32_t *>(x), bs * K / 2);
    __syncthreads();
    pdl_launch_dependents();
    const uint32_t tab_lane = (lane & T::kLaneMask) << T::kEntryShift;  // lane column; the table starts the dynamic shared memory (qp_dyn_smem)
    const uint32_t xs_addr = smem_u32(xs);
    const int row_words = BITS * K / 32 / VEC;
    const int nfull = K / kChunk, rem = K % kChunk;
    const int nchunks = nfull + (rem ? 1 : 0);

    // This warp's rows gwarp, gwarp + nwarps, ... are walked as one flat sequence of (row, chunk) units with the code words
    // of the NEXT unit already in flight (two register buffers): without it every chunk pays a full DRAM round trip
    // (measured 33-45 us for 4096 x 14336, 10-15 % of the HBM roofline).
    const int rows_w = gwarp < M ? (M - gwarp + nwarps - 1) / nwarps : 0;
    const int total = rows_w * nchunks;
    uint32_t w[2][BITS];
    auto load_unit = [&](uint32_t (&dst)[BITS], int row, int c) {
        const int eff = (c < nfull) ? 32 : rem / (32 * VEC);
        const uint32_t *q = codes + (size_t)row * row_words + (size_t)c * 32 * BITS + lane;
#pragma unroll
        for (int j = 0; j < BITS; ++j) dst[j] = (lane < eff) ? ldg_stream_u32(q + j * eff) : 0u;
    };
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    auto run_unit = [&](const uint32_t (&src)[BITS], int row, int c) {
        const int eff = (c < nfull) ? 32 : rem / (32 * VEC);
        if (lane < eff) {
            if constexpr (GEMV) simt_groups_gemv<BITS, VEC, 0, NG>(src, tab_lane, xs_addr, K, bs, c * kChunk, lane, eff, acc);
            else simt_groups_store<BITS, VEC, 0, NG>(src, tab_lane, out + (size_t)row * K, c * kChunk, lane, eff);
        }
        if (GEMV && c == nchunks - 1) {  // row finished: reduce over the warp, write, restart the accumulators
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (n < bs) {
                    float v = acc[n];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) {
                        if (out_f32) reinterpret_cast<float *>(out)[(size_t)n * M + row] = v;
                        else out[(size_t)n * M + row] = __float2half(v);
                    }
                }
                acc[n] = 0.f;
            }
        }
    };
    int row = gwarp, c = 0;        // unit being computed
    int nrow = gwarp, nc = 0;      // unit being loaded
    auto next = [&](int &r, int &cc) {
        if (++cc == nchunks) cc = 0, r += nwarps;
    };
    if (total > 0) load_unit(w[0], nrow, nc);
    for (int k = 0; k < total; k += 2) {
        next(nrow, nc);
        if (k + 1 < total) load_unit(w[1], nrow, nc);
        run_unit(w[0], row, c);
        next(row, c);
        if (k + 1 < total) {
            next(nrow, nc);
            if (k + 2 < total) load_unit(w[0], nrow, nc);
            run_unit(w[1], row, c);
            next(row, c);
        }
    }
}

template <int BITS, int VEC, bool GEMV>
static int launch_simt(__half *out, const void *codes, const void *x, const void *lut, int M, int K, int bs,
                       int out_f32, cudaStream_t st) {
    auto kern = simt_kernel<BITS, VEC, GEMV>;
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    }
    const size_t smem = (size_t)SimtTable<BITS, VEC>::kBytes + 4 * (size_t)(((1 << BITS) * VEC / 2) < 4 ? 4 : ((1 << BITS) * VEC / 2)) +
                        (GEMV ? (size_t)K * bs * 2 : 0);
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem, "bs*K = %d*%d does not fit the shared-memory x stage", bs, K);
    QP_CUDA(launch_pdl(kern, dim3(sm_count()), dim3(kSimtThreads), smem, st, out, (const uint32_t *)codes,
                       (const __half *)x, lut, M, K, bs, out_f32));
    return check_launch(GEMV ? "simt_gemv" : "simt_dequant");
}

template <bool GEMV>
static int dispatch_simt(int bits, int vec_sz, __half *out, const void *codes, const void *x, const void *lut, int M,
                         int K, int bs, int out_f32, cudaStream_t st) {
#define QP_C(B, V) \
    if (bits == B && vec_sz == V) return launch_simt<B, V, GEMV>(out, codes, x, lut, M, K, bs, out_f32, st);
    QP_C(2, 1) QP_C(3, 1) QP_C(4, 1) QP_C(5, 1) QP_C(6, 1) QP_C(7, 1) QP_C(8, 1)
    QP_C(2, 2) QP_C(3, 2) QP_C(4, 2) QP_C(5, 2) QP_C(6, 2) QP_C(7, 2) QP_C(8, 2) QP_C(9, 2) QP_C(10, 2) QP_C(11, 2)
    QP_C(12, 2)
    // ours_lib::vq_pack_{gemm,dequant}_simt_*_4_{6..12} (lib/linear/__init__.py:383-420)
    QP_C(6, 4) QP_C(7, 4) QP_C(8, 4) QP_C(9, 4) QP_C(10, 4) QP_C(11, 4) QP_C(12, 4)
#undef QP_C
    return fail(QP_ERR_ARG, "unsupported SIMT configuration bits=%d vec_sz=%d", bits, vec_sz);
}

// ---- tensor-core layout -> SIMT layout ---------------------------------------------------------------------------------
// one thread per output word.  code(row, pair) is read straight from the TC buffer (bit offset arithmetic of
// lib/quantizer/quant_op.py:101-162).
__device__ __forceinline__ uint32_t tc_read_code(const uint32_t *__restrict__ tc, int K, int bits, int vec, int row,
                                                 int col /* weight column, multiple of vec */) {
    const int E = bits * (2 / vec);
    const int mh = row >> 5, ml = (row >> 4) & 1, rr = row & 15;
    const int kh = col >> 5, kl = (col >> 4) & 1, cc = col & 15;
    const int lane = (rr & 7) * 4 + ((cc & 7) >> 1);
    const int j = (cc >> 3) * 2 + (rr >> 3);
    const int t = kl * 2 + ml;
    const size_t super = (size_t)mh * (K >> 5) + kh;
    const size_t bit = (super * 64 * E + (size_t)lane * 2 * E) * 8 + 4 * t * E + j * E + ((vec == 1) ? (cc & 1) * bits : 0);
    const size_t wi = bit >> 5;
    const int s = (int)(bit & 31);
    const uint32_t lo = tc[wi];
    const uint32_t hi = (s + bits > 32) ? tc[wi + 1] : 0u;
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> s) & ((1u << bits) - 1u);
}

__global__ void convert_tc_to_simt_kernel(uint32_t *__restrict__ simt, const uint32_t *__restrict__ tc, int M, int K,
                                          int bits, int vec) {
    const int row_words = bits * K / 32 / vec;
    const size_t total = (size_t)M * row_words;
    const int chunk = 32 * 32 * vec;
    const int nfull = K / chunk;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / row_words);
        const int wq = (int)(idx % row_words);
        const int c = wq / (32 * bits);
        const int eff = (c < nfull) ? 32 : (K % chunk) / (32 * vec);
        const int r = wq - c * 32 * bits;
        const int j = r / eff, t = r % eff;
        // bits [32j, 32j+32) of thread t's stream
        uint32_t word = 0;
        const int first = (32 * j) / bits, last = (32 * j + 31) / bits;
        const int cpg = 8 / vec;  // codes per group of 8 weights
        for (int i = first; i <= last; ++i) {
            const int g = i / cpg, within = i % cpg;
            const int col = c * chunk + (t + g * eff) * 8 + within * vec;
            const uint64_t code = tc_read_code(tc, K, bits, vec, row, col);
            const int pos = i * bits - 32 * j;
            if (pos >= 0) word |= (uint32_t)(code << pos);
            else word |= (uint32_t)(code >> (-pos));
        }
        simt[idx] = word;
    }
}

}  // namespace qp

using namespace qp;

static int simt_check(const void *codes, int M, int K, int bits, int vec_sz) {
    QP_CHECK_ARG(codes != nullptr, "codes is NULL");
    QP_CHECK_ARG(vec_sz == 1 || vec_sz == 2 || vec_sz == 4, "SIMT layout: vec_sz 1, 2 or 4 supported (got %d)", vec_sz);
    QP_CHECK_ARG(M > 0 && K > 0 && K % (32 * vec_sz) == 0, "SIMT layout needs K %% (32*vec_sz) == 0 (K=%d)", K);
    return check_align(codes, 4, "codes");
}

extern "C" int qp_simt_gemv(void *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs,
                            int bits, int vec_sz, int out_is_f32, void *stream) {
    void *out_f16 = out;
    QP_CHECK_ARG(out_f16 && x_f16 && lut_f16, "NULL pointer argument");
    QP_CHECK_ARG(bs >= 1 && bs <= 8, "bs = %d: the fused GEMV handles 1..8 rows", bs);
    int rc = simt_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(x_f16, 16, "x")) != QP_OK) return rc;
    return dispatch_simt<true>(bits, vec_sz, (__half *)out_f16, codes, x_f16, lut_f16, M, K, bs, out_is_f32,
                               (cudaStream_t)stream);
}

extern "C" int qp_simt_dequant(void *W_f16, const void *codes, const void *lut_f16, int M, int K, int bits, int vec_sz,
                               void *stream) {
    QP_CHECK_ARG(W_f16 && lut_f16, "NULL pointer argument");
    int rc = simt_check(codes, M, K, bits, vec_sz);
    if (rc != QP_OK) return rc;
    if ((rc = check_align(W_f16, 16, "W")) != QP_OK) return rc;
    return dispatch_simt<false>(bits, vec_sz, (__half *)W_f16, codes, nullptr, lut_f16, M, K, 1, 0, (cudaStream_t)stream);
}

extern "C" int qp_convert_tc_to_simt(void *simt_codes, const void *tc_codes, int M, int K, int bits, int vec_sz,
                                     void *stream) {
    QP_CHECK_ARG(simt_codes && tc_codes, "NULL pointer argument");
    QP_CHECK_ARG(vec_sz == 1 || vec_sz == 2, "vec_sz 1 or 2");
    QP_CHECK_ARG(M % 32 == 0 && K % 32 == 0 && K % (32 * vec_sz) == 0, "bad shape %d x %d", M, K);
    QP_CHECK_ARG(bits >= 2 && bits <= 12, "bits out of range");
    convert_tc_to_simt_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>((uint32_t *)simt_codes,
                                                                               (const uint32_t *)tc_codes, M, K, bits,
                                                                               vec_sz);
    return check_launch("convert_tc_to_simt");
}
