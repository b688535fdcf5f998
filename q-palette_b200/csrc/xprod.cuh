// xprod.cuh -- "x producer": the activation-side glue of an incoherent quantized linear, computed INSIDE the GEMV
// kernel's prologue by every CTA (redundantly: x is a few KB and L2 resident), so that
//     h' = h + fp16(acc)*Wscale*s        (residual add of the previous projection's output, optional)
//     y  = RMSNorm(h') * w                (optional)
//     x  = fp16( Hadamard(y * SU) * scale )
// costs no extra kernel launch and overlaps the first weight loads.  Replaces the standalone qp_fused_norm_had launch in
// front of a GEMV (reference: lib/linear/incoherent_linear.py:76-108,324-338 + LlamaDecoderLayer residual/RMSNorm).
// Only for bs = 1.  CTA 0 additionally stores h' (to a DIFFERENT buffer than h: other CTAs still read h) and, if asked, x.
#pragma once
#include "had_common.cuh"

namespace qp {

#ifdef QP_PROFILE_PHASES
// debug build only: %globaltimer stamps inside produce_x (CTA-wide thread 0), read back with qp_debug_xphases
__device__ unsigned long long g_xphase[256][8];
__device__ __forceinline__ void xphase_stamp(int i) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_xphase[blockIdx.x][i] = t;
    }
}
#define QP_XPHASE(i) xphase_stamp(i)
#else
#define QP_XPHASE(i)
#endif

struct XProd {
    int mode;                 // 0: x is given (plain staging); 1: produce x as described above; 2: ... with the LL operand
    const __half *src;        // h (n)
    __half *h_out;            // h' destination or NULL
    const float *acc;         // optional fp32 accumulators (n) of the previous projection
    const __half *wscale;     // its per-row scales (n)
    float acc_scale;
    const __half *norm_w;     // optional RMSNorm weight (n)
    float eps;
    const __half *su;         // optional signs (n)
    float had_scale;
    __half *x_out;            // optional copy of x in natural order (for sibling projections sharing x)
    float *zero1;             // accumulators to clear for later launches (not touched by this launch or its predecessor)
    int zero1_count;
    float *zero2;
    int zero2_count;
    int m, Kf;                // Hadamard block / 28-factor of n
    // row-sharded mode: the gathered operand arrives from the peers as "LL" words {2 x fp16, epoch, 2 x fp16, epoch} in a
    // local receive buffer (xchg_send_ll_kernel of every rank stores there over NVLink) and is polled here, in every CTA
    const uint4 *ll;          // n / 4 entries, or NULL
    const unsigned *ll_epoch; // epoch the entries must carry (written by the local sender launched in front of this kernel)
    int ll_kind;              // 1: the LL values are fp16(acc) (src is read normally); 2: they are src
    unsigned long long ll_spin_cycles;  // 0 = poll forever
};

// LL entries are read at L2 (ld.volatile), where the peers' NVLink stores land; an 8-byte {data, flag} half of an entry is
// written by one store, so a matching flag implies its data.  xp_ll_load issues the read; xp_ll_wait re-reads until both flags
// carry `ep` (normally zero times: the sender kernel of this rank has already completed, the peers' are a few us apart).
__device__ __forceinline__ uint4 xp_ll_load(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 xp_ll_wait(uint4 v, const uint4 *p, unsigned ep, unsigned long long limit) {
    const long long t0 = clock64();
    while (v.y != ep || v.w != ep) {
        if (limit != 0ull && (unsigned long long)(clock64() - t0) > limit) __trap();  // a peer died: fail, do not hang the GPU
        v = xp_ll_load(p);
    }
    return make_uint2(v.x, v.z);
}

__device__ __forceinline__ float xp_block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;  // `red` is written once per kernel: no barrier needed behind the reads
}

__device__ __forceinline__ void xp_unpack4(const uint2 u, float (&f)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__device__ __forceinline__ void xp_had_warp128(float (&y)[4]) {
    const float a = y[0] + y[1], b = y[0] - y[1], c = y[2] + y[3], d = y[2] - y[3];
    y[0] = a + c; y[1] = b + d; y[2] = a - c; y[3] = b - d;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        // o - y for the upper lane of a pair, y + o for the lower one, as ONE fused multiply-add by +-1 (bit-identical; the
        // select form costs two adds and a select per element and stage, and this runs in every CTA of a fused launch)
        const float sgn = ((lane >> s) & 1) ? -1.f : 1.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float o = __shfl_xor_sync(0xffffffffu, y[e], 1 << s);
            y[e] = fmaf(y[e], sgn, o);
        }
    }
}

#ifndef QP_XP_FINAL32
#define QP_XP_FINAL32 1
#endif
// strides 2^lh ... of a power-of-two n = p.m done in registers down to the last five, which one radix-32 pass finishes; its
// outputs are scaled, rounded to fp16 and stored in the B-fragment order of stage_x(): element i = 32*kh + pp lands in half
// ((kh*16 + q*4 + kl*2 + b) << 1) | e with kl = pp >> 4, b = (pp >> 3) & 1, q = (pp >> 1) & 3, e = pp & 1 -- a permutation inside
// each block of 32, so a thread's 32 results (stride >= 128 apart) share the in-block position.  No trailing barrier.
// FIXED: the pass starts at stride 2^7 with n = 4096 known at compile time (the Llama-8B case: every address offset is an
// immediate); WITH_XO: CTA 0 also stores x in natural order -- a separate instantiation, so that the other 147 CTAs do not issue 32
// predicated-off global stores and their address arithmetic (this pass runs on 4 warps while the other 20 wait at the barrier)
template <bool FIXED, bool WITH_XO>
__device__ __forceinline__ void xp_final32_pass(const float *v, __half *xh, __half *xo, int n, int lh_rt, float scale) {
    // Two threads per column set: both read the 32 inputs, the top stride is done first and each keeps one half of its results
    // (sums or differences, warp-uniform), then finishes 16 values.  A single warp per scheduler runs this dependent chain at
    // about a third of the issue rate (measured: 170 instructions fewer on this path were worth 0.28 us per launch), so two
    // shorter chains on 8 warps beat one long chain on 4.  (The top stride first instead of last: same transform, fp32 sums in
    // a different order than the stand-alone Hadamard kernel -- a 1e-7 effect, three orders below the run-to-run noise of the
    // accumulators it is applied to.)
    const int lh = FIXED ? 7 : lh_rt;
    const int lcnt = FIXED ? 7 : (31 - __clz(n)) - 5;  // log2 of the number of column sets
    const int h = 1 << lh;
    unsigned short *xh16 = reinterpret_cast<unsigned short *>(xh);
    unsigned short *xo16 = reinterpret_cast<unsigned short *>(xo);
    for (int w = threadIdx.x; w < (2 << lcnt); w += blockDim.x) {
        const int idx = w & ((1 << lcnt) - 1), half = w >> lcnt;
        const float sg = half ? -1.f : 1.f;
        const int low = idx & (h - 1), hi = idx >> lh;
        const int i0 = (hi << (5 + lh)) | low;
        float r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = fmaf(v[i0 + ((k + 16) << lh)], sg, v[i0 + (k << lh)]);
#pragma unroll
        for (int s = 1; s < 16; s <<= 1) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if ((k & s) == 0) {
                    const float a = r[k], b = r[k | s];
                    r[k] = a + b;
                    r[k | s] = a - b;
                }
            }
        }
        const int i1 = i0 + (half << (4 + lh));
        const int pp = i1 & 31;
        const int d0 = ((((i1 >> 5) << 4) + (((pp >> 1) & 3) << 2) + ((pp >> 4) << 1) + ((pp >> 3) & 1)) << 1) | (pp & 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned short o = __half_as_ushort(__float2half_rn(r[k] * scale));
            xh16[d0 + (k << lh)] = o;
            if (WITH_XO) xo16[i1 + (k << lh)] = o;
        }
    }
}

__device__ __forceinline__ void xp_final32_stage(float *v, uint32_t *xs, int n, int lh, const XProd &p) {
    const int lm = 31 - __clz(n);
    while (lm - lh > 5) {  // n > 4096: bring the remaining strides down to five
        if (lm - lh >= 8) fwht_pass<3>(v, n, lh), lh += 3;
        else if (lm - lh == 7) fwht_pass<2>(v, n, lh), lh += 2;
        else fwht_pass<1>(v, n, lh), lh += 1;
        __syncthreads();
    }
    __half *xh = reinterpret_cast<__half *>(xs);
    if (p.x_out && blockIdx.x == 0) xp_final32_pass<false, true>(v, xh, p.x_out, n, lh, p.had_scale);
    else if (n == 4096) xp_final32_pass<true, false>(v, xh, nullptr, n, lh, p.had_scale);
    else xp_final32_pass<false, false>(v, xh, nullptr, n, lh, p.had_scale);
}

__device__ __forceinline__ void xp_zero_slice(float *p, int count) {
    if (!p) return;
    float4 *p4 = reinterpret_cast<float4 *>(p);
    const int n4 = count >> 2;
    const int per = (n4 + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(n4, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// accumulator-clearing duty of a fused launch, run BEFORE the programmatic-dependency wait (measured: ~0.4-0.7 us on the
// critical path when done after it).  Safe for buffers that neither this launch nor the immediately preceding one touches:
// a kernel's pre-wait code only starts once every CTA of its predecessor has passed its own wait, i.e. once the
// predecessor's predecessor is complete.
__device__ __forceinline__ void xp_zero(const XProd &p) {
    xp_zero_slice(p.zero1, p.zero1_count);
    xp_zero_slice(p.zero2, p.zero2_count);
}

// the inputs of produce_x that do NOT depend on the preceding kernel (per-row scales, norm weight, signs): a GEMV kernel
// loads them before its programmatic-dependency wait, next to the weight prefetch, so that only h / acc are fetched in the
// dependent part (every CTA reads the same few KB at the same time: the fewer of them after the wait, the better)
template <int CH>
struct XPre {
    uint2 wv[CH], nv[CH], sv[CH];
};
template <int CH>
__device__ __forceinline__ void xp_preload(XPre<CH> &r, const XProd &p, int n) {
    const int nch = n >> 2, T = blockDim.x;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * T;
        const bool ok = c < nch;
        r.wv[j] = (ok && (p.acc || p.ll_kind == 1)) ? __ldg(reinterpret_cast<const uint2 *>(p.wscale) + c) : make_uint2(0u, 0u);
        r.nv[j] = (ok && p.norm_w) ? __ldg(reinterpret_cast<const uint2 *>(p.norm_w) + c) : make_uint2(0u, 0u);
        r.sv[j] = (ok && p.su) ? __ldg(reinterpret_cast<const uint2 *>(p.su) + c) : make_uint2(0u, 0u);
    }
}

// produce x (n = K values) into xs in the B-fragment order stage_x() uses (bs = 1).  v: n floats of shared scratch.
// CH = ceil(n / 4 / blockDim.x) chunks of 4 consecutive elements per thread.
// LL (compile time): one operand is polled out of the row-sharded receive buffer (p.ll); a separate instantiation because the
// polling registers and branches cost the single-GPU prologue 1.3 % of a decode step when they are merely present (measured)
template <int CH, bool LL = false>
__device__ __forceinline__ void produce_x(uint32_t *xs, float *v, float *red, const XProd &p, int n, const XPre<CH> &pre) {
    const int nch = n >> 2, T = blockDim.x;
    QP_XPHASE(0);
    uint2 hv[CH];
    float4 av[CH];
    const uint2 (&wv)[CH] = pre.wv, (&nv)[CH] = pre.nv, (&sv)[CH] = pre.sv;
    const uint4 *const p_ll = LL ? p.ll : nullptr;
    const unsigned ll_ep = p_ll ? __ldcg(p.ll_epoch) : 0u;
    uint4 lv[CH];
    if (p_ll) {  // row-sharded: one of the two operands comes out of the LL receive buffer; all reads in flight before the first check
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = threadIdx.x + j * T;
            lv[j] = c < nch ? xp_ll_load(p_ll + c) : make_uint4(0u, ll_ep, 0u, ll_ep);
            if (p.ll_kind != 2) hv[j] = c < nch ? __ldcg(reinterpret_cast<const uint2 *>(p.src) + c) : make_uint2(0u, 0u);
        }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * T;
        const bool ok = c < nch;
        if (p_ll) {
            const uint2 g = xp_ll_wait(lv[j], p_ll + (ok ? c : 0), ll_ep, p.ll_spin_cycles);
            if (p.ll_kind == 2) {
                hv[j] = g;
                av[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float a4[4];
                xp_unpack4(g, a4);  // fp16(acc): exactly what the arithmetic below makes of an fp32 accumulator first
                av[j] = make_float4(a4[0], a4[1], a4[2], a4[3]);
            }
            continue;
        }
        // activations / accumulators are rewritten by other kernels within one decode step and kernels overlap under programmatic
        // dependent launch: read them at L2 (ld.global.cg), never through the non-coherent L1 path (measured round 2: with
        // ld.global.nc a CTA could see a stale line of `acc` / `src` and the fused launch list diverged from the un-fused one)
        hv[j] = ok ? __ldcg(reinterpret_cast<const uint2 *>(p.src) + c) : make_uint2(0u, 0u);
        av[j] = (ok && p.acc) ? __ldcg(reinterpret_cast<const float4 *>(p.acc) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    QP_XPHASE(1);  // loads issued, zero slices stored
    // fp16 arithmetic on packed pairs (the rounding points of the reference's fp16 tensors: product by the row scale, by the
    // shared scale, residual add -- each rounded separately, never contracted into an fma).  Chunks past the end (the last of a
    // warp's CH chunks when n / 4 is not a multiple of the CTA size) are skipped warp-uniformly: the prologue runs in all 24
    // warps of all CTAs and is issue-bound like the loop behind it.
    const __half2 hs2 = __float2half2_rn(p.acc_scale);
    float y[CH][4];
    float ss = 0.f;
    const int wbase = (int)(threadIdx.x & ~31u);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * T;
        if (wbase + j * T >= nch) {
#pragma unroll
            for (int e = 0; e < 4; ++e) y[j][e] = 0.f;
            continue;
        }
        __half2 h01 = *reinterpret_cast<const __half2 *>(&hv[j].x), h23 = *reinterpret_cast<const __half2 *>(&hv[j].y);
        if (p.acc || p.ll_kind == 1) {
            const __half2 w01 = *reinterpret_cast<const __half2 *>(&wv[j].x), w23 = *reinterpret_cast<const __half2 *>(&wv[j].y);
            const __half2 a01 = __floats2half2_rn(av[j].x, av[j].y), a23 = __floats2half2_rn(av[j].z, av[j].w);
            h01 = __hadd2_rn(h01, __hmul2_rn(__hmul2_rn(a01, w01), hs2));
            h23 = __hadd2_rn(h23, __hmul2_rn(__hmul2_rn(a23, w23), hs2));
            if (p.h_out && blockIdx.x == 0 && c < nch) {
                uint2 u;
                u.x = *reinterpret_cast<const uint32_t *>(&h01);
                u.y = *reinterpret_cast<const uint32_t *>(&h23);
                reinterpret_cast<uint2 *>(p.h_out)[c] = u;
            }
        }
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        y[j][0] = f01.x; y[j][1] = f01.y; y[j][2] = f23.x; y[j][3] = f23.y;
#pragma unroll
        for (int e = 0; e < 4; ++e) ss += y[j][e] * y[j][e];
    }
    QP_XPHASE(2);  // inputs arrived, residual added
    if (p.norm_w) {
        const float rstd = rsqrtf(xp_block_sum(ss, red) / (float)n + p.eps);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (wbase + j * T >= nch) continue;
            const __half2 n01 = *reinterpret_cast<const __half2 *>(&nv[j].x), n23 = *reinterpret_cast<const __half2 *>(&nv[j].y);
            const float2 f01 = __half22float2(__hmul2_rn(n01, __floats2half2_rn(y[j][0] * rstd, y[j][1] * rstd)));
            const float2 f23 = __half22float2(__hmul2_rn(n23, __floats2half2_rn(y[j][2] * rstd, y[j][3] * rstd)));
            y[j][0] = f01.x; y[j][1] = f01.y; y[j][2] = f23.x; y[j][3] = f23.y;
        }
    }
    QP_XPHASE(3);  // normalised
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * T;
        if (wbase + j * T >= nch) continue;
        if (p.su) {
            float s4[4];
            xp_unpack4(sv[j], s4);
#pragma unroll
            for (int e = 0; e < 4; ++e) y[j][e] *= s4[e];
        }
        xp_had_warp128(y[j]);
        if (c < nch) reinterpret_cast<float4 *>(v)[c] = make_float4(y[j][0], y[j][1], y[j][2], y[j][3]);
    }
    __syncthreads();
    QP_XPHASE(4);  // strides < 128 done, in shared memory
#if QP_XP_FINAL32
    if (p.Kf == 1 && p.m >= 4096) {
        // power-of-two n >= 4096 (every fused launch of the Llama shapes): the last five strides as ONE radix-32 pass whose
        // results go straight into the fragment-ordered fp16 stage -- two barriers and two shared-memory round trips fewer than
        // radix-8 + radix-4 passes followed by a separate staging pass.  Same butterflies in the same order: bit-identical.
        xp_final32_stage(v, xs, n, 7, p);
        QP_XPHASE(5);
        return;  // the caller's barrier publishes xs
    }
#endif
    hadamard_smem(v, n, p.m, p.Kf, 7);
    QP_XPHASE(5);  // Hadamard done
    // fp16 x in B-fragment order: element i = 32*kh + 16*kl + 8*b + 2*q + e  ->  word ((kh*4 + q)*4 + kl*2 + b), half e
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * T;
        if (c < nch) {
            const float4 f = reinterpret_cast<const float4 *>(v)[c];
            const __half2 lo = __floats2half2_rn(f.x * p.had_scale, f.y * p.had_scale);
            const __half2 hi = __floats2half2_rn(f.z * p.had_scale, f.w * p.had_scale);
            const int i = c << 2;
            const int kh = i >> 5, pp = i & 31, kl = pp >> 4, b = (pp >> 3) & 1, q = (pp >> 1) & 3;  // q in {0, 2}
            uint32_t *d = xs + (kh * 16 + kl * 2 + b);
            d[q * 4] = *reinterpret_cast<const uint32_t *>(&lo);
            d[(q + 1) * 4] = *reinterpret_cast<const uint32_t *>(&hi);
            if (p.x_out && blockIdx.x == 0) {
                uint2 u;
                u.x = *reinterpret_cast<const uint32_t *>(&lo);
                u.y = *reinterpret_cast<const uint32_t *>(&hi);
                reinterpret_cast<uint2 *>(p.x_out)[c] = u;
            }
        }
    }
}

// dispatch on the chunk count (n <= 15360 for 768 threads; larger n does not fit the shared-memory budget anyway);
// this form loads everything after the caller's dependency wait
template <bool LL = false>
__device__ __forceinline__ void produce_x_dispatch(uint32_t *xs, float *v, float *red, const XProd &p, int n) {
    const int ch = ((n >> 2) + blockDim.x - 1) / blockDim.x;
    if (ch <= 2) {
        XPre<2> pre;
        xp_preload<2>(pre, p, n);
        produce_x<2, LL>(xs, v, red, p, n, pre);
    } else {  // host guarantees n <= 5 * 4 * blockDim.x
        XPre<5> pre;
        xp_preload<5>(pre, p, n);
        produce_x<5, LL>(xs, v, red, p, n, pre);
    }
}

}  // namespace qp
