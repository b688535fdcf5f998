// decode_kernels.cu -- the glue of one bs=1 decode step around the quantized GEMVs, as hand-written kernels
// (the reference leaves this to torch.compile/Inductor: eval/measure_latency.py:223-225; north_star forbids Triton).
//
//   qp_fused_norm_had : [h += fp16(acc)*Wscale*s] -> [RMSNorm] -> [*SU -> FWHT(/28-factor) -> *scale] -> fp16 x
//                       (IncoherentSdpaAttention.compute_qkv/compute_o prologues, IncoherentMLP.compute_ug prologue,
//                        the residual adds and LlamaRMSNorm of the decoder layer; lib/linear/incoherent_linear.py:76-108,
//                        324-338; model/llama.py LlamaDecoderLayer.forward)
//   qp_silu_mul_had   : up|gate epilogue -> SiLU*mul -> *SU -> FWHT -> fp16 x          (compute_ug tail + compute_dp head)
//   qp_rope_attention : q/k/v epilogue (Wscale) -> RoPE -> KV-cache append -> softmax(qK^T)V for one new token
//   qp_gemv_f16       : fp16 lm_head GEMV (128256 x 4096 is 27% of the bytes of a token)
//   qp_argmax, qp_embed, qp_step_advance : sampling / embedding / device-side position counter so a whole decode step is
//                       one CUDA graph that is replayed per token.
// Single-CTA kernels here are latency-bound by design (a few KB of data); they exist to remove launches.
#include "had_common.cuh"

namespace qp {

constexpr int kDecThreads = 512;

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (int)(blockDim.x >> 5)) ? red[l] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    return t;
}

__device__ __forceinline__ void zero_words(float *p, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) p[i] = 0.f;
}

// see header comment.  All pointers optional except x_out/h.  One CTA.
__global__ void __launch_bounds__(kDecThreads, 1)
fused_norm_had_kernel(__half *__restrict__ x_out, __half *__restrict__ h, int h_writeback,
                      const float *__restrict__ acc, const __half *__restrict__ wscale, float acc_scale,
                      const __half *__restrict__ norm_w, float eps, const __half *__restrict__ su, int n, int m, int Kf,
                      float had_scale, int do_had, float *__restrict__ zero_ptr, int zero_count) {
    extern __shared__ __align__(16) float v[];
    __shared__ float red[32];
    pdl_wait();
    pdl_launch_dependents();
    if (zero_ptr) zero_words(zero_ptr, zero_count);
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        __half hv = h[i];
        if (acc) {
            const __half t = __hmul(__hmul(__float2half(acc[i]), wscale[i]), __float2half(acc_scale));
            hv = __hadd(hv, t);
            if (h_writeback) h[i] = hv;
        }
        const float f = __half2float(hv);
        v[i] = f;
        ss += f * f;
    }
    if (norm_w) {
        const float tot = block_sum(ss, red);
        const float rstd = rsqrtf(tot / (float)n + eps);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            // HF LlamaRMSNorm: weight * (x.float() * rstd).to(fp16)
            const __half xn = __float2half(v[i] * rstd);
            v[i] = __half2float(__hmul(norm_w[i], xn));
        }
    }
    if (su) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] *= __half2float(su[i]);
    }
    __syncthreads();
    if (do_had) hadamard_smem(v, n, m, Kf);
    for (int i = threadIdx.x; i < n; i += blockDim.x) x_out[i] = __float2half(v[i] * had_scale);
}

// acc = [up (I) | gate (I)] fp32 -> y = silu(gate)*up (fp16 rounding points as the reference graph) -> *su -> had -> x
__global__ void __launch_bounds__(kDecThreads, 1)
silu_mul_had_kernel(__half *__restrict__ x_out, const float *__restrict__ acc, const __half *__restrict__ wscale,
                    float acc_scale, const __half *__restrict__ su, int I, int m, int Kf, float had_scale,
                    float *__restrict__ zero_ptr, int zero_count) {
    extern __shared__ __align__(16) float v[];
    pdl_wait();
    pdl_launch_dependents();
    if (zero_ptr) zero_words(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
        const __half up = __hmul(__hmul(__float2half(acc[i]), wscale[i]), hs);
        const __half gate = __hmul(__hmul(__float2half(acc[I + i]), wscale[I + i]), hs);
        const float g = __half2float(gate);
        const __half act = __float2half(g / (1.f + __expf(-g)));
        float y = __half2float(__hmul(act, up));
        if (su) y *= __half2float(su[i]);
        v[i] = y;
    }
    __syncthreads();
    hadamard_smem(v, I, m, Kf);
    for (int i = threadIdx.x; i < I; i += blockDim.x) x_out[i] = __float2half(v[i] * had_scale);
}

// One CTA per query head.  acc_qkv: fp32 [q (H*D) | k (Hkv*D) | v (Hkv*D)] raw GEMV sums; wscale same layout.
// RoPE (HF rotate_half convention, model/llama.py apply_rotary_pos_emb) with inv_freq table (D/2 floats, llama3 scaling
// already applied by the host).  KV cache: fp16 [2][max_seq][Hkv][D] for this layer.  pos read from device memory.
__global__ void __launch_bounds__(128, 1)
rope_attention_kernel(__half *__restrict__ attn_out, const float *__restrict__ acc_qkv, const __half *__restrict__ wscale,
                      float acc_scale, const float *__restrict__ inv_freq, __half *__restrict__ kcache,
                      __half *__restrict__ vcache, const int *__restrict__ pos_ptr, int H, int Hkv, int D, int max_seq,
                      float *__restrict__ zero_ptr, int zero_count) {
    extern __shared__ __align__(16) float sm[];  // q[D] | knew[D] | vnew[D] | scores[max_seq]
    __shared__ float red[32];
    float *q = sm, *kn = sm + D, *vn = sm + 2 * D, *sc = sm + 3 * D;
    pdl_wait();
    pdl_launch_dependents();
    const int head = blockIdx.x, kvh = head / (H / Hkv);
    const int pos = *pos_ptr;
    if (zero_ptr && blockIdx.x == 0) zero_words(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
    const int d = threadIdx.x;
    if (d < D) {
        auto val = [&](int idx) { return __half2float(__hmul(__hmul(__float2half(acc_qkv[idx]), wscale[idx]), hs)); };
        const int half = D / 2;
        const int pd = d < half ? d + half : d - half;
        const float sgn = d < half ? -1.f : 1.f;
        double snd, csd;  // precise range reduction (the build uses --use_fast_math)
        sincos((double)pos * (double)inv_freq[d % half], &snd, &csd);
        const float sn = (float)snd, cs = (float)csd;
        // fp16 rounding of cos/sin and of the products as the fp16 reference graph does
        const float c16 = __half2float(__float2half(cs)), s16 = __half2float(__float2half(sn));
        const float qa = val(head * D + d), qb = val(head * D + pd);
        const float ka = val(H * D + kvh * D + d), kb = val(H * D + kvh * D + pd);
        q[d] = __half2float(__float2half(qa * c16 + sgn * qb * s16));
        const __half kr = __float2half(ka * c16 + sgn * kb * s16);
        const __half vv = __float2half(val((H + Hkv) * D + kvh * D + d));
        kn[d] = __half2float(kr);
        vn[d] = __half2float(vv);
        if (head % (H / Hkv) == 0) {  // one CTA of the group appends to the cache
            kcache[((size_t)pos * Hkv + kvh) * D + d] = kr;
            vcache[((size_t)pos * Hkv + kvh) * D + d] = vv;
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)D);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // scores over cached positions [0, pos) + the new token (from shared memory, not the cache: no cross-CTA race)
    for (int t = warp; t <= pos; t += nw) {
        float s = 0.f;
        if (t < pos) {
            const __half *kr = kcache + ((size_t)t * Hkv + kvh) * D;
            for (int i = lane; i < D; i += 32) s += q[i] * __half2float(kr[i]);
        } else {
            for (int i = lane; i < D; i += 32) s += q[i] * kn[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sc[t] = s * scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = threadIdx.x; t <= pos; t += blockDim.x) mx = fmaxf(mx, sc[t]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int t = threadIdx.x; t <= pos; t += blockDim.x) {
        const float e = __expf(sc[t] - mx);
        sc[t] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    __syncthreads();
    if (d < D) {
        float o = 0.f;
        for (int t = 0; t < pos; ++t) o += sc[t] * __half2float(vcache[((size_t)t * Hkv + kvh) * D + d]);
        o += sc[pos] * vn[d];
        attn_out[head * D + d] = __float2half(o / sum);
    }
}

// out (rows) fp32 = W (rows x K, fp16 row-major) @ x (K fp16).  One warp per row, 16-byte streaming loads, x in smem.
__global__ void __launch_bounds__(512, 2)
gemv_f16_kernel(float *__restrict__ out, const __half *__restrict__ W, const __half *__restrict__ x, int rows, int K) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint4 *xs = reinterpret_cast<uint4 *>(smraw);
    pdl_wait();
    pdl_launch_dependents();
    const int kq = K / 8;
    for (int i = threadIdx.x; i < kq; i += blockDim.x) xs[i] = reinterpret_cast<const uint4 *>(x)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int r = gw; r < rows; r += nw) {
        const uint32_t *wr = reinterpret_cast<const uint32_t *>(W + (size_t)r * K);
        float acc = 0.f;
        for (int i = lane; i < kq; i += 32 * 4) {
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = i + u * 32;
                w[u] = (j < kq) ? ldg_stream_u128(wr + 4 * j) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = i + u * 32;
                if (j < kq) {
                    const uint4 xv = xs[j];
                    const __half2 *wh = reinterpret_cast<const __half2 *>(&w[u]);
                    const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = __half22float2(wh[e]), b = __half22float2(xh[e]);
                        acc = fmaf(a.x, b.x, acc);
                        acc = fmaf(a.y, b.y, acc);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[r] = acc;
    }
}

// two-stage argmax: stage 1 per-CTA partials (value, index), stage 2 by the last CTA (ticket counter)
__global__ void argmax_kernel(int *__restrict__ token_out, const float *__restrict__ logits, int n, float *part_val,
                              int *part_idx, unsigned *ticket) {
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ bool last;
    pdl_wait();
    pdl_launch_dependents();
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = logits[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    auto reduce = [&](float &v, int &ix) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
            if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
        }
    };
    reduce(bv, bi);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = bv; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        bv = l < (int)(blockDim.x >> 5) ? sv[l] : -INFINITY;
        bi = l < (int)(blockDim.x >> 5) ? si[l] : 0x7fffffff;
        reduce(bv, bi);
        if (l == 0) {
            part_val[blockIdx.x] = bv;
            part_idx[blockIdx.x] = bi;
            __threadfence();
            last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (last && w == 0) {
        __threadfence();
        bv = -INFINITY;
        bi = 0x7fffffff;
        for (int i = l; i < (int)gridDim.x; i += 32) {
            const float v = part_val[i];
            const int ix = part_idx[i];
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
        reduce(bv, bi);
        if (l == 0) {
            *token_out = bi;
            *ticket = 0;
        }
    }
}

__global__ void embed_kernel(__half *__restrict__ h, const __half *__restrict__ table, const int *__restrict__ token, int n) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t row = (size_t)(*token);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) h[i] = table[row * n + i];
}

__global__ void step_advance_kernel(int *pos, int *history, const int *token, int max_hist) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int p = *pos;
        if (history && p < max_hist) history[p] = *token;
        *pos = p + 1;
    }
}

}  // namespace qp

using namespace qp;

static int had_dims(int n, int &m, int &Kf) {
    Kf = 1;
    m = n;
    if ((n & (n - 1)) != 0) {
        QP_CHECK_ARG(n % 28 == 0 && (((n / 28) & (n / 28 - 1)) == 0), "Hadamard size %d is neither 2^k nor 28*2^k", n);
        Kf = 28;
        m = n / 28;
    }
    return QP_OK;
}

extern "C" int qp_fused_norm_had(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                                 float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n,
                                 float had_scale, int do_had, float *zero_ptr, int zero_count, void *stream) {
    QP_CHECK_ARG(x_out_f16 && h_f16, "NULL pointer argument");
    QP_CHECK_ARG(!acc || wscale_f16, "acc given without wscale");
    int m, Kf;
    int rc = had_dims(n, m, Kf);
    if (rc != QP_OK && do_had) return rc;
    const size_t smem = (size_t)n * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "n = %d too large", n);
    static bool configured = false;
    if (!configured) {
        QP_CUDA(cudaFuncSetAttribute(fused_norm_had_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
        configured = true;
    }
    QP_CUDA(launch_pdl(fused_norm_had_kernel, dim3(1), dim3(kDecThreads), smem, (cudaStream_t)stream, (__half *)x_out_f16,
                       (__half *)h_f16, h_writeback, acc, (const __half *)wscale_f16, acc_scale,
                       (const __half *)norm_w_f16, eps, (const __half *)su_f16, n, m, Kf, had_scale, do_had, zero_ptr,
                       zero_count));
    return check_launch("fused_norm_had");
}

extern "C" int qp_silu_mul_had(void *x_out_f16, const float *acc, const void *wscale_f16, float acc_scale,
                               const void *su_f16, int I, float had_scale, float *zero_ptr, int zero_count, void *stream) {
    QP_CHECK_ARG(x_out_f16 && acc && wscale_f16, "NULL pointer argument");
    int m, Kf;
    int rc = had_dims(I, m, Kf);
    if (rc != QP_OK) return rc;
    const size_t smem = (size_t)I * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "I = %d too large", I);
    static bool configured = false;
    if (!configured) {
        QP_CUDA(cudaFuncSetAttribute(silu_mul_had_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
        configured = true;
    }
    QP_CUDA(launch_pdl(silu_mul_had_kernel, dim3(1), dim3(kDecThreads), smem, (cudaStream_t)stream, (__half *)x_out_f16,
                       acc, (const __half *)wscale_f16, acc_scale, (const __half *)su_f16, I, m, Kf, had_scale, zero_ptr,
                       zero_count));
    return check_launch("silu_mul_had");
}

extern "C" int qp_rope_attention(void *attn_out_f16, const float *acc_qkv, const void *wscale_f16, float acc_scale,
                                 const float *inv_freq, void *kcache_f16, void *vcache_f16, const int *pos_ptr, int H,
                                 int Hkv, int D, int max_seq, float *zero_ptr, int zero_count, void *stream) {
    QP_CHECK_ARG(attn_out_f16 && acc_qkv && wscale_f16 && inv_freq && kcache_f16 && vcache_f16 && pos_ptr, "NULL pointer");
    QP_CHECK_ARG(D <= 128 && D % 2 == 0 && H % Hkv == 0, "unsupported head geometry H=%d Hkv=%d D=%d", H, Hkv, D);
    const size_t smem = (size_t)(3 * D + max_seq) * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "max_seq = %d too large for the single-pass attention kernel", max_seq);
    static bool configured = false;
    if (!configured) {
        QP_CUDA(cudaFuncSetAttribute(rope_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
        configured = true;
    }
    QP_CUDA(launch_pdl(rope_attention_kernel, dim3(H), dim3(128), smem, (cudaStream_t)stream, (__half *)attn_out_f16,
                       acc_qkv, (const __half *)wscale_f16, acc_scale, inv_freq, (__half *)kcache_f16,
                       (__half *)vcache_f16, pos_ptr, H, Hkv, D, max_seq, zero_ptr, zero_count));
    return check_launch("rope_attention");
}

extern "C" int qp_gemv_f16(float *out, const void *W_f16, const void *x_f16, int rows, int K, void *stream) {
    QP_CHECK_ARG(out && W_f16 && x_f16, "NULL pointer argument");
    QP_CHECK_ARG(K % 8 == 0 && (size_t)K * 2 <= 96 * 1024, "K = %d unsupported", K);
    static bool configured = false;
    if (!configured) {
        QP_CUDA(cudaFuncSetAttribute(gemv_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured = true;
    }
    QP_CUDA(launch_pdl(gemv_f16_kernel, dim3(sm_count() * 2), dim3(512), (size_t)K * 2, (cudaStream_t)stream, out,
                       (const __half *)W_f16, (const __half *)x_f16, rows, K));
    return check_launch("gemv_f16");
}

extern "C" int qp_argmax(int *token_out, const float *logits, int n, void *scratch /* >= 4 KiB, zeroed once */, void *stream) {
    QP_CHECK_ARG(token_out && logits && scratch, "NULL pointer argument");
    const int blocks = 128;
    float *pv = (float *)scratch;
    int *pi = (int *)scratch + 256;
    unsigned *ticket = (unsigned *)scratch + 512;
    QP_CUDA(launch_pdl(argmax_kernel, dim3(blocks), dim3(512), 0, (cudaStream_t)stream, token_out, logits, n, pv, pi, ticket));
    return check_launch("argmax");
}

extern "C" int qp_embed(void *h_f16, const void *table_f16, const int *token, int n, void *stream) {
    QP_CHECK_ARG(h_f16 && table_f16 && token, "NULL pointer argument");
    QP_CUDA(launch_pdl(embed_kernel, dim3(4), dim3(256), 0, (cudaStream_t)stream, (__half *)h_f16,
                       (const __half *)table_f16, token, n));
    return check_launch("embed");
}

extern "C" int qp_step_advance(int *pos, int *history, const int *token, int max_hist, void *stream) {
    QP_CHECK_ARG(pos && token, "NULL pointer argument");
    QP_CUDA(launch_pdl(step_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, pos, history, token, max_hist));
    return check_launch("step_advance");
}
