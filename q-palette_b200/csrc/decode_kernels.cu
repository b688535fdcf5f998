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
// Every buffer that another kernel of the same decode step writes (h, the fp32 accumulators, attention output, x, logits, token,
// position) is read with ld.global.cg (__ldcg): kernels overlap under programmatic dependent launch and the L1 is not coherent.
#include "had_common.cuh"

namespace qp {

constexpr int kDecThreads = 1024;

// spin limit (SM clock cycles) of the cross-GPU / cross-CTA flag waits before the kernel traps instead of hanging the GPU;
// 0 = wait forever.  Host-settable (qp_set_spin_timeout_ms): a rank under a profiler, a first-touch page fault or an
// 80-layer graph instantiation can legitimately lag its peers by seconds.
__device__ unsigned long long g_spin_limit_cycles = 120000000000ull;  // ~60 s at 2 GHz
__device__ __forceinline__ bool spin_expired(long long t0) {
    const unsigned long long lim = g_spin_limit_cycles;
    return lim != 0ull && (unsigned long long)(clock64() - t0) > lim;
}

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

__device__ __forceinline__ void zero_words4(float *p, int count) {
    float4 *p4 = reinterpret_cast<float4 *>(p);
    for (int i = threadIdx.x; i < count / 4; i += blockDim.x) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = (count & ~3) + threadIdx.x; i < count; i += blockDim.x) p[i] = 0.f;
}

__device__ __forceinline__ void unpack4(const uint2 u, float (&f)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
    uint2 u;
    *reinterpret_cast<__half2 *>(&u.x) = __floats2half2_rn(f[0], f[1]);
    *reinterpret_cast<__half2 *>(&u.y) = __floats2half2_rn(f[2], f[3]);
    return u;
}
// fp16(fp16(fp16(acc) * wscale) * s): the rounding points of `linear(x).half() * Wscale * scale`
__device__ __forceinline__ float scaled_acc(float acc, float ws, __half hs) {
    return __half2float(__hmul(__hmul(__float2half(acc), __float2half(ws)), hs));
}

// butterfly stages of element stride 1, 2 (registers) and 4..64 (warp shuffles) on a warp's 128 consecutive elements:
// lane l holds elements 4l..4l+3.  Shared memory only sees strides >= 128, which are bank-conflict free.
__device__ __forceinline__ void had_warp128(float (&y)[4]) {
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
constexpr int kHadLh0 = 7;

// ---- all-gather over NVLink peer memory, fused into the consumer kernel ------------------------------------------------
// Every rank maps the same "exchange region" of every peer (CUDA IPC).  A buffer that is produced row-sharded (rank r owns
// bytes [r*slice, (r+1)*slice)) is completed in place: the consumer kernel first stores its own slice straight into all
// peers' copies, publishes a per-(site, source) flag with release.sys semantics, then waits for the peers' flags.
// One call site per (layer, gather point); the flag value is the site's epoch (how many times the site has run), kept in
// local device memory, so CUDA-graph replays need no reset.  Replaces one ncclAllGather (~15-20 us in a graph) by
// ~3 us of stores + one NVLink round trip.
struct XchgDev {
    unsigned char *const *peer_base;   // device array [nranks]: base of the exchange region on each rank (own included)
    unsigned *const *peer_flags;       // device array [nranks]: flags[site * nranks + source] on each rank
    unsigned *epoch;                   // local: epoch[site]
    long long offset;                  // of the gathered buffer inside the region
    int slice_bytes, rank, nranks, site;
};

__device__ __forceinline__ void peer_allgather(const XchgDev &xc) {
    if (xc.nranks <= 1) return;
    const unsigned ep = *xc.epoch + 1u;
    const size_t off = (size_t)xc.offset + (size_t)xc.rank * xc.slice_bytes;
    const uint4 *src = reinterpret_cast<const uint4 *>(xc.peer_base[xc.rank] + off);
    const int n16 = xc.slice_bytes >> 4;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) {
        const uint4 val = __ldcg(src + i);  // produced by the previous kernel (atomics / stores): L2 is the point of truth
        for (int q = 1; q < xc.nranks; ++q) {
            const int peer = (xc.rank + q) % xc.nranks;
            reinterpret_cast<uint4 *>(xc.peer_base[peer] + off)[i] = val;
        }
    }
    // the CTA barrier orders every thread's stores before the flag writers; their release.sys store is cumulative over them
    // (one system-scope release per peer instead of a system fence in each of the 1024 threads)
    __syncthreads();
    const int t = threadIdx.x;
    if (t < xc.nranks && t != xc.rank) {
        unsigned *remote = xc.peer_flags[t] + xc.site * xc.nranks + xc.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(ep) : "memory");
        const unsigned *mine = xc.peer_flags[xc.rank] + xc.site * xc.nranks + t;
        unsigned cur;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(mine) : "memory");
            if (spin_expired(t0)) __trap();  // a peer died: fail instead of hanging the GPU (limit: qp_set_spin_timeout_ms)
        } while ((int)(cur - ep) < 0);
    }
    __syncthreads();
    if (t == 0) *xc.epoch = ep;
}
  // log2 of the first stride left for shared memory after had_warp128

// Single-CTA fused glue.  Every thread owns CH chunks of 4 consecutive elements; ALL global loads of the kernel are issued
// before the first dependent instruction (these kernels are pure latency: one L2 round trip instead of one per loop trip).
// The first two butterfly stages run in registers, the rest in shared memory.
template <int CH>
__global__ void __launch_bounds__(kDecThreads, 1)
fused_norm_had_kernel(__half *__restrict__ x_out, __half *h, int h_writeback, const float *acc /* may be completed by peers */,
                      const __half *__restrict__ wscale, float acc_scale, const __half *__restrict__ norm_w, float eps,
                      const __half *__restrict__ su, int n, int m, int Kf, float had_scale, int do_had,
                      float *__restrict__ zero_ptr, int zero_count, XchgDev xc) {
    extern __shared__ __align__(16) float v[];
    __shared__ float red[32];
    const int nch = n >> 2;
    uint2 hv[CH], wv[CH], nv[CH], sv[CH];
    float4 av[CH];
    // scales, norm weight and signs do not depend on the preceding kernel: fetched before the dependency wait
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        const bool ok = c < nch;
        wv[j] = (ok && acc) ? reinterpret_cast<const uint2 *>(wscale)[c] : make_uint2(0u, 0u);
        nv[j] = (ok && norm_w) ? reinterpret_cast<const uint2 *>(norm_w)[c] : make_uint2(0u, 0u);
        sv[j] = (ok && su) ? reinterpret_cast<const uint2 *>(su)[c] : make_uint2(0u, 0u);
    }
    pdl_wait();
    pdl_launch_dependents();
    if (xc.nranks > 1) {
        // clear first: a peer only pushes into the cleared buffer after it has seen this rank's flag of this site
        if (zero_ptr) zero_words4(zero_ptr, zero_count);
        zero_ptr = nullptr;
        peer_allgather(xc);  // completes `h` / `acc` (whichever is the row-sharded one) in place
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        const bool ok = c < nch;
        hv[j] = ok ? __ldcg(reinterpret_cast<const uint2 *>(h) + c) : make_uint2(0u, 0u);
        av[j] = (ok && acc) ? __ldcg(reinterpret_cast<const float4 *>(acc) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (zero_ptr) zero_words4(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
    float y[CH][4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        unpack4(hv[j], y[j]);
        if (acc) {
            float w4[4];
            unpack4(wv[j], w4);
            const float a4[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                y[j][e] = __half2float(__hadd(__float2half(y[j][e]), __float2half(scaled_acc(a4[e], w4[e], hs))));
            if (h_writeback && c < nch) reinterpret_cast<uint2 *>(h)[c] = pack4(y[j]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) ss += y[j][e] * y[j][e];
    }
    if (norm_w) {
        const float rstd = rsqrtf(block_sum(ss, red) / (float)n + eps);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            float w4[4];
            unpack4(nv[j], w4);
#pragma unroll
            for (int e = 0; e < 4; ++e)  // HF LlamaRMSNorm: weight * (x.float() * rstd).to(fp16)
                y[j][e] = __half2float(__hmul(__float2half(w4[e]), __float2half(y[j][e] * rstd)));
        }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        if (su) {
            float s4[4];
            unpack4(sv[j], s4);
#pragma unroll
            for (int e = 0; e < 4; ++e) y[j][e] *= s4[e];
        }
        if (do_had) had_warp128(y[j]);  // strides 1..64 in registers / shuffles (a warp's chunks are 128 consecutive elements)
        if (c < nch) {
            if (do_had) reinterpret_cast<float4 *>(v)[c] = make_float4(y[j][0], y[j][1], y[j][2], y[j][3]);
            else {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = y[j][e] * had_scale;
                reinterpret_cast<uint2 *>(x_out)[c] = pack4(o);
            }
        }
    }
    if (!do_had) return;
    __syncthreads();
    hadamard_smem(v, n, m, Kf, kHadLh0);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        if (c < nch) {
            const float4 f = reinterpret_cast<const float4 *>(v)[c];
            const float o[4] = {f.x * had_scale, f.y * had_scale, f.z * had_scale, f.w * had_scale};
            reinterpret_cast<uint2 *>(x_out)[c] = pack4(o);
        }
    }
}

// Fire-and-forget half of the LL exchange (the other half is polled inside the consumer GEMV's prologue, xprod.cuh): store
// this rank's slice of a gathered fp16 vector -- converted from the fp32 accumulators when `src_is_f32` -- as LL entries
// {2 x fp16, epoch, 2 x fp16, epoch} into the receive buffer of EVERY rank (own included), then bump the site's epoch.
// No fence, no flag, no wait: an 8-byte {data, epoch} store is a single NVLink write, the consumers check the epoch of every
// entry they read.  A receive buffer is rewritten one layer later at the earliest, and a peer can only get there after it
// has received this rank's contribution to the sites in between, i.e. after this rank's consumer kernel has completed.
__global__ void __launch_bounds__(kDecThreads, 1)
xchg_send_ll_kernel(const void *__restrict__ src_slice, int src_is_f32, int n, float *__restrict__ zero_ptr, int zero_count,
                    XchgDev xc) {
    pdl_wait();
    pdl_launch_dependents();
    if (zero_ptr) zero_words4(zero_ptr, zero_count);
    const unsigned ep = *xc.epoch + 1u;
    const size_t off = (size_t)xc.offset + (size_t)xc.rank * xc.slice_bytes;
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
        uint2 h;
        if (src_is_f32) {
            const float4 a = __ldcg(reinterpret_cast<const float4 *>(src_slice) + i);
            const float f[4] = {a.x, a.y, a.z, a.w};
            h = pack4(f);
        } else {
            h = __ldcg(reinterpret_cast<const uint2 *>(src_slice) + i);
        }
        for (int q = 0; q < xc.nranks; ++q) {
            const int peer = (xc.rank + q) % xc.nranks;
            uint4 *dst = reinterpret_cast<uint4 *>(xc.peer_base[peer] + off) + i;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(h.x), "r"(ep), "r"(h.y), "r"(ep) : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *xc.epoch = ep;
}

// acc = [up (I) | gate (I)] fp32 -> y = silu(gate)*up (fp16 rounding points as the reference graph) -> *su -> had -> x
template <int CH>
__global__ void __launch_bounds__(kDecThreads, 1)
silu_mul_had_kernel(__half *__restrict__ x_out, const float *__restrict__ acc, const __half *__restrict__ wscale,
                    float acc_scale, const __half *__restrict__ su, int I, int m, int Kf, float had_scale,
                    float *__restrict__ zero_ptr, int zero_count) {
    extern __shared__ __align__(16) float v[];
    pdl_wait();
    pdl_launch_dependents();
    const int nch = I >> 2;
    float4 au[CH], ag[CH];
    uint2 wu[CH], wg[CH], sv[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        const bool ok = c < nch;
        au[j] = ok ? __ldcg(reinterpret_cast<const float4 *>(acc) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        ag[j] = ok ? __ldcg(reinterpret_cast<const float4 *>(acc + I) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        wu[j] = ok ? reinterpret_cast<const uint2 *>(wscale)[c] : make_uint2(0u, 0u);
        wg[j] = ok ? reinterpret_cast<const uint2 *>(wscale + I)[c] : make_uint2(0u, 0u);
        sv[j] = (ok && su) ? reinterpret_cast<const uint2 *>(su)[c] : make_uint2(0u, 0u);
    }
    if (zero_ptr) zero_words4(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        float w_u[4], w_g[4], s4[4], y[4];
        unpack4(wu[j], w_u);
        unpack4(wg[j], w_g);
        unpack4(sv[j], s4);
        const float u4[4] = {au[j].x, au[j].y, au[j].z, au[j].w}, g4[4] = {ag[j].x, ag[j].y, ag[j].z, ag[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float up = scaled_acc(u4[e], w_u[e], hs), g = scaled_acc(g4[e], w_g[e], hs);
            const __half act = __float2half(g / (1.f + __expf(-g)));
            y[e] = __half2float(__hmul(act, __float2half(up)));
            if (su) y[e] *= s4[e];
        }
        had_warp128(y);
        if (c < nch) reinterpret_cast<float4 *>(v)[c] = make_float4(y[0], y[1], y[2], y[3]);
    }
    __syncthreads();
    hadamard_smem(v, I, m, Kf, kHadLh0);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int c = threadIdx.x + j * kDecThreads;
        if (c < nch) {
            const float4 f = reinterpret_cast<const float4 *>(v)[c];
            const float o[4] = {f.x * had_scale, f.y * had_scale, f.z * had_scale, f.w * had_scale};
            reinterpret_cast<uint2 *>(x_out)[c] = pack4(o);
        }
    }
}

// Multi-CTA variant for I = Kf * R * 512 (Llama-3.1-8B: 28 * 1 * 512; 70B: 28 * 2 * 512).  CTA b owns the 512 consecutive
// elements of block b: SiLU*mul, sign and the butterflies of stride < 512 stay inside the CTA; the block is published (fp32,
// in place over the `up` accumulators this CTA alone has read) and, after a grid-wide ticket barrier, every CTA applies the
// remaining cross-block factor (H_Kf (x) H_R on the block index) to its share of the 512 column positions.  28 SMs pull
// the 200 KB of inputs instead of one, and the 28-point factor runs 28-wide.
// Safe to spin: the grid (<= 64 CTAs of 128 threads, 2 KiB smem) is always co-resident, and dependents of a PDL launch are
// only scheduled after every CTA of this grid has started.  *sync_counter must be 0 before the first launch and is used by
// one launch at a time (launches of one stream serialise); it advances by gridDim.x per launch.
constexpr int kSiluBlk = 512, kSiluThreads = kSiluBlk / 4;
template <int KF, int R>
__global__ void __launch_bounds__(kSiluThreads)
silu_mul_had_grid_kernel(__half *__restrict__ x_out, float *acc, const __half *__restrict__ wscale, float acc_scale,
                         const __half *__restrict__ su, int I, float had_scale, float *__restrict__ zero_ptr, int zero_count,
                         unsigned *sync_counter) {
    __shared__ __align__(16) float v[kSiluBlk];
    constexpr int NB = KF * R;
    const int b = blockIdx.x, t = threadIdx.x;
    const int c = b * kSiluThreads + t;  // this thread's chunk of 4 consecutive elements
    // scales and signs do not depend on the preceding kernel: fetched before the dependency wait
    const uint2 wu = reinterpret_cast<const uint2 *>(wscale)[c];
    const uint2 wg = reinterpret_cast<const uint2 *>(wscale + I)[c];
    const uint2 sv = su ? reinterpret_cast<const uint2 *>(su)[c] : make_uint2(0u, 0u);
    pdl_wait();
    pdl_launch_dependents();
    const float4 au = __ldcg(reinterpret_cast<const float4 *>(acc) + c);
    const float4 ag = __ldcg(reinterpret_cast<const float4 *>(acc + I) + c);
    if (zero_ptr) {  // this CTA's slice of the accumulators to clear for later launches
        const int per = ((zero_count + NB - 1) / NB + 3) & ~3;
        const int lo = min(b * per, zero_count), hi = min(lo + per, zero_count);
        zero_words4(zero_ptr + lo, hi - lo);
    }
    const __half hs = __float2half(acc_scale);
    float w_u[4], w_g[4], s4[4], y[4];
    unpack4(wu, w_u);
    unpack4(wg, w_g);
    unpack4(sv, s4);
    const float u4[4] = {au.x, au.y, au.z, au.w}, g4[4] = {ag.x, ag.y, ag.z, ag.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float up = scaled_acc(u4[e], w_u[e], hs), g = scaled_acc(g4[e], w_g[e], hs);
        const __half act = __float2half(g / (1.f + __expf(-g)));
        y[e] = __half2float(__hmul(act, __float2half(up)));
        if (su) y[e] *= s4[e];
    }
    had_warp128(y);  // strides 1..64
    reinterpret_cast<float4 *>(v)[t] = make_float4(y[0], y[1], y[2], y[3]);
    __syncthreads();
    fwht_pass<2>(v, kSiluBlk, kHadLh0);  // strides 128, 256 across the 4 warps
    __syncthreads();
    reinterpret_cast<float4 *>(acc)[c] = reinterpret_cast<const float4 *>(v)[t];  // publish block b
    __syncthreads();
    if (t == 0) {
        __threadfence();  // cumulative over the CTA's stores (ordered before it by the barrier): one fence, not 128
        const unsigned old = atomicAdd(sync_counter, 1u);
        const unsigned target = old - old % NB + NB;
        unsigned cur;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(sync_counter) : "memory");
        } while ((int)(cur - target) < 0);
        __threadfence();
    }
    __syncthreads();
    // cross-block factor on columns [c_lo, c_hi) of the 512 positions: element (k, r, col) = acc[(k*R + r)*512 + col]
    const int c_lo = b * kSiluBlk / NB, c_hi = (b + 1) * kSiluBlk / NB;
    const int col = c_lo + t;
    if (col < c_hi) {
        float z[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) z[i] = __ldcg(acc + i * kSiluBlk + col);
        if (R > 1) {
#pragma unroll
            for (int k = 0; k < KF; ++k)
#pragma unroll
                for (int st = 1; st < R; st <<= 1)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if ((r & st) == 0) {
                            const float a0 = z[k * R + r], a1 = z[k * R + (r | st)];
                            z[k * R + r] = a0 + a1;
                            z[k * R + (r | st)] = a0 - a1;
                        }
        }
        if (KF == 28) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float u[14], w[14], sju[14], sjw[14];
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    u[j] = z[j * R + r];
                    w[j] = z[(14 + j) * R + r];
                }
                jacobsthal14(u, sju);
                jacobsthal14(w, sjw);
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    z[j * R + r] = (sju[j] + u[j]) + (sjw[j] - w[j]);          // (S+I)u + (S-I)w
                    z[(14 + j) * R + r] = (sju[j] - u[j]) - (sjw[j] + w[j]);   // (S-I)u - (S+I)w
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) x_out[i * kSiluBlk + col] = __float2half(z[i] * had_scale);
    }
}

// Thread-block-cluster form of the kernel above: ONE cluster of CL <= 8 CTAs, each owning NB / CL of the 512-element blocks.
// The transformed blocks stay in the owning CTA's shared memory; after a hardware cluster barrier every CTA reads the NB
// values of its column positions straight out of its peers' shared memory (DSMEM) -- no global exchange buffer, no atomic
// ticket, no L2 round trip between the two phases.  `acc` is not modified.
template <int KF, int R, int CL>
__global__ void __launch_bounds__((KF * R / CL) * kSiluThreads)
silu_mul_had_cluster_kernel(__half *__restrict__ x_out, const float *__restrict__ acc, const __half *__restrict__ wscale,
                            float acc_scale, const __half *__restrict__ su, int I, float had_scale,
                            float *__restrict__ zero_ptr, int zero_count) {
    constexpr int NB = KF * R, BPC = NB / CL, T = BPC * kSiluThreads, NE = BPC * kSiluBlk;
    static_assert(NB % CL == 0 && CL <= 8, "blocks must split evenly over a portable cluster");
    __shared__ __align__(16) float v[NE];
    const int b = blockIdx.x, t = threadIdx.x;  // the grid is one cluster: blockIdx.x = rank in the cluster
    const int c = b * T + t;                    // this thread's chunk of 4 consecutive elements
    const uint2 wu = reinterpret_cast<const uint2 *>(wscale)[c];
    const uint2 wg = reinterpret_cast<const uint2 *>(wscale + I)[c];
    const uint2 sv = su ? reinterpret_cast<const uint2 *>(su)[c] : make_uint2(0u, 0u);
    pdl_wait();
    pdl_launch_dependents();
    const float4 au = __ldcg(reinterpret_cast<const float4 *>(acc) + c);
    const float4 ag = __ldcg(reinterpret_cast<const float4 *>(acc + I) + c);
    if (zero_ptr) {
        const int per = ((zero_count + CL - 1) / CL + 3) & ~3;
        const int lo = min(b * per, zero_count), hi = min(lo + per, zero_count);
        zero_words4(zero_ptr + lo, hi - lo);
    }
    const __half hs = __float2half(acc_scale);
    float w_u[4], w_g[4], s4[4], y[4];
    unpack4(wu, w_u);
    unpack4(wg, w_g);
    unpack4(sv, s4);
    const float u4[4] = {au.x, au.y, au.z, au.w}, g4[4] = {ag.x, ag.y, ag.z, ag.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float up = scaled_acc(u4[e], w_u[e], hs), g = scaled_acc(g4[e], w_g[e], hs);
        const __half act = __float2half(g / (1.f + __expf(-g)));
        y[e] = __half2float(__hmul(act, __float2half(up)));
        if (su) y[e] *= s4[e];
    }
    had_warp128(y);  // strides 1..64
    reinterpret_cast<float4 *>(v)[t] = make_float4(y[0], y[1], y[2], y[3]);
    __syncthreads();
    fwht_pass<2>(v, NE, kHadLh0);  // strides 128, 256 inside every 512-block of this CTA
    cluster_sync_all();            // all blocks of all CTAs are final (also orders this CTA's own shared-memory writes)
    // cross-block factor on this CTA's share of the 512 column positions; block i lives in CTA i / BPC
    const int c_lo = b * kSiluBlk / CL, c_hi = (b + 1) * kSiluBlk / CL;
    for (int col = c_lo + t; col < c_hi; col += T) {
        float z[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) z[i] = ld_dsmem_f32(v + (i % BPC) * kSiluBlk + col, (unsigned)(i / BPC));
        if (R > 1) {
#pragma unroll
            for (int k = 0; k < KF; ++k)
#pragma unroll
                for (int st = 1; st < R; st <<= 1)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if ((r & st) == 0) {
                            const float a0 = z[k * R + r], a1 = z[k * R + (r | st)];
                            z[k * R + r] = a0 + a1;
                            z[k * R + (r | st)] = a0 - a1;
                        }
        }
        if (KF == 28) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float u[14], w[14], sju[14], sjw[14];
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    u[j] = z[j * R + r];
                    w[j] = z[(14 + j) * R + r];
                }
                jacobsthal14(u, sju);
                jacobsthal14(w, sjw);
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    z[j * R + r] = (sju[j] + u[j]) + (sjw[j] - w[j]);
                    z[(14 + j) * R + r] = (sju[j] - u[j]) - (sjw[j] + w[j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) x_out[i * kSiluBlk + col] = __float2half(z[i] * had_scale);
    }
    cluster_sync_all();  // nobody leaves while a peer may still read its blocks
}

// Row-sharded form of the kernel above: rank r holds the up / gate accumulators of rows [r, r+1) * I / nranks only.  It runs
// one CTA per LOCAL 512-element block: SiLU*mul, sign, butterflies of stride < 512, then the block (fp32, 2 KB) is stored
// into the exchange buffer `z` (I floats inside the peer-mapped region) of EVERY rank over NVLink and a per-(site, source)
// counter on each peer is bumped (red.release.sys).  Once all I / 512 blocks have arrived (own blocks: local ticket; the
// others: the peers' counters), every rank applies the cross-block factor to its share of the column positions and writes
// the complete x.  Replaces SiLU*mul epilogue + all-gather + single-CTA Hadamard (3 launches, one NCCL call).
// Counters only ever grow (launch k of a site waits for k * blocks-per-rank), so graph replays need no reset; the
// accumulator-clearing duty is done by each CTA before it publishes (a peer that has seen all counters of this site knows
// this rank's buffers are cleared).
template <int KF, int R>
__global__ void __launch_bounds__(kSiluThreads)
silu_mul_had_grid_xchg_kernel(__half *__restrict__ x_out, const float *__restrict__ acc, const __half *__restrict__ wscale,
                              float acc_scale, const __half *__restrict__ su, int I, float had_scale,
                              float *__restrict__ zero_ptr, int zero_count, unsigned *sync_counter, XchgDev xc) {
    __shared__ __align__(16) float v[kSiluBlk];
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NB = KF * R;
    const int nbl = gridDim.x;                 // local blocks = NB / nranks
    const int Il = nbl * kSiluBlk;             // local rows
    const int b = blockIdx.x, t = threadIdx.x;
    const int gb = xc.rank * nbl + b;          // global block
    const int c = b * kSiluThreads + t;        // local chunk of 4 consecutive elements
    const unsigned ep = *xc.epoch + 1u;        // every CTA reads it before any CTA can get past the barrier below
    const float4 au = __ldcg(reinterpret_cast<const float4 *>(acc) + c);
    const float4 ag = __ldcg(reinterpret_cast<const float4 *>(acc + Il) + c);
    const uint2 wu = reinterpret_cast<const uint2 *>(wscale)[c];
    const uint2 wg = reinterpret_cast<const uint2 *>(wscale + Il)[c];
    const uint2 sv = su ? reinterpret_cast<const uint2 *>(su)[gb * kSiluThreads + t] : make_uint2(0u, 0u);
    if (zero_ptr) {
        const int per = ((zero_count + nbl - 1) / nbl + 3) & ~3;
        const int lo = min(b * per, zero_count), hi = min(lo + per, zero_count);
        zero_words4(zero_ptr + lo, hi - lo);
    }
    const __half hs = __float2half(acc_scale);
    float w_u[4], w_g[4], s4[4], y[4];
    unpack4(wu, w_u);
    unpack4(wg, w_g);
    unpack4(sv, s4);
    const float u4[4] = {au.x, au.y, au.z, au.w}, g4[4] = {ag.x, ag.y, ag.z, ag.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float up = scaled_acc(u4[e], w_u[e], hs), g = scaled_acc(g4[e], w_g[e], hs);
        const __half act = __float2half(g / (1.f + __expf(-g)));
        y[e] = __half2float(__hmul(act, __float2half(up)));
        if (su) y[e] *= s4[e];
    }
    had_warp128(y);
    reinterpret_cast<float4 *>(v)[t] = make_float4(y[0], y[1], y[2], y[3]);
    __syncthreads();
    fwht_pass<2>(v, kSiluBlk, kHadLh0);
    __syncthreads();
    // publish block gb everywhere (own copy included)
    const float4 zb = reinterpret_cast<const float4 *>(v)[t];
    const size_t zoff = (size_t)xc.offset + ((size_t)gb * kSiluBlk + (size_t)t * 4) * sizeof(float);
    for (int q = 0; q < xc.nranks; ++q) {
        const int peer = (xc.rank + q) % xc.nranks;
        *reinterpret_cast<float4 *>(xc.peer_base[peer] + zoff) = zb;
    }
    __syncthreads();  // orders the block's stores before the (cumulative) release.sys increments below
    if (t < xc.nranks) {
        // counter [site][source = this rank] on rank t (t == rank: the local count of this rank's own blocks)
        unsigned *remote = xc.peer_flags[t] + xc.site * xc.nranks + xc.rank;
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(remote), "r"(1u) : "memory");
        const unsigned target = ep * (unsigned)nbl;
        const unsigned *mine = xc.peer_flags[xc.rank] + xc.site * xc.nranks + t;
        unsigned cur;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(mine) : "memory");
            if (spin_expired(t0)) __trap();  // a peer died: fail instead of hanging the GPU
        } while ((int)(cur - target) < 0);
        __threadfence_system();
    }
    __syncthreads();
    if (b == 0 && t == 0) *xc.epoch = ep;
    // cross-block factor on this CTA's share of the 512 column positions, all NB blocks from the local exchange buffer
    const float *z = reinterpret_cast<const float *>(xc.peer_base[xc.rank] + xc.offset);
    const int c_lo = b * kSiluBlk / nbl, c_hi = (b + 1) * kSiluBlk / nbl;
    for (int col = c_lo + t; col < c_hi; col += kSiluThreads) {
        float zz[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) zz[i] = __ldcg(z + i * kSiluBlk + col);
        if (R > 1) {
#pragma unroll
            for (int k = 0; k < KF; ++k)
#pragma unroll
                for (int st = 1; st < R; st <<= 1)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if ((r & st) == 0) {
                            const float a0 = zz[k * R + r], a1 = zz[k * R + (r | st)];
                            zz[k * R + r] = a0 + a1;
                            zz[k * R + (r | st)] = a0 - a1;
                        }
        }
        if (KF == 28) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float u[14], w[14], sju[14], sjw[14];
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    u[j] = zz[j * R + r];
                    w[j] = zz[(14 + j) * R + r];
                }
                jacobsthal14(u, sju);
                jacobsthal14(w, sjw);
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    zz[j * R + r] = (sju[j] + u[j]) + (sjw[j] - w[j]);
                    zz[(14 + j) * R + r] = (sju[j] - u[j]) - (sjw[j] + w[j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) x_out[i * kSiluBlk + col] = __float2half(zz[i] * had_scale);
    }
}

// Attention of one new token.  acc_qkv: fp32 [q (H*D) | k (Hkv*D) | v (Hkv*D)] raw GEMV sums; wscale same layout.
// RoPE (HF rotate_half convention, model/llama.py apply_rotary_pos_emb) with inv_freq table (D/2 floats, llama3 scaling
// already applied by the host).  KV cache: fp16 [max_seq][Hkv][D] per layer.  pos read from device memory.
// D = 128 (rope_attention_split_kernel): grid (heads, splits); a CTA takes kAttnChunk = 128 positions of one query head, a
// warp covers a cached position with one 8-byte load per lane and has ALL of its 16 K rows and 16 V rows in flight at once,
// issued before the dependency wait: a split costs one memory round trip whatever the context length.  Splits beyond the
// current position exit at once; with one active split (pos < 128) the result is written directly, otherwise the splits leave
// (max, sum, unnormalised output) partials and the last one to finish (a ticket per head) combines them.  Measured at
// position 2048: 4.45 ms / token with the single-CTA loop of round 1 (33 dependent round trips per layer) -- see DESIGN.md.
// Other head sizes: rope_attention_kernel, one CTA per head looping over the context.
constexpr int kAttnThreads = 256;
constexpr int kAttnChunk = 128;                                  // positions per split
constexpr int kAttnRows = kAttnChunk / (kAttnThreads / 32);      // cached rows per warp and split
constexpr int kAttnPart = 128 + 4;                               // floats per partial: o[128], max, sum, pad

__global__ void __launch_bounds__(kAttnThreads)
rope_attention_split_kernel(__half *__restrict__ attn_out, const float *__restrict__ acc_qkv, const __half *__restrict__ wscale,
                            float acc_scale, const float *__restrict__ inv_freq, __half *__restrict__ kcache,
                            __half *__restrict__ vcache, const int *__restrict__ pos_ptr, int H, int Hkv, int max_seq,
                            int qvk_order, float *__restrict__ zero_ptr, int zero_count, float *__restrict__ part_buf,
                            unsigned *__restrict__ tickets, int S) {
    constexpr int D = 128, C = kAttnChunk, nw = kAttnThreads / 32;
    __shared__ __align__(16) float q[D], kn[D], vn[D], part[nw][D], sc[C + 4];
    __shared__ float red[32];
    __shared__ unsigned s_ticket;
    const int head = blockIdx.x, split = blockIdx.y, kvh = head / (H / Hkv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a position past the cache capacity would write out of bounds: clamp (DecodeRunner.step refuses to go that far)
    const int pos = min(__ldcg(pos_ptr), max_seq - 1);
    const int nact = pos / C + 1;  // splits that hold a position <= pos
    if (split >= nact) return;
    const bool owner = split == nact - 1;                 // holds position pos itself (taken from shared memory, not the cache)
    const int t_lo = split * C, t_hi = min(pos, t_lo + C);  // cached rows [t_lo, t_hi) of this split
    // cached rows were written by earlier decode steps, not by the preceding kernel: fetched before the dependency wait
    uint2 kpre[kAttnRows], vpre[kAttnRows];
#pragma unroll
    for (int u = 0; u < kAttnRows; ++u) {
        const int t = t_lo + warp + u * nw;
        const size_t row = ((size_t)t * Hkv + kvh) * D;
        kpre[u] = (t < t_hi) ? __ldcg(reinterpret_cast<const uint2 *>(kcache + row) + lane) : make_uint2(0u, 0u);
        vpre[u] = (t < t_hi) ? __ldcg(reinterpret_cast<const uint2 *>(vcache + row) + lane) : make_uint2(0u, 0u);
    }
    // likewise independent of the preceding kernel: the per-row scales and the rotary angle of this position
    const int d = threadIdx.x;
    constexpr int half = D / 2;
    const int pd = d < half ? d + half : d - half;
    // accumulator / Wscale order: q | k | v, or q | v | k for a merge_qv layer (lib/linear/incoherent_linear.py:211-213)
    const int iq = head * D, ik = (qvk_order ? H + Hkv : H) * D + kvh * D, iv = (qvk_order ? H : H + Hkv) * D + kvh * D;
    float wq = 0.f, wqp = 0.f, wk = 0.f, wkp = 0.f, wv = 0.f, c16 = 0.f, s16 = 0.f;
    if (d < D) {
        wq = __half2float(wscale[iq + d]), wqp = __half2float(wscale[iq + pd]), wk = __half2float(wscale[ik + d]);
        wkp = __half2float(wscale[ik + pd]), wv = __half2float(wscale[iv + d]);
        const float fr = inv_freq[d % half];
        double snd, csd;  // precise range reduction (the build uses --use_fast_math)
        sincos((double)pos * (double)fr, &snd, &csd);
        // fp16 rounding of cos/sin as the fp16 reference graph does
        c16 = __half2float(__float2half((float)csd)), s16 = __half2float(__float2half((float)snd));
    }
    pdl_wait();
    pdl_launch_dependents();
    if (zero_ptr && blockIdx.x == 0 && split == 0) zero_words4(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
    if (d < D) {
        const float aq = __ldcg(acc_qkv + iq + d), aqp = __ldcg(acc_qkv + iq + pd), ak = __ldcg(acc_qkv + ik + d),
                    akp = __ldcg(acc_qkv + ik + pd), av = __ldcg(acc_qkv + iv + d);
        const float sgn = d < half ? -1.f : 1.f;
        const float qa = scaled_acc(aq, wq, hs), qb = scaled_acc(aqp, wqp, hs);
        const float ka = scaled_acc(ak, wk, hs), kb = scaled_acc(akp, wkp, hs);
        q[d] = __half2float(__float2half(qa * c16 + sgn * qb * s16));
        const __half kr = __float2half(ka * c16 + sgn * kb * s16);
        const __half vv = __float2half(scaled_acc(av, wv, hs));
        kn[d] = __half2float(kr);
        vn[d] = __half2float(vv);
        if (owner && head % (H / Hkv) == 0) {  // one CTA of the group appends to the cache
            kcache[((size_t)pos * Hkv + kvh) * D + d] = kr;
            vcache[((size_t)pos * Hkv + kvh) * D + d] = vv;
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)D);
    const float4 q4 = reinterpret_cast<const float4 *>(q)[lane];
#pragma unroll
    for (int u = 0; u < kAttnRows; ++u) {
        const int t = t_lo + warp + u * nw;
        float k4[4];
        unpack4(kpre[u], k4);
        float sdot = q4.x * k4[0] + q4.y * k4[1] + q4.z * k4[2] + q4.w * k4[3];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
        if (lane == 0 && t < t_hi) sc[t - t_lo] = sdot * scale;
    }
    const int n_loc = t_hi - t_lo + (owner ? 1 : 0);  // scores of this split; the owner's last one is the new token
    if (owner && warp == 0) {
        float sdot = 0.f;
        for (int i = lane; i < D; i += 32) sdot += q[i] * kn[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
        if (lane == 0) sc[n_loc - 1] = sdot * scale;
    }
    __syncthreads();
    float mx = threadIdx.x < n_loc ? sc[threadIdx.x] : -INFINITY;  // n_loc <= C + 1 <= blockDim.x
    mx = block_max(mx, red);
    float sum = 0.f;
    if (threadIdx.x < n_loc) {
        const float e = __expf(sc[threadIdx.x] - mx);
        sc[threadIdx.x] = e;
        sum = e;
    }
    sum = block_sum(sum, red);
    __syncthreads();
    float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < kAttnRows; ++u) {
        const int t = t_lo + warp + u * nw;
        if (t < t_hi) {
            float v4[4];
            unpack4(vpre[u], v4);
            const float w = sc[t - t_lo];
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] += w * v4[e];
        }
    }
    reinterpret_cast<float4 *>(part[warp])[lane] = make_float4(o4[0], o4[1], o4[2], o4[3]);
    __syncthreads();
    float o = 0.f;
    if (threadIdx.x < D) {
        o = owner ? sc[n_loc - 1] * vn[threadIdx.x] : 0.f;
#pragma unroll
        for (int w = 0; w < nw; ++w) o += part[w][threadIdx.x];
    }
    if (nact == 1) {
        if (threadIdx.x < D) attn_out[head * D + threadIdx.x] = __float2half(o / sum);
        return;
    }
    // several splits: leave the partial, the last split of this head to finish combines
    float *pb = part_buf + ((size_t)head * S + split) * kAttnPart;
    if (threadIdx.x < D) pb[threadIdx.x] = o;
    if (threadIdx.x == 0) pb[D] = mx, pb[D + 1] = sum;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(tickets + head, 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(nact - 1)) return;
    __threadfence();
    if (threadIdx.x < D) {
        const float *p0 = part_buf + (size_t)head * S * kAttnPart;
        float M = -INFINITY;
        for (int i = 0; i < nact; ++i) M = fmaxf(M, __ldcg(p0 + i * kAttnPart + D));
        float Lsum = 0.f, O = 0.f;
        for (int i = 0; i < nact; ++i) {
            const float w = __expf(__ldcg(p0 + i * kAttnPart + D) - M);
            Lsum += w * __ldcg(p0 + i * kAttnPart + D + 1);
            O += w * __ldcg(p0 + i * kAttnPart + threadIdx.x);
        }
        attn_out[head * D + threadIdx.x] = __float2half(O / Lsum);
    }
    if (threadIdx.x == 0) tickets[head] = 0u;  // ready for the next launch (stream-ordered behind this one)
}

__global__ void __launch_bounds__(kAttnThreads, 1)
rope_attention_kernel(__half *__restrict__ attn_out, const float *__restrict__ acc_qkv, const __half *__restrict__ wscale,
                      float acc_scale, const float *__restrict__ inv_freq, __half *__restrict__ kcache,
                      __half *__restrict__ vcache, const int *__restrict__ pos_ptr, int H, int Hkv, int D, int max_seq,
                      int qvk_order, float *__restrict__ zero_ptr, int zero_count) {
    extern __shared__ __align__(16) float sm[];  // q[D] | knew[D] | vnew[D] | part[8][D] | scores[max_seq]   (head sizes other than 128)
    __shared__ float red[32];
    float *q = sm, *kn = sm + D, *vn = sm + 2 * D, *part = sm + 3 * D, *sc = sm + 11 * D;
    const int head = blockIdx.x, kvh = head / (H / Hkv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kAttnThreads / 32;
    // a position past the cache capacity would write out of bounds (cache rows and the score array hold max_seq entries):
    // clamp to the last row -- the host API refuses to step that far (DecodeRunner.step), this only keeps memory safe
    const int pos = min(__ldcg(pos_ptr), max_seq - 1);
    // likewise independent of the preceding kernel: the per-row scales and the rotary angle of this position
    const int d = threadIdx.x;
    const int half = D / 2;
    const int pd = d < half ? d + half : d - half;
    // accumulator / Wscale order: q | k | v, or q | v | k for a merge_qv layer (lib/linear/incoherent_linear.py:211-213)
    const int iq = head * D, ik = (qvk_order ? H + Hkv : H) * D + kvh * D, iv = (qvk_order ? H : H + Hkv) * D + kvh * D;
    float wq = 0.f, wqp = 0.f, wk = 0.f, wkp = 0.f, wv = 0.f, c16 = 0.f, s16 = 0.f;
    if (d < D) {
        wq = __half2float(wscale[iq + d]), wqp = __half2float(wscale[iq + pd]), wk = __half2float(wscale[ik + d]);
        wkp = __half2float(wscale[ik + pd]), wv = __half2float(wscale[iv + d]);
        const float fr = inv_freq[d % half];
        double snd, csd;  // precise range reduction (the build uses --use_fast_math)
        sincos((double)pos * (double)fr, &snd, &csd);
        // fp16 rounding of cos/sin as the fp16 reference graph does
        c16 = __half2float(__float2half((float)csd)), s16 = __half2float(__float2half((float)snd));
    }
    pdl_wait();
    pdl_launch_dependents();
    if (zero_ptr && blockIdx.x == 0) zero_words4(zero_ptr, zero_count);
    const __half hs = __float2half(acc_scale);
    if (d < D) {
        // all loads first
        const float aq = __ldcg(acc_qkv + iq + d), aqp = __ldcg(acc_qkv + iq + pd), ak = __ldcg(acc_qkv + ik + d),
                    akp = __ldcg(acc_qkv + ik + pd), av = __ldcg(acc_qkv + iv + d);
        const float sgn = d < half ? -1.f : 1.f;
        const float qa = scaled_acc(aq, wq, hs), qb = scaled_acc(aqp, wqp, hs);
        const float ka = scaled_acc(ak, wk, hs), kb = scaled_acc(akp, wkp, hs);
        q[d] = __half2float(__float2half(qa * c16 + sgn * qb * s16));
        const __half kr = __float2half(ka * c16 + sgn * kb * s16);
        const __half vv = __float2half(scaled_acc(av, wv, hs));
        kn[d] = __half2float(kr);
        vn[d] = __half2float(vv);
        if (head % (H / Hkv) == 0) {  // one CTA of the group appends to the cache
            kcache[((size_t)pos * Hkv + kvh) * D + d] = kr;
            vcache[((size_t)pos * Hkv + kvh) * D + d] = vv;
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)D);
    for (int t = warp; t < pos; t += nw) {
        const __half *kr = kcache + ((size_t)t * Hkv + kvh) * D;
        float s = 0.f;
        for (int i = lane; i < D; i += 32) s += q[i] * __half2float(__ldcg(kr + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sc[t] = s * scale;
    }
    if (warp == 0) {  // the new token, from shared memory (not the cache: no cross-CTA race)
        float s = 0.f;
        for (int i = lane; i < D; i += 32) s += q[i] * kn[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sc[pos] = s * scale;
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
    // generic head size: two halves of the positions per dim, 8 loads in flight
    const int dd = threadIdx.x % D, part_id = threadIdx.x / D, nparts = (kAttnThreads / D) >= 2 ? 2 : 1;
    if (part_id < nparts) {
        float o = 0.f;
        for (int t0 = part_id; t0 < pos; t0 += nparts * 8) {
            __half vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * nparts;
                vv[u] = (t < pos) ? __ldcg(vcache + ((size_t)t * Hkv + kvh) * D + dd) : __float2half(0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * nparts;
                if (t < pos) o += sc[t] * __half2float(vv[u]);
            }
        }
        if (part_id == 0) o += sc[pos] * vn[dd];
        part[part_id * D + dd] = o;
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float o = part[threadIdx.x];
        if (nparts > 1) o += part[D + threadIdx.x];
        attn_out[head * D + threadIdx.x] = __float2half(o / sum);
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
    for (int i = threadIdx.x; i < kq; i += blockDim.x) xs[i] = __ldcg(reinterpret_cast<const uint4 *>(x) + i);
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
        const float v = __ldcg(logits + i);
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
    const size_t row = (size_t)__ldcg(token);
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
    QP_CHECK_ARG(n % 128 == 0, "Hadamard size %d must be a multiple of 128 for the fused kernels", n);
    Kf = 1;
    m = n;
    if ((n & (n - 1)) != 0) {
        QP_CHECK_ARG(n % 28 == 0 && (((n / 28) & (n / 28 - 1)) == 0), "Hadamard size %d is neither 2^k nor 28*2^k", n);
        Kf = 28;
        m = n / 28;
    }
    QP_CHECK_ARG(m >= 128, "Hadamard block %d < 128 is not supported by the fused decode kernels (use qp_hadamard)", m);
    return QP_OK;
}

static int fused_norm_had_impl(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                                 float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n,
                                 float had_scale, int do_had, float *zero_ptr, int zero_count, const XchgDev &xc, void *stream) {
    QP_CHECK_ARG(x_out_f16 && h_f16, "NULL pointer argument");
    QP_CHECK_ARG(!acc || wscale_f16, "acc given without wscale");
    int m, Kf;
    int rc = had_dims(n, m, Kf);
    if (rc != QP_OK && do_had) return rc;
    const size_t smem = (size_t)n * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "n = %d too large", n);
    QP_CHECK_ARG(n % 4 == 0 && n <= 8 * 4 * kDecThreads, "n = %d unsupported (needs n %% 4 == 0, n <= 32768)", n);
    const int ch = (n / 4 + kDecThreads - 1) / kDecThreads;
    void (*kern)(__half *, __half *, int, const float *, const __half *, float, const __half *, float, const __half *, int,
                 int, int, float, int, float *, int, XchgDev) =
        ch <= 1 ? fused_norm_had_kernel<1> : ch <= 2 ? fused_norm_had_kernel<2> : ch <= 4 ? fused_norm_had_kernel<4>
                                                                                       : fused_norm_had_kernel<8>;
    static DeviceOnce configured[4];
    const int ki = ch <= 1 ? 0 : ch <= 2 ? 1 : ch <= 4 ? 2 : 3;
    if (configured[ki].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
    }
    QP_CUDA(launch_pdl(kern, dim3(1), dim3(kDecThreads), smem, (cudaStream_t)stream, (__half *)x_out_f16, (__half *)h_f16,
                       h_writeback, acc, (const __half *)wscale_f16, acc_scale, (const __half *)norm_w_f16, eps,
                       (const __half *)su_f16, n, m, Kf, had_scale, do_had, zero_ptr, zero_count, xc));
    return check_launch("fused_norm_had");
}

extern "C" int qp_fused_norm_had(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                                 float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n,
                                 float had_scale, int do_had, float *zero_ptr, int zero_count, void *stream) {
    XchgDev none = {};
    return fused_norm_had_impl(x_out_f16, h_f16, h_writeback, acc, wscale_f16, acc_scale, norm_w_f16, eps, su_f16, n, had_scale,
                               do_had, zero_ptr, zero_count, none, stream);
}

extern "C" int qp_fused_norm_had_xchg(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                                      float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n,
                                      float had_scale, int do_had, float *zero_ptr, int zero_count, const qp_xchg *xc,
                                      void *stream) {
    QP_CHECK_ARG(xc && xc->peer_base && xc->peer_flags && xc->epoch, "NULL exchange descriptor");
    QP_CHECK_ARG(xc->nranks >= 1 && xc->nranks <= 32 && xc->rank >= 0 && xc->rank < xc->nranks, "bad rank %d of %d", xc->rank, xc->nranks);
    QP_CHECK_ARG(xc->slice_bytes > 0 && xc->slice_bytes % 16 == 0 && xc->offset % 16 == 0, "exchange slices must be 16-byte multiples");
    XchgDev d;
    d.peer_base = (unsigned char *const *)xc->peer_base;
    d.peer_flags = (unsigned *const *)xc->peer_flags;
    d.epoch = xc->epoch + xc->site;
    d.offset = xc->offset;
    d.slice_bytes = xc->slice_bytes;
    d.rank = xc->rank;
    d.nranks = xc->nranks;
    d.site = xc->site;
    return fused_norm_had_impl(x_out_f16, h_f16, h_writeback, acc, wscale_f16, acc_scale, norm_w_f16, eps, su_f16, n, had_scale,
                               do_had, zero_ptr, zero_count, d, stream);
}

static int make_xchg_dev(XchgDev &d, const qp_xchg *xc) {
    QP_CHECK_ARG(xc && xc->peer_base && xc->peer_flags && xc->epoch, "NULL exchange descriptor");
    QP_CHECK_ARG(xc->nranks >= 1 && xc->nranks <= 32 && xc->rank >= 0 && xc->rank < xc->nranks, "bad rank %d of %d", xc->rank, xc->nranks);
    QP_CHECK_ARG(xc->slice_bytes > 0 && xc->slice_bytes % 16 == 0 && xc->offset % 16 == 0, "exchange slices must be 16-byte multiples");
    d.peer_base = (unsigned char *const *)xc->peer_base;
    d.peer_flags = (unsigned *const *)xc->peer_flags;
    d.epoch = xc->epoch + xc->site;
    d.offset = xc->offset;
    d.slice_bytes = xc->slice_bytes;
    d.rank = xc->rank;
    d.nranks = xc->nranks;
    d.site = xc->site;
    return QP_OK;
}

extern "C" int qp_xchg_send_ll(const void *src_slice, int src_is_f32, int n, float *zero_ptr, int zero_count, const qp_xchg *xc,
                               void *stream) {
    QP_CHECK_ARG(src_slice && xc, "NULL pointer argument");
    QP_CHECK_ARG(n > 0 && n % 4 == 0, "slice of %d elements is not a multiple of 4", n);
    QP_CHECK_ARG(xc->slice_bytes == n * 4, "LL slice is 4 bytes per element: slice_bytes = %d for n = %d", xc->slice_bytes, n);
    XchgDev d;
    int rc = make_xchg_dev(d, xc);
    if (rc != QP_OK) return rc;
    QP_CHECK_ARG(((uintptr_t)src_slice) % 16 == 0, "src_slice must be 16-byte aligned");
    QP_CUDA(launch_pdl(xchg_send_ll_kernel, dim3(1), dim3(kDecThreads), 0, (cudaStream_t)stream, src_slice, src_is_f32, n, zero_ptr,
                       zero_count, d));
    return check_launch("xchg_send_ll");
}


// ---- exchange region management (CUDA IPC) -------------------------------------------------------------------------------
extern "C" int qp_peer_alloc(void **ptr, size_t bytes) {
    QP_CHECK_ARG(ptr && bytes > 0, "bad arguments");
    QP_CUDA(cudaMalloc(ptr, bytes));
    QP_CUDA(cudaMemset(*ptr, 0, bytes));
    QP_CUDA(cudaDeviceSynchronize());
    return QP_OK;
}
extern "C" int qp_peer_free(void *ptr) {
    QP_CUDA(cudaFree(ptr));
    return QP_OK;
}
extern "C" int qp_peer_export(void *ptr, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    QP_CHECK_ARG(ptr && handle64, "NULL pointer argument");
    QP_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), ptr));
    return QP_OK;
}
extern "C" int qp_peer_import(const void *handle64, void **ptr) {
    QP_CHECK_ARG(ptr && handle64, "NULL pointer argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    QP_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QP_OK;
}
extern "C" int qp_peer_close(void *ptr) {
    QP_CUDA(cudaIpcCloseMemHandle(ptr));
    return QP_OK;
}

extern "C" int qp_silu_mul_had(void *x_out_f16, const float *acc, const void *wscale_f16, float acc_scale,
                               const void *su_f16, int I, float had_scale, float *zero_ptr, int zero_count, void *stream) {
    QP_CHECK_ARG(x_out_f16 && acc && wscale_f16, "NULL pointer argument");
    int m, Kf;
    int rc = had_dims(I, m, Kf);
    if (rc != QP_OK) return rc;
    const size_t smem = (size_t)I * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "I = %d too large", I);
    QP_CHECK_ARG(I % 4 == 0 && I <= 8 * 4 * kDecThreads, "I = %d unsupported", I);
    const int ch = (I / 4 + kDecThreads - 1) / kDecThreads;
    void (*kern)(__half *, const float *, const __half *, float, const __half *, int, int, int, float, float *, int) =
        ch <= 1 ? silu_mul_had_kernel<1> : ch <= 2 ? silu_mul_had_kernel<2> : ch <= 4 ? silu_mul_had_kernel<4>
                                                                                     : silu_mul_had_kernel<8>;
    static DeviceOnce configured[4];
    const int ki = ch <= 1 ? 0 : ch <= 2 ? 1 : ch <= 4 ? 2 : 3;
    if (configured[ki].first()) {
        QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
    }
    QP_CUDA(launch_pdl(kern, dim3(1), dim3(kDecThreads), smem, (cudaStream_t)stream, (__half *)x_out_f16, acc,
                       (const __half *)wscale_f16, acc_scale, (const __half *)su_f16, I, m, Kf, had_scale, zero_ptr,
                       zero_count));
    return check_launch("silu_mul_had");
}

extern "C" int qp_silu_mul_had_grid(void *x_out_f16, float *acc, const void *wscale_f16, float acc_scale, const void *su_f16,
                                    int I, float had_scale, float *zero_ptr, int zero_count, unsigned *sync_counter,
                                    void *stream) {
    QP_CHECK_ARG(x_out_f16 && acc && wscale_f16 && sync_counter, "NULL pointer argument");
    int m, Kf;
    int rc = had_dims(I, m, Kf);
    if (rc != QP_OK) return rc;
    QP_CHECK_ARG(m % kSiluBlk == 0, "I = %d: the multi-CTA kernel needs a power-of-two factor >= 512 (use qp_silu_mul_had)", I);
    const int R = m / kSiluBlk;
    void (*kern)(__half *, float *, const __half *, float, const __half *, int, float, float *, int, unsigned *) = nullptr;
    if (Kf == 28 && R == 1) kern = silu_mul_had_grid_kernel<28, 1>;
    else if (Kf == 28 && R == 2) kern = silu_mul_had_grid_kernel<28, 2>;
    else if (Kf == 1 && R == 8) kern = silu_mul_had_grid_kernel<1, 8>;
    else if (Kf == 1 && R == 16) kern = silu_mul_had_grid_kernel<1, 16>;
    else if (Kf == 1 && R == 32) kern = silu_mul_had_grid_kernel<1, 32>;
    QP_CHECK_ARG(kern != nullptr, "I = %d = %d * %d * 512 is not instantiated for the multi-CTA kernel (use qp_silu_mul_had)", I, Kf, R);
    QP_CUDA(launch_pdl(kern, dim3(Kf * R), dim3(kSiluThreads), 0, (cudaStream_t)stream, (__half *)x_out_f16, acc,
                       (const __half *)wscale_f16, acc_scale, (const __half *)su_f16, I, had_scale, zero_ptr, zero_count,
                       sync_counter));
    return check_launch("silu_mul_had_grid");
}

extern "C" int qp_silu_mul_had_cluster(void *x_out_f16, const float *acc, const void *wscale_f16, float acc_scale,
                                       const void *su_f16, int I, float had_scale, float *zero_ptr, int zero_count,
                                       void *stream) {
    QP_CHECK_ARG(x_out_f16 && acc && wscale_f16, "NULL pointer argument");
    int m, Kf;
    int rc = had_dims(I, m, Kf);
    if (rc != QP_OK) return rc;
    QP_CHECK_ARG(m % kSiluBlk == 0, "I = %d: the cluster kernel needs a power-of-two factor >= 512 (use qp_silu_mul_had)", I);
    const int R = m / kSiluBlk;
#define QP_LAUNCH_CLUSTER(KF_, R_, CL_)                                                                                   \
    do {                                                                                                                  \
        auto kern = silu_mul_had_cluster_kernel<KF_, R_, CL_>;                                                            \
        static DeviceOnce configured;                                                                                   \
        if (configured.first()) {                                                                                                \
            QP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));                       \
        }                                                                                                                 \
        QP_CUDA(launch_pdl_cluster(kern, dim3(CL_), dim3((KF_ * R_ / CL_) * kSiluThreads), CL_, 0, (cudaStream_t)stream,  \
                                   (__half *)x_out_f16, acc, (const __half *)wscale_f16, acc_scale,                       \
                                   (const __half *)su_f16, I, had_scale, zero_ptr, zero_count));                          \
        return check_launch("silu_mul_had_cluster");                                                                      \
    } while (0)
    if (Kf == 28 && R == 1) QP_LAUNCH_CLUSTER(28, 1, 7);
    if (Kf == 28 && R == 2) QP_LAUNCH_CLUSTER(28, 2, 8);
    if (Kf == 1 && R == 8) QP_LAUNCH_CLUSTER(1, 8, 8);
    if (Kf == 1 && R == 16) QP_LAUNCH_CLUSTER(1, 16, 8);
    if (Kf == 1 && R == 32) QP_LAUNCH_CLUSTER(1, 32, 8);
#undef QP_LAUNCH_CLUSTER
    return fail(QP_ERR_ARG, "I = %d = %d * %d * 512 is not instantiated for the cluster kernel (use qp_silu_mul_had)", I, Kf, R);
}

extern "C" int qp_silu_mul_had_grid_xchg(void *x_out_f16, const float *acc_local, const void *wscale_local_f16, float acc_scale,
                                         const void *su_f16, int I, float had_scale, float *zero_ptr, int zero_count,
                                         unsigned *sync_counter, const qp_xchg *xc, void *stream) {
    QP_CHECK_ARG(x_out_f16 && acc_local && wscale_local_f16 && sync_counter, "NULL pointer argument");
    QP_CHECK_ARG(xc && xc->peer_base && xc->peer_flags && xc->epoch, "NULL exchange descriptor");
    QP_CHECK_ARG(xc->nranks >= 2 && xc->nranks <= 32 && xc->rank >= 0 && xc->rank < xc->nranks, "bad rank %d of %d", xc->rank, xc->nranks);
    int m, Kf;
    int rc = had_dims(I, m, Kf);
    if (rc != QP_OK) return rc;
    QP_CHECK_ARG(m % kSiluBlk == 0, "I = %d: needs a power-of-two factor >= 512", I);
    const int R = m / kSiluBlk, NB = Kf * R;
    QP_CHECK_ARG(NB % xc->nranks == 0, "I / 512 = %d blocks do not split over %d ranks", NB, xc->nranks);
    QP_CHECK_ARG(xc->offset % 16 == 0, "exchange buffer must be 16-byte aligned");
    void (*kern)(__half *, const float *, const __half *, float, const __half *, int, float, float *, int, unsigned *, XchgDev) =
        nullptr;
    if (Kf == 28 && R == 1) kern = silu_mul_had_grid_xchg_kernel<28, 1>;
    else if (Kf == 28 && R == 2) kern = silu_mul_had_grid_xchg_kernel<28, 2>;
    else if (Kf == 1 && R == 8) kern = silu_mul_had_grid_xchg_kernel<1, 8>;
    else if (Kf == 1 && R == 16) kern = silu_mul_had_grid_xchg_kernel<1, 16>;
    QP_CHECK_ARG(kern != nullptr, "I = %d = %d * %d * 512 is not instantiated for the multi-CTA kernel", I, Kf, R);
    XchgDev d;
    d.peer_base = (unsigned char *const *)xc->peer_base;
    d.peer_flags = (unsigned *const *)xc->peer_flags;
    d.epoch = xc->epoch + xc->site;
    d.offset = xc->offset;
    d.slice_bytes = xc->slice_bytes;
    d.rank = xc->rank;
    d.nranks = xc->nranks;
    d.site = xc->site;
    QP_CUDA(launch_pdl(kern, dim3(NB / xc->nranks), dim3(kSiluThreads), 0, (cudaStream_t)stream, (__half *)x_out_f16, acc_local,
                       (const __half *)wscale_local_f16, acc_scale, (const __half *)su_f16, I, had_scale, zero_ptr, zero_count,
                       sync_counter, d));
    return check_launch("silu_mul_had_grid_xchg");
}

// scratch of the split kernel: [H] tickets (zeroed once by the caller, left zero by every launch) | [H][S][kAttnPart] partials
static int attn_splits(int max_seq) { return (max_seq + kAttnChunk - 1) / kAttnChunk; }
extern "C" size_t qp_rope_attention_scratch_bytes(int H, int D, int max_seq) {
    if (D != 128 || attn_splits(max_seq) <= 1) return 0;
    return (((size_t)H * 4 + 15) & ~(size_t)15) + (size_t)H * attn_splits(max_seq) * kAttnPart * 4;
}

extern "C" int qp_rope_attention(void *attn_out_f16, const float *acc_qkv, const void *wscale_f16, float acc_scale,
                                 const float *inv_freq, void *kcache_f16, void *vcache_f16, const int *pos_ptr, int H,
                                 int Hkv, int D, int max_seq, int qvk_order, float *zero_ptr, int zero_count, void *scratch,
                                 void *stream) {
    QP_CHECK_ARG(attn_out_f16 && acc_qkv && wscale_f16 && inv_freq && kcache_f16 && vcache_f16 && pos_ptr, "NULL pointer");
    QP_CHECK_ARG(D <= 128 && D % 2 == 0 && H % Hkv == 0 && kAttnThreads % D == 0, "unsupported head geometry H=%d Hkv=%d D=%d", H, Hkv, D);
    QP_CHECK_ARG(max_seq >= 1, "max_seq = %d", max_seq);
    if (D == 128) {
        const int S = attn_splits(max_seq);
        QP_CHECK_ARG(S == 1 || scratch, "max_seq = %d needs the scratch buffer of qp_rope_attention_scratch_bytes()", max_seq);
        QP_CHECK_ARG(S <= 65535, "max_seq = %d too large", max_seq);
        unsigned *tickets = (unsigned *)scratch;
        float *part = scratch ? (float *)((unsigned char *)scratch + (((size_t)H * 4 + 15) & ~(size_t)15)) : nullptr;
        QP_CUDA(launch_pdl(rope_attention_split_kernel, dim3(H, S), dim3(kAttnThreads), 0, (cudaStream_t)stream,
                           (__half *)attn_out_f16, acc_qkv, (const __half *)wscale_f16, acc_scale, inv_freq, (__half *)kcache_f16,
                           (__half *)vcache_f16, pos_ptr, H, Hkv, max_seq, qvk_order, zero_ptr, zero_count, part, tickets, S));
        return check_launch("rope_attention_split");
    }
    const size_t smem = (size_t)(11 * D + max_seq) * 4;
    QP_CHECK_ARG(smem <= (size_t)kMaxSmem - 1024, "max_seq = %d too large for the single-pass attention kernel", max_seq);
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(rope_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 1024));
    }
    QP_CUDA(launch_pdl(rope_attention_kernel, dim3(H), dim3(kAttnThreads), smem, (cudaStream_t)stream, (__half *)attn_out_f16,
                       acc_qkv, (const __half *)wscale_f16, acc_scale, inv_freq, (__half *)kcache_f16,
                       (__half *)vcache_f16, pos_ptr, H, Hkv, D, max_seq, qvk_order, zero_ptr, zero_count));
    return check_launch("rope_attention");
}

extern "C" int qp_gemv_f16(float *out, const void *W_f16, const void *x_f16, int rows, int K, void *stream) {
    QP_CHECK_ARG(out && W_f16 && x_f16, "NULL pointer argument");
    QP_CHECK_ARG(K % 8 == 0 && (size_t)K * 2 <= 96 * 1024, "K = %d unsupported", K);
    static DeviceOnce configured;
    if (configured.first()) {
        QP_CUDA(cudaFuncSetAttribute(gemv_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
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

/* limit of the in-kernel flag waits of the peer-exchange kernels on the CURRENT device; ms <= 0 disables the trap */
static unsigned long long g_spin_limit_host = 120000000000ull;  // mirror of qp::g_spin_limit_cycles for descriptors built on the host
namespace qp {
unsigned long long qp_spin_limit_cycles_host() { return g_spin_limit_host; }
}  // namespace qp

extern "C" int qp_set_spin_timeout_ms(long long ms) {
    int dev = 0, khz = 0;
    QP_CUDA(cudaGetDevice(&dev));
    QP_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const unsigned long long cyc = ms <= 0 ? 0ull : (unsigned long long)ms * (unsigned long long)(khz > 0 ? khz : 2000000);
    QP_CUDA(cudaMemcpyToSymbol(qp::g_spin_limit_cycles, &cyc, sizeof(cyc)));
    g_spin_limit_host = cyc;
    return QP_OK;
}
