/* qpalette.h -- C ABI of libqpalette.so: B200 (sm_100a) kernels for Q-Palette's quantized-linear decode path.
 *
 * Every entry point takes raw DEVICE pointers, sizes and a CUDA stream (cudaStream_t passed as void*), never
 * allocates, never synchronises, is CUDA-graph capturable, and returns QP_OK or a negative error code
 * (qp_last_error() gives the message).  These are the functions the reference's pybind/torch FFI for this path
 * would bind; each one cites the reference interface it replaces (paths relative to the Q-Palette repo).
 * The `_host` variants take HOST buffers for the activations/result (weights stay device resident) and perform the
 * host<->device copies on the given stream; they are what bench.py's `e2e` number goes through.
 */
#ifndef QPALETTE_H
#define QPALETTE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QP_OK 0
#define QP_ERR_ARG (-1)    /* bad shape / unsupported parameter combination */
#define QP_ERR_CUDA (-2)   /* CUDA runtime error (launch, attribute, copy) */
#define QP_ERR_ALIGN (-3)  /* pointer alignment requirement violated */

/* split_mode for two-rate TCQ ("comb" quantizers) */
#define QP_SPLIT_NONE 0
#define QP_SPLIT_IN 1   /* tcomb_*: codes1 covers input columns [0,part1), codes2 the rest   (CombtLinearTCQ) */
#define QP_SPLIT_OUT 2  /* comb_*:  codes1 covers output rows   [0,part1), codes2 the rest   (CombLinearTCQ)  */

/* flags */
#define QP_FLAG_ACCUMULATE 1u /* GEMV: add into `out` instead of overwriting it (out must hold valid fp32 data) */

/* epilogue selector of qp_incoherent_* (fused layer ops) */
#define QP_EPI_NONE 0
#define QP_EPI_SILU_MUL 1 /* out[i] = silu(y[M/2 + i]) * y[i], i < M/2   (IncoherentMLP.compute_ug, merged up|gate) */

int qp_version(void);
const char *qp_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches claim) */
uint64_t qp_launch_count(void);
int qp_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * TCQ (bitshift trellis, L=16, V=2).  Packed layout = the `trellis` int16 buffer written by quantize_layer.py
 * (lib/quantizer/tcq_quant.py:47-60).  S = tlut_bits in {9,10,11}, KV in 2..10 bits per weight pair.
 * tlut: fp16 (2^S, 2).  x: fp16 (bs, K) row-major, bs <= 8.  out: fp32 (bs, M).
 *
 * Replaces tcq_kernels.decompress_gemm_16_{S}_{KV}_1_{M}_{N}_{K}, decompress_gemm_comb_*, decompress_gemm_combt_*
 * (kernels/tcq-kernels/src/inference.cu:408-1218; python side lib/linear/__init__.py:176-250), shape-generic.
 * For split modes codes2/KV2 describe the second part and part1 is its boundary (K1 or M1, multiple of 32/64).
 * ------------------------------------------------------------------------------------------------------------- */
int qp_tcq_gemv(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16,
                int M, int K, int bs, int S, int KV1, int KV2, int split_mode, int part1, unsigned flags,
                void *stream);

/* Replaces tcq_kernels.decompress_16_{S}_{KV}, decompress_comb_*, decompress_combt_*
 * (kernels/tcq-kernels/src/inference.cu:1222-1819; lib/linear/__init__.py:259-337).  W: fp16 (M, K) row-major. */
int qp_tcq_dequant(void *W_f16, const void *codes1, const void *codes2, const void *tlut_f16, int M, int K, int S,
                   int KV1, int KV2, int split_mode, int part1, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * VQ (vec_sz 2) / SQ (vec_sz 1) in the tensor-core layout (`qweight` int32 (M, bits*K/32/vec_sz),
 * lib/quantizer/quant_op.py:101-162).  lut: fp16 (2^bits, vec_sz).  out: fp32 (bs, M), bs <= 8.
 * Replaces vq_tensor_kernels.decompress_gemm_{bits}_{M}_{N}_{K}_{sq_dup|sq|vq2} and decompress_{bits}_{vtype}
 * (kernels/vq-tensor-kernels/src/inference.cu:570-1108; lib/linear/__init__.py:43-117).
 * ------------------------------------------------------------------------------------------------------------- */
int qp_lut_gemv(float *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs,
                int bits, int vec_sz, unsigned flags, void *stream);
int qp_lut_dequant(void *W_f16, const void *codes, const void *lut_f16, int M, int K, int bits, int vec_sz,
                   void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused dequantise + batched GEMM on tcgen05 tensor cores (TMEM accumulators), 1 <= bs <= 128 (meant for bs >= 16):
 * out (bs, M) fp32 = x (bs, K) @ decode(W)^T without materialising W.  Same packed inputs as the GEMV entry points;
 * needs M % 128 == 0, K % 64 == 0.  Replaces the reference's bs > 8 path `decompress_* + x @ dq.T`
 * (lib/linear/tcq_linear.py:75-84, comb_linear.py:100-125,246-268, vq_linear.py:58-66).  qp_lut_gemm_tc covers vq2
 * (2..12 bits) and SQ up to 5 bits; wider SQ falls back to qp_lut_dequant + GEMM on the host side.
 * ------------------------------------------------------------------------------------------------------------- */
int qp_tcq_gemm_tc(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16, int M,
                   int K, int bs, int S, int KV1, int KV2, int split_mode, int part1, unsigned flags, void *stream);
/* The same contraction for small batches (meant for 9 <= bs <= 32; any bs <= 128 runs, 32 batch rows per launch) on the
 * GEMV's decode loop: a 32x32 super-tile is decoded once and multiplied with the whole batch by mma.sync (fp32 accumulate);
 * the work is ordered by K-slabs so that a CTA keeps its slab of x fragments in shared memory, and a finished (strip, slab)
 * is added to out with fp32 atomics.  Needs M % 32 == 0 and parts of K % 32 == 0 only.  `scratch`: device buffer of qp_gemm_mma_scratch_bytes(K, bs) bytes (a fragment-ordered copy of x is built there).
 * qp_lut_gemm_mma: every VQ (vec_sz 2, 2..12 bits) and SQ (vec_sz 1, 2..8 bits) format of qp_lut_gemv.
 * TCQ: tlut_bits must be the reference's pairing for the rate (9 up to KV = 8, KV + 1 above; lib/utils/mem_op.py). */
size_t qp_gemm_mma_scratch_bytes(int K, int bs);
int qp_tcq_gemm_mma(float *out, const void *codes1, const void *codes2, const void *x_f16, const void *tlut_f16, void *scratch,
                    int M, int K, int bs, int S, int KV1, int KV2, int split_mode, int part1, unsigned flags, void *stream);
int qp_lut_gemm_mma(float *out, const void *codes, const void *x_f16, const void *lut_f16, void *scratch, int M, int K, int bs,
                    int bits, int vec_sz, unsigned flags, void *stream);
int qp_lut_gemm_tc(float *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs, int bits,
                   int vec_sz, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * VQ / SQ in the SIMT layout (lib/quantizer/pack_op.py:288-335, quant_op.py:33-86).  out: fp16 (bs, M).
 * Replaces sq_pack_gemm.pack_gemm / pack_dequant (kernels/sq-cuda-kernels/gemm_routines.cu:474-722) and
 * vq_pack_gemm.vq_pack_gemm_* / vq_pack_dequant_* (kernels/vq-cuda-kernels/src/gemm_routines.cu:1913-2120).
 * Accumulates in fp32 (the reference accumulates in fp16).  vec_sz 1 (SQ, 2..8 bits), 2 (3..12 bits; 2 accepted) and
 * 4 (6..12 bits): the op table of lib/linear/__init__.py:339-420.  K % (32*vec_sz) == 0; a last chunk shorter than
 * 32*32*vec_sz weights is packed with the reduced thread count the reference packer uses (quant_op.py:15-31).
 * ------------------------------------------------------------------------------------------------------------- */
int qp_simt_gemv(void *out, const void *codes, const void *x_f16, const void *lut_f16, int M, int K, int bs,
                 int bits, int vec_sz, int out_is_f32 /* 0: fp16 (bs,M) like the reference op; 1: fp32 */, void *stream);
int qp_simt_dequant(void *W_f16, const void *codes, const void *lut_f16, int M, int K, int bits, int vec_sz,
                    void *stream);
/* GPU version of lib/quantizer/quant_op.py:246-257 convert_tensor_core_to_simt (format conversion at load time) */
int qp_convert_tc_to_simt(void *simt_codes, const void *tc_codes, int M, int K, int bits, int vec_sz, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Incoherence processing.  y = ((hadK^T (x) H_{n/Kf}) (x * su)) * scale   along the last dim, Sylvester order;
 * Kf in {1, 28} (hadK = the reference's get_had28, applied transposed as every inference call site does).
 * Replaces hadamard::hadamard (lib/utils/matmul_had.py:124-134 -> fast_hadamard_transform) plus the dense
 * `hadK @ .` of matmul_hadU_cuda / matmul_hadU_head_cuda (lib/utils/matmul_had.py:94-106,137-147) and the `x*SU`
 * / `/scale` / `.half()` glue around them (lib/linear/incoherent_linear.py:82,106,325,336,491).
 * x, y: (rows, n) with dtype selected by *_is_f32 (0 = fp16, 1 = fp32); su: fp16 (n,) or NULL.
 * ------------------------------------------------------------------------------------------------------------- */
int qp_hadamard(void *y, const void *x, const void *su_f16, int rows, int n, float scale, int x_is_f32,
                int y_is_f32, void *stream);

/* out (bs, M) fp16 = fp16(acc) * wscale * scale   [+ epilogue]; acc = fp32 (bs, M) GEMV result.
 * The `.to(fp16) * Wscale * scale` (+ split + SiLU*mul) glue of lib/linear/incoherent_linear.py:83-108,326-338. */
int qp_scale_epilogue(void *out_f16, const float *acc, const void *wscale_f16, int bs, int M, float scale,
                      int epilogue, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Decode-step glue (bs = 1).  The reference runs these as torch eager ops fused by torch.compile/Inductor
 * (eval/measure_latency.py:223-225); north_star rules Triton out, so they are kernels here.  All are latency-bound
 * single-/few-CTA kernels that read their scalars (token, position) from DEVICE memory so that one decode step can be
 * captured once as a CUDA graph and replayed per token.
 * ------------------------------------------------------------------------------------------------------------- */
/* [h += fp16(acc)*wscale*acc_scale (written back if h_writeback)] -> [RMSNorm(norm_w, eps)] -> [*su] -> [Hadamard] ->
 * x_out = fp16(. * had_scale).  acc/wscale, norm_w, su may be NULL.  Also zeroes zero_count floats at zero_ptr (the next
 * GEMV's accumulators).  lib/linear/incoherent_linear.py:76-108,324-338 + LlamaDecoderLayer residual/RMSNorm. */
int qp_fused_norm_had(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                      float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n, float had_scale,
                      int do_had, float *zero_ptr, int zero_count, void *stream);
/* Row-sharded layers (one process per GPU): the buffer a layer boundary gathers (attention output, o / down accumulators,
 * SiLU*mul activations -- SURVEY 8e) lives at `offset` of an "exchange region" every rank has allocated with
 * qp_peer_alloc and mapped from every peer (qp_peer_export -> any host transport -> qp_peer_import).  Rank r has produced
 * bytes [r*slice_bytes, (r+1)*slice_bytes) of it; qp_fused_norm_had_xchg completes it in place by pushing that slice to all
 * peers over NVLink and waiting for theirs (flags[site*nranks + source] inside each rank's region, epoch[site] local),
 * then runs qp_fused_norm_had on it: all-gather + consumer in one kernel, no NCCL call.  A site is one (layer, gather
 * point); every rank must run the same sequence of sites.  zero_ptr is cleared BEFORE the flags are published. */
typedef struct qp_xchg {
    void *const *peer_base;     /* device array [nranks] of region base pointers (entry `rank` = the local region) */
    unsigned *const *peer_flags;/* device array [nranks] of flag-array pointers (inside the regions) */
    unsigned *epoch;            /* local device array [nsites], zero-initialised */
    long long offset;           /* of the gathered buffer inside the region, bytes, multiple of 16 */
    int slice_bytes;            /* per rank, multiple of 16 */
    int rank, nranks, site;
} qp_xchg;
int qp_fused_norm_had_xchg(void *x_out_f16, void *h_f16, int h_writeback, const float *acc, const void *wscale_f16,
                           float acc_scale, const void *norm_w_f16, float eps, const void *su_f16, int n, float had_scale,
                           int do_had, float *zero_ptr, int zero_count, const qp_xchg *xc, void *stream);
/* Low-latency ("LL") gather for a consumer GEMV with a fused prologue (the row-sharded decode step uses it at the three
 * sites in front of the qkv / o / up-gate projections): store this rank's slice (n elements;
 * fp16, or fp32 accumulators converted to fp16) as {2 x fp16, epoch, 2 x fp16, epoch} entries into the buffer at xc->offset of
 * EVERY rank's region (xc->slice_bytes = 4 * n per rank) and bump the site's epoch -- no fence, no flag, no wait.  The consumer
 * (qp_*_gemv_fused with qp_xprod.ll = local region base + xc->offset, ll_epoch = xc->epoch + xc->site) polls the epoch of each
 * entry it reads.  Clears zero_ptr[0..zero_count) first. */
int qp_xchg_send_ll(const void *src_slice, int src_is_f32, int n, float *zero_ptr, int zero_count, const qp_xchg *xc,
                    void *stream);
/* Row-sharded SiLU*mul + Hadamard: acc_local / wscale_local = [up | gate] of this rank's I / nranks rows; one CTA per local
 * 512-element block pushes its transformed block into the exchange buffer at xc->offset (I floats) of every rank, then every
 * rank finishes the transform from its own copy: x_out = the complete fp16 vector on every rank.  *sync_counter: a local
 * device word, 0 before the first call.  I / 512 must be a multiple of nranks (28*512: 2, 4; 28*1024: 2, 4, 8). */
int qp_silu_mul_had_grid_xchg(void *x_out_f16, const float *acc_local, const void *wscale_local_f16, float acc_scale,
                              const void *su_f16, int I, float had_scale, float *zero_ptr, int zero_count,
                              unsigned *sync_counter, const qp_xchg *xc, void *stream);
/* how long the in-kernel flag waits of the exchange kernels spin before the kernel traps (so that a dead peer fails the
 * CUDA context instead of hanging the GPU); per device, default ~60 s, ms <= 0 = wait forever */
int qp_set_spin_timeout_ms(long long ms);
int qp_peer_alloc(void **ptr, size_t bytes);             /* cudaMalloc + clear */
int qp_peer_free(void *ptr);
int qp_peer_export(void *ptr, void *handle64);           /* 64-byte CUDA IPC handle of a qp_peer_alloc region */
int qp_peer_import(const void *handle64, void **ptr);    /* map a peer's region (enables peer access) */
int qp_peer_close(void *ptr);

/* acc = [up | gate] (2*I fp32): x_out = fp16(Hadamard(silu(gate)*up * su) * had_scale)   (IncoherentMLP.compute_ug tail
 * + compute_dp head, lib/linear/incoherent_linear.py:324-338) */
int qp_silu_mul_had(void *x_out_f16, const float *acc, const void *wscale_f16, float acc_scale, const void *su_f16,
                    int I, float had_scale, float *zero_ptr, int zero_count, void *stream);
/* same result on I / 512 cooperating CTAs (I = 28*512, 28*1024 or 2^k >= 4096): `acc` is consumed (its first I words are
 * used as exchange space), *sync_counter is a device word that is 0 before the first call and reserved to one stream. */
/* same result on ONE thread-block cluster of <= 8 CTAs exchanging their blocks through distributed shared memory (no scratch,
 * `acc` untouched); same shapes as qp_silu_mul_had_grid */
int qp_silu_mul_had_cluster(void *x_out_f16, const float *acc, const void *wscale_f16, float acc_scale, const void *su_f16,
                            int I, float had_scale, float *zero_ptr, int zero_count, void *stream);
int qp_silu_mul_had_grid(void *x_out_f16, float *acc, const void *wscale_f16, float acc_scale, const void *su_f16, int I,
                         float had_scale, float *zero_ptr, int zero_count, unsigned *sync_counter, void *stream);
/* acc_qkv = [q | k | v] fp32 GEMV sums ([q | v | k] when qvk_order != 0: a merge_qv layer, whose Wscale_qkv the reference
 * stores in that order, lib/linear/incoherent_linear.py:211-213): Wscale epilogue, RoPE, KV-cache append at *pos_ptr, causal
 * attention of the new token over the cache (IncoherentSdpaAttention.forward, :110-203).  D = 128: grid (heads, ceil(max_seq /
 * 128)), each CTA takes 128 positions of one query head and the last one to finish combines the partial softmaxes; `scratch`
 * (qp_rope_attention_scratch_bytes(H, D, max_seq) bytes, zeroed ONCE by the caller; may be NULL when that is 0) holds the
 * partials and one ticket per head, and may be shared by all layers of a stream.  Other head sizes: one CTA per head.
 * *pos_ptr >= max_seq is clamped to the last cache row (memory safety only; the host API refuses to step that far). */
size_t qp_rope_attention_scratch_bytes(int H, int D, int max_seq);
int qp_rope_attention(void *attn_out_f16, const float *acc_qkv, const void *wscale_f16, float acc_scale,
                      const float *inv_freq, void *kcache_f16, void *vcache_f16, const int *pos_ptr, int H, int Hkv,
                      int D, int max_seq, int qvk_order, float *zero_ptr, int zero_count, void *scratch, void *stream);
/* fp16 GEMV for the unquantized lm_head: out (rows) fp32 = W (rows, K) @ x (K) */
int qp_gemv_f16(float *out, const void *W_f16, const void *x_f16, int rows, int K, void *stream);
int qp_argmax(int *token_out, const float *logits, int n, void *scratch, void *stream);
int qp_embed(void *h_f16, const void *table_f16, const int *token, int n, void *stream);
int qp_step_advance(int *pos, int *history, const int *token, int max_hist, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GEMV with the activation-side glue fused into its prologue (bs = 1): every CTA of the GEMV computes
 *     h' = src + fp16(acc)*wscale*acc_scale   (optional residual add; CTA 0 stores h' to h_out_f16, which must not alias src)
 *     y  = RMSNorm(h', norm_w, eps)            (optional)
 *     x  = fp16(Hadamard(y * su) * had_scale)  (su optional)
 * itself while its first weight loads are in flight, then runs out += decode(W) x (out must be pre-zeroed / hold the
 * value to add to).  x_out_f16 (optional) receives x for sibling projections; zero1/zero2 name fp32 buffers to clear for
 * LATER launches; they are cleared before the kernel waits for its predecessor, so neither this launch NOR the launch
 * immediately before it on the stream may read or write them.  Replaces a qp_fused_norm_had launch + the x staging.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct qp_xprod {
    const void *src_f16;
    void *h_out_f16;
    const float *acc;
    const void *wscale_f16;
    float acc_scale;
    const void *norm_w_f16;
    float eps;
    const void *su_f16;
    float had_scale;
    void *x_out_f16;
    float *zero1;
    int zero1_count;
    float *zero2;
    int zero2_count;
    /* row-sharded mode (all NULL / 0 otherwise): one operand is polled out of the LL receive buffer qp_xchg_send_ll fills.
     * ll_kind 1: the entries are fp16(acc) of the previous projection (acc must be NULL, wscale_f16 given);
     * ll_kind 2: the entries are src (src_f16 is then ignored).  ll_epoch: the exchange site's epoch word. */
    const void *ll;
    const unsigned *ll_epoch;
    int ll_kind;
} qp_xprod;
int qp_tcq_gemv_fused(float *out, const void *codes1, const void *codes2, const qp_xprod *xp, const void *tlut_f16,
                      int M, int K, int S, int KV1, int KV2, int split_mode, int part1, void *stream);
int qp_lut_gemv_fused(float *out, const void *codes, const qp_xprod *xp, const void *lut_f16, int M, int K, int bits,
                      int vec_sz, void *stream);

/* ---------------------------------------------------------------------------------------------------------------
 * host-buffer variant used for end-to-end timing: x_host (bs,K) fp16 pinned/pageable host memory, out_host (bs,M)
 * fp32 host memory; x_dev/out_dev are device scratch buffers of the same sizes.  Copies run on `stream`; the caller
 * synchronises the stream before reading out_host.
 * ------------------------------------------------------------------------------------------------------------- */
int qp_tcq_gemv_host(float *out_host, float *out_dev, const void *codes1, const void *codes2, const void *x_host,
                     void *x_dev, const void *tlut_f16, int M, int K, int bs, int S, int KV1, int KV2,
                     int split_mode, int part1, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QPALETTE_H */
