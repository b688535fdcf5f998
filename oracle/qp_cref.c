/* qp_cref.c -- multi-threaded C port of the oracle's TCQ decode + matvec (TEST / BASELINE INFRASTRUCTURE ONLY).
 * Restates oracle/qp_oracle.py:tcq_decode + gemv_ref (which follow lib/codebook/bitshift.py:71-79,296-329 and
 * lib/quantizer/tcq_quant.py:47-60 of the reference) in plain C with OpenMP so the CPU baseline can use every host
 * core.  Checked against the numpy oracle in tests/test_oracle_cref.py. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, f;
    if (exp == 0) {
        if (man == 0) f = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400));
            f = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) f = sign | 0x7F800000u | (man << 13);
    else f = sign | ((exp + 112) << 23) | (man << 13);
    float r;
    memcpy(&r, &f, 4);
    return r;
}

/* nibble j of the (lane, tile) chunk; layout [mh][kh][lane][kl][ml][KV] little-endian nibbles */
static inline uint64_t chunk_of(const uint8_t *buf, size_t nib0, int KV) {
    uint64_t c = 0;
    for (int j = 0; j < KV; ++j) {
        size_t n = nib0 + j;
        uint8_t b = buf[n >> 1];
        uint64_t v = (n & 1) ? (b >> 4) : (b & 15);
        c |= v << (4 * j);
    }
    return c;
}

typedef struct {
    const uint8_t *buf; const uint16_t *tlut; int M, K, KV, S; const uint16_t *x; int bs, ldx, col0, row0, ldo;
    float *out; uint16_t *Wout; int ldw; int mh0, mh1;
} cref_job;

static void cref_strips(const cref_job *jb) {
    const uint8_t *buf = jb->buf; const uint16_t *tlut = jb->tlut, *x = jb->x;
    const int K = jb->K, KV = jb->KV, S = jb->S, bs = jb->bs, ldx = jb->ldx, col0 = jb->col0, row0 = jb->row0;
    const int ldo = jb->ldo, ldw = jb->ldw;
    float *out = jb->out; uint16_t *Wout = jb->Wout;
    const int KH = K / 32;
    for (int mh = jb->mh0; mh < jb->mh1; ++mh) {
        double acc[8][32];
        memset(acc, 0, sizeof acc);
        for (int kh = 0; kh < KH; ++kh) {
            const size_t super_nib = ((size_t)mh * KH + kh) * 32 * 4 * KV;
            for (int t = 0; t < 4; ++t) {
                const int kl = t >> 1, ml = t & 1;
                uint64_t ch[32];
                for (int g = 0; g < 32; ++g) ch[g] = chunk_of(buf, super_nib + ((size_t)g * 4 + t) * KV, KV);
                for (int g = 0; g < 32; ++g) {
                    /* 64-bit window of the circular stream starting at lane g's chunk */
                    uint64_t X = 0;
                    int have = 0, gg = g;
                    while (have < 64) {
                        int B = 4 * KV, sh = 64 - have - B;
                        X |= (sh >= 0) ? (ch[gg & 31] << sh) : (ch[gg & 31] >> (-sh));
                        have += B;
                        gg++;
                    }
                    for (int j = 0; j < 4; ++j) {
                        uint32_t s = (uint32_t)((X >> (48 - j * KV)) & 0xFFFF);
                        uint32_t tt = s * (s + 1);
                        uint32_t c = (tt >> (15 - S)) & ((1u << S) - 1);
                        uint16_t w0 = tlut[2 * c], w1 = tlut[2 * c + 1];
                        if (tt & 0x8000u) w0 ^= 0x8000u;
                        const int r = 16 * ml + g / 4 + 8 * (j & 1);
                        const int col = 32 * kh + 16 * kl + 2 * (g % 4) + 8 * (j >> 1);
                        if (Wout) {
                            Wout[(size_t)(row0 + 32 * mh + r) * ldw + col0 + col] = w0;
                            Wout[(size_t)(row0 + 32 * mh + r) * ldw + col0 + col + 1] = w1;
                        }
                        if (out) {
                            const float f0 = half_to_float(w0), f1 = half_to_float(w1);
                            for (int n = 0; n < bs; ++n)
                                acc[n][r] += (double)f0 * half_to_float(x[(size_t)n * ldx + col0 + col]) +
                                             (double)f1 * half_to_float(x[(size_t)n * ldx + col0 + col + 1]);
                        }
                    }
                }
            }
        }
        if (out)
            for (int n = 0; n < bs; ++n)
                for (int r = 0; r < 32; ++r) out[(size_t)n * ldo + row0 + 32 * mh + r] += (float)acc[n][r];
    }
}

static void *cref_thread(void *p) { cref_strips((const cref_job *)p); return NULL; }

static int g_threads = 0;
int qp_cref_threads(void) {
    if (g_threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n > 0 ? (int)(n > 256 ? 256 : n) : 1;
    }
    return g_threads;
}
void qp_cref_set_threads(int n) { g_threads = n > 0 ? n : 0; }

/* out[n*ldo + row0 + row] += sum_k W[row][k] * x[n*ldx + col0 + k] for one TCQ part (M x K); decoded fp16 W is written
 * to Wout (row stride ldw, at (row0, col0)) if non-NULL; out may be NULL (decode only).  Rows are split over threads. */
int qp_cref_tcq(const uint8_t *buf, const uint16_t *tlut, int M, int K, int KV, int S, const uint16_t *x, int bs, int ldx,
                int col0, int row0, int ldo, float *out, uint16_t *Wout, int ldw) {
    const int strips = M / 32;
    int nt = qp_cref_threads();
    if (nt > strips) nt = strips;
    if (nt < 1) nt = 1;
    pthread_t th[256];
    cref_job jobs[256];
    for (int i = 0; i < nt; ++i) {
        cref_job jb = {buf, tlut, M, K, KV, S, x, bs, ldx, col0, row0, ldo, out, Wout, ldw,
                       (int)((long)strips * i / nt), (int)((long)strips * (i + 1) / nt)};
        jobs[i] = jb;
        if (i + 1 < nt) pthread_create(&th[i], NULL, cref_thread, &jobs[i]);
    }
    cref_strips(&jobs[nt - 1]);
    for (int i = 0; i + 1 < nt; ++i) pthread_join(th[i], NULL);
    return 0;
}
