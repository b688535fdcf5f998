// tcq_bits.cuh -- bit-stream extraction for the TCQ packed layout, shared by the device kernels and the host
// lane-emulator (tests/ link this header through csrc/host_emul.cpp so the exact extraction code the kernels run
// is checked against the oracle on CPU, where there is no GPU).
//
// Layout facts (reference: lib/quantizer/tcq_quant.py:47-60, lib/codebook/bitshift.py:296-329):
//   * a 32x32 "super-tile" (2x2 tiles of 16x16) is 64*KV contiguous bytes: [lane 32][kl 2][ml 2][KV nibbles]
//   * lane g's payload P = 16*KV bits (little-endian integer); tile t = kl*2+ml owns bits [4*t*KV, 4*(t+1)*KV) = chunk
//   * the tile's circular stream is the 32 lanes' chunks, each MSB first; state p = 4g+j = stream bits [p*KV, p*KV+16)
//   so lane g needs its own chunk followed by the top (16-KV) bits of lane g+1's chunk (two lanes when 5*KV < 16).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QP_HD __host__ __device__ __forceinline__
#else
#define QP_HD inline
#endif

// low 32 bits of (hi:lo) >> s, 0 <= s < 32
QP_HD uint32_t qp_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s == 0 ? lo : ((lo >> s) | (hi << (32 - s)));
#endif
}
// high 32 bits of (hi:lo) << s, 0 <= s < 32
QP_HD uint32_t qp_funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s == 0 ? hi : ((hi << s) | (lo >> (32 - s)));
#endif
}

template <int KV>
struct TcqGeom {
    static_assert(KV >= 2 && KV <= 16, "bits per weight pair out of range");  // TCQ uses 2..10; the LUT layouts reuse the geometry up to 16
    static constexpr int kChunkBits = 4 * KV;                        // per (lane, tile)
    static constexpr int kPayloadBits = 16 * KV;                     // per (lane, super-tile)
    static constexpr int kLaneBytes = 2 * KV;
    static constexpr int kSuperBytes = 64 * KV;
    static constexpr bool kOdd = (KV & 1) != 0;
    // raw 32-bit words fetched per lane: word-aligned window covering the payload (+2 bytes of slack for odd KV)
    static constexpr int kRawWords = kOdd ? (KV + 1) / 2 : KV / 2;
    static constexpr int kWords = (kPayloadBits + 31) / 32;         // aligned payload words
    static constexpr int kNeighbors = (5 * KV >= 16) ? 1 : 2;        // lanes whose chunk tops are needed
};

// word index (in 32-bit words, relative to the super-tile start) of lane g's first raw word, and the bit offset of
// its payload inside that word (0 or 16; 16 only for odd KV and odd g).
template <int KV>
QP_HD void tcq_lane_addr(int lane, int &word0, int &bitoff) {
    const int byte0 = lane * TcqGeom<KV>::kLaneBytes;
    word0 = byte0 >> 2;
    bitoff = (byte0 & 3) * 8;
}

// raw words -> aligned payload words P[0..kWords)
template <int KV>
QP_HD void tcq_align(const uint32_t (&raw)[TcqGeom<KV>::kRawWords], int bitoff, uint32_t (&P)[TcqGeom<KV>::kWords]) {
    using G = TcqGeom<KV>;
    if (!G::kOdd) {
#pragma unroll
        for (int i = 0; i < G::kWords; ++i) P[i] = raw[i];
    } else {
#pragma unroll
        for (int i = 0; i < G::kWords; ++i) {
            const uint32_t lo = raw[i];
            const uint32_t hi = (i + 1 < G::kRawWords) ? raw[i + 1] : 0u;
            P[i] = qp_funnel_r(lo, hi, (uint32_t)bitoff);
        }
    }
}

// 32-bit window of the payload starting at (compile-time) bit position POS; bits past the payload are unspecified.
template <int KV, int POS>
QP_HD uint32_t tcq_window(const uint32_t (&P)[TcqGeom<KV>::kWords]) {
    constexpr int W = TcqGeom<KV>::kWords;
    constexpr int w = POS / 32, s = POS % 32;
    static_assert(POS >= 0 && w < W, "window out of payload");
    if (s == 0) return P[w];
    const uint32_t hi = (w + 1 < W) ? P[w + 1] : 0u;
    return qp_funnel_r(P[w], hi, s);
}

// value a lane publishes to its predecessor(s) for tile T: its chunk left-aligned at bit 31.
// For chunks narrower than 32 bits the low bits are zero (needed by the two-neighbour merge).
template <int KV, int T>
QP_HD uint32_t tcq_send(const uint32_t (&P)[TcqGeom<KV>::kWords]) {
    constexpr int CB = 4 * KV;
    if constexpr (CB <= 32) {
        const uint32_t own = tcq_window<KV, T * CB>(P);
        return CB == 32 ? own : (own << ((32 - CB) & 31));
    } else {
        return tcq_window<KV, T * CB + CB - 32>(P);
    }
}

// one state: J-th 16-bit trellis state of (lane, tile T); low 16 bits of the result are the state, upper bits junk
// (harmless: only the low 16 bits of s*(s+1) are used).
//   own = chunk right-aligned (tcq_window at T*CB), z = own low bits : top (16-KV) bits of the next lane's chunk.
template <int KV, int T, int J>
QP_HD uint32_t tcq_state_j(const uint32_t (&P)[TcqGeom<KV>::kWords], uint32_t own, uint32_t z) {
    constexpr int CB = 4 * KV;
    constexpr int own_bits = (4 - J) * KV;  // stream bits of state J that lie in the own chunk
    if constexpr (own_bits >= 16) {
        if constexpr (CB <= 32) return own >> (own_bits - 16);
        else return tcq_window<KV, T * CB + own_bits - 16>(P);
    } else {
        return z >> ((3 - J) * KV);
    }
}

// the four states of (lane, tile T).  n1/n2 = tcq_send of lanes g+1 / g+2 (n2 only read when KV <= 3).
template <int KV, int T>
QP_HD void tcq_states(const uint32_t (&P)[TcqGeom<KV>::kWords], uint32_t n1, uint32_t n2, uint32_t (&u)[4]) {
    constexpr int CB = 4 * KV;
    if constexpr (TcqGeom<KV>::kNeighbors == 2) {
        // KV = 2, 3: 32 stream bits starting at the lane's chunk: own | next | next-next
        const uint32_t own = tcq_send<KV, T>(P);
        const uint32_t sw = own | (n1 >> CB) | (n2 >> (2 * CB));
        u[0] = sw >> 16;
        u[1] = sw >> (16 - KV);
        u[2] = sw >> (16 - 2 * KV);
        u[3] = sw >> (16 - 3 * KV);
    } else {
        const uint32_t own = tcq_window<KV, T * CB>(P);
        const uint32_t z = qp_funnel_l(n1, own, 16 - KV);
        u[0] = tcq_state_j<KV, T, 0>(P, own, z);
        u[1] = tcq_state_j<KV, T, 1>(P, own, z);
        u[2] = tcq_state_j<KV, T, 2>(P, own, z);
        u[3] = tcq_state_j<KV, T, 3>(P, own, z);
    }
}

// trellis hash: low 16 bits of s*(s+1); bit 15 = sign flip of component 0, bits [15-S, 15) = tlut index
// (lib/codebook/bitshift.py:71-79)
QP_HD uint32_t tcq_hash(uint32_t u) { return u * u + u; }
