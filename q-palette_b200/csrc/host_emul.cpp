// host_emul.cpp -- CPU lane-emulator of the kernels' integer decode logic (TEST SUPPORT, built by oracle/Makefile
// into oracle/_build/libqp_emul.so; never linked into libqpalette.so).  It runs the exact header the CUDA kernels
// compile (tcq_bits.cuh) with the warp shuffle replaced by an array lookup, so the extraction code can be checked
// against the oracle in the GPU-less build container.
#include <cstdint>
#include <cstring>
#include <vector>

#include "tcq_bits.cuh"
#include "lut_bits.cuh"

template <int KV, int T>
static void tile_states(const uint32_t (&P)[32][TcqGeom<KV>::kWords], const uint32_t (&send)[32][4], int g,
                        uint16_t *out /* [kl][ml][j] base for this lane */) {
    uint32_t u[4];
    tcq_states<KV, T>(P[g], send[(g + 1) & 31][T], send[(g + 2) & 31][T], u);
    for (int j = 0; j < 4; ++j) out[T * 4 + j] = (uint16_t)(u[j] & 0xFFFFu);
}

template <int KV>
static void emul(const uint8_t *buf, int M, int K, uint16_t *out) {
    using G = TcqGeom<KV>;
    const long nsuper = (long)(M / 32) * (K / 32);
    const size_t total = (size_t)nsuper * G::kSuperBytes;
    // pad so the word loads never run past the buffer in the emulator (the kernel proves this separately)
    std::vector<uint8_t> padded(total + 8, 0);
    std::memcpy(padded.data(), buf, total);
    for (long s = 0; s < nsuper; ++s) {
        const uint8_t *sp = padded.data() + (size_t)s * G::kSuperBytes;
        uint32_t P[32][G::kWords];
        uint32_t send[32][4];
        for (int g = 0; g < 32; ++g) {
            int w0, bo;
            tcq_lane_addr<KV>(g, w0, bo);
            uint32_t raw[G::kRawWords];
            for (int i = 0; i < G::kRawWords; ++i) std::memcpy(&raw[i], sp + 4 * (w0 + i), 4);
            tcq_align<KV>(raw, bo, P[g]);
            send[g][0] = tcq_send<KV, 0>(P[g]);
            send[g][1] = tcq_send<KV, 1>(P[g]);
            send[g][2] = tcq_send<KV, 2>(P[g]);
            send[g][3] = tcq_send<KV, 3>(P[g]);
        }
        for (int g = 0; g < 32; ++g) {
            uint16_t *o = out + ((size_t)s * 32 + g) * 16;
            tile_states<KV, 0>(P, send, g, o);
            tile_states<KV, 1>(P, send, g, o);
            tile_states<KV, 2>(P, send, g, o);
            tile_states<KV, 3>(P, send, g, o);
        }
    }
}

extern "C" int qp_emul_tcq_states(const uint8_t *buf, int M, int K, int KV, uint16_t *out) {
    switch (KV) {
        case 2: emul<2>(buf, M, K, out); break;
        case 3: emul<3>(buf, M, K, out); break;
        case 4: emul<4>(buf, M, K, out); break;
        case 5: emul<5>(buf, M, K, out); break;
        case 6: emul<6>(buf, M, K, out); break;
        case 7: emul<7>(buf, M, K, out); break;
        case 8: emul<8>(buf, M, K, out); break;
        case 9: emul<9>(buf, M, K, out); break;
        case 10: emul<10>(buf, M, K, out); break;
        default: return -1;
    }
    return 0;
}

// does the last lane of the last super-tile ever read past the end of the buffer?  returns max over lanes of
// (last byte read + 1) relative to the super-tile start; must be <= kSuperBytes.
extern "C" int qp_emul_tcq_max_read(int KV) {
    int mx = 0;
    auto f = [&](auto tag) {
        constexpr int kv = decltype(tag)::value;
        for (int g = 0; g < 32; ++g) {
            int w0, bo;
            tcq_lane_addr<kv>(g, w0, bo);
            const int end = 4 * (w0 + TcqGeom<kv>::kRawWords);
            if (end > mx) mx = end;
        }
        return TcqGeom<kv>::kSuperBytes;
    };
    int sb = 0;
    switch (KV) {
        case 2: sb = f(std::integral_constant<int, 2>{}); break;
        case 3: sb = f(std::integral_constant<int, 3>{}); break;
        case 4: sb = f(std::integral_constant<int, 4>{}); break;
        case 5: sb = f(std::integral_constant<int, 5>{}); break;
        case 6: sb = f(std::integral_constant<int, 6>{}); break;
        case 7: sb = f(std::integral_constant<int, 7>{}); break;
        case 8: sb = f(std::integral_constant<int, 8>{}); break;
        case 9: sb = f(std::integral_constant<int, 9>{}); break;
        case 10: sb = f(std::integral_constant<int, 10>{}); break;
        default: return -1;
    }
    return mx - sb;  // <= 0 means in-bounds
}

// ---- VQ / SQ tensor-core layout: pair codes per (super-tile, lane, tile, register) -----------------------------------
template <int E, int T>
static void lut_tile(const uint32_t (&P)[TcqGeom<E>::kWords], uint32_t *o) {
    o[T * 4 + 0] = lut_pair_offset<E, T, 0, 7>(P) >> 7;
    o[T * 4 + 1] = lut_pair_offset<E, T, 1, 7>(P) >> 7;
    o[T * 4 + 2] = lut_pair_offset<E, T, 2, 7>(P) >> 7;
    o[T * 4 + 3] = lut_pair_offset<E, T, 3, 7>(P) >> 7;
    if constexpr (E % 2 == 0) {  // the split lookups must agree with the pair code
        constexpr uint32_t m = (1u << (E / 2)) - 1u;
        const uint32_t c0 = lut_single_offset<E, T, 2, 0, 7>(P) >> 7, c1 = lut_single_offset<E, T, 2, 1, 7>(P) >> 7;
        if (c0 != (o[T * 4 + 2] & m) || c1 != (o[T * 4 + 2] >> (E / 2))) o[T * 4 + 2] = 0xFFFFFFFFu;
    }
}

template <int E>
static void lut_emul(const uint8_t *buf, int M, int K, uint32_t *out) {
    using G = TcqGeom<E>;
    const long nsuper = (long)(M / 32) * (K / 32);
    const size_t total = (size_t)nsuper * G::kSuperBytes;
    std::vector<uint8_t> padded(total + 8, 0);
    std::memcpy(padded.data(), buf, total);
    for (long s = 0; s < nsuper; ++s) {
        const uint8_t *sp = padded.data() + (size_t)s * G::kSuperBytes;
        for (int g = 0; g < 32; ++g) {
            int w0, bo;
            tcq_lane_addr<E>(g, w0, bo);
            uint32_t raw[G::kRawWords], P[G::kWords];
            for (int i = 0; i < G::kRawWords; ++i) std::memcpy(&raw[i], sp + 4 * (w0 + i), 4);
            tcq_align<E>(raw, bo, P);
            uint32_t *o = out + ((size_t)s * 32 + g) * 16;
            lut_tile<E, 0>(P, o);
            lut_tile<E, 1>(P, o);
            lut_tile<E, 2>(P, o);
            lut_tile<E, 3>(P, o);
        }
    }
}

// out[(super*32 + lane)*16 + tile*4 + j] = pair code
extern "C" int qp_emul_lut_pairs(const uint8_t *buf, int M, int K, int E, uint32_t *out) {
    switch (E) {
#define QP_CASE(e) case e: lut_emul<e>(buf, M, K, out); break;
        QP_CASE(2) QP_CASE(3) QP_CASE(4) QP_CASE(5) QP_CASE(6) QP_CASE(7) QP_CASE(8) QP_CASE(9)
        QP_CASE(10) QP_CASE(11) QP_CASE(12) QP_CASE(14) QP_CASE(16)
#undef QP_CASE
        default: return -1;
    }
    return 0;
}
