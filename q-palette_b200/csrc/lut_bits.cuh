// lut_bits.cuh -- code extraction for the VQ (vec_sz 2) / SQ (vec_sz 1) tensor-core packed layout
// (lib/quantizer/quant_op.py:101-162).  Host/device, like tcq_bits.cuh (same lane/super-tile geometry with
// E = bits per weight PAIR: E = R for vec_sz 2, E = 2R for vec_sz 1).
//   lane payload = 16*E bits, tile t = kl*2+ml owns bits [4tE, 4(t+1)E); register j's pair code = bits [jE, (j+1)E) of it,
//   LSB first.  vec_sz 1: the pair code is code(2j) | code(2j+1) << R.
#pragma once
#include "tcq_bits.cuh"

// (pair code of tile T, register J) << SL, i.e. the byte offset of its slot in a table with 2^SL-byte slots
template <int E, int T, int J, int SL>
QP_HD uint32_t lut_pair_offset(const uint32_t (&P)[TcqGeom<E>::kWords]) {
    constexpr int o = 4 * T * E + J * E;
    constexpr uint32_t mask = ((1u << E) - 1u) << SL;
    if constexpr (o >= SL) return tcq_window<E, o - SL>(P) & mask;
    else return (P[0] << (SL - o)) & mask;
}

// single code C (0/1) of R = E/2 bits inside the pair code (vec_sz 1, split lookup)
template <int E, int T, int J, int C, int SL>
QP_HD uint32_t lut_single_offset(const uint32_t (&P)[TcqGeom<E>::kWords]) {
    constexpr int R = E / 2;
    constexpr int o = 4 * T * E + J * E + C * R;
    constexpr uint32_t mask = ((1u << R) - 1u) << SL;
    if constexpr (o >= SL) return tcq_window<E, o - SL>(P) & mask;
    else return (P[0] << (SL - o)) & mask;
}
