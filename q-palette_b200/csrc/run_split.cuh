// run_split.cuh -- how a GEMV launch divides its super-tiles among the warps of the grid.  Plain integer code shared by the
// kernels and the CPU test harness (csrc/host_emul.cpp -> tests/test_host_emulation.py checks that every split tiles [0, T)).
#pragma once
#include "tcq_bits.cuh"  // QP_HD

namespace qp {

// Split of T work items (super-tiles in flat [strip][column] order) into contiguous ranges, computed on the host.
//  * one level (split_range): unit u of `units` owns base (+1 for `rem` of them) items.  CTA-granular in the batched GEMM,
//    warp-granular in the dequantise kernels and -- with flat = 1, the default -- in the GEMV kernels.
//  * two levels (split_range_cta with flat = 0): the items are first dealt to the CTAs, then a CTA's range to its warps, so that
//    every CTA carries the same number of super-tiles (+-1); `flip` sends the +1 items to the LAST units so that the second part
//    of a two-rate layer does not hand its remainder to the same CTAs as the first.  Built because the flat split looked
//    unbalanced (at 4096x14336 tcomb_6_7 CTAs 0..10 get 18 super-tiles per warp, all others 16, and finish 0.8 us after them) --
//    and measured slower: the first CTAs are also the first to start (gemv_common.cuh, gemv_splits).
//  * skew: the units from w2 on (the highest indices) form a second class with a smaller share.  CTAs are dispatched in index
//    order, so when the preceding kernel still holds some SMs (a GEMV CTA takes a whole SM) it is exactly the LAST CTAs of the
//    grid that start late: after the 7-CTA SiLU cluster CTAs 141..147 of down_proj start 3.3 us after the others.  Worth
//    +-0.2 % in the decode step on top of the flat split; off by default (qp_xprod.late_ctas).
struct RunSplit {
    unsigned base, rem;          // class 1: units [0, w2)
    unsigned w2, base2, rem2;    // class 2: units [w2, units), first item off2
    unsigned off2, units, flip;
    unsigned flat;               // GEMV kernels: the units are the grid's warps (one level), not its CTAs
};
inline RunSplit make_split(long T, int units, bool flip = false) {
    return RunSplit{(unsigned)(T / units), (unsigned)(T % units), (unsigned)units, 0u, 0u, (unsigned)T, (unsigned)units, flip ? 1u : 0u, 0u};
}
// the last late_units units get `permille` / 1000 of the others' share
inline RunSplit make_split_skewed(long T, int units, int late_units, int permille, bool flip = false) {
    if (late_units <= 0 || late_units >= units || permille >= 1000 || permille < 0) return make_split(T, units, flip);
    const double wl = (double)late_units * permille / 1000.0, we = (double)(units - late_units);
    const long T2 = (long)((double)T * wl / (we + wl)), T1 = T - T2;
    const long n1 = units - late_units, n2 = late_units;
    return RunSplit{(unsigned)(T1 / n1), (unsigned)(T1 % n1), (unsigned)n1, (unsigned)(T2 / n2), (unsigned)(T2 % n2), (unsigned)T1,
                    (unsigned)units, flip ? 1u : 0u, 0u};
}
// unit v of n: offset and count when `base` items each, one more for `rem` of them (the first, or the last when flip)
QP_HD void split_unit(unsigned base, unsigned rem, unsigned n, unsigned v, unsigned flip, unsigned &lo, unsigned &cnt) {
    if (!flip) {
        lo = v * base + (v < rem ? v : rem);
        cnt = base + (v < rem ? 1u : 0u);
    } else {
        const unsigned first = n - rem;  // first unit with an extra item
        lo = v * base + (v > first ? v - first : 0u);
        cnt = base + (v >= first ? 1u : 0u);
    }
}
QP_HD void split_range(const RunSplit s, int u, unsigned &lo, unsigned &hi) {
    const unsigned uu = (unsigned)u;
    unsigned l, c;
    if (uu < s.w2) {
        split_unit(s.base, s.rem, s.w2, uu, s.flip, l, c);
    } else {
        split_unit(s.base2, s.rem2, s.units - s.w2, uu - s.w2, s.flip, l, c);
        l += s.off2;
    }
    lo = l;
    hi = l + c;
}
// two levels: range of warp `warp` (of `warps`) inside CTA `cta`'s range
QP_HD void split_range_cta(const RunSplit s, int cta, int warp, int warps, unsigned &lo, unsigned &hi) {
    if (s.flat) {
        split_range(s, cta * warps + warp, lo, hi);
        return;
    }
    unsigned clo, chi, l, c;
    split_range(s, cta, clo, chi);
    const unsigned n = chi - clo;
    split_unit(n / (unsigned)warps, n % (unsigned)warps, (unsigned)warps, (unsigned)warp, s.flip, l, c);
    lo = clo + l;
    hi = lo + c;
}

}  // namespace qp
