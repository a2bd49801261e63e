// Exact order-statistic selection from window counts: the decision made for one (nucleus,
// quantile) after every pass of the fused predictive kernel.  Plain C++ (host + device) so the
// same code is unit-tested on the CPU (tests/select_harness.cpp) and run by predict_select_kernel.
//
// np.percentile's "linear" method (pybmc/sampling_utils.py:80-82) needs the order statistics
// r = floor(v) and, when v is fractional, r + 1.  A pass over the draws of one nucleus reports, for
// the current window [lo, hi):  cb = #(x < lo),  cw = #(lo <= x < hi),  the split of cw over
// kSelSlices equal-width slices, and the first `cap` draws that fell inside.  From that:
//   * both ranks inside and cw <= cap  -> sort the stored draws, read the answer        (done)
//   * ranks outside                    -> move the window towards them (hard bracket if known,
//                                         else geometric growth)
//   * inside but cw > cap (overflow)   -> keep the slice(s) holding the ranks: exact counts, so
//                                         the window shrinks >= 2x per pass whatever the law
//                                         (atoms, gaps, heavy tails); a window one representable
//                                         value wide is an atom and is the answer
//   * the two ranks fall into different slices of an overflowing window (a gap between adjacent
//     order statistics with mass on both sides): look for r alone (phase 1), then r + 1 (phase 2)
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define BMC_HD __host__ __device__ __forceinline__
#else
#define BMC_HD inline
#endif

namespace bmc {

constexpr int kSelSlices = 32;
constexpr int kSelGrowth = 8;

template <typename real>
struct SelLimits;
template <>
struct SelLimits<float> {
    static BMC_HD float big() { return FLT_MAX; }
    static BMC_HD float up(float x) { return nextafterf(x, FLT_MAX); }
    static BMC_HD float down(float x) { return nextafterf(x, -FLT_MAX); }
};
template <>
struct SelLimits<double> {
    // windows are parked at FLT_MAX in both precisions (draws never get near it)
    static BMC_HD double big() { return static_cast<double>(FLT_MAX); }
    static BMC_HD double up(double x) { return nextafter(x, DBL_MAX); }
    static BMC_HD double down(double x) { return nextafter(x, -DBL_MAX); }
};

template <typename real>
struct SelState {
    real lo, hi;        // current window [lo, hi)
    real blo, bhi;      // hard bracket: the rank(s) being looked for lie in [blo, bhi)
    real pair_hi;       // upper bound that also holds for rank r + 1 (split mode)
    real aux;           // order statistic r once found (phase 2)
    int phase;          // 0 pair / single, 1 split: looking for r, 2 split: looking for r + 1
};

enum SelAction { kSelResolved = 0, kSelAgain = 1 };

// `first`, `second`: the in-window draws of ranks t1 - cb and t1 - cb + 1 (ascending order) when
// cw <= cap; `second` is only meaningful when that rank exists (t1 - cb + 1 < cw);
// `slices` may be null when the pass did not count them (the first pass does not: overflow is rare);
// `stored_equal` / `stored_value`: all `cap` stored draws of an overflowing window are equal.
template <typename real>
BMC_HD SelAction sel_decide(SelState<real>& st, long long r, bool need_pair, long long cb, long long cw,
                            const unsigned int* slices, int cap, real first, real second, bool stored_equal,
                            real stored_value, double* v0, double* v1) {
    using L = SelLimits<real>;
    const real big = L::big();
    const long long t1 = st.phase == 2 ? r + 1 : r;
    const bool want2 = st.phase == 0 && need_pair;
    long long t2 = t1 + (want2 ? 1 : 0);
    const real lo = st.lo, hi = st.hi;
    const real width = hi - lo;
    const bool first_below = t1 < cb, last_above = t2 >= cb + cw;

    if (!first_below && !last_above) {
        if (cw <= cap) {                                            // ---- read the answer
            if (st.phase == 0) {
                *v0 = static_cast<double>(first);
                *v1 = want2 ? static_cast<double>(second) : *v0;
                return kSelResolved;
            }
            if (st.phase == 2) {
                *v0 = static_cast<double>(st.aux);
                *v1 = static_cast<double>(first);
                return kSelResolved;
            }
            if (t1 - cb + 1 < cw) {                                 // phase 1 and the successor is here too
                *v0 = static_cast<double>(first);
                *v1 = static_cast<double>(second);
                return kSelResolved;
            }
            st.aux = first;                                         // successor lies at or above hi
            st.phase = 2;
            st.blo = hi;
            st.bhi = st.pair_hi;
            st.lo = hi;
            st.hi = st.bhi < big ? st.bhi : hi + static_cast<real>(kSelGrowth) * width;
            if (!(st.hi > st.lo)) st.hi = L::up(st.lo);
            return kSelAgain;
        }
        // ---- overflow: the ranks are inside, more draws than the buffer holds
        const real prev_bhi = st.bhi;
        st.blo = lo;
        st.bhi = hi;
        if (st.phase == 0) st.pair_hi = hi;                         // both ranks are below hi
        if (!(L::up(lo) < hi)) {                                    // one representable value: an atom at lo
            if (st.phase == 0 || (st.phase == 1 && t1 + 1 < cb + cw)) {
                *v0 = *v1 = static_cast<double>(lo);
                return kSelResolved;
            }
            if (st.phase == 2) {
                *v0 = static_cast<double>(st.aux);
                *v1 = static_cast<double>(lo);
                return kSelResolved;
            }
            st.aux = lo;                                            // phase 1: r is the last draw of the atom
            st.phase = 2;
            st.blo = hi;
            st.bhi = st.pair_hi > hi ? st.pair_hi : prev_bhi;
            if (!(st.bhi > hi)) st.bhi = big;
            st.lo = hi;
            st.hi = st.bhi < big ? st.bhi : hi + static_cast<real>(kSelGrowth) * (fabs(hi) * real(1e-3) + width);
            if (!(st.hi > st.lo)) st.hi = L::up(st.lo);
            return kSelAgain;
        }
        if (stored_equal && stored_value >= lo && stored_value < hi) {   // probable atom: test that value
            st.lo = stored_value;
            st.hi = L::up(stored_value);
            return kSelAgain;
        }
        if (slices == nullptr) return kSelAgain;                  // same window again, slices counted
        long long cum = cb;
        int b0 = -1, b1 = -1;
        for (int b = 0; b < kSelSlices; ++b) {
            const long long c = slices[b];
            if (b0 < 0 && t1 < cum + c) b0 = b;
            if (b1 < 0 && t2 < cum + c) b1 = b;
            cum += c;
        }
        if (b0 < 0) b0 = kSelSlices - 1;
        if (b1 < 0) b1 = kSelSlices - 1;
        if (want2 && b1 != b0) {                                    // gap between the two ranks: split
            st.phase = 1;
            t2 = t1;
            b1 = b0;
        }
        const real w = width / static_cast<real>(kSelSlices);
        real nlo = L::down(lo + w * static_cast<real>(b0));         // slice edges are recomputed here:
        real nhi = L::up(lo + w * static_cast<real>(b1 + 1));       // pad by one representable value
        if (nlo < lo) nlo = lo;
        if (nhi > hi) nhi = hi;
        if (nlo <= lo && nhi >= hi) {                               // slices too fine for the mantissa: halve
            long long lower = 0;
            for (int b = 0; b < kSelSlices / 2; ++b) lower += slices[b];
            real mid = lo + width * real(0.5);
            if (!(mid > lo)) mid = L::up(lo);
            if (t1 < cb + lower && t2 >= cb + lower) {              // the pair straddles the middle
                st.phase = 1;
                t2 = t1;
            }
            if (t2 < cb + lower) nhi = mid;
            else nlo = mid;
        }
        if (!(nhi > nlo)) nhi = L::up(nlo);
        st.lo = nlo;
        st.hi = nhi;
        return kSelAgain;
    }

    // ---- the ranks are (partly) outside the window
    real nlo = lo, nhi = hi;
    if (first_below) {
        if (t2 < cb) {                                              // all below
            st.bhi = lo;
            nhi = lo;
        }
        nlo = st.blo > -big ? st.blo : lo - static_cast<real>(kSelGrowth) * width;
    }
    if (last_above) {
        if (t1 >= cb + cw) {                                        // all above
            st.blo = hi;
            nlo = first_below ? nlo : hi;
        }
        nhi = st.bhi < big ? st.bhi : hi + static_cast<real>(kSelGrowth) * width;
    }
    if (!(nhi > nlo)) nhi = L::up(nlo);
    st.lo = nlo;
    st.hi = nhi;
    return kSelAgain;
}

// np.percentile "linear": a + (b - a) t, taken from the upper end when t >= 1/2 (numpy's _lerp)
BMC_HD double sel_lerp(double a, double b, double t) {
    const double diff = b - a;
    double res = a + diff * t;
    if (t >= 0.5) res = b - diff * (1.0 - t);
    if (diff == 0.0) res = a;
    return res;
}

}  // namespace bmc
