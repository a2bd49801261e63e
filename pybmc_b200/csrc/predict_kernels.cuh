// Fused posterior prediction + uncertainty quantification.
//
// Replaces pybmc/sampling_utils.py:60-82 (rndm_m_random_calculator: weights, S-by-N predictive
// matrix, noise, three np.percentile calls) and pybmc/sampling_utils.py:24-35 (coverage: 21*N sorts)
// without ever storing the S-by-N matrix:
//
//   x[s][n] = u[n] . beta[s] + sigma[s] * z[s][n]            (centred draw; rndm_m = mu[n] + x)
//     u[n]  = preds[n] . Vt_hat'   and   mu[n] = mean_m preds[n][m]
//     since  (beta Vt_hat + 1/M) . preds[n]  =  beta . u[n] + mu[n]      (:64-72)
//   z[s][n] = Philox(seed; s, n)  -- a pure function of (s, n), so any later pass regenerates it
//
// One pass over (s, n) accumulates, per nucleus: sum and sum of squares (mean, variance), the two
// integers #(x < t) and #(x <= t) that decide every coverage level of :30-33 without sorting, and
// for each requested quantile the count below a narrow window plus the few draws inside it.  A
// second, tiny kernel sorts each window's candidates and reads off the exact order statistics
// floor(v), ceil(v) with v = q/100 (S-1), interpolated as np.percentile does.  Windows come from a
// normal approximation; a miss or an overflow re-centres the window from the counts and repeats the
// pass for the affected nuclei only, so the result is exact for any distribution.
//
// Work layout: lane <-> nucleus, four posterior draws per step (one Philox call gives the four normals
// of a nucleus for draws 4i .. 4i+3).  K <= 16 (and fp64): the contraction is an FFMA chain over rows
// broadcast from a TMA-staged tile (predict_pass_kernel).  fp32 with K > 16: the contraction runs on
// the tensor cores (predict_tc_kernels.cuh) and the same per-lane consumer reads it back from TMEM.
#pragma once
#include <cfloat>
#include "rng.cuh"
#include "select_logic.h"
#include "tma.cuh"

namespace bmc {

constexpr int kPredTile = 256;        // posterior draws per staged tile (fewer for wide rows, see pred_tile)
constexpr int kPredWarps = 8;         // 256 nuclei per block
constexpr int kMaxQuant = 8;
constexpr int kMaxSlots = 64;         // sample splits (gridDim.y) a launch may use
constexpr int kSubBins = kSelSlices;  // slices of a window counted for the overflow fallback

// draws per shared-memory tile: two stages must stay under ~100 KB so that two blocks fit an SM
template <typename real, int KP>
struct PredTile {
    static constexpr int kRowBytes = (KP + 4) * static_cast<int>(sizeof(real));
    static constexpr int value = 2 * 256 * kRowBytes <= 100 * 1024 ? 256 : (2 * 128 * kRowBytes <= 100 * 1024 ? 128 : 64);
};

struct PredictArgs {
    // per-nucleus inputs (this launch's chunk; index 0 is global nucleus point0)
    const void* u;            // [n][k] real
    const double* mu;         // [n] nullable (0)
    const double* truth;      // [n] nullable
    long long n_points;       // nuclei in this chunk
    unsigned long long point0;  // global index of nucleus 0
    // posterior draws: rows [beta_0 .. beta_{k-1}, 0 .., sigma at column KP, 0 ..], row stride KP + 4
    const void* theta;        // [n_draws][KP + 4] real; nullable => x = z (matrix mode)
    long long n_draws;
    int k;
    // noise
    int noise_mode;           // 0 none, 1 philox, 2 external
    uint32_t key0, key1;
    PhiloxKeys keys;          // round keys of (key0, key1): constant-bank operands of the Philox rounds
    const void* noise;        // [n_draws][ld_noise] real
    long long ld_noise;
    // windows / results, indexed [n * nq + j]
    int nq;
    void* win_lo;             // real
    void* win_hi;             // real
    unsigned int* cnt_below;  // += #(x < lo)
    unsigned int* cnt_slot;   // [n*nq][kMaxSlots] in-window draws seen by each sample split
    unsigned int* sub_cnt;    // [n*nq][kSubBins] hits per equal-width slice (retry passes only)
    void* cand;               // [n*nq][cand_stride] real, split s owns [s*seg_len, (s+1)*seg_len)
    int cand_stride, seg_len;
    int count_slices;
    // first-pass accumulators
    int first;
    const void* center;       // [n] real: shift used for the moment sums
    double* mom_part;         // [s_splits][2][n]
    unsigned int* c_lt;       // [n]
    unsigned int* c_le;       // [n]
    double* draws_out;        // [n_draws][ld_out] nullable (materialise rndm_m)
    long long ld_out;
    // active list (retry passes): nuclei to process
    const int* point_list;    // nullable
    int n_active;             // nuclei to process
    int s_splits;             // gridDim.y
};

// One (draw, window) step: below += (x < lo);  hits |= bit when lo <= x < hi.  Written as two compares
// and two predicated integer ops (the compiler's select + add form costs one instruction more, and this
// runs 4 x NQ times per step).
template <typename real>
__device__ __forceinline__ void window_step(real x, real lo, real hi, unsigned int& below, unsigned int& hits,
                                            unsigned int bit);
template <>
__device__ __forceinline__ void window_step<float>(float x, float lo, float hi, unsigned int& below,
                                                   unsigned int& hits, unsigned int bit) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f32 p, %2, %3;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "setp.lt.and.f32 q, %2, %4, !p;\n\t"
        "@q or.b32 %1, %1, %5;\n\t}"
        : "+r"(below), "+r"(hits)
        : "f"(x), "f"(lo), "f"(hi), "r"(bit));
}
template <>
__device__ __forceinline__ void window_step<double>(double x, double lo, double hi, unsigned int& below,
                                                    unsigned int& hits, unsigned int bit) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f64 p, %2, %3;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "setp.lt.and.f64 q, %2, %4, !p;\n\t"
        "@q or.b32 %1, %1, %5;\n\t}"
        : "+r"(below), "+r"(hits)
        : "d"(x), "d"(lo), "d"(hi), "r"(bit));
}

// nlt += (x < t); nle += (x <= t): the two order counts of the coverage rule, predicated adds again
template <typename real>
__device__ __forceinline__ void count_step(real x, real t, unsigned int& nlt, unsigned int& nle);
template <>
__device__ __forceinline__ void count_step<float>(float x, float t, unsigned int& nlt, unsigned int& nle) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f32 p, %2, %3;\n\t"
        "setp.le.f32 q, %2, %3;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "@q add.u32 %1, %1, 1;\n\t}"
        : "+r"(nlt), "+r"(nle)
        : "f"(x), "f"(t));
}
template <>
__device__ __forceinline__ void count_step<double>(double x, double t, unsigned int& nlt, unsigned int& nle) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f64 p, %2, %3;\n\t"
        "setp.le.f64 q, %2, %3;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "@q add.u32 %1, %1, 1;\n\t}"
        : "+r"(nlt), "+r"(nle)
        : "d"(x), "d"(t));
}

// What one lane (= one nucleus) accumulates over its draws, and the step that folds four draws in.
// Shared by the FFMA pass kernel below and the tensor-core pass kernel so both count identically.
template <typename real, int NQ>
struct LaneAcc {
    real wlo[NQ], whi[NQ];
    unsigned int below[NQ], inwin[NQ];
    real sx, sxx;
    unsigned int nlt, nle;
};

struct LaneCtx {
    int n, idx0, slot;
    bool live;
    unsigned int seg;
    double mu_d;
};

template <typename real, int NQ>
__device__ __forceinline__ void lane_init(const PredictArgs& a, const LaneCtx& c, LaneAcc<real, NQ>& st) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const bool on = c.live && j < a.nq;
        st.wlo[j] = on ? static_cast<const real*>(a.win_lo)[c.idx0 + j] : real(FLT_MAX);
        st.whi[j] = on ? static_cast<const real*>(a.win_hi)[c.idx0 + j] : real(FLT_MAX);
        st.below[j] = 0u;
        st.inwin[j] = 0u;
    }
    st.sx = real(0);
    st.sxx = real(0);
    st.nlt = 0u;
    st.nle = 0u;
}

// x[0..3] are draws s .. s+3 of nucleus c.n (FLT_MAX = no such draw).
// FAST: the caller has checked once (per eight draws) that this is a first pass that does not materialise the
// draws -- a.first and a.draws_out are then compile-time facts instead of two uniform branches per four draws.
template <typename real, int NQ, bool FULL = false, bool FAST = false>
__device__ __forceinline__ void consume4(const PredictArgs& a, const LaneCtx& c, LaneAcc<real, NQ>& st,
                                         const real (&x)[4], long long s, real tc, real ctr) {
    using M = Math<real>;
    if (FAST || a.first) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool valid = FULL || x[r] < real(FLT_MAX);      // FULL: all four draws exist, no padding marks
            const real dv = valid ? x[r] - ctr : real(0);
            st.sx += dv;
            st.sxx = M::fma(dv, dv, st.sxx);
            count_step<real>(x[r], tc, st.nlt, st.nle);
            if constexpr (!FAST) {
                if (a.draws_out && valid && c.live)
                    a.draws_out[(s + r) * a.ld_out + c.n] = static_cast<double>(x[r]) + c.mu_d;
            }
        }
    }
    // windows: two compares and a predicated add per (draw, window); hits only set a bit (bit 4 j + r)
    unsigned int hits = 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            window_step<real>(x[r], st.wlo[j], st.whi[j], st.below[j], hits, 1u << (4 * j + r));
        }
    }
    while (hits) {                                  // rare per lane: store the draw, count it
        const int bit = __ffs(hits) - 1;
        hits &= hits - 1u;
        const int r = bit & 3, j = bit >> 2;
        const real xv = r == 0 ? x[0] : (r == 1 ? x[1] : (r == 2 ? x[2] : x[3]));
        unsigned int have = 0u;
#pragma unroll
        for (int jj = 0; jj < NQ; ++jj) {
            if (jj == j) {
                have = st.inwin[jj];
                st.inwin[jj] = have + 1u;
            }
        }
        if (have < c.seg)
            static_cast<real*>(a.cand)[static_cast<long long>(c.idx0 + j) * a.cand_stride + c.slot * c.seg + have] = xv;
        if (a.count_slices) {
            // which 1/32 slice of the window: lets an overflowing window be narrowed with exact
            // counts whatever the distribution (atoms, gaps, heavy tails)
            real lo_j = real(0), hi_j = real(1);
#pragma unroll
            for (int jj = 0; jj < NQ; ++jj) {
                if (jj == j) {
                    lo_j = st.wlo[jj];
                    hi_j = st.whi[jj];
                }
            }
            const real rel = (xv - lo_j) * (real(kSubBins) / (hi_j - lo_j));
            int bin = static_cast<int>(rel);
            bin = bin < 0 ? 0 : (bin > kSubBins - 1 ? kSubBins - 1 : bin);
            atomicAdd(a.sub_cnt + static_cast<long long>(c.idx0 + j) * kSubBins + bin, 1u);
        }
    }
}

// what a lane leaves behind for the select kernel
template <typename real, int NQ>
__device__ __forceinline__ void lane_flush(const PredictArgs& a, const LaneCtx& c, const LaneAcc<real, NQ>& st) {
    if (!c.live) return;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        if (j < a.nq) {
            if (st.below[j]) atomicAdd(a.cnt_below + c.idx0 + j, st.below[j]);
            a.cnt_slot[static_cast<long long>(c.idx0 + j) * kMaxSlots + c.slot] = st.inwin[j];
        }
    }
    if (a.first) {
        a.mom_part[(static_cast<long long>(c.slot) * 2 + 0) * a.n_points + c.n] = static_cast<double>(st.sx);
        a.mom_part[(static_cast<long long>(c.slot) * 2 + 1) * a.n_points + c.n] = static_cast<double>(st.sxx);
        if (a.truth) {
            atomicAdd(a.c_lt + c.n, st.nlt);
            atomicAdd(a.c_le + c.n, st.nle);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// The pass kernel: lane <-> nucleus, four posterior draws per step (one Philox call = the four
// normals of this nucleus for draws 4i .. 4i+3), the draws' rows broadcast from a TMA-staged tile.
// Everything a lane accumulates is private to it, so the hot loop has no atomics, votes or branches.
template <typename real, int KP, int NQ>
__global__ void __launch_bounds__(kPredWarps * 32) predict_pass_kernel(const PredictArgs a) {
    using M = Math<real>;
    constexpr int LDT = KP + 4;
    constexpr int TILE = PredTile<real, KP>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    real* const tile0 = reinterpret_cast<real*>(smem_raw);
    real* const tile1 = tile0 + TILE * LDT;
    __shared__ __align__(8) uint64_t bars[2];

    const int slot = blockIdx.y;
    const int pslot = blockIdx.x * (kPredWarps * 32) + threadIdx.x;   // index into the active list
    const bool live = pslot < a.n_active;
    const int n = live ? (a.point_list ? a.point_list[pslot] : pslot) : 0;

    // sample range of this block: whole tiles, so every block's range starts on a multiple of 4
    const long long per = ((a.n_draws + a.s_splits - 1) / a.s_splits + TILE - 1) / TILE * TILE;
    const long long s_begin = static_cast<long long>(slot) * per;
    const long long s_end = min(a.n_draws, s_begin + per);
    const int n_tiles = s_end > s_begin ? static_cast<int>((s_end - s_begin + TILE - 1) / TILE) : 0;
    const bool use_theta = a.theta != nullptr;
    const real* theta = static_cast<const real*>(a.theta);
    const bool tma_ok = use_theta && (reinterpret_cast<uintptr_t>(theta) & 15) == 0;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_tile = [&](int t) {
        real* dst = (t & 1) ? tile1 : tile0;
        const long long s0 = s_begin + static_cast<long long>(t) * TILE;
        const int cnt = static_cast<int>(min(static_cast<long long>(TILE), s_end - s0));
        if (tma_ok) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = static_cast<uint32_t>(cnt * LDT * sizeof(real));   // LDT*sizeof % 16 == 0
                mbar_expect_tx(&bars[t & 1], bytes);
                tma_load_1d(dst, theta + s0 * LDT, bytes, &bars[t & 1]);
            }
        } else if (use_theta) {
            for (int i = threadIdx.x; i < cnt * LDT; i += blockDim.x) dst[i] = theta[s0 * LDT + i];
        }
    };

    // per-lane constants
    real u[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k)
        u[k] = (live && k < a.k) ? static_cast<const real*>(a.u)[static_cast<long long>(n) * a.k + k] : real(0);
    LaneCtx c;
    c.n = n;
    c.idx0 = n * a.nq;
    c.slot = slot;
    c.live = live;
    c.seg = static_cast<unsigned int>(a.seg_len);
    c.mu_d = (live && a.mu) ? a.mu[n] : 0.0;
    const real tc = (live && a.truth) ? static_cast<real>(a.truth[n] - c.mu_d) : real(0);
    const real ctr = (live && a.center) ? static_cast<const real*>(a.center)[n] : real(0);
    LaneAcc<real, NQ> st;
    lane_init<real, NQ>(a, c, st);
    const uint32_t nglob = static_cast<uint32_t>(a.point0 + static_cast<unsigned long long>(n));

    if (n_tiles > 0) issue_tile(0);
    if (!tma_ok) __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_tile(t + 1);           // the other stage was released by the barrier below
        if (tma_ok) mbar_wait(&bars[t & 1], (t >> 1) & 1);
        const real* tile = (t & 1) ? tile1 : tile0;
        const long long s0 = s_begin + static_cast<long long>(t) * TILE;
        const int cnt = static_cast<int>(min(static_cast<long long>(TILE), s_end - s0));
        for (int g = 0; g < cnt; g += 4) {
            real x[4] = {real(0), real(0), real(0), real(0)};
            real sigma[4] = {real(1), real(1), real(1), real(1)};
            if (use_theta) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const real* row = tile + (g + r) * LDT;          // warp-uniform address: broadcast
#pragma unroll
                    for (int k = 0; k < KP; ++k) x[r] = M::fma(u[k], row[k], x[r]);
                    sigma[r] = row[KP];
                }
            }
            const long long s = s0 + g;
            if (a.noise_mode == 1) {
                real z[4];
                normals4_k<real>(static_cast<uint32_t>(s >> 2), nglob, 0u, kTagNoise, a.keys, z);
#pragma unroll
                for (int r = 0; r < 4; ++r) x[r] = M::fma(sigma[r], z[r], x[r]);
            } else if (a.noise_mode == 2 && live) {
                const real* zc = static_cast<const real*>(a.noise) + n;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (g + r < cnt) x[r] = M::fma(sigma[r], zc[(s + r) * a.ld_noise], x[r]);
            }
            if (g + 4 > cnt) {
                // ragged end of the draws: park the missing ones where no window or truth can see them
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (g + r >= cnt) x[r] = real(FLT_MAX);
            }
            consume4<real, NQ>(a, c, st, x, s, tc, ctr);
        }
        __syncthreads();                                   // everyone is done with this stage
    }

    lane_flush<real, NQ>(a, c, st);
}

// ----------------------------------------------------------------------------------------------
struct SelectArgs {
    long long n_points;
    long long n_draws;
    int nq;
    const long long* rank;     // [nq] lower order-statistic index floor(v)
    const double* frac;        // [nq] v - floor(v)
    void* win_lo;
    void* win_hi;
    void* brk_lo;              // hard bracket known to contain the target
    void* brk_hi;
    void* pair_hi;             // split-mode state (select_logic.h)
    void* aux;
    unsigned char* phase;
    unsigned int* cnt_below;
    unsigned int* cnt_slot;    // [n*nq][kMaxSlots]
    unsigned int* sub_cnt;
    int slices_valid;          // the pass that just ran counted slices
    void* cand;
    int cand_stride, seg_len, n_slots;
    unsigned char* resolved;   // [n*nq]
    const double* mu;          // nullable
    double* out_quant;         // [nq][ld_quant]
    long long ld_quant;
    long long out_offset;      // chunk offset into the output arrays
    // moments (first select only)
    int first;
    const double* mom_part;
    const void* center;
    double* out_mean;
    double* out_var;
    const unsigned int* c_lt;
    const unsigned int* c_le;
    long long* out_c_lt;
    long long* out_c_le;
    // retry bookkeeping
    const int* point_list;     // nuclei examined by this launch (nullable = all)
    int n_active;
    int* point_flag;           // [n_points] set when the nucleus needs another pass
    int* next_list;
    int* next_count;
};

// order-preserving unsigned keys of floating-point values (for the radix select)
template <typename real>
struct RadixKey;
template <>
struct RadixKey<float> {
    using type = uint32_t;
    static constexpr int kBytes = 4;
    static __device__ __forceinline__ type encode(float x) {
        const uint32_t b = __float_as_uint(x);
        return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    }
    static __device__ __forceinline__ float decode(type k) {
        return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
    }
};
template <>
struct RadixKey<double> {
    using type = unsigned long long;
    static constexpr int kBytes = 8;
    static __device__ __forceinline__ type encode(double x) {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
        return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double decode(type k) {
        return __longlong_as_double(static_cast<long long>(k ^ ((k >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull)));
    }
};

// Warp-cooperative exact selection: the draws of ranks `want` and `want + 1` (0-based, ascending)
// among the candidates one (nucleus, quantile) collected, spread over `n_slots` segments.  MSB-first
// radix select on order-preserving keys: one 256-bin shared-memory histogram per byte, then one more
// sweep for the successor.  Reads the candidates straight from global memory (L1/L2 resident).
template <typename real>
__device__ void warp_radix_select(const real* src, const unsigned int* slot_cnt, int n_slots, int seg_len,
                                  long long want, unsigned int* hist, real* first, real* second) {
    using RK = RadixKey<real>;
    using key_t = typename RK::type;
    const int lane = threadIdx.x & 31;
    key_t prefix = 0;
    long long k = want;
    unsigned int n_equal = 0;
    for (int byte = RK::kBytes - 1; byte >= 0; --byte) {
        const int shift = 8 * byte;
        for (int b = lane; b < 256; b += 32) hist[b] = 0u;
        __syncwarp();
        for (int sl = 0; sl < n_slots; ++sl) {
            const int c = static_cast<int>(slot_cnt[sl]);
            const real* seg = src + static_cast<long long>(sl) * seg_len;
            for (int i = lane; i < c; i += 32) {
                const key_t key = RK::encode(seg[i]);
                const bool match = byte == RK::kBytes - 1 || (key >> (shift + 8)) == (prefix >> (shift + 8));
                if (match) atomicAdd(hist + static_cast<unsigned int>((key >> shift) & 0xff), 1u);
            }
        }
        __syncwarp();
        // lane l owns bins 8l .. 8l+7; find the bin where the running count passes k
        unsigned int mine[8], sum = 0u;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            mine[b] = hist[8 * lane + b];
            sum += mine[b];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned int excl = incl - sum;
        const bool owner = static_cast<long long>(excl) <= k && k < static_cast<long long>(incl);
        const unsigned int who = __ballot_sync(0xffffffffu, owner);
        const int src_lane = who ? __ffs(who) - 1 : 31;
        int bin = 0;
        unsigned int before = excl, in_bin = 0u;
        if (lane == src_lane) {
            unsigned int run = excl;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (static_cast<long long>(run) <= k && k < static_cast<long long>(run + mine[b])) {
                    bin = 8 * lane + b;
                    before = run;
                    in_bin = mine[b];
                }
                run += mine[b];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, src_lane);
        before = __shfl_sync(0xffffffffu, before, src_lane);
        in_bin = __shfl_sync(0xffffffffu, in_bin, src_lane);
        prefix |= static_cast<key_t>(bin) << shift;
        k -= before;
        n_equal = in_bin;
        __syncwarp();
    }
    *first = RK::decode(prefix);
    // successor: the same value if more copies remain, else the smallest larger key
    if (k + 1 < static_cast<long long>(n_equal)) {
        *second = *first;
        return;
    }
    key_t best = ~static_cast<key_t>(0);
    for (int sl = 0; sl < n_slots; ++sl) {
        const int c = static_cast<int>(slot_cnt[sl]);
        const real* seg = src + static_cast<long long>(sl) * seg_len;
        for (int i = lane; i < c; i += 32) {
            const key_t key = RK::encode(seg[i]);
            if (key > prefix && key < best) best = key;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const key_t other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    *second = RK::decode(best);
}

// one warp per (nucleus, quantile): exact selection among the window's candidates when they were
// all stored, then sel_decide (select_logic.h) reads the answer or chooses the next window
constexpr int kSelStageBytes = 8192;   // per warp: candidates gathered from their slots before the radix passes
template <typename real>
__global__ void __launch_bounds__(256) predict_select_kernel(const SelectArgs a) {
    __shared__ unsigned int hist_all[8][256];
    __shared__ unsigned int staged_cnt[8];
    extern __shared__ __align__(16) unsigned char select_smem[];
    constexpr int kStageCap = kSelStageBytes / static_cast<int>(sizeof(real));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* const hist = hist_all[warp];
    real* const stage = reinterpret_cast<real*>(select_smem) + warp * kStageCap;
    const long long item = static_cast<long long>(blockIdx.x) * 8 + warp;       // (active slot, j)
    const long long pslot = item / a.nq;
    if (pslot >= a.n_active) return;
    const int n = a.point_list ? a.point_list[pslot] : static_cast<int>(pslot);
    const int j = static_cast<int>(item % a.nq);
    const long long idx = static_cast<long long>(n) * a.nq + j;

    if (a.first && j == 0 && lane == 0) {
        double s1 = 0.0, s2 = 0.0;
        for (int s = 0; s < a.n_slots; ++s) {
            s1 += a.mom_part[(static_cast<long long>(s) * 2 + 0) * a.n_points + n];
            s2 += a.mom_part[(static_cast<long long>(s) * 2 + 1) * a.n_points + n];
        }
        const double cnt = static_cast<double>(a.n_draws);
        const double dm = s1 / cnt;
        const double c = a.center ? static_cast<double>(static_cast<const real*>(a.center)[n]) : 0.0;
        a.out_mean[a.out_offset + n] = (a.mu ? a.mu[n] : 0.0) + c + dm;
        a.out_var[a.out_offset + n] = fmax(s2 / cnt - dm * dm, 0.0);     // population variance (ddof = 0)
        if (a.out_c_lt) {
            a.out_c_lt[a.out_offset + n] = a.c_lt[n];
            a.out_c_le[a.out_offset + n] = a.c_le[n];
        }
    }
    if (a.resolved[idx]) return;

    SelState<real> st;
    st.lo = static_cast<real*>(a.win_lo)[idx];
    st.hi = static_cast<real*>(a.win_hi)[idx];
    st.blo = static_cast<real*>(a.brk_lo)[idx];
    st.bhi = static_cast<real*>(a.brk_hi)[idx];
    st.pair_hi = static_cast<real*>(a.pair_hi)[idx];
    st.aux = static_cast<real*>(a.aux)[idx];
    st.phase = a.phase[idx];
    const long long r = a.rank[j];
    const bool need_pair = a.frac[j] > 0.0;
    const long long cb = a.cnt_below[idx];
    const unsigned int* slot_cnt = a.cnt_slot + idx * kMaxSlots;
    long long cw = 0;
    bool storable = true;
    int first_slot = -1;
    for (int s = 0; s < a.n_slots; ++s) {
        const unsigned int c = slot_cnt[s];
        cw += c;
        storable = storable && c <= static_cast<unsigned int>(a.seg_len);
        if (first_slot < 0 && c > 0) first_slot = s;
    }
    const long long t1 = st.phase == 2 ? r + 1 : r;
    const long long t2 = t1 + ((st.phase == 0 && need_pair) ? 1 : 0);
    const bool inside = t1 >= cb && t2 < cb + cw;
    const real* src = static_cast<const real*>(a.cand) + idx * a.cand_stride;
    // every in-window draw was stored <=> the answer can be read; otherwise it is an overflow
    const int cap_eff = storable ? 0x7fffffff : 0;

    bool stored_equal = false;
    real stored_value = real(0), v_first = real(0), v_second = real(0);
    if (inside && storable && cw <= kStageCap) {
        // gather the slots' candidates once; the radix passes then run out of shared memory
        // four slots at a time, up to four values per lane and slot: sixteen independent loads in flight
        // instead of one slot's at a time (this gather is latency-bound: 33 % long-scoreboard stalls)
        int off = 0;
        for (int sl = 0; sl < a.n_slots; sl += 4) {
            int c[4];
            real v[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = sl + q < a.n_slots ? static_cast<int>(slot_cnt[sl + q]) : 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const real* seg = src + static_cast<long long>(sl + q) * a.seg_len;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = lane + 32 * t;
                    v[q][t] = i < c[q] ? seg[i] : real(0);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const real* seg = src + static_cast<long long>(sl + q) * a.seg_len;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = lane + 32 * t;
                    if (i < c[q]) stage[off + i] = v[q][t];
                }
                for (int i = lane + 128; i < c[q]; i += 32) stage[off + i] = seg[i];
                off += c[q];
            }
        }
        if (lane == 0) staged_cnt[warp] = static_cast<unsigned int>(off);
        __syncwarp();
        warp_radix_select<real>(stage, staged_cnt + warp, 1, kStageCap, t1 - cb, hist, &v_first, &v_second);
    } else if (inside && storable) {
        warp_radix_select<real>(src, slot_cnt, a.n_slots, a.seg_len, t1 - cb, hist, &v_first, &v_second);
    } else if (inside && first_slot >= 0) {
        const int c = min(static_cast<int>(slot_cnt[first_slot]), a.seg_len);
        real vmin = real(FLT_MAX), vmax = -real(FLT_MAX);
        for (int i = lane; i < c; i += 32) {
            const real v = src[first_slot * a.seg_len + i];
            vmin = fmin(vmin, v);
            vmax = fmax(vmax, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        stored_equal = c >= 8 && vmin == vmax;
        stored_value = vmin;
    }
    if (lane != 0) return;
    unsigned int* sub = a.sub_cnt + idx * kSubBins;
    double v0 = 0.0, v1 = 0.0;
    const SelAction act = sel_decide<real>(st, r, need_pair, cb, cw, a.slices_valid ? sub : nullptr, cap_eff, v_first,
                                           v_second, stored_equal, stored_value, &v0, &v1);
    if (act == kSelResolved) {
        a.out_quant[static_cast<long long>(j) * a.ld_quant + a.out_offset + n] =
            (a.mu ? a.mu[n] : 0.0) + sel_lerp(v0, v1, a.frac[j]);
        a.resolved[idx] = 1;
        // park the window so later passes over this nucleus collect nothing for it
        static_cast<real*>(a.win_lo)[idx] = real(FLT_MAX);
        static_cast<real*>(a.win_hi)[idx] = real(FLT_MAX);
        return;
    }
    static_cast<real*>(a.win_lo)[idx] = st.lo;
    static_cast<real*>(a.win_hi)[idx] = st.hi;
    static_cast<real*>(a.brk_lo)[idx] = st.blo;
    static_cast<real*>(a.brk_hi)[idx] = st.bhi;
    static_cast<real*>(a.pair_hi)[idx] = st.pair_hi;
    static_cast<real*>(a.aux)[idx] = st.aux;
    a.phase[idx] = static_cast<unsigned char>(st.phase);
    a.cnt_below[idx] = 0u;
    for (int b = 0; b < kSubBins; ++b) sub[b] = 0u;
    if (atomicExch(a.point_flag + n, 1) == 0) {
        const int pos = atomicAdd(a.next_count, 1);
        a.next_list[pos] = n;
    }
}

// ----------------------------------------------------------------------------------------------
// first guess of centre and spread per nucleus from the posterior sample moments:
//   E x = u.E[beta],  Var x = u' Cov[beta] u + noise * E[sigma^2]
// Eight threads per nucleus: thread q takes rows r = q, q + 8, ... of the quadratic form; the covariance
// and the block's u rows are staged in shared memory (row stride k + 2: conflict-free).  Block = 256
// threads = 32 nuclei; dynamic shared memory = ((k + 1) (k + 2) + 32 (k + 2)) doubles.
constexpr int kGuessSub = 8;
template <typename real>
__global__ void __launch_bounds__(256) predict_guess_kernel(const void* u_, long long n, int k,
                                                            const double* theta_mean,
                                                            const double* theta_cov /* [(k+1)^2] */, int noise_on,
                                                            void* center_, void* scale_) {
    extern __shared__ double guess_smem[];
    const int d = k + 1, ld = k + 2;
    double* const cov = guess_smem;                       // [d][ld]
    double* const us = guess_smem + d * ld;               // [32][ld]
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) cov[(i / d) * ld + i % d] = theta_cov[i];
    const long long p0 = static_cast<long long>(blockIdx.x) * 32;
    const real* u = static_cast<const real*>(u_);
    for (int i = threadIdx.x; i < 32 * k; i += blockDim.x) {
        const long long p = p0 + i / k;
        us[(i / k) * ld + i % k] = p < n ? static_cast<double>(u[p * k + i % k]) : 0.0;
    }
    __syncthreads();
    const int local = threadIdx.x / kGuessSub, q = threadIdx.x % kGuessSub;
    const double* ur = us + local * ld;
    double m = 0.0, v = 0.0;
    for (int r = q; r < k; r += kGuessSub) {
        const double x = ur[r];
        m = fma(x, theta_mean[r], m);
        double t = 0.0;
        for (int c = 0; c < k; ++c) t = fma(cov[r * ld + c], ur[c], t);
        v = fma(x, t, v);
    }
#pragma unroll
    for (int o = kGuessSub / 2; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    const long long i = p0 + local;
    if (q != 0 || i >= n) return;
    if (noise_on) v += cov[k * ld + k] + theta_mean[k] * theta_mean[k];
    static_cast<real*>(center_)[i] = static_cast<real>(m);
    static_cast<real*>(scale_)[i] = static_cast<real>(sqrt(fmax(v, 0.0)));
}

// windows from centre/scale:  [c + s (z_q - h_q), c + s (z_q + h_q))
template <typename real>
__global__ void predict_window_kernel(long long n, int nq, const void* center_, const void* scale_,
                                      const double* zq, const double* hw, void* win_lo, void* win_hi,
                                      void* brk_lo, void* brk_hi, void* pair_hi, void* aux, unsigned char* phase, unsigned int* cnt_below,
                                      unsigned int* sub_cnt, unsigned char* resolved) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * nq) return;
    const long long p = i / nq;
    const int j = static_cast<int>(i % nq);
    const double c = static_cast<double>(static_cast<const real*>(center_)[p]);
    double s = static_cast<double>(static_cast<const real*>(scale_)[p]);
    const double tiny = (sizeof(real) == 4 ? 1e-6 : 1e-14) * (fabs(c) + 1e-30);
    if (!(s > tiny)) s = tiny;
    real lo = static_cast<real>(c + s * (zq[j] - hw[j]));
    real hi = static_cast<real>(c + s * (zq[j] + hw[j]));
    if (!(hi > lo)) hi = SelLimits<real>::up(lo);
    static_cast<real*>(win_lo)[i] = lo;
    static_cast<real*>(win_hi)[i] = hi;
    static_cast<real*>(brk_lo)[i] = -real(FLT_MAX);
    static_cast<real*>(brk_hi)[i] = real(FLT_MAX);
    static_cast<real*>(pair_hi)[i] = real(FLT_MAX);
    static_cast<real*>(aux)[i] = real(0);
    phase[i] = 0;
    cnt_below[i] = 0u;
    for (int b = 0; b < kSubBins; ++b) sub_cnt[i * kSubBins + b] = 0u;
    resolved[i] = 0;
}

// ----------------------------------------------------------------------------------------------
// Streaming kernels on a materialised S-by-N matrix (the reference's rndm_m).
// #(x < t), #(x <= t) per column: pybmc/sampling_utils.py:28-33 without the sort.  HBM-bound.
__global__ void __launch_bounds__(256) coverage_counts_kernel(const double* __restrict__ mat, long long s_rows,
                                                              long long n_cols, long long ld,
                                                              const double* __restrict__ truth,
                                                              long long rows_per_block,
                                                              unsigned long long* __restrict__ c_lt,
                                                              unsigned long long* __restrict__ c_le) {
    const long long col = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (col >= n_cols) return;
    const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
    const long long r1 = min(s_rows, r0 + rows_per_block);
    const double t = truth[col];
    unsigned int lt = 0u, le = 0u;
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {
        const double v0 = mat[r * ld + col], v1 = mat[(r + 1) * ld + col];
        const double v2 = mat[(r + 2) * ld + col], v3 = mat[(r + 3) * ld + col];
        lt += (v0 < t) + (v1 < t) + (v2 < t) + (v3 < t);
        le += (v0 <= t) + (v1 <= t) + (v2 <= t) + (v3 <= t);
    }
    for (; r < r1; ++r) {
        const double v = mat[r * ld + col];
        lt += v < t;
        le += v <= t;
    }
    atomicAdd(c_lt + col, static_cast<unsigned long long>(lt));
    atomicAdd(c_le + col, static_cast<unsigned long long>(le));
}

// column mean and standard deviation (two-stage, shifted by the first row) for the window guess
__global__ void __launch_bounds__(256) column_moments_kernel(const double* __restrict__ mat, long long s_rows,
                                                             long long n_cols, long long ld, double* center,
                                                             double* scale) {
    const long long col = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (col >= n_cols) return;
    const double shift = mat[col];
    double s1 = 0.0, s2 = 0.0;
    for (long long r = 0; r < s_rows; ++r) {
        const double d = mat[r * ld + col] - shift;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    const double m = s1 / static_cast<double>(s_rows);
    center[col] = shift + m;
    scale[col] = sqrt(fmax(s2 / static_cast<double>(s_rows) - m * m, 0.0));
}

// covered[l] = #{ n : c_le[n] >= lo[l] + 1  and  c_lt[n] <= hi[l] }      (sampling_utils.py:30-33)
__global__ void coverage_levels_kernel(const long long* __restrict__ c_lt, const long long* __restrict__ c_le,
                                       long long n, const long long* __restrict__ lo_idx,
                                       const long long* __restrict__ hi_idx, int n_levels,
                                       unsigned long long* __restrict__ covered) {
    const int level = blockIdx.y;
    if (level >= n_levels) return;
    const long long lo = lo_idx[level], hi = hi_idx[level];
    unsigned int cnt = 0u;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        cnt += (c_le[i] >= lo + 1 && c_lt[i] <= hi) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(covered + level, static_cast<unsigned long long>(cnt));
}

}  // namespace bmc
