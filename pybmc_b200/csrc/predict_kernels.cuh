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
// Work layout: lane <-> posterior draw, warp <-> quad of 4 consecutive nuclei (one Philox call gives
// the quad's four normals), block = 8 warps sharing one TMA-staged tile of the transposed sample
// matrix thetaT[K+1][S] in shared memory.
#pragma once
#include <cfloat>
#include "rng.cuh"
#include "select_logic.h"
#include "tma.cuh"

namespace bmc {

constexpr int kPredTile = 256;        // posterior draws per staged tile
constexpr int kPredWarps = 8;
constexpr int kMaxQuant = 8;
constexpr int kSubBins = kSelSlices;  // slices of a window counted for the overflow fallback

struct PredictArgs {
    // per-nucleus inputs (this launch's chunk; index 0 is global nucleus point0)
    const void* u;            // [n][k] real
    const double* mu;         // [n] nullable (0)
    const double* truth;      // [n] nullable
    long long n_points;       // nuclei in this chunk
    unsigned long long point0;  // global index of nucleus 0 (multiple of 4)
    // posterior draws
    const void* theta_t;      // [k+1][n_draws] real; nullable => x = z (matrix mode)
    long long n_draws;
    int k;
    // noise
    int noise_mode;           // 0 none, 1 philox, 2 external
    uint32_t key0, key1;
    const void* noise;        // [n_draws][ld_noise] real
    long long ld_noise;
    // windows / results, all indexed [n * nq + j]
    int nq;
    void* win_lo;             // real
    void* win_hi;             // real
    unsigned int* cnt_below;
    unsigned int* cnt_in;
    unsigned int* sub_cnt;    // [n*nq][kSubBins] hits per equal-width slice of the window
    void* cand;               // [n*nq][cand_cap] real
    int cand_cap;
    // first-pass accumulators
    int first;
    const void* center;       // [n] real: shift used for the moment sums
    double* mom_part;         // [n_slots][2][n]
    unsigned int* c_lt;       // [n]
    unsigned int* c_le;       // [n]
    double* draws_out;        // [n_draws][ld_out] nullable (materialise rndm_m)
    long long ld_out;
    // active list (retry passes)
    const int* quad_list;     // nullable
    int n_quads;              // quads to process
    int warps_per_quad;       // 1, 2, 4 or 8
    int s_splits;             // gridDim.y
};

// ----------------------------------------------------------------------------------------------
// The pass kernel.  KP: compile-time bound on k (u kept in registers); NQ: bound on nq.
template <typename real, int KP, int NQ>
__global__ void __launch_bounds__(kPredWarps * 32) predict_pass_kernel(const PredictArgs a) {
    using M = Math<real>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // two stages of thetaT tile: [k+1][kPredTile]
    real* const tile0 = reinterpret_cast<real*>(smem_raw);
    const int rows = a.k + 1;
    real* const tile1 = tile0 + rows * kPredTile;
    __shared__ __align__(8) uint64_t bars[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpq = a.warps_per_quad;
    const int qpb = kPredWarps / wpq;                       // quads per block
    const int slice = warp % wpq;
    const int qslot = blockIdx.x * qpb + warp / wpq;        // index into the active list
    const bool has_quad = qslot < a.n_quads;
    const int quad = has_quad ? (a.quad_list ? a.quad_list[qslot] : qslot) : 0;
    const long long n0 = 4ll * quad;

    // sample range of this block
    const long long per = ((a.n_draws + a.s_splits - 1) / a.s_splits + kPredTile - 1) / kPredTile * kPredTile;
    const long long s_begin = static_cast<long long>(blockIdx.y) * per;
    const long long s_end = min(a.n_draws, s_begin + per);
    const int n_tiles = s_end > s_begin ? static_cast<int>((s_end - s_begin + kPredTile - 1) / kPredTile) : 0;
    const bool use_theta = a.theta_t != nullptr;
    const real* theta_t = static_cast<const real*>(a.theta_t);
    // TMA needs 16-byte aligned rows: n_draws % (16/sizeof(real)) == 0, else plain loads
    const bool tma_ok = use_theta && (a.n_draws % (16 / sizeof(real)) == 0) &&
                        ((reinterpret_cast<uintptr_t>(theta_t) & 15) == 0);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_tile = [&](int t) {
        // called by thread 0 (tma) or by all threads (fallback)
        real* dst = (t & 1) ? tile1 : tile0;
        const long long s0 = s_begin + static_cast<long long>(t) * kPredTile;
        const int cnt = static_cast<int>(min(static_cast<long long>(kPredTile), s_end - s0));
        if (tma_ok) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = static_cast<uint32_t>(((cnt * sizeof(real)) + 15) & ~15u);
                mbar_expect_tx(&bars[t & 1], bytes * rows);
                for (int r = 0; r < rows; ++r)
                    tma_load_1d(dst + r * kPredTile, theta_t + r * a.n_draws + s0, bytes, &bars[t & 1]);
            }
        } else if (use_theta) {
            for (int i = threadIdx.x; i < rows * kPredTile; i += blockDim.x) {
                const int r = i / kPredTile, c = i % kPredTile;
                dst[i] = c < cnt ? theta_t[r * a.n_draws + s0 + c] : real(0);
            }
        }
    };

    // per-warp constants
    real u[4][KP];
    real tc[4], ctr[4], muv[4];
    bool live[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long n = n0 + q;
        live[q] = has_quad && n < a.n_points;
        const long long nn = live[q] ? n : 0;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            u[q][k] = (live[q] && k < a.k) ? static_cast<const real*>(a.u)[nn * a.k + k] : real(0);
        const double m = a.mu ? a.mu[nn] : 0.0;
        muv[q] = static_cast<real>(m);
        tc[q] = a.truth ? static_cast<real>(a.truth[nn] - m) : real(0);
        ctr[q] = a.center ? static_cast<const real*>(a.center)[nn] : real(0);
    }
    real wlo[4][NQ], whi[4][NQ];
    unsigned int below[4][NQ];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const bool on = live[q] && j < a.nq;
            const long long idx = (n0 + q) * a.nq + j;
            wlo[q][j] = on ? static_cast<const real*>(a.win_lo)[idx] : real(FLT_MAX);
            whi[q][j] = on ? static_cast<const real*>(a.win_hi)[idx] : real(FLT_MAX);
            below[q][j] = 0u;
        }
    real sx[4] = {0, 0, 0, 0}, sxx[4] = {0, 0, 0, 0};
    unsigned int nlt[4] = {0, 0, 0, 0}, nle[4] = {0, 0, 0, 0};
    const uint32_t qglob = static_cast<uint32_t>((a.point0 >> 2) + static_cast<unsigned long long>(quad));

    if (n_tiles > 0) issue_tile(0);
    if (!tma_ok) __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_tile(t + 1);           // the other stage was released by the barrier below
        if (tma_ok) mbar_wait(&bars[t & 1], (t >> 1) & 1);
        const real* tile = (t & 1) ? tile1 : tile0;
        const long long s0 = s_begin + static_cast<long long>(t) * kPredTile;
        if (has_quad) {
            for (int sub = slice; sub < kPredTile / 32; sub += wpq) {
                const int sl = sub * 32 + lane;
                const long long s = s0 + sl;
                const bool valid = s < s_end;
                real x[4] = {0, 0, 0, 0};
                real sigma = real(1);
                if (use_theta) {
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        if (k < a.k) {
                            const real b = tile[k * kPredTile + sl];
#pragma unroll
                            for (int q = 0; q < 4; ++q) x[q] = M::fma(u[q][k], b, x[q]);
                        }
                    }
                    sigma = tile[a.k * kPredTile + sl];
                }
                if (a.noise_mode == 1) {
                    real z[4];
                    normals4<real>(static_cast<uint32_t>(s), qglob, 0u, kTagNoise, a.key0, a.key1, z);
#pragma unroll
                    for (int q = 0; q < 4; ++q) x[q] = M::fma(sigma, z[q], x[q]);
                } else if (a.noise_mode == 2 && valid) {
                    const real* zr = static_cast<const real*>(a.noise) + s * a.ld_noise + n0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (live[q]) x[q] = M::fma(sigma, zr[q], x[q]);
                }
                if (a.first) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const real dv = valid ? x[q] - ctr[q] : real(0);
                        sx[q] += dv;
                        sxx[q] = M::fma(dv, dv, sxx[q]);
                        nlt[q] += (valid && x[q] < tc[q]) ? 1u : 0u;
                        nle[q] += (valid && x[q] <= tc[q]) ? 1u : 0u;
                    }
                    if (a.draws_out && valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (live[q])
                                a.draws_out[s * a.ld_out + n0 + q] =
                                    static_cast<double>(x[q]) + (a.mu ? a.mu[n0 + q] : 0.0);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const bool b = valid && x[q] < wlo[q][j];
                        const bool w = valid && !b && x[q] < whi[q][j];
                        below[q][j] += b ? 1u : 0u;
                        const unsigned int mask = __ballot_sync(0xffffffffu, w);
                        if (mask) {
                            const long long idx = (n0 + q) * a.nq + j;
                            const int leader = __ffs(mask) - 1;
                            unsigned int base = 0u;
                            if (lane == leader) base = atomicAdd(a.cnt_in + idx, static_cast<unsigned int>(__popc(mask)));
                            base = __shfl_sync(0xffffffffu, base, leader);
                            const unsigned int pos = base + __popc(mask & ((1u << lane) - 1u));
                            if (w) {
                                if (pos < static_cast<unsigned int>(a.cand_cap))
                                    static_cast<real*>(a.cand)[idx * a.cand_cap + pos] = x[q];
                                // which 1/32 slice of the window: lets an overflowing window be narrowed
                                // with exact counts whatever the distribution (atoms, heavy tails)
                                const real rel = (x[q] - wlo[q][j]) * (real(kSubBins) / (whi[q][j] - wlo[q][j]));
                                int bin = static_cast<int>(rel);
                                bin = bin < 0 ? 0 : (bin > kSubBins - 1 ? kSubBins - 1 : bin);
                                atomicAdd(a.sub_cnt + idx * kSubBins + bin, 1u);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                                   // everyone is done with this stage
    }

    if (!has_quad) return;
    // warp reduction, then one atomic (integers) or one partial slot (moments) per nucleus
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            unsigned int v = below[q][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && live[q] && j < a.nq && v) atomicAdd(a.cnt_below + (n0 + q) * a.nq + j, v);
        }
        if (a.first) {
            double s1 = static_cast<double>(sx[q]), s2 = static_cast<double>(sxx[q]);
            unsigned int c1 = nlt[q], c2 = nle[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o);
            }
            if (lane == 0 && live[q]) {
                const long long slot = static_cast<long long>(blockIdx.y) * wpq + slice;
                a.mom_part[(slot * 2 + 0) * a.n_points + n0 + q] = s1;
                a.mom_part[(slot * 2 + 1) * a.n_points + n0 + q] = s2;
                if (a.truth) {
                    atomicAdd(a.c_lt + n0 + q, c1);
                    atomicAdd(a.c_le + n0 + q, c2);
                }
            }
        }
    }
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
    unsigned int* cnt_in;
    unsigned int* sub_cnt;
    void* cand;
    int cand_cap;
    unsigned char* resolved;   // [n*nq]
    const double* mu;          // nullable
    double* out_quant;         // [nq][ld_quant]
    long long ld_quant;
    long long out_offset;      // chunk offset into the output arrays
    // moments (first select only)
    int first;
    const double* mom_part;
    int n_slots;
    const void* center;
    double* out_mean;
    double* out_var;
    const unsigned int* c_lt;
    const unsigned int* c_le;
    long long* out_c_lt;
    long long* out_c_le;
    // retry bookkeeping
    const int* quad_list;      // quads examined by this launch (nullable = all)
    int n_quads;
    int* quad_flag;            // [n_quads_total] set when the quad needs another pass
    int* next_list;
    int* next_count;
};

// one warp per (nucleus, quantile): sort the window's candidates in shared memory when they fit,
// then let sel_decide (select_logic.h) read the answer or choose the next window
template <typename real>
__global__ void __launch_bounds__(128) predict_select_kernel(const SelectArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int p2cap = 1;
    while (p2cap < a.cand_cap) p2cap <<= 1;
    real* const buf = reinterpret_cast<real*>(smem_raw) + static_cast<size_t>(warp) * p2cap;
    const long long item = static_cast<long long>(blockIdx.x) * 4 + warp;       // (quad slot, q, j)
    const long long per_quad = 4ll * a.nq;
    const long long qslot = item / per_quad;
    if (qslot >= a.n_quads) return;
    const int quad = a.quad_list ? a.quad_list[qslot] : static_cast<int>(qslot);
    const int q = static_cast<int>((item % per_quad) / a.nq), j = static_cast<int>(item % a.nq);
    const long long n = 4ll * quad + q;
    if (n >= a.n_points) return;
    const long long idx = n * a.nq + j;

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
    const long long cb = a.cnt_below[idx], cw = a.cnt_in[idx];
    const long long t1 = st.phase == 2 ? r + 1 : r;
    const long long t2 = t1 + ((st.phase == 0 && need_pair) ? 1 : 0);
    const bool inside = t1 >= cb && t2 < cb + cw;
    const real* src = static_cast<const real*>(a.cand) + idx * a.cand_cap;

    bool stored_equal = false;
    real stored_value = real(0);
    if (inside && cw <= a.cand_cap) {
        int p2 = 1;
        while (p2 < cw) p2 <<= 1;
        for (int i = lane; i < p2; i += 32) buf[i] = i < cw ? src[i] : real(FLT_MAX);
        __syncwarp();
        for (int size = 2; size <= p2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (p2 >> 1); i += 32) {
                    const int lo_i = 2 * i - (i & (stride - 1));
                    const int hi_i = lo_i + stride;
                    const bool up = (lo_i & size) == 0;
                    const real x0 = buf[lo_i], x1 = buf[hi_i];
                    if ((x0 > x1) == up) {
                        buf[lo_i] = x1;
                        buf[hi_i] = x0;
                    }
                }
                __syncwarp();
            }
    } else if (inside) {
        real vmin = real(FLT_MAX), vmax = -real(FLT_MAX);
        for (int i = lane; i < a.cand_cap; i += 32) {
            vmin = fmin(vmin, src[i]);
            vmax = fmax(vmax, src[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        stored_equal = vmin == vmax;
        stored_value = vmin;
    }
    if (lane != 0) return;
    unsigned int* sub = a.sub_cnt + idx * kSubBins;
    double v0 = 0.0, v1 = 0.0;
    const SelAction act = sel_decide<real>(st, r, need_pair, cb, cw, sub, a.cand_cap, buf, stored_equal,
                                           stored_value, &v0, &v1);
    if (act == kSelResolved) {
        a.out_quant[static_cast<long long>(j) * a.ld_quant + a.out_offset + n] =
            (a.mu ? a.mu[n] : 0.0) + sel_lerp(v0, v1, a.frac[j]);
        a.resolved[idx] = 1;
        // park the window so later passes over this quad collect nothing for it
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
    a.cnt_in[idx] = 0u;
    for (int b = 0; b < kSubBins; ++b) sub[b] = 0u;
    if (atomicExch(a.quad_flag + quad, 1) == 0) {
        const int pos = atomicAdd(a.next_count, 1);
        a.next_list[pos] = quad;
    }
}

// ----------------------------------------------------------------------------------------------
// first guess of centre and spread per nucleus from the posterior sample moments:
//   E x = u.E[beta],  Var x = u' Cov[beta] u + noise * E[sigma^2]
template <typename real>
__global__ void predict_guess_kernel(const void* u_, long long n, int k, const double* theta_mean,
                                     const double* theta_cov /* [(k+1)^2] */, int noise_on, void* center_,
                                     void* scale_) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real* u = static_cast<const real*>(u_) + i * k;
    const int d = k + 1;
    double m = 0.0, v = 0.0;
    for (int r = 0; r < k; ++r) {
        const double ur = static_cast<double>(u[r]);
        m += ur * theta_mean[r];
        double t = 0.0;
        for (int c = 0; c < k; ++c) t += theta_cov[r * d + c] * static_cast<double>(u[c]);
        v += ur * t;
    }
    if (noise_on) v += theta_cov[k * d + k] + theta_mean[k] * theta_mean[k];
    static_cast<real*>(center_)[i] = static_cast<real>(m);
    static_cast<real*>(scale_)[i] = static_cast<real>(sqrt(fmax(v, 0.0)));
}

// windows from centre/scale:  [c + s (z_q - h_q), c + s (z_q + h_q))
template <typename real>
__global__ void predict_window_kernel(long long n, int nq, const void* center_, const void* scale_,
                                      const double* zq, const double* hw, void* win_lo, void* win_hi,
                                      void* brk_lo, void* brk_hi, void* pair_hi, void* aux, unsigned char* phase, unsigned int* cnt_below,
                                      unsigned int* cnt_in, unsigned int* sub_cnt, unsigned char* resolved) {
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
    cnt_in[i] = 0u;
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
