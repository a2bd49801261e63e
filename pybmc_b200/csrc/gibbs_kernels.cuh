// Batched samplers: one independent chain per thread, sufficient-statistic form.
//
//   gibbs_conjugate_kernel : pybmc/inference_utils.py:39-54 (gibbs_sampler hot loop)
//   gibbs_simplex_kernel   : pybmc/inference_utils.py:97-141 (gibbs_sampler_simplex loops)
//
// The reference recomputes X'y, the residual y - X b and a K-by-K inverse in every
// iteration.  Both are functions of b through K-sized statistics only:
//     RSS(b) = RSS_min + (b - b_ols)' G (b - b_ols),          G = X'X
// and, with W'(Lambda + 1e-6 I)W = I, W'GW = diag(d)  (simultaneous diagonalisation,
// done once on the host in fp64),
//     inv(G/s2 + Lambda + 1e-6 I) = W diag(1/(d/s2 + 1)) W'    (cf. :41)
// so in g = W^-1 b every coordinate is conditionally independent given s2.  The kernel
// iterates on the deviation e = g - g_ols, which has no cancellation in fp32:
//     e_k | s2 ~ N( pull_k / p_k , 1/p_k ),  p_k = d_k/s2 + 1,  pull = W'Lambda b0 - g_ols
//     RSS = RSS_min + sum_k d_k e_k^2
//     s2 | e = max( (nu0 s20 + RSS)/2 / Gamma((nu0+n)/2, 1), 1e-6 )          (:50-52)
// Moments of e (and of sigma) are accumulated per chain in fp64; samples b = W(g_ols + e)
// are written only for the iterations the caller keeps.
#pragma once
#include "rng.cuh"

namespace bmc {

struct GibbsArgs {
    // problem constants, device memory, fp64
    const double* d;       // [k]  generalised eigenvalues
    const double* pull;    // [k]  W' Lambda b0 - g_ols
    const double* g_ols;   // [k]
    const double* w;       // [k*k] row-major, b = W g   (dense_w) or [k] diagonal of W
    int k;
    int dense_w;
    double rss_min, shape, prior_scale;   // prior_scale = nu0 * sigma20
    double sigma2_init, sigma_ref;
    uint32_t key0, key1;
    unsigned long long chain0;            // global id of this launch's first chain
    long long n_chains;
    long long iterations;
    long long store_from, thin, n_kept;
    void* samples;                        // [n_kept][k+1][n_chains]  real, nullable
    double* chain_stats;                  // [n_stat][n_chains]       nullable
    int stats_mode;                       // 0 none, 1 diagonal second moments, 2 full
};

constexpr int kFlushEvery = 64;           // iterations between fp64 flushes of the moment sums

template <int KP, int MODE>
struct StatCount {
    static constexpr int D = KP + 1;
    static constexpr int value = MODE == 0 ? 0 : (MODE == 1 ? 2 * D : D + D * (D + 1) / 2);
};

template <typename real, int KP, int MODE>
__global__ void __launch_bounds__(128, sizeof(real) == 4 ? 4 : 2) gibbs_conjugate_kernel(const GibbsArgs a) {
    using M = Math<real>;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= a.n_chains) return;
    const unsigned long long gchain = a.chain0 + static_cast<unsigned long long>(tid);
    const uint32_t chain = static_cast<uint32_t>(gchain);
    constexpr int D = KP + 1;
    constexpr int NS = StatCount<KP, MODE>::value;

    real d[KP], pull[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        d[k] = k < a.k ? static_cast<real>(a.d[k]) : real(0);
        pull[k] = k < a.k ? static_cast<real>(a.pull[k]) : real(0);
    }
    const real rss_min = static_cast<real>(a.rss_min);
    const real prior_scale = static_cast<real>(a.prior_scale);
    const real sig_ref = static_cast<real>(a.sigma_ref);
    const GammaConst<real> gc = make_gamma_const<real>(a.shape);

    real acc[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < (NS > 0 ? NS : 1); ++j) acc[j] = real(0);

    real s2 = static_cast<real>(a.sigma2_init);
    real* const out = static_cast<real*>(a.samples);
    long long next_store = a.samples ? a.store_from : -1;
    long long slot = 0;

    for (long long it = 0; it < a.iterations; ++it) {
        const uint32_t it32 = static_cast<uint32_t>(it);
        const real inv_s2 = M::rcp(s2);
        real e[KP];
        real rss0 = rss_min, rss1 = real(0);
#pragma unroll
        for (int j = 0; j < (KP + 3) / 4; ++j) {
            real z[4];
            normals4<real>(it32, static_cast<uint32_t>(j), chain, kTagGibbs, a.key0, a.key1, z);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * j + q;
                if (k < KP) {
                    const real p = M::fma(d[k], inv_s2, real(1));
                    const real sd = M::rsqrt(p);
                    e[k] = sd * M::fma(pull[k], sd, z[q]);          // pull/p + z/sqrt(p)
                    if (k & 1) rss1 = M::fma(d[k] * e[k], e[k], rss1);
                    else rss0 = M::fma(d[k] * e[k], e[k], rss0);
                }
            }
        }
        const real scale = real(0.5) * (prior_scale + (rss0 + rss1));
        const real gm = gamma_unit_scale<real>(gc, it32, chain, kTagGibbs, a.key0, a.key1);
        s2 = M::div(scale, gm);
        s2 = s2 > real(1e-6) ? s2 : real(1e-6);
        const real sig = M::sqrt(s2);

        if (MODE != 0) {
            const real es = sig - sig_ref;
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] += e[k];
            acc[KP] += es;
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[D + k] = M::fma(e[k], e[k], acc[D + k]);
                acc[D + KP] = M::fma(es, es, acc[D + KP]);
            } else {
                int idx = D;
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    const real er = r < KP ? e[r < KP ? r : 0] : es;
#pragma unroll
                    for (int c = r; c < D; ++c) {
                        const real ec = c < KP ? e[c < KP ? c : 0] : es;
                        acc[idx] = M::fma(er, ec, acc[idx]);
                        ++idx;
                    }
                }
            }
            if (((it + 1) % kFlushEvery) == 0 || it + 1 == a.iterations) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    double* p = a.chain_stats + static_cast<long long>(j) * a.n_chains + tid;
                    *p += static_cast<double>(acc[j]);
                    acc[j] = real(0);
                }
            }
        }

        if (it == next_store) {
            // b = W (g_ols + e);  row layout [slot][component][chain] keeps lanes coalesced
            real* row = out + (slot * static_cast<long long>(a.k + 1)) * a.n_chains + tid;
            if (a.dense_w) {
                for (int r = 0; r < a.k; ++r) {
                    real b = real(0);
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < a.k)
                            b = M::fma(static_cast<real>(a.w[r * a.k + k]),
                                       static_cast<real>(a.g_ols[k]) + e[k], b);
                    row[static_cast<long long>(r) * a.n_chains] = b;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < a.k)
                        row[static_cast<long long>(k) * a.n_chains] =
                            static_cast<real>(a.w[k]) * (static_cast<real>(a.g_ols[k]) + e[k]);
            }
            row[static_cast<long long>(a.k) * a.n_chains] = sig;
            ++slot;
            next_store += a.thin;
        }
    }
}

// --------------------------------------------------------------------------------------
struct SimplexArgs {
    const double* gram;     // [k*k]
    const double* b_ols;    // [k]   any least-squares solution
    const double* step;     // [k]   S_hat * stepsize                      (:80)
    const double* vt;       // [k*m] row-major Vt_hat                      (:99)
    int k, m;
    double rss_min, shape, prior_scale;
    uint32_t key0, key1;
    unsigned long long chain0;
    long long n_chains;
    long long burn, iterations;
    long long thin, n_kept;
    void* samples;          // [n_kept][k+1][n_chains] real, nullable
    double* chain_stats;    // [n_stat][n_chains] nullable (moments of b - b_ols and sigma - sigma_ref)
    int stats_mode;
    double sigma_ref, sigma2_init;
    int* accepted;          // [n_chains] sampling-phase acceptances        (:135)
};

template <typename real, int KP, int MODE>
__global__ void __launch_bounds__(128) gibbs_simplex_kernel(const SimplexArgs a) {
    using M = Math<real>;
    constexpr int UR = KP <= 16 ? KP : 4;   // full unroll only while the state fits registers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* const vt_s = reinterpret_cast<real*>(smem_raw);           // [KP][m4]  zero padded
    const int m4 = (a.m + 3) & ~3;
    real* const gram_s = vt_s + KP * m4;                            // [KP][KP]
    for (int i = threadIdx.x; i < KP * m4; i += blockDim.x) {
        const int k = i / m4, m = i % m4;
        vt_s[i] = (k < a.k && m < a.m) ? static_cast<real>(a.vt[k * a.m + m]) : real(0);
    }
    for (int i = threadIdx.x; i < KP * KP; i += blockDim.x) {
        const int r = i / KP, c = i % KP;
        gram_s[i] = (r < a.k && c < a.k) ? static_cast<real>(a.gram[r * a.k + c]) : real(0);
    }
    __syncthreads();

    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= a.n_chains) return;
    const uint32_t chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(tid));
    constexpr int D = KP + 1;
    constexpr int NS = StatCount<KP, MODE>::value;

    real step[KP], b_ols[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        step[k] = k < a.k ? static_cast<real>(a.step[k]) : real(0);
        b_ols[k] = k < a.k ? static_cast<real>(a.b_ols[k]) : real(0);
    }
    const real bias0 = static_cast<real>(1.0 / static_cast<double>(a.m));       // :78
    const real prior_scale = static_cast<real>(a.prior_scale);
    const real sig_ref = static_cast<real>(a.sigma_ref);
    const GammaConst<real> gc = make_gamma_const<real>(a.shape);

    // state: dc = b - b_ols (b starts at 0, :82), gdc = G dc, rss = RSS(b)
    real dc[KP], gdc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) dc[k] = -b_ols[k];
    real rss = static_cast<real>(a.rss_min);
#pragma unroll UR
    for (int r = 0; r < KP; ++r) {
        real s = real(0);
#pragma unroll UR
        for (int c = 0; c < KP; ++c) s = M::fma(gram_s[r * KP + c], dc[c], s);
        gdc[r] = s;
        rss = M::fma(s, dc[r], rss);
    }
    real s2 = static_cast<real>(a.sigma2_init);                               // :86, RSS(0)/n

    real acc[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < (NS > 0 ? NS : 1); ++j) acc[j] = real(0);
    int n_acc = 0;
    real* const out = static_cast<real*>(a.samples);
    long long slot = 0;
    long long next_store = a.samples ? a.burn : -1;
    const long long total = a.burn + a.iterations;

    for (long long it = 0; it < total; ++it) {
        const uint32_t it32 = static_cast<uint32_t>(it);
        real delta[KP];
#pragma unroll
        for (int j = 0; j < (KP + 3) / 4; ++j) {
            real z[4];
            normals4<real>(it32, static_cast<uint32_t>(j), chain, kTagSimplex, a.key0, a.key1, z);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * j + q < KP) delta[4 * j + q] = step[4 * j + q] * z[q];           // :98 / :121
        }
        // weights of the proposal, w = b' Vt_hat + 1/M  (:99); only their minimum matters (:102)
        real prop[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) prop[k] = (b_ols[k] + dc[k]) + delta[k];
        real wmin = real(1);
        for (int m = 0; m < m4; m += 4) {
            real w0 = bias0, w1 = bias0, w2 = bias0, w3 = bias0;
#pragma unroll UR
            for (int k = 0; k < KP; ++k) {
                const real* v = vt_s + k * m4 + m;
                w0 = M::fma(prop[k], v[0], w0);
                w1 = M::fma(prop[k], v[1], w1);
                w2 = M::fma(prop[k], v[2], w2);
                w3 = M::fma(prop[k], v[3], w3);
            }
            wmin = fmin(wmin, fmin(fmin(w0, w1), fmin(w2, w3)));
        }
        const bool feasible = !(wmin < real(0));
        // RSS' - RSS = delta' G (2 dc + delta): no cancellation for small steps
        real gdelta[KP];
        real diff = real(0);
#pragma unroll UR
        for (int r = 0; r < KP; ++r) {
            real s = real(0);
#pragma unroll UR
            for (int c = 0; c < KP; ++c) s = M::fma(gram_s[r * KP + c], delta[c], s);
            gdelta[r] = s;
            diff = M::fma(delta[r], M::fma(real(2), gdc[r], s), diff);
        }
        const real u = M::u01(philox4x32_10(it32, kBlockUniform, chain, kTagSimplex, a.key0, a.key1).x);
        // min(1, exp((ll' - ll)/s2)) with ll = -RSS: note no factor 1/2 (:108 / :130)
        const real alpha = M::exp(M::div(-diff, s2));
        const bool accept = feasible && (u < fmin(real(1), alpha));
        if (accept) {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                dc[k] += delta[k];
                gdc[k] += gdelta[k];
            }
            rss += diff;
            n_acc += it >= a.burn ? 1 : 0;
        }
        const real scale = real(0.5) * (prior_scale + rss);                              // :116 / :139
        const real gm = gamma_unit_scale<real>(gc, it32, chain, kTagSimplex, a.key0, a.key1);
        s2 = M::div(scale, gm);                                                          // no floor (:117)
        if (it < a.burn) continue;
        const real sig = M::sqrt(s2);
        if (MODE != 0) {
            const real es = sig - sig_ref;
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] += dc[k];
            acc[KP] += es;
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[D + k] = M::fma(dc[k], dc[k], acc[D + k]);
                acc[D + KP] = M::fma(es, es, acc[D + KP]);
            } else {
                int idx = D;
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    const real er = r < KP ? dc[r < KP ? r : 0] : es;
#pragma unroll
                    for (int c = r; c < D; ++c) {
                        const real ec = c < KP ? dc[c < KP ? c : 0] : es;
                        acc[idx] = M::fma(er, ec, acc[idx]);
                        ++idx;
                    }
                }
            }
            const long long done = it - a.burn + 1;
            if ((done % kFlushEvery) == 0 || it + 1 == total) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    double* p = a.chain_stats + static_cast<long long>(j) * a.n_chains + tid;
                    *p += static_cast<double>(acc[j]);
                    acc[j] = real(0);
                }
            }
        }
        if (it == next_store) {
            real* row = out + (slot * static_cast<long long>(a.k + 1)) * a.n_chains + tid;
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < a.k) row[static_cast<long long>(k) * a.n_chains] = b_ols[k] + dc[k];
            row[static_cast<long long>(a.k) * a.n_chains] = sig;
            ++slot;
            next_store += a.thin;
        }
    }
    if (a.accepted) a.accepted[tid] = n_acc;
}

}  // namespace bmc
