/*
 * bmc_b200.h -- C ABI of libbmc_b200.so: sm_100a kernels for pyBMC's inference path.
 *
 * The upstream package (sudhanvalalit/pybmc) is pure Python/NumPy and has no FFI; the
 * interface each entry point replaces is therefore a Python function of the reference, cited
 * as file:line below (paths relative to the upstream tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer would add upstream.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, scalars; no torch/CUDA C++ types.  `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).  Every pointer marked "dev" is
 *     device memory owned by the caller; "host" pointers are small host arrays.
 *   - no hidden allocation: functions that need scratch take `workspace` (dev) and its size;
 *     the matching *_workspace_bytes() call returns the size to provide.
 *   - `dtype`: BMC_F32 or BMC_F64 selects the arithmetic type of sampler/predictive kernels
 *     ("real" below).  Orthogonalisation kernels are always fp64.
 *   - return value 0 on success, negative error code otherwise; bmc_last_error() returns a
 *     thread-local message for the last failure.
 *   - asynchronous w.r.t. the host unless stated.  No global mutable state: every switch between kernels
 *     that compute the same thing is a field of the problem struct of the call (`layout`, `tensor_min_k`);
 *     the only per-thread state is the message behind bmc_last_error().
 *   - launch shapes are derived from the SM count of the current device (cudaDevAttrMultiProcessorCount),
 *     never from a hard-coded 148.
 */
#ifndef BMC_B200_H
#define BMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMC_F32 0
#define BMC_F64 1

#define BMC_OK 0
#define BMC_ERR_ARG (-1)       /* invalid argument (maps to ValueError upstream)          */
#define BMC_ERR_CUDA (-2)      /* CUDA runtime failure                                    */
#define BMC_ERR_CONVERGE (-3)  /* quantile windows did not resolve within the pass limit  */
#define BMC_ERR_WORKSPACE (-4) /* workspace too small                                     */

#define BMC_STATS_NONE 0 /* no moment accumulation                                          */
#define BMC_STATS_DIAG 1 /* first moments + diagonal second moments                          */
#define BMC_STATS_FULL 2 /* first moments + all cross moments (k <= 16)                      */

#define BMC_NOISE_NONE 0
#define BMC_NOISE_PHILOX 1
#define BMC_NOISE_EXTERNAL 2

#define BMC_MAX_COMPONENTS 64
#define BMC_MAX_QUANTILES 8

/* `layout` of the sampler problems: which kernel family walks the chains.  Every layout draws the same
 * variates for the same (seed, global chain id, iteration), so they produce the same chains (up to the
 * summation order of RSS); AUTO picks by chain count.  Forcing one is for tests, A/B timing and the
 * parity checks of the benchmarked thread-per-chain kernels. */
#define BMC_LAYOUT_AUTO 0
#define BMC_LAYOUT_THREAD 1   /* one chain per thread                                                   */
#define BMC_LAYOUT_GROUP 2    /* eight lanes per chain (k <= 8); simplex: the general group kernel        */
#define BMC_LAYOUT_WARP 3     /* conjugate: a warp per chain (k <= 8); simplex: the <= 16-model group kernel
                                 that precomputes the state-independent part of 32 proposals at a time    */
#define BMC_HIST_BINS 512     /* bins of the sampler's marginal histograms (see bmc_gibbs_hist)          */

int bmc_version(void);
const char* bmc_last_error(void);
/* sm count, compute capability and opt-in shared memory of `device`. */
int bmc_device_caps(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);

/* ---- orthogonalisation: BayesianModelCombination.orthogonalize, pybmc/bmc.py:79-130 ------- */

/* mu[r] = mean_c preds[r][c] (bmc.py:106); y[r] = truth[r] - mu[r] (bmc.py:109-111);
 * xc[r][c] = preds[r][c] - mu[r] (bmc.py:114-116).  truth/y and xc may be NULL. */
int bmc_center_rows(const double* preds /*dev [n][ld]*/, int64_t n, int m, int64_t ld,
                    const double* truth /*dev [n]*/, double* mu /*dev [n]*/, double* y /*dev [n]*/,
                    double* xc /*dev [n][ldx]*/, int64_t ldx, void* stream);

/* gram = A'A with A = [a - mu | extra] (m1 = m + (extra != NULL) columns), fp64, reproducible.
 * Replaces the full-U dgesdd of bmc.py:119 by its M-by-M Gram eigenproblem, and X'X, X'y, y'y of
 * pybmc/inference_utils.py:25,28,43.  mu (row shift) and extra may be NULL. */
size_t bmc_gram_workspace_bytes(int64_t n, int m1);
int bmc_gram(const double* a /*dev [n][ld]*/, int64_t n, int m, int64_t ld, const double* mu /*dev [n]*/,
             const double* extra /*dev [n]*/, double* gram /*dev [m1][m1]*/, void* workspace, size_t workspace_bytes,
             void* stream);

/* out = (a - mu) vt' : U_hat = Xc Vt_hat' (bmc.py:122, inference_utils.py:164-166) and the K-space
 * coordinates u = preds Vt_hat' of new points (pybmc/sampling_utils.py:64-72).  mu may be NULL. */
int bmc_project_rows(const double* a /*dev [n][ld]*/, int64_t n, int m, int64_t ld, const double* mu /*dev [n]*/,
                     const double* vt /*dev [k][m]*/, int k, double* out /*dev [n][ldo]*/, int64_t ldo,
                     void* stream);

/* rss = |y - X b|^2 (inference_utils.py:29-31), reproducible two-stage sum. */
size_t bmc_rss_workspace_bytes(int64_t n);
int bmc_residual_ss(const double* x /*dev [n][ld]*/, int64_t n, int k, int64_t ld, const double* y /*dev [n]*/,
                    const double* b /*dev [k]*/, double* rss /*dev [1]*/, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- conjugate Gibbs sampler: gibbs_sampler hot loop, pybmc/inference_utils.py:39-54 ------- */

typedef struct {
    int k;                     /* components, <= BMC_MAX_COMPONENTS                               */
    const double* d;           /* dev [k]   generalised eigenvalues of (X'X, Lambda + 1e-6 I)      */
    const double* pull;        /* dev [k]   W' Lambda b0 - g_ols                                   */
    const double* g_ols;       /* dev [k]   W^-1 (X'X)^-1 X'y                                      */
    const double* w;           /* dev [k*k] row-major (dense_w) or [k] diagonal; b = W g           */
    int dense_w;
    double rss_min;            /* |y - X b_ols|^2                        (inference_utils.py:29-31) */
    double n_obs;              /* len(y)                                                    (:23)  */
    double nu0, sigma20;       /* prior_info[2], prior_info[3]                              (:21)  */
    double sigma2_init;        /* max(rss_min / n, 1e-6)                                  (:31,37) */
    int layout;                /* BMC_LAYOUT_*; 0 = choose by chain count                          */
    void* workspace;           /* dev, optional: bmc_gibbs_workspace_bytes(n_chains).  With it, a thread-per-chain
                                  launch whose chains would load the schedulers unevenly (65,536 chains: 3.46
                                  warps each) runs as persistent workers that pass the chain groups round --
                                  same chains, same results, no tail.  NULL: the plain launch.          */
    size_t workspace_bytes;
} bmc_gibbs_problem;
size_t bmc_gibbs_workspace_bytes(int64_t n_chains);

/* Runs chains [chain0, chain0 + n_chains) for `iterations` iterations each.
 * samples: real [n_kept][k+1][n_chains], iteration t is kept when t >= store_from and
 *          (t - store_from) % thin == 0; rows are [b_0..b_{k-1}, sigma] as in :54.  May be NULL.
 * chain_stats: fp64 [bmc_gibbs_n_stat(kp, stats_mode)][n_chains] sums over all iterations of the deviations
 *          e = W^-1 b - g_ols (k entries) and sigma - sqrt(sigma2_init), then second moments
 *          (diagonal, or upper triangle row-major); zeroed by this call.  May be NULL with
 *          stats_mode BMC_STATS_NONE. */
/* Kernels are compiled for component counts 4, 8, 16, 32, 64; k is padded up to the next one
 * (kp) and the moment rows are laid out for dimension kp + 1: rows [0, kp] first moments (entry kp
 * is sigma), then kp + 1 diagonal second moments (BMC_STATS_DIAG) or the upper triangle of the
 * (kp+1)-square cross-moment matrix, row-major (BMC_STATS_FULL).  Padded components stay 0. */
int bmc_padded_components(int k);
int64_t bmc_gibbs_n_stat(int kp, int stats_mode);

/* Marginal histograms of [b_0..b_{k-1}, sigma] over the run (the second thing, after the
 * moment sums, that ranks all-reduce): the state after every `every`-th iteration (t with (t+1) % every == 0) is
 * binned into BMC_HIST_BINS equal bins per coordinate, bin = floor((v - lo[c]) * inv_width[c]) clamped to the
 * edge bins.  Blocks count in shared memory (uint32) and merge into `counts` with 64-bit atomics when they
 * finish, so the result does not depend on the layout or the sharding.  `counts` is zeroed by the call; sum it
 * over ranks with one all-reduce.  k <= 16.  `every` must be a multiple of 64: the kernels bin where they flush
 * their moment sums, outside the arithmetic-only inner loop. */
typedef struct {
    int64_t every;             /* 0 = off                                                          */
    const double* lo;          /* dev [k+1]  lower edge of bin 0 per coordinate                   */
    const double* inv_width;   /* dev [k+1]  1 / bin width                                        */
    uint64_t* counts;          /* dev [k+1][BMC_HIST_BINS]                                        */
    void* workspace;           /* dev, bmc_gibbs_hist_workspace_bytes(k): replicas the blocks merge into (they all
                                  finish together; without it they queue on the words of `counts`)  */
    size_t workspace_bytes;
} bmc_gibbs_hist;
size_t bmc_gibbs_hist_workspace_bytes(int k);

int bmc_gibbs_run(int dtype, const bmc_gibbs_problem* problem /*host*/, uint64_t seed, uint64_t chain0,
                  int64_t n_chains, int64_t iterations, int64_t store_from, int64_t thin, int64_t n_kept,
                  void* samples /*dev*/, double* chain_stats /*dev*/, int stats_mode,
                  const bmc_gibbs_hist* hist /*host, may be NULL*/, void* stream);

/* Literal form of the same sampler, one chain per warp: X' ([k][n], real) and y ([n], real) are
 * staged in shared memory by TMA and every iteration factors X'X/s2 + Lambda + 1e-6 I (:41), draws b
 * (:45) and recomputes the residual y - X b over all n rows (:48-51) with a warp-shuffle reduction.
 * O(nK) per iteration: the parity anchor for bmc_gibbs_run, not the throughput path.  k <= 16 and
 * (k+1) n sizeof(real) must fit shared memory.  lam = inv(B0) [k*k], lam_b0 = lam b0 [k] (dev fp64).
 * samples: real [iterations][k+1][n_chains] (every iteration is kept), may be NULL. */
int bmc_gibbs_literal_run(int dtype, const void* xt /*dev*/, const void* y /*dev*/, int64_t n, int k,
                          const double* lam /*dev*/, const double* lam_b0 /*dev*/, double nu0, double sigma20,
                          double sigma2_init, uint64_t seed, uint64_t chain0, int64_t n_chains, int64_t iterations,
                          void* samples /*dev*/, void* stream);

/* ---- simplex sampler: gibbs_sampler_simplex loops, pybmc/inference_utils.py:97-141 --------- */

typedef struct {
    int k, m;                  /* components, models                                               */
    const double* gram;        /* dev [k*k]  X'X                                                   */
    const double* b_ols;       /* dev [k]    a least-squares solution of X b = y                   */
    const double* step;        /* dev [k]    S_hat * stepsize                               (:80)  */
    const double* vt_hat;      /* dev [k*m]  row-major                                      (:99)  */
    double rss_min;            /* |y - X b_ols|^2                                                  */
    double rss_zero;           /* |y|^2 = RSS at the start b = 0                          (:82-85) */
    double n_obs, nu0, sigma20;
    int layout;                /* BMC_LAYOUT_*; 0 = choose by chain count and model count          */
} bmc_simplex_problem;

/* Burn-in iterations [0, burn) are run but not recorded; iteration burn + t is kept when
 * t % thin == 0.  accepted: dev int32 [n_chains] sampling-phase acceptances (:135), may be NULL.
 * chain_stats as above with e = b - b_ols and sigma - sqrt(rss_zero / n). */
int bmc_gibbs_simplex_run(int dtype, const bmc_simplex_problem* problem /*host*/, uint64_t seed, uint64_t chain0,
                          int64_t n_chains, int64_t burn, int64_t iterations, int64_t thin, int64_t n_kept,
                          void* samples /*dev*/, double* chain_stats /*dev*/, int stats_mode,
                          int32_t* accepted /*dev*/, void* stream);

/* ---- fused prediction + UQ: rndm_m_random_calculator and coverage,
 *      pybmc/sampling_utils.py:40-84 and :4-37 ------------------------------------------------ */

typedef struct {
    int64_t n_points;          /* nuclei handled by this call                                      */
    uint64_t point0;           /* global index of the first one: noise is keyed on the global index so
                                  results do not depend on the sharding                            */
    int64_t n_draws;           /* S, posterior draws (reference: 10000, sampling_utils.py:57)      */
    int k;                     /* components, <= BMC_MAX_COMPONENTS                                 */
    const void* u;             /* dev real [n_points][k]   preds Vt_hat'                           */
    const double* mu;          /* dev [n_points] mean over models (the 1/M term of :64), or NULL   */
    const double* truth;       /* dev [n_points] or NULL (no coverage counts)                      */
    const void* theta;         /* dev real [n_draws][bmc_predict_theta_stride(k)] posterior draws, one row per
                                  draw: beta in columns 0..k-1, sigma (:60-61) in column stride-4, zeros elsewhere;
                                  NULL = "matrix mode": the draws are `noise` itself                  */
    int noise_mode;            /* BMC_NOISE_*                                                      */
    uint64_t seed;
    const void* noise;         /* dev real [n_draws][ld_noise] standard normals (external mode)    */
    int64_t ld_noise;
    int nq;                    /* number of quantiles, <= BMC_MAX_QUANTILES                        */
    const double* probs;       /* host [nq] in percent, np.percentile "linear" method (:80-82)     */
    const double* theta_mean;  /* dev [k+1]   sample mean of the draws (window guess)              */
    const double* theta_cov;   /* dev [(k+1)^2] sample covariance of the draws                     */
    const double* center;      /* dev [n_points] optional override of the guess (matrix mode)      */
    const double* scale;       /* dev [n_points]                                                   */
    int tensor_min_k;          /* fp32 contractions u . beta run on the tensor cores (tcgen05 kind::tf32, split
                                  TF32: fp32-level accuracy) when the padded component count reaches this
                                  threshold.  0 = the default (every k: measured faster for all of them);
                                  negative = never (FFMA kernels for every k); n >= 1 = threshold n           */
} bmc_predict_problem;

/* Outputs (all dev, fp64/int64): mean, var [n_points]; quant [nq][n_points]; c_lt, c_le [n_points]
 * = #(x < truth), #(x <= truth) (NULL without truth); draws_out [n_draws][ld_out] materialises
 * rndm_m = mu + x when not NULL.  Synchronises the stream (reads back the retry counter). */
size_t bmc_predict_workspace_bytes(int dtype, int64_t n_points, int nq, int64_t n_draws);
int bmc_predict_theta_stride(int k);
int bmc_predict_fused(int dtype, const bmc_predict_problem* problem /*host*/, double* mean, double* var,
                      double* quant, int64_t* c_lt, int64_t* c_le, double* draws_out, int64_t ld_out,
                      void* workspace, size_t workspace_bytes, int* passes_out /*host, may be NULL*/, void* stream);

/* Order counts of a materialised S-by-N matrix (the reference's rndm_m): sort-free form of
 * pybmc/sampling_utils.py:28-33.  c_lt/c_le dev int64 [n_cols]. */
int bmc_coverage_counts(const double* matrix /*dev [s_rows][ld]*/, int64_t s_rows, int64_t n_cols, int64_t ld,
                        const double* truth /*dev [n_cols]*/, int64_t* c_lt, int64_t* c_le, void* stream);

/* Column mean / standard deviation of a materialised matrix (window guess for matrix mode). */
int bmc_column_moments(const double* matrix, int64_t s_rows, int64_t n_cols, int64_t ld, double* center,
                       double* scale, void* stream);

/* covered[l] = #{ n : c_le[n] >= lo_idx[l] + 1 and c_lt[n] <= hi_idx[l] } with lo_idx/hi_idx the
 * reference's int((0.5 -+ p/200) S) indices computed on the host (sampling_utils.py:30-31).
 * lo_idx, hi_idx dev int64 [n_levels]; covered dev int64 [n_levels], zeroed by this call. */
int bmc_coverage_levels(const int64_t* c_lt, const int64_t* c_le, int64_t n_points, const int64_t* lo_idx,
                        const int64_t* hi_idx, int n_levels, int64_t* covered, void* stream);

/* ---- data split by distance: Dataset.separate_points_distance_allSets, pybmc/data.py:194-245 ---------
 * cls[i] = 0 if a reference point lies within d1 of point i (Euclidean, <=), 1 if one lies within d2 but
 * none within d1, 2 otherwise.  points dev [n][dim], refs dev [r][dim] (fp64), cls dev int32 [n]. */
int bmc_nearest_class(const double* points, int64_t n, const double* refs, int64_t r, int dim, double d1,
                      double d2, int32_t* cls, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BMC_B200_H */
