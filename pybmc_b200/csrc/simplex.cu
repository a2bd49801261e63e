// C ABI: simplex-constrained sampler (see include/bmc_b200.h; pybmc/inference_utils.py:59-144).
#include <algorithm>
#include "common.h"
#include "gibbs_kernels.cuh"

using namespace bmc;

namespace {

template <typename real, int KP>
int launch_simplex(const SimplexArgs& a, int stats_mode, int threads, cudaStream_t stream) {
    const unsigned blocks = static_cast<unsigned>((a.n_chains + threads - 1) / threads);
    const int m4 = (a.m + 3) & ~3;
    const size_t smem = sizeof(real) * (static_cast<size_t>(KP) * m4 + static_cast<size_t>(KP) * KP);
    if (smem > 200 * 1024) {
        set_error("bmc_gibbs_simplex_run: k=%d, m=%d needs %zu bytes of shared memory", a.k, a.m, smem);
        return BMC_ERR_ARG;
    }
#define BMC_SIMPLEX_LAUNCH(MODE)                                                                              \
    do {                                                                                                      \
        auto kern = gibbs_simplex_kernel<real, KP, MODE>;                                                     \
        if (smem > 32 * 1024)                                                                                 \
            BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<blocks, threads, smem, stream>>>(a);                                                           \
    } while (0)
    switch (stats_mode) {
        case BMC_STATS_NONE:
            BMC_SIMPLEX_LAUNCH(0);
            break;
        case BMC_STATS_DIAG:
            BMC_SIMPLEX_LAUNCH(1);
            break;
        default:
            if constexpr (KP <= 16) {
                BMC_SIMPLEX_LAUNCH(2);
            } else {
                set_error("bmc_gibbs_simplex_run: BMC_STATS_FULL needs k <= 16");
                return BMC_ERR_ARG;
            }
    }
#undef BMC_SIMPLEX_LAUNCH
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

template <typename real, int KP>
int launch_simplex_group(const SimplexArgs& a, int stats_mode, cudaStream_t stream) {
    const int wpb = 4, cpw = 32 / kSimplexGroup;
    const unsigned blocks = static_cast<unsigned>((a.n_chains + wpb * cpw - 1) / (wpb * cpw));
    const int mg = (a.m + kSimplexGroup - 1) / kSimplexGroup * kSimplexGroup;
    const size_t smem = sizeof(real) * (static_cast<size_t>(KP) * mg + static_cast<size_t>(wpb) * cpw * 32 * (KP + 2));
    if (smem > 200 * 1024) return 1;                       // caller falls back to one chain per thread
#define BMC_SIMPLEX_GROUP(MODE)                                                                               \
    do {                                                                                                      \
        auto kern = gibbs_simplex_group_kernel<real, KP, MODE>;                                               \
        if (smem > 32 * 1024)                                                                                 \
            BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<blocks, wpb * 32, smem, stream>>>(a);                                                          \
    } while (0)
    if (stats_mode == BMC_STATS_NONE) BMC_SIMPLEX_GROUP(0);
    else if (stats_mode == BMC_STATS_DIAG) BMC_SIMPLEX_GROUP(1);
    else BMC_SIMPLEX_GROUP(2);
#undef BMC_SIMPLEX_GROUP
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

// at most 16 models: the group kernel that precomputes the weight changes (gibbs_simplex_group16_kernel)
template <typename real, int KP>
int launch_simplex_group16(const SimplexArgs& a, int stats_mode, cudaStream_t stream) {
    const int wpb = 4, cpw = 32 / kSimplexGroup;
    const unsigned blocks = static_cast<unsigned>((a.n_chains + wpb * cpw - 1) / (wpb * cpw));
    const size_t smem = sizeof(real) * (static_cast<size_t>(KP) * kSimplexModels16 + static_cast<size_t>(KP) * KP +
                                        static_cast<size_t>(wpb) * cpw * 32 * SimplexRow16<KP>::kRow);
    if (smem > 200 * 1024) return 1;
#define BMC_SIMPLEX_GROUP16(MODE)                                                                             \
    do {                                                                                                      \
        auto kern = gibbs_simplex_group16_kernel<real, KP, MODE>;                                             \
        if (smem > 32 * 1024)                                                                                 \
            BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<blocks, wpb * 32, smem, stream>>>(a);                                                          \
    } while (0)
    if (stats_mode == BMC_STATS_NONE) BMC_SIMPLEX_GROUP16(0);
    else if (stats_mode == BMC_STATS_DIAG) BMC_SIMPLEX_GROUP16(1);
    else BMC_SIMPLEX_GROUP16(2);
#undef BMC_SIMPLEX_GROUP16
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

// below this many chains a thread-per-chain launch cannot fill the machine (SMs x 4 schedulers x
// a few warps each): give every chain eight lanes instead.  Measured crossover on B200
// (profiles/layout_sweep.py): 16,384 chains (4.08 vs 4.03 ms); 32,768: 4.8 vs 7.8 ms; 4096: 4.1 vs 1.7 ms.
constexpr long long kGroupPerChainBelow = 16384;

// layout (bmc_simplex_problem.layout): AUTO = by chain and model count; THREAD = one chain per thread;
// GROUP = the general eight-lanes kernel; WARP = the eight-lanes kernel for at most 16 models that precomputes
// the state-independent part of 32 proposals at a time.  A forced group layout that does not fit (shared
// memory) falls through to the next one, as AUTO does.
template <typename real>
int dispatch_simplex(const SimplexArgs& a, int layout, int stats_mode, int threads, cudaStream_t stream) {
    const bool few = a.n_chains < kGroupPerChainBelow;
    const bool group_fits = a.k <= 8 && a.burn + a.iterations < (1ll << 31) && a.thin < (1ll << 31);
    const bool try16 = group_fits && a.m <= kSimplexModels16 &&
                       (layout == BMC_LAYOUT_WARP || (layout == BMC_LAYOUT_AUTO && few));
    const bool try8 = group_fits && (layout == BMC_LAYOUT_GROUP || layout == BMC_LAYOUT_WARP ||
                                     (layout == BMC_LAYOUT_AUTO && few));
    if (try16) {
        const int rc = a.k <= 4 ? launch_simplex_group16<real, 4>(a, stats_mode, stream)
                                : launch_simplex_group16<real, 8>(a, stats_mode, stream);
        if (rc <= 0) return rc;
    }
    if (try8) {
        const int rc = a.k <= 4 ? launch_simplex_group<real, 4>(a, stats_mode, stream)
                                : launch_simplex_group<real, 8>(a, stats_mode, stream);
        if (rc <= 0) return rc;
    }
    if (a.k <= 4) return launch_simplex<real, 4>(a, stats_mode, threads, stream);
    if (a.k <= 8) return launch_simplex<real, 8>(a, stats_mode, threads, stream);
    if (a.k <= 16) return launch_simplex<real, 16>(a, stats_mode, threads, stream);
    if (a.k <= 32) return launch_simplex<real, 32>(a, stats_mode, threads, stream);
    return launch_simplex<real, 64>(a, stats_mode, threads, stream);
}

}  // namespace

extern "C" {

int bmc_gibbs_simplex_run(int dtype, const bmc_simplex_problem* p, uint64_t seed, uint64_t chain0,
                          int64_t n_chains, int64_t burn, int64_t iterations, int64_t thin, int64_t n_kept,
                          void* samples, double* chain_stats, int stats_mode, int32_t* accepted, void* stream) {
    BMC_REQUIRE(p, "bmc_gibbs_simplex_run: problem is NULL");
    BMC_REQUIRE(dtype == BMC_F32 || dtype == BMC_F64, "bmc_gibbs_simplex_run: bad dtype %d", dtype);
    // the reference's own checks, pybmc/inference_utils.py:91-94 (the Python layer raises ValueError)
    BMC_REQUIRE(burn >= 0, "Burn-in iterations must be non-negative.");
    BMC_REQUIRE(p->k >= 1 && p->k <= BMC_MAX_COMPONENTS && p->m >= 1, "bmc_gibbs_simplex_run: k=%d m=%d", p->k,
                p->m);
    BMC_REQUIRE(p->gram && p->b_ols && p->step && p->vt_hat, "bmc_gibbs_simplex_run: problem constants missing");
    BMC_REQUIRE(n_chains >= 1 && iterations >= 0, "bmc_gibbs_simplex_run: n_chains=%lld iterations=%lld",
                (long long)n_chains, (long long)iterations);
    BMC_REQUIRE(burn + iterations < (1ll << 32), "bmc_gibbs_simplex_run: burn + iterations must fit 32 bits");
    BMC_REQUIRE(stats_mode >= 0 && stats_mode <= 2, "bmc_gibbs_simplex_run: bad stats_mode");
    BMC_REQUIRE(stats_mode == 0 || chain_stats, "bmc_gibbs_simplex_run: chain_stats is NULL");
    BMC_REQUIRE(p->n_obs > 0 && p->nu0 + p->n_obs > 0, "bmc_gibbs_simplex_run: bad n_obs / nu0");
    BMC_REQUIRE(p->layout >= BMC_LAYOUT_AUTO && p->layout <= BMC_LAYOUT_WARP, "bmc_gibbs_simplex_run: bad layout %d",
                p->layout);
    if (samples) {
        BMC_REQUIRE(thin >= 1, "bmc_gibbs_simplex_run: thin=%lld", (long long)thin);
        BMC_REQUIRE(n_kept >= (iterations + thin - 1) / thin, "bmc_gibbs_simplex_run: n_kept too small");
    }
    cudaStream_t st = as_stream(stream);
    SimplexArgs a{};
    a.gram = p->gram;
    a.b_ols = p->b_ols;
    a.step = p->step;
    a.vt = p->vt_hat;
    a.k = p->k;
    a.m = p->m;
    a.rss_min = p->rss_min;
    a.shape = 0.5 * (p->nu0 + p->n_obs);                     // inference_utils.py:115 / :138
    a.prior_scale = p->nu0 * p->sigma20;                     // :116 / :139
    a.sigma2_init = p->rss_zero / p->n_obs;                  // :86
    a.sigma_ref = sqrt(a.sigma2_init > 0 ? a.sigma2_init : 1.0);
    a.cf = make_run_consts<float>(a.rss_min, a.prior_scale, a.sigma_ref, a.sigma2_init, a.shape);
    a.cd = make_run_consts<double>(a.rss_min, a.prior_scale, a.sigma_ref, a.sigma2_init, a.shape);
    a.gamma_boost = a.shape < 1.0;
    a.key0 = static_cast<uint32_t>(seed);
    a.key1 = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        a.keys.k0[r] = a.key0 + static_cast<uint32_t>(r) * kPhiloxW0;
        a.keys.k1[r] = a.key1 + static_cast<uint32_t>(r) * kPhiloxW1;
    }
    a.chain0 = chain0;
    a.n_chains = n_chains;
    a.burn = burn;
    a.iterations = iterations;
    a.thin = thin;
    a.n_kept = n_kept;
    a.samples = samples;
    a.chain_stats = chain_stats;
    a.stats_mode = stats_mode;
    a.accepted = accepted;
    if (stats_mode != 0) {
        const int kp = bmc_padded_components(p->k);
        BMC_CUDA(cudaMemsetAsync(chain_stats, 0,
                                 sizeof(double) * static_cast<size_t>(bmc_gibbs_n_stat(kp, stats_mode)) * n_chains, st));
    }
    if (accepted) BMC_CUDA(cudaMemsetAsync(accepted, 0, sizeof(int32_t) * n_chains, st));
    if (burn + iterations == 0) return BMC_OK;
    const long long sms = sm_count();
    const int threads = n_chains >= sms * 128 * 8 ? 128 : (n_chains >= sms * 64 * 2 ? 64 : 32);
    return dtype == BMC_F32 ? dispatch_simplex<float>(a, p->layout, stats_mode, threads, st)
                            : dispatch_simplex<double>(a, p->layout, stats_mode, threads, st);
}

}  // extern "C"
