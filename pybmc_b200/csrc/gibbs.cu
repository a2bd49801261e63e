// C ABI: conjugate Gibbs sampler (see include/bmc_b200.h; pybmc/inference_utils.py:4-56).
#include <algorithm>
#include <cmath>
#include "common.h"
#include "gibbs_kernels.cuh"

using namespace bmc;

namespace {

// with fewer chains than this a thread-per-chain launch leaves schedulers idle; eight lanes per chain
// cost ~50 % more instructions per chain-iteration but bring eight times the warps.  Measured crossover
// on B200 (profiles/layout_sweep.py): 16,384 chains (1.23 vs 1.18 ms); 32,768: 1.8 vs 2.3; 4096: 1.2 vs 0.5.
constexpr long long kConjGroupBelow = 16384;
// below this a chain gets the whole warp for generating its variates: one chain of 50,000 fp64 iterations
// 23.5 -> 16.2 ms; break-even near 2,048 chains (profiles/warp_sweep.py)
constexpr long long kConjWarpBelow = 1536;

constexpr int kHistReplicas = 64;          // copies of the histograms the blocks of a launch merge into

// dynamic shared memory of a launch: the block's marginal histograms, when they are on
size_t hist_smem(const GibbsArgs& a) {
    return a.hist_every ? sizeof(unsigned) * static_cast<size_t>(a.k + 1) * kHistBins : 0;
}

constexpr unsigned kItemIterations = 256;   // iterations per work item of the persistent launch (4 flush periods)

// One instantiation of the thread-per-chain kernel: plain (a warp per group of 32 chains, the whole run) or
// persistent (gibbs_kernels.cuh "Work items": every resident warp slot holds a worker, the chain groups rotate
// through them).  The persistent form pays when the plain one would leave the schedulers unevenly loaded -- a
// fractional number of warps per scheduler, or a partial last wave -- and needs the caller's workspace.
template <typename real, int KP, int MODE, bool HIST>
int launch_thread_kernel(GibbsArgs a, int threads, size_t smem, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
    auto kern = gibbs_conjugate_kernel<real, KP, MODE, HIST>;
    const long long groups = (a.n_chains + 31) / 32;
    const long long sched = 4ll * sm_count();
    if (workspace && workspace_bytes >= bmc_gibbs_workspace_bytes(a.n_chains) && a.iterations >= 2 * kItemIterations &&
        groups > sched) {
        int per_sm = 0;
        BMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 64, smem));
        // workers: a whole number of warps per scheduler, as many as fit and no more than there are groups for, so
        // that every worker always finds its next item finished (measured, 65,536 fp32 chains: 3 per scheduler 7.98
        // ms; 4 per scheduler, a sixth of them waiting at any time, 8.25; 3.5 -- two schedulers of an SM with 4 -- 9.09;
        // the plain launch 9.06)
        const long long per_sched = std::min<long long>(per_sm / 2, groups / sched);
        const long long workers = per_sched * sched;
        // (one worker per scheduler would serialise what the plain launch overlaps -- K = 64, 32,768 chains: 6.8 ms
        //  against 4.0 -- so at least two)
        if (per_sched >= 2 && groups % workers != 0) {
            const size_t flags = (static_cast<size_t>(groups) * sizeof(unsigned) + 255) / 256 * 256;
            a.item_its = kItemIterations;
            // who takes which item (gibbs_kernels.cuh): list order, or owned groups with the left-over ones going
            // round.  Measured on B200, fp32: 65,536 chains (left-over fraction 0.15 of W) 8.27 vs 8.40 ms, 100,000
            // chains (0.32) 13.5 vs 12.9 ms, 40,000 chains and the fp64 launches: no difference
            const double left = static_cast<double>(groups % workers) / static_cast<double>(workers);
            a.item_order = left >= 0.25 && left <= 0.6;
            a.item_done = static_cast<unsigned*>(workspace);
            a.item_state = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + flags);
            BMC_CUDA(cudaMemsetAsync(a.item_done, 0, flags, stream));
            // fewer blocks per SM than would fit: pad the dynamic shared memory so that exactly 2 per_sched blocks fit,
            // or the block scheduler may give one SM eight blocks and another four
            size_t smem_launch = smem;
            const int nb = static_cast<int>(2 * per_sched);
            if (nb < per_sm) {
                cudaFuncAttributes fa{};
                BMC_CUDA(cudaFuncGetAttributes(&fa, kern));         // static shared memory counts against the same budget
                const size_t each = (static_cast<size_t>(228) * 1024 / nb - 1024 - fa.sharedSizeBytes) / 128 * 128;
                smem_launch = std::max(smem, each);
                if (smem_launch > 48 * 1024)
                    BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  static_cast<int>(smem_launch)));
            }
            void* params[] = {&a};
            // cooperative: all workers are co-resident by construction, or the launch fails (never a deadlock)
            BMC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern),
                                                 dim3(static_cast<unsigned>(nb * sm_count())), dim3(64), params,
                                                 smem_launch, stream));
            return BMC_OK;
        }
    }
    const unsigned blocks = static_cast<unsigned>((a.n_chains + threads - 1) / threads);
    kern<<<blocks, threads, smem, stream>>>(a);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

template <typename real, int KP>
int launch_conjugate(const GibbsArgs& a, int stats_mode, int threads, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const size_t smem = hist_smem(a);
    if (a.hist_every) {
        if constexpr (KP <= 16) {
            switch (stats_mode) {
                case BMC_STATS_NONE: return launch_thread_kernel<real, KP, 0, true>(a, threads, smem, ws, ws_bytes, stream);
                case BMC_STATS_DIAG: return launch_thread_kernel<real, KP, 1, true>(a, threads, smem, ws, ws_bytes, stream);
                default: return launch_thread_kernel<real, KP, 2, true>(a, threads, smem, ws, ws_bytes, stream);
            }
        }
    }
    switch (stats_mode) {
        case BMC_STATS_NONE: return launch_thread_kernel<real, KP, 0, false>(a, threads, 0, ws, ws_bytes, stream);
        case BMC_STATS_DIAG: return launch_thread_kernel<real, KP, 1, false>(a, threads, 0, ws, ws_bytes, stream);
        default:
            if constexpr (KP <= 16) {
                return launch_thread_kernel<real, KP, 2, false>(a, threads, 0, ws, ws_bytes, stream);
            } else {
                set_error("bmc_gibbs_run: BMC_STATS_FULL needs k <= 16");
                return BMC_ERR_ARG;
            }
    }
}

template <typename real, int KP, int GEN>
int launch_group(const GibbsArgs& a, int stats_mode, cudaStream_t stream) {
    const int wpb = 4, cpw = 32 / GEN;
    const unsigned blocks = static_cast<unsigned>((a.n_chains + wpb * cpw - 1) / (wpb * cpw));
    const size_t smem = hist_smem(a);
    // the group kernels keep their iteration counters in 32-bit signed registers
    BMC_REQUIRE(a.iterations < (1ll << 31) && a.thin < (1ll << 31) && a.store_from < (1ll << 31),
                "bmc_gibbs_run: iterations, thin and store_from must fit 31 bits with fewer than %lld chains "
                "(eight lanes / a warp per chain)", kConjGroupBelow);
    if (a.hist_every) {
        // static variate buffer + the block's histograms can pass 48 KB in fp64: opt in to the larger carve-out
#define BMC_GROUP_HIST(MODE)                                                                                   \
    do {                                                                                                       \
        auto kern = gibbs_conjugate_group_kernel<real, KP, MODE, GEN, true>;                                   \
        BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kern<<<blocks, wpb * 32, smem, stream>>>(a);                                                           \
    } while (0)
        if (stats_mode == BMC_STATS_NONE) BMC_GROUP_HIST(0);
        else if (stats_mode == BMC_STATS_DIAG) BMC_GROUP_HIST(1);
        else BMC_GROUP_HIST(2);
#undef BMC_GROUP_HIST
        BMC_LAUNCH_CHECK();
        return BMC_OK;
    }
    switch (stats_mode) {
        case BMC_STATS_NONE:
            gibbs_conjugate_group_kernel<real, KP, 0, GEN><<<blocks, wpb * 32, 0, stream>>>(a);
            break;
        case BMC_STATS_DIAG:
            gibbs_conjugate_group_kernel<real, KP, 1, GEN><<<blocks, wpb * 32, 0, stream>>>(a);
            break;
        default:
            gibbs_conjugate_group_kernel<real, KP, 2, GEN><<<blocks, wpb * 32, 0, stream>>>(a);
    }
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

template <typename real>
int dispatch_conjugate(const GibbsArgs& a, int layout, int stats_mode, int threads, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
    // (a two-lanes-per-chain variant was measured in round 1: 46 % more instructions for 70 % instead of
    //  61 % issue utilisation and register-limited to 16 warps/SM -- slower; see profiles/r1_notes.md)
    if (layout == BMC_LAYOUT_AUTO) {
        const bool wide_counts = a.iterations >= (1ll << 31) || a.thin >= (1ll << 31) || a.store_from >= (1ll << 31);
        if (a.k > 8 || a.n_chains >= kConjGroupBelow || wide_counts) layout = BMC_LAYOUT_THREAD;
        else layout = a.n_chains < kConjWarpBelow ? BMC_LAYOUT_WARP : BMC_LAYOUT_GROUP;
    }
    if (layout == BMC_LAYOUT_WARP)
        return a.k <= 4 ? launch_group<real, 4, 32>(a, stats_mode, stream)
                        : launch_group<real, 8, 32>(a, stats_mode, stream);
    if (layout == BMC_LAYOUT_GROUP)
        return a.k <= 4 ? launch_group<real, 4, 8>(a, stats_mode, stream)
                        : launch_group<real, 8, 8>(a, stats_mode, stream);
    if (a.k <= 4) return launch_conjugate<real, 4>(a, stats_mode, threads, ws, ws_bytes, stream);
    if (a.k <= 8) return launch_conjugate<real, 8>(a, stats_mode, threads, ws, ws_bytes, stream);
    if (a.k <= 16) return launch_conjugate<real, 16>(a, stats_mode, threads, ws, ws_bytes, stream);
    if (a.k <= 32) return launch_conjugate<real, 32>(a, stats_mode, threads, ws, ws_bytes, stream);
    return launch_conjugate<real, 64>(a, stats_mode, threads, ws, ws_bytes, stream);
}

int pick_threads(long long n_chains) {
    // small blocks spread few chains over all SMs; 128 once there is plenty of work
    const long long sms = sm_count();
    if (n_chains >= sms * 128 * 8) return 128;
    if (n_chains >= sms * 64 * 2) return 64;
    return 32;
}

}  // namespace

extern "C" {

int bmc_padded_components(int k) { return k <= 4 ? 4 : k <= 8 ? 8 : k <= 16 ? 16 : k <= 32 ? 32 : 64; }

size_t bmc_gibbs_workspace_bytes(int64_t n_chains) {
    if (n_chains < 1) return 0;
    const size_t groups = static_cast<size_t>((n_chains + 31) / 32);
    return (groups * sizeof(unsigned) + 255) / 256 * 256 + static_cast<size_t>(n_chains) * sizeof(double);
}

size_t bmc_gibbs_hist_workspace_bytes(int k) {
    return sizeof(uint64_t) * static_cast<size_t>(kHistReplicas) * (k + 1) * kHistBins;
}

int64_t bmc_gibbs_n_stat(int k, int stats_mode) {
    const int64_t d = k + 1;
    if (stats_mode == BMC_STATS_NONE) return 0;
    if (stats_mode == BMC_STATS_DIAG) return 2 * d;
    return d + d * (d + 1) / 2;
}

int bmc_gibbs_run(int dtype, const bmc_gibbs_problem* p, uint64_t seed, uint64_t chain0, int64_t n_chains,
                  int64_t iterations, int64_t store_from, int64_t thin, int64_t n_kept, void* samples,
                  double* chain_stats, int stats_mode, const bmc_gibbs_hist* hist, void* stream) {
    BMC_REQUIRE(p, "bmc_gibbs_run: problem is NULL");
    BMC_REQUIRE(dtype == BMC_F32 || dtype == BMC_F64, "bmc_gibbs_run: bad dtype %d", dtype);
    BMC_REQUIRE(p->k >= 1 && p->k <= BMC_MAX_COMPONENTS, "bmc_gibbs_run: k=%d outside 1..%d", p->k,
                BMC_MAX_COMPONENTS);
    BMC_REQUIRE(p->d && p->pull && p->g_ols && p->w, "bmc_gibbs_run: problem constants missing");
    BMC_REQUIRE(n_chains >= 1 && iterations >= 0, "bmc_gibbs_run: n_chains=%lld iterations=%lld",
                (long long)n_chains, (long long)iterations);
    BMC_REQUIRE(iterations < (1ll << 32), "bmc_gibbs_run: iterations must fit 32 bits");
    BMC_REQUIRE(stats_mode >= 0 && stats_mode <= 2, "bmc_gibbs_run: bad stats_mode");
    BMC_REQUIRE(stats_mode == 0 || chain_stats, "bmc_gibbs_run: chain_stats is NULL");
    BMC_REQUIRE(p->nu0 + p->n_obs > 0 && p->sigma2_init > 0, "bmc_gibbs_run: bad variance prior");
    BMC_REQUIRE(p->layout >= BMC_LAYOUT_AUTO && p->layout <= BMC_LAYOUT_WARP, "bmc_gibbs_run: bad layout %d",
                p->layout);
    BMC_REQUIRE(p->layout == BMC_LAYOUT_AUTO || p->layout == BMC_LAYOUT_THREAD || p->k <= 8,
                "bmc_gibbs_run: the group layouts hold one component per lane (k <= 8), k=%d", p->k);
    if (hist && hist->every > 0) {
        BMC_REQUIRE(hist->lo && hist->inv_width && hist->counts, "bmc_gibbs_run: histogram arrays missing");
        BMC_REQUIRE(hist->every < (1ll << 31) && hist->every % kFlushEvery == 0,
                    "bmc_gibbs_run: hist.every must be a multiple of %d (the kernels bin where they flush their moment "
                    "sums) and fit 31 bits, got %lld", kFlushEvery, (long long)hist->every);
        BMC_REQUIRE(p->k <= 16, "bmc_gibbs_run: marginal histograms need k <= 16 (one block's bins live in "
                                "shared memory), k=%d", p->k);
    }
    if (samples) {
        BMC_REQUIRE(thin >= 1 && store_from >= 0, "bmc_gibbs_run: thin=%lld store_from=%lld", (long long)thin,
                    (long long)store_from);
        const int64_t need = iterations > store_from ? (iterations - store_from + thin - 1) / thin : 0;
        BMC_REQUIRE(n_kept >= need, "bmc_gibbs_run: n_kept=%lld < %lld kept iterations", (long long)n_kept,
                    (long long)need);
    }
    cudaStream_t st = as_stream(stream);
    GibbsArgs a{};
    a.d = p->d;
    a.pull = p->pull;
    a.g_ols = p->g_ols;
    a.w = p->w;
    a.k = p->k;
    a.dense_w = p->dense_w;
    a.rss_min = p->rss_min;
    a.shape = 0.5 * (p->nu0 + p->n_obs);                      // inference_utils.py:50
    a.prior_scale = p->nu0 * p->sigma20;                      // :51
    a.sigma2_init = p->sigma2_init;
    a.sigma_ref = sqrt(p->sigma2_init);
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
    a.iterations = iterations;
    a.store_from = store_from;
    a.thin = thin;
    a.n_kept = n_kept;
    a.samples = samples;
    a.chain_stats = chain_stats;
    a.stats_mode = stats_mode;
    if (hist && hist->every > 0) {
        a.hist_every = static_cast<uint32_t>(hist->every);
        a.hist_lo = hist->lo;
        a.hist_inv = hist->inv_width;
        const size_t words = static_cast<size_t>(p->k + 1) * kHistBins;
        if (hist->workspace && hist->workspace_bytes >= bmc_gibbs_hist_workspace_bytes(p->k)) {
            a.hist = static_cast<unsigned long long*>(hist->workspace);
            a.hist_replicas = kHistReplicas;
        } else {
            a.hist = reinterpret_cast<unsigned long long*>(hist->counts);      // correct, but the merge serialises
            a.hist_replicas = 1;
        }
        BMC_CUDA(cudaMemsetAsync(a.hist, 0, sizeof(uint64_t) * words * a.hist_replicas, st));
    }
    if (stats_mode != 0) {
        // the kernel accumulates with the padded component count kp = bmc_padded_components(k)
        const int kp = bmc_padded_components(p->k);
        BMC_CUDA(cudaMemsetAsync(chain_stats, 0,
                                 sizeof(double) * static_cast<size_t>(bmc_gibbs_n_stat(kp, stats_mode)) * n_chains, st));
    }
    if (iterations == 0) {
        if (a.hist_every) BMC_CUDA(cudaMemsetAsync(hist->counts, 0, sizeof(uint64_t) * (p->k + 1) * kHistBins, st));
        return BMC_OK;
    }
    const int threads = pick_threads(n_chains);
    const int rc = dtype == BMC_F32
                       ? dispatch_conjugate<float>(a, p->layout, stats_mode, threads, p->workspace, p->workspace_bytes, st)
                       : dispatch_conjugate<double>(a, p->layout, stats_mode, threads, p->workspace, p->workspace_bytes, st);
    if (rc == BMC_OK && a.hist_every && a.hist_replicas > 1) {
        const int words = (p->k + 1) * kHistBins;
        hist_reduce_kernel<<<(words + 255) / 256, 256, 0, st>>>(a.hist, a.hist_replicas, words,
                                                                reinterpret_cast<unsigned long long*>(hist->counts));
        BMC_LAUNCH_CHECK();
    }
    return rc;
}

}  // extern "C"
