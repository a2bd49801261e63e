// C ABI: literal one-chain-per-warp Gibbs sampler (parity anchor; pybmc/inference_utils.py:39-54).
#include "common.h"
#include "literal_kernels.cuh"

using namespace bmc;

namespace {
template <typename real, int KP>
int launch_literal(const LiteralArgs& a, cudaStream_t st) {
    const long long npad = (a.n + 3) & ~3ll;
    const size_t smem = sizeof(real) * (static_cast<size_t>(a.k + 1) * npad + KP * KP + KP);
    if (smem > 220 * 1024) {
        set_error("bmc_gibbs_literal_run: n=%lld, k=%d needs %zu bytes of shared memory (max 225280)",
                  (long long)a.n, a.k, smem);
        return BMC_ERR_ARG;
    }
    auto kern = gibbs_literal_kernel<real, KP>;
    BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int warps = 8;
    const unsigned blocks = static_cast<unsigned>((a.n_chains + warps - 1) / warps);
    kern<<<blocks, warps * 32, smem, st>>>(a);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}
template <typename real>
int dispatch_literal(const LiteralArgs& a, cudaStream_t st) {
    if (a.k <= 4) return launch_literal<real, 4>(a, st);
    if (a.k <= 8) return launch_literal<real, 8>(a, st);
    return launch_literal<real, 16>(a, st);
}
}  // namespace

extern "C" {

int bmc_gibbs_literal_run(int dtype, const void* xt, const void* y, int64_t n, int k, const double* lam,
                          const double* lam_b0, double nu0, double sigma20, double sigma2_init, uint64_t seed,
                          uint64_t chain0, int64_t n_chains, int64_t iterations, void* samples, void* stream) {
    BMC_REQUIRE(dtype == BMC_F32 || dtype == BMC_F64, "bmc_gibbs_literal_run: bad dtype %d", dtype);
    BMC_REQUIRE(xt && y && lam && lam_b0, "bmc_gibbs_literal_run: null pointer");
    BMC_REQUIRE(n >= 1 && k >= 1 && k <= 16, "bmc_gibbs_literal_run: n=%lld k=%d (k <= 16)", (long long)n, k);
    BMC_REQUIRE(n_chains >= 1 && iterations >= 0 && iterations < (1ll << 32), "bmc_gibbs_literal_run: bad counts");
    BMC_REQUIRE(sigma2_init > 0 && nu0 + n > 0, "bmc_gibbs_literal_run: bad variance prior");
    if (iterations == 0) return BMC_OK;
    LiteralArgs a{};
    a.xt = xt;
    a.y = y;
    a.n = n;
    a.k = k;
    a.lam = lam;
    a.lam_b0 = lam_b0;
    a.shape = 0.5 * (nu0 + static_cast<double>(n));
    a.prior_scale = nu0 * sigma20;
    a.sigma2_init = sigma2_init;
    a.key0 = static_cast<uint32_t>(seed);
    a.key1 = static_cast<uint32_t>(seed >> 32);
    a.chain0 = chain0;
    a.n_chains = n_chains;
    a.iterations = iterations;
    a.samples = samples;
    cudaStream_t st = as_stream(stream);
    return dtype == BMC_F32 ? dispatch_literal<float>(a, st) : dispatch_literal<double>(a, st);
}

}  // extern "C"
