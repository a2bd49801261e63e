// C ABI: fused posterior prediction + UQ (see include/bmc_b200.h;
// pybmc/sampling_utils.py:40-84 and :4-37).
#include <algorithm>
#include <cmath>
#include <cstring>
#include "common.h"
#include "predict_kernels.cuh"
#include "predict_tc_kernels.cuh"

using namespace bmc;

namespace {

constexpr int kMaxPasses = 64;
#ifndef BMC_WINDOW_SIGMAS
#define BMC_WINDOW_SIGMAS 3.0
#endif
#ifndef BMC_WINDOW_SLACK
#define BMC_WINDOW_SLACK 0.002
#endif
// Round 2: 4.0 / 0.005 -> 3.0 / 0.002 (a third fewer candidates: 1e5 x 1e5 31.0 -> 29.4 ms, still one pass: the
// guess is built from the moments of the same draws, so it follows the sample quantile more closely than an
// independent estimate would; a miss only costs a retry pass over the nuclei concerned).
constexpr double kWindowSigmas = BMC_WINDOW_SIGMAS;   // half-width in sampling standard errors of the order statistic
constexpr double kWindowSlack = BMC_WINDOW_SLACK;     // plus this fraction of the predictive spread (model error of the guess)

// Acklam's rational approximation of the standard normal quantile (|rel err| < 1.2e-9): only used to
// place the first window, never in a result.
double norm_ppf(double p) {
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01,  -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    const double plow = 0.02425;
    if (p < plow) {
        const double q = std::sqrt(-2 * std::log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    if (p > 1 - plow) return -norm_ppf(1 - p);
    const double q = p - 0.5, r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

// fp32 contractions with padded K >= this go to the tensor cores (bmc_predict_problem.tensor_min_k overrides it)
constexpr int kTensorDefaultMinK = 4;      // measured faster for every K (profiles/r1_notes.md)
inline bool use_tensor_path(int kp, int tensor_min_k) {
    if (tensor_min_k < 0) return false;
    return kp >= (tensor_min_k == 0 ? kTensorDefaultMinK : tensor_min_k);
}

// The quantile plan travels to the device as a kernel argument (no copy from the caller's stack, no
// synchronisation): layout of the destination = rank[kMaxQuant], frac[], zq[], hw[].
struct PlanImage {
    long long rank[kMaxQuant];
    double frac[kMaxQuant], zq[kMaxQuant], hw[kMaxQuant];
};
__global__ void store_plan_kernel(const PlanImage p, PlanImage* dst) {
    if (threadIdx.x < kMaxQuant) {
        dst->rank[threadIdx.x] = p.rank[threadIdx.x];
        dst->frac[threadIdx.x] = p.frac[threadIdx.x];
        dst->zq[threadIdx.x] = p.zq[threadIdx.x];
        dst->hw[threadIdx.x] = p.hw[threadIdx.x];
    }
}

struct QuantPlan {
    int nq;
    long long rank[kMaxQuant];
    double frac[kMaxQuant], zq[kMaxQuant], hw[kMaxQuant];
    double max_expected;      // largest expected number of draws inside a first window
};

// np.percentile(..., method="linear"): virtual index v = q/100 (S-1), order statistics floor(v), floor(v)+1
// (pybmc/sampling_utils.py:80-82).
QuantPlan make_plan(const double* probs, int nq, long long s) {
    QuantPlan pl{};
    pl.nq = nq;
    pl.max_expected = 16.0;
    for (int j = 0; j < nq; ++j) {
        const double v = (probs[j] / 100.0) * static_cast<double>(s - 1);
        long long r = static_cast<long long>(std::floor(v));
        double f = v - static_cast<double>(r);
        if (r >= s - 1) {
            r = s - 1;
            f = 0.0;
        }
        if (r < 0) r = 0;
        pl.rank[j] = r;
        pl.frac[j] = f;
        double p = (static_cast<double>(r) + 0.5 + f) / static_cast<double>(s);
        p = std::min(std::max(p, 0.5 / s), 1.0 - 0.5 / s);
        const double z = norm_ppf(p);
        const double pdf = std::exp(-0.5 * z * z) / std::sqrt(2.0 * M_PI);
        const double se = std::sqrt(p * (1.0 - p) / static_cast<double>(s)) / pdf;
        pl.zq[j] = z;
        pl.hw[j] = kWindowSigmas * se + kWindowSlack;
        pl.max_expected = std::max(pl.max_expected, 2.0 * pl.hw[j] * pdf * static_cast<double>(s));
    }
    return pl;
}

// room one sample split needs for its share of a window's draws (mean + 6 sigma of a Poisson share)
int segment_len(double expected_total, int splits) {
    const double e = expected_total / splits;
    return static_cast<int>(1.5 * e + 6.0 * std::sqrt(e) + 8.0);
}

struct PassShape {
    int s_splits;       // sample slots of the first pass
    int seg_len, cand_stride;
    int retry_splits;   // sample slots of a retry pass
};

// tensor-core pass: blocks of 128 nuclei, one block per SM, four slots per block
PassShape make_shape_tc(long long n_points, long long n_draws, double expected, int sms) {
    PassShape sh{};
    const int tiles = static_cast<int>((n_draws + kTcTile - 1) / kTcTile);
    const long long blocks_x = (n_points + kTcRows - 1) / kTcRows;
    const int gy_max = std::max(1, std::min(tiles, kMaxSlots / kTcSlotsPerBlock));
    int gy = 1;
    double best = 0.0;
    for (int g = 1; g <= gy_max; ++g) {          // fewest splits that fill whole waves of one block per SM
        const long long total = blocks_x * g;
        const double eff = static_cast<double>(total) / static_cast<double>((total + sms - 1) / sms * sms);
        if (eff > best + 1e-9) {
            best = eff;
            gy = g;
        }
        if (eff >= 0.92) break;
    }
    sh.s_splits = kTcSlotsPerBlock * gy;
    sh.seg_len = segment_len(expected, sh.s_splits);
    sh.cand_stride = (sh.s_splits * sh.seg_len + 3) / 4 * 4;
    int rs = kTcSlotsPerBlock * std::max(1, std::min(4, tiles));
    while (rs > kTcSlotsPerBlock && rs * segment_len(expected, rs) > sh.cand_stride) rs >>= 1;
    sh.retry_splits = rs;
    return sh;
}

PassShape make_shape(long long n_points, long long n_draws, double expected, bool tc = false, int sms = 148) {
    if (tc) return make_shape_tc(n_points, n_draws, expected, sms);
    PassShape sh{};
    const int tiles = static_cast<int>((n_draws + kPredTile - 1) / kPredTile);
    const long long blocks_x = (n_points + kPredWarps * 32 - 1) / (kPredWarps * 32);
    const int target = 3 * sms;                      // three resident blocks per SM
    int s = 1;
    if (blocks_x < target) s = static_cast<int>((target + blocks_x - 1) / blocks_x);
    s = std::max(1, std::min(s, std::min(tiles, kMaxSlots)));
    sh.s_splits = s;
    sh.seg_len = segment_len(expected, s);
    sh.cand_stride = (s * sh.seg_len + 3) / 4 * 4;
    // retry passes touch few nuclei: as many splits as the same buffer can host
    int rs = std::min(16, std::min(tiles, kMaxSlots));
    while (rs > 1 && rs * segment_len(expected, rs) > sh.cand_stride) rs >>= 1;
    sh.retry_splits = std::max(1, rs);
    return sh;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Layout {
    // byte offsets inside the workspace for a chunk of nc nuclei
    size_t center, scale, win_lo, win_hi, brk_lo, brk_hi, pair_hi, aux, phase, cnt_below, cnt_slot, sub, resolved,
        cand, mom, c_lt, c_le, flag, list_a, list_b, counter, plan, image, total;
};

Layout make_layout(long long nc, int nq, int cand_stride, size_t sz, size_t image_bytes = 0) {
    Layout l{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off = align_up(off + bytes);
        return at;
    };
    const size_t nqn = static_cast<size_t>(nc) * nq;
    l.plan = take(sizeof(PlanImage));
    l.counter = take(sizeof(int));
    l.center = take(nc * sz);
    l.scale = take(nc * sz);
    l.win_lo = take(nqn * sz);
    l.win_hi = take(nqn * sz);
    l.brk_lo = take(nqn * sz);
    l.brk_hi = take(nqn * sz);
    l.pair_hi = take(nqn * sz);
    l.aux = take(nqn * sz);
    l.phase = take(nqn);
    l.cnt_below = take(nqn * 4);
    l.cnt_slot = take(nqn * 4 * kMaxSlots);
    l.sub = take(nqn * 4 * kSubBins);
    l.resolved = take(nqn);
    l.cand = take(nqn * cand_stride * sz);
    l.mom = take(static_cast<size_t>(kMaxSlots) * 2 * nc * sizeof(double));
    l.c_lt = take(nc * 4);
    l.c_le = take(nc * 4);
    l.flag = take(nc * 4);
    l.list_a = take(nc * 4);
    l.list_b = take(nc * 4);
    l.image = take(image_bytes);
    l.total = off;
    return l;
}

#ifndef BMC_CHUNK_POINTS
#define BMC_CHUNK_POINTS 32768
#endif
constexpr long long kChunkPoints = BMC_CHUNK_POINTS;     // nuclei per chunk when the workspace allows

template <typename real>
__global__ void copy_guess_kernel(const double* c, const double* s, long long n, void* center, void* scale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    static_cast<real*>(center)[i] = static_cast<real>(c[i]);
    static_cast<real*>(scale)[i] = static_cast<real>(s[i]);
}

template <typename real, int KP, int NQ>
int launch_pass(const PredictArgs& a, cudaStream_t st) {
    dim3 grid((a.n_active + kPredWarps * 32 - 1) / (kPredWarps * 32), a.s_splits);
    const size_t smem = a.theta ? 2 * static_cast<size_t>(KP + 4) * PredTile<real, KP>::value * sizeof(real) : 16;
    auto kern = predict_pass_kernel<real, KP, NQ>;
    if (smem > 32 * 1024) BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPredWarps * 32, smem, st>>>(a);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

template <typename real, int KP>
int dispatch_nq(const PredictArgs& a, cudaStream_t st) {
    if (a.nq <= 3) return launch_pass<real, KP, 3>(a, st);
    if (a.nq <= 5) return launch_pass<real, KP, 5>(a, st);
    return launch_pass<real, KP, 8>(a, st);
}


template <int KP, int NQ>
int launch_pass_tc(const PredictArgs& a, const unsigned char* image, cudaStream_t st) {
    dim3 grid((a.n_active + kTcRows - 1) / kTcRows, a.s_splits / kTcSlotsPerBlock);
    auto kern = predict_pass_tc_kernel<KP, NQ>;
    BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcImage<KP>::kSmemBytes));
    kern<<<grid, kTcThreads + 32, TcImage<KP>::kSmemBytes, st>>>(a, image);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

template <int KP>
int dispatch_nq_tc(const PredictArgs& a, const unsigned char* image, cudaStream_t st) {
    if (a.nq <= 3) return launch_pass_tc<KP, 3>(a, image, st);
    if (a.nq <= 5) return launch_pass_tc<KP, 5>(a, image, st);
    return launch_pass_tc<KP, 8>(a, image, st);
}

template <typename real>
int run_predict(const bmc_predict_problem* p, double* mean, double* var, double* quant, int64_t* c_lt,
                int64_t* c_le, double* draws_out, int64_t ld_out, void* workspace, size_t workspace_bytes,
                int* passes_out, cudaStream_t st) {
    const size_t sz = sizeof(real);
    const QuantPlan plan = make_plan(p->probs, p->nq, p->n_draws);
    const int kp = bmc_padded_components(p->k > 0 ? p->k : 1);
    // wide bases in fp32: the contraction goes to the tensor cores (predict_tc_kernels.cuh)
    const bool tc = sizeof(real) == 4 && p->theta && use_tensor_path(kp, p->tensor_min_k) &&
                    p->noise_mode != BMC_NOISE_EXTERNAL;
    const int sms = sm_count();
    const int kt = std::max(kp, 8);               // operand width of the tensor path (one MMA spans 8 components)
    const long long tiles_tc = (p->n_draws + kTcTile - 1) / kTcTile;
    const size_t image_bytes = tc ? static_cast<size_t>(tiles_tc) * (2 * kt * 128 * 4 + kTcTile * 4) : 0;
    auto total_for = [&](long long n, const PassShape& sh) {
        return make_layout(n, p->nq, sh.cand_stride, sz, image_bytes).total;
    };
    // largest chunk of nuclei the workspace can hold (the shape depends on the chunk size)
    long long nc = std::min<long long>(p->n_points, kChunkPoints);
    PassShape shape = make_shape(nc, p->n_draws, plan.max_expected, tc, sms);
    while (nc > 32 && total_for(nc, shape) > workspace_bytes) {
        nc = std::max<long long>(32, nc / 2);
        shape = make_shape(nc, p->n_draws, plan.max_expected, tc, sms);
    }
    if (total_for(nc, shape) > workspace_bytes) {
        set_error("bmc_predict_fused: workspace of %zu bytes is too small (need %zu for %lld nuclei)",
                  workspace_bytes, total_for(nc, shape), nc);
        return BMC_ERR_WORKSPACE;
    }
    // equal chunks (a short last chunk would run at a fraction of the machine)
    {
        const long long n_chunks = (p->n_points + nc - 1) / nc;
        const long long even = ((p->n_points + n_chunks - 1) / n_chunks + 255) / 256 * 256;
        if (even < nc) {
            const PassShape es = make_shape(even, p->n_draws, plan.max_expected, tc, sms);
            if (total_for(even, es) <= workspace_bytes) {
                nc = even;
                shape = es;
            }
        }
    }
    const Layout lay = make_layout(nc, p->nq, shape.cand_stride, sz, image_bytes);
    if (lay.total > workspace_bytes) {
        set_error("bmc_predict_fused: workspace of %zu bytes is too small (need %zu)", workspace_bytes, lay.total);
        return BMC_ERR_WORKSPACE;
    }
    unsigned char* ws = static_cast<unsigned char*>(workspace);

    // quantile plan to the device
    PlanImage* d_plan = reinterpret_cast<PlanImage*>(ws + lay.plan);
    long long* d_rank = d_plan->rank;
    double* d_frac = d_plan->frac;
    double* d_zq = d_plan->zq;
    double* d_hw = d_plan->hw;
    {
        PlanImage img{};
        std::memcpy(img.rank, plan.rank, sizeof(img.rank));
        std::memcpy(img.frac, plan.frac, sizeof(img.frac));
        std::memcpy(img.zq, plan.zq, sizeof(img.zq));
        std::memcpy(img.hw, plan.hw, sizeof(img.hw));
        store_plan_kernel<<<1, 32, 0, st>>>(img, d_plan);     // by value in the launch: nothing to wait for
        BMC_LAUNCH_CHECK();
    }

    auto select_kern = predict_select_kernel<real>;
    BMC_CUDA(cudaFuncSetAttribute(select_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kSelStageBytes));
    const int tiles = tc ? static_cast<int>(tiles_tc) * kTcSlotsPerBlock
                         : static_cast<int>((p->n_draws + kPredTile - 1) / kPredTile);
    const unsigned char* image = tc ? ws + lay.image : nullptr;
    if (tc) {
        const float* th = static_cast<const float*>(p->theta);
        const unsigned gt = static_cast<unsigned>(tiles_tc);
        switch (kt) {
            case 8: theta_image_kernel<8><<<gt, 256, 0, st>>>(th, p->n_draws, kp + 4, kp, ws + lay.image); break;
            case 16: theta_image_kernel<16><<<gt, 256, 0, st>>>(th, p->n_draws, kp + 4, kp, ws + lay.image); break;
            case 32: theta_image_kernel<32><<<gt, 256, 0, st>>>(th, p->n_draws, kp + 4, kp, ws + lay.image); break;
            default: theta_image_kernel<64><<<gt, 256, 0, st>>>(th, p->n_draws, kp + 4, kp, ws + lay.image); break;
        }
        BMC_LAUNCH_CHECK();
    }

    int max_passes = 0;
    for (long long c0 = 0; c0 < p->n_points; c0 += nc) {
        const long long n = std::min<long long>(nc, p->n_points - c0);
        const size_t nqn = static_cast<size_t>(n) * p->nq;
        int* d_counter = reinterpret_cast<int*>(ws + lay.counter);
        BMC_CUDA(cudaMemsetAsync(ws + lay.c_lt, 0, n * 4, st));
        BMC_CUDA(cudaMemsetAsync(ws + lay.c_le, 0, n * 4, st));
        BMC_CUDA(cudaMemsetAsync(ws + lay.flag, 0, static_cast<size_t>(n) * 4, st));
        BMC_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), st));

        const unsigned gb = static_cast<unsigned>((n + 255) / 256);
        const real* u_chunk = static_cast<const real*>(p->u) + c0 * p->k;
        if (p->center && p->scale) {
            copy_guess_kernel<real><<<gb, 256, 0, st>>>(p->center + c0, p->scale + c0, n, ws + lay.center,
                                                        ws + lay.scale);
        } else {
            const size_t gs = (static_cast<size_t>(p->k + 1) + 32) * (p->k + 2) * sizeof(double);
            if (gs > 32 * 1024)
                BMC_CUDA(cudaFuncSetAttribute(predict_guess_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(gs)));
            predict_guess_kernel<real><<<static_cast<unsigned>((n + 31) / 32), 256, gs, st>>>(
                u_chunk, n, p->k, p->theta_mean, p->theta_cov, p->noise_mode != BMC_NOISE_NONE, ws + lay.center,
                ws + lay.scale);
        }
        BMC_LAUNCH_CHECK();
        predict_window_kernel<real><<<static_cast<unsigned>((nqn + 255) / 256), 256, 0, st>>>(
            n, p->nq, ws + lay.center, ws + lay.scale, d_zq, d_hw, ws + lay.win_lo, ws + lay.win_hi, ws + lay.brk_lo,
            ws + lay.brk_hi, ws + lay.pair_hi, ws + lay.aux, ws + lay.phase,
            reinterpret_cast<unsigned int*>(ws + lay.cnt_below), reinterpret_cast<unsigned int*>(ws + lay.sub),
            ws + lay.resolved);
        BMC_LAUNCH_CHECK();

        // the first pass of a short last chunk may use more sample splits than the full-size chunk
        PassShape cs = make_shape(n, p->n_draws, plan.max_expected, tc, sms);
        if (cs.cand_stride > shape.cand_stride) cs = shape;
        if (cs.s_splits * segment_len(plan.max_expected, cs.s_splits) > shape.cand_stride) cs = shape;

        PredictArgs a{};
        a.u = u_chunk;
        a.mu = p->mu ? p->mu + c0 : nullptr;
        a.truth = p->truth ? p->truth + c0 : nullptr;
        a.n_points = n;
        a.point0 = p->point0 + static_cast<unsigned long long>(c0);
        a.theta = p->theta;
        a.n_draws = p->n_draws;
        a.k = p->k;
        a.noise_mode = p->noise_mode;
        a.key0 = static_cast<uint32_t>(p->seed);
        a.key1 = static_cast<uint32_t>(p->seed >> 32);
        for (int r = 0; r < 10; ++r) {
            a.keys.k0[r] = a.key0 + static_cast<uint32_t>(r) * kPhiloxW0;
            a.keys.k1[r] = a.key1 + static_cast<uint32_t>(r) * kPhiloxW1;
        }
        a.noise = p->noise ? static_cast<const unsigned char*>(p->noise) + static_cast<size_t>(c0) * sz : nullptr;
        a.ld_noise = p->ld_noise;
        a.nq = p->nq;
        a.win_lo = ws + lay.win_lo;
        a.win_hi = ws + lay.win_hi;
        a.cnt_below = reinterpret_cast<unsigned int*>(ws + lay.cnt_below);
        a.cnt_slot = reinterpret_cast<unsigned int*>(ws + lay.cnt_slot);
        a.sub_cnt = reinterpret_cast<unsigned int*>(ws + lay.sub);
        a.cand = ws + lay.cand;
        a.cand_stride = shape.cand_stride;
        a.seg_len = shape.cand_stride / cs.s_splits;
        a.count_slices = 0;
        a.first = 1;
        a.center = ws + lay.center;
        a.mom_part = reinterpret_cast<double*>(ws + lay.mom);
        a.c_lt = reinterpret_cast<unsigned int*>(ws + lay.c_lt);
        a.c_le = reinterpret_cast<unsigned int*>(ws + lay.c_le);
        a.draws_out = draws_out ? draws_out + c0 : nullptr;
        a.ld_out = ld_out;
        a.point_list = nullptr;
        a.n_active = static_cast<int>(n);
        a.s_splits = cs.s_splits;

        SelectArgs s{};
        s.n_points = n;
        s.n_draws = p->n_draws;
        s.nq = p->nq;
        s.rank = d_rank;
        s.frac = d_frac;
        s.win_lo = a.win_lo;
        s.win_hi = a.win_hi;
        s.brk_lo = ws + lay.brk_lo;
        s.brk_hi = ws + lay.brk_hi;
        s.pair_hi = ws + lay.pair_hi;
        s.aux = ws + lay.aux;
        s.phase = ws + lay.phase;
        s.cnt_below = a.cnt_below;
        s.cnt_slot = a.cnt_slot;
        s.sub_cnt = a.sub_cnt;
        s.slices_valid = 0;
        s.cand = a.cand;
        s.cand_stride = a.cand_stride;
        s.seg_len = a.seg_len;
        s.n_slots = a.s_splits;
        s.resolved = ws + lay.resolved;
        s.mu = a.mu;
        s.out_quant = quant;
        s.ld_quant = p->n_points;
        s.out_offset = c0;
        s.first = 1;
        s.mom_part = a.mom_part;
        s.center = a.center;
        s.out_mean = mean;
        s.out_var = var;
        s.c_lt = a.c_lt;
        s.c_le = a.c_le;
        s.out_c_lt = p->truth ? reinterpret_cast<long long*>(c_lt) : nullptr;
        s.out_c_le = p->truth ? reinterpret_cast<long long*>(c_le) : nullptr;
        s.point_list = nullptr;
        s.n_active = static_cast<int>(n);
        s.point_flag = reinterpret_cast<int*>(ws + lay.flag);
        int* list_cur = reinterpret_cast<int*>(ws + lay.list_a);
        int* list_next = reinterpret_cast<int*>(ws + lay.list_b);
        s.next_list = list_next;
        s.next_count = d_counter;

        int pass = 0;
        for (;;) {
            ++pass;
            int rc = BMC_OK;
            if (tc) {
                switch (kt) {
                    case 8: rc = dispatch_nq_tc<8>(a, image, st); break;
                    case 16: rc = dispatch_nq_tc<16>(a, image, st); break;
                    case 32: rc = dispatch_nq_tc<32>(a, image, st); break;
                    default: rc = dispatch_nq_tc<64>(a, image, st); break;
                }
            } else {
                switch (kp) {
                    case 4: rc = dispatch_nq<real, 4>(a, st); break;
                    case 8: rc = dispatch_nq<real, 8>(a, st); break;
                    case 16: rc = dispatch_nq<real, 16>(a, st); break;
                    case 32: rc = dispatch_nq<real, 32>(a, st); break;
                    default: rc = dispatch_nq<real, 64>(a, st); break;
                }
            }
            if (rc != BMC_OK) return rc;
            const long long items = static_cast<long long>(s.n_active) * p->nq;
            select_kern<<<static_cast<unsigned>((items + 7) / 8), 256, 8 * kSelStageBytes, st>>>(s);
            BMC_LAUNCH_CHECK();
            int pending = 0;
            BMC_CUDA(cudaMemcpyAsync(&pending, d_counter, sizeof(int), cudaMemcpyDeviceToHost, st));
            BMC_CUDA(cudaStreamSynchronize(st));
            if (pending == 0) break;
            if (pass >= kMaxPasses) {
                set_error("bmc_predict_fused: %d nuclei unresolved after %d passes", pending, pass);
                return BMC_ERR_CONVERGE;
            }
            // next pass: only the nuclei that asked for it, with slice counting on
            std::swap(list_cur, list_next);
            BMC_CUDA(cudaMemsetAsync(ws + lay.flag, 0, static_cast<size_t>(n) * 4, st));
            BMC_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), st));
            a.first = 0;
            a.draws_out = nullptr;
            a.point_list = list_cur;
            a.n_active = pending;
            a.count_slices = 1;
            a.s_splits = std::min(shape.retry_splits, tiles);
            a.seg_len = shape.cand_stride / a.s_splits;
            s.first = 0;
            s.point_list = list_cur;
            s.n_active = pending;
            s.next_list = list_next;
            s.slices_valid = 1;
            s.n_slots = a.s_splits;
            s.seg_len = a.seg_len;
        }
        max_passes = std::max(max_passes, pass);
    }
    if (passes_out) *passes_out = max_passes;
    return BMC_OK;
}

}  // namespace

extern "C" {

size_t bmc_predict_workspace_bytes(int dtype, int64_t n_points, int nq, int64_t n_draws) {
    if (n_points < 1 || nq < 1 || nq > kMaxQuant || n_draws < 1) return 0;
    double probs[kMaxQuant];
    for (int j = 0; j < nq; ++j) probs[j] = 50.0;   // the median has the widest window
    const QuantPlan plan = make_plan(probs, nq, n_draws);
    const size_t sz = dtype == BMC_F32 ? 4 : 8;
    // chunks of 32768 nuclei keep the candidate buffers in the hundreds of MB
    const long long nc = std::min<long long>(n_points, kChunkPoints);
    const int sms = sm_count();
    const PassShape shape = make_shape(nc, n_draws, plan.max_expected, false, sms);
    size_t need = make_layout(nc, nq, shape.cand_stride, sz).total;
    if (dtype == BMC_F32) {
        // room for the tensor-core path (K > 16): its slot layout and the operand image of the draws
        const PassShape ts = make_shape(nc, n_draws, plan.max_expected, true, sms);
        const size_t image = static_cast<size_t>((n_draws + kTcTile - 1) / kTcTile) * TcImage<64>::kStride;
        need = std::max(need, make_layout(nc, nq, ts.cand_stride, sz, image).total);
    }
    return need;
}

int bmc_predict_theta_stride(int k) { return bmc_padded_components(k > 0 ? k : 1) + 4; }

int bmc_predict_fused(int dtype, const bmc_predict_problem* p, double* mean, double* var, double* quant,
                      int64_t* c_lt, int64_t* c_le, double* draws_out, int64_t ld_out, void* workspace,
                      size_t workspace_bytes, int* passes_out, void* stream) {
    BMC_REQUIRE(p, "bmc_predict_fused: problem is NULL");
    BMC_REQUIRE(dtype == BMC_F32 || dtype == BMC_F64, "bmc_predict_fused: bad dtype %d", dtype);
    BMC_REQUIRE(p->n_points >= 1 && p->n_draws >= 1, "bmc_predict_fused: n_points=%lld n_draws=%lld",
                (long long)p->n_points, (long long)p->n_draws);
    BMC_REQUIRE(p->n_draws < (1ll << 31), "bmc_predict_fused: n_draws must fit 31 bits");
    BMC_REQUIRE(p->nq >= 1 && p->nq <= kMaxQuant && p->probs, "bmc_predict_fused: nq=%d", p->nq);
    for (int j = 0; j < p->nq; ++j)
        BMC_REQUIRE(p->probs[j] >= 0.0 && p->probs[j] <= 100.0, "Percentiles must be in the range [0, 100]");
    BMC_REQUIRE(mean && var && quant && workspace, "bmc_predict_fused: null output");
    BMC_REQUIRE(!p->truth || (c_lt && c_le), "bmc_predict_fused: truth given without count outputs");
    if (p->theta) {
        BMC_REQUIRE(p->k >= 0 && p->k <= BMC_MAX_COMPONENTS, "bmc_predict_fused: k=%d outside 0..%d", p->k,
                    BMC_MAX_COMPONENTS);
        BMC_REQUIRE(p->k == 0 || p->u, "bmc_predict_fused: u is NULL");
        BMC_REQUIRE((p->theta_mean && p->theta_cov) || (p->center && p->scale),
                    "bmc_predict_fused: need theta moments or a centre/scale guess");
    } else {
        BMC_REQUIRE(p->noise_mode == BMC_NOISE_EXTERNAL && p->noise && p->center && p->scale,
                    "bmc_predict_fused: matrix mode needs noise, center and scale");
    }
    BMC_REQUIRE(p->noise_mode != BMC_NOISE_EXTERNAL || (p->noise && p->ld_noise >= p->n_points),
                "bmc_predict_fused: external noise needs a pointer and ld_noise >= n_points");
    BMC_REQUIRE(!draws_out || ld_out >= p->n_points, "bmc_predict_fused: ld_out < n_points");
    cudaStream_t st = as_stream(stream);
    return dtype == BMC_F32 ? run_predict<float>(p, mean, var, quant, c_lt, c_le, draws_out, ld_out, workspace,
                                                 workspace_bytes, passes_out, st)
                            : run_predict<double>(p, mean, var, quant, c_lt, c_le, draws_out, ld_out, workspace,
                                                  workspace_bytes, passes_out, st);
}

int bmc_coverage_counts(const double* matrix, int64_t s_rows, int64_t n_cols, int64_t ld, const double* truth,
                        int64_t* c_lt, int64_t* c_le, void* stream) {
    BMC_REQUIRE(matrix && truth && c_lt && c_le, "bmc_coverage_counts: null pointer");
    BMC_REQUIRE(s_rows >= 1 && n_cols >= 1 && ld >= n_cols, "bmc_coverage_counts: bad shape");
    cudaStream_t st = as_stream(stream);
    BMC_CUDA(cudaMemsetAsync(c_lt, 0, sizeof(int64_t) * n_cols, st));
    BMC_CUDA(cudaMemsetAsync(c_le, 0, sizeof(int64_t) * n_cols, st));
    const unsigned gx = static_cast<unsigned>((n_cols + 255) / 256);
    // enough row blocks to cover the machine a few times over
    long long gy = std::max<long long>(1, std::min<long long>((sm_count() * 8 + gx - 1) / gx, (s_rows + 63) / 64));
    const long long rows_per_block = (s_rows + gy - 1) / gy;
    gy = (s_rows + rows_per_block - 1) / rows_per_block;
    coverage_counts_kernel<<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, st>>>(
        matrix, s_rows, n_cols, ld, truth, rows_per_block, reinterpret_cast<unsigned long long*>(c_lt),
        reinterpret_cast<unsigned long long*>(c_le));
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

int bmc_column_moments(const double* matrix, int64_t s_rows, int64_t n_cols, int64_t ld, double* center,
                       double* scale, void* stream) {
    BMC_REQUIRE(matrix && center && scale, "bmc_column_moments: null pointer");
    BMC_REQUIRE(s_rows >= 1 && n_cols >= 1 && ld >= n_cols, "bmc_column_moments: bad shape");
    column_moments_kernel<<<static_cast<unsigned>((n_cols + 255) / 256), 256, 0, as_stream(stream)>>>(
        matrix, s_rows, n_cols, ld, center, scale);
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

int bmc_coverage_levels(const int64_t* c_lt, const int64_t* c_le, int64_t n_points, const int64_t* lo_idx,
                        const int64_t* hi_idx, int n_levels, int64_t* covered, void* stream) {
    BMC_REQUIRE(c_lt && c_le && lo_idx && hi_idx && covered, "bmc_coverage_levels: null pointer");
    BMC_REQUIRE(n_points >= 1 && n_levels >= 1, "bmc_coverage_levels: bad shape");
    cudaStream_t st = as_stream(stream);
    BMC_CUDA(cudaMemsetAsync(covered, 0, sizeof(int64_t) * n_levels, st));
    const unsigned gx = static_cast<unsigned>(std::min<long long>((n_points + 255) / 256, sm_count() * 4));
    coverage_levels_kernel<<<dim3(gx, n_levels), 256, 0, st>>>(
        reinterpret_cast<const long long*>(c_lt), reinterpret_cast<const long long*>(c_le), n_points,
        reinterpret_cast<const long long*>(lo_idx), reinterpret_cast<const long long*>(hi_idx), n_levels,
        reinterpret_cast<unsigned long long*>(covered));
    BMC_LAUNCH_CHECK();
    return BMC_OK;
}

}  // extern "C"
