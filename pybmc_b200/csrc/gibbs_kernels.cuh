// Batched samplers in sufficient-statistic form.
//
//   gibbs_conjugate_kernel        one chain per thread   pybmc/inference_utils.py:39-54 (gibbs_sampler loop)
//   gibbs_conjugate_group_kernel  eight lanes (or a warp) per chain  same loop, for fewer than 16,384 (1,536) chains
//   gibbs_simplex_kernel          one chain per thread   pybmc/inference_utils.py:97-141 (gibbs_sampler_simplex)
//   gibbs_simplex_group_kernel    eight lanes per chain  same loops, for fewer than 16,384 chains
//   gibbs_simplex_group16_kernel  eight lanes per chain, at most 16 models: the state-independent parts of a
//                                 proposal (weight changes, G delta) precomputed 32 iterations at a time
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

#ifndef BMC_F32_MIN_BLOCKS
// Blocks of 128 threads per SM the fp32 thread-per-chain kernels (K <= 16) are compiled for.  The persistent launch
// runs three warps per scheduler, so the kernel may have 168 registers instead of the 128 that four warps allowed:
// 65,536 x 10,000 with histograms 8.21 -> 7.89 ms, without 8.21 -> 7.67 ms (255 registers / two warps: 8.11 ms).
#define BMC_F32_MIN_BLOCKS 3
#endif


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
    RunConsts<float> cf;                  // the same scalars + Gamma constants, ready in either type
    RunConsts<double> cd;
    int gamma_boost;                      // shape < 1
    uint32_t key0, key1;
    PhiloxKeys keys;                      // the ten round keys of (key0, key1): constant-bank operands in the kernels
    unsigned long long chain0;            // global id of this launch's first chain
    long long n_chains;
    long long iterations;
    long long store_from, thin, n_kept;
    void* samples;                        // [n_kept][k+1][n_chains]  real, nullable
    double* chain_stats;                  // [n_stat][n_chains]       nullable
    int stats_mode;                       // 0 none, 1 diagonal second moments, 2 full
    // marginal histograms of [b_0..b_{k-1}, sigma]: the state after every
    // hist_every-th iteration is binned, per block in shared memory, merged into `hist` by 64-bit atomics
    uint32_t hist_every;                  // 0 = off
    const double* hist_lo;                // [k+1] lower edge of bin 0
    const double* hist_inv;               // [k+1] 1 / bin width
    unsigned long long* hist;             // [hist_replicas][k+1][kHistBins]; block b merges into replica b % hist_replicas
    int hist_replicas;
    // work items of the thread-per-chain kernel (see gibbs_conjugate_kernel): 0 = one item per warp, the whole run
    int item_order;                       // 0: items w, w + W, ... of the (block, group) list; 1: owned groups + rotation
    uint32_t item_its;                    // iterations per item, a multiple of kFlushEvery
    unsigned* item_done;                  // [groups] items finished per group of 32 chains (zeroed by the host)
    double* item_state;                   // [n_chains] sigma^2 handed from one item of a chain to the next
};

constexpr int kHistBins = BMC_HIST_BINS;

// bin of value v on coordinate c; values outside the range (and NaN) land in the edge bins
template <typename real>
__device__ __forceinline__ unsigned hist_bin(const double* lo, const double* inv, int c, real v) {
    real t = (v - static_cast<real>(lo[c])) * static_cast<real>(inv[c]);
    t = fmin(fmax(t, real(0)), real(kHistBins - 1));
    return static_cast<unsigned>(static_cast<int>(t));
}

// the same with the edges already in the kernel's arithmetic type (shared-memory copies, see gibbs_conjugate_kernel)
template <typename real>
__device__ __forceinline__ unsigned hist_bin_typed(real lo, real inv, real v) {
    real t = (v - lo) * inv;
    t = fmin(fmax(t, real(0)), real(kHistBins - 1));
    return static_cast<unsigned>(static_cast<int>(t));
}

// block-level histogram (dynamic shared memory, uint32 [coords][kHistBins]) -> global 64-bit counts
__device__ __forceinline__ void hist_zero(unsigned* hist_s, int coords) {
    for (int i = threadIdx.x; i < coords * kHistBins; i += blockDim.x) hist_s[i] = 0u;
    __syncthreads();
}
// The blocks of a launch finish together: merging straight into one histogram would put ~1000 atomics on each
// of its words at once (measured: +0.9 ms on a 9.3 ms launch).  Block b adds to replica b % replicas instead;
// hist_reduce_kernel sums the replicas afterwards.
// `n_active`: the threads 0 .. n_active-1 of the block are the ones that did not return early (threads past the
// last chain leave right after the zeroing barrier; from Volta on __syncthreads() waits for the non-exited
// threads only), so the merge loop strides over them and still covers every bin.
__device__ __forceinline__ void hist_merge(const unsigned* hist_s, int coords, unsigned long long* hist, int replicas,
                                           int n_active) {
    __syncthreads();
    unsigned long long* mine = hist + static_cast<size_t>(blockIdx.x % replicas) * coords * kHistBins;
    for (int i = threadIdx.x; i < coords * kHistBins; i += n_active) {
        const unsigned v = hist_s[i];
        if (v) atomicAdd(mine + i, static_cast<unsigned long long>(v));
    }
}
static __global__ void hist_reduce_kernel(const unsigned long long* replicas, int n_replicas, int words,
                                   unsigned long long* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= words) return;
    unsigned long long s = 0;
    for (int r = 0; r < n_replicas; ++r) s += replicas[static_cast<size_t>(r) * words + i];
    counts[i] = s;
}

constexpr int kFlushEvery = 64;           // iterations between fp64 flushes of the moment sums

// Add to a slot of the fp64 moment rows WITHOUT reading it back: a reduction (RED.E.ADD.F64) instead of
// load - add - store.  A slot belongs to one thread, so the sum is the same either way; but `*p += v` made a
// flush a chain of 54 dependent L2 round trips (the compiler may not move a load above the previous store:
// LDG DADD STG x 54 in SASS, ~270 cycles each = 12 % of the kernel's stall samples as long_scoreboard), where
// the reduction is 54 fire-and-forget instructions.
__device__ __forceinline__ void stat_add(double* p, double v) { atomicAdd(p, v); }

template <typename real, typename Args>
__device__ __forceinline__ const RunConsts<real>& run_consts(const Args& a) {
    if constexpr (sizeof(real) == 4) return a.cf;
    else return a.cd;
}

template <int KP, int MODE>
struct StatCount {
    static constexpr int D = KP + 1;
    static constexpr int value = MODE == 0 ? 0 : (MODE == 1 ? 2 * D : D + D * (D + 1) / 2);
};

// Work items.  A launch is cut into items of (32 chains) x (item_its iterations).  Worker (warp) w of the W in the
// grid owns chain groups w, w + W, ...; the G mod W groups left over go round the workers -- in iteration block b
// the j-th of them belongs to worker (j + (G mod W) b) mod W -- and a group's state (sigma^2: nothing else survives
// an iteration) passes from one item to the next through global memory and a per-group counter.  (First version:
// items w, w + W, ... of the list ordered by (block, group): nearly every item then waits for one that another
// worker started at the same moment, the workers fall into lock-step and 9 % of the stall samples were the wait.)  Why: one chain per thread gives 65,536 chains 2048 warps, 3.46 per scheduler -- four on
// 46 % of the schedulers, three on the others, and the launch ran at the pace of the fours (9.06 ms, the same as
// 75,776 chains; 56,832 chains, three everywhere, take 6.99 ms: profiles/r2_notes.md); the fp64 kernel (255
// registers, two warps per scheduler) paid two full waves for 1.73.  With every resident warp slot filled by a
// persistent worker and the chain groups rotating through them, every scheduler carries the same average load
// and the tail disappears.  item_its = 0 (or a grid that covers all groups with one item each) is the plain launch:
// no waiting, no state traffic.  The persistent form needs all W warps co-resident: the host launches it
// cooperatively (gibbs.cu).
//
// HIST: marginal histograms on (a second instantiation, so that the plain kernel's code is untouched by them).
template <typename real, int KP, int MODE, bool HIST = false>
__global__ void __launch_bounds__(128, (sizeof(real) == 4 && KP <= 16) ? BMC_F32_MIN_BLOCKS : 2) gibbs_conjugate_kernel(const GibbsArgs a) {
    using M = Math<real>;
    extern __shared__ unsigned hist_s[];                  // [k+1][kHistBins] when histograms are on
    // What the binning needs, once per block in the kernel's arithmetic type: W (zero-padded KP x KP; the diagonal
    // form on its diagonal), g_ols, the lower edges and 1 / widths.  The flush points used to read these as fp64 from
    // global memory and convert them every time -- 64 + 8 + 18 F2F on the XU pipe per flush in the fp32 kernel, one per
    // iteration on the pipe that is already at 57 %.  Same values, so the bins are the same.
    constexpr int HW = KP * KP, HG = HW + KP, HL = HG + KP + 1, HN = HL + KP + 1;
    __shared__ real hist_c[HIST ? HN : 1];
    if constexpr (HIST) {
        for (int i = threadIdx.x; i < HN; i += blockDim.x) {
            real v = real(0);
            if (i < HW) {
                const int r = i / KP, c = i % KP;
                if (r < a.k && c < a.k) {
                    if (a.dense_w) v = static_cast<real>(a.w[r * a.k + c]);
                    else if (r == c) v = static_cast<real>(a.w[r]);
                }
            } else if (i < HG) {
                if (i - HW < a.k) v = static_cast<real>(a.g_ols[i - HW]);
            } else if (i < HL) {
                if (i - HG <= a.k) v = static_cast<real>(a.hist_lo[i - HG]);
            } else {
                if (i - HL <= a.k) v = static_cast<real>(a.hist_inv[i - HL]);
            }
            hist_c[i] = v;
        }
        hist_zero(hist_s, a.k + 1);                       // ends with the block barrier
    }
    const uint32_t total = static_cast<uint32_t>(a.iterations);          // < 2^32 (checked by the host)
    const unsigned lane = threadIdx.x & 31u;
    const long long worker = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_workers = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long n_groups = (a.n_chains + 31) >> 5;
    const uint32_t item_its = a.item_its ? a.item_its : (total ? total : 1u);
    const long long n_blk = static_cast<long long>((total + item_its - 1u) / item_its);
    // worker w owns groups w, w + W, ... (q of them: their items never wait) and the `extra` = G - q W groups left
    // over go round: in iteration block b, left-over group j belongs to worker (j + extra b) mod W
    const long long own = n_groups / n_workers, extra = n_groups - own * n_workers;
    long long tid = 0;                                    // chain of the current item within this launch
    uint32_t chain = 0;                                   // its global id (low word: Philox counter)
    constexpr int D = KP + 1;
    constexpr int NS = StatCount<KP, MODE>::value;

    // (d and pull as constant-bank operands -- typed copies inside the kernel parameters -- were measured in
    //  round 2: 9.25 vs 9.28 ms, no gain for a synchronising read-back of the constants; profiles/r2_notes.md)
    // K >= 32: d and pull live in shared memory (every lane reads the same word: a broadcast), the thread keeps its
    // registers (255 of them: two blocks per SM) for e and the moment sums -- 64 + 64 + 64 + 130 values did not fit
    // in 128 registers (2.5 KB of spills per thread, round 1).
    constexpr bool BIGK = KP >= 32;
    __shared__ __align__(16) real cons_s[BIGK ? 2 * KP : 2];
    real d[BIGK ? 1 : KP], pull[BIGK ? 1 : KP];
    if constexpr (BIGK) {
        for (int k = threadIdx.x; k < KP; k += blockDim.x) {
            cons_s[k] = k < a.k ? static_cast<real>(a.d[k]) : real(0);
            cons_s[KP + k] = k < a.k ? static_cast<real>(a.pull[k]) : real(0);
        }
        __syncthreads();
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            d[k] = k < a.k ? static_cast<real>(a.d[k]) : real(0);
            pull[k] = k < a.k ? static_cast<real>(a.pull[k]) : real(0);
        }
    }
    const RunConsts<real>& rc = run_consts<real>(a);
    const real rss_min = rc.rss_min;
    const real prior_scale = rc.prior_scale;
    const real sig_ref = rc.sigma_ref;
    const GammaConst<real> gc = gamma_const_of(rc, a.gamma_boost);

    using VL = VariateLayout<KP>;                        // word stream of an iteration (rng.cuh)
    constexpr int NC = VL::kCalls;
    constexpr bool GIN = VL::kGammaInline;               // K = 8: the Gamma variates ride in the angle call

    // fp32: Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2) take two components at a time.
    // With all cross moments, the 45 + 9 sums are kept as register pairs (30 instructions instead of 54 for
    // K = 8): pair (i, c), c = 2i+1 .. KP, holds the products of e_2i and e_2i+1 with e_c (e_KP = sigma
    // deviation); pair i of `dg` holds their squares.  Every half is the same rounding as the scalar form.
    constexpr bool PACK2 = sizeof(real) == 4 && KP % 4 == 0;
    constexpr bool PACK = sizeof(real) == 4 && MODE == 2 && KP % 2 == 0;
    constexpr int H = KP / 2;
    constexpr int NX = PACK ? H * KP - H * (H - 1) : 1;
    f32x2 m1p[PACK ? H : 1], dg[PACK ? H : 1], cx[NX];
    float m1s = 0.f, m2s = 0.f;
    if (PACK) {
#pragma unroll
        for (int i = 0; i < H; ++i) m1p[i] = dg[i] = 0ull;
#pragma unroll
        for (int j = 0; j < NX; ++j) cx[j] = 0ull;
    }
    real acc[(NS > 0 && !PACK) ? NS : 1];
#pragma unroll
    for (int j = 0; j < ((NS > 0 && !PACK) ? NS : 1); ++j) acc[j] = real(0);
    auto row_of = [&](int r, int c) { return D + r * D - r * (r - 1) / 2 + (c - r); };
    auto flush = [&](int row, real v) {
        double* p = a.chain_stats + static_cast<long long>(row) * a.n_chains + tid;
        stat_add(p, static_cast<double>(v));
    };

    real* const out = static_cast<real*>(a.samples);
    const long long n_steps = a.item_order ? n_blk * (own + 1) : (n_blk * n_groups - worker + n_workers - 1) / n_workers;
    for (long long step = 0; step < n_steps; ++step) {
    long long blk, grp;
    if (a.item_order) {                              // owned groups, left-over groups rotating
        blk = step / (own + 1);
        const long long o = step - blk * (own + 1);
        grp = worker + o * n_workers;
        if (o == own) {
            const long long j = (worker + n_workers - (extra * blk) % n_workers) % n_workers;
            if (j >= extra) continue;
            grp = own * n_workers + j;
        }
    } else {                                         // items w, w + W, ... of the list ordered by (block, group)
        const long long item = worker + step * n_workers;
        blk = item / n_groups;
        grp = item - blk * n_groups;
    }
    tid = grp * 32 + lane;
    const bool live = tid < a.n_chains;
    chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(tid));
    // opaque to the compiler from here on: it used to re-derive the chain id from S2R SR_TID.X inside the loop
    // (once per pair of iterations, with the special-register latency) rather than keep it in a register
    asm volatile("" : "+r"(chain));
    const uint32_t it_begin = static_cast<uint32_t>(blk) * item_its;
    const uint32_t it_end = (total - it_begin > item_its) ? it_begin + item_its : total;
    real s2 = rc.sigma2_init;
    if (blk > 0) {
        // the group's previous item: wait for its counter, then read the state it left (L2: another SM wrote it)
        if (lane == 0) {
            unsigned seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.item_done + grp) : "memory");
                if (seen < static_cast<unsigned>(blk)) __nanosleep(256);
            } while (seen < static_cast<unsigned>(blk));
        }
        __syncwarp();
        if (live) s2 = static_cast<real>(__ldcg(a.item_state + tid));
    }
    real sig = M::sqrt(s2);
    real e[BIGK ? 1 : KP];
#pragma unroll
    for (int k = 0; k < (BIGK ? 1 : KP); ++k) e[k] = real(0);
    // first kept iteration at or after it_begin, and its slot
    long long next_store = -1, slot = 0;
    if (a.samples) {
        slot = static_cast<long long>(it_begin) > a.store_from
                   ? (static_cast<long long>(it_begin) - a.store_from + a.thin - 1) / a.thin : 0;
        next_store = a.store_from + slot * a.thin;
    }

    real gm_odd = real(1);                              // Gamma variate of iteration 2m+1, drawn with that of 2m
    Philox4 held{0u, 0u, 0u, 0u};
    // The iterations run in segments that end where something other than arithmetic happens (a flush
    // of the moment sums every kFlushEvery iterations, a kept draw, the end): the inner loop is pure
    // arithmetic with one 32-bit counter.
    // K >= 32.  One sweep over the components of iteration `it` from the state (s2v, sigv): e_k handed to `use` two at
    // a time and NOT kept -- 64 more live registers would not fit beside the moment sums; a kept draw repeats the
    // sweep of its iteration once more (same words, same arithmetic) to get its e back.  The Philox words are
    // drawn where they are used, eight pairs (two radius calls, one angle call) at a time; d and pull come from
    // shared memory.
    real s2_in = real(0), sig_in = real(0);
    auto big_update = [&](const uint32_t it, const real s2v, const real sigv, real& rss0, real& rss1, auto&& use) {
        constexpr int P = VL::kPairs, R = P / 4;
        f32x2 rssp = pack2(static_cast<float>(rss0), 0.f);
        const f32x2 s2b = pack2(static_cast<float>(s2v), static_cast<float>(s2v));
        const f32x2 sigb = pack2(static_cast<float>(sigv), static_cast<float>(sigv));
#pragma unroll
        for (int c = 0; c < P / 8; ++c) {
            const Philox4 r0 = philox4x32_10(it, static_cast<uint32_t>(2 * c), chain, kTagGibbs, a.keys);
            const Philox4 r1 = philox4x32_10(it, static_cast<uint32_t>(2 * c + 1), chain, kTagGibbs, a.keys);
            const Philox4 ra = philox4x32_10(it, static_cast<uint32_t>(R + c), chain, kTagGibbs, a.keys);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = 2 * (8 * c + q);
                const uint32_t rw = philox_word(q < 4 ? r0 : r1, q % 4);
                const uint32_t aw = philox_word(ra, q / 2);
                const uint32_t h = (q & 1) ? aw >> 16 : aw & 0xFFFFu;
                real e0, e1;
                if constexpr (sizeof(real) == 4) {
                    const f32x2 zp = M::box_muller2_h(rw, h);
                    const f32x2 dp = *reinterpret_cast<const f32x2*>(cons_s + k);
                    const f32x2 pp = *reinterpret_cast<const f32x2*>(cons_s + KP + k);
                    float t0, t1;
                    unpack2(add2(dp, s2b), t0, t1);
                    const f32x2 sdp = mul2(pack2(M::rsqrt(t0), M::rsqrt(t1)), sigb);
                    const f32x2 ep = mul2(sdp, fma2(pp, sdp, zp));
                    rssp = fma2(mul2(dp, ep), ep, rssp);
                    float f0, f1;
                    unpack2(ep, f0, f1);
                    e0 = f0;
                    e1 = f1;
                } else {
                    real z0, z1;
                    M::box_muller_h(rw, h, z0, z1);
                    const real d0 = cons_s[k], d1 = cons_s[k + 1];
                    const real sd0 = sigv * M::rsqrt(d0 + s2v), sd1 = sigv * M::rsqrt(d1 + s2v);
                    e0 = sd0 * M::fma(cons_s[KP + k], sd0, z0);
                    e1 = sd1 * M::fma(cons_s[KP + k + 1], sd1, z1);
                    rss0 = M::fma(d0 * e0, e0, rss0);
                    rss1 = M::fma(d1 * e1, e1, rss1);
                }
                use(k, e0, e1);
            }
        }
        if constexpr (sizeof(real) == 4) {
            float r0f, r1f;
            unpack2(rssp, r0f, r1f);
            rss0 = r0f;
            rss1 = r1f;
        }
    };
    uint32_t it32 = live ? it_begin : it_end;
    while (it32 < it_end) {
        uint32_t seg_end = (it32 | static_cast<uint32_t>(kFlushEvery - 1)) + 1u;
        if (seg_end > it_end || seg_end == 0u) seg_end = it_end;
        if (next_store >= static_cast<long long>(it32) && next_store < static_cast<long long>(seg_end))
            seg_end = static_cast<uint32_t>(next_store) + 1u;
        // one iteration; w = its Philox calls, gm = its Gamma(shape, 1) variate
        auto iterate = [&](const uint32_t it32, const Philox4 (&w)[NC], const real gm) {
            real rss0 = rss_min, rss1 = real(0);
            if constexpr (BIGK) {
                s2_in = s2;                      // a kept draw re-derives e from these (big_update below)
                sig_in = sig;
                big_update(it32, s2, sig, rss0, rss1, [&](int k, real e0, real e1) {
                    if constexpr (MODE != 0) {              // marginal moments (MODE 2 stops at K = 16)
                        acc[k] += e0;
                        acc[k + 1] += e1;
                        acc[D + k] = M::fma(e0, e0, acc[D + k]);
                        acc[D + k + 1] = M::fma(e1, e1, acc[D + k + 1]);
                    }
                });
            } else if constexpr (PACK2) {
                // two components at a time on packed fp32 instructions (same roundings as the scalar form)
                f32x2 rssp = pack2(rss_min, 0.f);
                const f32x2 s2b = pack2(s2, s2), sigb = pack2(sig, sig);
                constexpr int P = VL::kPairs;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint32_t ra = philox_word(w[p / 4], p % 4);
                    const uint32_t aw = philox_word(w[(P + p / 2) / 4], (P + p / 2) % 4);
                    const f32x2 zp = M::box_muller2_h(ra, (p & 1) ? aw >> 16 : aw & 0xFFFFu);
                    const int k = 2 * p;
                    const f32x2 dp = pack2(d[k], d[k + 1]);
                    float t0, t1;
                    unpack2(add2(dp, s2b), t0, t1);
                    const f32x2 sdp = mul2(pack2(M::rsqrt(t0), M::rsqrt(t1)), sigb);
                    const f32x2 ep = mul2(sdp, fma2(pack2(pull[k], pull[k + 1]), sdp, zp));
                    rssp = fma2(mul2(dp, ep), ep, rssp);
                    unpack2(ep, e[k], e[k + 1]);
                }
                unpack2(rssp, rss0, rss1);
            } else {
                real z[KP];
                normals_of_words<real, KP>(w, z);
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    // 1/sqrt(p_k) = sigma / sqrt(d_k + s2): one MUFU on the s2 -> e -> RSS -> s2 dependency
                    const real sd = sig * M::rsqrt(d[k] + s2);
                    e[k] = sd * M::fma(pull[k], sd, z[k]);      // pull/p + z/sqrt(p)
                    if (k & 1) rss1 = M::fma(d[k] * e[k], e[k], rss1);
                    else rss0 = M::fma(d[k] * e[k], e[k], rss0);
                }
            }
            const real scale = real(0.5) * (prior_scale + (rss0 + rss1));
            s2 = M::div(scale, gm);
            s2 = s2 > real(1e-6) ? s2 : real(1e-6);
            sig = M::sqrt(s2);

            if constexpr (PACK) {
                const float es = sig - sig_ref;
                const f32x2 esb = pack2(es, es);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const f32x2 ep = pack2(e[2 * i], e[2 * i + 1]);
                    m1p[i] = add2(m1p[i], ep);
                    dg[i] = fma2(ep, ep, dg[i]);
#pragma unroll
                    for (int c = 2 * i + 1; c <= KP; ++c) {
                        const int j = i * KP - i * (i - 1) + (c - 2 * i - 1);
                        cx[j] = fma2(ep, c < KP ? pack2(e[c < KP ? c : 0], e[c < KP ? c : 0]) : esb, cx[j]);
                    }
                }
                m1s += es;
                m2s = fmaf(es, es, m2s);
            } else if constexpr (BIGK) {
                if constexpr (MODE != 0) {
                    const real es = sig - sig_ref;
                    acc[KP] += es;
                    acc[D + KP] = M::fma(es, es, acc[D + KP]);
                }
            } else if (MODE != 0) {
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
            }
        };
        // iterations 2m and 2m+1 share one Gamma proposal pair: both variates are finished when 2m starts (one
        // rare branch for the pair, rng.cuh).  fp32 walks the iterations in pairs where the segment allows: the
        // Philox rounds of 2m+1 then sit in the same basic block as the MUFU work of 2m and overlap with it
        // (9.98 -> 9.32 ms); fp64 has no registers to spare for that and is faster one by one.
        // `held`: with the inline layout, the angle call of iteration 2m+1 (needed at 2m for the shared pair).
        // the Gamma variates of iterations it_even and it_even + 1 (a function of the words only)
        auto gammas_of = [&](const uint32_t it_even, const Philox4& last_even, const Philox4& last_odd, real& g0,
                             real& g1) {
            GammaPair<real> gp;
            if constexpr (GIN) gp = gamma_pair_inline<real>(last_even, last_odd);
            else gp = gamma_pair<real>(it_even, chain, kTagGibbs, a.keys);
            gamma_pair_finish<real>(gc, gp, it_even, chain, kTagGibbs, a.key0, a.key1, g0, g1);
        };
        if constexpr (sizeof(real) == 4 && KP <= 16) {
            if ((it32 & 1u) != 0u) {
                Philox4 w[NC];
                iteration_words<KP>(it32, chain, kTagGibbs, a.keys, w);
                iterate(it32, w, gm_odd);
                ++it32;
            }
            for (; it32 + 1u < seg_end; it32 += 2u) {
                Philox4 wa[NC], wb[NC];
                iteration_words<KP>(it32, chain, kTagGibbs, a.keys, wa);
                iteration_words<KP>(it32 + 1u, chain, kTagGibbs, a.keys, wb);
                real gm_even;
                gammas_of(it32, wa[NC - 1], wb[NC - 1], gm_even, gm_odd);
                iterate(it32, wa, gm_even);
                iterate(it32 + 1u, wb, gm_odd);
            }
            if (it32 < seg_end) {
                Philox4 w[NC];
                iteration_words<KP>(it32, chain, kTagGibbs, a.keys, w);
                if constexpr (GIN) held = philox4x32_10(it32 + 1u, static_cast<uint32_t>(NC - 1), chain, kTagGibbs, a.keys);
                real gm_even;
                gammas_of(it32, w[NC - 1], held, gm_even, gm_odd);
                iterate(it32, w, gm_even);
                ++it32;
            }
        } else {
            for (; it32 < seg_end; ++it32) {
                const bool odd = (it32 & 1u) != 0u;
                Philox4 w[NC];
                if constexpr (!BIGK) {
#pragma unroll
                    for (int b = 0; b < NC - (GIN ? 1 : 0); ++b)
                        w[b] = philox4x32_10(it32, static_cast<uint32_t>(b), chain, kTagGibbs, a.keys);
                }
                if constexpr (GIN) {
                    if (odd) {
                        w[NC - 1] = held;
                    } else {
                        w[NC - 1] = philox4x32_10(it32, static_cast<uint32_t>(NC - 1), chain, kTagGibbs, a.keys);
                        held = philox4x32_10(it32 + 1u, static_cast<uint32_t>(NC - 1), chain, kTagGibbs, a.keys);
                    }
                }
                real gm = gm_odd;
                if (!odd) gammas_of(it32, w[NC - 1], held, gm, gm_odd);
                iterate(it32, w, gm);
            }
        }

        // ---- end of a segment: it32 iterations are done
        if (MODE != 0 && ((it32 & static_cast<uint32_t>(kFlushEvery - 1)) == 0u || it32 == it_end)) {
            if constexpr (PACK) {
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    float lo, hi;
                    unpack2(m1p[i], lo, hi);
                    flush(2 * i, lo);
                    flush(2 * i + 1, hi);
                    unpack2(dg[i], lo, hi);
                    flush(row_of(2 * i, 2 * i), lo);             // hi repeats pair (i, 2i+1)'s second half
                    m1p[i] = dg[i] = 0ull;
#pragma unroll
                    for (int c = 2 * i + 1; c <= KP; ++c) {
                        const int j = i * KP - i * (i - 1) + (c - 2 * i - 1);
                        unpack2(cx[j], lo, hi);
                        flush(row_of(2 * i, c), lo);
                        flush(row_of(2 * i + 1, c), hi);
                        cx[j] = 0ull;
                    }
                }
                flush(KP, m1s);
                flush(row_of(KP, KP), m2s);
                m1s = m2s = 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    flush(j, acc[j]);
                    acc[j] = real(0);
                }
            }
        }
        if constexpr (HIST) if ((it32 % a.hist_every) == 0u) {
            // marginal histograms: bin b = W (g_ols + e) and sigma of this iteration (shared-memory atomics);
            // hist_every is a multiple of kFlushEvery, so this is always the end of a segment.  Unrolled over
            // the padded components so that the loads and conversions of all coordinates overlap.
            real bv[KP];
            if (a.dense_w) {
#pragma unroll
                for (int r = 0; r < KP; ++r) {
                    real b = real(0);                      // padded rows and columns of W are zero: they add nothing
#pragma unroll
                    for (int k = 0; k < KP; ++k) b = M::fma(hist_c[r * KP + k], hist_c[HW + k] + e[k], b);
                    bv[r] = b;
                }
            } else {
#pragma unroll
                for (int r = 0; r < KP; ++r) bv[r] = hist_c[r * KP + r] * (hist_c[HW + r] + e[r]);
            }
            unsigned bin[KP];
#pragma unroll
            for (int r = 0; r < KP; ++r) bin[r] = hist_bin_typed<real>(hist_c[HG + r], hist_c[HL + r], bv[r]);
            const unsigned bin_s = hist_bin_typed<real>(hist_c[HG + a.k], hist_c[HL + a.k], sig);
#pragma unroll
            for (int r = 0; r < KP; ++r)
                if (r < a.k) atomicAdd(hist_s + r * kHistBins + bin[r], 1u);
            atomicAdd(hist_s + a.k * kHistBins + bin_s, 1u);
        }
        if (static_cast<long long>(it32) - 1 == next_store) {
            // b = W (g_ols + e);  row layout [slot][component][chain] keeps lanes coalesced
            real* row = out + (slot * static_cast<long long>(a.k + 1)) * a.n_chains + tid;
            real ek[KP];                            // e of the iteration just done
            if constexpr (BIGK) {
                real t0 = real(0), t1 = real(0);
                big_update(it32 - 1u, s2_in, sig_in, t0, t1, [&](int k, real e0, real e1) {
                    ek[k] = e0;
                    ek[k + 1] = e1;
                });
            } else {
#pragma unroll
                for (int k = 0; k < KP; ++k) ek[k] = e[k];
            }
            if (a.dense_w) {
                for (int r = 0; r < a.k; ++r) {
                    real b = real(0);
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < a.k)
                            b = M::fma(static_cast<real>(a.w[r * a.k + k]),
                                       static_cast<real>(a.g_ols[k]) + ek[k], b);
                    row[static_cast<long long>(r) * a.n_chains] = b;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < a.k)
                        row[static_cast<long long>(k) * a.n_chains] =
                            static_cast<real>(a.w[k]) * (static_cast<real>(a.g_ols[k]) + ek[k]);
            }
            row[static_cast<long long>(a.k) * a.n_chains] = sig;
            ++slot;
            next_store += a.thin;
        }
    }
    if (it_end < total) {
        // hand the chains over: state first, then the group's counter (release: the stores of all lanes are
        // ordered before it by the warp barrier and the fence)
        if (live) __stcg(a.item_state + tid, static_cast<double>(s2));
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.item_done + grp), "r"(static_cast<unsigned>(blk) + 1u) : "memory");
        }
    }
    }   // items
    if constexpr (HIST) hist_merge(hist_s, a.k + 1, a.hist, a.hist_replicas, blockDim.x);
}

// --------------------------------------------------------------------------------------
// Conjugate sampler, eight lanes per chain: the layout for few chains (the reference's own use is ONE
// chain of 50,000 iterations, which a single thread would walk through alone).  As in the simplex
// group kernel below, the state-independent variates of 32 consecutive iterations are generated in
// parallel by the group (lane g: iterations base + g + 8t) and parked in shared memory; the 32 state
// updates then run in order with lane g owning component g and RSS summed by three shuffles.
// G = 32 gives a chain the whole warp: the 32 iterations of a batch are generated one per lane (lanes
// 8..31 only generate; the state lives in lanes 0..7), for launches of so few chains that latency, not
// issue slots, is what is scarce -- the reference's own single chain of 50,000 iterations.
constexpr int kConjGroup = 8;
constexpr int kConjLanes = 8;             // lanes that hold state: one per component

template <typename real, int KP, int MODE, int GEN = kConjGroup, bool HIST = false>
__global__ void __launch_bounds__(128) gibbs_conjugate_group_kernel(const GibbsArgs a) {
    using M = Math<real>;
    static_assert(KP <= kConjLanes, "one lane per component");
    static_assert(GEN == 8 || GEN == 32, "eight lanes or a warp per chain");
    constexpr int G = GEN;
    constexpr int D = KP + 1;
    constexpr int WPB = 4, CPW = 32 / G;
    constexpr int ROW = KP + 1;                                           // z[KP], 1 / gamma
    __shared__ real draws_s[WPB * CPW * 32 * ROW];
    extern __shared__ unsigned hist_s[];                  // [k+1][kHistBins] when histograms are on
    if constexpr (HIST) hist_zero(hist_s, a.k + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & (G - 1), grp = lane / G;
    const long long cid_raw = (static_cast<long long>(blockIdx.x) * WPB + warp) * CPW + grp;
    const bool chain_ok = cid_raw < a.n_chains;
    if (!__any_sync(0xffffffffu, chain_ok)) return;
    const long long cid = chain_ok ? cid_raw : a.n_chains - 1;            // idle groups shadow a real chain
    const uint32_t chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(cid));
    real* const mine = draws_s + (static_cast<size_t>(warp) * CPW + grp) * 32 * ROW;

    const bool comp = g < a.k;
    const real d = comp ? static_cast<real>(a.d[g]) : real(0);
    const real pull = comp ? static_cast<real>(a.pull[g]) : real(0);
    const real g_ols = comp ? static_cast<real>(a.g_ols[g]) : real(0);
    const RunConsts<real>& rc = run_consts<real>(a);
    const real rss_min = rc.rss_min;
    const real prior_scale = run_consts<real>(a).prior_scale;
    const real sig_ref = run_consts<real>(a).sigma_ref;
    const GammaConst<real> gc = gamma_const_of(run_consts<real>(a), a.gamma_boost);

    real acc1 = real(0), accs = real(0), acce = real(0), accee = real(0);
    real acc2[KP];
#pragma unroll
    for (int c = 0; c < KP; ++c) acc2[c] = real(0);
    real s2 = run_consts<real>(a).sigma2_init;
    real sig = M::sqrt(s2);
    real* const out = static_cast<real*>(a.samples);
    const int total = static_cast<int>(a.iterations);                     // fits 31 bits (checked by the host)
    int next_store = a.samples ? static_cast<int>(a.store_from) : -1;
    int next_hist = HIST ? static_cast<int>(a.hist_every) - 1 : -1;
    int slot = 0;

    for (int base = 0; base < total; base += 32) {
        // ---- phase 1: the variates of iterations base .. base+31 of this group's chain
#pragma unroll 1
        for (int t = 0; t < 32 / G; ++t) {
            const int j = g + G * t;
            const uint32_t it32 = static_cast<uint32_t>(base + j);
            real* row = mine + j * ROW;
            Philox4 w[VariateLayout<KP>::kCalls];
            iteration_words<KP>(it32, chain, kTagGibbs, a.keys, w);
            real z[KP];
            normals_of_words<real, KP>(w, z);
#pragma unroll
            for (int q = 0; q < KP; ++q) row[q] = z[q];
            // the reciprocal is taken here, in the parallel phase (a full division in fp64)
            row[KP] = M::rcp(gamma_unit_scale_w<real, KP>(gc, w, it32, chain, kTagGibbs, a.keys, a.key0, a.key1));
        }
        __syncwarp();
        // ---- phase 2: the state updates, in order (:41-52)
        const int n_here = min(32, total - base);
        for (int j = 0; j < n_here; ++j) {
            const int it = base + j;
            const real* row = mine + j * ROW;
            const real z = g < KP ? row[g] : real(0);
            const real inv_gm = row[KP];
            const real sd = sig * M::rsqrt(d + s2);                      // 1/sqrt(d/s2 + 1)
            const real e = sd * M::fma(pull, sd, z);                      // pull/p + z/sqrt(p)
            real rss = (d * e) * e;
#pragma unroll
            for (int o = kConjLanes / 2; o > 0; o >>= 1) rss += __shfl_xor_sync(0xffffffffu, rss, o, kConjLanes);
            s2 = (real(0.5) * (prior_scale + (rss_min + rss))) * inv_gm;
            s2 = s2 > real(1e-6) ? s2 : real(1e-6);
            sig = M::sqrt(s2);
            if (MODE != 0) {
                const real es = sig - sig_ref;
                acc1 += e;
                accs = M::fma(e, es, accs);
                acce += es;
                accee = M::fma(es, es, accee);
                if (MODE == 1) {
                    acc2[0] = M::fma(e, e, acc2[0]);
                } else {
#pragma unroll
                    for (int c = 0; c < KP; ++c) acc2[c] = M::fma(e, __shfl_sync(0xffffffffu, e, c, G), acc2[c]);
                }
                if (((it + 1) % kFlushEvery) == 0 || it + 1 == total) {
                    auto row2 = [&](int r, int c) { return D + r * D - r * (r - 1) / 2 + (c - r); };
                    auto flush = [&](int row_out, real& v) {
                        if (chain_ok) {
                            double* q = a.chain_stats + static_cast<long long>(row_out) * a.n_chains + cid;
                            stat_add(q, static_cast<double>(v));
                        }
                        v = real(0);
                    };
                    if (g < KP) {
                        flush(g, acc1);
                        if (MODE == 1) {
                            flush(D + g, acc2[0]);
                        } else {
#pragma unroll
                            for (int c = 0; c < KP; ++c) {
                                if (c >= g) flush(row2(g, c), acc2[c]);
                                else acc2[c] = real(0);
                            }
                            flush(row2(g, KP), accs);
                        }
                    }
                    if (g == 0) {
                        flush(KP, acce);
                        flush(MODE == 1 ? D + KP : row2(KP, KP), accee);
                    }
                    acc1 = accs = acce = accee = real(0);
#pragma unroll
                    for (int c = 0; c < KP; ++c) acc2[c] = real(0);
                }
            }
            if (it == next_store || (HIST && it == next_hist)) {
                // b = W (g_ols + e): lane r gathers the group's coordinates for its row of W
                const real gam = g_ols + e;
                real b;
                if (a.dense_w) {
                    b = real(0);
#pragma unroll
                    for (int c = 0; c < KP; ++c) {
                        const real gc_ = __shfl_sync(0xffffffffu, gam, c, G);
                        if (comp && c < a.k) b = M::fma(static_cast<real>(a.w[g * a.k + c]), gc_, b);
                    }
                } else {
                    b = comp ? static_cast<real>(a.w[g]) * gam : real(0);
                }
                if (HIST && it == next_hist) {
                    if (chain_ok && g < kConjLanes) {
                        if (comp) atomicAdd(hist_s + g * kHistBins + hist_bin<real>(a.hist_lo, a.hist_inv, g, b), 1u);
                        if (g == 0)
                            atomicAdd(hist_s + a.k * kHistBins + hist_bin<real>(a.hist_lo, a.hist_inv, a.k, sig), 1u);
                    }
                    next_hist += static_cast<int>(a.hist_every);
                }
                if (it == next_store && chain_ok) {
                    real* dst = out + static_cast<long long>(slot) * (a.k + 1) * a.n_chains + cid;
                    if (comp) dst[static_cast<long long>(g) * a.n_chains] = b;
                    if (g == 0) dst[static_cast<long long>(a.k) * a.n_chains] = sig;
                }
                if (it == next_store) {
                    ++slot;
                    next_store += static_cast<int>(a.thin);
                }
            }
        }
        __syncwarp();
    }
    if constexpr (HIST) {
        // whole warps without a chain returned above: the first `warps` warps of the block are still here
        const long long left = a.n_chains - static_cast<long long>(blockIdx.x) * WPB * CPW;
        const int warps = left >= WPB * CPW ? WPB : static_cast<int>((left + CPW - 1) / CPW);
        hist_merge(hist_s, a.k + 1, a.hist, a.hist_replicas, 32 * warps);
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
    PhiloxKeys keys;        // round keys of (key0, key1)
    unsigned long long chain0;
    long long n_chains;
    long long burn, iterations;
    long long thin, n_kept;
    void* samples;          // [n_kept][k+1][n_chains] real, nullable
    double* chain_stats;    // [n_stat][n_chains] nullable (moments of b - b_ols and sigma - sigma_ref)
    int stats_mode;
    double sigma_ref, sigma2_init;
    int* accepted;          // [n_chains] sampling-phase acceptances        (:135)
    RunConsts<float> cf;    // scalars + Gamma constants in either arithmetic type (constant-bank operands)
    RunConsts<double> cd;
    int gamma_boost;
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
    const real prior_scale = run_consts<real>(a).prior_scale;
    const real sig_ref = run_consts<real>(a).sigma_ref;
    const GammaConst<real> gc = gamma_const_of(run_consts<real>(a), a.gamma_boost);

    // state: dc = b - b_ols (b starts at 0, :82), gdc = G dc, rss = RSS(b)
    real dc[KP], gdc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) dc[k] = -b_ols[k];
    real rss = run_consts<real>(a).rss_min;
#pragma unroll UR
    for (int r = 0; r < KP; ++r) {
        real s = real(0);
#pragma unroll UR
        for (int c = 0; c < KP; ++c) s = M::fma(gram_s[r * KP + c], dc[c], s);
        gdc[r] = s;
        rss = M::fma(s, dc[r], rss);
    }
    real s2 = run_consts<real>(a).sigma2_init;                               // :86, RSS(0)/n

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
        Philox4 w[VariateLayout<KP>::kCalls];
        iteration_words<KP>(it32, chain, kTagSimplex, a.keys, w);
        {
            real z[KP];
            normals_of_words<real, KP>(w, z);
#pragma unroll
            for (int q = 0; q < KP; ++q) delta[q] = step[q] * z[q];                       // :98 / :121
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
        const real u = M::u01(philox4x32_10(it32, kBlockUniform, chain, kTagSimplex, a.keys).x);
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
        const real gm = gamma_unit_scale_w<real, KP>(gc, w, it32, chain, kTagSimplex, a.keys, a.key0, a.key1);
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
                    stat_add(p, static_cast<double>(acc[j]));
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

// --------------------------------------------------------------------------------------
// Simplex sampler, eight lanes per chain: the layout for few chains (BASELINE configs[1]: 4096 chains
// are 128 warps if a thread owns a chain -- under one warp per SM -- and the loop is a chain of
// dependent fixed-latency instructions).  The random variates of an iteration do not depend on the
// chain's state, so each group first generates those of 32 consecutive iterations in parallel (lane g
// draws the normals, the Metropolis uniform and the Gamma variate of iterations base + g + 8t:
// perfectly uniform SIMT code), parks them in shared memory, and then walks the 32 state updates in
// order with the state spread over the group: lane g owns component g (dc_g, (G dc)_g) and evaluates
// the weights of models g, g + 8, ...; reductions are three shuffle steps inside the group.
// Four chains per warp, everything predicated (no divergence between the four).
constexpr int kSimplexGroup = 8;

template <typename real, int KP, int MODE>
__global__ void __launch_bounds__(128) gibbs_simplex_group_kernel(const SimplexArgs a) {
    using M = Math<real>;
    static_assert(KP <= kSimplexGroup, "one lane per component");
    constexpr int G = kSimplexGroup;
    constexpr int D = KP + 1;
    constexpr int WPB = 4, CPW = 32 / G;                                  // warps per block, chains per warp
    constexpr int ROW = KP + 2;                                           // z[KP], uniform, gamma
    constexpr int RREG = 2;                                               // model passes kept in registers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int mg = (a.m + G - 1) / G * G;
    real* const vt_s = reinterpret_cast<real*>(smem_raw);                 // [KP][mg], zero padded
    real* const draws_s = vt_s + KP * mg;                                 // [WPB][CPW][32][ROW]
    for (int i = threadIdx.x; i < KP * mg; i += blockDim.x) {
        const int k = i / mg, m = i % mg;
        vt_s[i] = (k < a.k && m < a.m) ? static_cast<real>(a.vt[k * a.m + m]) : real(0);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & (G - 1), grp = lane / G;
    const long long cid_raw = (static_cast<long long>(blockIdx.x) * WPB + warp) * CPW + grp;
    const bool chain_ok = cid_raw < a.n_chains;
    if (!__any_sync(0xffffffffu, chain_ok)) return;                       // whole warp beyond the last chain
    const long long cid = chain_ok ? cid_raw : a.n_chains - 1;            // idle groups shadow a real chain
    const uint32_t chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(cid));
    real* const mine = draws_s + (static_cast<size_t>(warp) * CPW + grp) * 32 * ROW;

    // lane g: constants and state of component g
    const bool comp = g < a.k;
    const real step = comp ? static_cast<real>(a.step[g]) : real(0);
    const real b_ols = comp ? static_cast<real>(a.b_ols[g]) : real(0);
    real grow[KP];                                                        // row g of G
#pragma unroll
    for (int c = 0; c < KP; ++c) grow[c] = (comp && c < a.k) ? static_cast<real>(a.gram[g * a.k + c]) : real(0);
    const real bias0 = static_cast<real>(1.0 / static_cast<double>(a.m));  // :78
    const real prior_scale = run_consts<real>(a).prior_scale;
    const real sig_ref = run_consts<real>(a).sigma_ref;
    const GammaConst<real> gc = gamma_const_of(run_consts<real>(a), a.gamma_boost);
    // models g, g + 8 in registers; more (m > 16) are read from shared memory
    const int n_pass = mg / G;
    real vcol[RREG][KP];
    bool vok[RREG];
#pragma unroll
    for (int r = 0; r < RREG; ++r) {
        vok[r] = g + G * r < a.m;
#pragma unroll
        for (int k = 0; k < KP; ++k) vcol[r][k] = g + G * r < mg ? vt_s[k * mg + g + G * r] : real(0);
    }

    real dc = -b_ols;                                                     // b starts at 0 (:82)
    real gdc = real(0);
#pragma unroll
    for (int c = 0; c < KP; ++c) gdc = M::fma(grow[c], __shfl_sync(0xffffffffu, dc, c, G), gdc);
    real rss = comp ? gdc * dc : real(0);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) rss += __shfl_xor_sync(0xffffffffu, rss, o, G);
    rss += run_consts<real>(a).rss_min;
    real s2 = run_consts<real>(a).sigma2_init;                           // :86

    // moment sums of lane g: dc_g, dc_g * dc_c (c = 0..KP-1), dc_g * es, and (used from lane 0) es, es^2
    real acc1 = real(0), accs = real(0), acce = real(0), accee = real(0);
    real acc2[KP];
#pragma unroll
    for (int c = 0; c < KP; ++c) acc2[c] = real(0);
    int n_acc = 0;
    real* const out = static_cast<real*>(a.samples);

    // one state update (:98-117 / :121-140)
    auto update = [&](const real* row, bool count_accept) {
        const real z = g < KP ? row[g] : real(0);
        const real u = row[KP];
        const real gm = row[KP + 1];
        const real delta = step * z;                                      // :98 / :121
        const real prop = (b_ols + dc) + delta;
        real pk[KP], dk[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            pk[k] = __shfl_sync(0xffffffffu, prop, k, G);
            dk[k] = __shfl_sync(0xffffffffu, delta, k, G);
        }
        // weights of the proposal (:99); only the minimum matters (:102)
        real wmin = real(1);
#pragma unroll
        for (int r = 0; r < RREG; ++r) {
            real w = bias0;
#pragma unroll
            for (int k = 0; k < KP; ++k) w = M::fma(pk[k], vcol[r][k], w);
            wmin = vok[r] ? fmin(wmin, w) : wmin;
        }
        for (int r = RREG; r < n_pass; ++r) {
            real w = bias0;
#pragma unroll
            for (int k = 0; k < KP; ++k) w = M::fma(pk[k], vt_s[k * mg + g + G * r], w);
            if (g + G * r < a.m) wmin = fmin(wmin, w);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o, G));
        real gdelta = real(0);
#pragma unroll
        for (int c = 0; c < KP; ++c) gdelta = M::fma(grow[c], dk[c], gdelta);
        real diff = delta * M::fma(real(2), gdc, gdelta);                 // RSS' - RSS = delta'G(2 dc + delta)
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o, G);
        const real alpha = M::exp(M::div(-diff, s2));                     // no factor 1/2 (:108 / :130)
        const bool accept = !(wmin < real(0)) && u < fmin(real(1), alpha);   // :102, :110
        dc = accept ? dc + delta : dc;
        gdc = accept ? gdc + gdelta : gdc;
        rss = accept ? rss + diff : rss;
        n_acc += (accept && count_accept) ? 1 : 0;
        s2 = M::div(real(0.5) * (prior_scale + rss), gm);                 // :116-117, no floor
    };

    const int burn = static_cast<int>(a.burn), total = static_cast<int>(a.burn + a.iterations);
    int next_store = a.samples ? burn : -1;
    int slot = 0;
    for (int base = 0; base < total; base += 32) {
        // ---- phase 1: the variates of iterations base .. base+31 of this group's chain
#pragma unroll 1
        for (int t = 0; t < 32 / G; ++t) {
            const int j = g + G * t;
            const uint32_t it32 = static_cast<uint32_t>(base + j);
            real* row = mine + j * ROW;
            Philox4 w[VariateLayout<KP>::kCalls];
            iteration_words<KP>(it32, chain, kTagSimplex, a.keys, w);
            {
                real z[KP];
                normals_of_words<real, KP>(w, z);
#pragma unroll
                for (int q = 0; q < KP; ++q) row[q] = z[q];
            }
            row[KP] = M::u01(philox4x32_10(it32, kBlockUniform, chain, kTagSimplex, a.keys).x);
            row[KP + 1] = gamma_unit_scale_w<real, KP>(gc, w, it32, chain, kTagSimplex, a.keys, a.key0, a.key1);
        }
        __syncwarp();
        // ---- phase 2: the state updates, in order
        const int n_here = min(32, total - base);
        if (base + n_here <= burn) {
            for (int j = 0; j < n_here; ++j) update(mine + j * ROW, false);   // burn-in: nothing recorded
        } else {
            for (int j = 0; j < n_here; ++j) {
                const int it = base + j;
                update(mine + j * ROW, it >= burn);
                if (it < burn) continue;
                const real sig = M::sqrt(s2);
                if (MODE != 0) {
                    const real es = sig - sig_ref;
                    acc1 += dc;
                    accs = M::fma(dc, es, accs);
                    acce += es;
                    accee = M::fma(es, es, accee);
                    if (MODE == 1) {
                        acc2[0] = M::fma(dc, dc, acc2[0]);
                    } else {
#pragma unroll
                        for (int c = 0; c < KP; ++c)
                            acc2[c] = M::fma(dc, __shfl_sync(0xffffffffu, dc, c, G), acc2[c]);
                    }
                    const int done = it - burn + 1;
                    if ((done % kFlushEvery) == 0 || it + 1 == total) {
                        auto row2 = [&](int r, int c) { return D + r * D - r * (r - 1) / 2 + (c - r); };
                        auto flush = [&](int row_out, real& v) {
                            if (chain_ok) {
                                double* p = a.chain_stats + static_cast<long long>(row_out) * a.n_chains + cid;
                                stat_add(p, static_cast<double>(v));
                            }
                            v = real(0);
                        };
                        if (g < KP) {
                            flush(g, acc1);
                            if (MODE == 1) {
                                flush(D + g, acc2[0]);
                            } else {
#pragma unroll
                                for (int c = 0; c < KP; ++c) {
                                    if (c >= g) flush(row2(g, c), acc2[c]);
                                    else acc2[c] = real(0);
                                }
                                flush(row2(g, KP), accs);
                            }
                        }
                        if (g == 0) {
                            flush(KP, acce);
                            flush(MODE == 1 ? D + KP : row2(KP, KP), accee);
                        }
                        acc1 = accs = acce = accee = real(0);
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc2[c] = real(0);
                    }
                }
                if (it == next_store) {
                    if (chain_ok) {
                        real* dst = out + static_cast<long long>(slot) * (a.k + 1) * a.n_chains + cid;
                        if (g < a.k) dst[static_cast<long long>(g) * a.n_chains] = b_ols + dc;
                        if (g == 0) dst[static_cast<long long>(a.k) * a.n_chains] = sig;
                    }
                    ++slot;
                    next_store += static_cast<int>(a.thin);
                }
            }
        }
        __syncwarp();
    }
    if (a.accepted && g == 0 && chain_ok) a.accepted[cid] = n_acc;
}

// --------------------------------------------------------------------------------------
// Simplex sampler, eight lanes per chain, at most 16 models (BASELINE configs[1]: 15 models, K = 3).
// Everything about a proposal that does not depend on the chain's state is linear in the proposal step
// delta = S * stepsize * z and is worked out in the parallel phase, 32 iterations at a time:
//     dw = delta' Vt_hat         (the change of the 16 model weights, :99)
//     G delta, delta' G delta    (the pieces of RSS' - RSS = delta' G (2 dc + delta), :104-108)
// next to the Metropolis uniform and the Gamma variate.  The serial phase then carries, replicated in the
// eight lanes of the group, dc = b - b_ols, G dc, RSS and sigma^2, and in lane g the weights of models g
// and g + 8: an iteration is a K-term dot product, two adds, a vote ("any weight negative", :102), one
// exp and a handful of selects -- no shuffles and a dependency chain of ~100 cycles instead of ~650.
// Same random stream and the same decisions as the other simplex kernels (the weights and G dc are carried
// as running sums, as RSS always was; they are rebuilt from dc at the start of every batch of 32).
constexpr int kSimplexModels16 = 16;

template <int KP>
struct SimplexRow16 {
    static constexpr int kDelta = 0, kGDelta = KP, kDw = 2 * KP, kScal = 2 * KP + kSimplexModels16;   // dgd, u, gamma
    static constexpr int kRow = 2 * KP + kSimplexModels16 + 4;
};

template <typename real, int KP, int MODE>
__global__ void __launch_bounds__(128) gibbs_simplex_group16_kernel(const SimplexArgs a) {
    using M = Math<real>;
    using R = SimplexRow16<KP>;
    constexpr int G = kSimplexGroup, MW = kSimplexModels16;
    constexpr int D = KP + 1;
    constexpr int WPB = 4, CPW = 32 / G;
    constexpr int ROW = R::kRow;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* const vt_s = reinterpret_cast<real*>(smem_raw);                 // [KP][16], zero padded
    real* const gram_s = vt_s + KP * MW;                                  // [KP][KP], zero padded
    real* const draws_s = gram_s + KP * KP;                               // [WPB][CPW][32][ROW]
    for (int i = threadIdx.x; i < KP * MW; i += blockDim.x) {
        const int k = i / MW, m = i % MW;
        vt_s[i] = (k < a.k && m < a.m) ? static_cast<real>(a.vt[k * a.m + m]) : real(0);
    }
    for (int i = threadIdx.x; i < KP * KP; i += blockDim.x) {
        const int r = i / KP, c = i % KP;
        gram_s[i] = (r < a.k && c < a.k) ? static_cast<real>(a.gram[r * a.k + c]) : real(0);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & (G - 1), grp = lane / G;
    const long long cid_raw = (static_cast<long long>(blockIdx.x) * WPB + warp) * CPW + grp;
    const bool chain_ok = cid_raw < a.n_chains;
    if (!__any_sync(0xffffffffu, chain_ok)) return;                       // whole warp beyond the last chain
    const long long cid = chain_ok ? cid_raw : a.n_chains - 1;            // idle groups shadow a real chain
    const uint32_t chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(cid));
    real* const mine = draws_s + (static_cast<size_t>(warp) * CPW + grp) * 32 * ROW;

    real step[KP], b_ols[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        step[k] = k < a.k ? static_cast<real>(a.step[k]) : real(0);
        b_ols[k] = k < a.k ? static_cast<real>(a.b_ols[k]) : real(0);
    }
    const bool comp = g < a.k;
    const real b_ols_own = comp ? static_cast<real>(a.b_ols[g]) : real(0);
    const real bias0 = static_cast<real>(1.0 / static_cast<double>(a.m));  // :78
    const real prior_scale = run_consts<real>(a).prior_scale;
    const real sig_ref = run_consts<real>(a).sigma_ref;
    const GammaConst<real> gc = gamma_const_of(run_consts<real>(a), a.gamma_boost);
    const bool vok0 = g < a.m, vok1 = g + G < a.m;
    const unsigned int grp_shift = static_cast<unsigned int>(G * grp);

    // state, replicated in the group: dc = b - b_ols (b starts at 0, :82), G dc, RSS, sigma^2
    real dc[KP], gdc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) dc[k] = -b_ols[k];
    real dc_own = -b_ols_own;
    real w0 = bias0, w1 = bias0, rss = real(0);
    auto rebuild = [&](bool with_rss) {
        // G dc and the weights from dc itself (bounds the drift of the running sums)
        real acc = real(0);
#pragma unroll
        for (int r = 0; r < KP; ++r) {
            real t = real(0);
#pragma unroll
            for (int c = 0; c < KP; ++c) t = M::fma(gram_s[r * KP + c], dc[c], t);
            gdc[r] = t;
            acc = M::fma(t, dc[r], acc);
        }
        if (with_rss) rss = acc + run_consts<real>(a).rss_min;
        real a0 = bias0, a1 = bias0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const real bk = b_ols[k] + dc[k];
            a0 = M::fma(bk, vt_s[k * MW + g], a0);
            a1 = M::fma(bk, vt_s[k * MW + g + G], a1);
        }
        w0 = a0;
        w1 = a1;
    };
    rebuild(true);
    real s2 = run_consts<real>(a).sigma2_init;                           // :86

    real acc1 = real(0), accs = real(0), acce = real(0), accee = real(0);
    real acc2[KP];
#pragma unroll
    for (int c = 0; c < KP; ++c) acc2[c] = real(0);
    int n_acc = 0;
    real* const out = static_cast<real*>(a.samples);

    // one state update (:98-117 / :121-140)
    auto update = [&](const real* row, bool count_accept) {
        real delta[KP], gdelta[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            delta[k] = row[R::kDelta + k];
            gdelta[k] = row[R::kGDelta + k];
        }
        const real delta_own = g < KP ? row[R::kDelta + (g < KP ? g : 0)] : real(0);
        const real dgd = row[R::kScal], u = row[R::kScal + 1], gm = row[R::kScal + 2];
        const real wn0 = w0 + row[R::kDw + g], wn1 = w1 + row[R::kDw + g + G];
        const bool neg = (vok0 & (wn0 < real(0))) | (vok1 & (wn1 < real(0)));     // :102
        const unsigned int any_neg = (__ballot_sync(0xffffffffu, neg) >> grp_shift) & 0xffu;
        real t0 = real(0), t1 = real(0);
#pragma unroll
        for (int k = 0; k < KP; k += 2) {
            t0 = M::fma(delta[k], gdc[k], t0);
            if (k + 1 < KP) t1 = M::fma(delta[k + 1], gdc[k + 1], t1);
        }
        const real diff = M::fma(real(2), t0 + t1, dgd);                  // RSS' - RSS = delta'G(2 dc + delta)
        const real alpha = M::exp(M::div(-diff, s2));                     // no factor 1/2 (:108 / :130)
        const bool accept = (any_neg == 0u) & (u < fmin(real(1), alpha)); // :102, :110
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            dc[k] = accept ? dc[k] + delta[k] : dc[k];
            gdc[k] = accept ? gdc[k] + gdelta[k] : gdc[k];
        }
        dc_own = accept ? dc_own + delta_own : dc_own;
        w0 = accept ? wn0 : w0;
        w1 = accept ? wn1 : w1;
        rss = accept ? rss + diff : rss;
        n_acc += (accept && count_accept) ? 1 : 0;
        s2 = M::div(real(0.5) * (prior_scale + rss), gm);                 // :116-117, no floor
    };

    const int burn = static_cast<int>(a.burn), total = static_cast<int>(a.burn + a.iterations);
    int next_store = a.samples ? burn : -1;
    int slot = 0;
    for (int base = 0; base < total; base += 32) {
        // ---- phase 1: everything state-independent of iterations base .. base+31 of this group's chain
#pragma unroll 1
        for (int t = 0; t < 32 / G; ++t) {
            const int j = g + G * t;
            const uint32_t it32 = static_cast<uint32_t>(base + j);
            real* row = mine + j * ROW;
            real delta[KP];
            Philox4 w[VariateLayout<KP>::kCalls];
            iteration_words<KP>(it32, chain, kTagSimplex, a.keys, w);
            {
                real z[KP];
                normals_of_words<real, KP>(w, z);
#pragma unroll
                for (int q = 0; q < KP; ++q) delta[q] = step[q] * z[q];                       // :98 / :121
            }
            real dgd = real(0);
#pragma unroll
            for (int r = 0; r < KP; ++r) {
                real gd = real(0);
#pragma unroll
                for (int c = 0; c < KP; ++c) gd = M::fma(gram_s[r * KP + c], delta[c], gd);
                row[R::kDelta + r] = delta[r];
                row[R::kGDelta + r] = gd;
                dgd = M::fma(delta[r], gd, dgd);
            }
#pragma unroll
            for (int m = 0; m < MW; ++m) {
                real dw = real(0);
#pragma unroll
                for (int k = 0; k < KP; ++k) dw = M::fma(delta[k], vt_s[k * MW + m], dw);
                row[R::kDw + m] = dw;
            }
            row[R::kScal] = dgd;
            row[R::kScal + 1] = M::u01(philox4x32_10(it32, kBlockUniform, chain, kTagSimplex, a.keys).x);
            row[R::kScal + 2] = gamma_unit_scale_w<real, KP>(gc, w, it32, chain, kTagSimplex, a.keys, a.key0, a.key1);
        }
        __syncwarp();
        if (base > 0) rebuild(false);
        // ---- phase 2: the state updates, in order
        const int n_here = min(32, total - base);
        if (base + n_here <= burn) {
            for (int j = 0; j < n_here; ++j) update(mine + j * ROW, false);   // burn-in: nothing recorded
        } else {
            for (int j = 0; j < n_here; ++j) {
                const int it = base + j;
                update(mine + j * ROW, it >= burn);
                if (it < burn) continue;
                const real sig = M::sqrt(s2);
                if (MODE != 0) {
                    const real es = sig - sig_ref;
                    acc1 += dc_own;
                    accs = M::fma(dc_own, es, accs);
                    acce += es;
                    accee = M::fma(es, es, accee);
                    if (MODE == 1) {
                        acc2[0] = M::fma(dc_own, dc_own, acc2[0]);
                    } else {
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc2[c] = M::fma(dc_own, dc[c], acc2[c]);
                    }
                    const int done = it - burn + 1;
                    if ((done % kFlushEvery) == 0 || it + 1 == total) {
                        auto row2 = [&](int r, int c) { return D + r * D - r * (r - 1) / 2 + (c - r); };
                        auto flush = [&](int row_out, real& v) {
                            if (chain_ok) {
                                double* p = a.chain_stats + static_cast<long long>(row_out) * a.n_chains + cid;
                                stat_add(p, static_cast<double>(v));
                            }
                            v = real(0);
                        };
                        if (g < KP) {
                            flush(g, acc1);
                            if (MODE == 1) {
                                flush(D + g, acc2[0]);
                            } else {
#pragma unroll
                                for (int c = 0; c < KP; ++c) {
                                    if (c >= g) flush(row2(g, c), acc2[c]);
                                    else acc2[c] = real(0);
                                }
                                flush(row2(g, KP), accs);
                            }
                        }
                        if (g == 0) {
                            flush(KP, acce);
                            flush(MODE == 1 ? D + KP : row2(KP, KP), accee);
                        }
                        acc1 = accs = acce = accee = real(0);
#pragma unroll
                        for (int c = 0; c < KP; ++c) acc2[c] = real(0);
                    }
                }
                if (it == next_store) {
                    if (chain_ok) {
                        real* dst = out + static_cast<long long>(slot) * (a.k + 1) * a.n_chains + cid;
                        if (g < a.k) dst[static_cast<long long>(g) * a.n_chains] = b_ols_own + dc_own;
                        if (g == 0) dst[static_cast<long long>(a.k) * a.n_chains] = sig;
                    }
                    ++slot;
                    next_store += static_cast<int>(a.thin);
                }
            }
        }
        __syncwarp();
    }
    if (a.accepted && chain_ok && g == 0) a.accepted[cid] = n_acc;
}

}  // namespace bmc
