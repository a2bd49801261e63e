// Literal (reference-shaped) Gibbs sampler: one chain per warp, the design matrix resident in
// shared memory, every iteration doing what pybmc/inference_utils.py:39-54 does --
//   A     = X'X / s2 + Lambda + 1e-6 I        (:41, factored instead of inverted)
//   mean  = A^-1 (Lambda b0 + X'y / s2)       (:42-44)
//   b     ~ N(mean, A^-1)                     (:45)   b = L^-T (L^-1 rhs + z),  A = L L'
//   RSS   = sum_i (y_i - x_i . b)^2           (:48-51) lanes stride the rows, warp-shuffle reduce
//   s2    = max(scale / Gamma(shape, 1), 1e-6) (:50-52)
// It costs O(nK) per iteration against O(K) for the sufficient-statistic kernel and exists as the
// parity anchor between that kernel and the reference's literal arithmetic.  X' ([k][n], the
// caller passes it transposed) and y are staged once per block with 1-D TMA bulk copies.
#pragma once
#include "tma.cuh"
#include "rng.cuh"

namespace bmc {

struct LiteralArgs {
    const void* xt;          // [k][n] real, X transposed
    const void* y;           // [n] real
    long long n;
    int k;
    const double* lam;       // [k*k] prior precision
    const double* lam_b0;    // [k]   Lambda b0
    double shape, prior_scale, sigma2_init;
    uint32_t key0, key1;
    unsigned long long chain0;
    long long n_chains, iterations;
    void* samples;           // [iterations][k+1][n_chains] real
};

template <typename real, int KP>
__global__ void __launch_bounds__(256) gibbs_literal_kernel(const LiteralArgs a) {
    using M = Math<real>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const long long npad = (a.n + 3) & ~3ll;
    real* const xs = reinterpret_cast<real*>(smem_raw);            // [k][npad]
    real* const ys = xs + static_cast<size_t>(a.k) * npad;         // [npad]
    real* const gram = ys + npad;                                  // [KP][KP]
    real* const xty = gram + KP * KP;                              // [KP]
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const real* xt = static_cast<const real*>(a.xt);
    const real* yv = static_cast<const real*>(a.y);

    const bool tma_ok = (a.n * sizeof(real)) % 16 == 0 && (reinterpret_cast<uintptr_t>(xt) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(yv) & 15) == 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tma_ok) {
        if (threadIdx.x == 0) {
            const uint32_t row_bytes = static_cast<uint32_t>(a.n * sizeof(real));
            mbar_expect_tx(&bar, row_bytes * (a.k + 1));
            for (int k = 0; k < a.k; ++k) tma_load_1d(xs + k * npad, xt + k * a.n, row_bytes, &bar);
            tma_load_1d(ys, yv, row_bytes, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (long long i = threadIdx.x; i < a.n * a.k; i += blockDim.x) xs[(i / a.n) * npad + i % a.n] = xt[i];
        for (long long i = threadIdx.x; i < a.n; i += blockDim.x) ys[i] = yv[i];
    }
    __syncthreads();
    // X'X and X'y once per block (:25, :43), one (r, c) entry per warp at a time
    for (int e = warp; e < KP * (KP + 1); e += nwarp) {
        const int r = e / (KP + 1), c = e % (KP + 1);
        real s = real(0);
        if (r < a.k && (c < a.k || c == KP)) {
            const real* pr = xs + r * npad;
            const real* pc = c == KP ? ys : xs + c * npad;
            for (long long i = lane; i < a.n; i += 32) s = M::fma(pr[i], pc[i], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if (lane == 0) {
            if (c == KP) xty[r] = s;
            else gram[r * KP + c] = s;
        }
    }
    __syncthreads();

    const long long wid = static_cast<long long>(blockIdx.x) * nwarp + warp;
    if (wid >= a.n_chains) return;
    const uint32_t chain = static_cast<uint32_t>(a.chain0 + static_cast<unsigned long long>(wid));
    const GammaConst<real> gc = make_gamma_const<real>(a.shape);
    const PhiloxKeys ks = philox_keys(a.key0, a.key1);
    const real prior_scale = static_cast<real>(a.prior_scale);
    real s2 = static_cast<real>(a.sigma2_init);
    real* const out = static_cast<real*>(a.samples);

    for (long long it = 0; it < a.iterations; ++it) {
        const uint32_t it32 = static_cast<uint32_t>(it);
        const real inv_s2 = M::rcp(s2);
        // every lane factors the same K-by-K system in registers: no communication needed
        real L[KP][KP];
        real rhs[KP];
#pragma unroll
        for (int r = 0; r < KP; ++r) {
            rhs[r] = r < a.k ? static_cast<real>(a.lam_b0[r]) + xty[r] * inv_s2 : real(0);
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                real v = real(0);
                if (r < a.k && c < a.k) v = M::fma(gram[r * KP + c], inv_s2, static_cast<real>(a.lam[r * a.k + c]));
                if (r == c) v += r < a.k ? real(1e-6) : real(1);
                L[r][c] = v;
            }
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            real dj = L[j][j];
#pragma unroll
            for (int p = 0; p < j; ++p) dj = M::fma(-L[j][p], L[j][p], dj);
            const real inv = M::rsqrt(dj);
            L[j][j] = inv;                       // store 1 / L_jj
#pragma unroll
            for (int i = j + 1; i < KP; ++i) {
                real v = L[i][j];
#pragma unroll
                for (int p = 0; p < j; ++p) v = M::fma(-L[i][p], L[j][p], v);
                L[i][j] = v * inv;
            }
        }
        real z[KP];
        Philox4 w[VariateLayout<KP>::kCalls];
        iteration_words<KP>(it32, chain, kTagGibbs, ks, w);
        normals_of_words<real, KP>(w, z);
        // forward: t = L^-1 rhs + z ; backward: b = L^-T t
        real b[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            real v = rhs[i];
#pragma unroll
            for (int p = 0; p < i; ++p) v = M::fma(-L[i][p], b[p], v);
            b[i] = v * L[i][i];
        }
#pragma unroll
        for (int i = 0; i < KP; ++i) b[i] += z[i];
#pragma unroll
        for (int i = KP - 1; i >= 0; --i) {
            real v = b[i];
#pragma unroll
            for (int p = i + 1; p < KP; ++p) v = M::fma(-L[p][i], b[p], v);
            b[i] = v * L[i][i];
        }
        // residual sum of squares over the rows, lanes striding i
        real rss = real(0);
        for (long long i = lane; i < a.n; i += 32) {
            real f = ys[i];
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < a.k) f = M::fma(-xs[k * npad + i], b[k], f);
            rss = M::fma(f, f, rss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rss += __shfl_xor_sync(0xffffffffu, rss, o);
        const real scale = real(0.5) * (prior_scale + rss);
        const real gm = gamma_unit_scale_w<real, KP>(gc, w, it32, chain, kTagGibbs, ks, a.key0, a.key1);
        s2 = M::div(scale, gm);
        s2 = s2 > real(1e-6) ? s2 : real(1e-6);
        if (out && lane <= a.k) {
            real v = M::sqrt(s2);                                    // last column is sigma (:54)
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (lane == k && k < a.k) v = b[k];
            out[(it * (a.k + 1) + lane) * a.n_chains + wid] = v;
        }
    }
}

}  // namespace bmc
