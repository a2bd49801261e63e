// Orthogonalisation and "residual setup" kernels (fp64 throughout).
//
//   center_rows_kernel   : pybmc/bmc.py:106-116   mu = mean over models, y = truth - mu, Xc = P - mu
//   gram_partial_kernel  : the Gram matrix of [Xc | y]; replaces the n-by-n-U dgesdd of bmc.py:119
//                          (only the M-by-M eigenproblem is left to cuSOLVER) and the X'X, X'y of
//                          pybmc/inference_utils.py:25,28,43
//   project_rows_kernel  : U_hat = Xc Vt_hat'  (bmc.py:122 -> inference_utils.py:164) and
//                          u = P Vt_hat' for new points (pybmc/sampling_utils.py:64-72 contracted in
//                          K-space)
//   rss_partial_kernel   : RSS_min = |y - X b|^2  (inference_utils.py:29-31, 48-51)
//
// All reductions over rows are two-stage (per-block partials, then a fixed-order sum) so results
// are reproducible run to run.
#pragma once
#include <cuda_runtime.h>

namespace bmc {

constexpr int kGramTile = 32;   // columns per Gram tile
constexpr int kGramRows = 32;   // rows staged per step

__global__ void center_rows_kernel(const double* __restrict__ p, long long n, int m, long long ld,
                                   const double* __restrict__ truth, double* __restrict__ mu,
                                   double* __restrict__ y, double* __restrict__ xc, long long ldx) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long r = warp; r < n; r += nwarp) {
        const double* row = p + r * ld;
        double s = 0.0;
        for (int c = lane; c < m; c += 32) s += row[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double mean = s / static_cast<double>(m);
        if (lane == 0) {
            mu[r] = mean;
            if (y) y[r] = truth[r] - mean;
        }
        if (xc)
            for (int c = lane; c < m; c += 32) xc[r * ldx + c] = row[c] - mean;
    }
}

// partial[chunk][m1*m1] += A'A over this block's rows, for the tile pair (ti, tj), ti <= tj.
// A = [a - mu | b]; m1 = m + (b != nullptr).
__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ a, long long n, int m,
                                                           long long ld, const double* __restrict__ mu,
                                                           const double* __restrict__ b, int m1,
                                                           long long rows_per_chunk,
                                                           double* __restrict__ partial) {
    __shared__ double sa[kGramRows][kGramTile + 1];
    __shared__ double sb[kGramRows][kGramTile + 1];
    // decode the upper-triangular tile pair from blockIdx.y
    const int nt = (m1 + kGramTile - 1) / kGramTile;
    int ti = 0, rem = blockIdx.y;
    while (rem >= nt - ti) {
        rem -= nt - ti;
        ++ti;
    }
    const int tj = ti + rem;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 2 x 2 outputs each
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_chunk;
    const long long r1 = min(n, r0 + rows_per_chunk);
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (long long rb = r0; rb < r1; rb += kGramRows) {
        for (int i = threadIdx.x; i < kGramRows * kGramTile; i += 256) {
            const int rr = i / kGramTile, cc = i % kGramTile;
            const long long r = rb + rr;
            double va = 0.0, vb = 0.0;
            if (r < r1) {
                const double shift = mu ? mu[r] : 0.0;
                const int ca = ti * kGramTile + cc, cb = tj * kGramTile + cc;
                if (ca < m) va = a[r * ld + ca] - shift;
                else if (ca == m && b) va = b[r];
                if (cb < m) vb = a[r * ld + cb] - shift;
                else if (cb == m && b) vb = b[r];
            }
            sa[rr][cc] = va;
            sb[rr][cc] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < kGramRows; ++rr) {
            const double a0 = sa[rr][ty], a1 = sa[rr][ty + 16];
            const double b0 = sb[rr][tx], b1 = sb[rr][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
    double* out = partial + static_cast<long long>(blockIdx.x) * m1 * m1;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = ti * kGramTile + ty + 16 * i, gj = tj * kGramTile + tx + 16 * j;
            if (gi < m1 && gj < m1) {
                out[gi * m1 + gj] = acc[i][j];
                if (ti != tj) out[gj * m1 + gi] = acc[i][j];
            }
        }
}

// out[i] = sum over chunks of partial[chunk][i], in chunk order.
__global__ void sum_partials_kernel(const double* __restrict__ partial, long long n_chunks, long long width,
                                    double* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= width) return;
    double s = 0.0;
    for (long long c = 0; c < n_chunks; ++c) s += partial[c * width + i];
    out[i] = s;
}

// out[n][k] = (A - mu) Vt'   with Vt [k][m] row-major; 64 rows x 32 outputs per block.
__global__ void __launch_bounds__(256) project_rows_kernel(const double* __restrict__ a, long long n, int m,
                                                           long long ld, const double* __restrict__ mu,
                                                           const double* __restrict__ vt, int k,
                                                           double* __restrict__ out, long long ldo) {
    __shared__ double sa[64][33];
    __shared__ double sv[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 warps; warp ty owns rows ty*8 .. +7
    const long long r0 = static_cast<long long>(blockIdx.x) * 64;
    const int k0 = blockIdx.y * 32;
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    for (int m0 = 0; m0 < m; m0 += 32) {
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {
            const int rr = i >> 5, cc = i & 31;
            const long long r = r0 + rr;
            double v = 0.0;
            if (r < n && m0 + cc < m) v = a[r * ld + m0 + cc] - (mu ? mu[r] : 0.0);
            sa[rr][cc] = v;
        }
        for (int i = threadIdx.x; i < 32 * 32; i += 256) {
            const int kk = i >> 5, cc = i & 31;
            sv[kk][cc] = (k0 + kk < k && m0 + cc < m) ? vt[static_cast<long long>(k0 + kk) * m + m0 + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int cc = 0; cc < 32; ++cc) {
            const double v = sv[tx][cc];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(sa[ty * 8 + i][cc], v, acc[i]);
        }
        __syncthreads();
    }
    if (k0 + tx < k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = r0 + ty * 8 + i;
            if (r < n) out[r * ldo + k0 + tx] = acc[i];
        }
}

// partial[block] = sum over this block's rows of (y - X b)^2
__global__ void __launch_bounds__(256) rss_partial_kernel(const double* __restrict__ x, long long n, int k,
                                                          long long ld, const double* __restrict__ y,
                                                          const double* __restrict__ b,
                                                          double* __restrict__ partial) {
    __shared__ double red[8];
    double s = 0.0;
    for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < n;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        double f = y[r];
        for (int c = 0; c < k; ++c) f = fma(-x[r * ld + c], b[c], f);
        s = fma(f, f, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// class[i] = 0 if some reference point lies within d1 of point i, 1 if within d2 (but none within d1),
// 2 otherwise: Dataset.separate_points_distance_allSets, pybmc/data.py:194-245 (the O(N R) Python double
// loop with one np.linalg.norm per pair).  Distances are sqrt(sum of squares) in fp64, compared with <=
// exactly as the reference does.  Reference points are staged through shared memory in tiles.
__global__ void __launch_bounds__(256) nearest_class_kernel(const double* __restrict__ pts, long long n,
                                                            const double* __restrict__ refs, long long r, int dim,
                                                            double d1, double d2, int* __restrict__ cls) {
    extern __shared__ double tile[];                         // [tile_refs][dim]
    const int tile_refs = 1024 / dim > 0 ? 1024 / dim : 1;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool in1 = false, in2 = false;
    for (long long r0 = 0; r0 < r; r0 += tile_refs) {
        const int cnt = static_cast<int>(min(static_cast<long long>(tile_refs), r - r0));
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * dim; t += blockDim.x) tile[t] = refs[r0 * dim + t];
        __syncthreads();
        if (i < n && !in1) {
            for (int j = 0; j < cnt; ++j) {
                double s = 0.0;
                for (int c = 0; c < dim; ++c) {
                    const double diff = pts[i * dim + c] - tile[j * dim + c];
                    s += diff * diff;
                }
                const double dist = sqrt(s);
                in1 = in1 || dist <= d1;
                in2 = in2 || dist <= d2;
            }
        }
    }
    if (i < n) cls[i] = in1 ? 0 : (in2 ? 1 : 2);
}

}  // namespace bmc
