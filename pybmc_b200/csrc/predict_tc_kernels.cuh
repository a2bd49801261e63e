// Tensor-core variant of the fused prediction pass (fp32).
//
// The contraction  x[s][n] = u[n] . beta[s]  of pybmc/sampling_utils.py:64-77 is a GEMM whose output
// is consumed at once (noise added, compared, counted) and never stored.  With K = 64 the FFMA form
// spends 64 of its ~150 instructions per (draw, nucleus) on the dot product; here that part moves to
// the 5th-generation tensor cores:
//
//   D[128 nuclei][128 draws] (fp32, TMEM)  =  A[128][K] . B[128][K]'
//     A = u rows of this block's nuclei, B = one tile of posterior draws, both K-major in shared
//     memory, no swizzle (8-row x 16-byte core matrices; see umma_desc below)
//   tcgen05.mma kind::tf32, issued by one thread, three products per K-step (split TF32:
//     x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  A.B ~ Alo.Bhi + Ahi.Blo + Ahi.Bhi,
//     the dropped lo.lo term is 2^-22 relative) so the result keeps fp32-level accuracy
//   accumulator row i lives in TMEM lane i, so `tcgen05.ld.32x32b` hands every thread the draws of
//   ITS nucleus: exactly the lane <-> nucleus layout the per-lane consumer (consume4) wants.
//
// Block = 16 consumer warps + 1 producer warp: consumer warp w reads TMEM lanes 32 (w % 4) .. +31 (the quarter a warp may address) and
// columns 32 (w / 4) .. +31 of each 128-draw tile, i.e. four warps share a nucleus and act as four
// sample splits ("slots").  The draws arrive as a pre-split image (theta_image_kernel) moved by one
// 64 KB TMA bulk copy per tile, two stages; the accumulator is double-buffered in TMEM (2 x 128
// columns) so the MMAs of tile t+1 run under the consumer of tile t.
#pragma once
#include "predict_kernels.cuh"

namespace bmc {

constexpr int kTcRows = 128;      // nuclei per block  (UMMA M)
constexpr int kTcTile = 128;      // draws per tile    (UMMA N)
constexpr int kTcThreads = 512;
constexpr int kTcSlotsPerBlock = 4;
constexpr int kTcTmemCols = 256;  // two accumulator buffers

template <int KP>
struct TcImage {
    static constexpr int kOperandBytes = KP * 128 * 4;          // 128 rows of one operand, hi or lo
    static constexpr int kTileBytes = 2 * kOperandBytes;        // hi then lo: what one TMA copy moves
    static constexpr int kStride = kTileBytes + kTcTile * 4;    // + sigma[128]
    static constexpr int kSmemBytes = 2 * kOperandBytes + 2 * kTileBytes;   // A (hi, lo) + two B stages
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo, both representable in TF32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(to_tf32(x));
    lo = __uint_as_float(to_tf32(x - hi));
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// completion of every MMA issued so far by this thread arrives on the mbarrier
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// Shared-memory operand descriptor: K-major, SWIZZLE_NONE.  In 16-byte units the operand is
// ((8 rows, m groups), 2 chunks) : ((1, SBO), LBO): a core matrix is 8 rows x 16 bytes stored
// contiguously (128 B); SBO is the distance between 8-row groups, LBO the distance between the two
// 16-byte K chunks one MMA (K = 8 tf32 = 32 bytes) reads.  Our image is [K/4 chunks][128 rows][16 B]:
// SBO = 128 B, LBO = 128 rows * 16 B = 2048 B.  Bits: [0,14) address >> 4, [16,30) LBO >> 4,
// [32,46) SBO >> 4, [46,48) version = 1 (sm_100), [61,64) layout = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu) | (static_cast<uint64_t>(2048 >> 4) << 16) |
           (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::tf32: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both
// K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTcTile >> 3) << 17) | ((kTcRows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kTcIdesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_load32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_load8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- the draws as tensor-core operand images ----------------------------------------------------
// img[tile] = [hi: K/4 chunks x 128 rows x 16 B][lo: same][sigma: 128 floats]; rows past n_draws are 0.
template <int KP>
__global__ void __launch_bounds__(256) theta_image_kernel(const float* __restrict__ theta, long long n_draws,
                                                          int ldt, int sigma_col, unsigned char* __restrict__ img) {
    // theta rows: beta in columns 0 .. sigma_col-1 (zero padded), sigma in column sigma_col, stride ldt
    const long long t = blockIdx.x;
    unsigned char* out = img + t * TcImage<KP>::kStride;
    for (int i = threadIdx.x; i < (KP / 4) * 128; i += blockDim.x) {
        const int r = i & 127, c = i >> 7;
        const long long s = t * kTcTile + r;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (s < n_draws) {
            const float* row = theta + s * ldt;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * c + e < sigma_col) v[e] = row[4 * c + e];
        }
        float4 hi, lo;
        split_tf32(v[0], hi.x, lo.x);
        split_tf32(v[1], hi.y, lo.y);
        split_tf32(v[2], hi.z, lo.z);
        split_tf32(v[3], hi.w, lo.w);
        *reinterpret_cast<float4*>(out + (c * 128 + r) * 16) = hi;
        *reinterpret_cast<float4*>(out + TcImage<KP>::kOperandBytes + (c * 128 + r) * 16) = lo;
    }
    for (int r = threadIdx.x; r < kTcTile; r += blockDim.x) {
        const long long s = t * kTcTile + r;
        reinterpret_cast<float*>(out + TcImage<KP>::kTileBytes)[r] = s < n_draws ? theta[s * ldt + sigma_col] : 1.0f;
    }
}

// ---- the pass -------------------------------------------------------------------------------------
// grid = (ceil(n_active / 128), sample splits);  a.s_splits = 4 * gridDim.y slots.
// Warps 0..15 consume (lane <-> nucleus); warp 16 is the producer: one elected thread issues the TMA
// copies and the MMAs.  No block-wide barrier in the loop: mbarriers carry "tile landed" (full),
// "accumulator complete" (done) and "accumulator copied to registers by all 16 warps" (empty).
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int KP, int NQ>
__global__ void __launch_bounds__(kTcThreads + 32, 1) predict_pass_tc_kernel(const PredictArgs a,
                                                                             const unsigned char* __restrict__ img) {
    using IM = TcImage<KP>;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* const a_hi = tc_smem_raw;
    unsigned char* const a_lo = tc_smem_raw + IM::kOperandBytes;
    unsigned char* const stage0 = tc_smem_raw + 2 * IM::kOperandBytes;
    __shared__ __align__(8) uint64_t full_bar[2], done_bar[2], empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kConsumerWarps = kTcThreads / 32;
    const bool producer = warp == kConsumerWarps;

    const int gy = gridDim.y;
    const long long per = ((a.n_draws + gy - 1) / gy + kTcTile - 1) / kTcTile * kTcTile;
    const long long s_begin = static_cast<long long>(blockIdx.y) * per;
    const long long s_end = min(a.n_draws, s_begin + per);
    const int n_tiles = s_end > s_begin ? static_cast<int>((s_end - s_begin + kTcTile - 1) / kTcTile) : 0;
    const long long tile0 = s_begin / kTcTile;

    if (warp == 0) tmem_alloc(&tmem_slot, kTcTmemCols);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&done_bar[i], 1);
            mbar_init(&empty_bar[i], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // A operand: the u rows of this block's nuclei, split and laid out [chunk][row][16 B]
    for (int i = tid; i < (KP / 4) * kTcRows; i += kTcThreads + 32) {
        const int r = i & 127, c = i >> 7;
        const int ps = blockIdx.x * kTcRows + r;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (ps < a.n_active) {
            const int nn = a.point_list ? a.point_list[ps] : ps;
            const float* ur = static_cast<const float*>(a.u) + static_cast<long long>(nn) * a.k;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * c + e < a.k) v[e] = ur[4 * c + e];
        }
        float4 hi, lo;
        split_tf32(v[0], hi.x, lo.x);
        split_tf32(v[1], hi.y, lo.y);
        split_tf32(v[2], hi.z, lo.z);
        split_tf32(v[3], hi.w, lo.w);
        *reinterpret_cast<float4*>(a_hi + (c * 128 + r) * 16) = hi;
        *reinterpret_cast<float4*>(a_lo + (c * 128 + r) * 16) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (producer) {
        if (lane == 0) {
            auto issue_tma = [&](int t) {
                const int st = t & 1;
                mbar_expect_tx(&full_bar[st], IM::kTileBytes);
                tma_load_1d(stage0 + st * IM::kTileBytes, img + (tile0 + t) * IM::kStride, IM::kTileBytes,
                            &full_bar[st]);
            };
            if (n_tiles > 0) issue_tma(0);
            if (n_tiles > 1) issue_tma(1);
            const uint32_t ahi = smem_u32(a_hi), alo = smem_u32(a_lo);
            for (int t = 0; t < n_tiles; ++t) {
                const int st = t & 1;
                const uint32_t ph = static_cast<uint32_t>(t >> 1) & 1u;
                mbar_wait_backoff(&full_bar[st], ph);               // operands of tile t are in shared memory
                if (t >= 2) mbar_wait_backoff(&empty_bar[st], ph ^ 1u, 512);   // tile t-2 left this accumulator buffer
                tc_fence_after();
                const uint32_t d = tmem_base + static_cast<uint32_t>(st * kTcTile);
                const uint32_t bhi = smem_u32(stage0 + st * IM::kTileBytes), blo = bhi + IM::kOperandBytes;
                uint32_t acc = 0u;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {              // small terms first
                    const uint32_t ab = pass == 0 ? alo : ahi, bb = pass == 1 ? blo : bhi;
#pragma unroll
                    for (int j = 0; j < KP / 8; ++j) {
                        umma_tf32(d, umma_desc(ab + j * 4096), umma_desc(bb + j * 4096), acc);
                        acc = 1u;
                    }
                }
                umma_commit(&done_bar[st]);
                if (t + 2 < n_tiles) {
                    mbar_wait_backoff(&done_bar[st], ph, 256);      // the MMAs have read the stage: refill it
                    issue_tma(t + 2);
                }
            }
        }
    } else {
        const int quarter = warp & 3;             // TMEM lanes 32 q .. 32 q + 31 are this warp's
        const int colgrp = warp >> 2;             // columns 32 c .. 32 c + 31 of every tile
        const int row = 32 * quarter + lane;
        const int pslot = blockIdx.x * kTcRows + row;
        const bool live = pslot < a.n_active;
        const int n = live ? (a.point_list ? a.point_list[pslot] : pslot) : 0;
        LaneCtx c;
        c.n = n;
        c.idx0 = n * a.nq;
        c.slot = blockIdx.y * kTcSlotsPerBlock + colgrp;
        c.live = live;
        c.seg = static_cast<unsigned int>(a.seg_len);
        c.mu_d = (live && a.mu) ? a.mu[n] : 0.0;
        const float tc = (live && a.truth) ? static_cast<float>(a.truth[n] - c.mu_d) : 0.f;
        const float ctr = (live && a.center) ? static_cast<const float*>(a.center)[n] : 0.f;
        LaneAcc<float, NQ> acc;
        lane_init<float, NQ>(a, c, acc);
        const uint32_t nglob = static_cast<uint32_t>(a.point0 + static_cast<unsigned long long>(n));
        const bool fast_first = a.first && !a.draws_out;

        for (int t = 0; t < n_tiles; ++t) {
            const int st = t & 1;
            mbar_wait(&done_bar[st], static_cast<uint32_t>(t >> 1) & 1u);   // accumulator of tile t complete
            tc_fence_after();
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(32 * quarter) << 16) +
                                  static_cast<uint32_t>(st * kTcTile + 32 * colgrp);
            const long long s0 = s_begin + static_cast<long long>(t) * kTcTile + 32 * colgrp;
            const int rem = static_cast<int>(min(s_end - s0, static_cast<long long>(32)));   // draws of this warp's columns
            const float* sig =
                reinterpret_cast<const float*>(img + (tile0 + t) * IM::kStride + IM::kTileBytes) + 32 * colgrp;
            // the next tile's sigmas (one 128-byte line per column group, read straight from the image in global
            // memory) on their way to L1 while this tile is consumed: their L2 latency used to be paid by the
            // first eight draws of every tile (long_scoreboard, 2.4 % of the samples)
            if (t + 1 < n_tiles && lane == 0)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(sig + IM::kStride / sizeof(float)));
            // the common tile: a first pass that does not materialise the draws, 32 real draws for this warp, Philox
            // noise -- one flag instead of five uniform tests per eight draws
            const bool turbo = fast_first && a.noise_mode == 1 && rem >= 32;
            // eight draws at a time, rolled: the loop body stays small enough for the instruction cache
            // (the 32-draw unrolled form spent 13 % of its stall samples waiting for instructions)
#pragma unroll 1
            for (int g = 0; g < 32; g += 8) {
                float xs[8];
                tmem_load8(tcol + static_cast<uint32_t>(g), xs);
                if (turbo) {
                    // Two Philox calls side by side (independent chains for the scheduler to interleave),
                    // Box-Muller and x = xs + sigma z on packed fp32 pairs.
                    const uint32_t blk = static_cast<uint32_t>((s0 + g) >> 2);
                    const Philox4 ra = philox4x32_10(blk, nglob, 0u, kTagNoise, a.keys);
                    const Philox4 rb = philox4x32_10(blk + 1u, nglob, 0u, kTagNoise, a.keys);
                    const float4 s_a = *reinterpret_cast<const float4*>(sig + g);
                    const float4 s_b = *reinterpret_cast<const float4*>(sig + g + 4);
                    float x[8];
                    unpack2(fma2(pack2(s_a.x, s_a.y), Math<float>::box_muller2(ra.x, ra.y), pack2(xs[0], xs[1])), x[0], x[1]);
                    unpack2(fma2(pack2(s_a.z, s_a.w), Math<float>::box_muller2(ra.z, ra.w), pack2(xs[2], xs[3])), x[2], x[3]);
                    unpack2(fma2(pack2(s_b.x, s_b.y), Math<float>::box_muller2(rb.x, rb.y), pack2(xs[4], xs[5])), x[4], x[5]);
                    unpack2(fma2(pack2(s_b.z, s_b.w), Math<float>::box_muller2(rb.z, rb.w), pack2(xs[6], xs[7])), x[6], x[7]);
                    const float xa[4] = {x[0], x[1], x[2], x[3]}, xb[4] = {x[4], x[5], x[6], x[7]};
                    consume4<float, NQ, true, true>(a, c, acc, xa, s0 + g, tc, ctr);
                    consume4<float, NQ, true, true>(a, c, acc, xb, s0 + g + 4, tc, ctr);
                    continue;
                }
                if (g >= rem) continue;                                     // warp-uniform
                // tail of the draw range, or a pass without noise
                float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                float sg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (a.noise_mode == 1) {
                    const float4 s_a = *reinterpret_cast<const float4*>(sig + g);
                    const float4 s_b = *reinterpret_cast<const float4*>(sig + g + 4);
                    sg[0] = s_a.x; sg[1] = s_a.y; sg[2] = s_a.z; sg[3] = s_a.w;
                    sg[4] = s_b.x; sg[5] = s_b.y; sg[6] = s_b.z; sg[7] = s_b.w;
                    float za[4], zb[4];
                    const uint32_t blk = static_cast<uint32_t>((s0 + g) >> 2);
                    normals4_k<float>(blk, nglob, 0u, kTagNoise, a.keys, za);
                    normals4_k<float>(blk + 1u, nglob, 0u, kTagNoise, a.keys, zb);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        z[r] = za[r];
                        z[4 + r] = zb[r];
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (g + 4 * h < rem) {
                        float x[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            x[r] = fmaf(sg[4 * h + r], z[4 * h + r], xs[4 * h + r]);
                            if (g + 4 * h + r >= rem) x[r] = FLT_MAX;
                        }
                        consume4<float, NQ>(a, c, acc, x, s0 + g + 4 * h, tc, ctr);
                    }
                }
            }
            // this warp has read its part of the accumulator buffer (the producer is two tiles ahead: releasing it
            // here instead of right after the last tcgen05.ld costs nothing and takes a test out of the loop)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
        lane_flush<float, NQ>(a, c, acc);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, kTcTmemCols);
}

}  // namespace bmc
