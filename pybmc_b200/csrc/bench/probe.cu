// Peak probes for the non-HBM rooflines (benchmark tooling: built into libbmc_probe.so, NOT part of the
// product ABI in include/bmc_b200.h; bench.py and profiles/pipe_probes.py are the only callers).  SURVEY.md section 8(d): the sampler and predictive kernels
// are bound by the FP32 FMA pipe, the MUFU (XU) pipe and integer issue, and MEASURED_PEAKS.json only
// holds the HBM and bf16 tensor peaks -- so bench.py measures the pipe peaks it divides by with these
// kernels, on the same box and at the clocks of the same run.  Each probe is a grid of warps running
// `iters` iterations of independent register-only chains; the result is written so nothing is elided.
#include <cuda_runtime.h>
#include "probe.h"

namespace {

constexpr int kChains = 8;

__global__ void probe_ffma(long long iters, float* sink) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    const float m = 0.999999f, c = 1e-7f;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

// fp64 FMA pipe (the bound of the fp64 samplers)
__global__ void probe_dfma(long long iters, float* sink) {
    double a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double m = 0.999999, c = 1e-7;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 123.456) sink[0] = static_cast<float>(s);
}

// DFMA with three DISTINCT register operands per instruction (what real code issues: probe_dfma's multiplier and
// addend are the same two registers in every instruction), and the same interleaved with LOP3
template <int WITH_LOP>
__global__ void probe_dfma3(long long iters, float* sink) {
    double a[kChains], b[kChains], c[kChains];
    unsigned x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
        b[i] = 0.999999 - 1e-9 * (threadIdx.x + 3 * i);
        c[i] = 1e-7 * (1 + threadIdx.x + i);
        x[i] = threadIdx.x * 2654435761u + i;
    }
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            a[i] = fma(a[i], b[i], c[(i + 1) % kChains]);
            if (WITH_LOP) x[i] = (x[i] ^ x[(i + 3) % kChains]) & (x[(i + 5) % kChains] | 0x55555555u);
        }
    }
    double s = 0.0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        s += a[i];
        t ^= x[i];
    }
    if (s == 123.456 || t == 0x12345u) sink[0] = static_cast<float>(s);
}

__global__ void probe_mufu(long long iters, float* sink) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = 0.5f + 1e-3f * (threadIdx.x + i);
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
        for (int i = 0; i < kChains; ++i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

// the sampler's integer mix: 32x32->64 multiply (fma pipe) + three-input logic (alu pipe), as in Philox
__global__ void probe_philox_mix(long long iters, float* sink) {
    unsigned int a[kChains], b[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 2654435761u + i;
        b[i] = a[i] ^ 0x9E3779B9u;
    }
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            const unsigned long long p = static_cast<unsigned long long>(0xD2511F53u) * a[i];
            a[i] = static_cast<unsigned int>(p >> 32) ^ b[i] ^ 0xBB67AE85u;
            b[i] = static_cast<unsigned int>(p);
        }
    }
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] ^ b[i];
    if (s == 0x12345678u) sink[0] = static_cast<float>(s);
}

// issue peak: independent FFMA (fma pipe) and LOP3 (alu pipe) streams side by side
__global__ void probe_issue(long long iters, float* sink) {
    float a[kChains];
    unsigned int b[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
        b[i] = threadIdx.x * 2654435761u + i;
    }
    const float m = 0.999999f, c = 1e-7f;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            a[i] = fmaf(a[i], m, c);
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[(i + 1) % kChains]), "r"(0x9E3779B9u));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + static_cast<float>(b[i] & 1u);
    if (s == 123.456f) sink[0] = s;
}

// single-instruction streams, to price the pieces of Philox
template <int KIND>
__global__ void probe_single(long long iters, float* sink) {
    unsigned int a[kChains], b[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 2654435761u + i;
        b[i] = a[i] ^ 0x9E3779B9u;
    }
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (KIND == 0) {            // IMAD.WIDE.U32
                const unsigned long long p = static_cast<unsigned long long>(0xD2511F53u) * a[i];
                a[i] = static_cast<unsigned int>(p >> 32) + static_cast<unsigned int>(p);
            } else if (KIND == 1) {     // IMAD.HI.U32
                a[i] = __umulhi(a[i], 0xD2511F53u);
            } else if (KIND == 2) {     // IMAD (32-bit multiply-add)
                a[i] = a[i] * 0xD2511F53u + b[i];
            } else {                    // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(0x9E3779B9u));
            }
        }
    }
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] ^ b[i];
    if (s == 0x12345678u) sink[0] = static_cast<float>(s);
}

// Mixed streams: per chain and iteration W x IMAD.WIDE, F x FFMA, F2 x FFMA2 (packed fp32 pair), M x MUFU,
// L x LOP3, all independent across chains.  Shows which instruction classes share a pipe on this part
// (e.g. whether FFMA keeps flowing while IMAD.WIDE holds the heavy half of the FMA pipe).
template <int W, int F, int F2, int M, int L>
__global__ void probe_mix(long long iters, float* sink) {
    unsigned int a[kChains], b[kChains];
    float f[kChains], g[kChains];
    unsigned long long h[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 2654435761u + i;
        b[i] = a[i] ^ 0x9E3779B9u;
        f[i] = 1.0f + 1e-3f * (threadIdx.x + i);
        g[i] = 0.5f + 1e-3f * (threadIdx.x + i);
        h[i] = (static_cast<unsigned long long>(__float_as_uint(f[i])) << 32) | __float_as_uint(g[i]);
    }
    const float m = 0.999999f, c = 1e-7f;
    const unsigned long long m2 = (static_cast<unsigned long long>(__float_as_uint(m)) << 32) | __float_as_uint(m);
    const unsigned long long c2 = (static_cast<unsigned long long>(__float_as_uint(c)) << 32) | __float_as_uint(c);
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
#pragma unroll
            for (int r = 0; r < W; ++r) {
                unsigned long long p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[i]), "r"(0xD2511F53u));
                a[i] = static_cast<unsigned int>(p >> 32);
                b[i] = static_cast<unsigned int>(p);
            }
#pragma unroll
            for (int r = 0; r < F; ++r) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(m), "f"(c));
#pragma unroll
            for (int r = 0; r < F2; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(h[i]) : "l"(m2), "l"(c2));
#pragma unroll
            for (int r = 0; r < M; ++r) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g[i]));
#pragma unroll
            for (int r = 0; r < L; ++r)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[(i + 1) % kChains]), "r"(0x9E3779B9u));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i)
        s += f[i] + g[i] + static_cast<float>((a[i] ^ b[i]) & 1u) + static_cast<float>(h[i] & 1ull);
    if (s == 123.456f) sink[0] = s;
}

// FFMA (PACKED = 0) / FFMA2 (PACKED = 1) with three DISTINCT register operands per instruction
template <int PACKED>
__global__ void probe_fma3(long long iters, float* sink) {
    float a[kChains], b[kChains], c[kChains];
    unsigned long long ha[kChains], hb[kChains], hc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
        b[i] = 0.999999f - 1e-7f * (threadIdx.x + 3 * i);
        c[i] = 1e-7f * (1 + threadIdx.x + i);
        ha[i] = (static_cast<unsigned long long>(__float_as_uint(a[i])) << 32) | __float_as_uint(a[i]);
        hb[i] = (static_cast<unsigned long long>(__float_as_uint(b[i])) << 32) | __float_as_uint(b[i]);
        hc[i] = (static_cast<unsigned long long>(__float_as_uint(c[i])) << 32) | __float_as_uint(c[i]);
    }
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ha[i]) : "l"(hb[i]), "l"(hc[(i + 1) % kChains]));
            else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[(i + 1) % kChains]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] + static_cast<float>(ha[i] & 1ull);
    if (s == 123.456f) sink[0] = s;
}

}  // namespace

extern "C" {

int bmc_probe_ops_per_iteration(int kind) {
    // thread-level operations one loop iteration issues (bench.py turns time into a rate)
    switch (kind) {
        case 0: return kChains;          // FFMA
        case 1: return 2 * kChains;      // MUFU (ex2 + lg2)
        case 2: return 2 * kChains;      // IMAD.WIDE + LOP3
        case 3: return 2 * kChains;      // FFMA + LOP3
        case 4: case 5: case 6: case 7: return kChains;   // IMAD.WIDE(+IADD) / IMAD.HI / IMAD / LOP3 alone
        case 8: return kChains;          // FFMA2 alone (one packed instruction = two fp32 FMAs)
        case 9: return kChains;          // IMAD.WIDE alone
        case 10: return 2 * kChains;     // IMAD.WIDE + FFMA
        case 11: return 3 * kChains;     // IMAD.WIDE + 2 FFMA
        case 12: return 2 * kChains;     // IMAD.WIDE + FFMA2
        case 13: return 5 * kChains;     // MUFU + 4 FFMA
        case 14: return 3 * kChains;     // MUFU + 2 IMAD.WIDE
        case 15: return 2 * kChains;     // FFMA2 + FFMA
        case 16: return kChains;         // DFMA
        case 17: return kChains;         // DFMA, three distinct register operands
        case 18: return 2 * kChains;     // the same + one LOP3 each
        case 19: return kChains;         // FFMA, three distinct register operands
        case 20: return kChains;         // FFMA2, three distinct register pairs
        default: return 0;
    }
}

int bmc_probe(int kind, int64_t iters, int blocks, int threads, float* sink, void* stream) {
    if (!(kind >= 0 && kind <= 20 && iters > 0 && blocks > 0 && threads > 0 && threads <= 1024 && sink)) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (kind) {
        case 0: probe_ffma<<<blocks, threads, 0, st>>>(iters, sink); break;
        case 1: probe_mufu<<<blocks, threads, 0, st>>>(iters, sink); break;
        case 2: probe_philox_mix<<<blocks, threads, 0, st>>>(iters, sink); break;
        case 3: probe_issue<<<blocks, threads, 0, st>>>(iters, sink); break;
        case 4: probe_single<0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 5: probe_single<1><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 6: probe_single<2><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 7: probe_single<3><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 8: probe_mix<0, 0, 1, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 9: probe_mix<1, 0, 0, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 10: probe_mix<1, 1, 0, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 11: probe_mix<1, 2, 0, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 12: probe_mix<1, 0, 1, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 13: probe_mix<0, 4, 0, 1, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 14: probe_mix<2, 0, 0, 1, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 16: probe_dfma<<<blocks, threads, 0, st>>>(iters, sink); break;
        case 17: probe_dfma3<0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 18: probe_dfma3<1><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 19: probe_fma3<0><<<blocks, threads, 0, st>>>(iters, sink); break;
        case 20: probe_fma3<1><<<blocks, threads, 0, st>>>(iters, sink); break;
        default: probe_mix<0, 1, 1, 0, 0><<<blocks, threads, 0, st>>>(iters, sink); break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
