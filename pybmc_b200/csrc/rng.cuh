// Counter-based random numbers for the sampler and predictive kernels.
//
// Replaces the reference's NumPy call sites: np.random.multivariate_normal
// (pybmc/inference_utils.py:45,98,121), np.random.default_rng().gamma (:52,117,140),
// np.random.uniform (:110,132) and Generator.standard_normal (pybmc/sampling_utils.py:76).
// The stream layout (key, counter words, block numbers, tags) is the contract written
// down in DESIGN.md "RNG contract" and restated on the CPU in oracle/philox.py.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "fp64_tables.cuh"

namespace bmc {

#ifndef BMC_PHILOX_ROUNDS
#define BMC_PHILOX_ROUNDS 10     // timing experiments only (profiles/build_ab.sh): the contract is ten rounds
#endif
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

constexpr uint32_t kTagGibbs = 1u;
constexpr uint32_t kTagSimplex = 2u;
constexpr uint32_t kTagNoise = 3u;
constexpr uint32_t kBlockGamma = 0x10000u;
constexpr uint32_t kBlockUniform = 0x20000u;
constexpr int kGammaMaxAttempts = 64;

struct Philox4 {
    uint32_t x, y, z, w;
};

// Round keys of Philox-4x32-10: key + r * Weyl constant.  They only depend on the seed, so a kernel
// builds them once (they live in uniform registers) instead of re-deriving them in every call.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(uint32_t key0, uint32_t key1) {
    PhiloxKeys ks;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        ks.k0[r] = key0 + static_cast<uint32_t>(r) * kPhiloxW0;
        ks.k1[r] = key1 + static_cast<uint32_t>(r) * kPhiloxW1;
    }
    return ks;
}

// Philox-4x32, ten rounds (Salmon et al., SC'11).  Two IMAD.WIDE + two LOP3 per round.
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const PhiloxKeys& ks) {
#pragma unroll
    for (int r = 0; r < BMC_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ ks.k0[r];
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ ks.k1[r];
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    return philox4x32_10(c0, c1, c2, c3, philox_keys(k0, k1));
}

// (Keeping the loop-invariant products of the first two rounds in registers -- only counter word 0, the iteration,
//  changes inside a chain's loop -- was measured in round 2: 17 instead of 20 multiplies per call, but 9.47 against
//  9.27 ms: the nine registers cost more than the multiplies save.  profiles/r2_notes.md)

template <typename real>
struct Math;

// Two fp32 values in one 64-bit register pair, for Blackwell's packed FFMA2 / FADD2 / FMUL2 (PTX *.f32x2,
// sm_100+).  Each half is an ordinary round-to-nearest fp32 operation, so results are bit-identical to the
// scalar instructions; the gain is one issue slot for two operations.  ptxas folds pack2(x, x) and swapped
// halves into operand modifiers (R.F32, .F32x2.LO_HI): no moves are spent on broadcasts.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <>
struct Math<float> {
    // (r + 1/2) 2^-32, evaluated in one FFMA; never 0, at most 1.
    static __device__ __forceinline__ float u01(uint32_t r) {
        return fmaf(__uint2float_rn(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    }
    // Box-Muller on the MUFU pipe: lg2, sqrt, sin, cos (all .approx.ftz: the arguments are never denormal).
    static __device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& za, float& zb) {
        float l2, rad, sn, cs;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u01(ra)));                       // <= 0
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(l2 * -1.3862943611198906f));   // sqrt(-2 ln u)
        // angle in [0, 2 pi): inside the range where sin/cos.approx keep their 2^-20.9 absolute error
        const float ang = u01(rb) * 6.283185307179586f;
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
        za = rad * cs;
        zb = rad * sn;
    }
    // the same pair of normals (bit for bit) with the two-at-a-time steps packed: {za, zb}
    static __device__ __forceinline__ f32x2 box_muller2(uint32_t ra, uint32_t rb) {
        float ua, ub, l2, rad, sn, cs, arg, ang;
        unpack2(fma2(pack2(__uint2float_rn(ra), __uint2float_rn(rb)),
                     pack2(2.3283064365386963e-10f, 2.3283064365386963e-10f),
                     pack2(1.1641532182693481e-10f, 1.1641532182693481e-10f)), ua, ub);
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(ua));
        unpack2(mul2(pack2(l2, ub), pack2(-1.3862943611198906f, 6.283185307179586f)), arg, ang);
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(arg));
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
        return mul2(pack2(rad, rad), pack2(cs, sn));
    }
    // ---- sampler normals: radius from a 32-bit word, angle from a 16-bit half word, 2 pi (h + 1/2) 2^-16 -------
    // (A Box-Muller pair with N equally spaced angles has exactly normal marginals up to angular harmonics of
    //  order N, ~ (r/2)^N / N!: nothing at N = 65,536.  48 bits per pair instead of 64 is a quarter fewer Philox
    //  calls, the largest item of the thread-per-chain sampler.)
    // a 16-bit integer as a float, exactly, without the conversion unit: I2F.U16 runs on the XU pipe (a quarter of
    // the MUFU rate, and MUFU is the busiest pipe of the samplers); 2^23 + h as a bit pattern, minus 2^23, is one
    // LOP3 and one FADD
    static __device__ __forceinline__ float half_word_to_float(uint32_t h) {
        return __uint_as_float(h | 0x4B000000u) - 8388608.0f;
    }
    static __device__ __forceinline__ void box_muller_h(uint32_t ra, uint32_t h, float& za, float& zb) {
        float l2, rad, sn, cs;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u01(ra)));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(l2 * -1.3862943611198906f));
        const float ang = fmaf(half_word_to_float(h), 9.587379924285257e-05f, 4.7936899621426287e-05f);
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
        za = rad * cs;
        zb = rad * sn;
    }
    // the same pair, bit for bit, with the conversions packed: {za, zb}
    static __device__ __forceinline__ f32x2 box_muller2_h(uint32_t ra, uint32_t h) {
        float ua, ang, l2, rad, sn, cs;
        unpack2(fma2(pack2(__uint2float_rn(ra), half_word_to_float(h)),
                     pack2(2.3283064365386963e-10f, 9.587379924285257e-05f),
                     pack2(1.1641532182693481e-10f, 4.7936899621426287e-05f)), ua, ang);
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(ua));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(l2 * -1.3862943611198906f));
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
        return mul2(pack2(rad, rad), pack2(cs, sn));
    }
    // log of the uniform a 32-bit word stands for
    static __device__ __forceinline__ float log_u01(uint32_t r) { return log(u01(r)); }
    // natural log of a positive normal number: lg2.approx.ftz * ln 2 (what __logf does, minus its
    // denormal fix-up instructions)
    static __device__ __forceinline__ float log(float x) {
        float l2;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(x));
        return l2 * 0.6931471805599453f;
    }
    // e^x = 2^(x log2 e) on the MUFU pipe; results below 2^-126 flush to zero (they are compared with a
    // uniform that is never below 2^-33)
    static __device__ __forceinline__ float exp(float x) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
        return r;
    }
    static __device__ __forceinline__ float sqrt(float x) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float rsqrt(float x) {       // arguments are never denormal here
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    // a / b for a normal divisor: what div.approx does, minus its range fix-up instructions
    static __device__ __forceinline__ float div(float a, float b) { return a * rcp(b); }
    static __device__ __forceinline__ float pow(float a, float b) { return powf(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
};

#ifndef BMC_F64_SINCOS2
#define BMC_F64_SINCOS2 1      // 0: the round-2a form (one 128-entry table + series), for A/B runs
#endif
template <>
struct Math<double> {
    static __device__ __forceinline__ double u01(uint32_t r) {
        return (static_cast<double>(r) + 0.5) * 2.3283064365386963e-10;
    }
    // ---- fp64 variates straight from the random WORD.  The uniform is u = x 2^-33 with x = 2r + 1 a 33-bit odd
    // integer, so exponent, mantissa and a 7-bit table index come from integer instructions; what is left is a
    // degree-6 series on |q| <= 2^-8 (log) and degree-7/8 series on |delta| <= 2 pi / 256 (sine, cosine), instead
    // of libm's general-argument ::log (~40 instructions) and ::sincospi (~50).  Accuracy ~1e-16 absolute (the
    // oracle's math.log / math.cos on the rounded angle differ from these by <= 3e-15): the 2e-9 value-by-value
    // parity of the fp64 chains is untouched.  Tables: fp64_tables.cuh (generated: 2 x 2 KB for the log and the
    // whole-word angles, 2 x 4 KB for the samplers' half-word angles; read through L1).
    // log of the positive normal double with words (hi, lo), plus k_off ln 2.  Everything before the series is
    // 32-bit integer work on the upper word -- table index, "centre above sqrt 2" flag, the mantissa with its
    // exponent replaced -- and the exponent reaches the fp64 pipe through the 2^52 bit pattern (one DADD) instead of
    // an I2F conversion.  (Round 2a did the same on 64-bit integers: shifts, a count-leading-zeros and carries, ~31
    // dependent integer instructions per log with 2-5 cycles of stall each -- a quarter of the fp64 iteration's
    // statically scheduled stall cycles, and that kernel runs two warps per scheduler.)  Same values, bit for bit.
    static __device__ __forceinline__ double log_words(uint32_t hi, uint32_t lo, double k_off) {
        const uint32_t i = (hi >> 13) & 0x7Fu;
        const uint32_t up = i >= 53u ? 1u : 0u;                                       // centre above sqrt 2: take m / 2
        const double m = __hiloint2double(static_cast<int>((hi & 0x000FFFFFu) | (up ? 0x3FE00000u : 0x3FF00000u)),
                                          static_cast<int>(lo));
        const double2 t = __ldg(&kLogTab[i]);
        const double q = ::fma(m, t.x, -1.0);                                        // m / c_i - 1, one rounding
        double p = ::fma(q, -1.0 / 6.0, 0.2);
        p = ::fma(q, p, -0.25);
        p = ::fma(q, p, 1.0 / 3.0);
        p = ::fma(q, p, -0.5);
        // k = (biased exponent + up) - 1023 + k_off, exactly: 2^52 + n has n in its low word
        const double k = __hiloint2double(0x43300000, static_cast<int>((hi >> 20) + up)) - (4503599627370496.0 + 1023.0 - k_off);
        return ::fma(k, 0.6931471805599453, t.y) + ::fma(q * q, p, q);               // (k ln 2 + ln c_i) + log1p(q)
    }
    // log((r + 1/2) 2^-32) = log(x) - 33 ln 2 with x = 2r + 1 (33 bits): 2^52 + x is a bit pattern, x one DADD away
    static __device__ __forceinline__ double log_u01(uint32_t r) {
        const double x = __hiloint2double(static_cast<int>(0x43300000u | (r >> 31)), static_cast<int>((r << 1) | 1u)) -
                         4503599627370496.0;
        return log_words(static_cast<uint32_t>(__double2hiint(x)), static_cast<uint32_t>(__double2loint(x)), -33.0);
    }
    // natural log of a positive normal double, the same way (exponent and fraction straight from its bits)
    static __device__ __forceinline__ double log(double x) {
        return log_words(static_cast<uint32_t>(__double2hiint(x)), static_cast<uint32_t>(__double2loint(x)), 0.0);
    }
    // sine and cosine of 2 pi (r + 1/2) 2^-BITS for a BITS-bit integer r (32: a whole word, 16: half a word)
    template <int BITS>
    static __device__ __forceinline__ void sincos_index(uint32_t r, double& sn, double& cs) {
        if constexpr (BITS == 16 && BMC_F64_SINCOS2) {
            // half-word angles: w = 2r + 1 = 512 A + B, both parts from 256-entry tables (fp64_tables.cuh, 2 x 4 KB,
            // read through L1), sine and cosine by one rotation: 4 fp64 instructions instead of 15 + a conversion
            // (the series below), 2-level dependency instead of 8.  Correctly rounded table entries, so the result
            // is within ~2 ulp of the exact value of the angle.
            const double2 hi = __ldg(&kSinCosHi[r >> 8]);
            const double2 lo = __ldg(&kSinCosLo[r & 0xFFu]);
            sn = ::fma(hi.x, lo.y, hi.y * lo.x);
            cs = ::fma(hi.y, lo.y, -(hi.x * lo.x));
            return;
        }
        const unsigned long long w = 2ull * r + 1ull;                                // angle = 2 pi w 2^-(BITS+1)
        const int i = static_cast<int>(w >> (BITS - 6));
        const int d = static_cast<int>(w & ((1ull << (BITS - 6)) - 1ull)) - (1 << (BITS - 7));   // offset from the centre of slot i
        const double dl = static_cast<double>(d) * (BITS == 32 ? 0x1.921fb54442d18p-31 : 0x1.921fb54442d18p-15);   // 2 pi 2^-(BITS+1)
        const double d2 = dl * dl;
        double ps = ::fma(d2, -1.0 / 5040.0, 1.0 / 120.0);
        ps = ::fma(d2, ps, -1.0 / 6.0);
        const double sd = ::fma(d2 * dl, ps, dl);                                    // sin(delta)
        double pc = ::fma(d2, 1.0 / 40320.0, -1.0 / 720.0);
        pc = ::fma(d2, pc, 1.0 / 24.0);
        pc = ::fma(d2, pc, -0.5);
        const double cd = ::fma(d2, pc, 1.0);                                        // cos(delta)
        const double2 t = __ldg(&kSinCosTab[i]);                                     // (sin, cos) of the centre
        sn = ::fma(t.x, cd, t.y * sd);
        cs = ::fma(t.y, cd, -(t.x * sd));
    }
    static __device__ __forceinline__ void sincos_u01(uint32_t r, double& sn, double& cs) { sincos_index<32>(r, sn, cs); }
    static __device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, double& za, double& zb) {
        const double rad = sqrt(-2.0 * log_u01(ra));
        double s, c;
        sincos_u01(rb, s, c);
        za = rad * c;
        zb = rad * s;
    }
    // sampler normals: radius from a word, angle 2 pi (h + 1/2) 2^-16 from half a word (see Math<float>)
    static __device__ __forceinline__ void box_muller_h(uint32_t ra, uint32_t h, double& za, double& zb) {
        const double l2 = -2.0 * log_u01(ra);
        const double rad = BMC_F64_SINCOS2 ? l2 * rsqrt(l2) : sqrt(l2);   // ~2 ulp: sqrt()'s correction step is not needed here
        double s, c;
        sincos_index<16>(h, s, c);
        za = rad * c;
        zb = rad * s;
    }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    // ---- roots and quotients of POSITIVE NORMAL arguments, branch-free.  libm's ::sqrt / ::rsqrt / a / b test the
    // exponent range and call a slow path; 17 such call sites and 21 reconvergence points cut an fp64 iteration
    // into ~40 basic blocks that ptxas cannot schedule across (the fp64 sampler runs two warps per scheduler and
    // lives on instruction-level parallelism).  Every argument here is a variance, a Gamma variate, a Cholesky
    // pivot or -2 ln u: never zero, negative or subnormal.  MUFU seed (2^-20: it reads the upper word only) and one
    // step of third order, error ~ 0.3 (2^-20)^3 below the rounding of the last multiply: results within ~1.5 ulp
    // of the correctly rounded ones (the oracle's math.sqrt), against a 2e-9 parity requirement.
    static __device__ __forceinline__ double rsqrt(double x) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double e = ::fma(-(x * y), y, 1.0);                     // 1 - x y^2
        return ::fma(y * e, ::fma(e, 0.375, 0.5), y);                 // y (1 + e/2 + 3 e^2/8)
    }
    static __device__ __forceinline__ double sqrt(double x) {
        const double y = rsqrt(x);
        const double g = x * y;
        return ::fma(::fma(-g, g, x), 0.5 * y, g);                    // one correction: the residual is exact in the fma
    }
    static __device__ __forceinline__ double rcp(double x) {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double e = ::fma(-x, y, 1.0);
        return ::fma(y, ::fma(e, e, e), y);                           // y (1 + e + e^2)
    }
    static __device__ __forceinline__ double div(double a, double b) {
        const double y = rcp(b);
        const double q = a * y;
        return ::fma(::fma(-b, q, a), y, q);                          // residual correction
    }
    static __device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
};

template <typename real, typename Key>
__device__ __forceinline__ void normals4_k(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const Key& ks,
                                           real (&z)[4]) {
    const Philox4 r = philox4x32_10(c0, c1, c2, c3, ks);
    Math<real>::box_muller(r.x, r.y, z[0], z[1]);
    Math<real>::box_muller(r.z, r.w, z[2], z[3]);
}
template <typename real>
__device__ __forceinline__ void normals4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                         uint32_t k1, real (&z)[4]) {
    normals4_k<real>(c0, c1, c2, c3, philox_keys(k0, k1), z);
}

// Gamma(shape, 1): Marsaglia & Tsang (2000).  shape >= 1 is the sampler's case
// ((nu0 + n)/2); smaller shapes use the U^(1/a) boost.  d and c are passed in because they
// are the same for every draw of a run.
template <typename real>
struct GammaConst {
    real d, c, inv_shape;
    int boost;
};

template <typename real>
__host__ __device__ inline GammaConst<real> make_gamma_const(double shape) {
    GammaConst<real> g;
    g.boost = shape < 1.0;
    const double a = g.boost ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0;
    g.d = static_cast<real>(d);
#ifdef __CUDA_ARCH__
    g.c = static_cast<real>(1.0 / ::sqrt(9.0 * d));
#else
    g.c = static_cast<real>(1.0 / __builtin_sqrt(9.0 * d));
#endif
    g.inv_shape = static_cast<real>(1.0 / shape);
    return g;
}

// Scalar constants of a sampler run in the arithmetic type of the kernel.  They travel in the kernel
// parameters (constant bank), so the loop uses them as instruction operands: no registers, and no
// per-iteration fp64 -> fp32 conversions when the compiler rematerialises them.
template <typename T>
struct RunConsts {
    T rss_min, prior_scale, sigma_ref, sigma2_init;
    T gamma_d, gamma_c, gamma_inv_shape;
};
template <typename T>
inline RunConsts<T> make_run_consts(double rss_min, double prior_scale, double sigma_ref, double sigma2_init,
                                    double shape) {
    const GammaConst<T> g = make_gamma_const<T>(shape);
    return RunConsts<T>{static_cast<T>(rss_min), static_cast<T>(prior_scale), static_cast<T>(sigma_ref),
                        static_cast<T>(sigma2_init), g.d, g.c, g.inv_shape};
}
template <typename T>
__device__ __forceinline__ GammaConst<T> gamma_const_of(const RunConsts<T>& c, int boost) {
    GammaConst<T> g;
    g.d = c.gamma_d;
    g.c = c.gamma_c;
    g.inv_shape = c.gamma_inv_shape;
    g.boost = boost;
    return g;
}

// ---- the variates of one sampler iteration (tags 1 and 2) ------------------------------------
// With KP the padded component count (4, 8, 16, 32, 64) and P = KP/2 Box-Muller pairs, the iteration reads the
// word stream W[4 b + w] = word w of Philox(iteration, block b, chain, tag), b = 0, 1, ...:
//     pair p  (normals 2p, 2p+1):  radius from W[p], angle from half (p & 1) of W[P + p/2]      [48 bits a pair]
//     KP = 8 only (the stream ends two words short of a call):  W[6] = uniform of the iteration's first Gamma
//         proposal, W[7] = its Gamma word -- iterations 2m and 2m+1 share ONE Box-Muller pair, radius from W[7]
//         of 2m, angle from the low half of W[7] of 2m+1; 2m takes the cosine branch, 2m+1 the sine branch.
//     otherwise the Gamma block (2m, kBlockGamma) as described below.
// Calls per iteration: 1 (KP = 4), 2 (8, Gamma included), 3, 6, 12 (+ 1/2 for the Gamma block).
template <int KP>
struct VariateLayout {
    static constexpr int kPairs = KP / 2;
    static constexpr int kWords = kPairs + kPairs / 2;
    static constexpr int kCalls = (kWords + 3) / 4;
    static constexpr bool kGammaInline = 4 * kCalls - kWords >= 2;       // KP == 8
    static constexpr int kUniformWord = kWords, kGammaWord = kWords + 1;
};
__device__ __forceinline__ uint32_t philox_word(const Philox4& r, int w) {
    return w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w;
}
template <int KP, typename Key>
__device__ __forceinline__ void iteration_words(uint32_t it, uint32_t chain, uint32_t tag, const Key& ks,
                                                Philox4 (&w)[VariateLayout<KP>::kCalls]) {
#pragma unroll
    for (int b = 0; b < VariateLayout<KP>::kCalls; ++b) w[b] = philox4x32_10(it, static_cast<uint32_t>(b), chain, tag, ks);
}
template <typename real, int KP>
__device__ __forceinline__ void normals_of_words(const Philox4 (&w)[VariateLayout<KP>::kCalls], real (&z)[KP]) {
    constexpr int P = VariateLayout<KP>::kPairs;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t ra = philox_word(w[p / 4], p % 4);
        const uint32_t aw = philox_word(w[(P + p / 2) / 4], (P + p / 2) % 4);
        Math<real>::box_muller_h(ra, (p & 1) ? aw >> 16 : aw & 0xFFFFu, z[2 * p], z[2 * p + 1]);
    }
}
template <typename real, int KP, typename Key>
__device__ __forceinline__ void iteration_normals(uint32_t it, uint32_t chain, uint32_t tag, const Key& ks,
                                                  real (&z)[KP]) {
    Philox4 w[VariateLayout<KP>::kCalls];
    iteration_words<KP>(it, chain, tag, ks, w);
    normals_of_words<real, KP>(w, z);
}

// ---- Gamma(shape, 1) by Marsaglia & Tsang (2000) -------------------------------------------------
// Variates used: a standard normal x and a uniform u per attempt.  The FIRST attempts of iterations
// 2m and 2m+1 share one Philox block, (2m, kBlockGamma): the even iteration takes the cosine branch of
// the Box-Muller pair and word z, the odd one the sine branch and word w -- every bit of the block is
// used, and a kernel that walks the iterations in order makes the call only every other iteration.
// Later attempts t >= 1 (rejections: < 1 % for the sampler's shapes) have their own blocks
// (it, kBlockGamma + t); the boost uniform of shapes < 1 has block (it, kBlockBoost).
constexpr uint32_t kBlockBoost = kBlockGamma + 0x8000u;

template <typename real>
struct GammaPair {
    real x[2], u[2], lu[2];     // normal, uniform and log(uniform) of the first proposals of iterations 2m, 2m+1
};

// the pair from the words of block (it_even, kBlockGamma)
template <typename real>
__device__ __forceinline__ GammaPair<real> gamma_pair_of(const Philox4& r) {
    GammaPair<real> p;
    Math<real>::box_muller(r.x, r.y, p.x[0], p.x[1]);
    p.u[0] = Math<real>::u01(r.z);
    p.u[1] = Math<real>::u01(r.w);
    if constexpr (sizeof(real) == 8) {       // fp64: from the word (no general-argument log); fp32 takes lg2(u) on use
        p.lu[0] = Math<real>::log_u01(r.z);
        p.lu[1] = Math<real>::log_u01(r.w);
    } else {
        p.lu[0] = p.lu[1] = real(0);
    }
    return p;
}
template <typename real, typename Key>
__device__ __forceinline__ GammaPair<real> gamma_pair(uint32_t it_even, uint32_t chain, uint32_t tag, const Key& ks) {
    return gamma_pair_of<real>(philox4x32_10(it_even, kBlockGamma, chain, tag, ks));
}

// accept / reject one proposal; on acceptance v holds (1 + c x)^3.  Squeeze and full test are both
// evaluated (a warp nearly always has a lane that needs the full test) and combined without branches.
template <typename real>
__device__ __forceinline__ bool gamma_accept(const GammaConst<real>& g, real x, real u, real log_u, real& v) {
    using M = Math<real>;
    const real t = M::fma(g.c, x, real(1));
    const real v3 = t * t * t;
    const real x2 = x * x;
    const bool squeeze = u < real(1) - real(0.0331) * x2 * x2;                 // almost always
    const bool full = log_u < real(0.5) * x2 + g.d * (real(1) - v3 + M::log(v3));       // NaN (t <= 0): false
    const bool ok = (t > real(0)) & (squeeze | full);
    v = ok ? v3 : real(1);
    return ok;
}

// rejections are rare: out of line, keys rebuilt from the seed
template <typename real>
__device__ __noinline__ void gamma_retry(const GammaConst<real>& g, uint32_t it, uint32_t chain, uint32_t tag,
                                         uint32_t k0, uint32_t k1, real& v) {
    const PhiloxKeys ks = philox_keys(k0, k1);
    for (uint32_t t = 1; t < static_cast<uint32_t>(kGammaMaxAttempts); ++t) {
        const Philox4 r = philox4x32_10(it, kBlockGamma + t, chain, tag, ks);
        real x, unused;
        Math<real>::box_muller(r.x, r.y, x, unused);
        if (gamma_accept<real>(g, x, Math<real>::u01(r.z), Math<real>::log_u01(r.z), v)) return;
    }
}

// finish a draw whose first proposal (x, u) is already known
template <typename real>
__device__ __forceinline__ real gamma_from_first(const GammaConst<real>& g, real x, real u, real log_u, uint32_t it,
                                                 uint32_t chain, uint32_t tag, uint32_t k0, uint32_t k1) {
    using M = Math<real>;
    real v;
    if constexpr (sizeof(real) == 4) log_u = M::log(u);
    if (!gamma_accept<real>(g, x, u, log_u, v)) gamma_retry<real>(g, it, chain, tag, k0, k1, v);
    real out = g.d * v;
    if (g.boost) out *= M::pow(M::u01(philox4x32_10(it, kBlockBoost, chain, tag, k0, k1).x), g.inv_shape);
    return out;
}

// Both draws of an iteration pair at once, with ONE rarely-taken branch for everything that is rare (a rejected
// first proposal: < 1e-3 at the sampler's shapes; the boost of shapes < 1).  The thread-per-chain kernels used to
// branch twice per iteration (retry?, boost?): four basic-block cuts per pair that ptxas could not schedule
// across, with instruction-fetch stalls behind each (profiles/r2_notes.md).  Same values as gamma_from_first.
template <typename real>
__device__ __noinline__ real gamma_finish_slow(const GammaConst<real> g, uint32_t it, uint32_t chain, uint32_t tag,
                                               uint32_t k0, uint32_t k1, bool ok, real out) {
    if (!ok) {
        real v;
        gamma_retry<real>(g, it, chain, tag, k0, k1, v);
        out = g.d * v;
    }
    if (g.boost) out *= Math<real>::pow(Math<real>::u01(philox4x32_10(it, kBlockBoost, chain, tag, k0, k1).x), g.inv_shape);
    return out;
}
template <typename real>
__device__ __forceinline__ void gamma_pair_finish(const GammaConst<real>& g, const GammaPair<real>& p, uint32_t it_even,
                                                  uint32_t chain, uint32_t tag, uint32_t k0, uint32_t k1, real& g0,
                                                  real& g1) {
    using M = Math<real>;
    real v0, v1;
    const bool ok0 = gamma_accept<real>(g, p.x[0], p.u[0], sizeof(real) == 4 ? M::log(p.u[0]) : p.lu[0], v0);
    const bool ok1 = gamma_accept<real>(g, p.x[1], p.u[1], sizeof(real) == 4 ? M::log(p.u[1]) : p.lu[1], v1);
    g0 = g.d * v0;
    g1 = g.d * v1;
    if (!(ok0 & ok1) || g.boost) {
        g0 = gamma_finish_slow<real>(g, it_even, chain, tag, k0, k1, ok0, g0);      // by value: nothing of the
        g1 = gamma_finish_slow<real>(g, it_even + 1u, chain, tag, k0, k1, ok1, g1); // hot path goes through memory
    }
}

// KP = 8: the pair from the last words of the angle calls of iterations 2m (`even`) and 2m+1 (`odd`)
template <typename real>
__device__ __forceinline__ GammaPair<real> gamma_pair_inline(const Philox4& even, const Philox4& odd) {
    GammaPair<real> p;
    Math<real>::box_muller_h(even.w, odd.w & 0xFFFFu, p.x[0], p.x[1]);
    p.u[0] = Math<real>::u01(even.z);
    p.u[1] = Math<real>::u01(odd.z);
    if constexpr (sizeof(real) == 8) {
        p.lu[0] = Math<real>::log_u01(even.z);
        p.lu[1] = Math<real>::log_u01(odd.z);
    } else {
        p.lu[0] = p.lu[1] = real(0);
    }
    return p;
}

// stand-alone draw for iteration `it` of a sampler with KP padded components (kernels that do not walk the
// iterations in order)
template <typename real, int KP>
__device__ __forceinline__ real gamma_unit_scale(const GammaConst<real>& g, uint32_t it, uint32_t chain,
                                                 uint32_t tag, const PhiloxKeys& ks, uint32_t k0, uint32_t k1) {
    GammaPair<real> p;
    if constexpr (VariateLayout<KP>::kGammaInline) {
        constexpr uint32_t blk = VariateLayout<KP>::kCalls - 1;
        p = gamma_pair_inline<real>(philox4x32_10(it & ~1u, blk, chain, tag, ks), philox4x32_10(it | 1u, blk, chain, tag, ks));
    } else {
        p = gamma_pair<real>(it & ~1u, chain, tag, ks);
    }
    const int odd = static_cast<int>(it & 1u);
    return gamma_from_first<real>(g, odd ? p.x[1] : p.x[0], odd ? p.u[1] : p.u[0], odd ? p.lu[1] : p.lu[0], it, chain,
                                  tag, k0, k1);
}
template <typename real, int KP>
__device__ __forceinline__ real gamma_unit_scale(const GammaConst<real>& g, uint32_t it, uint32_t chain,
                                                 uint32_t tag, uint32_t k0, uint32_t k1) {
    return gamma_unit_scale<real, KP>(g, it, chain, tag, philox_keys(k0, k1), k0, k1);
}
// the same for a caller that already holds the iteration's own calls `w` (iteration_words): with the inline
// layout only the partner iteration's angle call is still to be made
template <typename real, int KP, typename Key>
__device__ __forceinline__ real gamma_unit_scale_w(const GammaConst<real>& g,
                                                   const Philox4 (&w)[VariateLayout<KP>::kCalls], uint32_t it,
                                                   uint32_t chain, uint32_t tag, const Key& ks, uint32_t k0,
                                                   uint32_t k1) {
    GammaPair<real> p;
    const bool odd = (it & 1u) != 0u;
    if constexpr (VariateLayout<KP>::kGammaInline) {
        constexpr int blk = VariateLayout<KP>::kCalls - 1;
        const Philox4 other = philox4x32_10(it ^ 1u, static_cast<uint32_t>(blk), chain, tag, ks);
        p = gamma_pair_inline<real>(odd ? other : w[blk], odd ? w[blk] : other);
    } else {
        p = gamma_pair<real>(it & ~1u, chain, tag, ks);
    }
    return gamma_from_first<real>(g, odd ? p.x[1] : p.x[0], odd ? p.u[1] : p.u[0], odd ? p.lu[1] : p.lu[0], it, chain,
                                  tag, k0, k1);
}

}  // namespace bmc
