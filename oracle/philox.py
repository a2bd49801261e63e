"""TEST INFRASTRUCTURE ONLY -- CPU statement of the device random-number contract.

This file is part of ``oracle/``: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it.
The product path (``pybmc_b200``) never does.

The reference draws from NumPy's MT19937 / PCG64 streams
(``pybmc/inference_utils.py:45,52,98,110,117,121,132,140`` and
``pybmc/sampling_utils.py:55-76``), whose per-iteration generator is seeded from
OS entropy, so no device stream can reproduce them.  The CUDA kernels instead use
a counter-based Philox4x32-10 stream.  This module restates that stream --
published algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11), Random123 ``philox4x32-10`` -- in plain Python integers, so that the
oracle sampler can be driven by exactly the variates a GPU chain sees and the two
can be compared number by number (tests/test_gibbs_parity_gpu.py).

Stream layout (DESIGN.md "RNG contract"):
  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (c0, c1, c2, c3)
     sampler : c0 = iteration index, c1 = block, c2 = global chain id, c3 = tag
               blocks 0, 1, ...    -> the iteration's word stream W[4 b + w].  With kp = K padded to 4, 8, 16, 32, 64
                                      and P = kp/2 Box-Muller pairs: pair p (normals 2p, 2p+1) takes its radius from
                                      W[p] and a 16-bit angle index from half (p & 1) of W[P + p//2] (48 bits a pair).
                                      kp = 8 only: W[6] = uniform of the first Gamma proposal, W[7] = Gamma word
                                      (iterations 2m, 2m+1 share one pair: radius W[7] of 2m, angle = low half of
                                      W[7] of 2m+1; even = cosine branch, odd = sine branch)
               block 0x10000       -> kp != 8: first Marsaglia-Tsang proposal of iterations 2m AND 2m+1 (counter
                                      c0 = 2m): even iteration = cosine branch + word 2, odd = sine branch + word 3
               block 0x10000 + t   -> attempt t >= 1 of the gamma draw of iteration c0 (after a rejection)
               block 0x18000       -> boost uniform for shapes < 1
               block 0x20000       -> Metropolis uniform (simplex sampler only)
     noise   : c0 = posterior-draw index s >> 2, c1 = global nucleus index, c2 = 0, c3 = tag
               -> 4 normals for draws 4*(s>>2) .. 4*(s>>2)+3 of that nucleus
  tag      : 1 conjugate Gibbs, 2 simplex sampler, 3 predictive noise
"""
import math

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

TAG_GIBBS = 1
TAG_SIMPLEX = 2
TAG_NOISE = 3
BLOCK_GAMMA = 0x10000
BLOCK_UNIFORM = 0x20000
GAMMA_MAX_ATTEMPTS = 64


def philox4x32_10(ctr, key):
    """Ten rounds of Philox-4x32; ``ctr`` four and ``key`` two 32-bit words."""
    c0, c1, c2, c3 = (int(x) & MASK for x in ctr)
    k0, k1 = (int(x) & MASK for x in key)
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, \
                         ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & MASK, seed >> 32


def u01(r):
    """32 random bits -> uniform on the open interval (0, 1): (r + 1/2) 2^-32."""
    return (int(r) + 0.5) * 2.0 ** -32


def box_muller(ra, rb):
    """Two 32-bit words -> two independent N(0,1) (cosine branch first)."""
    rad = math.sqrt(-2.0 * math.log(u01(ra)))
    ang = 2.0 * math.pi * u01(rb)
    return rad * math.cos(ang), rad * math.sin(ang)


def normals4(ctr, key):
    r = philox4x32_10(ctr, key)
    za, zb = box_muller(r[0], r[1])
    zc, zd = box_muller(r[2], r[3])
    return za, zb, zc, zd


def box_muller_h(ra, h):
    """A 32-bit radius word and a 16-bit angle index -> two independent N(0,1): angle 2 pi (h + 1/2) 2^-16."""
    rad = math.sqrt(-2.0 * math.log(u01(ra)))
    ang = 2.0 * math.pi * (int(h) + 0.5) * 2.0 ** -16
    return rad * math.cos(ang), rad * math.sin(ang)


def padded_components(k):
    """bmc_padded_components (pybmc_b200/csrc/gibbs.cu): the component count the kernels are instantiated for."""
    for kp in (4, 8, 16, 32, 64):
        if k <= kp:
            return kp
    raise ValueError("k > 64")


def variate_layout(kp):
    """(pairs, words, calls, gamma_inline) of an iteration's word stream -- rng.cuh VariateLayout<KP>."""
    pairs = kp // 2
    words = pairs + pairs // 2
    calls = (words + 3) // 4
    return pairs, words, calls, 4 * calls - words >= 2


def iteration_words(kp, it, chain, tag, key):
    calls = variate_layout(kp)[2]
    w = []
    for b in range(calls):
        w.extend(philox4x32_10((it, b, chain, tag), key))
    return w


def normal_vector(k, it, chain, tag, key):
    """The K standard normals of sampler iteration ``it`` (the first K of the kp the layout provides)."""
    kp = padded_components(k)
    pairs = variate_layout(kp)[0]
    w = iteration_words(kp, it, chain, tag, key)
    out = []
    for p in range(pairs):
        aw = w[pairs + p // 2]
        out.extend(box_muller_h(w[p], (aw >> 16) if p & 1 else (aw & 0xFFFF)))
    return out[:k]


BLOCK_BOOST = BLOCK_GAMMA + 0x8000


def gamma_unit_scale(shape, it, chain, tag, key, k=None):
    """Gamma(shape, 1) by Marsaglia & Tsang (2000) for a sampler with ``k`` components.

    The first proposals of iterations 2m and 2m+1 share one Box-Muller pair: the even iteration uses its
    cosine branch, the odd one its sine branch.  When the iteration's word stream leaves two words over
    (padded k = 8) the pair's radius is word 7 of iteration 2m, its angle index the low half of word 7 of
    iteration 2m+1, and each iteration's uniform its own word 6.  Otherwise (``k`` None or another padded
    size) they share Philox block (2m, BLOCK_GAMMA): words 0, 1 the pair, word 2 / word 3 the uniforms.
    Attempt t >= 1 (after a rejection) has block (it, BLOCK_GAMMA + t).  For shape < 1 the usual boost
    Gamma(a) = Gamma(a+1) * U^(1/a) takes U from block (it, BLOCK_BOOST).
    """
    a = float(shape)
    boost = a < 1.0
    if boost:
        a += 1.0
    d = a - 1.0 / 3.0
    c = 1.0 / math.sqrt(9.0 * d)
    odd = it & 1
    v = 1.0
    for t in range(GAMMA_MAX_ATTEMPTS):
        if t == 0 and k is not None and variate_layout(padded_components(k))[3]:
            kp = padded_components(k)
            _, words, _, _ = variate_layout(kp)
            even = iteration_words(kp, it - odd, chain, tag, key)
            nxt = iteration_words(kp, it - odd + 1, chain, tag, key)
            x = box_muller_h(even[words + 1], nxt[words + 1] & 0xFFFF)[odd]
            u = u01((nxt if odd else even)[words])
        elif t == 0:
            r = philox4x32_10((it - odd, BLOCK_GAMMA, chain, tag), key)
            x = box_muller(r[0], r[1])[odd]
            u = u01(r[2 + odd])
        else:
            r = philox4x32_10((it, BLOCK_GAMMA + t, chain, tag), key)
            x = box_muller(r[0], r[1])[0]
            u = u01(r[2])
        v = 1.0 + c * x
        if v <= 0.0:
            v = 1.0
            continue
        v = v * v * v
        x2 = x * x
        if u < 1.0 - 0.0331 * x2 * x2:
            break
        if math.log(u) < 0.5 * x2 + d * (1.0 - v + math.log(v)):
            break
    g = d * v
    if boost:
        g *= u01(philox4x32_10((it, BLOCK_BOOST, chain, tag), key)[0]) ** (1.0 / float(shape))
    return g


def metropolis_uniform(it, chain, tag, key):
    return u01(philox4x32_10((it, BLOCK_UNIFORM, chain, tag), key)[0])


def noise_block(sblock, nucleus, key):
    """Four N(0,1) for posterior draws 4*sblock .. 4*sblock+3 of one (global) nucleus."""
    return normals4((sblock, nucleus, 0, TAG_NOISE), key)
