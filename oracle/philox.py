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
               block j < 0x10000   -> normals 4j .. 4j+3 of the K-vector draw
               block 0x10000       -> first Marsaglia-Tsang proposal of iterations 2m AND 2m+1 (counter
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


def normal_vector(k, it, chain, tag, key):
    out = []
    for j in range((k + 3) // 4):
        out.extend(normals4((it, j, chain, tag), key))
    return out[:k]


BLOCK_BOOST = BLOCK_GAMMA + 0x8000


def gamma_unit_scale(shape, it, chain, tag, key):
    """Gamma(shape, 1) by Marsaglia & Tsang (2000).

    The first proposals of iterations 2m and 2m+1 share Philox block (2m, BLOCK_GAMMA): the even
    iteration uses the cosine branch of its Box-Muller pair and word 2 as the uniform, the odd one the
    sine branch and word 3.  Attempt t >= 1 (after a rejection) has block (it, BLOCK_GAMMA + t).  For
    shape < 1 the usual boost Gamma(a) = Gamma(a+1) * U^(1/a) takes U from block (it, BLOCK_BOOST).
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
        if t == 0:
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
