"""The generated tables behind the fp64 Box-Muller of rng.cuh (pybmc_b200/csrc/fp64_tables.cuh), checked on the CPU
against 40-digit arithmetic: every entry correctly rounded, and the two-table rotation the kernels use for the 16-bit
angle index reproduces sin / cos of 2 pi (h + 1/2) / 65536 for ALL 65,536 values of h to a few ulp."""
import os
import re

import mpmath
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(name):
    text = open(os.path.join(ROOT, "pybmc_b200", "csrc", "fp64_tables.cuh")).read()
    body = text.split(f"{name}[")[1].split("};")[0]
    rows = re.findall(r"\{([^,{}]+),\s*([^{}]+)\}", body)
    return np.array([[float.fromhex(a.strip()), float.fromhex(b.strip())] for a, b in rows])


def test_sincos_tables_are_correctly_rounded():
    mpmath.mp.dps = 40
    hi, lo, centre = _table("kSinCosHi"), _table("kSinCosLo"), _table("kSinCosTab")
    assert hi.shape == (256, 2) and lo.shape == (256, 2) and centre.shape == (128, 2)
    for tab, angle in ((hi, lambda i: 2 * mpmath.pi * i / 256), (lo, lambda j: 2 * mpmath.pi * (2 * j + 1) / 2 ** 17),
                       (centre, lambda i: 2 * mpmath.pi * (mpmath.mpf(i) + 0.5) / 128)):
        for i, (s, c) in enumerate(tab):
            a = angle(i)
            for got, want in ((s, mpmath.sin(a)), (c, mpmath.cos(a))):
                # half an ulp of the exact value (quarter turns are exact; mpmath's own pi leaves 1e-40 there)
                assert abs(mpmath.mpf(got) - want) <= abs(want) * mpmath.mpf(2) ** -53 + mpmath.mpf(10) ** -35, (i, got)


def test_two_table_rotation_gives_every_half_word_angle():
    """rng.cuh::sincos_index<16>: h = 256 A + j, angle = 2 pi A / 256 + 2 pi (2 j + 1) 2^-17."""
    hi, lo = _table("kSinCosHi"), _table("kSinCosLo")
    h = np.arange(65536)
    a, j = h >> 8, h & 255
    sn = hi[a, 0] * lo[j, 1] + hi[a, 1] * lo[j, 0]
    cs = hi[a, 1] * lo[j, 1] - hi[a, 0] * lo[j, 0]
    # reference in extended precision (x87 long double: 64-bit mantissa), argument reduced exactly: the angle is
    # a dyadic fraction of a turn
    turn = (2 * h.astype(np.longdouble) + 1) / np.longdouble(2 ** 17)
    two_pi = np.longdouble(2) * np.longdouble("3.14159265358979323846264338327950288")
    want_s, want_c = np.sin(two_pi * turn), np.cos(two_pi * turn)
    assert np.max(np.abs(sn - want_s)) < 4e-16 and np.max(np.abs(cs - want_c)) < 4e-16
    assert np.max(np.abs(sn * sn + cs * cs - 1.0)) < 1e-15


def test_log_table_identity():
    """kLogTab[i] = (fl(1/c_i), -ln of exactly that double): ln m = lg + log1p(m inv - 1) for m in the slot."""
    mpmath.mp.dps = 40
    tab = _table("kLogTab")
    assert tab.shape == (128, 2)
    for i, (inv, lg) in enumerate(tab):
        assert abs(mpmath.mpf(lg) + mpmath.log(mpmath.mpf(inv))) <= abs(mpmath.mpf(lg)) * mpmath.mpf(2) ** -53 + mpmath.mpf(10) ** -45
        c = 1 + (i + 0.5) / 128
        c = c / 2 if c > 2 ** 0.5 else c
        assert abs(inv * c - 1.0) < 1e-15            # the reciprocal of the slot's centre
