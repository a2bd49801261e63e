"""Statistical soundness of the device random-number contract, checked on its CPU statement
(oracle/philox.py).  The kernels reproduce these variates value by value (tests/test_gpu_gibbs.py,
tests/test_gpu_predict.py), so what holds here holds on the device."""
import numpy as np
from scipy import stats

from oracle import philox as px


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert px.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert px.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert px.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_variate_layout_counts():
    """Calls per iteration of the packed word stream (rng.cuh VariateLayout): 48 bits per Box-Muller pair."""
    assert [px.variate_layout(kp) for kp in (4, 8, 16, 32, 64)] == [
        (2, 3, 1, False), (4, 6, 2, True), (8, 12, 3, False), (16, 24, 6, False), (32, 48, 12, False)]
    assert [px.padded_components(k) for k in (1, 3, 4, 5, 8, 9, 16, 17, 33, 64)] == [4, 4, 4, 8, 8, 16, 16, 32, 64, 64]


def test_normals_are_standard_and_uncorrelated():
    key = px.seed_key(20261018)
    for k in (3, 8, 16):
        z = np.array([px.normal_vector(k, it, chain, px.TAG_GIBBS, key) for chain in range(4) for it in range(2500)])
        flat = z.ravel()
        assert stats.kstest(flat, "norm").pvalue > 1e-3, k
        assert abs(flat.mean()) < 4 / np.sqrt(flat.size) and abs(flat.var() - 1) < 4 * np.sqrt(2 / flat.size), k
        corr = np.corrcoef(z.T)
        assert np.max(np.abs(corr - np.eye(k))) < 5 / np.sqrt(len(z)), k
        lag = np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]            # consecutive iterations
        assert abs(lag) < 5 / np.sqrt(len(z)), k
        # squares too: a Box-Muller pair shares its radius, and pairs 2q, 2q+1 share an angle WORD (two halves)
        c2 = np.corrcoef((z ** 2).T)
        assert np.max(np.abs(c2 - np.eye(k))) < 5 / np.sqrt(len(z)), k


def test_sixteen_bit_angles_leave_the_normal_law_alone():
    """A Box-Muller pair with 65,536 equally spaced angles: the marginal of r cos(theta) deviates from N(0,1) only
    through angular harmonics of order 65,536.  Checked where it would show first -- the tails and the fine
    structure of 4e5 variates -- and on the exact grid sums of cos^2 / cos^4."""
    key = px.seed_key(31)
    z = np.array([px.normal_vector(8, it, 0, px.TAG_GIBBS, key) for it in range(50000)]).ravel()
    assert stats.kstest(z, "norm").pvalue > 1e-3
    for t in (1.0, 2.0, 3.0):
        p = stats.norm.sf(t) * 2
        assert abs(np.mean(np.abs(z) > t) - p) < 5 * np.sqrt(p * (1 - p) / z.size), t
    assert abs(stats.kurtosis(z)) < 5 * np.sqrt(24 / z.size)
    ang = 2 * np.pi * (np.arange(65536) + 0.5) / 65536
    assert abs(np.mean(np.cos(ang) ** 2) - 0.5) < 1e-15 and abs(np.mean(np.cos(ang) ** 4) - 0.375) < 1e-15
    assert abs(np.mean(np.cos(ang) * np.sin(ang))) < 1e-15


def test_gamma_variates_follow_the_gamma_law():
    key = px.seed_key(7)
    for k in (None, 3, 8):                                      # Gamma block / Gamma block / inline words (padded k = 8)
        for shape in (0.4, 1.0, 2.5, 189.0, 1500.5):            # (nu0 + n)/2 for n = 377 and 3000 among them
            g = np.array([px.gamma_unit_scale(shape, it, 3, px.TAG_GIBBS, key, k) for it in range(4000)])
            assert stats.kstest(g, "gamma", args=(shape,)).pvalue > 1e-3, (k, shape)
            assert abs(g.mean() - shape) < 5 * np.sqrt(shape / len(g)), (k, shape)
            assert abs(np.corrcoef(g[:-1], g[1:])[0, 1]) < 5 / np.sqrt(len(g)), (k, shape)    # 2m and 2m+1 share a pair


def test_inline_gamma_words_are_independent_of_the_normals():
    """Padded k = 8: the Gamma proposal of an iteration comes from words 6, 7 of the same two Philox calls as its
    eight normals; the variates must not notice."""
    key = px.seed_key(13)
    its = range(6000)
    z = np.array([px.normal_vector(8, it, 2, px.TAG_GIBBS, key) for it in its])
    g = np.array([px.gamma_unit_scale(2.5, it, 2, px.TAG_GIBBS, key, 8) for it in its])
    for j in range(8):
        assert abs(np.corrcoef(z[:, j], g)[0, 1]) < 5 / np.sqrt(len(g))
        assert abs(np.corrcoef(z[:, j] ** 2, g)[0, 1]) < 5 / np.sqrt(len(g))


def test_uniforms_and_noise_blocks():
    key = px.seed_key(99)
    u = np.array([px.metropolis_uniform(it, 1, px.TAG_SIMPLEX, key) for it in range(8000)])
    assert 0 < u.min() and u.max() < 1 and stats.kstest(u, "uniform").pvalue > 1e-3
    z = np.array([px.noise_block(sb, n, key) for sb in range(500) for n in range(8)]).reshape(500, 8, 4)
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    # neighbouring nuclei and neighbouring draw blocks are independent streams
    assert abs(np.corrcoef(z[:, 0, :].ravel(), z[:, 1, :].ravel())[0, 1]) < 0.08
    assert abs(np.corrcoef(z[:-1, 0, 0], z[1:, 0, 0])[0, 1]) < 0.2
