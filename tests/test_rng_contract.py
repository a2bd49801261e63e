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


def test_normals_are_standard_and_uncorrelated():
    key = px.seed_key(20261018)
    z = np.array([px.normal_vector(8, it, chain, px.TAG_GIBBS, key) for chain in range(4) for it in range(2500)])
    flat = z.ravel()
    assert stats.kstest(flat, "norm").pvalue > 1e-3
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size) and abs(flat.var() - 1) < 4 * np.sqrt(2 / flat.size)
    corr = np.corrcoef(z.T)
    assert np.max(np.abs(corr - np.eye(8))) < 5 / np.sqrt(len(z))
    lag = np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]            # consecutive iterations
    assert abs(lag) < 5 / np.sqrt(len(z))


def test_gamma_variates_follow_the_gamma_law():
    key = px.seed_key(7)
    for shape in (0.4, 1.0, 2.5, 189.0, 1500.5):            # (nu0 + n)/2 for n = 377 and 3000 among them
        g = np.array([px.gamma_unit_scale(shape, it, 3, px.TAG_GIBBS, key) for it in range(4000)])
        assert stats.kstest(g, "gamma", args=(shape,)).pvalue > 1e-3, shape
        assert abs(g.mean() - shape) < 5 * np.sqrt(shape / len(g))


def test_uniforms_and_noise_blocks():
    key = px.seed_key(99)
    u = np.array([px.metropolis_uniform(it, 1, px.TAG_SIMPLEX, key) for it in range(8000)])
    assert 0 < u.min() and u.max() < 1 and stats.kstest(u, "uniform").pvalue > 1e-3
    z = np.array([px.noise_block(sb, n, key) for sb in range(500) for n in range(8)]).reshape(500, 8, 4)
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    # neighbouring nuclei and neighbouring draw blocks are independent streams
    assert abs(np.corrcoef(z[:, 0, :].ravel(), z[:, 1, :].ravel())[0, 1]) < 0.08
    assert abs(np.corrcoef(z[:-1, 0, 0], z[1:, 0, 0])[0, 1]) < 0.2
