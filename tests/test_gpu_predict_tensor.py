"""fp32 prediction with a wide basis (K > 16): the contraction u . beta runs on the tensor cores
(tcgen05, split TF32).  Checked against the fp64 statement of pybmc/sampling_utils.py:64-77, against
the FFMA kernels on the same inputs, and for the exactness properties of the fused selection."""
import numpy as np
import pytest

from oracle import bmc_oracle as oc
from oracle import philox as px

pytestmark = pytest.mark.gpu


def _problem(k, n, s, m=None, seed=0, bimodal=False):
    rng = np.random.default_rng(500 + k + seed)
    m = m or k + 5
    preds = 800 + rng.normal(0, 3, size=(n, m))
    vt = rng.normal(size=(k, m)) * 0.03
    vt -= vt.mean(axis=1, keepdims=True)             # rows orthogonal to 1, as Vt_hat of row-centred predictions is
    beta = rng.normal(size=k)[None, :] + 0.15 * rng.normal(size=(s, k))
    if bimodal:
        beta[: s // 2] += 2.0                        # two posterior modes: the normal window guess misses
    theta = np.column_stack([beta, np.abs(rng.normal(0.2, 0.02, s))])
    truth = preds.mean(axis=1) + rng.normal(0, 0.5, n)
    return preds, vt, theta, truth


@pytest.mark.parametrize("k,n,s", [(17, 300, 3000), (40, 129, 1000), (64, 700, 2500), (32, 128, 128), (64, 5, 130),
                                   (3, 200, 1000), (8, 131, 900), (12, 64, 515), (16, 400, 2000)])
def test_contraction_accuracy_and_agreement_with_ffma(k, n, s):
    from pybmc_b200.sampling_utils import PredictiveProblem
    preds, vt, theta, truth = _problem(k, n, s)
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32")
    m = preds.shape[1]
    want = (theta[:, :k] @ vt + 1.0 / m) @ preds.T                   # sampling_utils.py:64-72 in fp64
    prob.tensor_min_k = 0                     # bmc_predict_problem.tensor_min_k: the default, tensor cores for every k
    tc = prob.run(noise="none", return_draws=True)
    prob.tensor_min_k = -1                    # the FFMA kernels
    ff = prob.run(noise="none", return_draws=True)
    # error budget of the fp32 path: inputs rounded to fp32 once, then K products
    u = preds @ vt.T
    scale = np.abs(theta[:, :k]) @ np.abs(u).T                       # sum_k |u||beta| per (draw, nucleus)
    err_tc = np.abs(tc.draws - want) / scale
    err_ff = np.abs(ff.draws - want) / scale
    assert err_tc.max() < 4e-7, err_tc.max()                         # fp32-level: 2^-23 = 1.2e-7 per rounding
    assert err_tc.max() < 3 * max(err_ff.max(), 1e-7)
    np.testing.assert_allclose(tc.draws, ff.draws, rtol=0, atol=6e-7 * scale.max())


@pytest.mark.parametrize("k", [6, 16, 24, 64])
def test_noise_counts_and_percentiles_are_exact_functions_of_the_draws(k):
    from pybmc_b200.sampling_utils import PredictiveProblem
    n, s = 333, 4000
    preds, vt, theta, truth = _problem(k, n, s, seed=1)
    q = [2.5, 16, 50, 84, 97.5]
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32", point0=11)
    full = prob.run(percentiles=q, seed=21, return_draws=True)
    lean = prob.run(percentiles=q, seed=21)
    assert np.array_equal(full.percentiles, lean.percentiles)
    np.testing.assert_allclose(lean.percentiles, np.percentile(full.draws, q, axis=0), rtol=1e-13)
    c_lt, c_le = oc.order_counts(full.draws, truth)
    assert np.array_equal(lean.c_lt, c_lt) and np.array_equal(lean.c_le, c_le)
    np.testing.assert_allclose(lean.mean, full.draws.mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(lean.var, full.draws.var(axis=0), rtol=2e-4)
    # the noise is the Philox contract of oracle/philox.py, keyed on the global nucleus index
    noiseless = prob.run(noise="none", return_draws=True)
    key = px.seed_key(21)
    z = np.empty((s, 8))
    for sb in range(s // 4):
        for j in range(8):
            z[4 * sb:4 * sb + 4, j] = px.noise_block(sb, 11 + j, key)
    got_z = (full.draws[:, :8] - noiseless.draws[:, :8]) / theta[:, -1:]
    np.testing.assert_allclose(got_z, z, atol=2e-3)                   # difference of two fp32 results / 0.2
    # against the fp64 path on the same stream
    f64 = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float64", point0=11).run(percentiles=q, seed=21)
    np.testing.assert_allclose(lean.percentiles, f64.percentiles, rtol=1e-5)
    # sharding: the second part of the nuclei on its own gives the same numbers
    shard = PredictiveProblem(preds[200:], theta, vt, truth=truth[200:], dtype="float32", point0=211).run(
        percentiles=q, seed=21)
    assert np.array_equal(shard.percentiles, lean.percentiles[:, 200:])
    assert np.array_equal(shard.c_lt, lean.c_lt[200:])


def test_retry_passes_regenerate_the_same_draws():
    """A bimodal posterior defeats the normal window guess: the retry passes (active list, slice
    counts) must see bit-identical draws from the tensor cores to land on the exact order statistics."""
    from pybmc_b200.sampling_utils import PredictiveProblem
    preds, vt, theta, truth = _problem(48, 260, 30000, seed=2, bimodal=True)
    q = [0, 2.5, 50, 97.5, 100]
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32")
    full = prob.run(percentiles=q, seed=4, return_draws=True)
    lean = prob.run(percentiles=q, seed=4)
    assert lean.passes > 1
    assert np.array_equal(full.percentiles, lean.percentiles)
    np.testing.assert_allclose(lean.percentiles, np.percentile(full.draws, q, axis=0), rtol=1e-13)
    c_lt, c_le = oc.order_counts(full.draws, truth)
    assert np.array_equal(lean.c_lt, c_lt) and np.array_equal(lean.c_le, c_le)


@pytest.mark.parametrize("k,n,s,q", [(1, 1, 5, [50.0]), (2, 3, 127, [0, 100]), (5, 130, 129, [1, 5, 25, 50, 75, 95, 99, 99.9]),
                                     (64, 257, 4, [2.5, 97.5]), (9, 1000, 1, [50.0])])
def test_small_and_ragged_shapes(k, n, s, q):
    """Fewer draws than one tile, one nucleus, eight percentiles (the widest template), a single draw."""
    from pybmc_b200.sampling_utils import PredictiveProblem
    preds, vt, theta, truth = _problem(k, n, s, seed=3)
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32", point0=5)
    full = prob.run(percentiles=q, seed=8, return_draws=True)
    lean = prob.run(percentiles=q, seed=8)
    assert full.draws.shape == (s, n)
    assert np.array_equal(full.percentiles, lean.percentiles)
    np.testing.assert_allclose(lean.percentiles, np.percentile(full.draws, q, axis=0), rtol=1e-13)
    c_lt, c_le = oc.order_counts(full.draws, truth)
    assert np.array_equal(lean.c_lt, c_lt) and np.array_equal(lean.c_le, c_le)
    f64 = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float64", point0=5).run(
        percentiles=q, seed=8, return_draws=True)
    np.testing.assert_allclose(full.draws, f64.draws, rtol=1e-5)
