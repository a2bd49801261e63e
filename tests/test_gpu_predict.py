"""Prediction / UQ parity on the GPU: the reference's own posterior rows and noise (regenerated
from the golden seeds, tests/golden/make_golden.py) are fed to the fused kernel; means, variances
and percentiles must match to 1e-10 (fp64) / 1e-5 (fp32) and the integer counts bit for bit."""
import numpy as np
import pandas as pd
import pytest

import cases
from oracle import bmc_oracle as oc
from oracle import philox as px

pytestmark = pytest.mark.gpu
LEVELS = np.arange(0, 101, 5)


def _reference_inputs(golden):
    g = golden("predict")
    preds, truth = cases.ensemble(12, 7, 5)
    theta_all = cases.posterior_like(13, 12000, 3)
    rng = cases.SeededFactory(3000)()
    theta = rng.choice(theta_all, 10000, replace=False)          # what the reference drew (:57)
    noise = rng.standard_normal((10000, preds.shape[0]))         # (:76)
    rndm_m, pct = oc.predictive_from_selected(preds, theta, g["Vt_hat"], noise)
    assert np.array_equal(rndm_m[:64], g["head"])                # same matrix the reference produced
    return g, preds, truth, theta, noise, rndm_m, pct


def test_fused_on_reference_samples_fp64(golden):
    from pybmc_b200.sampling_utils import PredictiveProblem, coverage_from_counts
    g, preds, truth, theta, noise, rndm_m, pct = _reference_inputs(golden)
    prob = PredictiveProblem(preds, theta, g["Vt_hat"], truth=truth, dtype="float64")
    res = prob.run(noise=noise, return_draws=True)
    np.testing.assert_allclose(res.draws, rndm_m, rtol=1e-12, atol=0)
    np.testing.assert_allclose(res.mean, rndm_m.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(res.var, rndm_m.var(axis=0), rtol=1e-9)
    for got, want in zip(res.percentiles, (g["lo"], g["med"], g["hi"])):
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=0)
    c_lt, c_le = oc.order_counts(rndm_m, truth)
    assert np.array_equal(res.c_lt, c_lt) and np.array_equal(res.c_le, c_le)
    assert coverage_from_counts(LEVELS, 10000, res.c_lt, res.c_le) == list(g["coverage"])
    assert res.passes >= 1


def test_fused_fp32_within_tolerance(golden):
    from pybmc_b200.sampling_utils import PredictiveProblem
    g, preds, truth, theta, noise, rndm_m, pct = _reference_inputs(golden)
    res = PredictiveProblem(preds, theta, g["Vt_hat"], truth=truth, dtype="float32").run(noise=noise)
    for got, want in zip(res.percentiles, (g["lo"], g["med"], g["hi"])):
        np.testing.assert_allclose(got, want, rtol=1e-5)
    np.testing.assert_allclose(res.mean, rndm_m.mean(axis=0), rtol=1e-5)
    c_lt, _ = oc.order_counts(rndm_m, truth)
    assert np.max(np.abs(res.c_lt - c_lt)) <= 3      # a draw within fp32 rounding of the truth may flip


def test_dropin_coverage_and_matrix_percentiles(golden):
    """coverage() and exact column percentiles on matrices the reference produced / was given."""
    import pybmc_b200 as pb
    g, preds, truth, theta, noise, rndm_m, pct = _reference_inputs(golden)
    assert pb.coverage(LEVELS, rndm_m, pd.DataFrame({"truth": truth}), "truth") == list(g["coverage"])
    res = pb.column_percentiles(rndm_m, [2.5, 50, 97.5, 0, 100, 16, 84, 33.3, 99.99, 0.01], truth=truth)
    want = np.percentile(rndm_m, [2.5, 50, 97.5, 0, 100, 16, 84, 33.3, 99.99, 0.01], axis=0)
    assert np.array_equal(res.percentiles, want)
    t = golden("coverage_ties")
    tdf = pd.DataFrame({"t": t["truth"]})
    assert pb.coverage(LEVELS, t["matrix"], tdf, "t") == list(t["coverage"])
    assert pb.coverage([1, 33, 68, 95, 99], t["matrix"][:137], tdf, "t") == list(t["coverage_odd"])
    res = pb.column_percentiles(t["matrix"], [2.5, 25, 50, 75, 97.5])          # heavy ties
    assert np.array_equal(res.percentiles, np.percentile(t["matrix"], [2.5, 25, 50, 75, 97.5], axis=0))
    res = pb.column_percentiles(t["matrix"][:5], [0, 50, 100])                  # five rows only
    assert np.array_equal(res.percentiles, np.percentile(t["matrix"][:5], [0, 50, 100], axis=0))


def test_philox_noise_matches_cpu_contract():
    """Noise regenerated on the device == oracle/philox.py, so the unmaterialised matrix is checkable."""
    from pybmc_b200.sampling_utils import PredictiveProblem
    preds, truth = cases.ensemble(31, 10, 4)                     # 10 points, shard starting at nucleus 7
    vt = np.linalg.qr(np.random.default_rng(1).normal(size=(4, 2)))[0].T * 0.05
    theta = cases.posterior_like(32, 600, 2)
    seed, point0 = 0xB203, 7
    res = PredictiveProblem(preds, theta, vt, truth=truth, point0=point0).run(seed=seed, return_draws=True)
    key = px.seed_key(seed)
    z = np.empty((600, 10))
    for sb in range(150):
        for n in range(10):
            z[4 * sb:4 * sb + 4, n] = px.noise_block(sb, point0 + n, key)
    want, pct = oc.predictive_from_selected(preds, theta, vt, z)
    np.testing.assert_allclose(res.draws, want, rtol=1e-11, atol=0)
    for got, w in zip(res.percentiles, pct):
        np.testing.assert_allclose(got, w, rtol=1e-10)
    c_lt, c_le = oc.order_counts(want, truth)
    assert np.array_equal(res.c_lt, c_lt) and np.array_equal(res.c_le, c_le)


def test_unmaterialised_equals_materialised_at_scale():
    """S = 10^5 draws (two-pass regime, no S-by-N matrix): the fused result equals the percentile of
    the matrix the same kernel can also write out, and does not depend on chunking / sharding."""
    from pybmc_b200.sampling_utils import PredictiveProblem
    rng = np.random.default_rng(44)
    n, k, s = 96, 16, 100_000
    preds = 1000 + rng.normal(0, 3, size=(n, 20))
    vt = rng.normal(size=(k, 20)) * 0.02
    theta = np.column_stack([rng.normal(0, 1, k)[None, :] + 0.1 * rng.normal(size=(s, k)),
                             np.abs(rng.normal(0.15, 0.01, s))])
    truth = preds.mean(axis=1) + rng.normal(0, 0.3, n)
    q = [2.5, 16, 50, 84, 97.5]
    prob = PredictiveProblem(preds, theta, vt, truth=truth)
    full = prob.run(percentiles=q, seed=7, return_draws=True)
    lean = prob.run(percentiles=q, seed=7, return_draws=False)
    assert np.array_equal(full.percentiles, lean.percentiles)
    np.testing.assert_allclose(lean.percentiles, np.percentile(full.draws, q, axis=0), rtol=1e-13, atol=0)
    c_lt, c_le = oc.order_counts(full.draws, truth)
    assert np.array_equal(lean.c_lt, c_lt) and np.array_equal(lean.c_le, c_le)
    np.testing.assert_allclose(lean.mean, full.draws.mean(axis=0), rtol=1e-12)
    # the second half of the points as its own shard (global index 48) gives the same numbers
    shard = PredictiveProblem(preds[48:], theta, vt, truth=truth[48:], point0=48).run(percentiles=q, seed=7)
    assert np.array_equal(shard.percentiles, lean.percentiles[:, 48:])
    assert np.array_equal(shard.c_lt, lean.c_lt[48:])
    # fp32 arithmetic stays within the fp32 tolerance of the north star
    f32 = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32").run(percentiles=q, seed=7)
    np.testing.assert_allclose(f32.percentiles, lean.percentiles, rtol=1e-5)


def test_windows_recover_from_bad_guesses():
    """Bimodal / heavy-tailed / degenerate columns: the normal-approximation window misses or
    overflows and the retry passes must still land on the exact order statistics."""
    import pybmc_b200 as pb
    rng = np.random.default_rng(9)
    s = 20000
    cols = [np.concatenate([rng.normal(-50, 0.1, s // 2), rng.normal(80, 0.1, s // 2)]),   # bimodal
            rng.standard_cauchy(s),                                                         # heavy tails
            np.full(s, 3.25),                                                               # constant
            np.repeat(rng.normal(size=20), s // 20),                                        # 20 atoms
            rng.exponential(1.0, s) ** 3,                                                   # skewed
            rng.normal(1e6, 1e-3, s)]                                                       # tiny spread, big offset
    mat = np.column_stack([rng.permutation(c) for c in cols])
    q = [0, 0.5, 2.5, 50, 97.5, 99.5, 100]
    res = pb.column_percentiles(mat, q)
    assert np.array_equal(res.percentiles, np.percentile(mat, q, axis=0))
    assert res.passes > 1


def test_rndm_m_random_calculator_dropin():
    import pybmc_b200 as pb
    from pybmc_b200.sampling_utils import rndm_m_random_calculator
    preds, truth = cases.ensemble(12, 7, 5)
    r = oc.orthogonalize_arrays(*cases.ensemble(11, 28, 5), 3)
    theta = cases.posterior_like(13, 12000, 3)
    rndm_m, (lo, med, hi) = rndm_m_random_calculator(preds, theta, r["Vt_hat"], seed=11)
    assert rndm_m.shape == (10000, 7) and rndm_m.dtype == np.float64
    np.testing.assert_allclose(lo, np.percentile(rndm_m, 2.5, axis=0), rtol=1e-13)
    np.testing.assert_allclose(med, np.percentile(rndm_m, 50, axis=0), rtol=1e-13)
    np.testing.assert_allclose(hi, np.percentile(rndm_m, 97.5, axis=0), rtol=1e-13)
    with pytest.raises(ValueError):                                  # < 10000 posterior rows (:57)
        rndm_m_random_calculator(preds, theta[:500], r["Vt_hat"])
    # distribution check against the reference path on NumPy's generator: 3 standard errors
    ref_m, (rlo, rmed, rhi) = oc.predictive_draws(preds, theta, r["Vt_hat"], np.random.default_rng(5))
    sd = ref_m.std(axis=0)
    assert np.all(np.abs(med - rmed) < 3 * 1.2533 * sd / np.sqrt(10000) * np.sqrt(2))
    assert np.all(np.abs(rndm_m.mean(axis=0) - ref_m.mean(axis=0)) < 3 * sd * np.sqrt(2 / 10000))
    assert np.all(np.abs(lo - rlo) < 3 * 2.7 * sd / np.sqrt(10000) * np.sqrt(2))


@pytest.mark.parametrize("k", [1, 5, 8, 12, 33])
def test_component_counts_and_precisions(k):
    """Every compiled component count (4, 8, 16, 32, 64) in both precisions; fp64 with 8 components
    needs exactly 48 KB of dynamic shared memory on top of the static barriers (opt-in path)."""
    from pybmc_b200.sampling_utils import PredictiveProblem
    rng = np.random.default_rng(100 + k)
    n, m, s = 37, k + 3, 1500
    preds = 500 + rng.normal(0, 2, size=(n, m))
    vt = rng.normal(size=(k, m)) * 0.05
    theta = np.column_stack([rng.normal(size=(s, k)) * 0.2 + rng.normal(size=k), np.abs(rng.normal(0.2, 0.02, s))])
    truth = preds.mean(axis=1) + rng.normal(0, 0.3, n)
    for dtype, tol in (("float64", 1e-12), ("float32", 1e-5)):
        prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype=dtype)
        res = prob.run(percentiles=[2.5, 50, 97.5], seed=3, return_draws=True)
        np.testing.assert_allclose(res.percentiles, np.percentile(res.draws, [2.5, 50, 97.5], axis=0), rtol=1e-13)
        noiseless = prob.run(noise="none", return_draws=True)
        want = (theta[:, :k] @ vt + 1.0 / m) @ preds.T             # sampling_utils.py:64-72
        np.testing.assert_allclose(noiseless.draws, want, rtol=tol if dtype == "float64" else 1e-5)
        c_lt, c_le = oc.order_counts(res.draws, truth)
        assert np.array_equal(res.c_lt, c_lt) and np.array_equal(res.c_le, c_le)
