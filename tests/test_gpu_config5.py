"""BASELINE configs[4] shape (256 models, K = 64 components) at a size the CPU oracle finishes in
seconds: orthogonalisation, sampler and fused prediction all through the 64-component kernels."""
import numpy as np
import pytest

from oracle import bmc_oracle as oc
from oracle import philox as px

pytestmark = pytest.mark.gpu


def _ensemble(n=3000, m=256, rank=64, seed=1005):
    rng = np.random.default_rng(seed)
    spectrum = np.logspace(0, -1.5, rank)
    latent = rng.normal(size=(n, rank)) * spectrum
    mix = rng.normal(size=(rank, m))
    base = rng.uniform(100, 2000, n)
    preds = base[:, None] + 30 * latent @ mix + 0.05 * rng.normal(size=(n, m))
    truth = base + 30 * latent @ rng.normal(size=rank) * 0.1 + rng.normal(0, 0.15, n)
    return preds, truth


def test_orthogonalize_sampler_predict_k64():
    import pybmc_b200 as pb
    from pybmc_b200.inference_utils import ConjugateSampler, _finish_samples
    from pybmc_b200.sampling_utils import PredictiveProblem
    preds, truth = _ensemble()
    k = 64
    ref = oc.orthogonalize_arrays(preds, truth, k, full_matrices=False)
    got = pb.orthogonalize_arrays(preds, truth, k)
    np.testing.assert_allclose(got["S_hat"], ref["S_hat"], rtol=1e-9)
    np.testing.assert_allclose(got["y"], ref["y"], rtol=1e-12, atol=1e-12)
    proj_g = got["U_hat"] @ got["U_hat"].T @ ref["y"]          # sign/rotation-free: projection of y
    proj_r = ref["U_hat"] @ ref["U_hat"].T @ ref["y"]
    np.testing.assert_allclose(proj_g, proj_r, rtol=0, atol=1e-9 * np.abs(proj_r).max())
    np.testing.assert_allclose(got["U_hat"].T @ got["U_hat"], np.eye(k), atol=1e-10)

    prior = [np.zeros(k), np.diag(got["S_hat"] ** 2), 1.0, 0.02]
    sampler = ConjugateSampler(got["y"], got["U_hat"], prior)
    samples, _, _ = sampler.run(12, n_chains=2, seed=9, dtype="float64", stats="none")
    chains = _finish_samples(samples, True).reshape(2, 12, k + 1)
    w_inv = np.linalg.inv(sampler.w)
    want = oc.gibbs_conjugate(got["y"], got["U_hat"], 12, prior, oc.PhiloxDraws(
        9, 1, px.TAG_GIBBS, lambda cov: sampler.w * np.sqrt(np.diag(w_inv @ cov @ w_inv.T))[None, :]))
    np.testing.assert_allclose(chains[1], want, rtol=1e-8, atol=1e-10)

    res = pb.run_gibbs(got["y"], got["U_hat"], 600, prior, n_chains=256, seed=3, dtype="float32", thin=60)
    c = got["U_hat"].T @ got["y"]
    s2 = res.mean[-1] ** 2
    assert np.allclose(res.mean[:k], c / (1 + s2 / got["S_hat"] ** 2), atol=4 * res.mean[-1] / np.sqrt(600 * 256) + 1e-3)

    theta = res.samples[:2000]
    new = preds[:70]
    prob = PredictiveProblem(new, theta, got["Vt_hat"], truth=truth[:70], dtype="float64")
    full = prob.run(percentiles=[2.5, 50, 97.5], seed=5, return_draws=True)
    np.testing.assert_allclose(full.percentiles, np.percentile(full.draws, [2.5, 50, 97.5], axis=0), rtol=1e-13)
    z = np.empty((2000, 70))
    key = px.seed_key(5)
    for sb in range(500):
        for n in range(70):
            z[4 * sb:4 * sb + 4, n] = px.noise_block(sb, n, key)
    want_draws, _ = oc.predictive_from_selected(new, theta, got["Vt_hat"], z)
    np.testing.assert_allclose(full.draws, want_draws, rtol=1e-10)
    c_lt, c_le = oc.order_counts(want_draws, truth[:70])
    assert np.array_equal(full.c_lt, c_lt) and np.array_equal(full.c_le, c_le)
    f32 = PredictiveProblem(new, theta, got["Vt_hat"], truth=truth[:70], dtype="float32").run(seed=5)
    np.testing.assert_allclose(f32.percentiles, full.percentiles, rtol=1e-5)
