"""Sampler parity on the GPU.

Deterministic part: a device chain and the oracle driven by the same Philox stream must agree
value by value (the oracle redoes the reference's per-iteration inverse / residual literally).
Stochastic part: device posterior moments vs replicated runs of the oracle on NumPy's own
generators (= the reference sampler, see tests/test_oracle_golden.py) within 3 Monte-Carlo
standard errors.
"""
import numpy as np
import pytest

import cases
from oracle import bmc_oracle as oc
from oracle import philox as px

pytestmark = pytest.mark.gpu


def _ens():
    frame, models = cases.ensemble_frame(11, 40, 5)
    tr = frame.iloc[:28]
    return oc.orthogonalize_arrays(tr[models].values, tr["truth"].values, 3)


def _w_factor(w):
    """cov = W diag(v) W'  ->  W diag(sqrt v): the colouring the sufficient-statistic kernel uses."""
    w_inv = np.linalg.inv(w)

    def factor(cov):
        v = np.diag(w_inv @ cov @ w_inv.T)
        return w * np.sqrt(v)[None, :]
    return factor


CASES = {
    "toy_identity_prior": lambda: (*cases.toy_regression(), (np.array([0.0, 0.0]), np.eye(2), 1.0, 1.0)),
    "toy_dense_prior": lambda: (*cases.toy_regression(),
                                (np.array([0.3, -0.2]), np.array([[2.0, 0.3], [0.3, 0.5]]), 2.5, 0.7)),
    "ensemble_default": lambda: (lambda r: (r["y"], r["U_hat"],
                                            [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02]))(_ens()),
}


def _wide(k, n=60, seed=5):
    """Non-orthonormal design with k components, dense prior covariance, non-zero prior mean."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, k)) @ (np.eye(k) + 0.3 * rng.normal(size=(k, k)))
    y = X @ rng.normal(size=k) + 0.3 * rng.normal(size=n)
    a = rng.normal(size=(k, k))
    return y, X, (0.1 * rng.normal(size=k), a @ a.T / k + np.eye(k), 1.5, 0.3)


CASES["wide_k8"] = lambda: _wide(8)          # kernels compiled for 8 components
CASES["wide_k11"] = lambda: _wide(11)        # padded to 16
CASES["wide_k20"] = lambda: _wide(20, n=90)  # padded to 32


@pytest.mark.parametrize("name", list(CASES))
def test_conjugate_chain_matches_oracle_fp64(name):
    from pybmc_b200.inference_utils import ConjugateSampler, _finish_samples
    y, X, prior = CASES[name]()
    X = np.asarray(X, dtype=float)
    sampler = ConjugateSampler(y, X, prior)
    T, C, seed = (150, 3, 0xB200) if X.shape[1] <= 8 else (40, 2, 0xB200)
    samples, _, _ = sampler.run(T, n_chains=C, seed=seed, dtype="float64", stats="none")
    got = _finish_samples(samples, True).reshape(C, T, -1)
    for c in range(C):
        ref = oc.gibbs_conjugate(y, X, T, prior, oc.PhiloxDraws(seed, c, px.TAG_GIBBS, _w_factor(sampler.w)))
        np.testing.assert_allclose(got[c], ref, rtol=2e-9, atol=1e-12)


def test_conjugate_fp32_tracks_fp64():
    import pybmc_b200 as pb
    r = _ens()
    prior = [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02]
    a = pb.gibbs_sampler(r["y"], r["U_hat"], 50, prior, seed=5, dtype="float64")
    b = pb.gibbs_sampler(r["y"], r["U_hat"], 50, prior, seed=5, dtype="float32")
    assert a.shape == b.shape == (50, 4) and a.dtype == np.float64
    # same Philox stream, fp32 arithmetic: 1e-5 relative (north-star fp32 tolerance) on the scale of each column
    scale = np.abs(a).max(axis=0)
    assert np.max(np.abs(a - b) / scale) < 2e-4
    assert np.median(np.abs(a - b) / scale) < 1e-5


def test_chain_ids_are_global():
    """Chains are keyed by their global id: one launch of 8 == two launches of 4 (what sharding relies on)."""
    import pybmc_b200 as pb
    y, X, prior = CASES["toy_dense_prior"]()
    whole = pb.run_gibbs(y, X, 40, prior, n_chains=8, seed=3).samples.reshape(8, 40, 3)
    lo = pb.run_gibbs(y, X, 40, prior, n_chains=4, seed=3, chain_offset=0).samples.reshape(4, 40, 3)
    hi = pb.run_gibbs(y, X, 40, prior, n_chains=4, seed=3, chain_offset=4).samples.reshape(4, 40, 3)
    assert np.array_equal(whole[:4], lo) and np.array_equal(whole[4:], hi)


def test_conjugate_layouts_agree():
    """A handful of chains get a warp each, a few thousand eight lanes each, many chains one thread each:
    same stream, same chains (up to the summation order of RSS), same moment sums."""
    import pybmc_b200 as pb
    for name in ("ensemble_default", "wide_k8"):
        y, X, prior = CASES[name]()
        width = np.asarray(X).shape[1] + 1
        many = pb.run_gibbs(y, X, 50, prior, n_chains=16384, seed=13, stats="full")
        few = pb.run_gibbs(y, X, 50, prior, n_chains=6, seed=13, stats="full")
        mid = pb.run_gibbs(y, X, 50, prior, n_chains=2000, seed=13, stats="full")
        a = many.samples.reshape(16384, 50, width)[:6]
        b = few.samples.reshape(6, 50, width)
        c = mid.samples.reshape(2000, 50, width)[:6]
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(c, b, rtol=1e-9, atol=1e-12)
        for res in (few, mid, many):
            smp = res.samples
            np.testing.assert_allclose(res.mean, smp.mean(axis=0), rtol=1e-9, atol=1e-12)
            want = np.cov(smp.T, ddof=0)
            np.testing.assert_allclose(res.cov, want, rtol=1e-7, atol=1e-9 * np.abs(want).max())
        d = pb.run_gibbs(y, X, 50, prior, n_chains=6, seed=13, stats="diag")
        np.testing.assert_allclose(d.mean, few.mean, rtol=1e-12)


def test_rhat_from_device_moments():
    """R-hat from per-chain device moments == R-hat computed from the stored samples; the fast-mixing
    conjugate sampler sits at 1, short simplex chains started at b = 0 do not."""
    import pybmc_b200 as pb
    y, X, prior = CASES["wide_k8"]()
    res = pb.run_gibbs(y, X, 400, prior, n_chains=32, seed=2, stats="full")
    s = res.samples.reshape(32, 400, 9)
    w = s.var(axis=1, ddof=1).mean(axis=0)
    b_over_n = s.mean(axis=1).var(axis=0, ddof=1)
    want = np.sqrt((399.0 / 400.0 * w + b_over_n) / w)
    np.testing.assert_allclose(res.rhat, want, rtol=1e-8)
    assert np.all(np.abs(res.rhat - 1.0) < 0.02)
    ys, Xs, Vt, S = _simplex_case()
    slow = pb.run_gibbs_simplex(ys, Xs, Vt, S, 300, [1.0, 0.02], burn=0, stepsize=0.002, n_chains=32, seed=2)
    assert slow.rhat is not None and slow.rhat[:3].max() > 1.05
    # effective sample size: the conjugate sampler mixes in a few iterations, the random walk does not
    big = pb.run_gibbs(y, X, 400, prior, n_chains=4096, seed=3, stats="full", keep_samples=False)
    assert np.all(big.ess > 0.5 * 4096 * 400) and np.all(big.ess < 1.3 * 4096 * 400)
    assert slow.ess[:3].max() < 0.05 * 32 * 300


def test_thinning_and_discard_select_the_same_iterates():
    import pybmc_b200 as pb
    y, X, prior = CASES["toy_identity_prior"]()
    full = pb.gibbs_sampler(y, X, 60, prior, seed=9)
    thin = pb.gibbs_sampler(y, X, 60, prior, seed=9, thin=7, discard=4)
    assert np.array_equal(thin, full[4::7])


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 2e-4)])
def test_device_moments_match_stored_samples(dtype, tol):
    """fp64 moment sums accumulated in the kernel == moments of the samples it wrote."""
    import pybmc_b200 as pb
    for name, stats in (("toy_dense_prior", "full"), ("wide_k8", "full"), ("wide_k11", "full"), ("wide_k20", "diag")):
        y, X, prior = CASES[name]()
        width = np.asarray(X).shape[1] + 1
        res = pb.run_gibbs(y, X, 300, prior, n_chains=64, seed=21, dtype=dtype, stats=stats)
        s = res.samples
        scale = np.abs(s).max()
        np.testing.assert_allclose(res.mean, s.mean(axis=0), rtol=tol, atol=tol * scale)
        want = np.cov(s.T, ddof=0)
        if stats == "diag":      # only the marginal variances are accumulated (in the rotated coordinates)
            continue
        np.testing.assert_allclose(res.cov, want, rtol=10 * tol, atol=10 * tol * np.abs(want).max())
        per_chain = s.reshape(64, 300, width).mean(axis=1)
        np.testing.assert_allclose(res.chain_mean, per_chain, rtol=tol, atol=tol * scale)


def test_literal_kernel_matches_oracle_and_sufficient_statistic_law():
    import pybmc_b200 as pb
    r = _ens()
    prior = [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02]
    T, seed = 80, 77
    got = pb.gibbs_sampler_literal(r["y"], r["U_hat"], T, prior, n_chains=2, seed=seed).reshape(2, T, 4)

    def precision_cholesky(cov):          # b = mean + L^-T z with A = inv(cov) = L L'
        return np.linalg.inv(np.linalg.cholesky(np.linalg.inv(cov))).T
    for c in range(2):
        ref = oc.gibbs_conjugate(r["y"], r["U_hat"], T, prior,
                                 oc.PhiloxDraws(seed, c, px.TAG_GIBBS, precision_cholesky))
        np.testing.assert_allclose(got[c], ref, rtol=5e-8, atol=1e-10)
    # non-orthonormal toy design, dense prior
    y, X, prior = CASES["toy_dense_prior"]()
    got = pb.gibbs_sampler_literal(y, X, 50, prior, n_chains=1, seed=4)
    ref = oc.gibbs_conjugate(y, np.asarray(X, float), 50, prior, oc.PhiloxDraws(4, 0, px.TAG_GIBBS, precision_cholesky))
    np.testing.assert_allclose(got, ref, rtol=5e-8, atol=1e-10)


def _replicated_reference(fn, reps, base_seed):
    out = []
    for r in range(reps):
        np.random.seed(base_seed + r)
        out.append(fn(oc.NumpyDraws(cases.SeededFactory(10_000 * (base_seed + r)))))
    return out


def _assert_within_mcse(gpu_value, rep_values, gpu_se, label, k=3.0):
    rep_values = np.asarray(rep_values)
    centre = rep_values.mean(axis=0)
    se = rep_values.std(axis=0, ddof=1) / np.sqrt(len(rep_values))
    z = np.abs(gpu_value - centre) / np.sqrt(se ** 2 + gpu_se ** 2)
    assert np.all(z < k), f"{label}: z = {z}"


def test_conjugate_posterior_matches_reference_sampler_statistically():
    import pybmc_b200 as pb
    r = _ens()
    prior = [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02]
    T_ref, reps = 1500, 40
    runs = _replicated_reference(lambda d: oc.gibbs_conjugate(r["y"], r["U_hat"], T_ref, prior, d), reps, 100)
    for dtype in ("float64", "float32"):
        res = pb.run_gibbs(r["y"], r["U_hat"], 400, prior, n_chains=2048, seed=1, dtype=dtype, keep_samples=False)
        n_eff = 400 * 2048
        sd = np.sqrt(np.diag(res.cov))
        _assert_within_mcse(res.mean, [s.mean(axis=0) for s in runs], sd / np.sqrt(n_eff), f"mean {dtype}")
        _assert_within_mcse(sd, [s.std(axis=0) for s in runs], sd / np.sqrt(2 * n_eff), f"sd {dtype}")
        corr = res.cov / np.outer(sd, sd)
        ref_corr = [np.corrcoef(s.T) for s in runs]
        _assert_within_mcse(corr[np.triu_indices(4, 1)], [c[np.triu_indices(4, 1)] for c in ref_corr],
                            1.0 / np.sqrt(n_eff), f"corr {dtype}")
        # analytic anchor for default priors: E[b_k] ~ c_k / (1 + s2 / S_k^2)
        c = r["U_hat"].T @ r["y"]
        s2 = res.mean[-1] ** 2
        assert np.allclose(res.mean[:3], c / (1 + s2 / r["S_hat"] ** 2), rtol=2e-3, atol=2e-3)


# ---------------------------------------------------------------------------------------------
def _simplex_case():
    r = _ens()
    return r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"]


def test_simplex_chain_matches_oracle_fp64(capsys):
    import pybmc_b200 as pb
    y, X, Vt, S = _simplex_case()
    burn, T, seed = 120, 200, 0xB202
    res = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=3, seed=seed)
    got = res.samples.reshape(3, T, 4)
    for c in range(3):
        ref, acc = oc.gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02,
                                    draws=oc.PhiloxDraws(seed, c, px.TAG_SIMPLEX, lambda cov: np.sqrt(cov)),
                                    return_acceptance=True)
        np.testing.assert_allclose(got[c], ref, rtol=1e-8, atol=1e-11)
        assert round(res.acceptance[c] * T) == acc
        assert 0 < acc < T
    # toy case of the upstream tests (non-orthonormal X, weights on the boundary)
    y, X, Vt, S = cases.toy_simplex()
    got = pb.gibbs_sampler_simplex(y, X, Vt, S, 100, [1.0, 1.0], burn=10, stepsize=0.01, seed=6)
    assert "Acceptance rate:" in capsys.readouterr().out
    ref = oc.gibbs_simplex(y, X, Vt, S, 100, [1.0, 1.0], burn=10, stepsize=0.01,
                           draws=oc.PhiloxDraws(6, 0, px.TAG_SIMPLEX, lambda cov: np.sqrt(cov)))
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-11)


def test_simplex_chain_matches_oracle_k8_m16():
    import pybmc_b200 as pb
    rng = np.random.default_rng(12)
    preds, truth = cases.ensemble(77, 120, 16)
    r = oc.orthogonalize_arrays(preds, truth, 8)
    burn, T, seed = 60, 120, 31
    res = pb.run_gibbs_simplex(r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"], T, [1.0, 0.02], burn=burn,
                               stepsize=0.01, n_chains=2, seed=seed)
    got = res.samples.reshape(2, T, 9)
    for c in range(2):
        ref, acc = oc.gibbs_simplex(r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"], T, [1.0, 0.02], burn=burn,
                                    stepsize=0.01, draws=oc.PhiloxDraws(seed, c, px.TAG_SIMPLEX, lambda cov: np.sqrt(cov)),
                                    return_acceptance=True)
        np.testing.assert_allclose(got[c], ref, rtol=1e-8, atol=1e-11)
        assert round(res.acceptance[c] * T) == acc


def test_simplex_layouts_agree_with_oracle_and_each_other():
    """Few chains run eight lanes per chain, many chains one chain per thread: same stream, same law.
    Chain ids are global, so chain 5 of a 16384-chain launch is the oracle's chain 5."""
    import pybmc_b200 as pb
    y, X, Vt, S = _simplex_case()
    burn, T, seed = 40, 60, 77
    many = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=16384, seed=seed,
                                stats="full")
    few = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=8, seed=seed,
                               stats="full")
    got_many = many.samples.reshape(16384, T, 4)
    got_few = few.samples.reshape(8, T, 4)
    np.testing.assert_allclose(got_many[:8], got_few, rtol=1e-9, atol=1e-12)
    for c in (0, 5):
        ref, acc = oc.gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02,
                                    draws=oc.PhiloxDraws(seed, c, px.TAG_SIMPLEX, lambda cov: np.sqrt(cov)),
                                    return_acceptance=True)
        np.testing.assert_allclose(got_many[c], ref, rtol=1e-8, atol=1e-11)
        assert round(many.acceptance[c] * T) == acc == round(few.acceptance[c] * T)
    # device moment sums of both layouts against the samples they wrote
    for res in (few, many):
        smp = res.samples
        np.testing.assert_allclose(res.mean, smp.mean(axis=0), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(res.cov, np.cov(smp.T, ddof=0), rtol=1e-7, atol=1e-12)
    diag = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=8, seed=seed,
                                stats="diag")
    np.testing.assert_allclose(np.diag(diag.cov), np.diag(few.cov), rtol=1e-10)
    np.testing.assert_allclose(diag.mean, few.mean, rtol=1e-12)


def test_simplex_validation_errors():
    import pybmc_b200 as pb
    y, X, Vt, S = cases.toy_simplex()
    with pytest.raises(ValueError):
        pb.gibbs_sampler_simplex(y, X, Vt, S, 10, [1.0, 1.0], burn=-1)
    with pytest.raises(ValueError):
        pb.gibbs_sampler_simplex(y, X, Vt, S, 10, [1.0, 1.0], stepsize=-0.01)


def test_simplex_posterior_matches_reference_sampler_statistically():
    """Slow-mixing chain: Monte-Carlo errors come from replicated reference chains, not a formula."""
    import pybmc_b200 as pb
    y, X, Vt, S = _simplex_case()
    burn, T, reps = 400, 1500, 96
    runs = _replicated_reference(
        lambda d: oc.gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, draws=d), reps, 300)
    for dtype in ("float64", "float32"):
        res = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=1024, seed=2,
                                   dtype=dtype, keep_samples=False)
        # the device estimate averages 1024 chains of the same length: its error is the replicate spread / 32
        rep_means = np.array([s.mean(axis=0) for s in runs])
        gpu_se = rep_means.std(axis=0, ddof=1) / np.sqrt(1024)
        _assert_within_mcse(res.mean, rep_means, gpu_se, f"simplex mean {dtype}")
        rep_sd = np.array([s.std(axis=0) for s in runs])
        chain_sd_se = rep_sd.std(axis=0, ddof=1) / np.sqrt(1024)
        # pooled sd over chains includes between-chain spread of the means; compare within-chain sd instead
        within = np.sqrt(np.maximum(np.diag(res.cov) - res.chain_mean.var(axis=0), 0.0))
        _assert_within_mcse(within, rep_sd, chain_sd_se + 0.02 * rep_sd.mean(axis=0), f"simplex sd {dtype}")
        w = res.mean[:3] @ Vt + 1.0 / Vt.shape[1]
        assert np.all(w > -1e-6) and abs(w.sum() - 1.0) < 1e-8
        acc_ref = np.mean([np.mean(np.any(np.diff(s[:, :3], axis=0) != 0, axis=1)) for s in runs])
        assert abs(res.acceptance.mean() - acc_ref) < 0.03


def test_simplex_few_model_kernels_match_general_group_kernel():
    """At most 16 models: the kernel that precomputes the weight changes of 32 iterations at a time (the
    default) walks the same chains as the general eight-lanes-per-chain kernel -- running sums instead of
    per-iteration dot products, so fp64 round-off apart -- with the same moment sums and acceptance counts;
    chain counts that leave groups of the last warp idle, burn-in that ends inside a batch, thinning."""
    import pybmc_b200 as pb
    y, X, Vt, S = _simplex_case()
    layouts = {1: "warp", 0: "group"}          # bmc_simplex_problem.layout: precomputed rows / general group kernel
    for n_chains, T, burn, thin in ((37, 500, 70, 3), (1, 203, 0, 1), (64, 61, 11, 7)):
        out = {}
        for mode in (1, 0):
            out[mode] = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02,
                                             n_chains=n_chains, seed=5, thin=thin, stats="full",
                                             layout=layouts[mode])
        b = out[0]
        for mode in (1,):
            a = out[mode]
            np.testing.assert_allclose(a.samples, b.samples, rtol=1e-9, atol=1e-12)
            assert np.array_equal(a.acceptance, b.acceptance)
            np.testing.assert_allclose(a.mean, b.mean, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(a.cov, b.cov, rtol=1e-7, atol=1e-12)
    f32 = {}
    for mode in (1, 0):
        f32[mode] = pb.run_gibbs_simplex(y, X, Vt, S, 4000, [1.0, 0.02], burn=500, stepsize=0.02, n_chains=512,
                                         seed=9, dtype="float32", keep_samples=False, stats="full",
                                         layout=layouts[mode])
    # fp32: same law (chains may part ways after a borderline decision); moments of 2e6 draws agree
    np.testing.assert_allclose(f32[1].mean, f32[0].mean, rtol=2e-3, atol=2e-4)
    assert abs(f32[1].acceptance.mean() - f32[0].acceptance.mean()) < 5e-3


def test_conjugate_layouts_agree_in_fp32():
    """The fp32 thread-per-chain kernel (packed FFMA2 arithmetic, iterations in pairs, segmented loop, packed
    moment pairs) against the fp32 group kernels (scalar arithmetic) on the same stream, and its device moment
    sums -- every entry of the packed layout -- against the samples it wrote; kept draws at odd and even
    iterations, a keep point inside a pair, flush boundaries inside the run."""
    import pybmc_b200 as pb
    for name in ("ensemble_default", "wide_k8"):
        y, X, prior = CASES[name]()
        width = np.asarray(X).shape[1] + 1
        T = 150
        many = pb.run_gibbs(y, X, T, prior, n_chains=16384, seed=13, stats="full", dtype="float32")
        few = pb.run_gibbs(y, X, T, prior, n_chains=6, seed=13, stats="full", dtype="float32")
        mid = pb.run_gibbs(y, X, T, prior, n_chains=2000, seed=13, stats="full", dtype="float32")
        a = many.samples.reshape(16384, T, width)[:6].astype(np.float64)
        b = few.samples.reshape(6, T, width).astype(np.float64)
        c = mid.samples.reshape(2000, T, width)[:6].astype(np.float64)
        scale = np.abs(b).max(axis=(0, 1))
        assert np.max(np.abs(a - b) / scale) < 2e-5 and np.max(np.abs(c - b) / scale) < 2e-5
        smp = many.samples.astype(np.float64)
        np.testing.assert_allclose(many.mean, smp.mean(axis=0), rtol=2e-5, atol=2e-6 * scale.max())
        want = np.cov(smp.T, ddof=0)
        assert np.max(np.abs(many.cov - want)) < 2e-3 * np.abs(want).max()
        thin = pb.run_gibbs(y, X, T, prior, n_chains=16384, seed=13, stats="diag", dtype="float32", thin=7, discard=5)
        kept = thin.samples.reshape(16384, -1, width)
        full = many.samples.reshape(16384, T, width)
        assert np.array_equal(kept, full[:, 5::7])
        np.testing.assert_allclose(thin.mean, many.mean, rtol=1e-6, atol=1e-7 * scale.max())
