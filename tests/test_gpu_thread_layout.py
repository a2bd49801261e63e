"""Parity of the kernels bench.py measures: ONE CHAIN PER THREAD (>= 16,384 chains, or ``layout="thread"``).

The headline number comes from ``gibbs_conjugate_kernel<float, 8, 2>`` (packed FFMA2 arithmetic, iterations
in pairs, ``.approx.ftz`` MUFU forms) and its fp64 sibling; with fewer than 16,384 chains the dispatcher
picks the eight-lane / whole-warp kernels instead, so tests with a few thousand chains never touched them.
Here the thread-per-chain kernels are checked on the benchmark's own problem (BASELINE configs[2]: 3000
points, K = 8) -- value by value against the oracle on the shared Philox stream (fp64), and within three
Monte-Carlo standard errors of replicated runs of the oracle on NumPy's generators (= the reference sampler,
pybmc/inference_utils.py:39-54; fp32 and fp64) -- plus the simplex sampler's thread-per-chain kernel
(pybmc/inference_utils.py:97-141) and the fp32 Gamma draw at configs[4]'s shape (n = 1e5 -> shape 5e4).
"""
import ctypes as C

import numpy as np
import pytest

import bench
import cases
from oracle import bmc_oracle as oc
from oracle import philox as px

pytestmark = pytest.mark.gpu


def _config2_problem():
    preds, truth = bench.config3_ensemble()
    r = oc.orthogonalize_arrays(preds, truth, 8, full_matrices=False)
    prior = [np.zeros(8), np.diag(r["S_hat"] ** 2), 1.0, 0.02]
    return r["y"], np.ascontiguousarray(r["U_hat"]), prior


def _w_factor(w):
    w_inv = np.linalg.inv(w)

    def factor(cov):
        return w * np.sqrt(np.diag(w_inv @ cov @ w_inv.T))[None, :]
    return factor


@pytest.mark.parametrize("n_chains,layout", [(16384, None), (300, "thread")])
def test_thread_per_chain_fp64_matches_oracle_value_by_value(n_chains, layout):
    """gibbs_conjugate_kernel<double, 8, *>: first, last and a middle chain of the launch against the oracle's
    literal algorithm (per-iteration inverse, X'y, residual over 3000 rows) on the same Philox stream."""
    from pybmc_b200.inference_utils import ConjugateSampler, _finish_samples
    y, X, prior = _config2_problem()
    sampler = ConjugateSampler(y, X, prior)
    T, seed = 150, 0xB203
    samples, cstats, meta = sampler.run(T, n_chains=n_chains, seed=seed, dtype="float64", stats="full", layout=layout)
    got = samples.permute(2, 0, 1).cpu().numpy()                # [chain, iteration, K+1]
    for c in (0, n_chains // 2 + 1, n_chains - 1):
        ref = oc.gibbs_conjugate(y, X, T, prior, oc.PhiloxDraws(seed, c, px.TAG_GIBBS, _w_factor(sampler.w)))
        np.testing.assert_allclose(got[c], ref, rtol=2e-9, atol=1e-12)
    # device moment sums of this layout against the samples it wrote
    mean, cov, _ = sampler.summarise(cstats, meta, T, n_chains)
    flat = got.reshape(-1, 9)
    np.testing.assert_allclose(mean, flat.mean(axis=0), rtol=1e-9, atol=1e-12)
    want = np.cov(flat.T, ddof=0)
    np.testing.assert_allclose(cov, want, rtol=1e-6, atol=1e-9 * np.abs(want).max())


def test_forced_layouts_walk_the_same_chains():
    """layout is a field of bmc_gibbs_problem: thread / group / warp kernels on the same (seed, chain ids)."""
    import pybmc_b200 as pb
    y, X, prior = _config2_problem()
    out = {lay: pb.run_gibbs(y, X, 70, prior, n_chains=40, seed=5, layout=lay, stats="full")
           for lay in ("thread", "group", "warp", None)}
    for lay in ("group", "warp", None):
        np.testing.assert_allclose(out[lay].samples, out["thread"].samples, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out[lay].mean, out["thread"].mean, rtol=1e-10)
    f32 = {lay: pb.run_gibbs(y, X, 70, prior, n_chains=40, seed=5, layout=lay, dtype="float32").samples
           for lay in ("thread", "group")}
    scale = np.abs(f32["group"]).max(axis=0)
    assert np.max(np.abs(f32["thread"].astype(np.float64) - f32["group"]) / scale) < 2e-5
    with pytest.raises(ValueError):
        yk, Xk, pk = np.zeros(30), np.eye(30)[:, :12], [np.zeros(12), np.eye(12), 1.0, 1.0]
        pb.run_gibbs(yk + 1.0, Xk, 5, pk, layout="group")       # one lane per component: k <= 8


def _replicated_reference(fn, reps, base_seed):
    out = []
    for r in range(reps):
        np.random.seed(base_seed + r)
        out.append(fn(oc.NumpyDraws(cases.SeededFactory(10_000 * (base_seed + r)))))
    return out


def _z_scores(gpu_value, rep_values, gpu_se):
    rep_values = np.asarray(rep_values)
    centre = rep_values.mean(axis=0)
    se = rep_values.std(axis=0, ddof=1) / np.sqrt(len(rep_values))
    return np.abs(gpu_value - centre) / np.sqrt(se ** 2 + gpu_se ** 2)


def test_headline_kernels_match_reference_sampler_statistically():
    """65,536 chains on the configs[2] data -- exactly bench.py's launch shape -- fp32 and fp64: posterior
    mean, sd and correlations of (b, sigma) within 3 MCSE of 40 replicated runs of the oracle on NumPy's own
    generators (the reference sampler by the bit-exact golden pin)."""
    import pybmc_b200 as pb
    y, X, prior = _config2_problem()
    T_ref, reps = 1500, 40
    runs = _replicated_reference(lambda d: oc.gibbs_conjugate(y, X, T_ref, prior, d), reps, 700)
    iu = np.triu_indices(9, 1)
    for dtype in ("float32", "float64"):
        res = pb.run_gibbs(y, X, T_ref, prior, n_chains=bench.CHAINS_PER_GPU, seed=11, dtype=dtype,
                           keep_samples=False, stats="full")
        n_eff = T_ref * bench.CHAINS_PER_GPU
        sd = np.sqrt(np.diag(res.cov))
        z = _z_scores(res.mean, [s.mean(axis=0) for s in runs], sd / np.sqrt(n_eff))
        assert np.all(z < 3.0), f"mean {dtype}: z = {z}"
        z = _z_scores(sd, [s.std(axis=0) for s in runs], sd / np.sqrt(2 * n_eff))
        assert np.all(z < 3.0), f"sd {dtype}: z = {z}"
        corr = res.cov / np.outer(sd, sd)
        z = _z_scores(corr[iu], [np.corrcoef(s.T)[iu] for s in runs], 1.0 / np.sqrt(n_eff))
        assert np.sum(z > 3.0) <= 1 and np.all(z < 4.0), f"corr {dtype}: z = {z}"      # 36 entries
        # R-hat of 65,536 chains and the analytic anchor of SURVEY.md section 7
        assert np.all(np.abs(res.rhat - 1.0) < 5e-3)
        c = X.T @ y
        s2 = res.mean[-1] ** 2
        s_hat = np.sqrt(np.diag(prior[1]))
        assert np.allclose(res.mean[:8], c / (1 + s2 / s_hat ** 2), rtol=1e-3, atol=1e-3 * np.abs(c).max())


def test_simplex_thread_per_chain_matches_reference_sampler_statistically():
    """gibbs_simplex_kernel (one chain per thread, what >= 16,384 simplex chains run on), fp32 and fp64, against
    replicated oracle chains: means within 3 replicate standard errors, within-chain sd and acceptance rate."""
    import pybmc_b200 as pb
    frame, models = cases.ensemble_frame(11, 40, 5)
    tr = frame.iloc[:28]
    r = oc.orthogonalize_arrays(tr[models].values, tr["truth"].values, 3)
    y, X, Vt, S = r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"]
    burn, T, reps, chains = 400, 1500, 96, 4096
    runs = _replicated_reference(
        lambda d: oc.gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, draws=d), reps, 300)
    rep_means = np.array([s.mean(axis=0) for s in runs])
    rep_var = np.array([s.var(axis=0) for s in runs])
    acc_ref = np.array([np.mean(np.any(np.diff(s[:, :3], axis=0) != 0, axis=1)) for s in runs])
    for dtype in ("float64", "float32"):
        res = pb.run_gibbs_simplex(y, X, Vt, S, T, [1.0, 0.02], burn=burn, stepsize=0.02, n_chains=chains, seed=3,
                                   dtype=dtype, keep_samples=False, layout="thread")
        gpu_se = rep_means.std(axis=0, ddof=1) / np.sqrt(chains)
        z = _z_scores(res.mean, rep_means, gpu_se)
        assert np.all(z < 3.0), f"simplex mean {dtype}: z = {z}"
        # like with like: pooled variance = mean within-chain variance + variance of the chain means (all ddof = 0),
        # so the mean within-chain VARIANCE of 4096 device chains is compared with that of the 96 replicates
        within_var = np.diag(res.cov) - res.chain_mean.var(axis=0)
        z = _z_scores(within_var, rep_var, rep_var.std(axis=0, ddof=1) / np.sqrt(chains))
        assert np.all(z < 3.0), f"simplex within-chain variance {dtype}: z = {z}"
        se_acc = acc_ref.std(ddof=1) / np.sqrt(reps)
        assert abs(res.acceptance.mean() - acc_ref.mean()) < 3.0 * se_acc + 1.0 / T, dtype


def _gamma_only_run(dtype, shape, n_chains, iters, seed, layout, k=1):
    """A conjugate-sampler problem whose RSS does not depend on the coefficient (d = 0): sigma^2 of iteration t is
    exactly scale / Gamma_t, so the kept sigmas give back the device's Gamma(shape, 1) variates.  ``k`` picks the
    variate layout: padded k = 8 takes the Gamma words from the iteration's own Philox calls, the others from the
    Gamma block."""
    import torch
    from pybmc_b200 import _device as D
    from pybmc_b200 import _lib
    lib = _lib.load()
    dev = D.device(None)
    tdt, code = D.resolve_dtype(dtype)
    nu0, sigma20, rss = 1.0, 0.02, 777.0
    consts = torch.tensor([0.0] * (3 * k) + [1.0] * k, dtype=torch.float64, device=dev)   # d, pull, g_ols, w
    base = consts.data_ptr()
    prob = _lib.GibbsProblem(k=k, d=base, pull=base + 8 * k, g_ols=base + 16 * k, w=base + 24 * k, dense_w=0,
                             rss_min=rss, n_obs=2.0 * shape - nu0, nu0=nu0, sigma20=sigma20,
                             sigma2_init=rss / (2.0 * shape), layout=_lib.LAYOUTS[layout])
    out = torch.empty((iters, k + 1, n_chains), dtype=tdt, device=dev)
    _lib.check(lib.bmc_gibbs_run(code, C.byref(prob), seed, 0, n_chains, iters, 0, 1, iters, out.data_ptr(), None,
                                 _lib.STATS_NONE, None, D.stream_ptr(dev)))
    sig = out[:, k, :].double().cpu().numpy()                   # [iteration, chain]
    return 0.5 * (nu0 * sigma20 + rss) / sig ** 2


@pytest.mark.parametrize("k", [1, 8])
@pytest.mark.parametrize("shape", [5.0e4, 1500.5, 14.5])
def test_gamma_variates_at_large_shape(shape, k):
    """Marsaglia-Tsang on the device at shape (nu0 + n)/2 = 5e4 (configs[4]: n = 1e5), where the fp32 acceptance
    test 0.5 x^2 + d (1 - v + log v) cancels 30 times harder than at configs[2]'s 1500.5: value by value against
    oracle/philox.py::gamma_unit_scale (fp64 <= 1e-9; fp32 to fp32 rounding of sigma), and the moments of 2e6
    fp32 variates against Gamma(shape, 1) within 4 standard errors."""
    seed, key = 99, px.seed_key(99)
    g64 = _gamma_only_run("float64", shape, 64, 40, seed, "thread", k)
    want = np.array([[px.gamma_unit_scale(shape, it, c, px.TAG_GIBBS, key, k) for c in range(64)] for it in range(40)])
    if k == 8:      # the eight-lane layout reads the same words through the stand-alone path
        np.testing.assert_allclose(_gamma_only_run("float64", shape, 64, 40, seed, "group", k), want, rtol=1e-9)
    np.testing.assert_allclose(g64, want, rtol=1e-9)
    g32 = _gamma_only_run("float32", shape, 64, 40, seed, "thread", k)
    # the variate is recovered from a stored fp32 sigma (2^-24 relative, doubled by the square) on top of the
    # fp32 evaluation of d (1 + c x)^3; a different accept/reject decision would show as an O(1/sqrt(shape)) jump
    assert np.max(np.abs(g32 / want - 1.0)) < 3e-6
    big = _gamma_only_run("float32", shape, 32768, 64, seed + 1, None, k)
    n = big.size
    assert abs(big.mean() - shape) < 4.0 * np.sqrt(shape / n) + 2e-7 * shape
    assert abs(big.var() / shape - 1.0) < 4.0 * np.sqrt(2.0 / n) + 6.0 / shape      # kurtosis term 6/shape
    # skewness 2 / sqrt(shape); its standard error from the spread over the 64 iterations (independent batches)
    per_it = np.array([np.mean((b - shape) ** 3) / shape ** 1.5 for b in big])
    assert abs(per_it.mean() - 2.0 / np.sqrt(shape)) < 4.0 * per_it.std(ddof=1) / np.sqrt(len(per_it))


def test_histogram_quantiles_match_reference_sampler():
    """Marginal histograms collected on the device (bmc_gibbs_hist): the 2.5 / 16 / 50 / 84 / 97.5 % points of every
    coordinate of (b, sigma) against the same percentiles of replicated oracle runs, within 3 replicate standard
    errors plus one bin width; counts are exact integers and independent of layout, sharding and arithmetic
    of the other statistics."""
    import pybmc_b200 as pb
    y, X, prior = _config2_problem()
    q = [2.5, 16.0, 50.0, 84.0, 97.5]
    T_ref, reps = 1500, 40
    runs = _replicated_reference(lambda d: oc.gibbs_conjugate(y, X, T_ref, prior, d), reps, 900)
    rep_q = np.array([np.percentile(s, q, axis=0) for s in runs])                     # [reps, Q, 9]
    for dtype, chains, layout in (("float32", 65536, None), ("float64", 4096, None)):
        res = pb.run_gibbs(y, X, 1024, prior, n_chains=chains, seed=21, dtype=dtype, keep_samples=False,
                           hist_every=64, layout=layout)
        assert res.hist.shape == (9, 512) and res.hist.dtype == np.int64
        assert np.all(res.hist.sum(axis=1) == chains * (1024 // 64))
        got = res.quantiles(q)                                                        # [Q, 9]
        assert np.all(np.isfinite(got))
        se = rep_q.std(axis=0, ddof=1) / np.sqrt(reps)
        assert np.all(np.abs(got - rep_q.mean(axis=0)) < 3.0 * se + res.hist_width[None, :]), dtype
        # the central 68 % interval against mean +- sd of the moment sums (near-Gaussian posterior)
        sd = np.sqrt(np.diag(res.cov))
        np.testing.assert_allclose(0.5 * (got[3] - got[1]), sd, rtol=0.03)
    # the binned states are the iterates 63, 127, ... (hist_every = 64) -- stored here with discard=63, thin=64 --
    # and the histogram of those stored samples is reproduced bit for bit, in every layout
    for layout, chains, every in (("thread", 70, 64), ("group", 70, 64), ("warp", 5, 64), ("thread", 16384, 64),
                                  ("thread", 33, 128), ("group", 9, 192)):
        res = pb.run_gibbs(y, X, 400, prior, n_chains=chains, seed=4, hist_every=every, layout=layout,
                           discard=every - 1, thin=every)
        s = res.samples
        assert s.shape[0] == chains * (400 // every)
        for c in range(9):
            idx = np.clip(np.floor((s[:, c] - res.hist_lo[c]) * (1.0 / res.hist_width[c])), 0, 511).astype(int)
            assert np.array_equal(np.bincount(idx, minlength=512), res.hist[c]), (layout, c)
    with pytest.raises(ValueError):
        pb.run_gibbs(y, X, 100, prior, n_chains=4, seed=4, hist_every=7)      # multiples of 64 only


def test_fp64_normals_from_the_tables_match_the_oracle_to_rounding():
    """The fp64 kernels take log and sine / cosine of the uniforms straight from the random word (integer exponent
    and table index, short series: rng.cuh).  With d = 0 and pull = 0 the sampler's coefficient IS the standard
    normal vector of the iteration, so the kept samples expose the device normals: they must equal
    oracle/philox.py's math.log / math.cos evaluation to a few ulp of the intermediate quantities."""
    import torch
    from pybmc_b200 import _device as D
    from pybmc_b200 import _lib
    lib = _lib.load()
    dev = D.device(None)
    k, chains, iters, seed = 8, 96, 50, 4242
    consts = torch.tensor([0.0] * k + [0.0] * k + [0.0] * k + [1.0] * k, dtype=torch.float64, device=dev)
    base = consts.data_ptr()
    for layout in ("thread", "group"):
        prob = _lib.GibbsProblem(k=k, d=base, pull=base + 8 * k, g_ols=base + 16 * k, w=base + 24 * k, dense_w=0,
                                 rss_min=10.0, n_obs=50.0, nu0=1.0, sigma20=0.02, sigma2_init=0.2,
                                 layout=_lib.LAYOUTS[layout])
        out = torch.empty((iters, k + 1, chains), dtype=torch.float64, device=dev)
        _lib.check(lib.bmc_gibbs_run(_lib.F64, C.byref(prob), seed, 0, chains, iters, 0, 1, iters, out.data_ptr(), None,
                                     _lib.STATS_NONE, None, D.stream_ptr(dev)))
        z = out[:, :k, :].cpu().numpy()                           # [iteration, component, chain]
        key = px.seed_key(seed)
        worst = 0.0
        for c in (0, 31, 95):
            for it in range(iters):
                want = np.array(px.normal_vector(k, it, c, px.TAG_GIBBS, key))
                worst = max(worst, np.max(np.abs(z[it, :, c] - want)))
        assert worst < 2e-14, (layout, worst)
    # tails and the extreme words: u01 of 0 and of 2^32 - 1 through the same code (noise of the prediction path)
    import pybmc_b200 as pb
    theta = np.column_stack([np.zeros((64, 2)), np.ones(64)])      # beta = 0, sigma = 1: the draws are the noise
    preds = np.array([[1.0, -1.0], [2.0, -2.0], [0.5, -0.5], [3.0, -3.0]])
    vt = np.array([[0.5, -0.5], [0.25, -0.25]])
    res = pb.predictive_summary(preds, theta, vt, seed=77, subsample=False, dtype="float64", return_draws=True)
    key = px.seed_key(77)
    want = np.array([[px.noise_block(sb, n, key)[j] for n in range(4)] for sb in range(16) for j in range(4)])
    np.testing.assert_allclose(res.draws, want, rtol=0, atol=2e-14)


@pytest.mark.parametrize("dtype,n_chains", [("float32", 65536), ("float64", 40000), ("float32", 100000)])
def test_persistent_launch_equals_plain_launch(dtype, n_chains):
    """The persistent launch (work items handed round resident warps, gibbs_kernels.cuh) against the plain one warp
    per 32 chains: chains are pure functions of (seed, chain id, iteration), so kept draws, per-chain moment sums
    and histogram counts must be IDENTICAL -- bit for bit -- whichever warp on whichever SM ran which stretch of a
    chain, in both item orders (65,536 chains: list order; 100,000: owned groups + rotating left-overs).  Then one
    chain of the persistent launch against the oracle, value by value, across item boundaries (items are 256
    iterations)."""
    import torch
    from pybmc_b200.inference_utils import ConjugateSampler
    y, X, prior = _config2_problem()
    sampler = ConjugateSampler(y, X, prior)
    T, seed, thin = 1100, 0xB2, 100                                    # 5 items per chain group, the last one short
    out = {}
    for persistent in (True, False):
        samples, cstats, meta = sampler.run(T, n_chains=n_chains, seed=seed, dtype=dtype, thin=thin, stats="full",
                                            layout="thread", hist_every=64, persistent=persistent)
        out[persistent] = (samples.clone(), cstats.clone(), meta["hist"].clone())
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)
    assert int(out[True][2].sum()) == (T // 64) * n_chains * 9        # every chain binned at every flush point
    if dtype == "float64":
        got = out[True][0].permute(2, 0, 1).cpu().numpy()              # [chain, kept, K+1]
        for c in (0, n_chains - 1):
            ref = oc.gibbs_conjugate(y, X, T, prior, oc.PhiloxDraws(seed, c, px.TAG_GIBBS, _w_factor(sampler.w)))
            np.testing.assert_allclose(got[c], ref[::thin], rtol=2e-9, atol=1e-12)
