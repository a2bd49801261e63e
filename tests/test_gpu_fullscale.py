"""BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle cannot run
them): exactness of the unmaterialised percentiles via a second pass of order counts, invariance to
sharding / chunking, analytic anchors of the posterior."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def test_config3_full_size_sampler():
    """16 models x 3000 points, K = 8, 65,536 chains x 10,000 iterations (fp32 arithmetic)."""
    import bench
    import pybmc_b200 as pb
    from pybmc_b200.inference_utils import ConjugateSampler
    preds, truth = bench.config3_ensemble()
    o = pb.orthogonalize_arrays(preds, truth, 8)
    prior = [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02]
    s = ConjugateSampler(o["y"], o["U_hat"], prior)
    _, cstats, meta = s.run(10000, 65536, 11, "float32", 1, 0, False, "full", 0)
    mean, cov, chain_mean = s.summarise(cstats, meta, 10000, 65536)
    # analytic anchors for the default prior
    c = o["U_hat"].T @ o["y"]
    s2 = mean[-1] ** 2 + cov[-1, -1]
    n_tot = 10000.0 * 65536
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(mean[:8] - c / (1 + s2 / o["S_hat"] ** 2)) < 6 * sd[:8] / np.sqrt(n_tot) + 2e-6 * np.abs(c))
    np.testing.assert_allclose(sd[:8], (1 / s2 + 1 / o["S_hat"] ** 2) ** -0.5, rtol=2e-3)
    rss_min = s.rss_min
    np.testing.assert_allclose(s2, (1.0 * 0.02 + rss_min + 8 * s2) / (1.0 + 3000 - 2), rtol=2e-3)
    assert np.all(np.abs(s.last_rhat - 1.0) < 1e-3)
    # chains are keyed by global id: two half launches sum to the full launch exactly (fp64 sums of the
    # same per-chain rows), i.e. what two GPUs would all-reduce
    _, lo, _ = s.run(10000, 32768, 11, "float32", 1, 0, False, "full", 0)
    lo_sum = lo.sum(dim=1)
    _, hi, _ = s.run(10000, 32768, 11, "float32", 1, 0, False, "full", 32768)
    both = (lo_sum + hi.sum(dim=1)).cpu().numpy()
    full = cstats.sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(both, full, rtol=1e-12, atol=1e-9)
    # the same seed reproduces the run bit for bit (moment sums are accumulated in a fixed order)
    _, again, _ = s.run(10000, 32768, 11, "float32", 1, 0, False, "full", 0)
    assert torch_equal(again, lo)


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))


def test_config4_full_size_prediction_is_exact():
    """1e5 nuclei x 1e5 draws x K = 16 without the S x N matrix.  The returned percentile q of a nucleus
    lies between its order statistics r and r + 1, so a second pass that counts draws below q must
    return exactly r + 1 (and the pair of counts for an un-interpolated percentile brackets r)."""
    import bench
    from pybmc_b200.sampling_utils import PredictiveProblem
    n, s = 100_000, 100_000
    preds, vt, theta, truth = bench.config4_inputs(n, s)
    q = [2.5, 16.0, 50.0, 84.0, 97.5]
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32")
    first = prob.run(percentiles=q, seed=21)
    assert np.all(np.diff(first.percentiles, axis=0) > 0)
    assert np.all(first.c_lt <= first.c_le) and first.c_le.max() <= s
    mu = preds.mean(axis=1)
    u = preds @ vt.T
    np.testing.assert_allclose(first.mean, mu + u @ theta[:, :16].mean(axis=0), atol=5 * np.sqrt(first.var.max() / s))
    import torch
    for j, p in enumerate(q):
        v = p / 100.0 * (s - 1)
        r, frac = int(np.floor(v)), v - np.floor(v)
        # fp32 arithmetic compares centred values: hand the kernel the centred percentile it selected
        centred = first.percentiles[j] - mu
        prob.truth = torch.from_numpy(mu + centred.astype(np.float32).astype(np.float64)).to(prob.dev)
        again = prob.run(percentiles=[50.0], seed=21)
        if frac > 0:
            ok = (again.c_lt >= r) & (again.c_lt <= r + 1) & (again.c_le >= r + 1) & (again.c_le <= r + 2)
        else:
            ok = (again.c_lt <= r) & (again.c_le >= r + 1)
        assert ok.mean() > 0.999, (p, ok.mean())      # fp32 re-centring may move a draw across the threshold
    # the same check in fp64 on the first 8192 nuclei is exact for every nucleus
    p64 = PredictiveProblem(preds[:8192], theta, vt, truth=truth[:8192], dtype="float64")
    f64 = p64.run(percentiles=q, seed=21)
    np.testing.assert_allclose(f64.percentiles, first.percentiles[:, :8192], rtol=1e-5)
    for j, p in enumerate(q):
        v = p / 100.0 * (s - 1)
        r = int(np.floor(v))
        p64.truth = torch.from_numpy(f64.percentiles[j] - mu[:8192]).to(p64.dev) + p64.mu
        again = p64.run(percentiles=[50.0], seed=21)
        # q = mu + lerp(x_r, x_r+1): re-adding mu may round onto x_r or x_r+1 itself, never past them
        assert np.all((again.c_lt >= r) & (again.c_lt <= r + 1) & (again.c_le >= r + 1) & (again.c_le <= r + 2)), p
    # nuclei 40,000 .. 59,999 as their own shard (global ids kept) give the same numbers
    shard = PredictiveProblem(preds[40000:60000], theta, vt, truth=truth[40000:60000], dtype="float32",
                              point0=40000).run(percentiles=q, seed=21)
    assert np.array_equal(shard.percentiles, first.percentiles[:, 40000:60000])
    assert np.array_equal(shard.c_lt, first.c_lt[40000:60000])
    # the sort-free coverage rule on the counts agrees with the percentile band (they can differ only when
    # the truth falls between the two adjacent order statistics at an edge of the band)
    from pybmc_b200.sampling_utils import coverage_indices
    (lo_idx,), (hi_idx,) = coverage_indices([95], s)
    by_counts = (first.c_le >= lo_idx + 1) & (first.c_lt <= hi_idx)
    by_band = (truth >= first.percentiles[0]) & (truth <= first.percentiles[4])
    assert np.mean(by_counts != by_band) < 1e-3
