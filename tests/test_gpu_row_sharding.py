"""Rows of the prediction table split over ranks (SURVEY.md 8e, configs[4]): the device path with its
all-reduce emulated on one GPU.  Each "rank" is run in turn; a replay reducer feeds it the sum of all
ranks' partials (recorded in earlier sweeps), which is exactly what NCCL would hand it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Replay:
    """reduce callable: call i returns totals[i] when known, else the local tensor; records locals."""

    def __init__(self, totals):
        self.totals, self.local, self.i = totals, [], 0

    def __call__(self, t):
        self.local.append(t.clone())
        out = self.totals[self.i] if self.i < len(self.totals) else t
        self.i += 1
        return out.clone()


def _emulate(run_rank, n_ranks, n_calls):
    """Sweep until every reduce call has its total: call c's partials are valid once calls < c were replayed."""
    totals, results = [], None
    for sweep in range(n_calls + 1):
        reducers = [Replay(totals) for _ in range(n_ranks)]
        results = []
        for r in range(n_ranks):
            try:
                results.append(run_rank(r, reducers[r]))
            except Exception:                       # a lone shard may be singular before its totals are known
                if sweep == n_calls:
                    raise
                results.append(None)
        if sweep < n_calls:
            totals = totals + [sum(red.local[sweep] for red in reducers)]
    return results


def _table(n=3000, m=40, seed=5):
    rng = np.random.default_rng(seed)
    base = rng.uniform(100, 2000, n)
    latent = rng.normal(size=(n, 6)) * np.logspace(0, -1, 6)
    preds = base[:, None] + 20 * latent @ rng.normal(size=(6, m)) + 0.1 * rng.normal(size=(n, m))
    truth = base + 2 * latent[:, 0] + rng.normal(0, 0.2, n)
    return preds, truth


def test_row_sharded_orthogonalize_and_sampler_match_single_rank():
    import pybmc_b200 as pb
    from pybmc_b200.inference_utils import ConjugateSampler
    preds, truth = _table()
    k, cuts = 6, [0, 1100, 1101, 3000]                      # three uneven "ranks", one holding a single row
    whole = pb.orthogonalize_arrays(preds, truth, k, method="gram")

    def orth(r, reduce):
        return pb.orthogonalize_arrays(preds[cuts[r]:cuts[r + 1]], truth[cuts[r]:cuts[r + 1]], k, reduce=reduce)
    parts = _emulate(orth, 3, 1)
    for key in ("S_hat",):
        np.testing.assert_allclose(parts[0][key], whole[key], rtol=1e-12)
    sign = np.sign(np.sum(parts[0]["Vt_hat"] * whole["Vt_hat"], axis=1))
    for p in parts:                                            # replicated results are identical on every rank
        assert np.array_equal(p["Vt_hat"], parts[0]["Vt_hat"]) and np.array_equal(p["S_hat"], parts[0]["S_hat"])
    np.testing.assert_allclose(parts[0]["Vt_hat"] * sign[:, None], whole["Vt_hat"], rtol=1e-7, atol=1e-12)
    u = np.concatenate([p["U_hat"] for p in parts])
    np.testing.assert_allclose(u * sign[None, :], whole["U_hat"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(np.concatenate([p["y"] for p in parts]), whole["y"], rtol=0, atol=0)
    np.testing.assert_allclose(u.T @ u, np.eye(k), atol=1e-9)

    prior = [np.zeros(k), np.diag(whole["S_hat"] ** 2), 1.0, 0.02]
    single = ConjugateSampler(whole["y"], whole["U_hat"], prior)

    def setup(r, reduce):
        lo, hi = cuts[r], cuts[r + 1]
        return ConjugateSampler(whole["y"][lo:hi], whole["U_hat"][lo:hi], prior, reduce=reduce)
    ranks = _emulate(setup, 3, 3)                              # n, Gram, RSS
    for s in ranks:
        assert s.n == single.n
        np.testing.assert_allclose(s.gram, single.gram, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(s.xty, single.xty, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(s.rss_min, single.rss_min, rtol=1e-11)
        np.testing.assert_allclose(s.d, single.d, rtol=1e-10)
        np.testing.assert_allclose(s.pull, single.pull, rtol=1e-8, atol=1e-12)
    # same constants -> same chains, whichever rank runs them (global chain ids)
    a, _, _ = ranks[0].run(20, n_chains=4, seed=3, dtype="float64", stats="none", chain_offset=8)
    b, _, _ = ranks[2].run(20, n_chains=4, seed=3, dtype="float64", stats="none", chain_offset=8)
    c, _, _ = single.run(20, n_chains=4, seed=3, dtype="float64", stats="none", chain_offset=8)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a.cpu().numpy(), c.cpu().numpy(), rtol=1e-7, atol=1e-10)
