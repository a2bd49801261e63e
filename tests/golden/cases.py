"""Seeded inputs shared by ``make_golden.py`` (which feeds them to the upstream
reference) and by the tests (which feed them to the oracle and the CUDA path).

Inputs are rebuilt from seeds rather than stored; each golden file also keeps a
checksum of its inputs so a changed NumPy stream fails loudly instead of silently
comparing different problems.
"""
import numpy as np
import pandas as pd


class SeededFactory:
    """Stand-in for ``numpy.random.default_rng`` whose k-th call returns a PCG64
    generator seeded with ``base + k`` (the reference builds a fresh, OS-seeded
    generator on every variance draw, pybmc/inference_utils.py:52,117,140)."""

    def __init__(self, base):
        self.base = int(base)
        self.calls = 0

    def __call__(self, *args, **kwargs):
        g = np.random.Generator(np.random.PCG64(self.base + self.calls))
        self.calls += 1
        return g


def toy_frame():
    """The 6-row frame of the upstream tests (tests/test_bmc.py:10-19)."""
    return pd.DataFrame({
        "x": [1, 2, 3, 4, 5, 6],
        "y": [10, 11, 12, 13, 14, 15],
        "truth": [11, 21, 31, 41, 51, 61],
        "model1": [10, 20, 30, 40, 50, 60],
        "model2": [15, 25, 35, 45, 55, 65],
        "model3": [12, 30, 32, 43, 58, 67],
    })


def toy_regression():
    """tests/test_inference_utils.py:6-14 (non-orthonormal design)."""
    y = np.array([1.0, 2.0, 3.0])
    X = np.array([[1, 0], [0, 1], [1, 1]])
    return y, X


def toy_simplex():
    """tests/test_inference_utils.py:22-27."""
    y, X = toy_regression()
    Vt_hat = np.array([[0.5, 0.5], [0.5, -0.5]])
    S_hat = np.array([1.0, 0.5])
    return y, X, Vt_hat, S_hat


def ensemble(seed, n, n_models, scale=100.0, spread=0.01, offset=0.5, noise=0.2,
             truth_noise=0.1):
    """A small model ensemble: every model is the truth curve with its own scale
    error, offset and white noise.  Returns (preds[n, M], truth[n])."""
    rng = np.random.default_rng(seed)
    t = scale + np.cumsum(rng.uniform(0.5, 1.5, n))
    preds = (t[:, None] * (1 + rng.normal(0, spread, n_models))[None, :]
             + rng.normal(0, offset, n_models)[None, :]
             + rng.normal(0, noise, (n, n_models)))
    truth = t + rng.normal(0, truth_noise, n)
    return preds, truth


def ensemble_frame(seed, n, n_models):
    preds, truth = ensemble(seed, n, n_models)
    cols = {"N": np.arange(n) + 8, "Z": np.arange(n) // 2 + 8}
    for m in range(n_models):
        cols[f"m{m}"] = preds[:, m]
    cols["truth"] = truth
    return pd.DataFrame(cols), [f"m{m}" for m in range(n_models)]


def long_table(seed=41, n_points=60):
    """A CSV-style long table (one row per model and point, pybmc/data.py:86-107): three models that each
    miss a few points, two properties, and a ``truth`` pseudo-model covering everything."""
    rng = np.random.default_rng(seed)
    n = np.arange(n_points) % 12 + 8
    z = np.arange(n_points) // 12 + 20
    base = 8.0 * (n + z) - 0.05 * (n - z) ** 2
    rows = []
    for j, model in enumerate(["mA", "mB", "mC", "truth"]):
        keep = np.ones(n_points, bool) if model == "truth" else rng.random(n_points) > 0.15
        order = rng.permutation(np.flatnonzero(keep))
        rows.append(pd.DataFrame({
            "model": model, "N": n[order], "Z": z[order],
            "BE": np.round(base[order] * (1 + 0.002 * j) + rng.normal(0, 0.3, order.size), 6),
            "Rad": np.round(1.2 * (n[order] + z[order]) ** (1 / 3) + rng.normal(0, 0.01, order.size), 6)}))
    return pd.concat(rows, ignore_index=True)


def posterior_like(seed, rows, k):
    """Rows shaped like sampler output: K coefficients and a positive sigma."""
    rng = np.random.default_rng(seed)
    beta = rng.normal(0.0, 1.0, k)[None, :] + 0.2 * rng.normal(size=(rows, k))
    sig = np.abs(0.3 + 0.03 * rng.normal(size=rows))
    return np.column_stack([beta, sig])


def checksum(*arrays):
    h = 0.0
    for a in arrays:
        a = np.asarray(a, dtype=np.float64).ravel()
        h += float(np.dot(a, np.cos(np.arange(a.size) + 1.0)))
    return h
