"""The oracle must reproduce the upstream reference bit for bit on the golden
files written by tests/golden/make_golden.py (which ran the unmodified reference
with a seeded ``default_rng`` factory).  This is what pins the oracle."""
import numpy as np
import pandas as pd

import cases
from oracle import bmc_oracle as oc


def seeded_draws(legacy_seed, base):
    np.random.seed(legacy_seed)
    return oc.NumpyDraws(cases.SeededFactory(base))


def test_truncate_svd(golden):
    g = golden("usvt")
    uh, sh, vh, vn = oc.truncate_svd(g["U"], g["S"], g["Vt"], 2)
    assert np.array_equal(uh, g["U_hat"]) and np.array_equal(sh, g["S_hat"])
    assert np.array_equal(vh, g["Vt_hat"]) and np.array_equal(vn, g["Vt_norm"])


def test_orthogonalize_toy(golden):
    g = golden("orth_toy")
    df = cases.toy_frame().iloc[:4]
    r = oc.orthogonalize_arrays(df[["model1", "model2", "model3"]].values, df["truth"].values, 2)
    for key, name in [("y", "y"), ("U_hat", "U_hat"), ("S_hat", "S_hat"), ("Vt_hat", "Vt_hat"),
                      ("Vt_hat_normalized", "Vt_norm"), ("mu", "mu")]:
        assert np.array_equal(r[key], g[name]), key


def _ens():
    frame, models = cases.ensemble_frame(11, 40, 5)
    tr = frame.iloc[:28]
    return frame, models, oc.orthogonalize_arrays(tr[models].values, tr["truth"].values, 3)


def test_orthogonalize_ensemble(golden):
    g = golden("orth_ens")
    frame, models, r = _ens()
    assert cases.checksum(frame[models].values, frame["truth"].values) == float(g["insum"])
    for key, name in [("y", "y"), ("U_hat", "U_hat"), ("S_hat", "S_hat"), ("Vt_hat", "Vt_hat"),
                      ("Vt_hat_normalized", "Vt_norm"), ("mu", "mu")]:
        assert np.array_equal(r[key], g[name]), key
    # thin SVD keeps the same columns (what the large configs use)
    thin = oc.orthogonalize_arrays(frame.iloc[:28][models].values, frame.iloc[:28]["truth"].values, 3,
                                   full_matrices=False)
    assert np.allclose(thin["U_hat"], r["U_hat"], rtol=0, atol=1e-13)


def test_gibbs_conjugate_toy(golden):
    y, X = cases.toy_regression()
    s = oc.gibbs_conjugate(y, X, 60, (np.array([0.0, 0.0]), np.eye(2), 1.0, 1.0), seeded_draws(7, 1000))
    assert np.array_equal(s, golden("gibbs_toy")["samples"])
    s = oc.gibbs_conjugate(y, X, 40, (np.array([0.3, -0.2]), np.array([[2.0, 0.3], [0.3, 0.5]]), 2.5, 0.7),
                           seeded_draws(17, 1100))
    assert np.array_equal(s, golden("gibbs_toy_dense_prior")["samples"])


def test_gibbs_conjugate_ensemble(golden):
    _, _, r = _ens()
    s = oc.gibbs_conjugate(r["y"], r["U_hat"], 300, [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02],
                           seeded_draws(9, 1200))
    assert np.array_equal(s, golden("gibbs_ens")["samples"])


def test_gibbs_simplex(golden):
    y, X, Vt_hat, S_hat = cases.toy_simplex()
    s, acc = oc.gibbs_simplex(y, X, Vt_hat, S_hat, 40, [1.0, 1.0], burn=100, stepsize=0.01,
                              draws=seeded_draws(8, 2000), return_acceptance=True)
    g = golden("simplex_toy")
    assert np.array_equal(s, g["samples"])
    assert abs(acc / 40 * 100 - float(g["acceptance_pct"])) < 0.006
    _, _, r = _ens()
    s, acc = oc.gibbs_simplex(r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"], 250, [1.0, 0.02], burn=300,
                              stepsize=0.02, draws=seeded_draws(10, 2100), return_acceptance=True)
    g = golden("simplex_ens")
    assert np.array_equal(s, g["samples"])
    assert abs(acc / 250 * 100 - float(g["acceptance_pct"])) < 0.006
    assert 0 < acc < 250   # the case exercises both accepted and rejected moves


def test_simplex_validation():
    import pytest
    y, X, Vt_hat, S_hat = cases.toy_simplex()
    with pytest.raises(ValueError):
        oc.gibbs_simplex(y, X, Vt_hat, S_hat, 10, [1.0, 1.0], burn=-1)
    with pytest.raises(ValueError):
        oc.gibbs_simplex(y, X, Vt_hat, S_hat, 10, [1.0, 1.0], stepsize=-0.01)


def test_predictive_and_coverage(golden):
    g = golden("predict")
    preds, truth = cases.ensemble(12, 7, 5)
    theta = cases.posterior_like(13, 12000, 3)
    assert cases.checksum(preds, truth, theta) == float(g["insum"])
    rng = cases.SeededFactory(3000)()
    rndm_m, (lo, med, hi) = oc.predictive_draws(preds, theta, g["Vt_hat"], rng)
    assert np.array_equal(rndm_m[:64], g["head"])
    assert np.array_equal(rndm_m.sum(axis=0), g["colsum"])
    assert cases.checksum(rndm_m) == float(g["total"])
    assert np.array_equal(lo, g["lo"]) and np.array_equal(med, g["med"]) and np.array_equal(hi, g["hi"])
    levels = np.arange(0, 101, 5)
    cov = oc.coverage_levels(levels, rndm_m, truth)
    assert cov == list(g["coverage"])
    # sort-free form used on the device: two integer counts per point
    c_lt, c_le = oc.order_counts(rndm_m, truth)
    assert oc.coverage_from_counts(levels, len(rndm_m), c_lt, c_le) == cov


def test_coverage_ties_and_truncation(golden):
    g = golden("coverage_ties")
    levels = np.arange(0, 101, 5)
    assert oc.coverage_levels(levels, g["matrix"], g["truth"]) == list(g["coverage"])
    assert oc.coverage_levels([1, 33, 68, 95, 99], g["matrix"][:137], g["truth"]) == list(g["coverage_odd"])
    c_lt, c_le = oc.order_counts(g["matrix"], g["truth"])
    assert oc.coverage_from_counts(levels, 200, c_lt, c_le) == list(g["coverage"])
    c_lt, c_le = oc.order_counts(g["matrix"][:137], g["truth"])
    assert oc.coverage_from_counts([1, 33, 68, 95, 99], 137, c_lt, c_le) == list(g["coverage_odd"])
    # the uneven truncation SURVEY.md quotes for S = 10^4 and 10^5
    lo, hi = oc.coverage_indices(levels, 10000)
    assert lo[16] == 999 and lo[18] == 499 and lo[0] == 5000 and hi[0] == 4999
    lo, hi = oc.coverage_indices(levels, 100000)
    assert hi[3] == 57498 and lo[11] == 22499 and lo[16] == 9999 and lo[18] == 4999


def test_simultaneous_diagonalisation_matches_reference_covariance():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(30, 4)); gram = a.T @ a
    b = rng.normal(size=(4, 4)); lam = np.linalg.inv(b @ b.T + np.eye(4))
    w, d = oc.simultaneous_diagonalisation(gram, lam)
    for s2 in (1e-6, 0.02, 3.7):
        ref = np.linalg.inv(gram / s2 + lam + np.eye(4) * oc.RIDGE)   # inference_utils.py:41
        mine = (w / (d / s2 + 1.0)) @ w.T
        assert np.allclose(mine, ref, rtol=1e-11, atol=0)


def test_distance_split(golden):
    g = golden("split_distance")
    tr, va, te = oc.distance_classes(g["points"], g["stable"], 3.0, 7.5)
    assert tr == list(g["train"]) and va == list(g["val"]) and te == list(g["test"])
    assert len(tr) and len(va) and len(te) and sorted(tr + va + te) == list(range(300))
