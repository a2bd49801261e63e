"""Orthogonalisation parity on the GPU (deterministic: 1e-10 relative after sign alignment)."""
import numpy as np
import pytest

import cases
from oracle import bmc_oracle as oc

pytestmark = pytest.mark.gpu


def _aligned(got, ref):
    """Singular vectors are defined up to a sign per component."""
    sign = np.sign(np.sum(got["Vt_hat_normalized"] * ref["Vt_hat_normalized"], axis=1))
    return dict(U_hat=got["U_hat"] * sign[None, :], Vt_hat=got["Vt_hat"] * sign[:, None],
                Vt_hat_normalized=got["Vt_hat_normalized"] * sign[:, None], S_hat=got["S_hat"], y=got["y"],
                mu=got["mu"])


def _check(got, ref, tol=1e-10):
    a = _aligned(got, ref)
    for key in ("y", "mu", "S_hat"):
        np.testing.assert_allclose(a[key], ref[key], rtol=tol, atol=0)
    for key in ("U_hat", "Vt_hat", "Vt_hat_normalized"):
        scale = np.abs(ref[key]).max()
        assert np.max(np.abs(a[key] - ref[key])) <= tol * scale, key
    # sign-free invariants
    np.testing.assert_allclose(got["U_hat"] @ got["U_hat"].T, ref["U_hat"] @ ref["U_hat"].T, atol=1e-11)


@pytest.mark.parametrize("method", ["auto", "gram", "svd"])
def test_golden_cases(golden, method):
    import pybmc_b200 as pb
    df = cases.toy_frame().iloc[:4]
    ref = golden("orth_toy")
    ref = dict(y=ref["y"], mu=ref["mu"], S_hat=ref["S_hat"], U_hat=ref["U_hat"], Vt_hat=ref["Vt_hat"],
               Vt_hat_normalized=ref["Vt_norm"])
    got = pb.orthogonalize_arrays(df[["model1", "model2", "model3"]].values, df["truth"].values, 2, method=method)
    _check(got, ref)
    frame, models = cases.ensemble_frame(11, 40, 5)
    g = golden("orth_ens")
    ref = dict(y=g["y"], mu=g["mu"], S_hat=g["S_hat"], U_hat=g["U_hat"], Vt_hat=g["Vt_hat"],
               Vt_hat_normalized=g["Vt_norm"])
    got = pb.orthogonalize_arrays(frame.iloc[:28][models].values, frame.iloc[:28]["truth"].values, 3, method=method)
    _check(got, ref)


def test_config3_shape_and_graded_spectrum():
    """3000 x 16, K = 8 (BASELINE config 3) and a spectrum graded over 1e6: the Gram route is only
    taken where it can deliver 1e-10; otherwise the thin SVD is."""
    import pybmc_b200 as pb
    rng = np.random.default_rng(1003)
    t = np.cumsum(rng.uniform(5, 15, 3000))
    preds = t[:, None] * (1 + rng.normal(0, 0.003, 16))[None, :] + rng.normal(0, 2, 16)[None, :] \
        + rng.normal(0, 0.5, (3000, 16))
    truth = t + rng.normal(0, 0.15, 3000)
    ref = oc.orthogonalize_arrays(preds, truth, 8, full_matrices=False)
    got = pb.orthogonalize_arrays(preds, truth, 8)
    _check(got, ref, tol=1e-9)
    np.testing.assert_allclose(got["U_hat"].T @ got["U_hat"], np.eye(8), atol=1e-11)
    np.testing.assert_allclose(preds @ got["Vt_hat"].T, got["U_hat"], atol=1e-9)   # Vt_hat projects raw predictions
    # graded: singular values 1 .. 1e-6
    q1 = np.linalg.qr(rng.normal(size=(500, 12)))[0]
    q2 = np.linalg.qr(rng.normal(size=(12, 12)))[0]
    graded = (q1 * np.logspace(0, -6, 12)) @ q2.T
    graded = graded - graded.mean(axis=1, keepdims=True) + 5.0
    ref = oc.orthogonalize_arrays(graded, np.zeros(500), 6, full_matrices=False)
    got = pb.orthogonalize_arrays(graded, np.zeros(500), 6)
    assert got["method"] == "svd"
    np.testing.assert_allclose(got["S_hat"], ref["S_hat"], rtol=1e-9)


def test_device_sufficient_statistics():
    """bmc_gram / bmc_residual_ss against NumPy (X'X, X'y, y'y, RSS_min), non-orthonormal X."""
    from pybmc_b200.inference_utils import ConjugateSampler
    rng = np.random.default_rng(8)
    X = rng.normal(size=(777, 5)) @ rng.normal(size=(5, 5))
    y = X @ rng.normal(size=5) + rng.normal(size=777)
    s = ConjugateSampler(y, X, (np.zeros(5), np.eye(5) * 3.0, 1.0, 0.5))
    np.testing.assert_allclose(s.gram, X.T @ X, rtol=1e-12)
    np.testing.assert_allclose(s.xty, X.T @ y, rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(s.yty, y @ y, rtol=1e-12)
    b = np.linalg.solve(X.T @ X, X.T @ y)
    np.testing.assert_allclose(s.rss_min, np.sum((y - X @ b) ** 2), rtol=1e-11)
    lam = np.linalg.inv(np.eye(5) * 3.0)
    for s2 in (1e-6, 0.02, 3.7):                                   # the reference's covariance, :41
        ref = np.linalg.inv(X.T @ X / s2 + lam + np.eye(5) * 1e-6)
        np.testing.assert_allclose((s.w / (s.d / s2 + 1.0)) @ s.w.T, ref, rtol=1e-10)


def test_distance_split_matches_reference(golden):
    """Dataset.separate_points_distance_allSets (pybmc/data.py:194-245) on the device: same three index
    lists as the reference produced, and as the oracle gives on non-integer coordinates in 3-D."""
    import pandas as pd
    from pybmc_b200.data_utils import separate_points_distance_allSets, split_inside_to_outside
    g = golden("split_distance")
    pts = [tuple(p) for p in g["points"]]
    stable = [tuple(p) for p in g["stable"]]
    tr, va, te = separate_points_distance_allSets(pts, stable, 3.0, 7.5)
    assert tr == list(g["train"]) and va == list(g["val"]) and te == list(g["test"])
    rng = np.random.default_rng(4)
    p3, r3 = rng.normal(size=(5000, 3)) * 4, rng.normal(size=(1500, 3))
    assert separate_points_distance_allSets(p3, r3, 0.8, 2.5) == tuple(oc.distance_classes(p3, r3, 0.8, 2.5))
    assert separate_points_distance_allSets([], stable, 1.0, 2.0) == ([], [], [])
    frame = pd.DataFrame(g["points"], columns=["N", "Z"])
    a, b, c = split_inside_to_outside(frame, stable, 3.0, 7.5)
    assert list(a.index) == list(g["train"]) and list(b.index) == list(g["val"]) and list(c.index) == list(g["test"])
