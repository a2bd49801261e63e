"""CPU-side checks: the C-ABI library loads and exports every symbol include/bmc_b200.h declares
(no compute calls), the host-side logic of the package, and that nothing falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pybmc_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bmc_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.bmc_version() >= 100
    assert lib.bmc_last_error() is not None


def test_pure_host_entry_points():
    from pybmc_b200 import _lib
    lib = _lib.load()
    assert [lib.bmc_padded_components(k) for k in (1, 3, 4, 5, 8, 9, 16, 17, 33, 64)] == \
        [4, 4, 4, 8, 8, 16, 16, 32, 64, 64]
    assert lib.bmc_gibbs_n_stat(8, _lib.STATS_NONE) == 0
    assert lib.bmc_gibbs_n_stat(8, _lib.STATS_DIAG) == 18
    assert lib.bmc_gibbs_n_stat(8, _lib.STATS_FULL) == 9 + 45
    assert lib.bmc_gram_workspace_bytes(3000, 17) >= 17 * 17 * 8
    assert lib.bmc_rss_workspace_bytes(3000) >= 8
    small = lib.bmc_predict_workspace_bytes(_lib.F32, 629, 3, 10000)
    big = lib.bmc_predict_workspace_bytes(_lib.F64, 100000, 5, 100000)
    assert 0 < small < big < 8 * 2 ** 30
    assert lib.bmc_predict_workspace_bytes(_lib.F32, 629, 9, 10000) == 0     # more than BMC_MAX_QUANTILES


def test_argument_errors_surface_as_value_errors_without_a_gpu():
    """Validation happens before any CUDA call, and maps to the exceptions the reference raises."""
    from pybmc_b200 import _lib
    lib = _lib.load()
    prob = _lib.SimplexProblem(k=2, m=2, n_obs=3.0, nu0=1.0, sigma20=1.0)
    rc = lib.bmc_gibbs_simplex_run(_lib.F64, ctypes.byref(prob), 0, 0, 1, -1, 10, 1, 10, None, None, 0, None, None)
    assert rc == _lib.ERR_ARG
    assert lib.bmc_last_error().decode() == "Burn-in iterations must be non-negative."
    with pytest.raises(ValueError):
        _lib.check(rc)
    gp = _lib.GibbsProblem(k=65)
    assert lib.bmc_gibbs_run(_lib.F32, ctypes.byref(gp), 0, 0, 1, 1, 0, 1, 1, None, None, 0, None, None) == _lib.ERR_ARG
    assert "k=65" in lib.bmc_last_error().decode()


def test_no_cpu_fallback():
    import torch
    import pybmc_b200 as pb
    from pybmc_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CUDA present: nothing to refuse")
    y, X = cases.toy_regression()
    with pytest.raises(_lib.BmcError):
        pb.gibbs_sampler(y, X, 10, (np.zeros(2), np.eye(2), 1.0, 1.0))
    with pytest.raises(_lib.BmcError):
        pb.coverage([50], np.zeros((10, 2)), __import__("pandas").DataFrame({"t": [0.0, 1.0]}), "t")
    with pytest.raises(_lib.BmcError):
        pb.orthogonalize_arrays(np.ones((4, 3)), np.ones(4), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pybmc_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_usvt_hat_extraction_matches_reference(golden):
    import pybmc_b200 as pb
    g = golden("usvt")
    uh, sh, vh, vn = pb.USVt_hat_extraction(g["U"], g["S"], g["Vt"], 2)
    assert np.array_equal(uh, g["U_hat"]) and np.array_equal(sh, g["S_hat"])
    assert np.array_equal(vh, g["Vt_hat"]) and np.array_equal(vn, g["Vt_norm"])
    # upstream test (tests/test_inference_utils.py:34-45)
    uh, sh, vh, vn = pb.USVt_hat_extraction(np.eye(2), np.array([2.0, 1.0]), np.eye(2), 2)
    assert uh.shape == (2, 2) and len(sh) == 2 and vh.shape == (2, 2) and vn.shape == (2, 2)


def test_coverage_indices_follow_the_reference_expression():
    from pybmc_b200.sampling_utils import coverage_indices
    from oracle import bmc_oracle as oc
    for s in (137, 200, 10000, 100000):
        assert coverage_indices(np.arange(0, 101, 5), s) == oc.coverage_indices(np.arange(0, 101, 5), s)
    lo, hi = coverage_indices(np.arange(0, 101, 5), 10000)
    assert lo[16] == 999 and lo[18] == 499


def test_moment_layout_roundtrip():
    """Host unpacking of the device moment rows (diag and full layouts)."""
    from pybmc_b200 import _lib
    from pybmc_b200.inference_utils import _moments_from_stats
    rng = np.random.default_rng(0)
    k, kp, n = 3, 4, 5000
    e = rng.normal(size=(n, kp + 1)) @ rng.normal(size=(kp + 1, kp + 1))
    e[:, 3] = 0.0                                   # padded component
    d = kp + 1
    full = np.concatenate([e.sum(0), [np.sum(e[:, r] * e[:, c]) for r in range(d) for c in range(r, d)]])
    mean, cov = _moments_from_stats(full, k, kp, _lib.STATS_FULL, n)
    keep = [0, 1, 2, 4]
    np.testing.assert_allclose(mean, e[:, keep].mean(0))
    np.testing.assert_allclose(cov, np.cov(e[:, keep].T, ddof=0), atol=1e-12)
    diag = np.concatenate([e.sum(0), (e ** 2).sum(0)])
    mean, cov = _moments_from_stats(diag, k, kp, _lib.STATS_DIAG, n)
    np.testing.assert_allclose(np.diag(cov), e[:, keep].var(0))
    assert np.count_nonzero(cov - np.diag(np.diag(cov))) == 0


def test_chain_diagnostics_on_ar1_chains():
    """R-hat and the between-chain effective sample size from per-chain moment sums (host logic, torch on
    the CPU): AR(1) chains with coefficient rho have integrated autocorrelation time (1+rho)/(1-rho)."""
    import torch
    from pybmc_b200.inference_utils import _chain_diagnostics, _stat_layout
    rng = np.random.default_rng(5)
    k, kp, chains, n = 2, 4, 600, 4000
    rhos = np.array([0.0, 0.8, 0.5])                      # two components + sigma
    x = np.empty((chains, n, 3))
    x[:, 0] = rng.standard_normal((chains, 3))
    innov = rng.standard_normal((chains, n, 3)) * np.sqrt(1 - rhos ** 2)
    for t in range(1, n):
        x[:, t] = rhos * x[:, t - 1] + innov[:, t]
    d, comp, second = _stat_layout(k, kp, 2)
    stats = np.zeros((d + d * (d + 1) // 2, chains))
    for a, ia in enumerate(comp):
        stats[ia] = x[:, :, a].sum(axis=1)
        for b, ib in enumerate(comp):
            if ib >= ia:
                stats[second[(ia, ib)]] = (x[:, :, a] * x[:, :, b]).sum(axis=1)
    cs = torch.from_numpy(stats)
    chain_mean = torch.from_numpy(x.mean(axis=1))
    rhat, ess = _chain_diagnostics(cs, comp, second, torch.eye(3, dtype=torch.float64), chain_mean, n)
    assert np.all(np.abs(rhat - 1.0) < 5e-3)
    want = chains * n * (1 - rhos) / (1 + rhos)
    assert np.all(np.abs(ess / want - 1.0) < 0.2), (ess, want)        # sd of the estimate ~ sqrt(2/chains) = 6 %
    one = _chain_diagnostics(cs[:, :1], comp, second, torch.eye(3, dtype=torch.float64), chain_mean[:1], n)
    assert one == (None, None)


def test_install_as_pybmc_makes_upstream_imports_resolve_here():
    """Scripts written against upstream (docs/usage.md:10-12) run unchanged after one call; an upstream
    package that is already imported is not shadowed."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import pybmc_b200; pybmc_b200.install_as_pybmc()\n"
        "from pybmc.data import Dataset\n"
        "from pybmc.bmc import BayesianModelCombination\n"
        "from pybmc.inference_utils import gibbs_sampler, gibbs_sampler_simplex, USVt_hat_extraction\n"
        "from pybmc.sampling_utils import coverage, rndm_m_random_calculator\n"
        "import pybmc\n"
        "assert pybmc is pybmc_b200 and pybmc.Dataset is Dataset is pybmc_b200.Dataset\n"
        "assert BayesianModelCombination is pybmc_b200.BayesianModelCombination\n"
        "assert gibbs_sampler is pybmc_b200.gibbs_sampler and coverage is pybmc_b200.coverage\n"
        "import types; sys.modules['pybmc'] = types.ModuleType('pybmc')\n"
        "try:\n"
        "    pybmc_b200.install_as_pybmc()\n"
        "except RuntimeError:\n"
        "    print('refused')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "refused"


def test_predict_before_train_raises_like_upstream():
    """Upstream never initialises ``samples`` / ``Vt_hat`` (pybmc/bmc.py:73-77), so predicting on a fresh object
    raises AttributeError at bmc.py:212 before its ValueError can; the drop-in does the same (no GPU involved)."""
    import pandas as pd
    import pybmc_b200 as pb
    df = pd.DataFrame({"m1": [1.0, 2.0], "m2": [2.0, 3.0], "truth": [1.5, 2.5]})
    bmc = pb.BayesianModelCombination(["m1", "m2"], {"p": df}, "truth")
    with pytest.raises(AttributeError):
        bmc.predict(df)
    with pytest.raises(AttributeError):
        bmc.evaluate()
