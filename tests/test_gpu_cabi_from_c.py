"""The drop-in boundary is a C ABI: a plain-C program (gcc, CUDA runtime only -- no Python, no torch in
the process) calls libbmc_b200.so and its output is checked against the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases
from oracle import bmc_oracle as oc
from oracle import philox as px

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_c_program_against_oracle(tmp_path):
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    libdir = os.path.join(ROOT, "pybmc_b200", "csrc")
    exe = str(tmp_path / "cabi_demo")
    subprocess.run([gcc, "-O1", "-o", exe, os.path.join(ROOT, "tests", "cabi_demo.c"), f"-I{cuda}/include",
                    f"-L{libdir}", "-lbmc_b200", f"-L{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{cuda}/lib64"], check=True)
    # the caller's share of the set-up (host, fp64): OLS, RSS_min and the simultaneous diagonalisation
    y, X = cases.toy_regression()
    X = np.asarray(X, float)
    b0, B0 = np.array([0.0, 0.0]), np.eye(2)
    gram = X.T @ X
    lam = np.linalg.inv(B0)
    w, d = oc.simultaneous_diagonalisation(gram, lam)
    b_ols = np.linalg.solve(gram, X.T @ y)
    g_ols = np.linalg.solve(w, b_ols)
    pull = w.T @ (lam @ b0) - g_ols
    rss_min = float(np.sum((y - X @ b_ols) ** 2))
    s2_init = max(rss_min / 3, 1e-6)
    args = [*d, *pull, *g_ols, *w.reshape(-1), rss_min, s2_init]
    r = subprocess.run([exe, *[repr(float(v)) for v in args]], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("version 1")
    got_gram = np.array(lines[1].split()[1:], dtype=float).reshape(3, 3)
    aug = np.column_stack([X, y])
    np.testing.assert_allclose(got_gram, aug.T @ aug, rtol=1e-15)
    got = np.array([ln.split()[3:] for ln in lines if ln.startswith("sample")], dtype=float).reshape(4, 5, 3)
    w_inv = np.linalg.inv(w)
    for chain in range(4):
        want = oc.gibbs_conjugate(y, X, 5, (b0, B0, 1.0, 1.0), oc.PhiloxDraws(
            42, chain, px.TAG_GIBBS, lambda cov: w * np.sqrt(np.diag(w_inv @ cov @ w_inv.T))[None, :]))
        np.testing.assert_allclose(got[chain], want, rtol=1e-9, atol=1e-12)
    assert [ln for ln in lines if ln.startswith("counts")][0] == "counts 2 2 0 3"
    bad = [ln for ln in lines if ln.startswith("bad call")][0]
    assert "-> -1" in bad and "n_chains=0" in bad
