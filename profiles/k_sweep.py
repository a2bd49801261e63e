"""Sampler throughput against the number of components and the moment mode (65,536 chains, fp32)."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import pybmc_b200 as pb
from pybmc_b200.inference_utils import ConjugateSampler
rng = np.random.default_rng(0)
for k in (3, 8, 12, 16, 24, 32, 64):
    n = 3000
    X = np.linalg.qr(rng.normal(size=(n, k)))[0]
    y = X @ rng.normal(size=k) * 3 + rng.normal(0, 0.2, n)
    s = ConjugateSampler(y, X, [np.zeros(k), np.eye(k) * 100.0, 1.0, 0.02])
    row = [k]
    for mode in ("none", "diag", "full"):
        if mode == "full" and k > 16:
            row.append(float("nan")); continue
        its = 1000
        s.run(its, 65536, 1, "float32", its, 0, False, mode, 0); torch.cuda.synchronize()
        t0 = time.perf_counter(); s.run(its, 65536, 1, "float32", its, 0, False, mode, 0); torch.cuda.synchronize()
        row.append(65536 * its / (time.perf_counter() - t0))
    print("K=%2d  none %.2e  diag %.2e  full %.2e chain-iters/s" % tuple(row))
