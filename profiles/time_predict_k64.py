"""configs[4] prediction (1e5 nuclei x 1e4 draws, K = 64, fp32): tensor-core contraction vs FFMA."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from pybmc_b200 import _lib
from pybmc_b200.sampling_utils import PredictiveProblem

lib = _lib.load()
rng = np.random.default_rng(1005)
for k, n, n_draws in [(16, 100000, 100000), (64, 100000, 10000), (16, 629, 10000), (64, 100000, 100000)]:
    pr = rng.uniform(100, 2000, n)[:, None] + rng.normal(0, 3.0, (n, 80))
    vt = rng.normal(size=(k, 80)) * 0.02
    vt -= vt.mean(axis=1, keepdims=True)
    theta = np.column_stack([rng.normal(size=k)[None, :] + 0.1 * rng.normal(size=(n_draws, k)),
                             np.abs(rng.normal(0.15, 0.01, n_draws))])
    prob = PredictiveProblem(pr, theta, vt, truth=pr.mean(axis=1), dtype="float32")
    ws = torch.empty(int(lib.bmc_predict_workspace_bytes(_lib.F32, n, 3, n_draws)), dtype=torch.uint8, device="cuda")
    for mode in (1, 0):
        prob.tensor_min_k = 0 if mode else -1
        times = []
        for it in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = prob.run(percentiles=[2.5, 50.0, 97.5], seed=3, as_numpy=False, workspace=ws)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        ms = float(np.median(times[2:]))
        print(f"K={k} n={n} S={n_draws} tensor={mode}: {ms:.2f} ms  {n * n_draws / ms / 1e6:.1f} G units/s  passes={res.passes}",
              flush=True)
