"""One fused prediction at K = 64 (tensor-core pass) for an ncu capture: 25,088 nuclei x 10,000 draws."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pybmc_b200 import _lib
from pybmc_b200.sampling_utils import PredictiveProblem

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
rng = np.random.default_rng(1005)
n, n_draws = 25088, 10000
pr = rng.uniform(100, 2000, n)[:, None] + rng.normal(0, 3.0, (n, 80))
vt = rng.normal(size=(k, 80)) * 0.02
vt -= vt.mean(axis=1, keepdims=True)
theta = np.column_stack([rng.normal(size=k)[None, :] + 0.1 * rng.normal(size=(n_draws, k)),
                         np.abs(rng.normal(0.15, 0.01, n_draws))])
prob = PredictiveProblem(pr, theta, vt, truth=pr.mean(axis=1), dtype="float32")
for _ in range(2):
    res = prob.run(percentiles=[2.5, 50.0, 97.5], seed=3, as_numpy=False)
torch.cuda.synchronize()
print("ok", res.passes)
