"""Time the K = 64 conjugate sampler of BASELINE configs[4] alone (thread-per-chain, marginal moments):
16,384 chains x 1,000 iterations on a 4,000 x 64 orthonormal design.  usage: python profiles/time_gibbs_k64.py [dtype] [iterations] [chains]"""
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
from pybmc_b200.inference_utils import ConjugateSampler

dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
rng = np.random.default_rng(64)
q, _ = np.linalg.qr(rng.normal(size=(4000, 64)))
s_hat = np.logspace(2, 0, 64)
y = q @ (rng.normal(size=64) * s_hat) + 0.15 * rng.normal(size=4000)
prior = [np.zeros(64), np.diag(s_hat ** 2), 1.0, 0.02]
s = ConjugateSampler(y, q, prior)
ms = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    samples, cstats, meta = s.run(iters, chains, 0xB205, dtype, max(1, iters // 10), 0, True, "diag", 0)
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
mean, cov, _ = s.summarise(cstats, meta, iters, chains)
print(f"{dtype} K=64 {chains} chains x {iters} iterations: ms {[round(m, 3) for m in ms]}; {chains * iters / (min(ms[1:]) * 1e-3):.4g} chain-iters/s")
print("sigma mean", mean[-1], "b[:3]", mean[:3])
