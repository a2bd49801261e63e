"""Time the headline launch alone: conjugate sampler, BASELINE configs[2] (16 models x 3000 points, K = 8),
65,536 chains x 10,000 iterations, full cross moments, 10 kept draws per chain.  usage: python profiles/time_gibbs.py [dtype] [iterations] [chains] [hist_every] [plain]"""
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
import os
import bench
import pybmc_b200 as pb
from pybmc_b200 import _lib
if os.environ.get("BMC_LIB"):          # A/B runs: a differently built library (profiles/ab/*.so, not tracked)
    _lib.LIB_PATH = os.path.abspath(os.environ["BMC_LIB"])
from pybmc_b200.inference_utils import ConjugateSampler

dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
hist = int(sys.argv[4]) if len(sys.argv) > 4 else bench.HIST_EVERY
persistent = (sys.argv[5] != "plain") if len(sys.argv) > 5 else True
preds, truth = bench.config3_ensemble()
orth = pb.orthogonalize_arrays(preds, truth, 8)
prior = [np.zeros(8), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
s = ConjugateSampler(orth["y"], orth["U_hat"], prior)
ms = []
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    samples, cstats, meta = s.run(iters, chains, 0xB203, dtype, iters // 10, 0, True, "full", 0, None, hist, persistent)
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
mean, cov, _ = s.summarise(cstats, meta, iters, chains)
print(f"{dtype} {chains} chains x {iters} iterations, hist_every {hist}{'' if persistent else ' (plain launch)'}: ms per launch {[round(m, 3) for m in ms]}; {chains * iters / (min(ms[1:]) * 1e-3):.4g} chain-iters/s")
print("mean", np.array2string(mean, precision=6), "sd", np.array2string(np.sqrt(np.diag(cov)), precision=6))
