"""Where the headline launch spends its time: the same 65,536 x 10,000 fp32 launch with the moment sums switched
to marginal-only / off (and, with an A/B library built with -DBMC_PHILOX_ROUNDS=5, with half the Philox rounds:
timing only, the variates are then NOT the contract's).  usage: [BMC_LIB=...] python profiles/sensitivity.py [dtype]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
import bench
import pybmc_b200 as pb
from pybmc_b200 import _lib
if os.environ.get("BMC_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["BMC_LIB"])
from pybmc_b200.inference_utils import ConjugateSampler

dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
iters, chains = 10000, 65536
preds, truth = bench.config3_ensemble()
orth = pb.orthogonalize_arrays(preds, truth, 8)
prior = [np.zeros(8), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
s = ConjugateSampler(orth["y"], orth["U_hat"], prior)
for stats in ("full", "diag", "none"):
    ms = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run(iters, chains, 0xB203, dtype, iters // 10, 0, True, stats, 0, None, 0)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"{os.environ.get('BMC_LIB', 'default')} {dtype} stats={stats}: {min(ms[1:]):.3f} ms")
