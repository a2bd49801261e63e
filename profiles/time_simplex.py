"""Time the simplex sampler of BASELINE configs[1] (surrogate 377 x 15, K = 3; 4096 chains x (10,000 burn + 50,000)),
fp32, for the two few-chains kernels (bmc_simplex_problem.layout: 0 general group kernel, 1 precomputed rows).
usage: python profiles/time_simplex.py [n_chains] [modes]"""
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
import bench
import pybmc_b200 as pb
from pybmc_b200 import _lib
from pybmc_b200.inference_utils import SimplexSampler

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
modes = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "01")]
lib = _lib.load()
preds, truth = bench.config1_ensemble()
idx = np.random.default_rng(1).permutation(len(truth))[:377]
o = pb.orthogonalize_arrays(preds[idx], truth[idx], 3)
s = SimplexSampler(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], [1.0, 0.02], 0.001)
for mode in modes:
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        samples, cstats, accepted, meta = s.run(50000, 10000, n_chains, 7, "float32", 5000, True, "full", 0,
                                                "warp" if mode else "group")
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    mean, cov, _ = s.summarise(cstats, meta, 50000, n_chains)
    print(f"mode {mode}: {n_chains} chains: ms {[round(m, 2) for m in ms]}; {n_chains * 60000 / (min(ms[1:]) * 1e-3):.4g} chain-iters/s; "
          f"acceptance {accepted.float().mean().item() / 50000:.4f}; mean {np.array2string(mean, precision=5)}")
