"""Conjugate sampler with few chains: eight lanes per chain vs a whole warp per chain (fp64 and fp32)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench, pybmc_b200 as pb
from pybmc_b200.inference_utils import ConjugateSampler
p1, t1 = bench.config1_ensemble(); idx = np.random.default_rng(1).permutation(len(t1))[:377]
o1 = pb.orthogonalize_arrays(p1[idx], t1[idx], 3)
prior = [np.zeros(3), np.diag(o1["S_hat"] ** 2), 1.0, 0.02]
cs = ConjugateSampler(o1["y"], o1["U_hat"], prior)
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
cs8 = ConjugateSampler(o["y"], o["U_hat"], [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02])
def tm(f):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
for name, s, iters in (("K=3", cs, 50000), ("K=8", cs8, 10000)):
    for dtype in ("float64", "float32"):
        for chains in (1, 8, 64, 256, 592, 1024, 2048, 4096, 8192):
            row = []
            for thr in ("0", "100000000"):
                os.environ["BMC_X_CONJ_WARP"] = thr
                row.append(tm(lambda: s.run(iters, chains, 1, dtype, iters, 0, False, "full", 0)))
            print(f"{name} {dtype} chains {chains:5d} | 8 lanes {row[0]:8.2f} ms | warp {row[1]:8.2f} ms", flush=True)
