import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench, pybmc_b200 as pb
from pybmc_b200.inference_utils import ConjugateSampler, SimplexSampler
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
prior = [np.zeros(8), np.diag(o["S_hat"]**2), 1.0, 0.02]
cs = ConjugateSampler(o["y"], o["U_hat"], prior)
p1, t1 = bench.config1_ensemble(); idx = np.random.default_rng(1).permutation(len(t1))[:377]
o1 = pb.orthogonalize_arrays(p1[idx], t1[idx], 3)
ss = SimplexSampler(o1["y"], o1["U_hat"], o1["Vt_hat"], o1["S_hat"], [1.0, 0.02], 0.001)
def tm(f):
    f(); torch.cuda.synchronize(); t0=time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter()-t0)*1e3
for chains in (1024, 4096, 16384, 32768, 65536, 131072):
    row = [chains]
    for thr in ("0", "100000000"):
        os.environ["BMC_X_CONJ"] = thr; os.environ["BMC_X_SIMPLEX"] = thr
        row.append(round(tm(lambda: cs.run(2000, chains, 1, "float32", 2000, 0, False, "full", 0)), 2))
        row.append(round(tm(lambda: ss.run(4000, 1000, chains, 1, "float32", 4000, False, "full", 0)), 2))
    print("chains %6d | conj thread %7.2f group %7.2f | simplex thread %7.2f group %7.2f ms" % (row[0], row[1], row[3], row[2], row[4]))
