"""Where the end-to-end milliseconds go (host clock with a device synchronisation after every stage).
usage: python profiles/e2e_breakdown2.py [gibbs|predict]"""
import cProfile
import pstats
import sys
import time
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
import bench
import pybmc_b200 as pb
from pybmc_b200 import parallel as par

what = sys.argv[1] if len(sys.argv) > 1 else "gibbs"


def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


if what == "gibbs":
    preds, truth = bench.config3_ensemble()
    o = pb.orthogonalize_arrays(preds, truth, 8)
    y, X = np.ascontiguousarray(o["y"]), np.ascontiguousarray(o["U_hat"])
    prior = [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02]
    fn = lambda: par.sharded_gibbs(y, X, 10000, prior, 65536, seed=1, dtype="float32", thin=1000, keep_samples=True, hist_every=64)
else:
    preds, vt, theta, truth = bench.config4_inputs(100000, 100000, 16)
    fn = lambda: par.sharded_predictive_summary(preds, theta, vt, truth=truth, percentiles=bench.PRED_Q, seed=1, dtype="float32")
for rep in range(4):
    t, _ = T(fn)
    print(f"{what} e2e call {rep}: {t:.2f} ms")
pr = cProfile.Profile()
pr.enable()
fn()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
