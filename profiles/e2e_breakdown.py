import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench, pybmc_b200 as pb
from pybmc_b200.inference_utils import ConjugateSampler, _finish_samples
from pybmc_b200 import _device as D
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
y, X = o["y"], o["U_hat"]; prior = [np.zeros(8), np.diag(o["S_hat"]**2), 1.0, 0.02]
def T(f):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); return (time.perf_counter()-t0)*1e3, r
for rep in range(3):
    t_setup, s = T(lambda: ConjugateSampler(y, X, prior))
    t_run, (samples, cstats, meta) = T(lambda: s.run(10000, 65536, 1, "float32", 1000, 0, True, "full", 0))
    t_sum, _ = T(lambda: s.summarise(cstats, meta, 10000, 65536))
    t_fin, arr = T(lambda: _finish_samples(samples, True))
    t_perm, rows = T(lambda: samples.permute(2, 0, 1).reshape(-1, 9).to(torch.float64))
    t_d2h, _ = T(lambda: rows.cpu())
    print(f"setup {t_setup:.2f} run {t_run:.2f} summarise {t_sum:.2f} finish {t_fin:.2f} (permute {t_perm:.2f} d2h {t_d2h:.2f}) ms")

import cProfile, pstats
for rep in range(2):
    t, res = T(lambda: pb.run_gibbs(y, X, 10000, prior, n_chains=65536, seed=1, dtype="float32", thin=1000, stats="full"))
    print(f"run_gibbs total {t:.2f} ms")
pr = cProfile.Profile(); pr.enable()
pb.run_gibbs(y, X, 10000, prior, n_chains=65536, seed=1, dtype="float32", thin=1000, stats="full")
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
