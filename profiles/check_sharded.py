"""2-rank check of pybmc_b200.parallel on GPUs: sharded results == single-GPU results.
torchrun --nproc-per-node 2 profiles/check_sharded.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, pybmc_b200 as pb
from pybmc_b200 import parallel as par
from pybmc_b200.sampling_utils import PredictiveProblem

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
prior = [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02]
mean, cov, local_res = par.sharded_gibbs(o["y"], o["U_hat"], 200, prior, 4096, seed=5, dtype="float64", keep_samples=True)
ok = True
if rank == 0:
    one = pb.run_gibbs(o["y"], o["U_hat"], 200, prior, n_chains=4096, seed=5, dtype="float64", stats="full")
    ok &= np.allclose(mean, one.mean, rtol=1e-12) and np.allclose(cov, one.cov, rtol=1e-9, atol=1e-18)
    ok &= np.array_equal(local_res.samples, one.samples[: len(local_res.samples)])
    print("sharded_gibbs == single GPU:", ok)
theta = pb.run_gibbs(o["y"], o["U_hat"], 4000, prior, n_chains=1, seed=6).samples
res = par.sharded_predictive_summary(preds[:1003], theta, o["Vt_hat"], truth=truth[:1003], seed=9, dtype="float64")
if rank == 0:
    one = PredictiveProblem(preds[:1003], theta, o["Vt_hat"], truth=truth[:1003], dtype="float64").run(seed=9)
    ok2 = (np.array_equal(res.percentiles, one.percentiles) and np.array_equal(res.c_lt, one.c_lt)
           and np.array_equal(res.c_le, one.c_le) and np.allclose(res.mean, one.mean, rtol=1e-13))
    print("sharded_predictive_summary == single GPU:", ok2)
    print("SHARDED OK" if ok and ok2 else "SHARDED MISMATCH")
dist.destroy_process_group()
