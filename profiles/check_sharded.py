"""N-rank check of pybmc_b200.parallel on GPUs: sharded results == single-GPU results.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/check_sharded.py
Prints "SHARDED OK" on rank 0 when every comparison holds (tests/test_gpu_multi.py runs it when >= 2 GPUs are visible)."""
import os
import sys
import warnings
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pybmc_b200 as pb
from pybmc_b200 import parallel as par
from pybmc_b200.sampling_utils import PredictiveProblem

rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
checks = {}

# ---- chains sharded: moments, histograms, kept samples ---------------------------------------------------------------
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
prior = [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02]
mean, cov, loc = par.sharded_gibbs(o["y"], o["U_hat"], 256, prior, 4096 + 3, seed=5, dtype="float64", keep_samples=True,
                                   hist_every=64)
one = pb.run_gibbs(o["y"], o["U_hat"], 256, prior, n_chains=4096 + 3, seed=5, dtype="float64", stats="full", hist_every=64)
lo, hi = loc.info["chain_range"]
checks["gibbs moments"] = np.allclose(mean, one.mean, rtol=1e-12) and np.allclose(cov, one.cov, rtol=1e-9, atol=1e-18)
checks["gibbs samples"] = np.array_equal(loc.samples, one.samples.reshape(4096 + 3, 256, 9)[lo:hi].reshape(-1, 9))
checks["gibbs histograms"] = np.array_equal(loc.hist, one.hist) and np.allclose(loc.quantiles([2.5, 50, 97.5]),
                                                                                one.quantiles([2.5, 50, 97.5]))
try:        # fewer chains than ranks: every rank raises, nobody hangs in a collective
    par.sharded_gibbs(o["y"], o["U_hat"], 10, prior, world - 1, seed=5)
    checks["empty share raises"] = False
except ValueError:
    checks["empty share raises"] = True

# ---- simplex chains sharded: moment sums, count and accepted proposals in one all-reduce ------------------------------
ms_, cs_, acc_, loc_ = par.sharded_gibbs_simplex(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], 400, [1.0, 0.02], 2048, burn=200,
                                                 stepsize=0.02, seed=8, dtype="float64", keep_samples=True, thin=100)
one_s = pb.run_gibbs_simplex(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], 400, [1.0, 0.02], burn=200, stepsize=0.02,
                             n_chains=2048, seed=8, dtype="float64", thin=100)
clo, chi = par.chain_range(2048)
checks["simplex: moments"] = np.allclose(ms_, one_s.mean, rtol=1e-10, atol=1e-12) and np.allclose(cs_, one_s.cov, rtol=1e-8, atol=1e-14)
checks["simplex: acceptance"] = (abs(acc_ - one_s.acceptance.mean()) < 1e-12
                                 and np.array_equal(loc_.acceptance, one_s.acceptance[clo:chi]))
checks["simplex: kept samples"] = np.array_equal(loc_.samples, one_s.samples.reshape(2048, -1, o["U_hat"].shape[1] + 1)[clo:chi]
                                                 .reshape(-1, o["U_hat"].shape[1] + 1))
# ---- nuclei sharded: broadcast of the draws, packed all-gather ---------------------------------------------------------
theta = pb.run_gibbs(o["y"], o["U_hat"], 4000, prior, n_chains=1, seed=6).samples
n_pts = 1003
ref = PredictiveProblem(preds[:n_pts], theta, o["Vt_hat"], truth=truth[:n_pts], dtype="float64").run(seed=9)
# (a) only rank 0's values count: broadcast; (b) every rank holds the rows: 1/N uploaded each, all-gather
for name, kw, th_in in (("predict gathered (broadcast)", dict(draws_from=0), theta if rank == 0 else np.zeros_like(theta)),
                        ("predict gathered (split upload)", dict(), theta)):
    full = par.sharded_predictive_summary(preds[:n_pts], th_in, o["Vt_hat"], truth=truth[:n_pts], seed=9,
                                          dtype="float64", **kw)
    checks[name] = (np.array_equal(full.percentiles, ref.percentiles) and np.array_equal(full.c_lt, ref.c_lt)
                    and np.array_equal(full.c_le, ref.c_le) and np.allclose(full.mean, ref.mean, rtol=1e-13)
                    and np.allclose(full.var, ref.var, rtol=1e-12))
# results gathered to rank 0 only: the same values there, nothing on the other ranks
rooted = par.sharded_predictive_summary(preds[:n_pts], theta, o["Vt_hat"], truth=truth[:n_pts], seed=9, dtype="float64",
                                        gather="root")
checks["predict gathered to rank 0"] = (
    (np.array_equal(rooted.percentiles, ref.percentiles) and np.array_equal(rooted.c_lt, ref.c_lt)
     and np.allclose(rooted.mean, ref.mean, rtol=1e-13)) if rank == 0 else (rooted.mean is None and rooted.percentiles is None))
plo, phi = par.point_range(n_pts)
mine = par.sharded_predictive_summary(preds[plo:phi], theta, o["Vt_hat"], truth=truth[plo:phi], seed=9, dtype="float32",
                                      gather=False, n_points_total=n_pts)
ref32 = PredictiveProblem(preds[:n_pts], theta, o["Vt_hat"], truth=truth[:n_pts], dtype="float32").run(seed=9)
checks["predict local block"] = (np.array_equal(mine.percentiles, ref32.percentiles[:, plo:phi])
                                 and np.array_equal(mine.c_lt, ref32.c_lt[plo:phi]))

# ---- rows sharded: Gram all-reduce, TSQR for a graded spectrum, sampler on all-reduced statistics ---------------------
n_rows = preds.shape[0]
rlo, rhi = par.row_range(n_rows)
so = par.sharded_orthogonalize(preds[rlo:rhi], truth[rlo:rhi], 8)
sign = np.sign(np.sum(so["Vt_hat"] * o["Vt_hat"], axis=1))
checks["rows: orthogonalize"] = (np.allclose(so["S_hat"], o["S_hat"], rtol=1e-10)
                                 and np.allclose(so["U_hat"] * sign, o["U_hat"][rlo:rhi], rtol=0, atol=1e-9)
                                 and so["method"] in ("gram", "tsqr"))
pri = [np.zeros(8), np.diag(so["S_hat"] ** 2), 1.0, 0.02]
m2, c2, _ = par.sharded_gibbs(so["y"], so["U_hat"], 200, pri, 2048, seed=7, dtype="float64", rows_sharded=True)
one2 = pb.run_gibbs(o["y"], o["U_hat"] * sign, 200, pri, n_chains=2048, seed=7, dtype="float64", stats="full")
checks["rows: sampler"] = np.allclose(m2, one2.mean, rtol=1e-8, atol=1e-10) and np.allclose(c2, one2.cov, rtol=1e-6, atol=1e-12)
# graded spectrum (lambda_K / lambda_1 far below 1e-5): the TSQR route must give what the thin SVD gives
rng = np.random.default_rng(3)
graded = rng.normal(size=(600, 12)) @ np.diag(np.logspace(0, -6, 12)) @ rng.normal(size=(12, 12)) + 500.0
gt = graded.mean(axis=1) + rng.normal(size=600)
glo, ghi = par.row_range(600)
with warnings.catch_warnings():
    warnings.simplefilter("error")
    sg = par.sharded_orthogonalize(graded[glo:ghi], gt[glo:ghi], 10)
og = pb.orthogonalize_arrays(graded, gt, 10)
checks["rows: tsqr"] = sg["method"] == "tsqr" and np.allclose(sg["S_hat"], og["S_hat"], rtol=1e-9) and og["method"] == "svd"

# ---- device= names a GPU that is not the current one (one rank is enough) -------------------------------------------
if torch.cuda.device_count() >= 2 and rank == 0:
    other = torch.device("cuda", (local + 1) % torch.cuda.device_count())
    r_other = pb.run_gibbs(o["y"], o["U_hat"], 64, prior, n_chains=40, seed=5, device=other)
    r_here = pb.run_gibbs(o["y"], o["U_hat"], 64, prior, n_chains=40, seed=5)
    p_other = pb.predictive_summary(preds[:200], theta, o["Vt_hat"], truth=truth[:200], seed=3, subsample=False, device=other)
    p_here = pb.predictive_summary(preds[:200], theta, o["Vt_hat"], truth=truth[:200], seed=3, subsample=False)
    checks["device guard"] = (np.array_equal(r_other.samples, r_here.samples) and torch.cuda.current_device() == local
                              and np.array_equal(p_other.percentiles, p_here.percentiles))

flags = torch.tensor([int(all(checks.values()))], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
for k, v in checks.items():
    if not v or rank == 0:
        print(f"[rank {rank}] {k}: {'ok' if v else 'MISMATCH'}", flush=True)
if rank == 0:
    print("SHARDED OK" if int(flags.item()) else "SHARDED MISMATCH", flush=True)
dist.destroy_process_group()
