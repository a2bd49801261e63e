"""Stage times of the end-to-end sharded prediction (configs[3]) on real ranks: run under torch.distributed.run.
Every stage is closed by a device synchronisation, so the sum exceeds the unsynchronised call; the shares are the point.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/e2e_predict_multi.py"""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, "/root/repo")
import bench
from pybmc_b200 import _device as D
from pybmc_b200 import parallel as par
from pybmc_b200.sampling_utils import PredictiveProblem

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
preds, vt, theta, truth = bench.config4_inputs(100000, 100000, 16)
lo, hi = par.point_range(100000, rank, world)
p_h, t_h = np.ascontiguousarray(preds[lo:hi]), np.ascontiguousarray(truth[lo:hi])


def T(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


def whole():
    return par.sharded_predictive_summary(p_h, theta, vt, truth=t_h, percentiles=bench.PRED_Q, seed=1, dtype="float32",
                                          device=dev, n_points_total=100000)


for rep in range(4):
    t, _ = T(whole)
    if rank == 0:
        print(f"whole call {rep}: {t:.2f} ms", flush=True)
for rep in range(2):
    t_b, th = T(lambda: par.broadcast_draws(theta, 16, device=dev))
    t_s, prob = T(lambda: PredictiveProblem(p_h, th, vt, truth=t_h, dtype="float32", device=dev, point0=lo))
    t_r, res = T(lambda: prob.run(percentiles=bench.PRED_Q, seed=1, as_numpy=False))
    rows = 2 + 5 + 2
    per = par.point_range(100000, 0, world)[1]

    def gather():
        block = torch.zeros((rows, per), dtype=torch.float64, device=dev)
        block[0, : hi - lo] = res.mean
        block[1, : hi - lo] = res.var
        block[2:7, : hi - lo] = res.percentiles
        block[7, : hi - lo] = res.c_lt.view(torch.float64)
        block[8, : hi - lo] = res.c_le.view(torch.float64)
        if world > 1:
            parts = torch.empty((world, rows, per), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(parts, block)
            block = parts.permute(1, 0, 2).reshape(rows, world * per)[:, :100000]
        return block.contiguous()
    t_g, block = T(gather)
    t_h2, _ = T(lambda: D.to_host(block))
    if rank == 0:
        print(f"draws to every GPU {t_b:.2f} | problem set-up (uploads, u = P Vt', moments of the draws) {t_s:.2f} | "
              f"fused kernels {t_r:.2f} | pack + all-gather {t_g:.2f} | to host {t_h2:.2f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
