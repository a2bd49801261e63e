"""Stage times of one sharded sampler step (configs[2], what bench.py times as `value` at N GPUs) on real ranks.
Every stage is closed by a device synchronisation; per-rank lines, so that waiting for a slower rank shows up in the
first collective of the rank that waited.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/e2e_gibbs_multi.py [dtype]"""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, "/root/repo")
import bench
import pybmc_b200 as pb
from pybmc_b200 import parallel as par
from pybmc_b200.inference_utils import ConjugateSampler

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
preds, truth = bench.config3_ensemble()
orth = pb.orthogonalize_arrays(preds, truth, 8, device=dev)
prior = [np.zeros(8), np.diag(orth["S_hat"] ** 2), 1.0, 0.02]
s = ConjugateSampler(orth["y"], orth["U_hat"], prior, device=dev)
C, IT = bench.CHAINS_PER_GPU, bench.ITERATIONS


def now():
    torch.cuda.synchronize()
    return time.perf_counter()


def whole():
    return par.sharded_gibbs(None, None, IT, prior, C * world, seed=1, dtype=dtype, thin=IT // 10, keep_samples=True,
                             hist_every=bench.HIST_EVERY, as_numpy=False, sampler=s, device=dev)


for rep in range(5):
    if world > 1:
        dist.barrier()
    t0 = now()
    whole()
    t1 = now()
    if rank == 0:
        print(f"whole step {rep}: {1e3 * (t1 - t0):.3f} ms", flush=True)
for rep in range(3):
    if world > 1:
        dist.barrier()
    t = [now()]
    samples, cstats, meta = s.run(IT, C, 1, dtype, IT // 10, 0, True, "full", rank * C, None, bench.HIST_EVERY)
    t.append(now())
    sums = cstats.sum(dim=1)
    t.append(now())
    total, count = par.merge_moment_sums(sums, float(IT) * C)
    t.append(now())
    hist = meta["hist"]
    if world > 1:
        dist.all_reduce(hist)
    t.append(now())
    th = total.cpu().numpy()
    hh = hist.cpu().numpy()
    t.append(now())
    d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    print(f"rank {rank} rep {rep}: kernel {d[0]:.3f} | sum over chains {d[1]:.3f} | all-reduce moments (+ .item()) {d[2]:.3f} | "
          f"all-reduce histograms {d[3]:.3f} | to host {d[4]:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
