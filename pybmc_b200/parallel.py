"""Sharding of chains and nuclei over ranks (one process per GPU, torch.distributed).

The path has no data-path exchange step: chains and nuclei are independent, and every random
variate is keyed on the *global* chain / nucleus id, so a rank only needs
its own range.  The collectives are the ones SURVEY.md section 8(e) lists: one all-reduce of the
fp64 moment sums after sampling, one all-gather of per-nucleus outputs after prediction.  The
functions take an optional process group, so the same code runs over NCCL on GPUs and over gloo in
the CPU tests (where the per-shard inputs come from the oracle).
"""
import numpy as np
import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def chain_range(n_chains_total, rank=None, world=None, group=None):
    """Global chain ids [start, stop) owned by ``rank``: contiguous, sizes differ by at most one."""
    if rank is None or world is None:
        rank, world = _world(group)
    base, rem = divmod(int(n_chains_total), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def point_range(n_points_total, rank=None, world=None, group=None, align=4):
    """Nuclei [start, stop) owned by ``rank``; starts are multiples of 4 because one Philox call
    produces the noise of a quad of consecutive nuclei."""
    if rank is None or world is None:
        rank, world = _world(group)
    n = int(n_points_total)
    per = -(-n // world)
    per = -(-per // align) * align
    start = min(rank * per, n)
    return start, min(start + per, n)


def merge_moment_sums(local_sums, local_count, group=None):
    """All-reduce (sum) of per-rank moment sums and sample counts -> global (sums, count).

    ``local_sums`` is the fp64 vector ``chain_stats.sum(dim=1)`` of this rank's chains; the result is
    what a single rank holding every chain would have computed (up to fp64 summation order)."""
    t = torch.cat([torch.as_tensor(local_sums, dtype=torch.float64).reshape(-1),
                   torch.tensor([float(local_count)], dtype=torch.float64,
                                device=getattr(local_sums, "device", None))])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[:-1], float(t[-1].item())


def gather_points(local, n_points_total, group=None, align=4):
    """All-gather per-nucleus outputs (last dimension = this rank's nuclei) into the global order."""
    rank, world = _world(group)
    local = torch.as_tensor(local)
    if world == 1:
        return local
    per = point_range(n_points_total, 0, world, align=align)[1]
    pad = torch.zeros(local.shape[:-1] + (per,), dtype=local.dtype, device=local.device)
    pad[..., : local.shape[-1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat(parts, dim=-1)
    return out[..., : int(n_points_total)]


def merge_coverage_counts(local_covered, group=None):
    """All-reduce (sum) of the per-level covered counts (int64): exact."""
    t = torch.as_tensor(local_covered, dtype=torch.int64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def posterior_from_sums(sums, count, k):
    """First / full second moment sums of a (k+1)-vector -> (mean, covariance)."""
    sums = np.asarray(sums, dtype=np.float64)
    d = k + 1
    mean = sums[:d] / count
    cov = np.zeros((d, d))
    idx = d
    for r in range(d):
        for c in range(r, d):
            cov[r, c] = cov[c, r] = sums[idx] / count - mean[r] * mean[c]
            idx += 1
    return mean, cov
