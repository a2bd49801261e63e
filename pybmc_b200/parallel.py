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


def sum_over_ranks(group=None):
    """A ``reduce`` callable for ``orthogonalize_arrays`` / ``ConjugateSampler``: in-place all-reduce
    (sum) of a tensor over ``group``; the identity when not distributed."""
    def reduce(t):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t
    return reduce


def row_range(n_rows_total, rank=None, world=None, group=None):
    """Training rows [start, stop) owned by ``rank`` (contiguous, sizes differ by at most one)."""
    return chain_range(n_rows_total, rank, world, group)


def sharded_orthogonalize(preds_rows, truth_rows, components_kept, *, group=None, device=None):
    """pybmc/bmc.py:102-130 with the ROWS of the prediction table split over ranks (configs[4]: 1e5 x 256):
    local centring and partial Gram, one all-reduce of the M-by-M matrix (512 KB at M = 256), the
    same small eigenproblem on every rank, local projection.  Returns this rank's y, mu, U_hat rows
    and the replicated S_hat, Vt_hat, Vt_hat_normalized."""
    from .bmc import orthogonalize_arrays
    return orthogonalize_arrays(preds_rows, truth_rows, components_kept, method="gram", device=device,
                                reduce=sum_over_ranks(group))


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


# ------------------------------------------------------------------------------------------------
# the two sharded entry points (one process per GPU; call them from every rank)
# ------------------------------------------------------------------------------------------------
def sharded_gibbs(y, X, iterations, prior_info, n_chains_total, *, seed, dtype="float32", thin=1, discard=0,
                  keep_samples=False, group=None, device=None, rows_sharded=False):
    """Run this rank's share of ``n_chains_total`` conjugate chains and all-reduce the moment sums.

    Every rank returns the same posterior mean / covariance of [b, sigma] (what a single GPU running
    all chains would report) plus its own ``GibbsResult`` (samples of its chains, if kept).
    ``rows_sharded=True``: (y, X) are this rank's rows (from ``sharded_orthogonalize``); X'X, X'y,
    y'y and RSS_min are all-reduced first, so every rank samples the same posterior."""
    from .inference_utils import ConjugateSampler, GibbsResult, _finish_samples, _moments_from_stats
    rank, world = _world(group)
    lo, hi = chain_range(n_chains_total, rank, world)
    sampler = ConjugateSampler(y, X, prior_info, device, reduce=sum_over_ranks(group) if rows_sharded else None)
    stats = "full" if sampler.k <= 8 else "auto"           # cross moments only while they fit in registers
    samples, cstats, meta = sampler.run(iterations, hi - lo, seed, dtype, thin, discard, keep_samples, stats, lo)
    total, count = merge_moment_sums(cstats.sum(dim=1), float(iterations) * (hi - lo), group)
    k, kp = sampler.k, meta["kp"]
    mean_e, cov_e = _moments_from_stats(total.cpu().numpy(), k, kp, meta["mode"], count)
    jac = np.zeros((k + 1, k + 1))
    jac[:k, :k] = sampler.w
    jac[k, k] = 1.0
    base = np.concatenate([sampler.w @ sampler.g_ols, [np.sqrt(sampler.sigma2_init)]])
    local = GibbsResult(samples=_finish_samples(samples, True), mean=None, cov=None, chain_mean=None,
                        n_chains=hi - lo, iterations=int(iterations), n_kept=meta["n_kept"], seed=int(seed),
                        dtype=str(dtype), info=dict(chain_range=(lo, hi)))
    return base + jac @ mean_e, jac @ cov_e @ jac.T, local


def sharded_predictive_summary(preds, theta, Vt_hat, *, truth=None, percentiles=(2.5, 50.0, 97.5), seed=0,
                               dtype="float32", group=None, device=None):
    """Fused prediction with the nuclei split over ranks; every rank returns the full-length outputs
    (all-gather).  ``theta`` are the posterior rows to use (already selected), identical on all ranks."""
    from .sampling_utils import PredictiveProblem, PredictiveResult
    rank, world = _world(group)
    n = int(np.asarray(preds).shape[0])
    lo, hi = point_range(n, rank, world)
    if hi > lo:
        prob = PredictiveProblem(np.asarray(preds)[lo:hi], theta, Vt_hat,
                                 truth=None if truth is None else np.asarray(truth)[lo:hi], dtype=dtype,
                                 device=device, point0=lo)
        res = prob.run(percentiles=percentiles, seed=seed, as_numpy=False)
        dev = res.mean.device
        parts = dict(mean=res.mean, var=res.var, percentiles=res.percentiles, c_lt=res.c_lt, c_le=res.c_le)
        passes = res.passes
    else:   # more ranks than 4-aligned blocks of nuclei
        dev = torch.device("cuda", torch.cuda.current_device())
        nq = len(tuple(percentiles))
        parts = dict(mean=torch.zeros(0, dtype=torch.float64, device=dev),
                     var=torch.zeros(0, dtype=torch.float64, device=dev),
                     percentiles=torch.zeros((nq, 0), dtype=torch.float64, device=dev),
                     c_lt=None if truth is None else torch.zeros(0, dtype=torch.int64, device=dev),
                     c_le=None if truth is None else torch.zeros(0, dtype=torch.int64, device=dev))
        passes = 0
    out = {k: (None if v is None else gather_points(v, n, group).cpu().numpy()) for k, v in parts.items()}
    return PredictiveResult(mean=out["mean"], var=out["var"], percentiles=out["percentiles"], c_lt=out["c_lt"],
                            c_le=out["c_le"], draws=None, n_draws=int(np.asarray(theta).shape[0]), passes=passes,
                            seed=int(seed))
