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


def all_ranks_agree(ok, what, group=None, device=None):
    """Raise ValueError(what) on EVERY rank if ``ok`` is false on any of them (one MIN all-reduce), so that
    a rank with an unusable share -- no chains, no rows -- can never leave the others waiting inside a later
    collective."""
    flag = 1 if ok else 0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([flag], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        flag = int(t.item())
    if not flag:
        raise ValueError(what)


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
    reduce.rank, reduce.world = _world(group)      # lets callers lay out an all-gather as a sum of slots
    return reduce


def row_range(n_rows_total, rank=None, world=None, group=None):
    """Training rows [start, stop) owned by ``rank`` (contiguous, sizes differ by at most one)."""
    return chain_range(n_rows_total, rank, world, group)


def sharded_orthogonalize(preds_rows, truth_rows, components_kept, *, group=None, device=None):
    """pybmc/bmc.py:102-130 with the ROWS of the prediction table split over ranks (configs[4]: 1e5 x 256):
    local centring and partial Gram, one all-reduce of the M-by-M matrix (512 KB at M = 256), the
    same small eigenproblem on every rank, local projection.  Returns this rank's y, mu, U_hat rows
    and the replicated S_hat, Vt_hat, Vt_hat_normalized."""
    from . import _device as D
    from .bmc import orthogonalize_arrays
    n_rows = int(np.asarray(preds_rows).shape[0])
    all_ranks_agree(n_rows >= 1, "sharded_orthogonalize: every rank needs at least one row of the table", group,
                    D.device(device))
    return orthogonalize_arrays(preds_rows, truth_rows, components_kept, method="auto", device=device,
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
# the sharded entry points (one process per GPU; call them from every rank)
# ------------------------------------------------------------------------------------------------
def sharded_gibbs(y, X, iterations, prior_info, n_chains_total, *, seed, dtype="float32", thin=1, discard=0,
                  keep_samples=False, group=None, device=None, rows_sharded=False, hist_every=0, as_numpy=True,
                  sampler=None):
    """Run this rank's share of ``n_chains_total`` conjugate chains and all-reduce the moment sums.

    Every rank returns the same posterior mean / covariance of [b, sigma] (what a single GPU running
    all chains would report) plus its own ``GibbsResult`` (samples of its chains, if kept).
    ``rows_sharded=True``: (y, X) are this rank's rows (from ``sharded_orthogonalize``); X'X, X'y,
    y'y and RSS_min are all-reduced first, so every rank samples the same posterior.
    ``hist_every`` > 0: marginal histograms of [b, sigma] are collected on every rank and summed by a second
    all-reduce (uint64 counts as int64: exact); the local result then carries the GLOBAL histograms, so
    ``local.quantiles(q)`` are the credible-interval endpoints of the whole run on every rank.
    ``sampler``: a ``ConjugateSampler`` already built for (y, X, prior_info) on this rank (skips the set-up).

    Collective: ONE all-reduce(sum, fp64) of [moment sums, count, histogram counts] -- 55 doubles at K = 8 plus,
    with histograms, (K+1) x 512 counts (36 KB at K = 8; exact in fp64)."""
    from . import _device as D
    from .inference_utils import ConjugateSampler, GibbsResult, _finish_samples, _moments_from_stats
    rank, world = _world(group)
    lo, hi = chain_range(n_chains_total, rank, world)
    dev = D.device(device) if sampler is None else sampler.dev
    # every rank can see that a share is empty from the totals alone: all of them raise, nobody enters a collective
    if int(n_chains_total) < world:
        raise ValueError(f"sharded_gibbs: {n_chains_total} chains do not cover all {world} ranks")
    if rows_sharded and sampler is None:      # row counts are only known locally: agree on them first
        all_ranks_agree(int(np.asarray(y).shape[0]) >= 1,
                        f"sharded_gibbs: the rows of the table do not cover all {world} ranks", group, dev)
    if sampler is None:
        sampler = ConjugateSampler(y, X, prior_info, dev, reduce=sum_over_ranks(group) if rows_sharded else None)
    stats = "full" if sampler.k <= 8 else "auto"           # cross moments only while they fit in registers
    samples, cstats, meta = sampler.run(iterations, hi - lo, seed, dtype, thin, discard, keep_samples, stats, lo,
                                        None, hist_every)
    with D.on(dev):
        # ONE collective and ONE read-back per step: [moment sums | count | histogram counts] travel together as
        # fp64 (a bin holds at most iterations / hist_every x chains draws, far below 2^53: the sum is exact)
        hist = meta["hist"]
        n_mom = int(cstats.shape[0])
        parts = [cstats.sum(dim=1), torch.full((1,), float(iterations) * (hi - lo), dtype=torch.float64, device=dev)]
        if hist is not None:
            parts.append(hist.reshape(-1).to(torch.float64))
        vec = torch.cat(parts)
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        host = vec.cpu().numpy()
        total, count = host[:n_mom], float(host[n_mom])
        k, kp = sampler.k, meta["kp"]
        mean_e, cov_e = _moments_from_stats(total, k, kp, meta["mode"], count)
        rows = _finish_samples(samples, as_numpy)
        hist = None if hist is None else np.rint(host[n_mom + 1:]).astype(np.int64).reshape(tuple(hist.shape))
    jac = np.zeros((k + 1, k + 1))
    jac[:k, :k] = sampler.w
    jac[k, k] = 1.0
    base = np.concatenate([sampler.w @ sampler.g_ols, [np.sqrt(sampler.sigma2_init)]])
    local = GibbsResult(samples=rows, mean=None, cov=None, chain_mean=None,
                        n_chains=hi - lo, iterations=int(iterations), n_kept=meta["n_kept"], seed=int(seed),
                        dtype=str(dtype), info=dict(chain_range=(lo, hi)), hist=hist,
                        hist_lo=sampler.hist_lo if hist is not None else None,
                        hist_width=sampler.hist_width if hist is not None else None)
    return base + jac @ mean_e, jac @ cov_e @ jac.T, local


def sharded_gibbs_simplex(y, X, Vt_hat, S_hat, iterations, prior_info, n_chains_total, *, burn=10000, stepsize=0.001,
                          seed, dtype="float32", thin=1, keep_samples=False, group=None, device=None, as_numpy=True,
                          sampler=None):
    """This rank's share of ``n_chains_total`` simplex-constrained chains (pybmc/inference_utils.py:59-144, every
    chain with its own burn-in) and ONE all-reduce (sum, fp64) of [moment sums | count | accepted proposals] --
    SURVEY.md section 8e's sufficient statistics and accept counts.  Every rank returns the same posterior mean /
    covariance of [b, sigma] and the overall acceptance rate of the sampling phase (what upstream prints, :143),
    plus its own ``GibbsResult`` (kept samples and per-chain acceptance of ITS chains)."""
    from . import _device as D
    from .inference_utils import GibbsResult, SimplexSampler, _finish_samples, _moments_from_stats
    if burn < 0:
        raise ValueError("Burn-in iterations must be non-negative.")
    if stepsize <= 0:
        raise ValueError("Stepsize must be positive.")
    rank, world = _world(group)
    if int(n_chains_total) < world:
        raise ValueError(f"sharded_gibbs_simplex: {n_chains_total} chains do not cover all {world} ranks")
    lo, hi = chain_range(n_chains_total, rank, world)
    if sampler is None:
        sampler = SimplexSampler(y, X, Vt_hat, S_hat, prior_info, stepsize, device)
    dev = sampler.dev
    samples, cstats, accepted, meta = sampler.run(iterations, burn, hi - lo, seed, dtype, thin, keep_samples, "auto", lo)
    with D.on(dev):
        n_mom = int(cstats.shape[0])
        vec = torch.cat([cstats.sum(dim=1),
                         torch.full((1,), float(iterations) * (hi - lo), dtype=torch.float64, device=dev),
                         accepted.sum(dtype=torch.float64).reshape(1)])
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        host = vec.cpu().numpy()
        count = float(host[n_mom])
        mean_e, cov_e = _moments_from_stats(host[:n_mom], sampler.k, meta["kp"], meta["mode"], count)
        rows = _finish_samples(samples, as_numpy)
        acc_local = D.to_host(accepted).astype(np.float64) / max(int(iterations), 1)
    sig_ref = np.sqrt(sampler.rss_zero / sampler.n) if sampler.rss_zero > 0 else 1.0
    base = np.concatenate([sampler.b_ols, [sig_ref]])
    local = GibbsResult(samples=rows, mean=None, cov=None, chain_mean=None, n_chains=hi - lo,
                        iterations=int(iterations), n_kept=meta["n_kept"], seed=int(seed), dtype=str(dtype),
                        acceptance=acc_local, info=dict(chain_range=(lo, hi)))
    return base + mean_e, cov_e, float(host[n_mom + 1]) / max(count, 1.0), local


def broadcast_draws(theta, k, *, src=0, group=None, device=None, split_upload=True):
    """The posterior rows used for prediction, resident on every rank's GPU after ONE trip of each row over PCIe
    (SURVEY.md section 8e asks for one broadcast of ``samples`` over NVLink instead of N host uploads).

    ``split_upload=True`` (``theta`` -- [S, K+1] host array -- is identical on all ranks, as
    ``sharded_predictive_summary`` requires): rank r uploads rows [r S/N, (r+1) S/N) and one all-gather over
    NVLink completes the table everywhere -- the staging and the PCIe copies of the N slices run side by side
    (13.6 MB at S = 1e5, K = 16: 1.7 MB per rank at N = 8 instead of all of it on rank ``src``).
    ``split_upload=False``: only rank ``src`` holds the values; it uploads them all and broadcasts."""
    from . import _device as D
    rank, world = _world(group)
    dev = D.device(device)
    s_rows = int(np.shape(theta)[0])
    with D.on(dev):
        if world == 1:
            return D.to_device(np.asarray(theta, dtype=np.float64), dev)
        if split_upload:
            per = -(-s_rows // world)
            full = torch.empty((world * per, k + 1), dtype=torch.float64, device=dev)
            lo, hi = min(rank * per, s_rows), min((rank + 1) * per, s_rows)
            if hi > lo:
                full[lo:hi].copy_(D.to_device(np.asarray(theta[lo:hi], dtype=np.float64), dev))
            if hi - lo < per:
                full[hi:(rank + 1) * per].zero_()
            dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per].clone(), group=group)
            return full[:s_rows]
        if rank == src:
            th = D.to_device(np.asarray(theta, dtype=np.float64), dev)
        else:
            th = torch.empty((s_rows, k + 1), dtype=torch.float64, device=dev)
        dist.broadcast(th, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return th


def sharded_predictive_summary(preds, theta, Vt_hat, *, truth=None, percentiles=(2.5, 50.0, 97.5), seed=0,
                               dtype="float32", group=None, device=None, gather=True, n_points_total=None,
                               draws_from=None):
    """Fused prediction with the nuclei split over ranks.  ``preds`` / ``truth`` are the FULL tables (every
    rank slices its own 4-aligned block of nuclei); ``theta`` are the posterior rows to use (already
    selected).  ``draws_from=None``: every rank holds them -- each uploads 1/N of the rows and ``broadcast_draws``
    completes the table with one all-gather over NVLink (rank r contributes rows [r S/N, (r+1) S/N) of ITS array, so
    ranks holding different draws of the same posterior end up with one common table).  ``draws_from=r``: only rank
    r's values count; it uploads them all and broadcasts.

    ``gather=True``: every rank returns the full-length outputs (ONE all-gather of the packed per-nucleus block
    [mean | var | percentiles | c_lt | c_le], (4 + Q) x 8 bytes per nucleus); ``gather="root"``: one gather to rank 0
    of the group, which alone copies the full-length outputs to its host (the other ranks return a result whose
    arrays are None) -- eight ranks reading 7 MB each back at once share the host link, one reads at full speed;
    ``gather=False``: each rank returns its own block (``PredictiveResult`` over nuclei ``point_range(...)``), no
    collective after the broadcast.  ``n_points_total``: ``preds`` / ``truth`` are ALREADY this rank's block ``point_range(n_points_total)``
    (tables too large to replicate on every host process)."""
    from . import _device as D
    from .sampling_utils import PredictiveProblem, PredictiveResult
    rank, world = _world(group)
    dev = D.device(device)
    sliced = n_points_total is not None
    n = int(n_points_total) if sliced else int(np.asarray(preds).shape[0])
    lo, hi = point_range(n, rank, world)
    if sliced and int(np.asarray(preds).shape[0]) != hi - lo:
        raise ValueError(f"rank {rank} was given {np.asarray(preds).shape[0]} rows for its block [{lo}, {hi})")
    k = int(np.asarray(Vt_hat).shape[0])
    nq = len(tuple(percentiles))
    th = broadcast_draws(theta, k, group=group, device=dev, src=draws_from or 0, split_upload=draws_from is None)
    with D.on(dev):
        rows = 2 + nq + 2
        per = point_range(n, 0, world)[1]
        block = torch.zeros((rows, per), dtype=torch.float64, device=dev)
        passes = 0
        if hi > lo:
            rows_ = slice(None) if sliced else slice(lo, hi)
            prob = PredictiveProblem(np.asarray(preds)[rows_], th, Vt_hat,
                                     truth=None if truth is None else np.asarray(truth)[rows_], dtype=dtype,
                                     device=dev, point0=lo)
            res = prob.run(percentiles=percentiles, seed=seed, as_numpy=False)
            passes = res.passes
            block[0, : hi - lo] = res.mean
            block[1, : hi - lo] = res.var
            block[2:2 + nq, : hi - lo] = res.percentiles
            if truth is not None:       # int64 counts travel as their bit patterns
                block[2 + nq, : hi - lo] = res.c_lt.view(torch.float64)
                block[3 + nq, : hi - lo] = res.c_le.view(torch.float64)
        to_root = isinstance(gather, str)
        if to_root and gather != "root":
            raise ValueError(f"gather must be True, False or 'root', got {gather!r}")
        holder = True
        if gather and world > 1:
            if to_root:
                holder = rank == 0
                parts = torch.empty((world, rows, per), dtype=torch.float64, device=dev) if holder else None
                root = dist.get_global_rank(group, 0) if group is not None else 0
                dist.gather(block, list(parts.unbind(0)) if holder else None, dst=root, group=group)
            else:
                parts = torch.empty((world, rows, per), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(parts, block, group=group)
            if holder:
                block = parts.permute(1, 0, 2).reshape(rows, world * per)[:, :n]
        elif not gather:
            block = block[:, : hi - lo]
        else:
            block = block[:, :n]
        if not holder:
            return PredictiveResult(mean=None, var=None, percentiles=None, c_lt=None, c_le=None, draws=None,
                                    n_draws=int(np.shape(theta)[0]), passes=passes, seed=int(seed))
        out = D.to_host(block.contiguous()) if block.numel() else block.cpu().numpy()
    counts = None if truth is None else out[2 + nq:].view(np.int64)
    return PredictiveResult(mean=out[0], var=out[1], percentiles=out[2:2 + nq],
                            c_lt=None if counts is None else counts[0], c_le=None if counts is None else counts[1],
                            draws=None, n_draws=int(np.shape(theta)[0]), passes=passes, seed=int(seed))
