"""Samplers and SVD truncation with the reference's signatures
(pybmc/inference_utils.py:4, :59, :147), computed by sm_100a kernels.

``gibbs_sampler`` and ``gibbs_sampler_simplex`` take and return the same NumPy arrays as
upstream.  Extra keyword-only options (all optional; the defaults reproduce the reference's
single chain) expose what the GPU adds: many independent chains, a seed, fp32 arithmetic,
thinning, and posterior moments accumulated on the device.

The per-iteration K-by-K inverse, X'y and residual of the reference are replaced by
sufficient statistics computed once (``bmc_gram``, ``bmc_residual_ss``) and a simultaneous
diagonalisation of (X'X, Lambda + 1e-6 I) done on the host in fp64 -- O(K^3) set-up work, not
a CPU fallback: every iteration of every chain runs in ``bmc_gibbs_run`` /
``bmc_gibbs_simplex_run``.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _device as D
from . import _lib

RIDGE = 1e-6          # pybmc/inference_utils.py:41
SIGMA2_FLOOR = 1e-6   # pybmc/inference_utils.py:37,52


@dataclass
class GibbsResult:
    """Outcome of a batched sampler run.

    samples : [n_chains * n_kept, K+1] rows [b_0..b_{K-1}, sigma], chain-major (chain c owns rows
              c*n_kept .. (c+1)*n_kept-1); a device tensor when ``as_numpy=False``; None if not kept
    mean, cov : posterior mean / covariance of [b, sigma] from device-side fp64 moment sums over
              ALL iterations of all chains (not only the kept ones); None when stats are off
    chain_mean : [n_chains, K+1] per-chain means (for between-chain diagnostics)
    acceptance : per-chain sampling-phase acceptance fraction (simplex sampler only)
    """
    samples: object
    mean: np.ndarray
    cov: np.ndarray
    chain_mean: np.ndarray
    n_chains: int
    iterations: int
    n_kept: int
    seed: int
    dtype: str
    acceptance: np.ndarray = None
    info: dict = field(default_factory=dict)
    rhat: np.ndarray = None   # [K+1] Gelman-Rubin potential scale reduction (needs >= 2 chains, full stats)
    ess: np.ndarray = None    # [K+1] effective sample size of the whole run (all chains), same requirements
    hist: np.ndarray = None   # [K+1, HIST_BINS] int64 marginal histograms of [b, sigma] (hist_every > 0)
    hist_lo: np.ndarray = None      # [K+1] lower edge of bin 0
    hist_width: np.ndarray = None   # [K+1] bin width

    def quantiles(self, probs):
        """Marginal quantiles of [b_0..b_{K-1}, sigma] from the device histograms: ``[len(probs), K+1]``
        (probs in percent, like ``np.percentile``).  NaN where the quantile falls into an edge bin (the
        posterior reached beyond the binned range)."""
        if self.hist is None:
            raise ValueError("no histograms were collected: run with hist_every > 0")
        return histogram_quantiles(self.hist, self.hist_lo, self.hist_width, probs)


def histogram_quantiles(counts, lo, width, probs):
    """Quantiles of binned data, linear inside the bin that holds the target rank.  counts [D, B]."""
    counts = np.asarray(counts, dtype=np.float64)
    probs = np.atleast_1d(np.asarray(probs, dtype=np.float64))
    if np.any(probs < 0) or np.any(probs > 100):
        raise ValueError("Percentiles must be in the range [0, 100]")
    d, nb = counts.shape
    out = np.full((probs.shape[0], d), np.nan)
    cum = np.cumsum(counts, axis=1)
    for c in range(d):
        total = cum[c, -1]
        if total <= 0:
            continue
        for j, p in enumerate(probs):
            target = p / 100.0 * total
            b = int(np.searchsorted(cum[c], target, side="left"))
            b = min(b, nb - 1)
            if b == 0 or b == nb - 1 or counts[c, b] == 0:
                continue                      # edge bins collect everything beyond the range
            below = cum[c, b - 1]
            out[j, c] = lo[c] + (b + (target - below) / counts[c, b]) * width[c]
    return out


def USVt_hat_extraction(U, S, Vt, components_kept):
    """Truncate an SVD to ``components_kept`` components (pybmc/inference_utils.py:147-168).

    Returns ``(U_hat, S_hat, Vt_hat, Vt_hat_normalized)`` where -- as upstream --
    ``Vt_hat`` holds the right singular vectors divided by their singular values and
    ``Vt_hat_normalized`` the plain ones.  Pure slicing of host arrays: no kernel involved.
    """
    U = np.asarray(U)
    S = np.asarray(S)
    Vt = np.asarray(Vt)
    k = int(components_kept)
    U_hat = U.T[:k].copy().T
    S_hat = S[:k]
    Vt_hat_normalized = np.array(Vt[:k], copy=True)
    Vt_hat = Vt_hat_normalized / S_hat[:, None]
    return U_hat, S_hat, Vt_hat, Vt_hat_normalized


# --------------------------------------------------------------------------------------------
# sufficient statistics on the device
# --------------------------------------------------------------------------------------------
def _gram_with_response(Xd, yd, dev, reduce=None):
    """[X | y]'[X | y] by ``bmc_gram`` -> (X'X, X'y, y'y) on the host (fp64).  ``reduce`` sums the
    (K+1)-by-(K+1) matrix over the ranks holding the other rows (pybmc_b200.parallel)."""
    lib = _lib.load()
    n, k = Xd.shape
    out = torch.empty((k + 1, k + 1), dtype=torch.float64, device=dev)
    nbytes = lib.bmc_gram_workspace_bytes(n, k + 1)
    ws = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)
    _lib.check(lib.bmc_gram(D.ptr(Xd), n, k, Xd.stride(0), None, D.ptr(yd), D.ptr(out), D.ptr(ws), ws.numel(),
                            D.stream_ptr(dev)), "bmc_gram")
    if reduce is not None:
        out = reduce(out)
    a = D.to_host(out)
    return a[:k, :k].copy(), a[:k, k].copy(), float(a[k, k])


def _residual_ss(Xd, yd, b, dev, reduce=None):
    """|y - X b|^2 by ``bmc_residual_ss`` (pybmc/inference_utils.py:29-31)."""
    lib = _lib.load()
    n, k = Xd.shape
    bd = D.to_device(b, dev)
    out = torch.empty(1, dtype=torch.float64, device=dev)
    ws = torch.empty(max(int(lib.bmc_rss_workspace_bytes(n)), 8), dtype=torch.uint8, device=dev)
    _lib.check(lib.bmc_residual_ss(D.ptr(Xd), n, k, Xd.stride(0), D.ptr(yd), D.ptr(bd), D.ptr(out), D.ptr(ws),
                                   ws.numel(), D.stream_ptr(dev)), "bmc_residual_ss")
    if reduce is not None:
        out = reduce(out)
    return float(out.item())


def _as_design(y, X, dev):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2 or X.shape[0] != y.shape[0]:
        raise ValueError(f"X must be [len(y), K]; got X {X.shape}, y {y.shape}")
    if X.shape[1] < 1 or X.shape[1] > _lib.MAX_COMPONENTS:
        raise ValueError(f"number of components must be in 1..{_lib.MAX_COMPONENTS}, got {X.shape[1]}")
    return D.to_device(y, dev), D.to_device(X, dev)


def _stat_layout(k, kp, mode):
    """Index helpers for the moment rows described in include/bmc_b200.h."""
    d = kp + 1
    comp = list(range(k)) + [kp]           # real components, then sigma
    if mode == _lib.STATS_DIAG:
        second = {(i, i): d + i for i in range(d)}
    else:
        second, idx = {}, d
        for r in range(d):
            for c in range(r, d):
                second[(r, c)] = idx
                idx += 1
    return d, comp, second


def _moments_from_stats(stats, k, kp, mode, count):
    """Sum of deviations -> (mean_e [k+1], cov_e [k+1, k+1]) for components + sigma."""
    d, comp, second = _stat_layout(k, kp, mode)
    s1 = stats[:d] / count
    mean = s1[comp]
    cov = np.zeros((k + 1, k + 1))
    for a, ia in enumerate(comp):
        for b, ib in enumerate(comp):
            key = (min(ia, ib), max(ia, ib))
            if key in second:
                cov[a, b] = stats[second[key]] / count - s1[ia] * s1[ib]
    return mean, cov


class ConjugateSampler:
    """Device-resident set-up of pybmc/inference_utils.py:21-37 plus launches of the hot loop.

    Building the object uploads (y, X), computes X'X, X'y, y'y, b_ols, RSS_min and the
    diagonalising transform; ``run`` only launches ``bmc_gibbs_run`` (what bench.py times as the
    HBM-resident step).

    ``reduce`` (optional): a callable summing a device tensor over ranks.  With it, (y, X) are this
    rank's ROWS of the training set and the sufficient statistics are the all-reduced ones, so every
    rank ends up with the same sampler constants (SURVEY.md section 8e, configs[4]).
    """

    def __init__(self, y, X, prior_info, device=None, reduce=None):
        self.dev = D.device(device)
        with D.on(self.dev):
            self._setup(y, X, prior_info, reduce)

    def _setup(self, y, X, prior_info, reduce):
        b0, B0, nu0, sigma20 = prior_info
        yd, Xd = _as_design(y, X, self.dev)
        self.n, self.k = Xd.shape
        if reduce is not None:
            self.n = int(round(float(reduce(torch.tensor([float(self.n)], dtype=torch.float64,
                                                         device=self.dev)).item())))
        b0 = np.asarray(b0, dtype=np.float64).reshape(-1)
        B0 = np.asarray(B0, dtype=np.float64)
        if b0.shape[0] != self.k or B0.shape != (self.k, self.k):
            raise ValueError("prior mean / covariance do not match the number of components")
        lam = np.linalg.inv(B0)                                   # :22
        gram, xty, yty = _gram_with_response(Xd, yd, self.dev, reduce)    # :25 (+ X'y, y'y)
        gram_inv = np.linalg.inv(gram)                            # :26 (LinAlgError on singular X'X)
        b_ols = gram_inv @ xty                                    # :28
        rss_min = _residual_ss(Xd, yd, b_ols, self.dev, reduce)   # :29-31
        self.sigma2_init = max(rss_min / self.n, SIGMA2_FLOOR)    # :31, :37
        low = np.linalg.cholesky(lam + RIDGE * np.eye(self.k))    # prior precision incl. the ridge of :41
        low_inv = np.linalg.inv(low)
        a = low_inv @ gram @ low_inv.T
        a = 0.5 * (a + a.T)
        off = a - np.diag(np.diag(a))
        if np.max(np.abs(off), initial=0.0) <= 1e-14 * np.max(np.abs(np.diag(a))):
            d, q = np.diag(a).copy(), np.eye(self.k)
        else:
            d, q = np.linalg.eigh(a)
        w = low_inv.T @ q
        g_ols = q.T @ (low.T @ b_ols)
        pull = w.T @ (lam @ b0) - g_ols
        w_off = w - np.diag(np.diag(w))
        self.dense_w = bool(np.any(w_off != 0.0))
        self.gram, self.xty, self.yty, self.b_ols, self.rss_min = gram, xty, yty, b_ols, rss_min
        self.lam, self.b0, self.nu0, self.sigma20 = lam, b0, float(nu0), float(sigma20)
        self.w, self.d, self.g_ols, self.pull = w, d, g_ols, pull
        wflat = w.reshape(-1) if self.dense_w else np.diag(w).copy()
        self.hist_lo, self.hist_width = self.histogram_range()
        self._consts = D.to_device(np.concatenate([d, pull, g_ols, self.hist_lo, 1.0 / self.hist_width, wflat]),
                                   self.dev)
        self._Xd, self._yd = Xd, yd

    def histogram_range(self, span=10.0):
        """Binned range of [b, sigma] for the marginal histograms: centre +- ``span`` posterior standard
        deviations, both taken at sigma^2 = sigma2_init from the conditional law of :41-45 (b) and from the
        Inverse-Gamma law of :50-52 (sigma); _lib.HIST_BINS equal bins.  Returns (lo [K+1], width [K+1])."""
        k, s2 = self.k, self.sigma2_init
        var_e = s2 / (self.d + s2)                                # 1 / p_k
        centre = self.w @ (self.g_ols + self.pull * var_e)
        sd = np.sqrt((self.w ** 2) @ var_e)
        a = 0.5 * (self.nu0 + self.n)
        sig_c = np.sqrt((self.nu0 * self.sigma20 + self.rss_min + k * s2) / (2.0 * a))
        spread = np.exp(span / (2.0 * np.sqrt(a)))                # sd of log sigma^2 ~ 1 / sqrt(a)
        lo = np.concatenate([centre - span * sd, [sig_c / spread]])
        hi = np.concatenate([centre + span * sd, [sig_c * spread]])
        width = np.maximum(hi - lo, 1e-300) / _lib.HIST_BINS
        return lo, width

    def problem(self, layout=None):
        base = self._consts.data_ptr()
        k = self.k
        return _lib.GibbsProblem(k=k, d=base, pull=base + 8 * k, g_ols=base + 16 * k, w=base + 8 * (5 * k + 2),
                                 dense_w=int(self.dense_w), rss_min=self.rss_min, n_obs=float(self.n),
                                 nu0=self.nu0, sigma20=self.sigma20, sigma2_init=self.sigma2_init,
                                 layout=_lib.LAYOUTS[layout])

    @D.on_own_device
    def run(self, iterations, n_chains=1, seed=0, dtype="float64", thin=1, discard=0, keep_samples=True,
            stats="auto", chain_offset=0, layout=None, hist_every=0, persistent=True):
        """Launch the sampler; returns device tensors (samples [n_kept, K+1, C] or None, chain_stats) and a
        dict with the launch's layout constants (and ``hist``: uint64 counts [K+1, HIST_BINS] as an int64
        tensor when ``hist_every`` > 0).  ``layout``: None/"auto", "thread", "group", "warp".  ``persistent=False``
        withholds the workspace of the persistent launch (A/B runs: the plain one-warp-per-32-chains launch)."""
        lib = _lib.load()
        tdt, code = D.resolve_dtype(dtype)
        iterations, n_chains, thin, discard = int(iterations), int(n_chains), int(thin), int(discard)
        if iterations < 0 or n_chains < 1 or thin < 1 or discard < 0:
            raise ValueError("iterations >= 0, n_chains >= 1, thin >= 1 and discard >= 0 are required")
        kp = lib.bmc_padded_components(self.k)
        # "auto": all cross moments while they fit in registers (K <= 8: 45 sums), marginal moments beyond
        # (K = 16 would need 153 sums per thread and runs 2.3x slower, profiles/k_sweep.py); "full" is
        # available on request up to K = 16
        mode = {"auto": _lib.STATS_FULL if self.k <= 8 else _lib.STATS_DIAG, "full": _lib.STATS_FULL,
                "diag": _lib.STATS_DIAG, "none": _lib.STATS_NONE, None: _lib.STATS_NONE}[stats]
        n_kept = max(0, -(-(iterations - discard) // thin)) if keep_samples else 0
        samples = (torch.empty((n_kept, self.k + 1, n_chains), dtype=tdt, device=self.dev)
                   if keep_samples else None)
        n_stat = lib.bmc_gibbs_n_stat(kp, mode)
        cstats = torch.empty((n_stat, n_chains), dtype=torch.float64, device=self.dev) if mode else None
        prob = self.problem(layout)
        # room for the persistent launch (chain groups handed round the resident warps: no uneven tail); the library
        # only uses it for thread-per-chain launches that would load the schedulers unevenly
        if persistent:
            ws = torch.empty(int(lib.bmc_gibbs_workspace_bytes(n_chains)), dtype=torch.uint8, device=self.dev)
            prob.workspace, prob.workspace_bytes = ws.data_ptr(), ws.numel()
        hist_every = int(hist_every)
        if hist_every < 0:
            raise ValueError("hist_every must be >= 0")
        hist = hist_arg = None
        if hist_every:
            base, k = self._consts.data_ptr(), self.k
            hist = torch.empty((k + 1, _lib.HIST_BINS), dtype=torch.int64, device=self.dev)
            hws = torch.empty(int(lib.bmc_gibbs_hist_workspace_bytes(k)), dtype=torch.uint8, device=self.dev)
            hist_arg = C.byref(_lib.GibbsHist(every=hist_every, lo=base + 24 * k, inv_width=base + 8 * (4 * k + 1),
                                              counts=hist.data_ptr(), workspace=hws.data_ptr(),
                                              workspace_bytes=hws.numel()))
        _lib.check(lib.bmc_gibbs_run(code, C.byref(prob), int(seed) & (2 ** 64 - 1), int(chain_offset), n_chains,
                                     iterations, discard, thin, n_kept, D.ptr(samples), D.ptr(cstats), mode,
                                     hist_arg, D.stream_ptr(self.dev)), "bmc_gibbs_run")
        return samples, cstats, dict(kp=kp, mode=mode, n_kept=n_kept, hist=hist)

    @D.on_own_device
    def summarise(self, cstats, meta, iterations, n_chains):
        """Device moment sums -> posterior mean/cov of [b, sigma] and per-chain means (host, fp64)."""
        if cstats is None or iterations == 0:
            return None, None, None
        kp, mode, k = meta["kp"], meta["mode"], self.k
        total = D.to_host(cstats.sum(dim=1))
        mean_e, cov_e = _moments_from_stats(total, k, kp, mode, float(iterations) * n_chains)
        jac = np.zeros((k + 1, k + 1))
        jac[:k, :k] = self.w
        jac[k, k] = 1.0
        base = np.concatenate([self.w @ self.g_ols, [np.sqrt(self.sigma2_init)]])
        mean = base + jac @ mean_e
        cov = jac @ cov_e @ jac.T
        comp = list(range(k)) + [kp]
        jac_d = torch.from_numpy(jac).to(cstats.device)
        base_d = torch.from_numpy(base).to(cstats.device)
        # [C, K+1] x [K+1, K+1] as a broadcast product: element-wise kernels only, no cuBLAS handle for this
        chain_mean_d = base_d[None, :] + ((cstats[comp, :].t() / float(iterations))[:, None, :]
                                          * jac_d[None, :, :]).sum(dim=2)
        self.last_rhat, self.last_ess = _chain_diagnostics(cstats, comp, _stat_layout(k, kp, mode)[2], jac_d,
                                                           chain_mean_d, iterations)
        return mean, cov, D.to_host(chain_mean_d)


def _chain_diagnostics(cstats, comp, second, jac_d, chain_mean_d, iterations):
    """Between-chain diagnostics per coordinate of [b, sigma] from the per-chain moment sums (device, fp64).

    R-hat (Gelman-Rubin): sqrt(((n-1)/n W + B/n) / W), W the mean within-chain variance and B/n the
    variance of the chain means.  Effective sample size of the whole run: a chain mean over n iterations
    has variance var * tau / n (tau = integrated autocorrelation time), and with many independent chains
    that variance is *observed* as B/n, so ESS = chains * n / tau = chains * var+ / (B/n), var+ the
    pooled variance.  Returns (None, None) when there is a single chain or only diagonal moments of
    rotated coordinates."""
    n_chains = cstats.shape[1]
    if n_chains < 2 or iterations < 2:
        return None, None
    d = len(comp)
    idx = [[second.get((min(a, b), max(a, b))) for b in comp] for a in comp]
    m1 = cstats[comp, :].t() / float(iterations)
    # W, the mean over chains of the within-chain variances, is linear in the second-moment sums: only their TOTALS
    # over the chains are needed, together with the per-chain first moments (one [d, d] matrix instead of a
    # [chains, d, d] intermediate).
    scale = iterations / (iterations - 1.0)
    if any(i is None for row in idx for i in row):
        # diagonal moments only: enough when the coordinates are not rotated (diagonal W)
        if bool((jac_d - torch.diag(torch.diagonal(jac_d))).abs().max() > 0):
            return None, None
        diag_rows = torch.tensor([second[(a, a)] for a in comp], device=cstats.device)
        var_e = cstats[diag_rows, :].sum(dim=1) / (float(iterations) * n_chains) - (m1 * m1).mean(dim=0)
        w = var_e * torch.diagonal(jac_d) ** 2 * scale
    else:
        flat = torch.tensor([i for row in idx for i in row], device=cstats.device)
        m2 = cstats[flat, :].sum(dim=1).reshape(d, d) / (float(iterations) * n_chains)     # E[e e'], all chains
        cov_e = m2 - (m1[:, :, None] * m1[:, None, :]).mean(dim=0)                          # mean within-chain covariance
        w = (jac_d[:, :, None] * cov_e[None, :, :] * jac_d[:, None, :]).sum(dim=(1, 2)) * scale   # diag(J C J'), no cuBLAS
    b_over_n = chain_mean_d.var(dim=0, unbiased=True)
    var_plus = (iterations - 1.0) / iterations * w + b_over_n
    rhat = torch.sqrt(var_plus / w)
    ess = n_chains * var_plus / b_over_n
    return rhat.cpu().numpy(), ess.cpu().numpy()


def _finish_samples(samples, as_numpy):
    """[kept, K+1, chains] on the device -> the reference's row layout [chains*kept, K+1] (chain-major).
    The array keeps the arithmetic type of the run: float64 by default, float32 when asked for."""
    if samples is None:
        return None
    n_kept, width, n_chains = samples.shape
    rows = samples.permute(2, 0, 1).reshape(n_chains * n_kept, width)
    return D.to_host(rows) if as_numpy else rows


def run_gibbs(y, X, iterations, prior_info, *, n_chains=1, seed=None, dtype="float64", thin=1, discard=0,
              keep_samples=True, stats="auto", device=None, chain_offset=0, as_numpy=True, layout=None,
              hist_every=0):
    """Batched conjugate sampler: ``n_chains`` independent copies of the reference's chain.

    The moment sums (``mean``, ``cov``, ``rhat``, ``ess``) and the histograms cover ALL iterations of all chains,
    including the first ``discard`` ones -- ``discard`` / ``thin`` only select which iterates are *stored*; the
    conjugate sampler starts at the OLS variance and has no burn-in upstream either (:37-39).
    ``hist_every`` > 0 (a multiple of 64: the kernels bin where they flush their moment sums) bins the state
    after every hist_every-th iteration into marginal histograms (``GibbsResult.hist`` / ``.quantiles``)."""
    seed = D.fresh_seed() if seed is None else int(seed)
    sampler = ConjugateSampler(y, X, prior_info, device)
    samples, cstats, meta = sampler.run(iterations, n_chains, seed, dtype, thin, discard, keep_samples, stats,
                                        chain_offset, layout, hist_every)
    mean, cov, chain_mean = sampler.summarise(cstats, meta, int(iterations), int(n_chains))
    with D.on(sampler.dev):
        hist = None if meta["hist"] is None else meta["hist"].cpu().numpy()
        rows = _finish_samples(samples, as_numpy)
    return GibbsResult(samples=rows, mean=mean, cov=cov, chain_mean=chain_mean,
                       n_chains=int(n_chains), iterations=int(iterations), n_kept=meta["n_kept"], seed=seed,
                       dtype=str(dtype), info=dict(rss_min=sampler.rss_min, b_ols=sampler.b_ols,
                                                   sigma2_init=sampler.sigma2_init),
                       rhat=getattr(sampler, "last_rhat", None), ess=getattr(sampler, "last_ess", None),
                       hist=hist, hist_lo=sampler.hist_lo if hist is not None else None,
                       hist_width=sampler.hist_width if hist is not None else None)


def gibbs_sampler(y, X, iterations, prior_info, *, n_chains=1, seed=None, dtype="float64", thin=1, discard=0,
                  device=None):
    """Gibbs sampling for Bayesian linear regression (pybmc/inference_utils.py:4-56).

    Args:
        y: response vector (centred).
        X: design matrix ``[len(y), K]``.
        iterations: iterations per chain.
        prior_info: ``(b_mean_prior, b_mean_cov, nu0, sigma20)``.
        n_chains, seed, dtype, thin, discard, device: optional GPU controls; the defaults run one
            fp64 chain from a fresh seed, like upstream.

    Returns:
        ``[n_chains * kept, K+1]`` array of ``[beta, sigma]`` rows (``[iterations, K+1]`` float64 by
        default; float32 if ``dtype="float32"`` was asked for), no burn-in, first iterate kept -- as
        upstream.
    """
    return run_gibbs(y, X, iterations, prior_info, n_chains=n_chains, seed=seed, dtype=dtype, thin=thin,
                     discard=discard, stats="none", device=device).samples


# --------------------------------------------------------------------------------------------
class SimplexSampler:
    """Set-up of pybmc/inference_utils.py:78-94 on the device + launches of the two loops."""

    def __init__(self, y, X, Vt_hat, S_hat, prior_info, stepsize=0.001, device=None):
        self.dev = D.device(device)
        with D.on(self.dev):
            self._setup(y, X, Vt_hat, S_hat, prior_info, stepsize)

    def _setup(self, y, X, Vt_hat, S_hat, prior_info, stepsize):
        nu0, sigma20 = prior_info
        Vt_hat = np.asarray(Vt_hat, dtype=np.float64)
        S_hat = np.asarray(S_hat, dtype=np.float64).reshape(-1)
        yd, Xd = _as_design(y, X, self.dev)
        self.n, self.k = Xd.shape
        if Vt_hat.ndim != 2 or Vt_hat.shape[0] != self.k or S_hat.shape[0] != self.k:
            raise ValueError("Vt_hat must be [K, M] and S_hat [K] with K = X.shape[1]")
        self.m = Vt_hat.shape[1]
        gram, xty, yty = _gram_with_response(Xd, yd, self.dev)
        # any least-squares solution serves RSS(b) = RSS_min + (b-b_ols)'G(b-b_ols); pinv also covers
        # rank-deficient designs, which the reference's simplex sampler accepts (it never inverts X'X)
        b_ols = np.linalg.pinv(gram) @ xty
        self.rss_min = _residual_ss(Xd, yd, b_ols, self.dev)
        self.gram, self.b_ols, self.rss_zero = gram, b_ols, yty
        self.nu0, self.sigma20 = float(nu0), float(sigma20)
        step = S_hat * float(stepsize)                            # sqrt of diag(S^2 step^2), :80
        self._consts = D.to_device(np.concatenate([gram.reshape(-1), b_ols, step, Vt_hat.reshape(-1)]), self.dev)

    def problem(self, layout=None):
        base, k = self._consts.data_ptr(), self.k
        return _lib.SimplexProblem(k=k, m=self.m, gram=base, b_ols=base + 8 * k * k, step=base + 8 * (k * k + k),
                                   vt_hat=base + 8 * (k * k + 2 * k), rss_min=self.rss_min, rss_zero=self.rss_zero,
                                   n_obs=float(self.n), nu0=self.nu0, sigma20=self.sigma20,
                                   layout=_lib.LAYOUTS[layout])

    @D.on_own_device
    def run(self, iterations, burn, n_chains=1, seed=0, dtype="float64", thin=1, keep_samples=True,
            stats="auto", chain_offset=0, layout=None):
        lib = _lib.load()
        tdt, code = D.resolve_dtype(dtype)
        iterations, burn, n_chains, thin = int(iterations), int(burn), int(n_chains), int(thin)
        kp = lib.bmc_padded_components(self.k)
        mode = {"auto": _lib.STATS_FULL if self.k <= 8 else _lib.STATS_DIAG, "full": _lib.STATS_FULL,
                "diag": _lib.STATS_DIAG, "none": _lib.STATS_NONE, None: _lib.STATS_NONE}[stats]
        n_kept = -(-iterations // thin) if keep_samples else 0
        samples = (torch.empty((n_kept, self.k + 1, n_chains), dtype=tdt, device=self.dev)
                   if keep_samples else None)
        n_stat = lib.bmc_gibbs_n_stat(kp, mode)
        cstats = torch.empty((n_stat, n_chains), dtype=torch.float64, device=self.dev) if mode else None
        accepted = torch.empty(n_chains, dtype=torch.int32, device=self.dev)
        prob = self.problem(layout)
        _lib.check(lib.bmc_gibbs_simplex_run(code, C.byref(prob), int(seed) & (2 ** 64 - 1), int(chain_offset),
                                             n_chains, burn, iterations, thin, n_kept, D.ptr(samples),
                                             D.ptr(cstats), mode, D.ptr(accepted), D.stream_ptr(self.dev)),
                   "bmc_gibbs_simplex_run")
        return samples, cstats, accepted, dict(kp=kp, mode=mode, n_kept=n_kept)

    @D.on_own_device
    def summarise(self, cstats, meta, iterations, n_chains):
        if cstats is None or iterations == 0:
            return None, None, None
        kp, mode, k = meta["kp"], meta["mode"], self.k
        total = D.to_host(cstats.sum(dim=1))
        mean_e, cov_e = _moments_from_stats(total, k, kp, mode, float(iterations) * n_chains)
        sig_ref = np.sqrt(self.rss_zero / self.n) if self.rss_zero > 0 else 1.0
        base = np.concatenate([self.b_ols, [sig_ref]])
        comp = list(range(k)) + [kp]
        base_d = torch.from_numpy(base).to(cstats.device)
        chain_mean_d = base_d[None, :] + cstats[comp, :].t() / float(iterations)
        eye = torch.eye(k + 1, dtype=torch.float64, device=cstats.device)
        self.last_rhat, self.last_ess = _chain_diagnostics(cstats, comp, _stat_layout(k, kp, mode)[2], eye,
                                                           chain_mean_d, iterations)
        return base + mean_e, cov_e, D.to_host(chain_mean_d)


def run_gibbs_simplex(y, X, Vt_hat, S_hat, iterations, prior_info, burn=10000, stepsize=0.001, *, n_chains=1,
                      seed=None, dtype="float64", thin=1, keep_samples=True, stats="auto", device=None,
                      chain_offset=0, as_numpy=True, layout=None):
    """Batched simplex-constrained sampler (every chain runs its own burn-in)."""
    # the reference validates after its set-up and before any iteration (:91-94)
    if burn < 0:
        raise ValueError("Burn-in iterations must be non-negative.")
    if stepsize <= 0:
        raise ValueError("Stepsize must be positive.")
    seed = D.fresh_seed() if seed is None else int(seed)
    sampler = SimplexSampler(y, X, Vt_hat, S_hat, prior_info, stepsize, device)
    samples, cstats, accepted, meta = sampler.run(iterations, burn, n_chains, seed, dtype, thin, keep_samples,
                                                  stats, chain_offset, layout)
    mean, cov, chain_mean = sampler.summarise(cstats, meta, int(iterations), int(n_chains))
    with D.on(sampler.dev):
        acc = D.to_host(accepted).astype(np.float64) / max(int(iterations), 1)
        rows = _finish_samples(samples, as_numpy)
    return GibbsResult(samples=rows, mean=mean, cov=cov, chain_mean=chain_mean,
                       n_chains=int(n_chains), iterations=int(iterations), n_kept=meta["n_kept"], seed=seed,
                       dtype=str(dtype), acceptance=acc, info=dict(rss_min=sampler.rss_min, b_ols=sampler.b_ols),
                       rhat=getattr(sampler, "last_rhat", None), ess=getattr(sampler, "last_ess", None))


def gibbs_sampler_simplex(y, X, Vt_hat, S_hat, iterations, prior_info, burn=10000, stepsize=0.001, *,
                          n_chains=1, seed=None, dtype="float64", thin=1, device=None):
    """Gibbs sampling with simplex constraints on the model weights
    (pybmc/inference_utils.py:59-144).

    Same arguments, return value, ``ValueError``s and ``Acceptance rate`` print as upstream; the
    keyword-only options add chains / seed / precision / thinning.
    """
    res = run_gibbs_simplex(y, X, Vt_hat, S_hat, iterations, prior_info, burn, stepsize, n_chains=n_chains,
                            seed=seed, dtype=dtype, thin=thin, stats="none", device=device)
    rate = float(np.mean(res.acceptance)) * 100 if iterations else float("nan")
    print(f"Acceptance rate: {rate:.2f}%")                        # :143
    return res.samples


# --------------------------------------------------------------------------------------------
def gibbs_sampler_literal(y, X, iterations, prior_info, *, n_chains=1, seed=None, dtype="float64",
                          device=None, chain_offset=0):
    """One-chain-per-warp sampler that redoes the reference's per-iteration algebra (K-by-K
    factorisation, residual over all n rows).  Parity anchor for ``gibbs_sampler``; see
    ``bmc_gibbs_literal_run`` in include/bmc_b200.h.  Returns ``[n_chains*iterations, K+1]``."""
    lib = _lib.load()
    dev = D.device(device)
    with D.on(dev):
        return _literal(lib, dev, y, X, iterations, prior_info, n_chains, seed, dtype, chain_offset)


def _literal(lib, dev, y, X, iterations, prior_info, n_chains, seed, dtype, chain_offset):
    tdt, code = D.resolve_dtype(dtype)
    seed = D.fresh_seed() if seed is None else int(seed)
    b0, B0, nu0, sigma20 = prior_info
    yd, Xd = _as_design(y, X, dev)
    n, k = Xd.shape
    lam = np.linalg.inv(np.asarray(B0, dtype=np.float64))
    gram, xty, _ = _gram_with_response(Xd, yd, dev)
    b_ols = np.linalg.inv(gram) @ xty
    sigma2_init = max(_residual_ss(Xd, yd, b_ols, dev) / n, SIGMA2_FLOOR)
    xt = Xd.t().contiguous().to(tdt)
    yr = yd.to(tdt)
    consts = D.to_device(np.concatenate([lam.reshape(-1), lam @ np.asarray(b0, dtype=np.float64)]), dev)
    out = torch.empty((int(iterations), k + 1, int(n_chains)), dtype=tdt, device=dev)
    _lib.check(lib.bmc_gibbs_literal_run(code, D.ptr(xt), D.ptr(yr), n, k, consts.data_ptr(),
                                         consts.data_ptr() + 8 * k * k, float(nu0), float(sigma20), sigma2_init,
                                         seed & (2 ** 64 - 1), int(chain_offset), int(n_chains), int(iterations),
                                         D.ptr(out), D.stream_ptr(dev)), "bmc_gibbs_literal_run")
    return _finish_samples(out, True)
