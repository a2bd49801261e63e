"""Posterior prediction and coverage with the reference's signatures
(pybmc/sampling_utils.py:4, :40), computed by the fused sm_100a kernels.

``rndm_m_random_calculator`` and ``coverage`` are drop-ins.  ``predictive_summary`` is the
engine behind them: means, variances, exact percentiles and the two order counts per point
that decide every coverage level, without storing the S-by-N matrix unless asked to.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _device as D
from . import _lib

DEFAULT_DRAWS = 10000                      # pybmc/sampling_utils.py:57
DEFAULT_PERCENTILES = (2.5, 50.0, 97.5)    # pybmc/sampling_utils.py:80-82


@dataclass
class PredictiveResult:
    mean: np.ndarray          # [N]
    var: np.ndarray           # [N] population variance of the draws
    percentiles: np.ndarray   # [Q, N]
    c_lt: np.ndarray          # [N] #(draw < truth)   (None without truth)
    c_le: np.ndarray          # [N] #(draw <= truth)
    draws: np.ndarray         # [S, N] when materialised, else None
    n_draws: int
    passes: int
    seed: int


def truth_for(truth, n_points, what="truth"):
    """The first ``n_points`` truth values as a float64 vector.  Upstream indexes ``data_true[i]`` for every
    point (pybmc/sampling_utils.py:26-33) and raises IndexError when the column is shorter; a longer one is
    read only as far as the points go.  Checked on the host: the kernels read truth[n] for every n."""
    t = np.asarray(truth, dtype=np.float64).reshape(-1)
    if t.shape[0] < n_points:
        raise IndexError(f"{what} has {t.shape[0]} entries for {n_points} points")
    return t[:n_points]


def coverage_indices(percentiles, n_draws):
    """Order-statistic indices exactly as the reference computes them
    (pybmc/sampling_utils.py:30-31): Python float arithmetic on each element of ``percentiles``,
    truncated by ``int``.  The truncation is uneven (S=10^4: p=80 -> 999), so it is not re-derived."""
    lo = [int((0.5 - p / 200) * n_draws) for p in percentiles]
    hi = [int((0.5 + p / 200) * n_draws) - 1 for p in percentiles]
    return lo, hi


def coverage_from_counts(percentiles, n_draws, c_lt, c_le, device=None):
    """``sorted[l] <= t <= sorted[u]``  <=>  ``#(x <= t) >= l+1 and #(x < t) <= u``; counted on the
    device by ``bmc_coverage_levels``.  Returns the reference's list of percentages (:35)."""
    lib = _lib.load()
    dev = D.device(device)
    lo, hi = coverage_indices(percentiles, n_draws)
    with D.on(dev):
        c_lt = c_lt if isinstance(c_lt, torch.Tensor) else torch.from_numpy(np.asarray(c_lt, dtype=np.int64))
        c_le = c_le if isinstance(c_le, torch.Tensor) else torch.from_numpy(np.asarray(c_le, dtype=np.int64))
        c_lt, c_le = c_lt.to(dev).contiguous(), c_le.to(dev).contiguous()
        n = int(c_lt.numel())
        if n == 0:
            raise ZeroDivisionError("division by zero")     # upstream divides by the number of points (:35)
        if int(c_le.numel()) != n:
            raise ValueError("c_lt and c_le must have one entry per point")
        lo_d = torch.tensor(lo, dtype=torch.int64, device=dev)
        hi_d = torch.tensor(hi, dtype=torch.int64, device=dev)
        out = torch.empty(len(lo), dtype=torch.int64, device=dev)
        _lib.check(lib.bmc_coverage_levels(D.ptr(c_lt), D.ptr(c_le), n, D.ptr(lo_d), D.ptr(hi_d), len(lo),
                                           D.ptr(out), D.stream_ptr(dev)), "bmc_coverage_levels")
        return [int(c) / n * 100 for c in out.cpu().tolist()]


def coverage(percentiles, rndm_m, models_output, truth_column, *, device=None):
    """Coverage of credible intervals (pybmc/sampling_utils.py:4-37).

    Args:
        percentiles: interval levels in percent (``evaluate`` passes ``np.arange(0, 101, 5)``).
        rndm_m: ``[S, N]`` posterior predictive draws.
        models_output: DataFrame holding the truth column.
        truth_column: its name.

    Returns:
        list[float]: percentage of points whose truth lies within each interval.

    The 21*N sorts of the reference are replaced by two integer counts per point
    (``bmc_coverage_counts``), bit-exact with the sorted form including ties.
    """
    lib = _lib.load()
    dev = D.device(device)
    mat = np.asarray(rndm_m, dtype=np.float64)
    if mat.ndim != 2:
        raise ValueError("rndm_m must be [n_draws, n_points]")
    s_rows, n_cols = mat.shape
    truth = truth_for(models_output[truth_column].tolist(), n_cols, f"column {truth_column!r}")   # :26-33
    with D.on(dev):
        c_lt, c_le = _order_counts_device(lib, dev, mat, truth)
    return coverage_from_counts(percentiles, s_rows, c_lt, c_le, dev)


def _order_counts_device(lib, dev, mat, truth):
    md = D.to_device(mat, dev)
    td = D.to_device(truth, dev)
    s_rows, n_cols = md.shape
    c_lt = torch.empty(n_cols, dtype=torch.int64, device=dev)
    c_le = torch.empty(n_cols, dtype=torch.int64, device=dev)
    _lib.check(lib.bmc_coverage_counts(D.ptr(md), s_rows, n_cols, md.stride(0), D.ptr(td), D.ptr(c_lt), D.ptr(c_le),
                                       D.stream_ptr(dev)), "bmc_coverage_counts")
    return c_lt, c_le


def order_counts(rndm_m, truth, *, device=None):
    """#(x < t), #(x <= t) per column of a materialised matrix (``bmc_coverage_counts``)."""
    lib = _lib.load()
    dev = D.device(device)
    mat = np.asarray(rndm_m, dtype=np.float64)
    if mat.ndim != 2:
        raise ValueError("rndm_m must be [n_draws, n_points]")
    with D.on(dev):
        c_lt, c_le = _order_counts_device(lib, dev, mat, truth_for(truth, mat.shape[1]))
        return c_lt.cpu().numpy(), c_le.cpu().numpy()


def column_percentiles(matrix, percentiles, *, truth=None, device=None):
    """Exact ``np.percentile(matrix, q, axis=0)`` (linear method) of a materialised ``[S, N]`` matrix
    through the same window/select kernels as the fused path ("matrix mode").  Returns a
    ``PredictiveResult`` (draws=None)."""
    lib = _lib.load()
    dev = D.device(device)
    mat = np.asarray(matrix, dtype=np.float64)
    if mat.ndim != 2:
        raise ValueError("matrix must be [n_draws, n_points]")
    with D.on(dev):
        md = D.to_device(mat, dev)
        s_rows, n_cols = md.shape
        center = torch.empty(n_cols, dtype=torch.float64, device=dev)
        scale = torch.empty(n_cols, dtype=torch.float64, device=dev)
        _lib.check(lib.bmc_column_moments(D.ptr(md), s_rows, n_cols, md.stride(0), D.ptr(center), D.ptr(scale),
                                          D.stream_ptr(dev)), "bmc_column_moments")
        td = D.to_device(truth_for(truth, n_cols), dev) if truth is not None else None
        return _launch_fused(dev, "float64", n_points=n_cols, point0=0, n_draws=s_rows, k=0, u=None, mu=None,
                             truth=td, theta=None, noise_mode=_lib.NOISE_EXTERNAL, seed=0, noise=md,
                             ld_noise=md.stride(0), percentiles=percentiles, theta_mean=None, theta_cov=None,
                             center=center, scale=scale, return_draws=False)


def _launch_fused(dev, dtype, *, n_points, point0, n_draws, k, u, mu, truth, theta, noise_mode, seed, noise,
                  ld_noise, percentiles, theta_mean, theta_cov, center, scale, return_draws, as_numpy=True,
                  workspace=None, tensor_min_k=0):
    lib = _lib.load()
    _, code = D.resolve_dtype(dtype)
    probs = np.asarray(percentiles, dtype=np.float64).reshape(-1)
    nq_total = probs.shape[0]
    if nq_total < 1:
        raise ValueError("at least one percentile is required")
    if np.any(probs < 0) or np.any(probs > 100):
        raise ValueError("Percentiles must be in the range [0, 100]")
    if n_points == 0:      # nothing to predict: empty outputs of the right shapes, as NumPy would give
        empty = np.zeros(0)
        return PredictiveResult(mean=empty, var=empty, percentiles=np.zeros((nq_total, 0)),
                                c_lt=None if truth is None else np.zeros(0, dtype=np.int64),
                                c_le=None if truth is None else np.zeros(0, dtype=np.int64),
                                draws=np.zeros((n_draws, 0)) if return_draws else None, n_draws=n_draws, passes=0,
                                seed=int(seed))
    # one fp64 block [mean | var | quantiles] and one int64 block [c_lt | c_le]: two device-to-host copies
    outf = torch.empty((2 + nq_total, n_points), dtype=torch.float64, device=dev)
    mean, var, quant = outf[0], outf[1], outf[2:]
    outi = torch.empty((2, n_points), dtype=torch.int64, device=dev) if truth is not None else None
    c_lt = outi[0] if truth is not None else None
    c_le = outi[1] if truth is not None else None
    draws = torch.empty((n_draws, n_points), dtype=torch.float64, device=dev) if return_draws else None
    passes = 0
    # the kernel takes up to MAX_QUANTILES percentiles per call; longer lists go in batches
    for q0 in range(0, nq_total, _lib.MAX_QUANTILES):
        batch = np.ascontiguousarray(probs[q0:q0 + _lib.MAX_QUANTILES])
        nq = batch.shape[0]
        nbytes = int(lib.bmc_predict_workspace_bytes(code, n_points, nq, n_draws))
        ws = workspace if workspace is not None and workspace.numel() >= nbytes else torch.empty(
            nbytes, dtype=torch.uint8, device=dev)
        prob = _lib.PredictProblem(
            n_points=n_points, point0=point0, n_draws=n_draws, k=k, u=D.ptr(u), mu=D.ptr(mu), truth=D.ptr(truth),
            theta=D.ptr(theta), noise_mode=noise_mode, seed=int(seed) & (2 ** 64 - 1), noise=D.ptr(noise),
            ld_noise=ld_noise, nq=nq, probs=batch.ctypes.data_as(C.POINTER(C.c_double)),
            theta_mean=D.ptr(theta_mean), theta_cov=D.ptr(theta_cov), center=D.ptr(center), scale=D.ptr(scale),
            tensor_min_k=int(tensor_min_k))
        n_pass = C.c_int(0)
        first = q0 == 0
        _lib.check(lib.bmc_predict_fused(code, C.byref(prob), D.ptr(mean), D.ptr(var), quant[q0:].data_ptr(),
                                         D.ptr(c_lt), D.ptr(c_le), D.ptr(draws) if first else None,
                                         n_points, D.ptr(ws), ws.numel(), C.byref(n_pass), D.stream_ptr(dev)),
                   "bmc_predict_fused")
        passes = max(passes, n_pass.value)
    if as_numpy:
        outf = D.to_host(outf)
        outi = None if outi is None else D.to_host(outi)
        draws = None if draws is None else D.to_host(draws)
    return PredictiveResult(mean=outf[0], var=outf[1], percentiles=outf[2:],
                            c_lt=None if outi is None else outi[0], c_le=None if outi is None else outi[1],
                            draws=draws, n_draws=n_draws, passes=passes, seed=int(seed))


class PredictiveProblem:
    """Device-resident inputs of the fused kernel: K-space coordinates of the points and the
    transposed posterior draws.  ``run`` only launches kernels (what bench.py times)."""

    def __init__(self, preds, theta, Vt_hat, truth=None, dtype="float64", device=None, point0=0, tensor_min_k=0):
        self.dev = D.device(device)
        self.dtype = dtype
        self.tensor_min_k = int(tensor_min_k)      # bmc_predict_problem.tensor_min_k (0 default, < 0 FFMA kernels)
        with D.on(self.dev):
            self._setup(preds, theta, Vt_hat, truth, point0)

    def _setup(self, preds, theta, Vt_hat, truth, point0):
        lib = _lib.load()
        tdt, _ = D.resolve_dtype(self.dtype)
        Vt_hat = np.asarray(Vt_hat, dtype=np.float64)
        preds = np.asarray(preds, dtype=np.float64)
        if preds.ndim != 2 or Vt_hat.ndim != 2 or preds.shape[1] != Vt_hat.shape[1]:
            raise ValueError(f"shapes do not align: predictions {preds.shape}, Vt_hat {Vt_hat.shape}")
        self.k, self.m = Vt_hat.shape
        if self.k > _lib.MAX_COMPONENTS:
            raise ValueError(f"at most {_lib.MAX_COMPONENTS} components are supported")
        self.n_points = preds.shape[0]
        self.point0 = int(point0)
        if truth is not None:
            truth = truth_for(truth, self.n_points)
        if self.n_points == 0:
            self.mu = torch.zeros(0, dtype=torch.float64, device=self.dev)
            self.u = torch.zeros((0, self.k), dtype=tdt, device=self.dev)
            self.truth = None if truth is None else torch.zeros(0, dtype=torch.float64, device=self.dev)
            self.set_draws(theta)
            return
        pd_ = D.to_device(preds, self.dev)
        vd = D.to_device(Vt_hat, self.dev)
        # mu = mean over models: the 1/M default weights of :64;  u = preds Vt_hat' (:64-72 in K-space)
        self.mu = torch.empty(self.n_points, dtype=torch.float64, device=self.dev)
        _lib.check(lib.bmc_center_rows(D.ptr(pd_), self.n_points, self.m, pd_.stride(0), None, D.ptr(self.mu), None,
                                       None, 0, D.stream_ptr(self.dev)), "bmc_center_rows")
        u64 = torch.empty((self.n_points, self.k), dtype=torch.float64, device=self.dev)
        _lib.check(lib.bmc_project_rows(D.ptr(pd_), self.n_points, self.m, pd_.stride(0), None, D.ptr(vd), self.k,
                                        D.ptr(u64), self.k, D.stream_ptr(self.dev)), "bmc_project_rows")
        self.u = u64.to(tdt).contiguous()
        self.truth = D.to_device(truth, self.dev) if truth is not None else None
        self.set_draws(theta)

    @D.on_own_device
    def set_draws(self, theta):
        """theta: [S, K+1] rows [beta, sigma] (host array or device tensor)."""
        tdt, _ = D.resolve_dtype(self.dtype)
        th = theta if isinstance(theta, torch.Tensor) else D.to_device(np.asarray(theta, dtype=np.float64), self.dev)
        th = th.to(device=self.dev, dtype=torch.float64)
        if th.ndim != 2 or th.shape[1] != self.k + 1:
            raise ValueError(f"samples must be [S, {self.k + 1}], got {tuple(th.shape)}")
        self.n_draws = th.shape[0]
        self.theta_mean = th.mean(dim=0).contiguous()
        cen = (th - self.theta_mean).contiguous()
        # covariance of the draws by the library's own Gram kernel (no cuBLAS handle for a (K+1)^2 result)
        lib = _lib.load()
        cov = torch.zeros((self.k + 1, self.k + 1), dtype=torch.float64, device=self.dev)
        if self.n_draws > 0:
            ws = torch.empty(max(int(lib.bmc_gram_workspace_bytes(self.n_draws, self.k + 1)), 8), dtype=torch.uint8,
                             device=self.dev)
            _lib.check(lib.bmc_gram(D.ptr(cen), self.n_draws, self.k + 1, cen.stride(0), None, None, D.ptr(cov),
                                    D.ptr(ws), ws.numel(), D.stream_ptr(self.dev)), "bmc_gram")
        self.theta_cov = (cov / max(self.n_draws, 1)).contiguous()
        # rows padded for 16-byte broadcast loads: beta | zeros | sigma at column stride-4 | zeros
        stride = _lib.load().bmc_predict_theta_stride(self.k)
        padded = torch.zeros((self.n_draws, stride), dtype=tdt, device=self.dev)
        padded[:, : self.k] = th[:, : self.k].to(tdt)
        padded[:, stride - 4] = th[:, self.k].to(tdt)
        self.theta = padded

    @D.on_own_device
    def run(self, percentiles=DEFAULT_PERCENTILES, noise="philox", seed=0, return_draws=False, as_numpy=True,
            workspace=None):
        noise_t, ld = None, 0
        if isinstance(noise, str):
            mode = {"philox": _lib.NOISE_PHILOX, "none": _lib.NOISE_NONE}[noise]
        else:
            tdt, _ = D.resolve_dtype(self.dtype)
            noise_t = D.to_device(noise, self.dev, tdt)
            if noise_t.shape != (self.n_draws, self.n_points):
                raise ValueError("external noise must be [n_draws, n_points]")
            mode, ld = _lib.NOISE_EXTERNAL, noise_t.stride(0)
        return _launch_fused(self.dev, self.dtype, n_points=self.n_points, point0=self.point0,
                             n_draws=self.n_draws, k=self.k, u=self.u, mu=self.mu, truth=self.truth,
                             theta=self.theta, noise_mode=mode, seed=seed, noise=noise_t, ld_noise=ld,
                             percentiles=percentiles, theta_mean=self.theta_mean, theta_cov=self.theta_cov,
                             center=None, scale=None, return_draws=return_draws, as_numpy=as_numpy,
                             workspace=workspace, tensor_min_k=self.tensor_min_k)


def select_draws(samples, n_draws, rng):
    """``rng.choice(samples, n_draws, replace=False)`` (pybmc/sampling_utils.py:57): raises NumPy's
    ValueError when fewer than ``n_draws`` rows are available, as upstream."""
    samples = np.asarray(samples)
    idx = rng.choice(samples.shape[0], n_draws, replace=False)
    return samples[idx]


def predictive_summary(filtered_model_predictions, samples, Vt_hat, *, truth=None, n_draws=DEFAULT_DRAWS,
                       percentiles=DEFAULT_PERCENTILES, seed=None, dtype="float64", noise="philox",
                       return_draws=False, subsample=True, device=None, point0=0, tensor_min_k=0):
    """Fused prediction + UQ for ``N`` points (see module docstring).

    ``subsample=True`` draws ``n_draws`` posterior rows without replacement like the reference;
    ``subsample=False`` uses ``samples`` as given (``n_draws`` is then its length).
    """
    seed = D.fresh_seed() if seed is None else int(seed)
    if subsample:
        theta = select_draws(samples, int(n_draws), np.random.default_rng(seed))
    else:
        theta = np.asarray(samples, dtype=np.float64)
    prob = PredictiveProblem(filtered_model_predictions, theta, Vt_hat, truth=truth, dtype=dtype, device=device,
                             point0=point0, tensor_min_k=tensor_min_k)
    return prob.run(percentiles=percentiles, noise=noise, seed=seed, return_draws=return_draws)


def rndm_m_random_calculator(filtered_model_predictions, samples, Vt_hat, *, n_draws=DEFAULT_DRAWS, seed=None,
                             dtype="float64", return_draws=True, device=None):
    """Posterior predictive samples and credible intervals (pybmc/sampling_utils.py:40-84).

    Returns ``(rndm_m, [lower, median, upper])`` with ``rndm_m`` the ``[10000, N]`` float64 draws and
    the 2.5 / 50 / 97.5 percentiles per point.  ``return_draws=False`` skips materialising
    ``rndm_m`` (returned as None) -- the only way to run sizes where it would not fit.
    """
    np.random.seed(142858)     # side effect kept from :54 (it never influenced the draws themselves)
    res = predictive_summary(filtered_model_predictions, samples, Vt_hat, n_draws=n_draws, seed=seed, dtype=dtype,
                             return_draws=return_draws, device=device)
    lower, median, upper = res.percentiles
    return res.draws, [lower, median, upper]
