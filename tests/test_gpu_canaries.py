"""Out-of-bounds guard: compute-sanitizer is closed on this GPU pool, so the entry points are run
on buffers with sentinel-filled guard bands on both sides (ragged sizes: chains, nuclei and draws that
are not multiples of the block / tile sizes) and the bands must come back untouched."""
import ctypes as C

import numpy as np
import pytest
import torch

import cases
from oracle import bmc_oracle as oc

pytestmark = pytest.mark.gpu
GUARD = 4096


def guarded(shape, dtype, dev, fill):
    n = int(np.prod(shape))
    raw = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device=dev)
    return raw, raw[GUARD:GUARD + n].view(*shape)


def intact(raw, n, fill):
    lo, hi = raw[:GUARD], raw[GUARD + n:]
    return bool((lo == fill).all() and (hi == fill).all())


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n_chains", [1, 37, 16385])       # group layout (ragged) and thread layout (ragged)
def test_sampler_buffers(dtype, n_chains):
    from pybmc_b200 import _device as D, _lib
    from pybmc_b200.inference_utils import ConjugateSampler, SimplexSampler
    lib = _lib.load()
    r = oc.orthogonalize_arrays(*cases.ensemble(11, 28, 5), 3)
    dev = D.device()
    tdt, code = D.resolve_dtype(dtype)
    k, iters, thin = 3, 70, 7
    n_kept = -(-iters // thin)
    kp = lib.bmc_padded_components(k)
    n_stat = lib.bmc_gibbs_n_stat(kp, _lib.STATS_FULL)
    cs = ConjugateSampler(r["y"], r["U_hat"], [np.zeros(3), np.diag(r["S_hat"] ** 2), 1.0, 0.02])
    raw_s, samples = guarded((n_kept, k + 1, n_chains), tdt, dev, -777.0)
    raw_c, stats = guarded((n_stat, n_chains), torch.float64, dev, -777.0)
    prob = cs.problem()
    _lib.check(lib.bmc_gibbs_run(code, C.byref(prob), 5, 0, n_chains, iters, 0, thin, n_kept, samples.data_ptr(),
                                 stats.data_ptr(), _lib.STATS_FULL, None, D.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert intact(raw_s, samples.numel(), -777.0) and intact(raw_c, stats.numel(), -777.0)
    assert bool(torch.isfinite(samples).all()) and bool((samples[:, k, :] > 0).all())
    ss = SimplexSampler(r["y"], r["U_hat"], r["Vt_hat"], r["S_hat"], [1.0, 0.02], 0.02)
    raw_s, samples = guarded((n_kept, k + 1, n_chains), tdt, dev, -777.0)
    raw_c, stats = guarded((n_stat, n_chains), torch.float64, dev, -777.0)
    raw_a, acc = guarded((n_chains,), torch.int32, dev, -7)
    sp = ss.problem()
    _lib.check(lib.bmc_gibbs_simplex_run(code, C.byref(sp), 5, 0, n_chains, 33, iters, thin, n_kept,
                                         samples.data_ptr(), stats.data_ptr(), _lib.STATS_FULL, acc.data_ptr(),
                                         D.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert intact(raw_s, samples.numel(), -777.0) and intact(raw_c, stats.numel(), -777.0) and intact(raw_a, n_chains, -7)
    assert bool(torch.isfinite(samples).all()) and bool(((acc >= 0) & (acc <= iters)).all())


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n_points,n_draws", [(1, 5), (257, 1001), (1030, 4099)])
def test_predict_buffers(dtype, n_points, n_draws):
    from pybmc_b200 import _device as D, _lib
    from pybmc_b200.sampling_utils import PredictiveProblem
    lib = _lib.load()
    rng = np.random.default_rng(n_points)
    k, m = 5, 9
    preds = 50 + rng.normal(size=(n_points, m))
    vt = rng.normal(size=(k, m)) * 0.1
    theta = np.column_stack([rng.normal(size=(n_draws, k)), np.abs(rng.normal(0.3, 0.05, n_draws))])
    truth = preds.mean(axis=1) + rng.normal(0, 0.5, n_points)
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype=dtype)
    dev = prob.dev
    _, code = D.resolve_dtype(dtype)
    q = np.array([2.5, 50.0, 97.5])
    raw_m, mean = guarded((n_points,), torch.float64, dev, -777.0)
    raw_v, var = guarded((n_points,), torch.float64, dev, -777.0)
    raw_q, quant = guarded((3, n_points), torch.float64, dev, -777.0)
    raw_l, c_lt = guarded((n_points,), torch.int64, dev, -7)
    raw_e, c_le = guarded((n_points,), torch.int64, dev, -7)
    raw_d, draws = guarded((n_draws, n_points), torch.float64, dev, -777.0)
    nbytes = int(lib.bmc_predict_workspace_bytes(code, n_points, 3, n_draws))
    raw_w = torch.full((nbytes + 2 * GUARD,), 0x5A, dtype=torch.uint8, device=dev)
    ws = raw_w[GUARD:GUARD + nbytes]
    pp = _lib.PredictProblem(n_points=n_points, point0=0, n_draws=n_draws, k=k, u=prob.u.data_ptr(),
                             mu=prob.mu.data_ptr(), truth=prob.truth.data_ptr(), theta=prob.theta.data_ptr(),
                             noise_mode=_lib.NOISE_PHILOX, seed=3, noise=None, ld_noise=0, nq=3,
                             probs=q.ctypes.data_as(C.POINTER(C.c_double)), theta_mean=prob.theta_mean.data_ptr(),
                             theta_cov=prob.theta_cov.data_ptr(), center=None, scale=None)
    passes = C.c_int(0)
    _lib.check(lib.bmc_predict_fused(code, C.byref(pp), mean.data_ptr(), var.data_ptr(), quant.data_ptr(),
                                     c_lt.data_ptr(), c_le.data_ptr(), draws.data_ptr(), n_points, ws.data_ptr(),
                                     nbytes, C.byref(passes), D.stream_ptr(dev)))
    torch.cuda.synchronize()
    for raw, n, fill in ((raw_m, n_points, -777.0), (raw_v, n_points, -777.0), (raw_q, 3 * n_points, -777.0),
                         (raw_l, n_points, -7), (raw_e, n_points, -7), (raw_d, n_draws * n_points, -777.0)):
        assert intact(raw, n, fill)
    assert bool((raw_w[:GUARD] == 0x5A).all()) and bool((raw_w[GUARD + nbytes:] == 0x5A).all())
    d = draws.cpu().numpy()
    np.testing.assert_allclose(quant.cpu().numpy(), np.percentile(d, q, axis=0), rtol=1e-12 if dtype == "float64" else 1e-6)
    assert np.array_equal(c_lt.cpu().numpy(), (d < truth[None, :]).sum(axis=0))
