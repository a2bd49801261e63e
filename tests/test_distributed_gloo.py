"""World-size-2 checks of the sharding / merge logic on the CPU (gloo).  The per-shard numbers come
from the oracle (the kernels need a GPU); what is tested is that shards keyed by global ids merge
into exactly what one rank holding everything computes."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _chains(first, last, iters):
    """Oracle chains with global ids [first, last) on the device's Philox stream."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    from oracle import bmc_oracle as oc
    from oracle import philox as px
    y, X = cases.toy_regression()
    prior = (np.array([0.3, -0.2]), np.array([[2.0, 0.3], [0.3, 0.5]]), 2.5, 0.7)
    return [oc.gibbs_conjugate(y, np.asarray(X, float), iters, prior, oc.PhiloxDraws(11, c, px.TAG_GIBBS))
            for c in range(first, last)]


def _table():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    return cases.ensemble(11, 41, 6)


def _sums(samples):
    s = np.concatenate(samples)
    d = s.shape[1]
    second = [np.sum(s[:, r] * s[:, c]) for r in range(d) for c in range(r, d)]
    return np.concatenate([s.sum(axis=0), second]), len(s)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from pybmc_b200 import parallel as par
    n_chains, iters, n_points = 5, 12, 10
    lo, hi = par.chain_range(n_chains)
    sums, count = _sums(_chains(lo, hi, iters))
    total, n = par.merge_moment_sums(torch.from_numpy(sums), count)
    plo, phi = par.point_range(n_points)
    local = torch.arange(plo, phi, dtype=torch.float64).repeat(3, 1) * 1.5
    gathered = par.gather_points(local, n_points)
    covered = par.merge_coverage_counts(torch.tensor([phi - plo, rank + 1], dtype=torch.int64))
    # rows of the prediction table sharded (SURVEY.md 8e, configs[4]): partial Gram of the row-centred
    # matrix, all-reduced by the callable the device path hands to orthogonalize_arrays / ConjugateSampler
    preds, truth = _table()
    rlo, rhi = par.row_range(len(truth))
    xc = preds[rlo:rhi] - preds[rlo:rhi].mean(axis=1, keepdims=True)
    gram = par.sum_over_ranks()(torch.from_numpy(xc.T @ xc))
    if rank == 0:
        torch.save(dict(total=total, n=n, gathered=gathered, covered=covered, ranges=(lo, hi, plo, phi),
                        gram=gram, rows=(rlo, rhi)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    sys.path.insert(0, ROOT)
    from pybmc_b200 import parallel as par
    want, n = _sums(_chains(0, 5, 12))
    assert got["n"] == n == 60
    np.testing.assert_allclose(got["total"].numpy(), want, rtol=1e-13)
    mean, cov = par.posterior_from_sums(got["total"].numpy(), got["n"], 2)
    allc = np.concatenate(_chains(0, 5, 12))
    np.testing.assert_allclose(mean, allc.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(cov, np.cov(allc.T, ddof=0), rtol=1e-9, atol=1e-12)
    assert torch.equal(got["gathered"], torch.arange(10, dtype=torch.float64).repeat(3, 1) * 1.5)
    assert got["covered"].tolist() == [10, 3]
    assert got["ranges"] == (0, 3, 0, 8)
    # the all-reduced Gram matrix gives the singular values / right vectors of the whole table
    from oracle import bmc_oracle as oc
    preds, truth = _table()
    assert got["rows"] == (0, 21)
    ref = oc.orthogonalize_arrays(preds, truth, 3, full_matrices=False)
    lam, vec = np.linalg.eigh(got["gram"].numpy())
    np.testing.assert_allclose(np.sqrt(lam[::-1][:3]), ref["S_hat"], rtol=1e-10)
    for i in range(3):
        v = vec[:, ::-1][:, i]
        np.testing.assert_allclose(abs(v @ ref["Vt_hat_normalized"][i]), 1.0, rtol=1e-10)


def test_ranges_cover_everything_once():
    sys.path.insert(0, ROOT)
    from pybmc_b200 import parallel as par
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 8, 65536, 100000):
            chains = [par.chain_range(total, r, world) for r in range(world)]
            assert chains[0][0] == 0 and chains[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(chains, chains[1:]))
            pts = [par.point_range(total, r, world) for r in range(world)]
            assert pts[0][0] == 0 and max(p[1] for p in pts) == total
            assert all(p[0] % 4 == 0 or p[0] == p[1] for p in pts)
            assert all(a[1] == b[0] or b[0] == b[1] == total for a, b in zip(pts, pts[1:]))
