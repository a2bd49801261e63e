"""pybmc_b200.parallel on real ranks: one process per GPU over NCCL, launched with torch.distributed.run exactly as
the driver launches bench.py.  profiles/check_sharded.py compares every sharded entry point with its single-GPU
result (moments, kept samples, histograms; the broadcast + packed all-gather of the prediction; row-sharded
orthogonalisation incl. the TSQR route and the sampler on all-reduced statistics; `device=` naming a GPU that is
not current).  Skipped with fewer than two GPUs; the merge arithmetic itself also runs over two gloo ranks on
the CPU (test_distributed_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_entry_points_on_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "profiles", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED OK" in r.stdout, r.stdout[-3000:]
