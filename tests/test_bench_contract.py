"""bench.py on the CPU: the reference arm prints the contract's JSON line; the synthetic inputs are
deterministic.  (The native arm needs a GPU and is exercised by the driver.)"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gibbs_chain_iters_per_sec"
    assert line["unit"] == "chain-iters/s" and line["higher_is_better"] is True and line["value"] > 1e3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("BASELINE configs[2]")


def test_synthetic_inputs_are_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    a, b = bench.config3_ensemble(), bench.config3_ensemble()
    assert a[0].shape == (3000, 16) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    p, t = bench.config1_ensemble()
    assert p.shape == (629, 15) and t.shape == (629,)
    preds, vt, theta, truth = bench.config4_inputs(100, 50)
    assert preds.shape == (100, 24) and vt.shape == (16, 24) and theta.shape == (50, 17) and truth.shape == (100,)


def test_native_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
