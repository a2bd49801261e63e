"""Time the fused prediction of BASELINE configs[3] (nuclei x draws x K = 16, 5 percentiles + coverage counts,
no S x N matrix) alone.  usage: python profiles/time_predict.py [n_points] [n_draws] [k] [reps]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
import bench
from pybmc_b200 import _lib
if os.environ.get("BMC_LIB"):          # A/B runs: a differently built library (profiles/ab/*.so, not tracked)
    _lib.LIB_PATH = os.path.abspath(os.environ["BMC_LIB"])
from pybmc_b200.sampling_utils import PredictiveProblem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
s = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
lib = _lib.load()
preds, vt, theta, truth = bench.config4_inputs(n, s, k)
prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32")
ws = torch.empty(int(lib.bmc_predict_workspace_bytes(_lib.F32, n, 5, s)), dtype=torch.uint8, device="cuda")
q = [2.5, 16.0, 50.0, 84.0, 97.5]
ms = []
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = prob.run(percentiles=q, seed=0xB204, as_numpy=False, workspace=ws)
    e1.record()
    e1.synchronize()
    ms.append(e0.elapsed_time(e1))
best = min(ms[1:]) if reps > 1 else ms[0]
print(f"{n} x {s} x K={k}: ms {[round(m, 2) for m in ms]}; {n * s / (best * 1e-3):.4g} samples x points/s; passes {res.passes}")
qs = res.quantiles if hasattr(res, "quantiles") else None
print("mean[:3]", res.mean[:3].tolist() if hasattr(res.mean, "tolist") else res.mean[:3])
