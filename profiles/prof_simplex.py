import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench, pybmc_b200 as pb
from pybmc_b200.inference_utils import SimplexSampler
preds, truth = bench.config1_ensemble()
idx = np.random.default_rng(1).permutation(len(truth))[:377]
o = pb.orthogonalize_arrays(preds[idx], truth[idx], 3)
s = SimplexSampler(o["y"], o["U_hat"], o["Vt_hat"], o["S_hat"], [1.0, 0.02], 0.001)
for _ in range(2):
    s.run(5000, 1000, 4096, 7, "float32", 500, True, "full", 0)
torch.cuda.synchronize()
