"""Which arithmetic does the tensor-core contraction actually perform?  Compares the device result
with NumPy emulations of 1xTF32 and split-TF32 (3 products)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_predict_tensor import _problem
from pybmc_b200.sampling_utils import PredictiveProblem
from pybmc_b200 import _lib


def tf32(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000            # round to nearest, ties away (cvt.rna)
    return b.astype(np.uint32).view(np.float32)


for k, n, s in [(17, 300, 3000), (64, 700, 2500)]:
    preds, vt, theta, truth = _problem(k, n, s)
    prob = PredictiveProblem(preds, theta, vt, truth=truth, dtype="float32")
    prob.tensor_min_k = 0            # bmc_predict_problem.tensor_min_k: tensor cores for every k (default)
    tc = prob.run(noise="none", return_draws=True).draws
    prob.tensor_min_k = -1           # FFMA kernels
    ff = prob.run(noise="none", return_draws=True).draws
    mu = preds.mean(axis=1)
    u = (preds @ vt.T).astype(np.float32)
    b = theta[:, :k].astype(np.float32)
    exact = b.astype(np.float64) @ u.astype(np.float64).T
    scale = np.abs(b).astype(np.float64) @ np.abs(u).astype(np.float64).T
    uh, bh = tf32(u), tf32(b)
    ul, bl = tf32(u - uh), tf32(b - bh)
    one = bh.astype(np.float64) @ uh.astype(np.float64).T
    three = one + bl.astype(np.float64) @ uh.astype(np.float64).T + bh.astype(np.float64) @ ul.astype(np.float64).T
    for name, arr in (("tensor", tc - mu), ("ffma", ff - mu)):
        print(k, name, "vs exact", np.abs(arr - exact).max() / 1, (np.abs(arr - exact) / scale).max(),
              "vs 1xTF32", (np.abs(arr - one) / scale).max(), "vs 3xTF32", (np.abs(arr - three) / scale).max())
    print(k, "emulated 1x err", (np.abs(one - exact) / scale).max(), "3x err", (np.abs(three - exact) / scale).max())
    bad = np.argwhere(np.abs(tc - mu - exact) / scale > 1e-5)
    print("bad entries", len(bad), "of", tc.size, "first", bad[:10].tolist())
    if len(bad):
        print("bad draws mod 128", np.nonzero(np.bincount(bad[:, 0] % 128, minlength=128))[0].tolist())
        print("bad nuclei mod 128", np.nonzero(np.bincount(bad[:, 1] % 128, minlength=128))[0].tolist())
