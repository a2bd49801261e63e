import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import bench
import pybmc_b200 as pb
preds, truth = bench.config3_ensemble()
o = pb.orthogonalize_arrays(preds, truth, 8)
y, X = o["y"], o["U_hat"]
prior = [np.zeros(8), np.diag(o["S_hat"] ** 2), 1.0, 0.02]
for layout, chains, every, dtype in (("thread", 70, 64, "float64"), ("thread", 64, 64, "float64"), ("thread", 70, 64, "float32"), ("group", 70, 64, "float64")):
    res = pb.run_gibbs(y, X, 400, prior, n_chains=chains, seed=4, hist_every=every, layout=layout, discard=every - 1, thin=every, dtype=dtype)
    s = res.samples.astype(np.float64)
    print(layout, chains, dtype, "hist totals", res.hist.sum(axis=1).tolist(), "expected", chains * (400 // every))
    for c in (0, 8):
        idx = np.clip(np.floor((s[:, c] - res.hist_lo[c]) * (1.0 / res.hist_width[c])), 0, 511).astype(int)
        want = np.bincount(idx, minlength=512)
        d = res.hist[c] - want
        nz = np.flatnonzero(d)
        print("  coord", c, "mismatching bins", len(nz), [(int(i), int(res.hist[c][i]), int(want[i])) for i in nz[:12]])
