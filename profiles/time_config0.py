"""Wall-clock of the reference's own workflow (BASELINE configs[0], surrogate data) through the drop-in class."""
import sys, time, io, contextlib
import numpy as np, pandas as pd, torch
sys.path.insert(0, "/root/repo")
import bench, pybmc_b200 as pb

preds, truth = bench.config1_ensemble()
models = [f"m{i}" for i in range(preds.shape[1])]
df = pd.DataFrame(preds, columns=models)
df["N"] = np.arange(len(truth)); df["Z"] = np.arange(len(truth)) // 3
df["truth"] = truth
rng = np.random.default_rng(1)
train = df.iloc[rng.permutation(len(df))[:377]]
bmc = pb.BayesianModelCombination(models, {"BE": df}, "truth")
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return time.perf_counter() - t0, r
with contextlib.redirect_stdout(io.StringIO()):
    t(lambda: bmc.orthogonalize("BE", train, 3)); t(lambda: bmc.train({"iterations": 20000}))   # warm-up
    res = {}
    res["orthogonalize"] = t(lambda: bmc.orthogonalize("BE", train, 3))[0]
    res["train_50k_1chain_f64"] = t(lambda: bmc.train({"iterations": 50000}))[0]
    res["train_50k_1chain_f32"] = t(lambda: bmc.train({"iterations": 50000, "dtype": "float32"}))[0]
    res["train_50k_64chains_f64"] = t(lambda: bmc.train({"iterations": 50000, "n_chains": 64, "thin": 64}))[0]
    bmc.train({"iterations": 50000})
    res["predict2_first_call"] = t(lambda: bmc.predict2("BE"))[0]
    res["predict2"] = t(lambda: bmc.predict2("BE"))[0]
    res["predict2_nodraws"] = t(lambda: bmc.predict2("BE", return_draws=False))[0]
    res["evaluate"] = t(lambda: bmc.evaluate())[0]
    res["train_simplex_10k+50k_1chain"] = t(lambda: bmc.train({"iterations": 50000, "sampler": "simplex"}))[0]
print({k: round(v * 1e3, 2) for k, v in res.items()}, "ms")
