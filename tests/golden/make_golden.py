"""Generate tests/golden/*.npz by running the UNMODIFIED upstream pyBMC.

Run in the authoring container only (the reference tree is not on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

The reference draws its variance and predictive noise from
``numpy.random.default_rng()`` with no seed (pybmc/inference_utils.py:52,117,140;
pybmc/sampling_utils.py:55), so the only way to obtain repeatable outputs from the
reference's own code is to hand it a seeded factory in place of
``numpy.random.default_rng`` for the duration of the call.  Nothing in the
reference is edited; the legacy global stream it also uses is seeded with
``numpy.random.seed``.
"""
import contextlib
import io
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

REF = next((a for a in sys.argv[1:] if not a.startswith("--")), "/root/reference")
sys.path.insert(0, REF)
from pybmc.bmc import BayesianModelCombination  # noqa: E402
from pybmc.inference_utils import (  # noqa: E402
    USVt_hat_extraction, gibbs_sampler, gibbs_sampler_simplex)
from pybmc.sampling_utils import coverage, rndm_m_random_calculator  # noqa: E402


@contextlib.contextmanager
def seeded(legacy_seed, factory_base):
    real = np.random.default_rng
    fac = cases.SeededFactory(factory_base)
    np.random.default_rng = fac
    np.random.seed(legacy_seed)
    try:
        yield fac
    finally:
        np.random.default_rng = real


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes")


def main():
    rng = np.random.default_rng(5)

    # -- USVt_hat_extraction (inference_utils.py:147) ---------------------------
    U = rng.normal(size=(5, 5)); S = np.array([4.0, 2.0, 1.0, 0.5]); Vt = rng.normal(size=(4, 4))
    uh, sh, vh, vn = USVt_hat_extraction(U, S, Vt, 2)
    save("usvt", U=U, S=S, Vt=Vt, U_hat=uh, S_hat=sh, Vt_hat=vh, Vt_norm=vn)

    # -- orthogonalize (bmc.py:79) ------------------------------------------------
    df = cases.toy_frame()
    bmc = BayesianModelCombination(["model1", "model2", "model3", "truth"], {"target": df}, "truth")
    bmc.orthogonalize("target", df.iloc[:4], 2)
    save("orth_toy", y=bmc.centered_experiment_train, U_hat=bmc.U_hat, S_hat=bmc.S_hat,
         Vt_hat=bmc.Vt_hat, Vt_norm=bmc.Vt_hat_normalized, mu=bmc._predictions_mean_train)

    frame, models = cases.ensemble_frame(11, 40, 5)
    ens = BayesianModelCombination(models, {"BE": frame}, "truth")
    train = frame.iloc[:28]
    ens.orthogonalize("BE", train, 3)
    save("orth_ens", y=ens.centered_experiment_train, U_hat=ens.U_hat, S_hat=ens.S_hat,
         Vt_hat=ens.Vt_hat, Vt_norm=ens.Vt_hat_normalized, mu=ens._predictions_mean_train,
         insum=cases.checksum(frame[models].values, frame["truth"].values))

    # -- gibbs_sampler (inference_utils.py:4) ---------------------------------------
    y, X = cases.toy_regression()
    with seeded(7, 1000):
        s = gibbs_sampler(y, X, 60, (np.array([0.0, 0.0]), np.eye(2), 1.0, 1.0))
    save("gibbs_toy", samples=s)
    with seeded(17, 1100):
        s = gibbs_sampler(y, X, 40, (np.array([0.3, -0.2]), np.array([[2.0, 0.3], [0.3, 0.5]]), 2.5, 0.7))
    save("gibbs_toy_dense_prior", samples=s)
    with seeded(9, 1200):
        s = gibbs_sampler(ens.centered_experiment_train, ens.U_hat, 300,
                          [np.zeros(3), np.diag(ens.S_hat ** 2), 1.0, 0.02])
    save("gibbs_ens", samples=s)

    # -- gibbs_sampler_simplex (inference_utils.py:59) ---------------------------------
    y, X, Vt_hat, S_hat = cases.toy_simplex()
    out = io.StringIO()
    with seeded(8, 2000), contextlib.redirect_stdout(out):
        s = gibbs_sampler_simplex(y, X, Vt_hat, S_hat, 40, [1.0, 1.0], burn=100, stepsize=0.01)
    acc = float(re.search(r"([0-9.]+)%", out.getvalue()).group(1))
    save("simplex_toy", samples=s, acceptance_pct=acc)
    out = io.StringIO()
    with seeded(10, 2100), contextlib.redirect_stdout(out):
        s = gibbs_sampler_simplex(ens.centered_experiment_train, ens.U_hat, ens.Vt_hat, ens.S_hat,
                                  250, [1.0, 0.02], burn=300, stepsize=0.02)
    acc = float(re.search(r"([0-9.]+)%", out.getvalue()).group(1))
    save("simplex_ens", samples=s, acceptance_pct=acc)

    # -- rndm_m_random_calculator + coverage (sampling_utils.py:40, :4) ------------------
    preds, truth = cases.ensemble(12, 7, 5)
    theta = cases.posterior_like(13, 12000, 3)
    with seeded(1, 3000):
        rndm_m, (lo, med, hi) = rndm_m_random_calculator(preds, theta, ens.Vt_hat)
    import pandas as pd
    tdf = pd.DataFrame({"truth": truth})
    cov = coverage(np.arange(0, 101, 5), rndm_m, tdf, "truth")
    save("predict", lo=lo, med=med, hi=hi, head=rndm_m[:64], colsum=rndm_m.sum(axis=0),
         total=cases.checksum(rndm_m), coverage=np.array(cov), Vt_hat=ens.Vt_hat,
         insum=cases.checksum(preds, truth, theta))

    # coverage with ties and the uneven index truncation (sampling_utils.py:30-33)
    g = np.random.default_rng(21)
    mat = g.integers(0, 12, size=(200, 6)).astype(float)
    tr = np.array([5.0, 0.0, 11.0, 5.5, -1.0, 6.0])
    cov_ties = coverage(np.arange(0, 101, 5), mat, pd.DataFrame({"t": tr}), "t")
    cov_odd = coverage([1, 33, 68, 95, 99], mat[:137], pd.DataFrame({"t": tr}), "t")
    save("coverage_ties", matrix=mat, truth=tr, coverage=np.array(cov_ties), coverage_odd=np.array(cov_odd))

    # -- inside_to_outside split (data.py:194-245) -------------------------------------------------
    from pybmc.data import Dataset
    g2 = np.random.default_rng(31)
    pts = [tuple(map(float, p)) for p in g2.integers(0, 40, size=(300, 2))]
    stable = [tuple(map(float, p)) for p in g2.integers(10, 30, size=(12, 2))]
    tr, va, te = Dataset().separate_points_distance_allSets(pts, stable, 3.0, 7.5)
    save("split_distance", points=np.array(pts), stable=np.array(stable), train=np.array(tr), val=np.array(va),
         test=np.array(te))

    # -- whole class pipeline (bmc.py:79-376) -----------------------------------------------
    pipe = BayesianModelCombination(models, {"BE": frame}, "truth")
    pipe.orthogonalize("BE", train, 3)
    with seeded(3, 4000), contextlib.redirect_stdout(io.StringIO()):
        pipe.train({"iterations": 10500, "sampler": "gibbs_sampling"})
        _, lo_df, med_df, up_df = pipe.predict2("BE")
        cov = pipe.evaluate()
        cov_f = pipe.evaluate({"N": (10, 30)})
    save("pipeline", samples_head=pipe.samples[:32], samples_mean=pipe.samples.mean(axis=0),
         lower=lo_df["Predicted_Lower"].values, median=med_df["Predicted_Median"].values,
         upper=up_df["Predicted_Upper"].values, coverage=np.array(cov), coverage_filtered=np.array(cov_f))


def dataset_golden():
    """Dataset.load_data / split_data / get_subset / view_data (data.py:30-374) on a CSV written from
    ``cases.long_table``."""
    import tempfile
    from pybmc.data import Dataset
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ensemble.csv")
        cases.long_table().to_csv(path, index=False)
        ds = Dataset(path)
        with contextlib.redirect_stdout(io.StringIO()):
            data = ds.load_data(models=["mA", "mB", "mC", "truth"], keys=["BE", "Rad"], domain_keys=["N", "Z"])
    out = {"BE": data["BE"].values.astype(float), "Rad": data["Rad"].values.astype(float),
           "columns": np.array(list(data["BE"].columns))}
    tr, va, te = ds.split_data(data, "BE", "random", train_size=0.6, val_size=0.2, test_size=0.2)
    out.update(train_idx=tr.index.values, val_idx=va.index.values, test_idx=te.index.values)
    tr2, va2, te2 = ds.split_data(data, "Rad", "random", train_size=0.5, val_size=0.3, test_size=0.2)
    out.update(train_idx2=tr2.index.values, val_idx2=va2.index.values, test_idx2=te2.index.values)
    subsets = {
        "tuple": ds.get_subset("BE", filters={"N": (10, 14)}),
        "list": ds.get_subset("BE", filters={"Z": [20, 22]}),
        "scalar": ds.get_subset("BE", filters={"Z": 23}, models_to_include=["mB", "truth"]),
        "callable": ds.get_subset("BE", filters={"N": lambda c: c % 2 == 0, "Z": (21, 24)}),
        "multi": ds.get_subset("Rad", filters={"multi": lambda r: r["N"] + r["Z"] > 35}, models_to_include=["mC"]),
    }
    for name, frame in subsets.items():
        out["subset_" + name] = frame.values.astype(float)
        out["subset_" + name + "_index"] = frame.index.values
        out["subset_" + name + "_columns"] = np.array(list(frame.columns))
    view = ds.view_data()
    out["view_models"] = np.array(view["available_models"])
    out["view_properties"] = np.array(view["available_properties"])
    out["view_series"] = ds.view_data("Rad", "mB").values
    save("dataset", **out)


if __name__ == "__main__":
    if "--dataset-only" not in sys.argv:
        main()
    dataset_golden()
