"""BayesianModelCombination on the GPU: the upstream test-suite's checks (tests/test_bmc.py,
tests/test_inference_utils.py -- shapes, types, column names, exceptions, tiny degenerate shapes)
against the drop-in class, then the whole pipeline against the reference's golden run."""
import numpy as np
import pandas as pd
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture
def toy():
    import pybmc_b200 as pb
    df = cases.toy_frame()
    bmc = pb.BayesianModelCombination(["model1", "model2", "model3", "truth"], {"target": df}, "truth")
    return bmc, df, df.iloc[:4]


def _two_model():
    import pybmc_b200 as pb
    data = {"property": pd.DataFrame({"model1": [1, 2], "model2": [3, 4], "truth": [5, 6]})}
    bmc = pb.BayesianModelCombination(["model1", "model2"], data, "truth")
    train = pd.DataFrame({"model1": [1, 2], "model2": [3, 4], "truth": [5, 6]})
    return bmc, train


def test_init_validation():
    import pybmc_b200 as pb
    data = {"property": pd.DataFrame({"model1": [1, 2], "model2": [3, 4]})}
    bmc = pb.BayesianModelCombination(["model1", "model2"], data, "truth")
    assert bmc.models_list == ["model1", "model2"] and bmc.data_dict == data and bmc.truth_column_name == "truth"
    with pytest.raises(ValueError):
        pb.BayesianModelCombination("not_a_list", data, "truth")
    with pytest.raises(ValueError):
        pb.BayesianModelCombination(["model1"], "not_a_dict", "truth")


def test_orthogonalize_shapes(toy):
    bmc, df, train = toy
    bmc.orthogonalize("target", train, 2)
    assert bmc.U_hat.shape == (4, 2) and bmc.Vt_hat.shape == (2, 3) and bmc.S_hat.shape == (2,)
    assert bmc.centered_experiment_train.shape == (4,) and bmc.current_property == "target"
    small, train2 = _two_model()
    small.orthogonalize("property", train2, 1)
    for name in ("centered_experiment_train", "U_hat", "Vt_hat", "S_hat", "_predictions_mean_train"):
        assert getattr(small, name) is not None


def test_train_default_and_simplex(toy, capsys):
    bmc, df, train = toy
    bmc.orthogonalize("target", train, 2)
    bmc.train()
    out = capsys.readouterr().out
    assert out.count("[INFO] Using default value for") == 8      # every defaulted key is announced
    assert bmc.samples.shape == (50000, 3) and not np.any(np.isnan(bmc.samples))
    small, train2 = _two_model()
    small.orthogonalize("property", train2, 1)
    small.train()
    assert small.samples.shape == (50000, 2) and np.all(np.isfinite(small.samples))
    small.train({"iterations": 100, "sampler": "simplex", "burn": 10, "stepsize": 0.01,
                 "b_mean_prior": np.zeros(1), "b_mean_cov": np.eye(1), "nu0_chosen": 1.0, "sigma20_chosen": 0.02})
    assert small.samples.shape == (100, 2) and np.all(np.isfinite(small.samples))
    assert "Acceptance rate:" in capsys.readouterr().out


def test_predict_predict2_evaluate(toy):
    bmc, df, train = toy
    bmc.orthogonalize("target", train, 2)
    bmc.train({"iterations": 12000})
    rndm_m, lower_df, median_df, upper_df = bmc.predict2("target")
    assert rndm_m.shape == (10000, 6)
    assert list(lower_df.columns) == ["x", "y", "Predicted_Lower"]
    assert "Predicted_Median" in median_df.columns and "Predicted_Upper" in upper_df.columns
    assert np.all(lower_df["Predicted_Lower"].values <= median_df["Predicted_Median"].values)
    assert np.all(median_df["Predicted_Median"].values <= upper_df["Predicted_Upper"].values)
    X = df[["x", "y", "model1", "model2", "model3"]]
    rndm_m, lower_df, median_df, upper_df = bmc.predict(X)
    assert rndm_m.shape == (10000, 6) and not lower_df.empty and not median_df.empty and not upper_df.empty
    res = bmc.evaluate()
    assert isinstance(res, list) and len(res) == 21 and all(isinstance(v, float) for v in res)
    assert res[0] == 0.0 or res[0] <= res[-1]
    assert all(0.0 <= v <= 100.0 for v in res)
    filt = bmc.evaluate({"x": (2, 5)})
    assert len(filt) == 21
    with pytest.raises(ValueError):
        bmc.predict(X.values)
    with pytest.raises(KeyError):
        bmc.predict2("nope")
    with pytest.raises(ValueError):                     # fewer than 10000 posterior rows (sampling_utils.py:57)
        bmc.train({"iterations": 500})
        bmc.predict2("target")


def test_predict2_model_subset(capsys):
    import pybmc_b200 as pb
    frame, models = cases.ensemble_frame(11, 40, 5)
    other = frame.drop(columns=["m3"])
    bmc = pb.BayesianModelCombination(models, {"BE": frame, "Other": other}, "truth")
    bmc.orthogonalize("BE", frame.iloc[:28], 3)
    bmc.train({"iterations": 10500, "seed": 4})
    capsys.readouterr()
    rndm_m, lo, med, hi = bmc.predict2("Other", seed=3)
    assert "WARNING: Predicting on property 'Other' with missing models: ['m3']" in capsys.readouterr().out
    assert rndm_m.shape == (10000, 40)
    extra = frame.copy()
    bmc2 = pb.BayesianModelCombination(models[:4], {"BE": frame.drop(columns=["m4"]), "More": extra}, "truth")
    bmc2.models = models[:4]
    bmc2.orthogonalize("BE", frame.iloc[:28], 2)
    bmc2.train({"iterations": 10000, "seed": 4})
    # "More" holds a model column (m4) the combination was not trained with -> upstream treats it as a domain key
    out = bmc2.predict2("More", return_draws=False)
    assert out[0] is None and "m4" in out[1].columns


def test_pipeline_against_reference_golden_run(golden):
    """orthogonalize -> train -> predict2 -> evaluate on the ensemble the golden run used.  Stochastic
    outputs: within 3 Monte-Carlo standard errors of the reference's run (its own error included)."""
    import pybmc_b200 as pb
    g = golden("pipeline")
    frame, models = cases.ensemble_frame(11, 40, 5)
    bmc = pb.BayesianModelCombination(models, {"BE": frame}, "truth")
    bmc.orthogonalize("BE", frame.iloc[:28], 3)
    bmc.train({"iterations": 10500, "sampler": "gibbs_sampling", "seed": 12, "n_chains": 1})
    # coefficient signs follow the arbitrary sign of each singular vector: compare model weights
    w_ref = g["samples_mean"][:3]
    mine = bmc.samples.mean(axis=0)
    sd = bmc.samples.std(axis=0)
    se = sd * np.sqrt(2.0 / 10500)
    assert np.all(np.abs(np.abs(mine[:3]) - np.abs(w_ref)) < 3 * se[:3])
    assert abs(mine[3] - g["samples_mean"][3]) < 3 * se[3]
    _, lo, med, hi = bmc.predict2("BE", seed=8)
    spread = (g["upper"] - g["lower"]) / 3.92
    assert np.all(np.abs(med["Predicted_Median"].values - g["median"]) < 3 * 1.2533 * spread * np.sqrt(2 / 10000) + 3 * spread * np.sqrt(2 / 10500))
    assert np.all(np.abs(lo["Predicted_Lower"].values - g["lower"]) < 0.12 * spread)
    assert np.all(np.abs(hi["Predicted_Upper"].values - g["upper"]) < 0.12 * spread)
    cov = np.array(bmc.evaluate(seed=9))
    # 40 points: one point changing sides moves a level by 2.5 %
    assert np.max(np.abs(cov - g["coverage"])) <= 5.0
    assert cov[0] == g["coverage"][0] == 0.0


def test_edge_cases_empty_and_single_point(toy):
    import pybmc_b200 as pb
    from pybmc_b200.sampling_utils import rndm_m_random_calculator
    bmc, df, train = toy
    bmc.orthogonalize("target", train, 2)
    bmc.train({"iterations": 10000, "dtype": "float32", "seed": 1})
    assert bmc.samples.dtype == np.float32 and bmc.samples.shape == (10000, 3)
    one = df.iloc[[2]][["x", "model1", "model2", "model3"]]
    rndm_m, lo, med, hi = bmc.predict(one, seed=2)                         # a single point
    assert rndm_m.shape == (10000, 1) and len(lo) == len(med) == len(hi) == 1
    none = df.iloc[:0][["x", "model1", "model2", "model3"]]
    rndm_m, lo, med, hi = bmc.predict(none)                                # no points at all
    assert rndm_m.shape == (10000, 0) and lo.empty and list(lo.columns) == ["x", "Predicted_Lower"]
    with pytest.raises(ZeroDivisionError):                                 # upstream divides by the point count
        bmc.evaluate({"x": (100, 200)})
    empty = pb.gibbs_sampler(np.array([1.0, 2.0, 3.0]), np.eye(3)[:, :2], 0, (np.zeros(2), np.eye(2), 1.0, 1.0))
    assert empty.shape == (0, 3)
    with pytest.raises(np.linalg.LinAlgError):                             # singular X'X, as np.linalg.inv upstream
        pb.gibbs_sampler(np.ones(3), np.ones((3, 2)), 5, (np.zeros(2), np.eye(2), 1.0, 1.0))
