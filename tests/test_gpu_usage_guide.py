"""The upstream usage guide (docs/usage.md:10-93) run line by line against ``pybmc_b200`` -- Dataset ->
split_data -> BayesianModelCombination.orthogonalize / train / predict2 / evaluate -- on a six-model CSV in
which, as in the guide, the truth column ("AME2020") is also listed as a model.  Deterministic stages are
checked against the oracle to 1e-10; the stochastic ones against the oracle's sampler run on NumPy's
generators (= the reference, by the golden pin) within Monte-Carlo error."""
import numpy as np
import pandas as pd
import pytest

from oracle import bmc_oracle as oc

pytestmark = pytest.mark.gpu

MODELS = ["FRDM12", "HFB24", "D1M", "UNEDF1", "BCPM", "AME2020"]


def _mass_table(seed=77, n_points=240):
    """Long CSV table of six "mass models": a smooth binding-energy surface, model-specific scale and offset
    errors, and white noise; AME2020 is the measured value."""
    rng = np.random.default_rng(seed)
    n = rng.integers(8, 160, n_points)
    z = rng.integers(8, 110, n_points)
    pts = np.unique(np.column_stack([n, z]), axis=0)
    n, z = pts[:, 0], pts[:, 1]
    a = n + z
    be = 15.8 * a - 18.3 * a ** (2 / 3) - 0.714 * z * (z - 1) / a ** (1 / 3) - 23.2 * (n - z) ** 2 / a
    rows = []
    for j, model in enumerate(MODELS):
        if model == "AME2020":
            val = be + rng.normal(0, 0.15, a.size)
        else:
            val = be * (1 + rng.normal(0, 2e-3)) + rng.normal(0, 2.0) + rng.normal(0, 0.5, a.size)
        rows.append(pd.DataFrame({"model": model, "N": n, "Z": z, "BE": val}))
    return pd.concat(rows, ignore_index=True)


def test_usage_guide_end_to_end(tmp_path, capsys):
    from pybmc_b200 import BayesianModelCombination, Dataset
    path = str(tmp_path / "selected_data.csv")
    _mass_table().to_csv(path, index=False)

    # 1. load (docs/usage.md:16-25)
    dataset = Dataset(path)
    data_dict = dataset.load_data(models=MODELS, keys=["BE"], domain_keys=["N", "Z"])
    frame = data_dict["BE"]
    assert list(frame.columns) == ["N", "Z"] + MODELS and len(frame) > 200
    # 2. split (:33-41)
    train_df, val_df, test_df = dataset.split_data(data_dict, "BE", splitting_algorithm="random", train_size=0.6,
                                                   val_size=0.2, test_size=0.2)
    assert len(train_df) + len(val_df) + len(test_df) == len(frame)
    # 3. model, orthogonalize, train (:49-71)
    bmc = BayesianModelCombination(models_list=MODELS, data_dict=data_dict, truth_column_name="AME2020")
    assert bmc.models == MODELS                              # only the literal "truth" is dropped (bmc.py:75)
    bmc.orthogonalize("BE", train_df, components_kept=3)
    ref = oc.orthogonalize_arrays(train_df[MODELS].values, train_df["AME2020"].values, 3, full_matrices=False)
    sign = np.sign(np.sum(ref["Vt_hat"] * bmc.Vt_hat, axis=1))
    np.testing.assert_allclose(bmc.S_hat, ref["S_hat"], rtol=1e-10)
    np.testing.assert_allclose(bmc.Vt_hat * sign[:, None], ref["Vt_hat"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(bmc.U_hat * sign[None, :], ref["U_hat"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(bmc.centered_experiment_train, ref["y"], rtol=1e-10, atol=1e-10)
    bmc.train(training_options={"iterations": 50000, "sampler": "gibbs_sampling"})
    assert "[INFO] Using default value for" in capsys.readouterr().out
    assert bmc.samples.shape == (50000, 4) and np.all(bmc.samples[:, -1] > 0)
    # the oracle's chain on NumPy's generators, same problem (sign-aligned to the device basis)
    prior = [np.zeros(3), np.diag(ref["S_hat"] ** 2), 1.0, 0.02]
    want = oc.gibbs_conjugate(ref["y"], ref["U_hat"] * sign[None, :], 20000, prior)
    for col in range(4):
        mcse = want[:, col].std() / np.sqrt(20000 / 3.0) + bmc.samples[:, col].std() / np.sqrt(50000 / 3.0)
        assert abs(bmc.samples[:, col].mean() - want[:, col].mean()) < 5 * mcse, col
        assert abs(bmc.samples[:, col].std() / want[:, col].std() - 1) < 0.05, col
    # 4. predict2 (:79-83)
    rndm_m, lower_df, median_df, upper_df = bmc.predict2("BE")
    assert rndm_m.shape == (10000, len(frame))
    assert list(median_df.columns) == ["N", "Z", "Predicted_Median"]
    assert list(lower_df.columns) == ["N", "Z", "Predicted_Lower"] and list(upper_df.columns) == ["N", "Z", "Predicted_Upper"]
    np.testing.assert_allclose(np.percentile(rndm_m, 50, axis=0), median_df["Predicted_Median"].values, rtol=1e-12)
    np.testing.assert_allclose(np.percentile(rndm_m, 2.5, axis=0), lower_df["Predicted_Lower"].values, rtol=1e-12)
    # the same predictive law as the oracle's draws from the oracle's chain: compare medians and band widths
    o_draws, _ = oc.predictive_draws(frame[MODELS].values, want, ref["Vt_hat"] * sign[:, None],
                                     np.random.default_rng(5), n_draws=10000)
    o_med = np.percentile(o_draws, 50, axis=0)
    sd = o_draws.std(axis=0)
    assert np.max(np.abs(median_df["Predicted_Median"].values - o_med) / sd) < 0.1
    width = upper_df["Predicted_Upper"].values - lower_df["Predicted_Lower"].values
    o_width = np.percentile(o_draws, 97.5, axis=0) - np.percentile(o_draws, 2.5, axis=0)
    assert np.max(np.abs(width / o_width - 1)) < 0.08
    # 5. evaluate (:91-95): 21 levels, monotone, and the same numbers as coverage() on the returned draws
    cov = bmc.evaluate()
    assert isinstance(cov, list) and len(cov) == 21 and cov[0] == 0.0 and all(isinstance(c, float) for c in cov)
    assert all(b >= a for a, b in zip(cov, cov[1:]))
    o_cov = oc.coverage_levels(np.arange(0, 101, 5), o_draws, frame["AME2020"].values)
    assert np.max(np.abs(np.array(cov) - np.array(o_cov))) < 6.0          # percent of ~230 nuclei, two independent runs
