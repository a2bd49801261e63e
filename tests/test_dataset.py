"""``pybmc_b200.Dataset`` against the upstream ``pybmc.data.Dataset``: golden outputs written by
tests/golden/make_golden.py (load_data on a CSV, both random splits, every filter rule, view_data), the
upstream test-suite's own checks (tests/test_data.py), and the device distance split."""
import contextlib
import io

import numpy as np
import pandas as pd
import pytest

import cases


@pytest.fixture
def loaded(tmp_path):
    from pybmc_b200 import Dataset
    path = str(tmp_path / "ensemble.csv")
    cases.long_table().to_csv(path, index=False)
    ds = Dataset(path)
    data = ds.load_data(models=["mA", "mB", "mC", "truth"], keys=["BE", "Rad"], domain_keys=["N", "Z"])
    return ds, data


def test_load_csv_matches_upstream(loaded, golden):
    g = golden("dataset")
    ds, data = loaded
    assert list(data) == ["BE", "Rad"] and ds.data is not None and list(ds.data) == ["BE", "Rad"]
    assert list(data["BE"].columns) == list(g["columns"])
    assert np.array_equal(data["BE"].values.astype(float), g["BE"])
    assert np.array_equal(data["Rad"].values.astype(float), g["Rad"])
    assert 0 < len(data["BE"]) < 60                       # the inner join dropped points some model lacks
    assert ds.domain_keys == ["N", "Z"]


def test_random_split_matches_upstream(loaded, golden):
    g = golden("dataset")
    ds, data = loaded
    tr, va, te = ds.split_data(data, "BE", "random", train_size=0.6, val_size=0.2, test_size=0.2)
    assert np.array_equal(tr.index.values, g["train_idx"])
    assert np.array_equal(va.index.values, g["val_idx"])
    assert np.array_equal(te.index.values, g["test_idx"])
    tr, va, te = ds.split_data(data, "Rad", "random", train_size=0.5, val_size=0.3, test_size=0.2)
    assert np.array_equal(tr.index.values, g["train_idx2"])
    assert np.array_equal(va.index.values, g["val_idx2"])
    assert np.array_equal(te.index.values, g["test_idx2"])


def test_subsets_match_upstream(loaded, golden):
    g = golden("dataset")
    ds, _ = loaded
    subsets = {
        "tuple": ds.get_subset("BE", filters={"N": (10, 14)}),
        "list": ds.get_subset("BE", filters={"Z": [20, 22]}),
        "scalar": ds.get_subset("BE", filters={"Z": 23}, models_to_include=["mB", "truth"]),
        "callable": ds.get_subset("BE", filters={"N": lambda c: c % 2 == 0, "Z": (21, 24)}),
        "multi": ds.get_subset("Rad", filters={"multi": lambda r: r["N"] + r["Z"] > 35}, models_to_include=["mC"]),
    }
    for name, frame in subsets.items():
        assert list(frame.columns) == list(g["subset_" + name + "_columns"]), name
        assert np.array_equal(frame.index.values, g["subset_" + name + "_index"]), name
        assert np.array_equal(frame.values.astype(float), g["subset_" + name]), name
    before = ds.data["BE"].copy()
    ds.get_subset("BE", filters={"N": (10, 14)})
    assert ds.data["BE"].equals(before)                   # value semantics: the stored frame is untouched
    with pytest.raises(ValueError, match="not found in dataset"):
        ds.get_subset("Q")


def test_view_data_matches_upstream(loaded, golden):
    g = golden("dataset")
    ds, data = loaded
    view = ds.view_data()
    assert view == {"available_properties": list(g["view_properties"]), "available_models": list(g["view_models"])}
    assert ds.view_data("BE") is ds.data["BE"]
    assert np.array_equal(ds.view_data("Rad", "mB").values, g["view_series"])
    per_model = ds.view_data(model_name="mA")
    assert list(per_model) == ["BE", "Rad"] and list(per_model["BE"].columns) == ["N", "Z", "mA"]
    assert ds.view_data(model_name="nope") == {p: "[Model 'nope' not available]" for p in ("BE", "Rad")}
    with pytest.raises(KeyError):
        ds.view_data("Q")
    with pytest.raises(KeyError):
        ds.view_data("BE", "nope")
    from pybmc_b200 import Dataset
    with pytest.raises(RuntimeError, match="No data loaded"):
        Dataset("x.csv").view_data()


def test_load_data_errors_and_skips(tmp_path):
    """pybmc/data.py:57-121: argument errors, unsupported formats, skipped models, the empty frame."""
    from pybmc_b200 import Dataset
    with pytest.raises(ValueError, match="Data source must be specified"):
        Dataset().load_data(["a"], keys=["BE"], domain_keys=["N"])
    with pytest.raises(FileNotFoundError):
        Dataset(str(tmp_path / "missing.csv")).load_data(["a"], keys=["BE"], domain_keys=["N"])
    txt = tmp_path / "table.txt"
    txt.write_text("x\n1\n")
    with pytest.raises(ValueError, match="Unsupported file format"):
        Dataset(str(txt)).load_data(["a"], keys=["BE"], domain_keys=["N"])
    csv = tmp_path / "t.csv"
    pd.DataFrame({"N": [1, 2], "BE": [1.0, 2.0], "kind": ["a", "b"]}).to_csv(csv, index=False)
    with pytest.raises(ValueError, match="You must specify which properties"):
        Dataset(str(csv)).load_data(["a"], domain_keys=["N"])
    with pytest.raises(ValueError, match="Expected column 'model' not found"):
        Dataset(str(csv)).load_data(["a"], keys=["BE"], domain_keys=["N"])
    ds = Dataset(str(csv))
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        data = ds.load_data(["a", "b"], keys=["Rad"], domain_keys=["N"], model_column="kind")
    assert out.getvalue().count("[Skipped] Model") == 2 and "[Warning] No models with property 'Rad'" in out.getvalue()
    assert list(data["Rad"].columns) == ["N"] and len(data["Rad"]) == 0
    assert ds.data == {}                                   # upstream only stores frames it could build (:128)
    data = ds.load_data(["a", "b"], keys=["BE"], domain_keys=["N"], model_column="kind")
    assert list(data["BE"].columns) == ["N", "a", "b"] and len(data["BE"]) == 0    # disjoint points


def test_load_h5_path_reads_one_frame_per_model(tmp_path, monkeypatch):
    """HDF5 needs PyTables, which this image lacks: as upstream's own test does (tests/test_data.py:55-77),
    stand in for ``pandas.read_hdf`` and check the per-model keys, renaming, skipping and the join."""
    from pybmc_b200 import Dataset
    import pybmc_b200.data as mod
    frames = {"m1": pd.DataFrame({"x": [1, 2, 3], "y": [1, 2, 3], "target": [10., 20., 30.]}),
              "m2": pd.DataFrame({"x": [2, 3, 4], "y": [2, 3, 4], "target": [21., 31., 41.]}),
              "m3": pd.DataFrame({"x": [1], "y": [1], "other": [0.]})}
    seen = []

    def fake_read_hdf(path, key):
        seen.append(key)
        return frames[key]
    monkeypatch.setattr(mod.pd, "read_hdf", fake_read_hdf)
    h5 = tmp_path / "d.h5"
    h5.write_bytes(b"")
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        data = Dataset(str(h5)).load_data(["m1", "m2", "m3"], keys=["target"], domain_keys=["x", "y"])
    assert seen == ["m1", "m2", "m3"]
    assert "[Skipped] Model 'm3' missing columns ['target'] for property 'target'." in out.getvalue()
    assert data["target"].to_dict("list") == {"x": [2, 3], "y": [2, 3], "m1": [20., 30.], "m2": [21., 31.]}


def test_split_data_argument_errors():
    from pybmc_b200 import Dataset
    ds = Dataset()
    frame = cases.toy_frame()
    with pytest.raises(ValueError, match="not found in the provided data dictionary"):
        ds.split_data({"a": frame}, "b")
    with pytest.raises(TypeError):
        ds.split_data({"a": frame.values}, "a")
    with pytest.raises(ValueError, match="Missing required kwargs for 'random'"):
        ds.split_data({"a": frame}, "a", "random", train_size=0.5)
    with pytest.raises(ValueError, match="must equal 1.0"):
        ds.split_data({"a": frame}, "a", "random", train_size=0.5, val_size=0.3, test_size=0.3)
    with pytest.raises(ValueError, match="Missing required kwargs for 'inside_to_outside'"):
        ds.split_data({"a": frame}, "a", "inside_to_outside", distance1=1.0)
    with pytest.raises(ValueError, match="either 'random' or 'inside_to_outside'"):
        ds.split_data({"a": frame}, "a", "by_magic")
    tr, va, te = ds.split_data({"a": frame}, "a", "random", train_size=0.6, val_size=0.2, test_size=0.2)
    assert sorted(list(tr.index) + list(va.index) + list(te.index)) == list(range(6))


@pytest.mark.gpu
def test_inside_to_outside_split_on_device(golden):
    """Upstream's two distance tests (tests/test_data.py:108-127, 154-172) and the golden index lists of
    ``split_distance.npz``, through the Dataset methods."""
    from pybmc_b200 import Dataset
    ds = Dataset("fake_path.h5")
    coords = pd.DataFrame({"x": [1, 2, 3, 4], "y": [1, 2, 3, 4]})
    tr, va, te = ds.split_data({"target": coords}, "target", "inside_to_outside", stable_points=[(1, 1)],
                               distance1=0.1, distance2=100)
    assert len(tr) + len(va) + len(te) == 4
    assert tr.values.tolist() == [[1, 1]] and len(va) == 3 and len(te) == 0
    tr, va, te = ds.separate_points_distance_allSets(list1=[(1, 1), (2, 2)], list2=[(1.1, 1.1), (3, 3)],
                                                     distance1=0.2, distance2=1.5)
    assert (tr, va, te) == ([0], [1], [])
    g = golden("split_distance")
    frame = pd.DataFrame(g["points"], columns=["N", "Z"])
    stable = [tuple(p) for p in g["stable"]]
    tr, va, te = ds.split_data({"BE": frame}, "BE", "inside_to_outside", stable_points=stable, distance1=3.0,
                               distance2=7.5)
    assert np.array_equal(tr.index.values, g["train"]) and np.array_equal(va.index.values, g["val"])
    assert np.array_equal(te.index.values, g["test"])
