"""``Dataset``: the data layer either side of the accelerated path (drop-in for ``pybmc.data.Dataset``,
pybmc/data.py:7-374).

Loading, joining, viewing and filtering are pandas work done once per run and stay on the host, with
upstream's signatures, return types, messages and exceptions.  The one super-linear step -- the
"inside_to_outside" split, an O(N R) double loop of ``np.linalg.norm`` upstream (pybmc/data.py:194-245)
-- runs on the device through ``bmc_nearest_class`` (``pybmc_b200.data_utils``); there is no CPU
version of it here.
"""
import os
from functools import reduce

import numpy as np
import pandas as pd

from . import data_utils
from .bmc import _filter_rows

_SPLIT_ARGS = {"random": ["train_size", "val_size", "test_size"],
               "inside_to_outside": ["stable_points", "distance1", "distance2"]}


def _model_frames_h5(path, models):
    """One stored frame per model key (pybmc/data.py:70-72)."""
    for model in models:
        yield model, pd.read_hdf(path, key=model)


def _model_frames_csv(path, models, model_column):
    """One long table, rows tagged by ``model_column`` (pybmc/data.py:86-93).  The missing-column error is
    raised while visiting the first model, so an empty model list reads the file and raises nothing."""
    table = pd.read_csv(path)
    for model in models:
        if model_column not in table.columns:
            raise ValueError(f"Expected column '{model_column}' not found in CSV.")
        yield model, table[table[model_column] == model]


class Dataset:
    """Datasets for Bayesian model combination: load from HDF5/CSV, view, split, filter.

    Attributes:
        data_source (str): path of the data file.
        data (dict[str, pandas.DataFrame]): loaded frames by property.
        domain_keys (list[str]): the domain columns the models are aligned on.
    """

    def __init__(self, data_source=None):
        self.data_source = data_source
        self.data = {}
        self.domain_keys = ["X1", "X2"]                      # pybmc/data.py:28

    # ------------------------------------------------------------------ loading
    def load_data(self, models, keys=None, domain_keys=None, model_column="model"):
        """One DataFrame per property in ``keys``: the domain columns plus one column per model, on the
        points every model covers (inner join on ``domain_keys``; pybmc/data.py:30-129).

        Raises ``ValueError`` (no data source, no ``keys``, unsupported extension, CSV without
        ``model_column``) and ``FileNotFoundError`` as upstream; models lacking a column are reported
        with a ``[Skipped]`` line and left out.
        """
        self.domain_keys = domain_keys                        # set before validation, as upstream (:55)
        source = self.data_source
        if source is None:
            raise ValueError("Data source must be specified.")
        if not os.path.exists(source):
            raise FileNotFoundError(f"Data source '{source}' not found.")
        if keys is None:
            raise ValueError("You must specify which properties to extract via 'keys'.")

        loaded = {}
        for prop in keys:
            if source.endswith(".h5"):
                frames, noun = _model_frames_h5(source, models), "property"
            elif source.endswith(".csv"):
                frames, noun = _model_frames_csv(source, models, model_column), "key"
            else:
                raise ValueError("Unsupported file format. Only .h5 and .csv are supported.")
            wanted = domain_keys + [prop]
            columns, skipped = [], []
            for model, frame in frames:
                missing = [c for c in wanted if c not in frame.columns]
                if missing:
                    print(f"[Skipped] Model '{model}' missing columns {missing} for {noun} '{prop}'.")
                    skipped.append(model)
                    continue
                columns.append(frame[wanted].rename(columns={prop: model}))
            if not columns:
                print(f"[Warning] No models with property '{prop}'. Resulting DataFrame will be empty.")
                loaded[prop] = pd.DataFrame(columns=domain_keys + [m for m in models if m not in skipped])
                continue                                      # note: self.data is not updated (:115-121)
            loaded[prop] = reduce(lambda left, right: pd.merge(left, right, on=domain_keys, how="inner"), columns)
            self.data = loaded
        return loaded

    # ------------------------------------------------------------------ viewing
    def view_data(self, property_name=None, model_name=None):
        """No arguments: ``{"available_properties", "available_models"}``; a model: ``{property: frame of
        the domain columns + that model}``; a property: its frame; both: that model's column
        (pybmc/data.py:131-192).  ``RuntimeError`` before ``load_data``, ``KeyError`` for unknown names."""
        if not self.data:
            raise RuntimeError("No data loaded. Run `load_data(...)` first.")
        if property_name is not None:
            if property_name not in self.data:
                raise KeyError(f"Property '{property_name}' not found.")
            frame = self.data[property_name]
            if model_name is None:
                return frame
            if model_name not in frame.columns:
                raise KeyError(f"Model '{model_name}' not found in property '{property_name}'.")
            return frame[model_name]
        if model_name is not None:
            return {prop: (frame[self.domain_keys + [model_name]] if model_name in frame.columns
                           else f"[Model '{model_name}' not available]")
                    for prop, frame in self.data.items()}
        models = {c for frame in self.data.values() for c in frame.columns if c not in self.domain_keys}
        return {"available_properties": list(self.data), "available_models": sorted(models)}

    # ------------------------------------------------------------------ splitting
    def separate_points_distance_allSets(self, list1, list2, distance1, distance2):
        """Indices of ``list1`` within ``distance1`` of some point of ``list2`` / within ``distance2`` only /
        beyond (pybmc/data.py:194-245), classified on the device by ``bmc_nearest_class``."""
        tr, va, te = data_utils.separate_points_distance_allSets(list1, list2, distance1, distance2)
        return tr, va, te

    def split_data(self, data_dict, property_name, splitting_algorithm="random", **kwargs):
        """(train, validation, test) frames of one property (pybmc/data.py:247-330).

        ``"random"``: ``train_size``, ``val_size``, ``test_size`` summing to 1; two chained
        ``train_test_split`` calls with ``random_state=1``, as upstream, so the same rows come out.
        ``"inside_to_outside"``: ``stable_points``, ``distance1``, ``distance2``; whole rows are the
        points, distances on the device.
        """
        if property_name not in data_dict:
            raise ValueError(f"Property '{property_name}' not found in the provided data dictionary.")
        frame = data_dict[property_name]
        if not isinstance(frame, pd.DataFrame):
            raise TypeError("Data for the specified property must be a pandas DataFrame.")
        if splitting_algorithm not in _SPLIT_ARGS:
            raise ValueError("splitting_algorithm must be either 'random' or 'inside_to_outside'")
        required = _SPLIT_ARGS[splitting_algorithm]
        if any(name not in kwargs for name in required):
            raise ValueError(f"Missing required kwargs for '{splitting_algorithm}': {required}")
        if splitting_algorithm == "inside_to_outside":
            return data_utils.split_inside_to_outside(frame, *(kwargs[name] for name in required))

        from sklearn.model_selection import train_test_split          # upstream's splitter (:3)
        train_size, val_size, test_size = (kwargs[name] for name in required)
        if not np.isclose(train_size + val_size + test_size, 1.0):
            raise ValueError("train_size + val_size + test_size must equal 1.0")
        rows = frame.reset_index(drop=True)
        train_idx, rest = train_test_split(rows.index, train_size=train_size, random_state=1)
        val_idx, test_idx = train_test_split(rest, test_size=1 - val_size / (val_size + test_size), random_state=1)
        return rows.iloc[train_idx], rows.iloc[val_idx], rows.iloc[test_idx]

    # ------------------------------------------------------------------ filtering
    def get_subset(self, property_name, filters=None, models_to_include=None):
        """Rows of a property passing ``filters`` (the rules of ``evaluate``: callable on a column, (low,
        high) tuple, list of values, single value, ``"multi"`` row callable), optionally restricted to the
        columns N, Z and ``models_to_include`` (pybmc/data.py:332-374)."""
        if property_name not in self.data:
            raise ValueError(f"Property '{property_name}' not found in dataset.")
        frame = _filter_rows(self.data[property_name].copy(), filters)
        if models_to_include is not None:
            keep = [c for c in ("N", "Z") if c in frame.columns]          # fixed names upstream (:368)
            frame = frame[keep + [m for m in models_to_include if m in frame.columns]]
        return frame
