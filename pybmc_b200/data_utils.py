"""The numeric step of the reference's data layer that sits right before ``orthogonalize``:
the "inside_to_outside" split (pybmc/data.py:194-245, 303-320).  Loading and joining files stays with
upstream's ``Dataset``; this module only replaces its O(N R) distance double loop."""
import numpy as np
import torch

from . import _device as D
from . import _lib


def separate_points_distance_allSets(list1, list2, distance1, distance2, *, device=None):
    """Classify the points of ``list1`` by their distance to the nearest point of ``list2``.

    Same arguments and return value as ``Dataset.separate_points_distance_allSets``
    (pybmc/data.py:194): three lists of indices into ``list1`` -- within ``distance1`` of some
    reference point, within ``distance2`` but not ``distance1``, beyond ``distance2`` -- each in
    ascending order, Euclidean distance compared with ``<=`` as upstream."""
    lib = _lib.load()
    dev = D.device(device)
    pts = np.asarray(list1, dtype=np.float64)
    refs = np.asarray(list2, dtype=np.float64)
    if pts.size == 0:
        return [], [], []
    if pts.ndim != 2 or refs.ndim != 2 or pts.shape[1] != refs.shape[1]:
        raise ValueError(f"operands could not be broadcast together with shapes {pts.shape[1:]} {refs.shape[1:]}")
    if refs.shape[0] == 0:
        return [], [], list(range(len(pts)))
    pd_, rd = D.to_device(pts, dev), D.to_device(refs, dev)
    cls = torch.empty(pts.shape[0], dtype=torch.int32, device=dev)
    _lib.check(lib.bmc_nearest_class(D.ptr(pd_), pts.shape[0], D.ptr(rd), refs.shape[0], pts.shape[1],
                                     float(distance1), float(distance2), D.ptr(cls), D.stream_ptr(dev)),
               "bmc_nearest_class")
    c = cls.cpu().numpy()
    return tuple(np.flatnonzero(c == k).tolist() for k in (0, 1, 2))


def split_inside_to_outside(frame, stable_points, distance1, distance2, *, device=None):
    """``Dataset.split_data(..., splitting_algorithm="inside_to_outside")`` on one DataFrame
    (pybmc/data.py:303-330): rows are compared as whole tuples, as upstream does."""
    indexable = frame.reset_index(drop=True)
    points = list(indexable.itertuples(index=False, name=None))
    tr, va, te = separate_points_distance_allSets(points, stable_points, distance1, distance2, device=device)
    return indexable.iloc[tr], indexable.iloc[va], indexable.iloc[te]
