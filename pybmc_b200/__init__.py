"""pybmc_b200: pyBMC's Bayesian-model-combination inference path on B200 (sm_100a) kernels.

Same public names as upstream ``pybmc`` (pybmc/__init__.py:11-24) for the inference path:

    Dataset, BayesianModelCombination, gibbs_sampler, USVt_hat_extraction, coverage

plus, importable from their submodules as upstream, ``gibbs_sampler_simplex`` and
``rndm_m_random_calculator``.  ``Dataset`` (pybmc/data.py) is host-side pandas glue except for its
distance split, which runs on the device.  Importing the package does not touch the GPU; the first call
loads ``csrc/libbmc_b200.so`` and raises if it or a CUDA device is missing (no CPU path).
"""
from .bmc import BayesianModelCombination, orthogonalize_arrays
from .data import Dataset
from .inference_utils import (ConjugateSampler, GibbsResult, SimplexSampler, USVt_hat_extraction, gibbs_sampler,
                              gibbs_sampler_literal, gibbs_sampler_simplex, run_gibbs, run_gibbs_simplex)
from .sampling_utils import (PredictiveProblem, PredictiveResult, column_percentiles, coverage,
                             coverage_from_counts, predictive_summary, rndm_m_random_calculator)

__version__ = "0.1.0"

__all__ = [
    "Dataset",
    "BayesianModelCombination",
    "gibbs_sampler",
    "gibbs_sampler_simplex",
    "USVt_hat_extraction",
    "coverage",
    "rndm_m_random_calculator",
]


def install_as_pybmc():
    """Make ``import pybmc`` (and ``pybmc.bmc``, ``pybmc.data``, ``pybmc.inference_utils``,
    ``pybmc.sampling_utils``) resolve to this package for the rest of the process, so that scripts written
    against upstream run unchanged: call it once before they import ``pybmc``.  Refuses to shadow an upstream
    ``pybmc`` that is already imported."""
    import sys
    from . import bmc, data, inference_utils, sampling_utils
    this = sys.modules[__name__]
    loaded = sys.modules.get("pybmc")
    if loaded is not None and loaded is not this:
        raise RuntimeError("another 'pybmc' is already imported; call install_as_pybmc() before importing it")
    sys.modules["pybmc"] = this
    for name, module in (("bmc", bmc), ("data", data), ("inference_utils", inference_utils),
                         ("sampling_utils", sampling_utils)):
        sys.modules["pybmc." + name] = module
    return this
