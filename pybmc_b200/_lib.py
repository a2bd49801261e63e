"""ctypes binding of libbmc_b200.so (the C ABI declared in include/bmc_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is not
available, every entry point of the package raises.  Build the library with
``python -m pybmc_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbmc_b200.so")

F32, F64 = 0, 1
STATS_NONE, STATS_DIAG, STATS_FULL = 0, 1, 2
NOISE_NONE, NOISE_PHILOX, NOISE_EXTERNAL = 0, 1, 2
ERR_ARG, ERR_CUDA, ERR_CONVERGE, ERR_WORKSPACE = -1, -2, -3, -4
MAX_COMPONENTS = 64
MAX_QUANTILES = 8
LAYOUT_AUTO, LAYOUT_THREAD, LAYOUT_GROUP, LAYOUT_WARP = 0, 1, 2, 3
LAYOUTS = {None: 0, "auto": 0, "thread": 1, "group": 2, "warp": 3}
HIST_BINS = 512

_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_int = C.c_int
_dbl = C.c_double
_sz = C.c_size_t


class GibbsProblem(C.Structure):
    _fields_ = [("k", _int), ("d", _p), ("pull", _p), ("g_ols", _p), ("w", _p), ("dense_w", _int),
                ("rss_min", _dbl), ("n_obs", _dbl), ("nu0", _dbl), ("sigma20", _dbl), ("sigma2_init", _dbl),
                ("layout", _int), ("workspace", _p), ("workspace_bytes", _sz)]


class GibbsHist(C.Structure):
    _fields_ = [("every", _i64), ("lo", _p), ("inv_width", _p), ("counts", _p), ("workspace", _p),
                ("workspace_bytes", _sz)]


class SimplexProblem(C.Structure):
    _fields_ = [("k", _int), ("m", _int), ("gram", _p), ("b_ols", _p), ("step", _p), ("vt_hat", _p),
                ("rss_min", _dbl), ("rss_zero", _dbl), ("n_obs", _dbl), ("nu0", _dbl), ("sigma20", _dbl),
                ("layout", _int)]


class PredictProblem(C.Structure):
    _fields_ = [("n_points", _i64), ("point0", _u64), ("n_draws", _i64), ("k", _int), ("u", _p), ("mu", _p),
                ("truth", _p), ("theta", _p), ("noise_mode", _int), ("seed", _u64), ("noise", _p),
                ("ld_noise", _i64), ("nq", _int), ("probs", C.POINTER(_dbl)), ("theta_mean", _p),
                ("theta_cov", _p), ("center", _p), ("scale", _p), ("tensor_min_k", _int)]


# name -> (restype, argtypes); every symbol include/bmc_b200.h declares
SIGNATURES = {
    "bmc_version": (_int, []),
    "bmc_last_error": (C.c_char_p, []),
    "bmc_device_caps": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_sz)]),
    "bmc_center_rows": (_int, [_p, _i64, _int, _i64, _p, _p, _p, _p, _i64, _p]),
    "bmc_gram_workspace_bytes": (_sz, [_i64, _int]),
    "bmc_gram": (_int, [_p, _i64, _int, _i64, _p, _p, _p, _p, _sz, _p]),
    "bmc_project_rows": (_int, [_p, _i64, _int, _i64, _p, _p, _int, _p, _i64, _p]),
    "bmc_rss_workspace_bytes": (_sz, [_i64]),
    "bmc_residual_ss": (_int, [_p, _i64, _int, _i64, _p, _p, _p, _p, _sz, _p]),
    "bmc_padded_components": (_int, [_int]),
    "bmc_gibbs_n_stat": (_i64, [_int, _int]),
    "bmc_gibbs_workspace_bytes": (_sz, [_i64]),
    "bmc_gibbs_hist_workspace_bytes": (_sz, [_int]),
    "bmc_gibbs_run": (_int, [_int, C.POINTER(GibbsProblem), _u64, _u64, _i64, _i64, _i64, _i64, _i64, _p, _p,
                             _int, C.POINTER(GibbsHist), _p]),
    "bmc_gibbs_simplex_run": (_int, [_int, C.POINTER(SimplexProblem), _u64, _u64, _i64, _i64, _i64, _i64, _i64,
                                     _p, _p, _int, _p, _p]),
    "bmc_gibbs_literal_run": (_int, [_int, _p, _p, _i64, _int, _p, _p, _dbl, _dbl, _dbl, _u64, _u64, _i64, _i64,
                                     _p, _p]),
    "bmc_predict_workspace_bytes": (_sz, [_int, _i64, _int, _i64]),
    "bmc_predict_theta_stride": (_int, [_int]),
    "bmc_predict_fused": (_int, [_int, C.POINTER(PredictProblem), _p, _p, _p, _p, _p, _p, _i64, _p, _sz,
                                 C.POINTER(_int), _p]),
    "bmc_coverage_counts": (_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "bmc_column_moments": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "bmc_coverage_levels": (_int, [_p, _p, _i64, _p, _p, _int, _p, _p]),
    "bmc_nearest_class": (_int, [_p, _i64, _p, _i64, _int, _dbl, _dbl, _p, _p]),
}

# benchmark tooling (libbmc_probe.so, pybmc_b200/csrc/bench/probe.h): not part of the product ABI
PROBE_LIB_PATH = os.path.join(_HERE, "csrc", "bench", "libbmc_probe.so")
PROBE_SIGNATURES = {
    "bmc_probe_ops_per_iteration": (_int, [_int]),
    "bmc_probe": (_int, [_int, _i64, _int, _int, _p, _p]),
}

_lib = None


class BmcError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BmcError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python -m pybmc_b200.build`). pybmc_b200 has no CPU path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_probes():
    """The pipe-peak probes of the benchmark (built by pybmc_b200.build.build_probe_library)."""
    if not os.path.exists(PROBE_LIB_PATH):
        raise BmcError(f"{PROBE_LIB_PATH} is missing (run `python -m pybmc_b200.build`)")
    lib = C.CDLL(PROBE_LIB_PATH)
    for name, (res, args) in PROBE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def check(rc, what=""):
    """Map a C-ABI status to the exception the reference would raise."""
    if rc == 0:
        return
    msg = load().bmc_last_error().decode("utf-8", "replace")
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise BmcError(f"{what or 'libbmc_b200'} failed ({rc}): {msg}")
