"""``BayesianModelCombination`` with the reference's surface (pybmc/bmc.py:11-376); every numeric
step runs on the GPU through libbmc_b200.so.

Same constructor, methods, attributes, return types, prints and exceptions as upstream, so a
script written against pyBMC runs unchanged.  Optional extras are additive: new keys in
``training_options`` (``n_chains``, ``seed``, ``dtype``, ``thin``) and keyword arguments on
``predict`` / ``predict2`` / ``evaluate`` (``n_draws``, ``seed``, ``dtype``, ``return_draws``).
"""
import numpy as np
import pandas as pd
import torch

from . import _device as D
from . import _lib
from .inference_utils import gibbs_sampler, gibbs_sampler_simplex
from .sampling_utils import (DEFAULT_DRAWS, coverage_from_counts, predictive_summary,
                             rndm_m_random_calculator)

# lambda_K / lambda_1 of the Gram matrix below which its eigen-decomposition cannot give 1e-10
# singular vectors (error ~ eps * lambda_1 / lambda_K); the thin SVD of Xc is used instead
GRAM_RATIO_MIN = 1e-5


def orthogonalize_arrays(preds, truth, components_kept, method="auto", device=None, reduce=None):
    """Centre the model predictions per point and project on the truncated SVD basis
    (pybmc/bmc.py:102-130 + pybmc/inference_utils.py:147-168) on the device.

    Kernels: ``bmc_center_rows`` (mu, y, Xc), ``bmc_gram`` (Xc'Xc), cuSOLVER for the small
    M-by-M eigenproblem (or the thin SVD when the spectrum is too graded for the Gram route),
    ``bmc_project_rows`` (U_hat = Xc Vt_hat').  The reference's n-by-n U is never formed.

    ``reduce`` (optional): a callable summing a device tensor over ranks; ``preds`` / ``truth`` are
    then this rank's rows, the M-by-M Gram matrix is all-reduced, every rank solves the same small
    eigenproblem and projects its own rows (Gram route only).

    Returns dict(y, mu, U_hat, S_hat, Vt_hat, Vt_hat_normalized, method); with ``reduce`` the
    per-point entries (y, mu, U_hat) cover this rank's rows.
    """
    dev = D.device(device)
    with D.on(dev):
        return _orthogonalize(dev, preds, truth, components_kept, method, reduce)


def _tsqr_svd(xc, reduce):
    """Singular values / right singular vectors of the row-sharded centred matrix without squaring its
    condition number: local QR, the M-by-M R factors of all ranks stacked (exchanged as one summed
    [world, M, M] tensor -- 512 KB per rank at M = 256), SVD of the stack on every rank."""
    m = xc.shape[1]
    r_local = torch.linalg.qr(xc, mode="r")[1] if xc.shape[0] > 0 else xc.new_zeros((0, m))
    stack = xc.new_zeros((reduce.world, m, m))
    stack[reduce.rank, : r_local.shape[0]] = r_local
    stack = reduce(stack)
    _, s, vt = torch.linalg.svd(stack.reshape(reduce.world * m, m), full_matrices=False)
    return s, vt


def _orthogonalize(dev, preds, truth, components_kept, method, reduce):
    lib = _lib.load()
    preds = np.asarray(preds, dtype=np.float64)
    if preds.ndim != 2:
        raise ValueError("model predictions must be [n_points, n_models]")
    n, m = preds.shape
    k = int(components_kept)
    if reduce is not None and method == "auto":
        method = "gram_checked"          # Gram route, TSQR when the spectrum is too graded for it
    if k < 1 or k > (m if reduce is not None else min(n, m)):
        raise IndexError(f"components_kept={k} outside 1..{min(n, m)}")   # upstream: IndexError from U.T[i]
    truth = np.asarray(truth, dtype=np.float64).reshape(-1)
    if truth.shape[0] != n:
        # upstream subtracts two pandas columns of the same frame (bmc.py:109-111); arrays must match
        raise ValueError(f"truth has {truth.shape[0]} entries for {n} rows of predictions")
    pd_ = D.to_device(preds, dev)
    td = D.to_device(truth, dev)
    st = D.stream_ptr(dev)
    mu = torch.empty(n, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    xc = torch.empty((n, m), dtype=torch.float64, device=dev)
    _lib.check(lib.bmc_center_rows(D.ptr(pd_), n, m, pd_.stride(0), D.ptr(td), D.ptr(mu), D.ptr(y), D.ptr(xc), m, st),
               "bmc_center_rows")
    used = method
    vt = s = None
    if method in ("auto", "gram", "gram_checked"):
        gram = torch.empty((m, m), dtype=torch.float64, device=dev)
        ws = torch.empty(max(int(lib.bmc_gram_workspace_bytes(n, m)), 8), dtype=torch.uint8, device=dev)
        _lib.check(lib.bmc_gram(D.ptr(xc), n, m, m, None, None, D.ptr(gram), D.ptr(ws), ws.numel(), st), "bmc_gram")
        if reduce is not None:
            gram = reduce(gram)
        lam, vec = torch.linalg.eigh(gram)                       # cuSOLVER, M-by-M
        lam = torch.flip(lam, dims=[0]).clamp_min(0.0)
        vec = torch.flip(vec, dims=[1])
        ratio = float(lam[k - 1] / lam[0]) if float(lam[0]) > 0 else 0.0
        if method == "gram" or ratio >= GRAM_RATIO_MIN:
            s, vt, used = torch.sqrt(lam), vec.t().contiguous(), "gram"
        elif method == "gram_checked" and not hasattr(reduce, "world"):
            import warnings
            warnings.warn(f"row-sharded orthogonalisation: lambda_K/lambda_1 = {ratio:.1e} of the Gram matrix is below "
                          f"{GRAM_RATIO_MIN:g}; its eigenvectors carry errors ~ eps*lambda_1/lambda_K and the reducer "
                          "offers no rank/world for the TSQR route", RuntimeWarning, stacklevel=3)
            s, vt, used = torch.sqrt(lam), vec.t().contiguous(), "gram"
    if vt is None and reduce is not None:
        s, vt = _tsqr_svd(xc, reduce)                             # graded spectrum, rows on several ranks
        used = "tsqr"
    if vt is None:
        # graded spectrum (or method="svd"): thin SVD of the centred matrix, still on the device
        _, s, vt = torch.linalg.svd(xc, full_matrices=False)
        used = "svd"
    s_hat = s[:k].contiguous()
    vt_norm = vt[:k].contiguous()
    vt_hat = (vt_norm / s_hat[:, None]).contiguous()              # inference_utils.py:166
    u_hat = torch.empty((n, k), dtype=torch.float64, device=dev)
    _lib.check(lib.bmc_project_rows(D.ptr(xc), n, m, m, None, D.ptr(vt_hat), k, D.ptr(u_hat), k, st),
               "bmc_project_rows")
    return dict(y=D.to_host(y), mu=D.to_host(mu), U_hat=D.to_host(u_hat), S_hat=D.to_host(s_hat),
                Vt_hat=D.to_host(vt_hat), Vt_hat_normalized=D.to_host(vt_norm), method=used)


# training options in the order upstream announces them (bmc.py:160-171); each default may depend on
# the instance (number of components, singular values)
_TRAIN_OPTIONS = (
    ("iterations", lambda self: 50000),
    ("sampler", lambda self: "gibbs_sampling"),
    ("burn", lambda self: 10000),
    ("stepsize", lambda self: 0.001),
    ("b_mean_prior", lambda self: np.zeros(self.U_hat.shape[1])),
    ("b_mean_cov", lambda self: np.diag(self.S_hat ** 2)),
    ("nu0_chosen", lambda self: 1.0),
    ("sigma20_chosen", lambda self: 0.02),
)
# GPU controls: optional, silent, defaults reproduce upstream's single fp64 chain
_GPU_OPTIONS = {"n_chains": 1, "seed": None, "dtype": "float64", "thin": 1}
_BANDS = (("Predicted_Lower", 0), ("Predicted_Median", 1), ("Predicted_Upper", 2))


def _filter_rows(df, domain_filter):
    """Row selection rules of ``evaluate`` (bmc.py:352-364): per column a callable on the column, a
    (low, high) tuple, a list of admissible values or a single value; the key "multi" with a callable
    is applied row-wise."""
    for column, rule in (domain_filter or {}).items():
        if callable(rule):
            mask = df.apply(rule, axis=1) if column == "multi" else rule(df[column])
        elif isinstance(rule, tuple) and len(rule) == 2:
            mask = df[column].between(*rule)
        elif isinstance(rule, list):
            mask = df[column].isin(rule)
        else:
            mask = df[column] == rule
        df = df[mask]
    return df


class BayesianModelCombination:
    """Bayesian model combination of several models' predictions (drop-in for ``pybmc.bmc``).

    Args:
        models_list (list[str]): model (column) names to combine.
        data_dict (dict[str, pandas.DataFrame]): one DataFrame per property.
        truth_column_name (str): column holding the ground truth.
        weights (list[float], optional): kept for compatibility (unused upstream as well).

    ``orthogonalize`` sets ``centered_experiment_train``, ``U_hat``, ``Vt_hat`` (right singular vectors
    divided by their singular values), ``S_hat``, ``Vt_hat_normalized``, ``_predictions_mean_train``,
    ``current_property``; ``train`` sets ``samples``.
    """

    def __init__(self, models_list, data_dict, truth_column_name, weights=None):
        names_ok = isinstance(models_list, list) and all(isinstance(m, str) for m in models_list)
        if not names_ok:
            raise ValueError("The 'models' should be a list of model names (strings) for Bayesian Combination.")
        frames_ok = isinstance(data_dict, dict) and all(isinstance(v, pd.DataFrame) for v in data_dict.values())
        if not frames_ok:
            raise ValueError("The 'data_dict' should be a dictionary of pandas DataFrames, one per property.")
        self.models_list = models_list
        self.data_dict = data_dict
        self.truth_column_name = truth_column_name
        self.weights = weights
        # only the literal name "truth" is dropped (bmc.py:75): a truth column with another name that is
        # listed in models_list is used as a model, exactly as upstream
        self.models = [name for name in models_list if name != "truth"]

    # -- orthogonalize ------------------------------------------------------------------------------
    def orthogonalize(self, property, train_df, components_kept, *, method="auto", device=None):
        """SVD-orthogonalise the centred training predictions on the GPU (bmc.py:79-130)."""
        self.current_property = property
        self.selected_models_dataset = self.data_dict[property].copy()
        result = orthogonalize_arrays(train_df[self.models].values, train_df[self.truth_column_name].values,
                                      components_kept, method=method, device=device)
        self._svd_method = result["method"]
        for attribute, key in (("centered_experiment_train", "y"), ("U_hat", "U_hat"), ("Vt_hat", "Vt_hat"),
                               ("S_hat", "S_hat"), ("Vt_hat_normalized", "Vt_hat_normalized"),
                               ("_predictions_mean_train", "mu")):
            setattr(self, attribute, result[key])

    # -- train ----------------------------------------------------------------------------------------
    def train(self, training_options=None):
        """Sample the posterior of the combination coefficients (bmc.py:132-193).

        ``training_options`` takes upstream's keys (iterations, sampler, burn, stepsize, b_mean_prior,
        b_mean_cov, nu0_chosen, sigma20_chosen); every key left out is announced with upstream's
        ``[INFO]`` line.  Extra keys ``n_chains``, ``seed``, ``dtype``, ``thin`` steer the GPU run."""
        given = dict(training_options or {})
        opt = {}
        for key, default in _TRAIN_OPTIONS:
            if key in given:
                opt[key] = given[key]
            else:
                opt[key] = default(self)
                print(f"[INFO] Using default value for '{key}': {opt[key]}")
        gpu = {key: given.get(key, default) for key, default in _GPU_OPTIONS.items()}
        variance_prior = [opt["nu0_chosen"], opt["sigma20_chosen"]]
        if opt["sampler"] == "simplex":
            self.samples = gibbs_sampler_simplex(self.centered_experiment_train, self.U_hat, self.Vt_hat, self.S_hat,
                                                 opt["iterations"], variance_prior, burn=opt["burn"],
                                                 stepsize=opt["stepsize"], **gpu)
        else:   # any other name selects the conjugate sampler, as upstream (bmc.py:187)
            self.samples = gibbs_sampler(self.centered_experiment_train, self.U_hat, opt["iterations"],
                                         [opt["b_mean_prior"], opt["b_mean_cov"], *variance_prior], **gpu)

    # -- predict --------------------------------------------------------------------------------------
    def _predict_frames(self, model_preds, vt_hat, domain_df, **kwargs):
        """Shared tail of ``predict`` / ``predict2``: draws + the three band frames (bmc.py:226-242)."""
        rndm_m, bands = rndm_m_random_calculator(model_preds, self.samples, vt_hat, **kwargs)
        domain_df = domain_df.reset_index(drop=True)
        frames = []
        for column, which in _BANDS:
            frame = domain_df.copy()
            frame[column] = bands[which]
            frames.append(frame)
        return (rndm_m, *frames)

    def _check_trained(self):
        # bmc.py:212-215 as written upstream.  Upstream's __init__ (bmc.py:73-77) never creates ``samples`` or
        # ``Vt_hat``, so on a fresh object this line raises AttributeError before the ValueError can -- the
        # de-facto behaviour, kept: neither attribute exists here either until orthogonalize() / train() ran.
        if self.samples is None or self.Vt_hat is None:
            raise ValueError("Must call `orthogonalize()` and `train()` before predicting.")

    def predict(self, X, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64", return_draws=True):
        """Posterior predictive draws and 2.5/50/97.5 % bands for the rows of ``X`` (bmc.py:195-242).

        Returns ``(rndm_m, lower_df, median_df, upper_df)``; ``return_draws=False`` skips building the
        ``[n_draws, N]`` matrix (``rndm_m`` is then None)."""
        self._check_trained()
        if not isinstance(X, pd.DataFrame):
            raise ValueError("X must be a pandas DataFrame containing model predictions and domain info.")
        domain = X[[c for c in X.columns if c not in self.models]]
        return self._predict_frames(X[self.models].values, self.Vt_hat, domain, n_draws=n_draws, seed=seed,
                                    dtype=dtype, return_draws=return_draws)

    def predict2(self, property, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64", return_draws=True):
        """``predict`` for a whole property table; models absent from that table are tolerated with a
        warning, unknown ones are an error (bmc.py:244-337)."""
        self._check_trained()
        if property not in self.data_dict:
            raise KeyError(f"Property '{property}' not found in data_dict.")
        table = self.data_dict[property].copy()
        present = [c for c in table.columns if c in self.models]
        trained_models_set, available_models_set = set(self.models), set(present)
        print(f"Available models: {available_models_set}")
        print(f"Trained models: {trained_models_set}")
        unknown = available_models_set - trained_models_set
        if unknown:
            raise ValueError(
                f"ERROR: Property '{property}' contains extra models not present during training: "
                f"{list(unknown)}. You must retrain if using a larger model space.")
        absent = trained_models_set - available_models_set
        if absent:
            print(f"WARNING: Predicting on property '{property}' with missing models: {list(absent)}")
            print("         The trained model weights include these models — prediction will proceed, "
                  "but results may not be statistically accurate.")
        if not present:
            raise ValueError("No available trained models are present in prediction DataFrame.")
        columns = [self.models.index(name) for name in present]
        domain = table[[c for c in table.columns if c not in self.models and c != self.truth_column_name]]
        return self._predict_frames(table[present].values, self.Vt_hat[:, columns], domain, n_draws=n_draws,
                                    seed=seed, dtype=dtype, return_draws=return_draws)

    # -- evaluate -------------------------------------------------------------------------------------
    def evaluate(self, domain_filter=None, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64"):
        """Coverage of the 0, 5, ..., 100 % credible intervals over the current property's table,
        optionally filtered (bmc.py:339-376).  The predictive matrix is never materialised: the fused
        kernel returns the two order counts per point that decide every level."""
        table = _filter_rows(self.data_dict[self.current_property], domain_filter)
        truth = np.asarray(table[self.truth_column_name].tolist(), dtype=np.float64)
        np.random.seed(142858)   # side effect of rndm_m_random_calculator upstream (sampling_utils.py:54)
        summary = predictive_summary(table[self.models].to_numpy(), self.samples, self.Vt_hat, truth=truth,
                                     n_draws=n_draws, seed=seed, dtype=dtype, return_draws=False)
        return coverage_from_counts(np.arange(0, 101, 5), summary.n_draws, summary.c_lt, summary.c_le)
