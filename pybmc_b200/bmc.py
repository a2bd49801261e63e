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


def orthogonalize_arrays(preds, truth, components_kept, method="auto", device=None):
    """Centre the model predictions per point and project on the truncated SVD basis
    (pybmc/bmc.py:102-130 + pybmc/inference_utils.py:147-168) on the device.

    Kernels: ``bmc_center_rows`` (mu, y, Xc), ``bmc_gram`` (Xc'Xc), cuSOLVER for the small
    M-by-M eigenproblem (or the thin SVD when the spectrum is too graded for the Gram route),
    ``bmc_project_rows`` (U_hat = Xc Vt_hat').  The reference's n-by-n U is never formed.

    Returns dict(y, mu, U_hat, S_hat, Vt_hat, Vt_hat_normalized, method).
    """
    lib = _lib.load()
    dev = D.device(device)
    preds = np.asarray(preds, dtype=np.float64)
    if preds.ndim != 2:
        raise ValueError("model predictions must be [n_points, n_models]")
    n, m = preds.shape
    k = int(components_kept)
    if k < 1 or k > min(n, m):
        raise IndexError(f"components_kept={k} outside 1..{min(n, m)}")   # upstream: IndexError from U.T[i]
    pd_ = D.to_device(preds, dev)
    td = D.to_device(np.asarray(truth, dtype=np.float64).reshape(-1), dev)
    st = D.stream_ptr(dev)
    mu = torch.empty(n, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    xc = torch.empty((n, m), dtype=torch.float64, device=dev)
    _lib.check(lib.bmc_center_rows(D.ptr(pd_), n, m, pd_.stride(0), D.ptr(td), D.ptr(mu), D.ptr(y), D.ptr(xc), m, st),
               "bmc_center_rows")
    used = method
    vt = s = None
    if method in ("auto", "gram"):
        gram = torch.empty((m, m), dtype=torch.float64, device=dev)
        ws = torch.empty(max(int(lib.bmc_gram_workspace_bytes(n, m)), 8), dtype=torch.uint8, device=dev)
        _lib.check(lib.bmc_gram(D.ptr(xc), n, m, m, None, None, D.ptr(gram), D.ptr(ws), ws.numel(), st), "bmc_gram")
        lam, vec = torch.linalg.eigh(gram)                       # cuSOLVER, M-by-M
        lam = torch.flip(lam, dims=[0]).clamp_min(0.0)
        vec = torch.flip(vec, dims=[1])
        ratio = float(lam[k - 1] / lam[0]) if float(lam[0]) > 0 else 0.0
        if method == "gram" or ratio >= GRAM_RATIO_MIN:
            s, vt, used = torch.sqrt(lam), vec.t().contiguous(), "gram"
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


class BayesianModelCombination:
    """Bayesian model combination of several models' predictions.

    Args:
        models_list (list[str]): model (column) names to combine.
        data_dict (dict[str, pandas.DataFrame]): one DataFrame per property.
        truth_column_name (str): column holding the ground truth.
        weights (list[float], optional): initial weights (kept for compatibility; unused upstream too).

    Attributes set by ``orthogonalize``: ``centered_experiment_train``, ``U_hat``, ``Vt_hat`` (right
    singular vectors divided by the singular values), ``S_hat``, ``Vt_hat_normalized``,
    ``_predictions_mean_train``, ``current_property``; by ``train``: ``samples``.
    """

    def __init__(self, models_list, data_dict, truth_column_name, weights=None):
        if not isinstance(models_list, list) or not all(isinstance(m, str) for m in models_list):
            raise ValueError("The 'models' should be a list of model names (strings) for Bayesian Combination.")
        if not isinstance(data_dict, dict) or not all(isinstance(df, pd.DataFrame) for df in data_dict.values()):
            raise ValueError("The 'data_dict' should be a dictionary of pandas DataFrames, one per property.")
        self.data_dict = data_dict
        self.models_list = models_list
        # only the literal name "truth" is dropped (bmc.py:75); a truth column with another name
        # that is listed in models_list is used as a model, as upstream
        self.models = [m for m in models_list if m != "truth"]
        self.weights = weights if weights is not None else None
        self.truth_column_name = truth_column_name

    # -- orthogonalize ------------------------------------------------------------------------
    def orthogonalize(self, property, train_df, components_kept, *, method="auto", device=None):
        """SVD-orthogonalise the centred training predictions (bmc.py:79-130)."""
        self.current_property = property
        self.selected_models_dataset = self.data_dict[property].copy()
        preds = train_df[self.models].values
        truth = train_df[self.truth_column_name].values
        r = orthogonalize_arrays(preds, truth, components_kept, method=method, device=device)
        self.centered_experiment_train = r["y"]
        self.U_hat = r["U_hat"]
        self.Vt_hat = r["Vt_hat"]
        self.S_hat = r["S_hat"]
        self.Vt_hat_normalized = r["Vt_hat_normalized"]
        self._predictions_mean_train = r["mu"]
        self._svd_method = r["method"]

    # -- train ------------------------------------------------------------------------------------
    def train(self, training_options=None):
        """Sample the posterior of the combination coefficients (bmc.py:132-193).

        ``training_options`` keys as upstream (iterations, sampler, burn, stepsize, b_mean_prior,
        b_mean_cov, nu0_chosen, sigma20_chosen), each announced with an ``[INFO]`` line when
        defaulted; optional GPU keys ``n_chains``, ``seed``, ``dtype``, ``thin`` are silent.
        """
        if training_options is None:
            training_options = {}

        def get_option(key, default):
            if key not in training_options:
                print(f"[INFO] Using default value for '{key}': {default}")
            return training_options.get(key, default)

        iterations = get_option("iterations", 50000)
        sampler = get_option("sampler", "gibbs_sampling")
        burn = get_option("burn", 10000)
        stepsize = get_option("stepsize", 0.001)
        num_components = self.U_hat.shape[1]
        b_mean_prior = get_option("b_mean_prior", np.zeros(num_components))
        b_mean_cov = get_option("b_mean_cov", np.diag(self.S_hat ** 2))
        nu0_chosen = get_option("nu0_chosen", 1.0)
        sigma20_chosen = get_option("sigma20_chosen", 0.02)
        extra = dict(n_chains=training_options.get("n_chains", 1), seed=training_options.get("seed"),
                     dtype=training_options.get("dtype", "float64"), thin=training_options.get("thin", 1))
        if sampler == "simplex":
            self.samples = gibbs_sampler_simplex(self.centered_experiment_train, self.U_hat, self.Vt_hat, self.S_hat,
                                                 iterations, [nu0_chosen, sigma20_chosen], burn=burn,
                                                 stepsize=stepsize, **extra)
        else:  # any other string selects the conjugate sampler, as upstream (bmc.py:187)
            self.samples = gibbs_sampler(self.centered_experiment_train, self.U_hat, iterations,
                                         [b_mean_prior, b_mean_cov, nu0_chosen, sigma20_chosen], **extra)

    # -- predict ----------------------------------------------------------------------------------
    def _require_trained(self):
        if getattr(self, "samples", None) is None or getattr(self, "Vt_hat", None) is None:
            raise ValueError("Must call `orthogonalize()` and `train()` before predicting.")

    @staticmethod
    def _frames(domain_df, lower, median, upper):
        lower_df = domain_df.copy()
        lower_df["Predicted_Lower"] = lower
        median_df = domain_df.copy()
        median_df["Predicted_Median"] = median
        upper_df = domain_df.copy()
        upper_df["Predicted_Upper"] = upper
        return lower_df, median_df, upper_df

    def predict(self, X, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64", return_draws=True):
        """Posterior predictive draws and 2.5/50/97.5 % bands for the rows of ``X`` (bmc.py:195-242).

        Returns ``(rndm_m, lower_df, median_df, upper_df)``; ``return_draws=False`` skips building the
        ``[n_draws, N]`` matrix (``rndm_m`` is then None).
        """
        self._require_trained()
        if not isinstance(X, pd.DataFrame):
            raise ValueError("X must be a pandas DataFrame containing model predictions and domain info.")
        domain_keys = [c for c in X.columns if c not in self.models]
        rndm_m, (lower, median, upper) = rndm_m_random_calculator(
            X[self.models].values, self.samples, self.Vt_hat, n_draws=n_draws, seed=seed, dtype=dtype,
            return_draws=return_draws)
        domain_df = X[domain_keys].reset_index(drop=True)
        return (rndm_m, *self._frames(domain_df, lower, median, upper))

    def predict2(self, property, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64", return_draws=True):
        """Same as ``predict`` for a whole property table, tolerating missing models (bmc.py:244-337)."""
        self._require_trained()
        if property not in self.data_dict:
            raise KeyError(f"Property '{property}' not found in data_dict.")
        df = self.data_dict[property].copy()
        domain_keys = [c for c in df.columns if c not in self.models and c != self.truth_column_name]
        available_models = [m for m in df.columns if m in self.models]
        trained_models_set = set(self.models)
        available_models_set = set(available_models)
        missing_models = trained_models_set - available_models_set
        extra_models = available_models_set - trained_models_set
        print(f"Available models: {available_models_set}")
        print(f"Trained models: {trained_models_set}")
        if len(extra_models) > 0:
            raise ValueError(
                f"ERROR: Property '{property}' contains extra models not present during training: "
                f"{list(extra_models)}. You must retrain if using a larger model space.")
        if len(missing_models) > 0:
            print(f"WARNING: Predicting on property '{property}' with missing models: {list(missing_models)}")
            print("         The trained model weights include these models — prediction will proceed, "
                  "but results may not be statistically accurate.")
        if len(available_models) == 0:
            raise ValueError("No available trained models are present in prediction DataFrame.")
        model_indices = [self.models.index(m) for m in available_models]
        vt_reduced = self.Vt_hat[:, model_indices]
        rndm_m, (lower, median, upper) = rndm_m_random_calculator(
            df[available_models].values, self.samples, vt_reduced, n_draws=n_draws, seed=seed, dtype=dtype,
            return_draws=return_draws)
        domain_df = df[domain_keys].reset_index(drop=True)
        return (rndm_m, *self._frames(domain_df, lower, median, upper))

    # -- evaluate ---------------------------------------------------------------------------------
    def evaluate(self, domain_filter=None, *, n_draws=DEFAULT_DRAWS, seed=None, dtype="float64"):
        """Coverage of the 0, 5, ..., 100 % credible intervals over the current property's table,
        optionally filtered (bmc.py:339-376).  The predictive matrix is never materialised: the
        fused kernel returns the two order counts per point that decide every level."""
        df = self.data_dict[self.current_property]
        if domain_filter:
            for col, cond in domain_filter.items():
                if col == "multi" and callable(cond):
                    df = df[df.apply(cond, axis=1)]
                elif callable(cond):
                    df = df[cond(df[col])]
                elif isinstance(cond, tuple) and len(cond) == 2:
                    df = df[df[col].between(*cond)]
                elif isinstance(cond, list):
                    df = df[df[col].isin(cond)]
                else:
                    df = df[df[col] == cond]
        preds = df[self.models].to_numpy()
        truth = np.asarray(df[self.truth_column_name].tolist(), dtype=np.float64)
        np.random.seed(142858)   # side effect of rndm_m_random_calculator upstream (sampling_utils.py:54)
        res = predictive_summary(preds, self.samples, self.Vt_hat, truth=truth, n_draws=n_draws, seed=seed,
                                 dtype=dtype, return_draws=False)
        return coverage_from_counts(np.arange(0, 101, 5), res.n_draws, res.c_lt, res.c_le)
