"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, fp64) of pyBMC's inference path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` leg may import this module; the shipped package
``pybmc_b200`` never does and fails loudly when its CUDA library is missing.

Every function names the reference lines it restates (paths relative to the
upstream tree, ``pybmc/...``).  The arithmetic boundary of the reference is NumPy
(pinned 2.3.2 in the upstream ``poetry.lock``; 2.3.5 in this image): ``svd``,
``inv``, ``multivariate_normal``, ``Generator.gamma``, ``percentile``, ``sort``.

PINNING.  The reference's own tests hold no golden numbers for this path
(shape/type/exception checks only), so the oracle is pinned against the
reference ITSELF: ``tests/golden/make_golden.py`` imports the unmodified upstream
package in the authoring container, makes its unseedable draws deterministic by
swapping ``numpy.random.default_rng`` for a seeded factory, and stores inputs'
seeds + outputs under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
requires this module to reproduce those files bit for bit (same NumPy calls in
the same order), which is what licenses using it as the checker for the CUDA
path.  The Philox-driven mode below (``PhiloxDraws``) runs the same algorithm on
the device's random-number contract so a GPU chain can be checked value by value.
"""
import numpy as np

from . import philox as px

RIDGE = 1e-6        # inference_utils.py:41  (added to the precision matrix)
SIGMA2_FLOOR = 1e-6  # inference_utils.py:37,52


# --------------------------------------------------------------------------- #
# draw providers                                                              #
# --------------------------------------------------------------------------- #
class NumpyDraws:
    """The reference's own random call sites.

    * coefficient draw / proposal: legacy global ``np.random.multivariate_normal``
      (inference_utils.py:45,98,121)
    * variance draw: ``np.random.default_rng().gamma`` on a generator built anew
      for every call (inference_utils.py:52,117,140); ``rng_factory`` stands in for
      ``np.random.default_rng`` so that the golden files can seed it
    * Metropolis uniform: legacy global ``np.random.uniform`` (:110,132)
    """

    def __init__(self, rng_factory=None):
        self.rng_factory = rng_factory or np.random.default_rng

    def start(self, it):
        pass

    def gaussian(self, mean, cov):
        return np.random.multivariate_normal(mean, cov)

    def gamma(self, shape, scale):
        return self.rng_factory().gamma(shape, scale)

    def uniform(self):
        return np.random.uniform()


class PhiloxDraws:
    """Device random-number contract (oracle/philox.py) behind the same interface.

    ``factor(cov) -> F`` with ``F @ F.T == cov`` decides how the K standard
    normals are coloured; any valid factor samples the reference's N(mean, cov).
    ``factor=None`` uses the lower Cholesky factor.
    """

    def __init__(self, seed, chain, tag, factor=None):
        self.key = px.seed_key(seed)
        self.chain = int(chain)
        self.tag = int(tag)
        self.factor = factor
        self.it = 0
        self.k = None          # component count of the sampler (set by the first gaussian draw)

    def start(self, it):
        self.it = int(it)

    def gaussian(self, mean, cov):
        k = len(mean)
        self.k = k
        z = np.array(px.normal_vector(k, self.it, self.chain, self.tag, self.key))
        f = np.linalg.cholesky(cov) if self.factor is None else self.factor(cov)
        # a wrong factor would silently sample another law: refuse it
        if not np.allclose(f @ f.T, cov, rtol=1e-9, atol=1e-12 * np.abs(cov).max()):
            raise AssertionError("factor does not reproduce the covariance")
        return np.asarray(mean, dtype=float) + f @ z

    def gamma(self, shape, scale):
        return scale * px.gamma_unit_scale(shape, self.it, self.chain, self.tag, self.key, self.k)

    def uniform(self):
        return px.metropolis_uniform(self.it, self.chain, self.tag, self.key)


# --------------------------------------------------------------------------- #
# orthogonalisation                                                            #
# --------------------------------------------------------------------------- #
def truncate_svd(u, s, vt, keep):
    """inference_utils.py:147-168.  Note the naming: the third output is Vt/S."""
    # :164 -- the kept columns are gathered row-wise and transposed back, so U_hat is
    # column-major in memory; BLAS rounding in X'X depends on that layout
    u_hat = np.asarray(u).T[:keep].copy().T
    s_hat = s[:keep]                                                      # :165
    vt_scaled = np.array([vt[i] / s[i] for i in range(keep)])             # :166
    vt_plain = np.array([vt[i] for i in range(keep)])                     # :167
    return u_hat, s_hat, vt_scaled, vt_plain


def orthogonalize_arrays(preds, truth, keep, full_matrices=True):
    """bmc.py:102-130 on bare arrays.

    ``full_matrices=False`` is the thin SVD used where the reference's n-by-n U
    cannot be formed; the kept columns are the same up to rounding.
    Returns dict(y, mu, U_hat, S_hat, Vt_hat, Vt_hat_normalized).
    """
    preds = np.asarray(preds)
    mu = np.mean(preds, axis=1)                                           # :106
    y = np.asarray(truth) - mu                                            # :109-111
    xc = preds - mu[:, None]                                              # :114-116
    u, s, vt = np.linalg.svd(xc, full_matrices=full_matrices)             # :119
    u_hat, s_hat, vt_hat, vt_norm = truncate_svd(u, s, vt, keep)          # :122
    return dict(y=y, mu=mu, U_hat=u_hat, S_hat=s_hat, Vt_hat=vt_hat,
                Vt_hat_normalized=vt_norm)


# --------------------------------------------------------------------------- #
# samplers                                                                     #
# --------------------------------------------------------------------------- #
def gibbs_conjugate(y, X, iterations, prior, draws=None):
    """inference_utils.py:4-56: Normal / inverse-gamma Gibbs for y = X b + e.

    Rows of the result are [b_0 .. b_{K-1}, sigma] (sigma = sqrt of the variance
    draw, :54); there is no burn-in and the first iterate is kept (:39-56).
    """
    draws = draws or NumpyDraws()
    b0, B0, nu0, sigma20 = prior
    lam = np.linalg.inv(B0)                                               # :22
    n = len(y)
    gram = X.T.dot(X)                                                     # :25
    gram_inv = np.linalg.inv(gram)                                        # :26
    b_ols = gram_inv.dot(X.T).dot(y)                                      # :28
    res = y - X.dot(b_ols)                                                # :29-30
    sigma2 = np.sum(res ** 2) / len(res)                                  # :31
    sigma2 = max(sigma2, SIGMA2_FLOOR)                                    # :37
    out = []
    for it in range(iterations):                                          # :39
        draws.start(it)
        cov = np.linalg.inv(gram / sigma2 + lam + np.eye(gram.shape[0]) * RIDGE)  # :41
        mean = cov.dot(lam.dot(b0) + X.T.dot(y) / sigma2)                 # :42-44
        b = draws.gaussian(mean, cov)                                     # :45
        res = y - X.dot(b)                                                # :48-49
        shape_post = (nu0 + n) / 2.0                                      # :50
        scale_post = (nu0 * sigma20 + np.sum(res ** 2)) / 2.0             # :51
        sigma2 = max(1 / draws.gamma(shape_post, 1 / scale_post), SIGMA2_FLOOR)   # :52
        out.append(np.append(b, np.sqrt(sigma2)))                         # :54
    return np.array(out)                                                  # :56


def gibbs_simplex(y, X, Vt_hat, S_hat, iterations, prior, burn=10000,
                  stepsize=0.001, draws=None, return_acceptance=False):
    """inference_utils.py:59-144: random-walk Metropolis inside Gibbs with the
    model weights w = b Vt_hat + 1/M kept non-negative.

    Reference behaviours kept on purpose: acceptance ratio exp((ll'-ll)/sigma2)
    with ll = -RSS, i.e. no factor 1/2 (:108,130); sigma2 is redrawn even when the
    proposal is skipped (:102,115-117); start at b = 0, sigma2 = y'y/n (:82-86);
    no floor on sigma2; burn-in acceptances are not counted (:110-112).
    """
    draws = draws or NumpyDraws()
    n_models = len(Vt_hat.T)
    bias0 = np.full(n_models, 1 / n_models)                               # :78
    nu0, sigma20 = prior                                                  # :79
    step_cov = np.diag(S_hat ** 2 * stepsize ** 2)                        # :80
    n = len(y)
    b = np.full(len(X.T), 0)                                              # :82
    ll = -np.sum((y - X.dot(b)) ** 2)                                     # :83-85
    sigma2 = -ll / n                                                      # :86
    out = []
    accepted = 0
    if burn < 0:                                                          # :91-92
        raise ValueError("Burn-in iterations must be non-negative.")
    if stepsize <= 0:                                                     # :93-94
        raise ValueError("Stepsize must be positive.")
    for it in range(burn + iterations):                                   # :97, :120
        draws.start(it)
        keep = it >= burn
        prop = draws.gaussian(b, step_cov)                                # :98 / :121
        w = np.dot(prop, Vt_hat) + bias0                                  # :99 / :122
        if not np.any(w < 0):                                             # :102 / :124
            ll_prop = -np.sum((y - X.dot(prop)) ** 2)                     # :103-105
            alpha = min(1, np.exp((ll_prop - ll) / sigma2))               # :106-109
            if draws.uniform() < alpha:                                   # :110 / :132
                b = np.copy(prop)
                ll = ll_prop
                accepted += 1 if keep else 0                              # :135
        shape_post = (nu0 + n) / 2.0                                      # :115 / :138
        scale_post = (nu0 * sigma20 - ll) / 2.0                           # :116 / :139
        sigma2 = 1 / draws.gamma(shape_post, 1 / scale_post)              # :117 / :140
        if keep:
            out.append(np.append(b, np.sqrt(sigma2)))                     # :141
    res = np.array(out)
    return (res, accepted) if return_acceptance else res


# --------------------------------------------------------------------------- #
# posterior prediction and coverage                                            #
# --------------------------------------------------------------------------- #
def predictive_from_selected(preds, theta, Vt_hat, noise):
    """sampling_utils.py:60-82 once the S rows ``theta`` and the S-by-N standard
    normals ``noise`` are fixed.  Returns (rndm_m, [lo, med, hi])."""
    betas = theta[:, :-1]                                                 # :60
    sig = theta[:, -1]                                                    # :61
    w = betas @ Vt_hat + np.full(Vt_hat.shape[1], 1 / Vt_hat.shape[1])    # :64-67
    yv = w @ preds.T                                                      # :70-72
    rndm_m = yv + noise * sig[:, None]                                    # :76-77
    lo = np.percentile(rndm_m, 2.5, axis=0)                               # :80
    med = np.percentile(rndm_m, 50, axis=0)                               # :81
    hi = np.percentile(rndm_m, 97.5, axis=0)                              # :82
    return rndm_m, [lo, med, hi]


def predictive_draws(preds, samples, Vt_hat, rng, n_draws=10000):
    """sampling_utils.py:40-84 with the generator passed in (the reference builds
    an unseeded one, :55; its ``np.random.seed(142858)`` on :54 does not reach it).
    ``n_draws`` is 10000 in the reference (:57)."""
    theta = rng.choice(samples, n_draws, replace=False)                   # :57
    noise = rng.standard_normal((n_draws, preds.shape[0]))                # :76
    return predictive_from_selected(preds, theta, Vt_hat, noise)


def coverage_indices(percentiles, n_draws):
    """sampling_utils.py:30-31 -- the float expression truncates unevenly, so it
    is evaluated exactly as written, on the element type of ``percentiles``."""
    lo = [int((0.5 - p / 200) * n_draws) for p in percentiles]
    hi = [int((0.5 + p / 200) * n_draws) - 1 for p in percentiles]
    return lo, hi


def coverage_levels(percentiles, rndm_m, truth):
    """sampling_utils.py:4-37 with the truth column already extracted."""
    n_points = len(rndm_m.T)                                              # :18
    n_draws = len(rndm_m)                                                 # :19
    truth = list(truth)                                                   # :20
    lo, hi = coverage_indices(percentiles, n_draws)
    cols = np.sort(rndm_m.T, axis=1)                                      # :28 (once, not 21x)
    res = []
    for j, _ in enumerate(percentiles):                                   # :24
        covered = 0
        for i in range(n_points):                                         # :26
            if cols[i][lo[j]] <= truth[i] <= cols[i][hi[j]]:              # :33
                covered += 1
        res.append(covered / n_points * 100)                              # :35
    return res


def order_counts(rndm_m, truth):
    """#(x < t) and #(x <= t) per column -- the two integers from which the
    sort-free form of sampling_utils.py:33 is decided."""
    t = np.asarray(truth)[None, :]
    return (rndm_m < t).sum(axis=0).astype(np.int64), (rndm_m <= t).sum(axis=0).astype(np.int64)


def coverage_from_counts(percentiles, n_draws, c_lt, c_le):
    """sorted[l] <= t <= sorted[u]  <=>  #(x <= t) >= l+1  and  #(x < t) <= u."""
    lo, hi = coverage_indices(percentiles, n_draws)
    n_points = len(c_lt)
    out = []
    for l, u in zip(lo, hi):
        covered = int(np.sum((c_le >= l + 1) & (c_lt <= u)))
        out.append(covered / n_points * 100)
    return out


# --------------------------------------------------------------------------- #
# data split by distance                                                       #
# --------------------------------------------------------------------------- #
def distance_classes(points, refs, distance1, distance2):
    """data.py:194-245: indices of ``points`` whose nearest reference point is within ``distance1``,
    within ``distance2`` only, or farther.  One norm per pair, `<=`, as the reference's double loop."""
    points = np.asarray(points, dtype=float)
    refs = np.asarray(refs, dtype=float)
    near, mid, far = [], [], []
    for i, p in enumerate(points):
        dist = np.array([np.linalg.norm(p - q) for q in refs])           # :218-221, :228-231
        if np.any(dist <= distance1):
            near.append(i)                                               # :223-225
        elif np.any(dist <= distance2):
            mid.append(i)                                                # :233-235
        else:
            far.append(i)                                                # :237-238
    return near, mid, far


# --------------------------------------------------------------------------- #
# sufficient-statistic identities used by the device kernels (checked in tests)  #
# --------------------------------------------------------------------------- #
def simultaneous_diagonalisation(gram, lam):
    """W with W'(lam + RIDGE I)W = I and W' gram W = diag(d), so that
    inv(gram/s2 + lam + RIDGE I) = W diag(1/(d/s2 + 1)) W'   (cf. :41)."""
    k = gram.shape[0]
    low = np.linalg.cholesky(lam + RIDGE * np.eye(k))
    linv = np.linalg.inv(low)
    d, q = np.linalg.eigh(linv @ gram @ linv.T)
    return linv.T @ q, d


def moments(samples):
    """Posterior summaries compared against device sufficient statistics."""
    return dict(mean=samples.mean(axis=0), cov=np.cov(samples.T, ddof=0).reshape(
        samples.shape[1], samples.shape[1]))
