"""Linear mixed model with a random intercept per group
(``genetest/statistics/models/mixedlm.py``: ``sm.MixedLM(y, X, groups).fit(reml=...)``).

This test is NOT on the B200 hot path: the reference itself avoids fitting it per
marker (``genetest/scripts/cli.py:183-310``: the "optimized" MixedLM analysis fits the
model once without the SNP, runs a plain linear GWAS of the predicted random effects --
which here is the GPU path -- and refits the mixed model only for the markers that pass
a p-value threshold).  The fit below runs on the host, one model at a time, like the
reference's.

A random-intercept model ``y = X b + u_g + e`` has ``V = s2 (I + t Z Z')`` with one
variance ratio ``t = var(u) / s2``; the (restricted) likelihood profiled over ``b`` and
``s2`` is a function of ``t`` alone::

    V_g^-1 = (I - c_g 1 1') / s2,   c_g = t / (1 + t n_g),   log |V_g / s2| = log(1 + t n_g)
    b(t)   = (X' W X)^-1 X' W y,    W = I - blockdiag(c_g 1 1')
    REML:  -2 l = (n - p) log rss + sum_g log(1 + t n_g) + log |X' W X| + const
    ML:    -2 l =  n      log rss + sum_g log(1 + t n_g)                + const

so the optimum is found by a bounded scalar search, polished on the score equation,
instead of statsmodels' BFGS over its Cholesky parameterisation: the same estimates (the
likelihood is the same function), within 0.03 units of the stated decimal place of every
R / lme4 number the reference's ``test_mixedlm.py`` asserts (``tests/test_mixedlm.py``).  Standard errors are those of generalised least squares at
the optimum, ``s2 (X' W X)^-1``; z statistics, normal p-values and confidence intervals
as in ``MixedLMResults``; the random effects are the BLUPs ``c_g sum_{i in g} r_i``.
"""

import logging

import numpy as np
import pandas as pd

from ..core import StatsError, StatsModels

__all__ = ["StatsMixedLM"]

logger = logging.getLogger(__name__)

_Z975 = 1.959963984540054


def _norm_sf2(z):
    """2 * P(Z > |z|)."""
    from math import erfc, sqrt
    return erfc(abs(z) / sqrt(2.0))


class _Profile(object):
    """Sufficient statistics of a random-intercept model: everything the profiled
    likelihood needs at a given variance ratio is O(groups x p^2)."""

    def __init__(self, y, X, groups):
        self.n, self.p = X.shape
        labels, inv = np.unique(groups, return_inverse=True)
        self.labels = labels
        G = labels.shape[0]
        self.ng = np.bincount(inv, minlength=G).astype(float)
        self.S = np.zeros((G, self.p))
        np.add.at(self.S, inv, X)                      # sum of the rows of X per group
        self.t = np.bincount(inv, weights=y, minlength=G)
        self.XtX = X.T @ X
        self.Xty = X.T @ y
        self.yty = float(y @ y)

    def at(self, theta):
        c = theta / (1.0 + theta * self.ng)
        A = self.XtX - (self.S * c[:, None]).T @ self.S            # X' W X
        b = self.Xty - self.S.T @ (c * self.t)                      # X' W y
        beta = np.linalg.solve(A, b)
        rss = self.yty - float(np.sum(c * self.t ** 2)) - float(beta @ b)
        return c, A, beta, rss

    def score(self, theta, reml):
        """d criterion / d theta (envelope theorem: beta and the scale are at their optima)."""
        c, A, beta, rss = self.at(theta)
        dc = 1.0 / (1.0 + theta * self.ng) ** 2                     # d c_g / d theta
        rg = self.t - self.S @ beta                                 # sum of the residuals per group
        dof = self.n - self.p if reml else self.n
        val = -dof * float(np.sum(dc * rg ** 2)) / rss + float(np.sum(self.ng / (1.0 + theta * self.ng)))
        if reml:
            Ainv_S = np.linalg.solve(A, self.S.T)                   # p x G
            val -= float(np.sum(dc * np.einsum("gp,pg->g", self.S, Ainv_S)))
        return val

    def criterion(self, theta, reml):
        """-2 log-likelihood up to its constant."""
        c, A, beta, rss = self.at(theta)
        if not rss > 0.0:
            return np.inf
        dof = self.n - self.p if reml else self.n
        val = dof * np.log(rss) + float(np.sum(np.log1p(theta * self.ng)))
        if reml:
            sign, ld = np.linalg.slogdet(A)
            if sign <= 0:
                return np.inf
            val += ld
        return val


def fit_random_intercept(y, X, groups, reml=True):
    """Estimates of ``sm.MixedLM(y, X, groups).fit(reml=reml)``.

    Returns a dict: ``params``, ``bse``, ``scale`` (residual variance), ``cov_re``
    (variance of the random intercept), ``llf``, ``n_groups``, ``random_effects``
    (group label -> BLUP)."""
    from scipy.optimize import minimize_scalar
    y = np.asarray(y, dtype=float).ravel()
    X = np.asarray(X, dtype=float)
    prof = _Profile(y, X, np.asarray(groups))
    n, p = prof.n, prof.p
    if n <= p:
        raise np.linalg.LinAlgError("Singular matrix")
    # theta = exp(s): a coarse scan brackets the minimum, Brent refines it; the boundary
    # theta = 0 (no group effect) is compared at the end
    grid = np.linspace(-18.0, 18.0, 73)
    vals = np.array([prof.criterion(np.exp(s), reml) for s in grid])
    k = int(np.argmin(vals))
    lo, hi = grid[max(k - 1, 0)], grid[min(k + 1, len(grid) - 1)]
    res = minimize_scalar(lambda s: prof.criterion(np.exp(s), reml), bounds=(lo, hi),
                          method="bounded", options={"xatol": 1e-12, "maxiter": 500})
    theta, best = float(np.exp(res.x)), float(res.fun)
    if prof.criterion(0.0, reml) <= best:
        theta = 0.0
    else:
        # polish on the score equation: the minimum of a flat function is only located to
        # ~1e-7, a sign change of its derivative to machine precision
        from scipy.optimize import brentq
        a, b = theta * (1.0 - 1e-4), theta * (1.0 + 1e-4)
        fa, fb = prof.score(a, reml), prof.score(b, reml)
        if fa * fb < 0.0:
            theta = float(brentq(lambda t: prof.score(t, reml), a, b, xtol=1e-14 * theta, rtol=1e-14))
    c, A, beta, rss = prof.at(theta)
    dof = (n - p) if reml else n
    scale = rss / dof
    cov = scale * np.linalg.inv(A)
    bse = np.sqrt(np.diag(cov))
    # log-likelihood with statsmodels' constants (MixedLM.loglike with the scale profiled out)
    llf = -0.5 * float(np.sum(np.log1p(theta * prof.ng))) - 0.5 * dof * np.log(rss)
    if reml:
        llf -= 0.5 * np.linalg.slogdet(A)[1]
    llf += -0.5 * dof * np.log(2.0 * np.pi) + 0.5 * (dof * np.log(dof) - dof)
    blup = c * (prof.t - prof.S @ beta)
    return {"params": beta, "bse": bse, "scale": scale, "cov_re": theta * scale, "llf": float(llf),
            "n_groups": int(prof.labels.shape[0]),
            "random_effects": pd.Series(blup, index=prof.labels)}


class StatsMixedLM(StatsModels):
    def __init__(self, reml=True):
        """Initializes a 'StatsMixedLM' instance.

        Args:
            reml (bool): Whether to use REML or ML for the test.

        """
        self._reml = reml

    @staticmethod
    def _prepare_data(y, X):
        """``outcome`` and ``groups`` columns of y (``mixedlm.py:42-57``)."""
        missing_cols = {"outcome", "groups"} - set(y.columns)
        if missing_cols:
            raise ValueError(
                "missing column in y: mixedlm requires 'outcome' and 'groups'"
            )
        if len(y.columns) > 2:
            extra = set(y.columns) - {"outcome", "groups"}
            logger.warning("{}: unknown column in y, will be ignored".format(
                ",".join(extra),
            ))
        return y[["outcome"]], X, y.groups

    def fit(self, y, X):
        """Fit the model; same dict as the reference (``mixedlm.py:66-117``)."""
        y, X, groups = self._prepare_data(y, X)
        try:
            fitted = fit_random_intercept(y.values[:, 0], X.values, groups.values, reml=self._reml)
        except np.linalg.LinAlgError as e:
            raise StatsError(str(e))
        out = {
            "MODEL": {
                "log_likelihood": fitted["llf"],
                "nobs": X.shape[0],
                "n_groups": fitted["n_groups"],
                "random_effects": fitted["random_effects"],
            },
        }
        for j, param in enumerate(X.columns):
            coef, se = float(fitted["params"][j]), float(fitted["bse"][j])
            z = coef / se if se > 0 else float("nan")
            pval = _norm_sf2(z) if z == z else float("nan")
            # If GWAS, check that inference could be done on the SNP
            if param == "SNPs" and np.isnan(pval):
                raise StatsError("Inference did not converge.")
            out[param] = {
                "coef": coef, "std_err": se,
                "lower_ci": coef - _Z975 * se, "upper_ci": coef + _Z975 * se,
                "z_value": z, "p_value": pval,
            }
        return out
