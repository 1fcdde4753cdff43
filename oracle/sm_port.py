"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product.

numpy restatement of the three statsmodels estimators that own the arithmetic
of genetest's per-SNP hot path.  statsmodels is a third-party dependency that
is NOT under /root/reference (``setup.py:102-105`` pins ``statsmodels>=0.8.0``,
no lock file; CI era <= 0.13), so its published algorithms are restated here
and parity is anchored on the reference's own call sites and golden vectors:

* ``ols_pinv``           <- ``sm.OLS(y, X).fit()``
                            (call site ``genetest/statistics/models/linear.py:102-104``,
                            fields read at ``linear.py:50-88``)
* ``glm_binomial_irls``  <- ``sm.GLM(y, X, family=Binomial()).fit()``
                            (``genetest/statistics/models/logistic.py:38-40``, fields ``:49-72``)
* ``phreg_efron_newton`` <- ``sm.PHReg(tte, X, status=event, ties="efron").fit()``
                            (``genetest/statistics/models/survival.py:77-82``, fields ``:100-120``)

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks these functions
against every SAS / R number asserted by the reference's
``tests/test_linear.py``, ``test_logistic.py``, ``test_coxph.py`` and
``test_gwaswriter.py`` (extracted into ``tests/golden/expected.json``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.
"""

import numpy as np
from scipy import stats

FLOAT_EPS = np.finfo(float).eps


class OracleLinAlgError(Exception):
    """Stands in for ``numpy.linalg.LinAlgError`` raised inside statsmodels."""


class OraclePerfectSeparation(Exception):
    """statsmodels (<0.14) ``PerfectSeparationError``; message pinned by
    ``genetest/tests/test_logistic.py:209,661``."""
    def __init__(self):
        super().__init__("Perfect separation detected, results not available")


class OracleConvergence(Exception):
    """statsmodels ``ConvergenceWarning`` promoted to an error
    (``survival.py:80-88``)."""


# --------------------------------------------------------------------------
# OLS, method="pinv"
# --------------------------------------------------------------------------
def _pinv_extended(x, rcond=1e-15):
    """statsmodels.tools.tools.pinv_extended: SVD pseudo-inverse that also
    returns the singular values."""
    u, s, vt = np.linalg.svd(x, full_matrices=False)
    s_orig = s.copy()
    cutoff = rcond * (s.max() if s.size else 0.0)
    with np.errstate(divide="ignore"):
        s_inv = np.where(s > cutoff, 1.0 / s, 0.0)
    return (vt.T * s_inv) @ u.T, s_orig


def _has_constant(X):
    """statsmodels.base.data ``_handle_constant``: explicit constant column
    (zero peak-to-peak, non-zero value) or an implicit one (adding a column of
    ones does not raise the rank)."""
    if X.shape[1] == 0:
        return 0
    ptp = X.max(axis=0) - X.min(axis=0)
    if np.any((ptp == 0) & (X[0] != 0)):
        return 1
    aug = np.column_stack((np.ones(X.shape[0]), X))
    return int(np.linalg.matrix_rank(aug) == np.linalg.matrix_rank(X))


def ols_pinv(y, X):
    """``sm.OLS(y, X).fit()`` restated (RegressionModel.fit, method="pinv" and
    the RegressionResults properties read by ``linear.py:50-88``)."""
    y = np.asarray(y, dtype=float).ravel()
    X = np.asarray(X, dtype=float)
    nobs = float(X.shape[0])
    pinv, sing = _pinv_extended(X)
    params = pinv @ y
    ncov = pinv @ pinv.T
    rank = np.linalg.matrix_rank(np.diag(sing))
    k_const = _has_constant(X)
    df_resid = nobs - rank
    resid = y - X @ params
    ssr = float(resid @ resid)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = ssr / df_resid
        bse = np.sqrt(np.diag(ncov * scale))
        tvalues = params / bse
        pvalues = stats.t.sf(np.abs(tvalues), df_resid) * 2
        q = stats.t.ppf(1 - 0.05 / 2, df_resid)
        lower = params - q * bse
        upper = params + q * bse
        eigenvals = np.sort(sing ** 2)[::-1]
        cond = np.sqrt(eigenvals[0] / eigenvals[-1])
        if k_const:
            tss = float(np.sum((y - y.mean()) ** 2))
        else:
            tss = float(y @ y)
        rsq = 1 - ssr / tss
        rsq_adj = 1 - np.divide(nobs - k_const, df_resid) * (1 - rsq)
        nobs2 = nobs / 2.0
        llf = -nobs2 * np.log(2 * np.pi) - nobs2 * np.log(ssr / nobs) - nobs2
    return dict(params=params, bse=bse, tvalues=tvalues, pvalues=pvalues,
                lower=lower, upper=upper, eigenvals=eigenvals,
                condition_number=cond, rsquared_adj=rsq_adj, llf=llf,
                nobs=nobs, df_resid=df_resid, ssr=ssr, scale=scale)


# --------------------------------------------------------------------------
# GLM Binomial / logit, IRLS
# --------------------------------------------------------------------------
def _logit_deriv(mu):
    p = np.clip(mu, FLOAT_EPS, 1 - FLOAT_EPS)
    return 1.0 / (p * (1 - p))


def _binom_variance(mu):
    p = np.clip(mu, FLOAT_EPS, 1 - FLOAT_EPS)
    return p * (1 - p)


def _binom_deviance(y, mu):
    endog_mu = np.clip(y / (mu + 1e-20), FLOAT_EPS, np.inf)
    n_endog_mu = np.clip((1.0 - y) / (1.0 - mu + 1e-20), FLOAT_EPS, np.inf)
    return float(np.sum(2 * (y * np.log(endog_mu) +
                             (1 - y) * np.log(n_endog_mu))))


def glm_binomial_irls(y, X, maxiter=100, tol=1e-8, return_history=False):
    """``sm.GLM(y, X, family=Binomial()).fit()`` restated (GLM._fit_irls):
    start mu=(y+.5)/2, WLS by least squares each step, perfect-separation
    check ``allclose(mu - y, 0)``, stop on abs(delta deviance) <= 1e-8, then
    params/cov from the LAST step's working weights."""
    y = np.asarray(y, dtype=float).ravel()
    X = np.asarray(X, dtype=float)
    mu = (y + 0.5) / 2.0
    lin_pred = np.log(mu / (1 - mu))
    dev = [np.inf, _binom_deviance(y, mu)]
    if np.isnan(dev[1]):
        raise ValueError("The first guess on the deviance function returned "
                         "a nan.")
    niter = 0
    for it in range(maxiter):
        d = _logit_deriv(mu)
        w = 1.0 / (d ** 2 * _binom_variance(mu))
        z = lin_pred + d * (y - mu)
        sw = np.sqrt(w)
        params = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=-1)[0]
        lin_pred = X @ params
        mu = 1.0 / (1.0 + np.exp(-lin_pred))
        dev.append(_binom_deviance(y, mu))
        niter = it + 1
        if np.allclose(mu - y, 0):
            raise OraclePerfectSeparation()
        if np.allclose(dev[it + 1], dev[it + 2], atol=tol, rtol=0.0):
            break
    # final WLS (same working response / weights), method="pinv"
    pinv, _ = _pinv_extended(X * sw[:, None])
    params = pinv @ (z * sw)
    cov = pinv @ pinv.T
    bse = np.sqrt(np.diag(cov))
    zval = params / bse
    pvalues = stats.norm.sf(np.abs(zval)) * 2
    q = stats.norm.ppf(1 - 0.05 / 2)
    llf = float(np.sum(y * np.log(mu / (1 - mu) + 1e-20) +
                       np.log(1 - mu + 1e-20)))
    out = dict(params=params, bse=bse, tvalues=zval, pvalues=pvalues,
               lower=params - q * bse, upper=params + q * bse, llf=llf,
               nobs=float(X.shape[0]), niter=niter)
    if return_history:
        out["deviance"] = dev
    return out


# --------------------------------------------------------------------------
# PHReg, Efron ties, Newton
# --------------------------------------------------------------------------
def _efron_pieces(time, status, X, params, want, strata=None):
    """Log partial likelihood (Efron), gradient and Hessian at ``params``.
    With ``strata`` (PHReg(strata=...), genetest survival.py:54-56,77-82;
    statsmodels PHSurvivalTime splits the rows by stratum label and sums the
    per-stratum partial likelihoods, strata without events contribute nothing)
    the pieces of every stratum are added up."""
    if strata is not None:
        strata = np.asarray(strata)
        p = X.shape[1]
        like, grad, hess = 0.0, np.zeros(p), np.zeros((p, p))
        for lab in np.unique(strata):
            sel = strata == lab
            if status[sel].sum() == 0:
                continue
            a, b, c = _efron_pieces_one(time[sel], status[sel], X[sel], params, want)
            like += a
            grad += b
            hess += c
        return like, grad, hess
    return _efron_pieces_one(time, status, X, params, want)


def _efron_pieces_one(time, status, X, params, want):
    """One stratum (or the whole, unstratified, sample).

    statsmodels.duration.hazard_regression.PHReg.efron_{loglike,gradient,
    hessian}: samples sorted by time, unique failure times walked backwards
    while the risk-set sums S0, S1, S2 grow."""
    order = np.argsort(time, kind="stable")
    t = time[order]
    d = status[order]
    Z = X[order]
    n, p = Z.shape
    linpred = Z @ params
    linpred = linpred - linpred.max()
    e = np.exp(linpred)
    like = 0.0
    grad = np.zeros(p)
    hess = np.zeros((p, p))
    ufail = np.unique(t[d == 1])
    s0 = 0.0
    s1 = np.zeros(p)
    s2 = np.zeros((p, p))
    hi = n
    for ft in ufail[::-1]:
        lo = np.searchsorted(t, ft, side="left")
        if lo < hi:
            ee = e[lo:hi]
            zz = Z[lo:hi]
            s0 += ee.sum()
            if want > 0:
                s1 += ee @ zz
            if want > 1:
                s2 += (zz * ee[:, None]).T @ zz
            hi = lo
        tied = np.flatnonzero((t == ft) & (d == 1))
        m = len(tied)
        ef = e[tied]
        zf = Z[tied]
        f0 = ef.sum()
        like += linpred[tied].sum()
        if want > 0:
            f1 = ef @ zf
            grad += zf.sum(axis=0)
        if want > 1:
            f2 = (zf * ef[:, None]).T @ zf
        for j in range(m):
            c = j / float(m)
            den = s0 - c * f0
            like -= np.log(den)
            if want > 0:
                num1 = s1 - c * f1
                grad -= num1 / den
            if want > 1:
                num2 = s2 - c * f2
                hess -= num2 / den - np.outer(num1, num1) / den ** 2
    return like, grad, hess


def phreg_efron_newton(time, status, X, maxiter=100, tol=1e-8, ridge=1e-10,
                       strata=None):
    """``sm.PHReg(..., ties="efron").fit()`` restated: LikelihoodModel.fit
    with method="newton" works on loglike/nobs with a +1e-10 ridge on the
    (un-negated) Hessian diagonal, from beta=0, until max abs(step) <= 1e-8;
    cov = inv(-H(beta_hat)) without the ridge."""
    time = np.asarray(time, dtype=float).ravel()
    status = np.asarray(status, dtype=float).ravel()
    X = np.asarray(X, dtype=float)
    nobs = X.shape[0]
    p = X.shape[1]
    new = np.zeros(p)
    old = np.full(p, np.inf)
    it = 0
    while it < maxiter and np.any(np.abs(new - old) > tol):
        _, g, H = _efron_pieces(time, status, X, new, 2, strata)
        H = H / nobs
        H[np.diag_indices(p)] += ridge
        old = new
        try:
            new = old - np.linalg.solve(H, g / nobs)
        except np.linalg.LinAlgError as exc:
            raise OracleLinAlgError(str(exc))
        it += 1
    if it == maxiter:
        raise OracleConvergence("Maximum Likelihood optimization failed to "
                                "converge.")
    llf, _, H = _efron_pieces(time, status, X, new, 2, strata)
    try:
        cov = np.linalg.inv(-H)
    except np.linalg.LinAlgError as exc:
        raise OracleLinAlgError(str(exc))
    with np.errstate(invalid="ignore"):
        bse = np.sqrt(np.diag(cov))
        z = new / bse
    pvalues = stats.norm.sf(np.abs(z)) * 2
    q = stats.norm.ppf(1 - 0.05 / 2)
    return dict(params=new, bse=bse, tvalues=z, pvalues=pvalues,
                lower=new - q * bse, upper=new + q * bse, llf=float(llf),
                nobs=nobs, niter=it)
