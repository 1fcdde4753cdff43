"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product.

Restates genetest's per-SNP association loop on plain numpy arrays:

* ``get_maf`` / ``get_sex_maf``  <- ``genetest/statistics/descriptive.py:19-48`` / ``:51-93``
* ``stats_linear``               <- ``genetest/statistics/models/linear.py:46-104``
* ``stats_logistic``             <- ``genetest/statistics/models/logistic.py:29-74``
* ``stats_coxph``                <- ``genetest/statistics/models/survival.py:36-122``
* ``gwas_one`` / ``gwas``        <- ``genetest/analysis.py:68-205`` (``_gwas_worker``)

The estimators themselves are in ``oracle/sm_port.py`` (statsmodels restated).
Parity status: PINNED by ``tests/test_oracle_golden.py`` (reference goldens).
"""

import numpy as np

from . import sm_port


class OracleStatsError(Exception):
    """Mirror of ``genetest.statistics.core.StatsError`` (``core.py:54-60``)."""
    def __init__(self, msg):
        super().__init__(str(msg))
        self.message = str(msg)

    def __str__(self):
        return self.message


# descriptive.py:19-48
def get_maf(genotypes, minor, major):
    with np.errstate(invalid="ignore"), _quiet():
        maf = np.nanmean(genotypes) / 2
    if maf > 0.5:
        return 1 - maf, major, minor, True
    return maf, minor, major, False


# descriptive.py:51-93
def get_sex_maf(genotypes, sexes, minor, major):
    genotypes = np.asarray(genotypes, dtype=float)
    sexes = np.asarray(sexes, dtype=float)
    ok = ~(np.isnan(genotypes) | np.isnan(sexes))
    if (~ok).all():
        return np.nan, minor, major, False
    g = genotypes[ok]
    s = sexes[ok]
    nb_males = np.sum(s)
    nb_females = s.shape[0] - nb_males
    maf = (-0.5 * np.dot(g, s) + np.sum(g)) / (nb_males + 2 * nb_females)
    if maf > 0.5:
        return 1 - maf, major, minor, True
    return maf, minor, major, False


class _quiet(object):
    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        self._cm.__exit__(*a)


# linear.py:46-104
def stats_linear(y, X, names, condition_value_t=1000, eigenvals_t=1e-10):
    fit = sm_port.ols_pinv(y, X)
    if fit["condition_number"] > condition_value_t:
        raise OracleStatsError("condition number is large, {}".format(
            fit["condition_number"]))
    if fit["eigenvals"].min() < eigenvals_t:
        raise OracleStatsError("smallest eigenvalue is small, {}".format(
            fit["eigenvals"].min()))
    out = {"MODEL": {"r_squared_adj": fit["rsquared_adj"],
                     "log_likelihood": fit["llf"], "nobs": fit["nobs"]}}
    for i, name in enumerate(names):
        if name == "SNPs" and np.isnan(fit["pvalues"][i]):
            raise OracleStatsError("Inference did not converge.")
        out[name] = {"coef": fit["params"][i], "std_err": fit["bse"][i],
                     "lower_ci": fit["lower"][i], "upper_ci": fit["upper"][i],
                     "t_value": fit["tvalues"][i],
                     "p_value": fit["pvalues"][i]}
    return out


# logistic.py:29-74
def stats_logistic(y, X, names):
    y = np.asarray(y, dtype=float).ravel()
    try:
        fit = sm_port.glm_binomial_irls(y, X)
    except sm_port.OraclePerfectSeparation as e:
        raise OracleStatsError(str(e))
    out = {"MODEL": {"log_likelihood": fit["llf"], "nobs": fit["nobs"],
                     "nevents": np.sum(y)}}
    for i, name in enumerate(names):
        if name == "SNPs" and np.isnan(fit["pvalues"][i]):
            raise OracleStatsError("Inference did not converge.")
        out[name] = {"coef": fit["params"][i], "std_err": fit["bse"][i],
                     "lower_ci": fit["lower"][i], "upper_ci": fit["upper"][i],
                     "t_value": fit["tvalues"][i],
                     "p_value": fit["pvalues"][i]}
    return out


# survival.py:36-122
def stats_coxph(tte, event, X, names, strata=None):
    X = np.asarray(X, dtype=float)
    keep = [i for i, nm in enumerate(names) if nm != "intercept"]
    names = [names[i] for i in keep]
    X = X[:, keep]
    try:
        fit = sm_port.phreg_efron_newton(tte, event, X, strata=strata)
    except sm_port.OracleLinAlgError as e:
        raise OracleStatsError(str(e))
    except sm_port.OracleConvergence as e:
        raise OracleStatsError(str(e))
    out = {"MODEL": {"log_likelihood": fit["llf"], "nobs": X.shape[0],
                     "nevents": np.sum(event)}}
    for i, name in enumerate(names):
        if name == "SNPs" and np.isnan(fit["pvalues"][i]):
            raise OracleStatsError("Inference did not converge.")
        out[name] = {"coef": fit["params"][i], "std_err": fit["bse"][i],
                     "hr": np.exp(fit["params"][i]),
                     "hr_lower_ci": np.exp(fit["lower"][i]),
                     "hr_upper_ci": np.exp(fit["upper"][i]),
                     "z_value": fit["tvalues"][i],
                     "p_value": fit["pvalues"][i]}
    return out


def gwas_one(model, y, C, cov_names, g, maf_t=None, interaction=None,
             sexes=None, model_kwargs=None, coded="A1", reference="A2"):
    """One pass of ``_gwas_worker``'s loop body (``analysis.py:108-197``).

    ``y``: (n, 1) outcome, or (n, 2) ``[tte, event]`` / (n, 3) ``[tte, event,
    strata]`` for coxph.
    ``C``: (n, k) covariates (complete: ``execute`` already ran ``dropna``),
    ``cov_names`` its column ids (``"intercept"`` last when present).
    ``g``: (n,) dosages aligned to the rows of ``C``; NaN = missing genotype
    or sample absent from the genotype file.
    ``interaction``: ordered ``{out_col: [covariate ids]}``
    (``modelspec.gwas_interaction``).

    Returns ``("ok", results_dict)`` or ``("failed", reason)``."""
    model_kwargs = model_kwargs or {}
    g = np.asarray(g, dtype=float)
    ok = ~np.isnan(g)                                   # analysis.py:119
    if ok.sum() == 0:                                    # :121-124
        return "failed", "All genotypes are missing"
    if sexes is None:                                    # :135-148
        maf, minor, major, flip = get_maf(g[ok], coded, reference)
    else:
        maf, minor, major, flip = get_sex_maf(g[ok], np.asarray(sexes)[ok],
                                              coded, reference)
    if maf_t is not None and maf < maf_t:                # :151-153
        return "failed", "MAF: {} < {}".format(maf, maf_t)
    if flip:                                             # :156-157
        g = 2 - g
    cols = [C, g[:, None]]
    names = list(cov_names) + ["SNPs"]
    if interaction:                                      # :159-166
        for key, mult in interaction.items():
            v = g.copy()
            for c in mult:
                v = v * C[:, cov_names.index(c)]
            cols.append(v[:, None])
            names.append(key)
    X = np.hstack(cols)[ok]
    yy = np.asarray(y, dtype=float)[ok]
    try:                                                 # :171-181
        if model == "linear":
            res = stats_linear(yy[:, 0], X, names, **model_kwargs)
        elif model == "logistic":
            res = stats_logistic(yy[:, 0], X, names)
        elif model == "coxph":
            # a third outcome column is y.strata (survival.py:54-56)
            res = stats_coxph(yy[:, 0], yy[:, 1], X, names,
                              strata=yy[:, 2] if yy.shape[1] > 2 else None)
        else:
            raise ValueError(model)
    except OracleStatsError as e:
        return "failed", str(e)
    res["SNPs"].update({"major": major, "minor": minor, "maf": maf})  # :191-195
    res["_flip"] = flip
    return "ok", res


def gwas(model, y, C, cov_names, genotypes, **kw):
    """Run ``gwas_one`` over an iterable of ``(name, dosages)``; returns
    ``(results: {name: dict}, failed: [(name, reason)])``."""
    results, failed = {}, []
    for name, g in genotypes:
        status, payload = gwas_one(model, y, C, cov_names, g, **kw)
        if status == "ok":
            results[name] = payload
        else:
            failed.append((name, payload))
    return results, failed
