"""Cox proportional hazards, Efron ties
(``genetest/statistics/models/survival.py``)."""

import logging

import numpy as np

from ..core import BatchedGWAS, StatsError, StatsModels
from ._single import fit_single

__all__ = ["StatsCoxPH"]

logger = logging.getLogger(__name__)

COX_FIELDS = ("coef", "std_err", "hr", "hr_lower_ci", "hr_upper_ci",
              "z_value", "p_value")


def cox_param_dict(vals):
    """(coef, se, lower, upper, z, p) -> the reference's per-parameter dict
    (``survival.py:112-120``)."""
    coef, se, lo, hi, z, p = (float(v) for v in vals)
    return {"coef": coef, "std_err": se, "hr": float(np.exp(coef)),
            "hr_lower_ci": float(np.exp(lo)), "hr_upper_ci": float(np.exp(hi)),
            "z_value": z, "p_value": p}


class StatsCoxPH(StatsModels, BatchedGWAS):
    gwas_model = "coxph"

    @staticmethod
    def _prepare_data(y, X):
        """``y`` needs ``tte`` and ``event``; the intercept is dropped."""
        if {"tte", "event"} - set(y.columns):
            raise ValueError(
                "missing column in y: coxph requires 'tte' and 'event'")
        extra = set(y.columns) - {"tte", "event", "strata"}
        if extra:
            logger.warning("{}: unknown column in y, will be ignored".format(
                ",".join(sorted(extra))))
        if "intercept" in X.columns:
            X = X.drop("intercept", axis=1)
        strata = y.strata if "strata" in y.columns else None     # survival.py:54-56
        return y.tte, y.event, X, strata

    def fit(self, y, X):
        tte, event, X, strata = self._prepare_data(y, X)
        res, order = fit_single(
            "coxph", [tte.values.astype(float), event.values.astype(float)],
            X, {} if strata is None else {"strata": strata.values})
        out = {"MODEL": {"log_likelihood": float(res.llf[0]),
                         "nobs": X.shape[0], "nevents": np.sum(event)}}
        for j, col in enumerate(X.columns):
            vals = res.param(order[j])[:, 0]
            if col == "SNPs" and np.isnan(vals[5]):
                raise StatsError("Inference did not converge.")
            out[col] = cox_param_dict(vals)
        return out
