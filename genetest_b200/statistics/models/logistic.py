"""Logistic regression, GLM binomial / logit by IRLS
(``genetest/statistics/models/logistic.py``)."""

import numpy as np

from ..core import BatchedGWAS, StatsError, StatsModels
from ._single import fit_single
from .linear import PARAM_FIELDS

__all__ = ["StatsLogistic"]


class StatsLogistic(StatsModels, BatchedGWAS):
    gwas_model = "logistic"

    def fit(self, y, X):
        """Same dict as the reference (``logistic.py:49-72``); the
        ``t_value`` is a z statistic."""
        yv = y.iloc[:, 0].values.astype(float)
        res, order = fit_single("logistic", [yv], X, {})
        out = {"MODEL": {"log_likelihood": float(res.llf[0]),
                         "nobs": float(res.nobs[0]),
                         "nevents": np.sum(y.y) if "y" in y.columns
                         else float(yv.sum())}}
        for j, col in enumerate(X.columns):
            vals = res.param(order[j])[:, 0]
            if col == "SNPs" and np.isnan(vals[5]):
                raise StatsError("Inference did not converge.")
            out[col] = {f: float(v) for f, v in zip(PARAM_FIELDS, vals)}
        return out
