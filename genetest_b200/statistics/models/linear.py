"""Ordinary least squares (``genetest/statistics/models/linear.py``)."""

import numpy as np

from ..core import BatchedGWAS, StatsError, StatsModels
from ._single import fit_single

__all__ = ["StatsLinear"]

PARAM_FIELDS = ("coef", "std_err", "lower_ci", "upper_ci", "t_value",
                "p_value")


class StatsLinear(StatsModels, BatchedGWAS):
    gwas_model = "linear"

    def __init__(self, condition_value_t=1000, eigenvals_t=1e-10):
        """
        Args:
            condition_value_t: reject a fit whose design has a condition
                number above this (strong multicollinearity).
            eigenvals_t: reject a fit whose smallest eigenvalue of X'X is
                below this (singular design).
        """
        self._condition_value_t = condition_value_t
        self._eigenvals_t = eigenvals_t

    def gwas_options(self):
        return {"condition_value_t": self._condition_value_t,
                "eigenvals_t": self._eigenvals_t}

    def fit(self, y, X):
        """Fit ``y ~ X`` on the GPU engine; same dict as the reference
        (``linear.py:64-89``)."""
        res, order = fit_single("linear", [y.iloc[:, 0].values.astype(float)],
                               X, self.gwas_options())
        out = {"MODEL": {"r_squared_adj": float(res.stat[0]),
                         "log_likelihood": float(res.llf[0]),
                         "nobs": float(res.nobs[0])}}
        for j, col in enumerate(X.columns):
            vals = res.param(order[j])[:, 0]
            if col == "SNPs" and np.isnan(vals[5]):
                raise StatsError("Inference did not converge.")
            out[col] = {f: float(v) for f, v in zip(PARAM_FIELDS, vals)}
        return out
