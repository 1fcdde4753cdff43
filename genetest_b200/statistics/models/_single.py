"""Single fit ``Stats*.fit(y, X)`` on the engine.

The reference refits statsmodels for one design matrix
(``genetest/analysis.py:562-590`` ``_execute_simple``).  Here the same call
is a batch of ONE on the GPU engine in raw mode: the last column of ``X``
plays the marker's role (no MAF, no flip), the others are the covariates.
No CPU path exists.
"""

import numpy as np

from ...engine import get_engine
from ..core import StatsError

FAILURES = {
    1: lambda r, i: "All genotypes are missing",
    3: lambda r, i: "condition number is large, {}".format(r.aux[i]),
    4: lambda r, i: "smallest eigenvalue is small, {}".format(r.aux[i]),
    5: lambda r, i: "Inference did not converge.",
    6: lambda r, i: "Perfect separation detected, results not available",
    7: lambda r, i: "Singular matrix",
    8: lambda r, i: "Maximum Likelihood optimization failed to converge.",
}


def failure_reason(res, i, maf_t=None):
    """The reference's failure string for marker ``i`` of a result block."""
    status = int(res.status[i])
    if status == 2:
        return "MAF: {} < {}".format(res.maf[i], maf_t)
    return FAILURES[status](res, i)


def fit_single(model, y_cols, X, options, device=0):
    """One-marker raw-mode batch.  Returns ``(results, order)`` where
    ``order[j]`` is the engine parameter index of column ``j`` of ``X``.

    The column handed to the engine as "the marker" is the last
    non-constant one, so that an intercept stays among the covariates (the
    engine detects the constant there, as statsmodels' ``k_constant`` does).
    """
    Xv = np.ascontiguousarray(X.values, dtype=float)
    p = Xv.shape[1]
    marker = p - 1
    for j in range(p - 1, -1, -1):
        if Xv.shape[0] and np.ptp(Xv[:, j]) != 0:
            marker = j
            break
    rest = [j for j in range(p) if j != marker]
    order = [0] * p
    for pos, j in enumerate(rest):
        order[j] = pos
    order[marker] = p - 1
    eng = get_engine(device)
    kwargs = dict(options)
    if model == "coxph":
        eng.set_design(model, y_cols[0], Xv[:, rest], event=y_cols[1],
                       raw_mode=True, **kwargs)
    else:
        eng.set_design(model, y_cols[0], Xv[:, rest], raw_mode=True, **kwargs)
    res = eng.run_dosage_host(np.ascontiguousarray(Xv[:, marker][None, :]))
    if res.status[0] != 0:
        raise StatsError(failure_reason(res, 0))
    return res, order
