"""Rounding noise of the float64 oracle (numpy SVD pseudo-inverse, as statsmodels'
OLS) at the config-2 shape: its coefficients against the normal equations
accumulated in 80-bit long double and solved with 40 digits (mpmath).
Result (2026-10-20): up to 1.3e-11 standard errors -- the floor behind the
0.01-standard-error denominator guard of tests/gpu_support.compare(strict=True)."""
import sys
import numpy as np
import mpmath
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from gpu_support import synth_design
from oracle import gwas as og

rng = np.random.default_rng(20261020)
n = 100000
C, names = synth_design(rng, n, k_extra=8)
y = 0.1 * C[:, 0] + rng.normal(size=n) + 10.0
r2 = np.random.default_rng(1)
mpmath.mp.dps = 40
worst = 0.0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    maf = r2.uniform(0.05, 0.5)
    g = r2.binomial(2, maf, n).astype(float)
    g[r2.random(n) < 0.01] = np.nan
    st, res = og.gwas_one("linear", y[:, None], C, names, g)
    ok = ~np.isnan(g)
    X = np.column_stack((C, g))[ok].astype(np.longdouble)
    G, b = X.T @ X, X.T @ y[ok].astype(np.longdouble)

    def mp(v):
        hi = float(v)
        return mpmath.mpf(hi) + mpmath.mpf(float(v - np.longdouble(hi)))
    beta = mpmath.lu_solve(mpmath.matrix([[mp(v) for v in row] for row in G]),
                           mpmath.matrix([mp(v) for v in b]))
    for i, nm in enumerate(names + ["SNPs"]):
        err = abs(res[nm]["coef"] - float(beta[i])) / res[nm]["std_err"]
        worst = max(worst, err)
print("worst |oracle - exact| = %.2e standard errors" % worst)
