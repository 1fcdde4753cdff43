"""Helpers for the GPU parity tests: run the oracle over a batch of markers
and compare it field by field with the engine's result blocks."""

import numpy as np

from oracle import gwas as og

STATUS_BY_REASON = [
    ("All genotypes are missing", 1), ("MAF:", 2),
    ("condition number is large", 3), ("smallest eigenvalue is small", 4),
    ("Inference did not converge.", 5), ("Perfect separation", 6),
    ("Singular matrix", 7), ("Maximum Likelihood", 8),
]

LIN_FIELDS = ("coef", "std_err", "lower_ci", "upper_ci", "t_value", "p_value")
COX_FIELDS = ("coef", "std_err", "coef_lower", "coef_upper", "z_value",
              "p_value")


def reason_to_status(reason):
    for prefix, code in STATUS_BY_REASON:
        if reason.startswith(prefix):
            return code
    raise AssertionError("unknown failure reason: " + reason)


def synth_design(rng, n, k_extra=0, standardise=True):
    """age, sex, k_extra N(0,1) covariates, intercept."""
    age = rng.normal(50, 10, n)
    if standardise:
        age = (age - age.mean()) / age.std()
    sex = rng.integers(0, 2, n).astype(float)
    cols = [age, sex] + [rng.normal(size=n) for _ in range(k_extra)]
    names = ["age", "sex"] + ["c{}".format(i) for i in range(k_extra)]
    cols.append(np.ones(n))
    names.append("intercept")
    return np.column_stack(cols), names


def synth_genotypes(rng, n_snps, n, missing=0.0, maf_lo=0.05, maf_hi=0.5):
    maf = rng.uniform(maf_lo, maf_hi, n_snps)
    g = rng.binomial(2, maf[:, None], size=(n_snps, n)).astype(float)
    if missing > 0:
        g[rng.random((n_snps, n)) < missing] = np.nan
    return g


def run_oracle(model, y, C, names, G, interaction=None, maf_t=None,
               model_kwargs=None, sexes=None):
    out = []
    for j in range(G.shape[0]):
        out.append(og.gwas_one(model, y, C, names, G[j], maf_t=maf_t,
                               interaction=interaction, sexes=sexes,
                               model_kwargs=model_kwargs))
    return out


# ---------------------------------------------------------------------------
# oracle over many markers: one process per core, rows shipped as 2-bit packs
# ---------------------------------------------------------------------------
_POOL_STATE = {}


def _pool_init(model, y, C, names, n, pos, interaction, maf_t, model_kwargs):
    import os
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _POOL_STATE["limit"] = threadpool_limits(1)
    except Exception:
        pass
    _POOL_STATE.update(model=model, y=y, C=C, names=names, n=n, pos=pos,
                       interaction=interaction, maf_t=maf_t,
                       model_kwargs=model_kwargs)


def _pool_work(packed_rows):
    from genetest_b200.genotypes import unpack_rows
    st = _POOL_STATE
    G = unpack_rows(packed_rows, st["n"])
    if st["pos"] is not None:
        G = G[:, st["pos"]]
    return [og.gwas_one(st["model"], st["y"], st["C"], st["names"], g,
                        maf_t=st["maf_t"], interaction=st["interaction"],
                        model_kwargs=st["model_kwargs"]) for g in G]


def run_oracle_packed(model, y, C, names, packed, n_file, pos=None,
                      interaction=None, maf_t=None, model_kwargs=None,
                      procs=None):
    """The oracle over the 2-bit rows ``packed`` ([m, ceil(n_file / 4)] uint8),
    spread over the host cores (2,000-marker samples take seconds
    that way).  ``pos``: genotype-file position of every design row."""
    import multiprocessing as mp
    import os
    procs = procs or max(1, min(len(os.sched_getaffinity(0)), 32))
    packed = np.ascontiguousarray(packed)
    chunks = [c for c in np.array_split(packed, min(len(packed), 4 * procs)) if len(c)]
    with mp.get_context("spawn").Pool(
            procs, _pool_init, (model, y, C, names, n_file, pos, interaction,
                                maf_t, model_kwargs)) as pool:
        parts = pool.map(_pool_work, chunks)
    return [r for part in parts for r in part]


def stratified_sample(res, rng, n_low_maf, n_most_missing, n_random):
    """Marker indices of a parity sample: the lowest MAFs, the most missing
    genotypes, and a random draw."""
    n = res.maf.shape[0]
    maf = np.where(np.isnan(res.maf), np.inf, res.maf)
    return np.unique(np.concatenate([
        np.argsort(maf, kind="stable")[:n_low_maf],
        np.argsort(res.nobs, kind="stable")[:n_most_missing],
        rng.choice(n, min(n, n_random), replace=False)]))


def compare(model, res, oracle, names, inter_keys=(), rtol=1e-8,
            check_model=True, strict=False):
    """``res``: engine Results; ``oracle``: list from run_oracle.

    ``strict``: plain relative error |got - want| / |want| on every one of coef,
    std_err, the CI bounds, t/z and p (BASELINE's "relative 1e-8 / 1e-6").  The
    only guard is the denominator of coef / CI bounds / t, which is floored at
    0.01 standard errors: the float64 oracle itself (numpy SVD pseudo-inverse,
    like statsmodels) carries up to ~1.3e-11 standard errors of rounding noise
    at n = 100,000 (measured against 40-digit arithmetic by
    scripts/oracle_noise.py), so below |t| = 0.01 a 1e-8 relative
    neighbourhood of the coefficient is finer than the checker resolves.  Without ``strict`` the errors of coef / CI are measured in units of
    max(|coef|, std_err) and the p-value error is divided by its condition
    number max(1, t^2) (the older, weaker form)."""
    keys = list(names) + ["SNPs"] + list(inter_keys)
    if model == "coxph":
        keys = [k for k in keys if k != "intercept"]
    problems = []
    worst = 0.0
    n_ok = 0
    for j, (st, payload) in enumerate(oracle):
        if st == "failed":
            want = reason_to_status(payload)
            if res.status[j] != want:
                problems.append((j, "status", int(res.status[j]), payload))
            elif want == 2:      # MAF below: the value is printed
                maf = float(payload.split()[1])
                if res.maf[j] != maf:
                    problems.append((j, "maf", res.maf[j], maf))
            elif want == 3 and payload.endswith("inf"):
                if not np.isinf(res.aux[j]):
                    problems.append((j, "aux", res.aux[j], payload))
            continue
        if res.status[j] != 0:
            problems.append((j, "status", int(res.status[j]), "ok"))
            continue
        n_ok += 1
        o = payload
        if res.nobs[j] != o["MODEL"]["nobs"]:
            problems.append((j, "nobs", int(res.nobs[j]), o["MODEL"]["nobs"]))
        if bool(res.flip[j]) != bool(o["_flip"]):
            problems.append((j, "flip", int(res.flip[j]), o["_flip"]))
        both_nan = np.isnan(res.maf[j]) and np.isnan(o["SNPs"]["maf"])
        if res.maf[j] != o["SNPs"]["maf"] and not both_nan:
            problems.append((j, "maf", res.maf[j], o["SNPs"]["maf"]))
        for i, key in enumerate(keys):
            got = res.param(i)[:, j]
            ref = o[key]
            coef, se = ref["coef"], ref["std_err"]
            scale = max(abs(coef), abs(se))
            if model == "coxph":
                want = [coef, se, np.log(ref["hr_lower_ci"]),
                        np.log(ref["hr_upper_ci"]), ref["z_value"],
                        ref["p_value"]]
            else:
                want = [ref[f] for f in LIN_FIELDS]
            for f in range(6):
                if strict:
                    den = abs(want[f])
                    if f in (0, 2, 3):
                        den = max(den, 1e-2 * abs(se))
                    elif f == 4:
                        den = max(den, 1e-2)
                    elif f == 5:
                        den = max(den, 1e-290)
                    err = abs(got[f] - want[f]) / den if den > 0 else abs(got[f] - want[f])
                elif f in (0, 2, 3):
                    err = abs(got[f] - want[f]) / scale
                elif f == 4:
                    err = abs(got[f] - want[f]) / max(1.0, abs(want[f]))
                elif f == 5:
                    # p-values: relative, conditioning ~ t^2 (and floor for
                    # underflowing tails)
                    den = max(abs(want[f]), 1e-290)
                    err = abs(got[f] - want[f]) / den / max(1.0, want[4] ** 2)
                else:
                    err = abs(got[f] - want[f]) / abs(want[f])
                if not (err <= rtol):
                    problems.append((j, key, f, got[f], want[f], err))
                worst = max(worst, err if err == err else np.inf)
        if check_model:
            m = o["MODEL"]
            pairs = [("llf", res.llf[j], m["log_likelihood"])]
            if model == "linear":
                pairs.append(("r2adj", res.stat[j], m["r_squared_adj"]))
            else:
                pairs.append(("nevents", res.stat[j], m["nevents"]))
            for nm, g, w in pairs:
                err = abs(g - w) / max(abs(w), 1e-300)
                if nm == "r2adj":
                    err = abs(g - w) / max(abs(w), 1e-3)
                if not (err <= rtol):
                    problems.append((j, nm, g, w, err))
                worst = max(worst, err)
    return problems, worst, n_ok
