"""Pin the CPU oracle (``oracle/``) to the reference's golden vectors.

Every SAS / R number asserted by the reference's test_linear.py,
test_logistic.py and test_coxph.py (``tests/golden/expected.json``) and the
coefficient vectors of test_gwaswriter.py must be reproduced by the numpy
restatement before it may be used as the checker for the CUDA path.
"""

import numpy as np
import pytest

from oracle import gwas as og
from oracle import sm_port

from golden_support import (CASES, EXPECTED, MARKERS, OUTCOME, NSNPS,
                            case_ids, check_record, covariates, load_table)


def _run_gwas_case(case):
    df = load_table(case["data"])
    model = case["model"]
    C, names = covariates(df)
    y = df[OUTCOME[model]].values.astype(float)
    inter = None
    if case["inter"] == "var1":
        inter = {"INTER": ["var1"]}
    elif case["inter"] == "gender":
        inter = {"INTER:level.2": ["gender:level.2"]}
    results, failed = {}, []
    for k in range(NSNPS[model]):
        snp = "snp{}".format(k + 1)
        chrom, pos, a1, a2 = MARKERS[snp]
        status, res = og.gwas_one(
            model, y, C, names, df[snp].values.astype(float),
            interaction=inter, model_kwargs=case["kwargs"],
            coded=a1, reference=a2)
        if status == "failed":
            failed.append([snp, res])
            continue
        res["SNPs"].update(name=snp, chrom=chrom, pos=pos)
        results[snp] = res
    return results, failed


def _run_single_case(case):
    df = load_table(case["data"])
    model = case["model"]
    snp = case["snp"]
    chrom, pos, a1, a2 = MARKERS[snp]
    cols = OUTCOME[model] + ["age", "var1", "gender", snp]
    extra = case.get("inter_snp", snp)
    if extra != snp:
        cols.append(extra)
    df = df[cols].copy()
    # modelspec/core.py:438-446: MAF / flip on the joined samples
    for s in {snp, extra}:
        m = og.get_maf(df[s].values.astype(float), *MARKERS[s][2:])
        if m[3]:
            df[s] = 2 - df[s]
        if s == snp:
            maf, minor, major, _ = m
    C, cnames = covariates(df)
    g = df[snp].values.astype(float)
    X = [g, C[:, 0], C[:, 1], C[:, 2]]
    names = [snp, "age", "var1", "gender:level.2"]
    if case["inter"] == "var1":
        X.append(df["var1"].values * df[extra].values.astype(float))
        names.append("INTER")
    elif case["inter"] == "gender":
        X.append(C[:, 2] * g)
        names.append("INTER:level.2")
    X.append(C[:, 3])
    names.append("intercept")
    X = np.column_stack(X)
    y = df[OUTCOME[model]].values.astype(float)
    ok = ~(np.isnan(X).any(axis=1) | np.isnan(y).any(axis=1))   # dropna
    X, y = X[ok], y[ok]
    if model == "linear":
        res = og.stats_linear(y[:, 0], X, names, **case["kwargs"])
    elif model == "logistic":
        res = og.stats_logistic(y[:, 0], X, names)
    else:
        res = og.stats_coxph(y[:, 0], y[:, 1], X, names)
    res[snp].update(name=snp, chrom=chrom, pos=pos, minor=minor, major=major,
                    maf=maf)
    return res


@pytest.mark.parametrize("model", ["linear", "logistic", "coxph"])
def test_reference_goldens(model):
    for name in case_ids(model):
        case = CASES[model][name]
        rec = EXPECTED[model][name]
        if case["mode"] == "gwas":
            results, failed = _run_gwas_case(case)
            assert len(results) == rec["n_results"], name
            assert failed == rec["failed"], name
            problems = check_record(
                rec, lambda snp, ent, fld: results[snp][ent][fld])
        elif "raises" in rec:
            with pytest.raises(og.OracleStatsError) as err:
                _run_single_case(case)
            assert str(err.value) == rec["raises"], name
            continue
        else:
            res = _run_single_case(case)
            problems = check_record(
                rec, lambda snp, ent, fld: res[ent][fld])
        assert not problems, (name, problems)


def test_iteration_counts_match_survey():
    """logistic snp1 converges in 5 IRLS steps, Cox
    snp1 in 6 Newton steps (flip=True, nobs 31)."""
    df = load_table("logistic")
    C, names = covariates(df)
    X = np.column_stack((C, df.snp1.values))
    fit = sm_port.glm_binomial_irls(df.pheno2.values, X)
    assert fit["niter"] == 5
    np.testing.assert_allclose(fit["params"][-1], -2.2419876, rtol=1e-6)
    np.testing.assert_allclose(fit["bse"][-1], 0.5975958, rtol=1e-6)

    df = load_table("coxph")
    C, names = covariates(df)
    status, res = og.gwas_one("coxph", df[["tte", "event"]].values, C, names,
                              df.snp1.values, coded="T", reference="C")
    assert status == "ok" and res["_flip"] is True
    assert res["MODEL"]["nobs"] == 31
    np.testing.assert_allclose(res["SNPs"]["coef"], 3.4982415732348975,
                               rtol=1e-9)
    np.testing.assert_allclose(res["SNPs"]["std_err"], 1.057449647346056,
                               rtol=1e-8)


def test_gwaswriter_coefficients():
    """test_gwaswriter.py:131-134, 211-214, 258-267 (plain, SNPs*var1,
    SNPs*factor(var3)) and the var4 subgroups (:163-177)."""
    df = load_table("factors")

    def design(d):
        cols, names = [d.var1.values, d.var2.values], ["var1", "var2"]
        for var in ("var3", "var4", "var5"):
            levels = sorted(d[var].unique())
            for lv in levels[1:]:
                col = (d[var].values == lv).astype(float)
                if col.sum() == 0:
                    continue
                cols.append(col)
                names.append("{}:level.{}".format(var, lv))
        cols.append(np.ones(d.shape[0]))
        names.append("intercept")
        return np.column_stack(cols).astype(float), names

    C, names = design(df)
    y = df[["pheno"]].values.astype(float)

    def coefs(inter, key, d=df, CC=C, nn=names, yy=y):
        out = []
        for s in ("snp1", "snp2", "snp3"):
            st, res = og.gwas_one("linear", yy, CC, nn,
                                  d[s].values.astype(float),
                                  interaction=inter)
            assert st == "ok", res
            out.append(res[key]["coef"])
        return np.array(out)

    np.testing.assert_array_almost_equal(
        coefs(None, "SNPs"),
        [6.48616317675085, 4.29701960321627, -1.44891512389655])
    np.testing.assert_array_almost_equal(
        coefs({"I": ["var1"]}, "I"),
        [0.4073975765317, -0.04253919820585, -0.05814970176832])
    inter = {"I:level.x1": ["var3:level.x1"], "I:level.x2": ["var3:level.x2"]}
    np.testing.assert_array_almost_equal(
        coefs(inter, "I:level.x1"),
        [6.15574045825434, 2.2178331615269, -1.94057337856961])
    np.testing.assert_array_almost_equal(
        coefs(inter, "I:level.x2"),
        [21.47033225862596, 2.7136560668229, -0.05622018358578])

    # subgroup analysis pheno | var4 ~ ... (stratified: var4 columns dropped)
    expected = {"y0": [-0.6798617801022, 5.9538885434287, -15.1758958532932],
                "y1": [12.69000471889982, 4.2496912087917, 7.5291802846083]}
    for level, exp in expected.items():
        sub = df[df.var4 == level]
        Cs, ns = design(sub.drop(columns=["var4"]).assign(var4="z"))
        keep = [i for i, n in enumerate(ns) if not n.startswith("var4")]
        Cs, ns = Cs[:, keep], [ns[i] for i in keep]
        got = coefs(None, "SNPs", d=sub, CC=Cs, nn=ns,
                    yy=sub[["pheno"]].values.astype(float))
        np.testing.assert_array_almost_equal(got, exp)


def test_maf_vectors():
    """test_descriptive_stats.py:25-116."""
    g = np.array([0, 0, 0, 1, 0, 0, 1, 2, 0, 0], dtype=float)
    assert og.get_maf(g, "B", "A") == (4 / 20, "B", "A", False)
    g = np.array([2, 2, 2, 1, 2, 2, 1, 0, 2, 2], dtype=float)
    maf, minor, major, flip = og.get_maf(g, "B", "A")
    assert (minor, major, flip) == ("A", "B", True)
    assert abs(maf - 4 / 20) < 1e-15
    g = np.array([0, np.nan, 0, 1, 0, np.nan, 1, 2, 0, 0], dtype=float)
    assert og.get_maf(g, "B", "A")[0] == 4 / 16
    maf, minor, major, flip = og.get_maf(np.full(5, np.nan), "B", "A")
    assert np.isnan(maf) and (minor, major, flip) == ("B", "A", False)
    # sex-aware (males are hemizygous)
    g = np.array([0, 0, 0, 2, 0, 0, 1, 2, 0, 0], dtype=float)
    sex = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0], dtype=float)
    maf, minor, major, flip = og.get_sex_maf(g, sex, "B", "A")
    assert abs(maf - 3 / 15) < 1e-15 and not flip
    maf, _, _, _ = og.get_sex_maf(np.full(4, np.nan), np.ones(4), "B", "A")
    assert np.isnan(maf)


def test_stratified_cox_against_a_brute_force_partial_likelihood():
    """PHReg(strata=...) (survival.py:54-56, 77-82).  The reference's tests hold no
    stratified case, so the oracle's extension is checked against a direct
    restatement of the stratified Efron partial likelihood (explicit loops over
    strata, event times and tied events; numerical derivatives) and against
    the identities it has to satisfy."""
    from oracle import sm_port
    rng = np.random.default_rng(3)
    n, p = 90, 3
    X = rng.normal(size=(n, p))
    tte = np.ceil(rng.exponential(1.0, n) * 6) / 6         # ties
    event = (rng.random(n) < 0.7).astype(float)
    strata = rng.integers(0, 4, n).astype(float)
    event[strata == 3] = 0.0                                # a stratum without events

    def brute(beta):
        ll = 0.0
        for lab in np.unique(strata):
            idx = np.flatnonzero(strata == lab)
            for t in np.unique(tte[idx][event[idx] == 1]):
                risk = idx[tte[idx] >= t]
                tied = idx[(tte[idx] == t) & (event[idx] == 1)]
                m = len(tied)
                s0 = np.exp(X[risk] @ beta).sum()
                f0 = np.exp(X[tied] @ beta).sum()
                ll += (X[tied] @ beta).sum()
                for j in range(m):
                    ll -= np.log(s0 - j / m * f0)
        return ll

    beta = rng.normal(size=p) * 0.3
    like, grad, hess = sm_port._efron_pieces(tte, event, X, beta, 2, strata)
    assert abs(like - brute(beta)) < 1e-10 * abs(like)
    h = 1e-5
    for a in range(p):
        e = np.zeros(p)
        e[a] = h
        num = (brute(beta + e) - brute(beta - e)) / (2 * h)
        assert abs(grad[a] - num) < 1e-6 * max(1.0, abs(num))
        for b in range(p):
            e2 = np.zeros(p)
            e2[b] = h
            num2 = (brute(beta + e + e2) - brute(beta + e - e2) -
                    brute(beta - e + e2) + brute(beta - e - e2)) / (4 * h * h)
            assert abs(hess[a, b] - num2) < 2e-4 * max(1.0, abs(num2))
    # one stratum = no strata; the fit maximises the stratified likelihood
    one = sm_port.phreg_efron_newton(tte, event, X, strata=np.zeros(n))
    none = sm_port.phreg_efron_newton(tte, event, X)
    np.testing.assert_allclose(one["params"], none["params"], rtol=1e-12)
    fit = sm_port.phreg_efron_newton(tte, event, X, strata=strata)
    assert abs(fit["llf"] - brute(fit["params"])) < 1e-10 * abs(fit["llf"])
    for a in range(p):
        e = np.zeros(p)
        e[a] = 1e-4
        assert brute(fit["params"]) >= brute(fit["params"] + e)
        assert brute(fit["params"]) >= brute(fit["params"] - e)
    assert np.max(np.abs(fit["params"] - none["params"])) > 1e-3
