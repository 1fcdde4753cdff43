"""Replay the reference's known-answer MixedLM tests (R / lme4 numbers asserted by
genetest/tests/test_mixedlm.py, extracted into tests/golden/expected.json) through
``analysis.execute`` with the host implementation of ``StatsMixedLM``: repeated
measurements (three visits of 60 samples), phenotype rows AND columns re-permuted before
every case, genotypes in a permuted sample order -- the same fixture as the reference's
(test_mixedlm.py:36-92).  CPU only: the mixed model is not on the B200 hot path."""

import numpy as np
import pandas as pd
import pytest

from genetest_b200 import analysis, subscribers
from genetest_b200 import modelspec as spec
from genetest_b200.genotypes import DataFrameReader
from genetest_b200.phenotypes import _DummyPhenotypes
from genetest_b200.statistics.models.mixedlm import StatsMixedLM, fit_random_intercept

from golden_support import EXPECTED, check_record, load_table

# test_mixedlm.py:66-73
MARKERS = {"snp1": ("3", 1234, "T", "G"), "snp2": ("3", 2345, "C", "A"),
           "snp3": ("2", 3456, "G", "T"), "snp4": ("22", 4567, "A", "C")}
SNPS = ["snp1", "snp2", "snp3", "snp4"]


def _fixture(rng):
    df = load_table("mixedlm")
    df.index = df["sampleid"].values
    ph = _DummyPhenotypes()
    data = df.drop(SNPS, axis=1)
    ph.data = data.iloc[rng.permutation(data.shape[0]), rng.permutation(data.shape[1])]
    ph._repeated = True
    geno = df.loc[rng.permutation(df.index.values), SNPS].copy()
    geno = geno[~geno.index.duplicated()]
    info = pd.DataFrame({"chrom": [MARKERS[s][0] for s in SNPS], "pos": [MARKERS[s][1] for s in SNPS],
                         "a1": [MARKERS[s][2] for s in SNPS], "a2": [MARKERS[s][3] for s in SNPS]},
                        index=SNPS)
    return ph, DataFrameReader(geno, info)


def _modelspec(name):
    """The model of reference test ``name`` (test_mixedlm.py:100-1200)."""
    spec._reset()
    gender = spec.factor(spec.phenotypes.gender)
    visit = spec.factor(spec.phenotypes.visit)
    outcome = {"outcome": spec.phenotypes.pheno3, "groups": spec.phenotypes.sampleid}
    inter = None
    if "gwas" in name:
        first = spec.SNPs
        if name.endswith("inter"):
            inter = spec.gwas_interaction(spec.phenotypes.var1)
    else:
        snp = name.split("_")[2]
        first = spec.genotypes[snp]
        if "_inter_" in name:
            inter = spec.interaction(spec.genotypes[snp], spec.phenotypes.var1)
    predictors = [first, spec.phenotypes.age, spec.phenotypes.var1, gender, visit]
    if inter is not None:
        predictors.append(inter)
    test = (lambda: StatsMixedLM(reml=False)) if name.endswith("_ml") else "mixedlm"
    return spec.ModelSpec(outcome=outcome, predictors=predictors, test=test), inter


@pytest.mark.parametrize("name", sorted(EXPECTED["mixedlm"]))
def test_reference_mixedlm_goldens(name, tmp_path):
    rec = EXPECTED["mixedlm"][name]
    assert not rec["expected_failure"]
    rng = np.random.default_rng(sum(name.encode()))
    phenotypes, genotypes = _fixture(rng)
    modelspec, inter = _modelspec(name)
    sub = subscribers.ResultsMemory()
    analysis.execute(phenotypes, genotypes, modelspec, subscribers=[sub],
                     output_prefix=str(tmp_path / "results"))

    def key(entity):
        return entity.replace("INTER", inter.id) if inter else entity

    if "gwas" in name:
        results = sub._get_gwas_results()
        assert len(results) == rec["n_results"]
        problems = check_record(rec, lambda snp, ent, fld: results[snp][key(ent)][fld])
    else:
        res = sub.results[0]
        problems = check_record(rec, lambda snp, ent, fld: res[key(ent)][fld])
    assert not problems, problems
    assert len(rec["checks"]) >= 14


def test_random_intercept_properties():
    """Estimates are those of the closed forms where they exist: without a group effect the
    fit is ordinary least squares; a balanced one-way layout has the ANOVA estimators; the
    BLUPs shrink the group means of the residuals; labels may be anything hashable."""
    rng = np.random.default_rng(5)
    G, m = 40, 4
    groups = np.repeat(np.arange(G), m)
    u = rng.normal(0, 2.0, G)
    x = rng.normal(size=G * m)
    y = 1.0 + 0.5 * x + u[groups] + rng.normal(0, 1.0, G * m)
    X = np.column_stack((x, np.ones(G * m)))
    fit = fit_random_intercept(y, X, groups, reml=True)
    assert abs(fit["params"][0] - 0.5) < 0.2 and fit["n_groups"] == G
    assert 1.0 < fit["cov_re"] < 9.0 and 0.5 < fit["scale"] < 2.0
    # BLUP of group g = c_g * sum of its residuals
    r = y - X @ fit["params"]
    theta = fit["cov_re"] / fit["scale"]
    want = np.array([theta / (1 + theta * m) * r[groups == g].sum() for g in range(G)])
    np.testing.assert_allclose(fit["random_effects"].values, want, rtol=1e-9)
    # intercept-only balanced layout: REML = ANOVA estimators
    y0 = u[groups] + rng.normal(0, 1.0, G * m)
    f0 = fit_random_intercept(y0, np.ones((G * m, 1)), groups, reml=True)
    means = y0.reshape(G, m).mean(axis=1)
    msw = ((y0.reshape(G, m) - means[:, None]) ** 2).sum() / (G * (m - 1))
    msb = m * ((means - means.mean()) ** 2).sum() / (G - 1)
    np.testing.assert_allclose(f0["scale"], msw, rtol=1e-6)
    np.testing.assert_allclose(f0["cov_re"], (msb - msw) / m, rtol=1e-6)
    np.testing.assert_allclose(f0["params"][0], y0.mean(), rtol=1e-10)
    # no group effect at all: theta = 0 and the OLS coefficients
    y1 = 1.0 + 0.5 * x + rng.normal(size=G * m)
    grp1 = np.array(["g{}".format(i) for i in rng.permutation(G * m) % G])
    f1 = fit_random_intercept(y1, X, grp1, reml=False)
    ols = np.linalg.lstsq(X, y1, rcond=None)[0]
    if f1["cov_re"] == 0.0:
        np.testing.assert_allclose(f1["params"], ols, rtol=1e-10)
    assert np.all(np.abs(f1["params"] - ols) < 0.05)


def test_mixedlm_requires_outcome_and_groups():
    y = pd.DataFrame({"outcome": [1.0, 2.0, 3.0]})
    X = pd.DataFrame({"intercept": [1.0, 1.0, 1.0]})
    with pytest.raises(ValueError):
        StatsMixedLM().fit(y, X)
