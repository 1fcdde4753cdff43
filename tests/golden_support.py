"""Shared helpers for replaying the reference's known-answer tests.

The numbers live in ``tests/golden/expected.json`` (made by
``tests/golden/make_golden.py`` from ``genetest/tests/test_{linear,logistic,
coxph}.py``); the data tables in ``tests/golden/data/*.tsv``.  ``CASES`` says
how each reference test builds its model (read off the reference test
sources: predictors, interaction term, model kwargs, which genotype file).
"""

import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

with open(os.path.join(GOLDEN, "expected.json")) as _f:
    EXPECTED = json.load(_f)

# marker map used by the reference fixtures (test_linear.py:96-101,
# test_logistic.py:90-96, test_coxph.py:66-73): name -> (chrom, pos, a1, a2)
MARKERS = {
    "snp1": ("3", 1234, "T", "C"), "snp2": ("3", 9618, "C", "A"),
    "snp3": ("2", 1519, "G", "T"), "snp4": ("1", 5871, "G", "A"),
    "snp5": ("23", 2938, "T", "C"),
}
OUTCOME = {"linear": ["pheno1"], "logistic": ["pheno2"],
           "coxph": ["tte", "event"]}
NSNPS = {"linear": 5, "logistic": 3, "coxph": 4}


def _cases():
    cases = {}
    big = {"condition_value_t": 15000}
    cases["linear"] = {
        "test_linear_gwas": dict(data="linear", mode="gwas"),
        "test_linear_gwas_inter": dict(data="linear", mode="gwas",
                                       inter="var1", kwargs=big),
        "test_linear_gwas_inter_with_missing": dict(
            data="linear_missing", mode="gwas", inter="gender", kwargs=big),
        "test_linear_snp1_inter_categorical": dict(
            data="linear", mode="single", snp="snp1", inter="gender"),
    }
    for k in range(1, 6):
        s = "snp{}".format(k)
        cases["linear"]["test_linear_" + s] = dict(
            data="linear", mode="single", snp=s)
        cases["linear"]["test_linear_" + s + "_inter"] = dict(
            data="linear", mode="single", snp=s, inter="var1",
            kwargs=big if k == 1 else {})
    cases["logistic"] = {
        "test_logistic_gwas": dict(data="logistic", mode="gwas"),
        "test_logistic_gwas_inter": dict(data="logistic", mode="gwas",
                                         inter="var1"),
        "test_logistic_gwas_inter_with_missing": dict(
            data="logistic_missing", mode="gwas", inter="gender"),
        "test_logistic_snp1_inter_categorical": dict(
            data="logistic", mode="single", snp="snp1", inter="gender"),
    }
    for k in range(1, 4):
        s = "snp{}".format(k)
        cases["logistic"]["test_logistic_" + s] = dict(
            data="logistic", mode="single", snp=s)
        cases["logistic"]["test_logistic_" + s + "_inter"] = dict(
            data="logistic", mode="single", snp=s, inter="var1",
            # reference quirk (test_logistic.py:670): snp3's test multiplies
            # var1 with snp2
            inter_snp="snp2" if k == 3 else s)
    cases["coxph"] = {
        "test_coxph_gwas": dict(data="coxph", mode="gwas"),
        "test_coxph_snp1_inter_categorical": dict(
            data="coxph", mode="single", snp="snp1", inter="gender"),
    }
    for k in range(1, 5):
        s = "snp{}".format(k)
        cases["coxph"]["test_coxph_" + s] = dict(
            data="coxph", mode="single", snp=s)
        cases["coxph"]["test_coxph_" + s + "_inter"] = dict(
            data="coxph", mode="single", snp=s, inter="var1")
    for model in cases:
        for name, c in cases[model].items():
            c.setdefault("inter", None)
            c.setdefault("kwargs", {})
            c["model"] = model
            c["name"] = name
    return cases


CASES = _cases()


def case_ids(model):
    return [n for n in sorted(CASES[model])
            if not EXPECTED[model][n]["expected_failure"]]


def load_table(name):
    df = pd.read_csv(os.path.join(GOLDEN, "data", name + ".tsv"), sep="\t")
    if "sample_id" in df.columns:
        return df.set_index("sample_id")
    df.index = ["s{}".format(i + 1) for i in range(df.shape[0])]
    df.index.name = "sample"
    return df


def covariates(df):
    """[age, var1, gender:level.2, intercept] as the reference's
    create_data_matrix lays them out (predictor order, intercept last;
    gender in {1, 2}, first sorted level is the reference level)."""
    C = np.column_stack((df.age.values.astype(float),
                         df.var1.values.astype(float),
                         (df.gender.values == 2).astype(float),
                         np.ones(df.shape[0])))
    return C, ["age", "var1", "gender:level.2", "intercept"]


def almost_equal(a, b, places):
    """unittest.assertAlmostEqual semantics."""
    if a == b:
        return True
    return round(abs(a - b), places) == 0


def transform(value, how):
    if how == "identity":
        return value
    if how == "neglog10":
        return -np.log10(value)
    if how == "square":
        return value ** 2
    raise ValueError(how)


def check_record(rec, lookup):
    """Apply every extracted assertion of one reference test.

    ``lookup(snp, entity, field)`` returns the produced value (``snp`` is None
    for single-variant tests)."""
    problems = []
    for c in rec["checks"]:
        got = lookup(c["snp"], c["entity"], c["field"])
        exp = c["expected"]
        if c["places"] is None:
            ok = (got == exp)
        else:
            ok = almost_equal(transform(got, c["transform"]), exp, c["places"])
        if not ok:
            problems.append((c["snp"], c["entity"], c["field"], got, exp))
    return problems
