"""The cases of genetest/tests/test_modelspec.py replayed on
genetest_b200.modelspec: ``ModelSpec.create_data_matrix`` is the one-time
preparation of ``analysis.execute`` -- phenotype fetch,
label alignment with the genotype container, log / pow transformations,
factor encoding (first sorted level = reference), interaction products and
the ``gwas_interaction`` multiplication table."""

import numpy as np
import pandas as pd
import pytest

from genetest_b200 import modelspec as spec
from genetest_b200.genotypes import PlinkReader, PlinkWriter
from genetest_b200.phenotypes import _DummyPhenotypes


def _write_plink(prefix, snp, samples):
    with PlinkWriter(prefix) as w:
        w.write_genotypes(snp.loc[samples].values)
        w.write_bim([(1, "snp", 1, "B", "A")])
        w.write_fam(list(samples))


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    rng = np.random.default_rng(2016)
    n = 100
    data = pd.DataFrame(dict(
        pheno=rng.integers(1, 100, n), var1=rng.integers(1, 100, n),
        var2=rng.random(n),
        var3=["x{}".format(i) for i in rng.integers(0, 3, n)],
        var4=["y{}".format(i) for i in rng.integers(0, 2, n)],
        var5=rng.integers(0, 4, n), snp=rng.binomial(2, 0.3, n)),
        index=["sample_{}".format(i + 1) for i in range(n)])
    data["var5"] = data.var5.astype("category")
    prefix = str(tmp_path_factory.mktemp("spec") / "input")
    _write_plink(prefix, data.snp, rng.permutation(data.index.values))
    geno = PlinkReader(prefix)
    yield data, geno, prefix, rng
    geno.close()


@pytest.fixture
def ctx(world):
    data, geno, prefix, rng = world
    spec._reset()
    ph = _DummyPhenotypes()
    cols = ["pheno"] + ["var{}".format(i + 1) for i in range(5)]
    ph.data = data[cols].iloc[rng.permutation(data.shape[0]),
                              rng.permutation(len(cols))].copy()
    return data, geno, ph, prefix, rng


def _common(matrix, data, n_cols):
    assert matrix.shape == (data.shape[0], n_cols)
    assert matrix.intercept.unique().tolist() == [1]
    assert matrix.loc[data.index, spec.phenotypes.pheno.id].equals(data.pheno)


def test_reset():
    spec._reset()
    assert len(spec.core.TransformationManager.transformations) == 0
    assert len(spec.core.DependencyManager.dependencies) == 0


def test_simple_modelspec(ctx):
    data, geno, ph, _, _ = ctx
    predictors = [spec.genotypes.snp, spec.phenotypes.var1, spec.phenotypes.var2]
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno, predictors=predictors,
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 5)
    names = m.get_translations()
    for pr in predictors:
        np.testing.assert_array_equal(matrix.loc[data.index, pr.id].values,
                                      data[names[pr.id]].values)


def test_no_intersect(ctx):
    data, geno, _, _, _ = ctx
    ph = _DummyPhenotypes()
    ph.data = data[["pheno", "var1"]].copy()
    ph.data.index = ["s" + s for s in ph.data.index]
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=[spec.genotypes.snp, spec.phenotypes.var1],
                       test="linear")
    with pytest.raises(ValueError):
        m.create_data_matrix(ph, geno)


def test_smaller_intersect(ctx):
    data, _, _, prefix, rng = ctx
    drop = rng.choice(data.index, 10, replace=False)
    ph = _DummyPhenotypes()
    ph.data = data.drop(drop[:5], axis=0)
    geno_data = data.drop(drop[5:], axis=0)
    _write_plink(prefix + "_less", geno_data.snp, geno_data.index)
    predictors = [spec.genotypes.snp, spec.phenotypes.var1, spec.phenotypes.var2]
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno, predictors=predictors,
                       test="linear")
    with PlinkReader(prefix + "_less") as geno:
        matrix = m.create_data_matrix(ph, geno)
    kept = data.drop(drop, axis=0)
    _common(matrix, kept, 5)
    names = m.get_translations()
    for pr in predictors:
        np.testing.assert_array_equal(matrix.loc[kept.index, pr.id].values,
                                      kept[names[pr.id]].values)


def test_pow_and_log(ctx):
    data, geno, ph, _, _ = ctx
    predictors = [spec.genotypes.snp, spec.pow(spec.phenotypes.var1, 3),
                  spec.log10(spec.phenotypes.var1), spec.ln(spec.phenotypes.var1)]
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno, predictors=predictors,
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 6)
    want = [data.snp.values, data.var1.values ** 3,
            np.log10(data.var1.values), np.log(data.var1.values)]
    for pr, w in zip(predictors, want):
        np.testing.assert_array_equal(matrix.loc[data.index, pr.id].values, w)


FACTORS = (("var3", ("x1", "x2")), ("var4", ("y1",)), ("var5", (1, 2, 3)))


def test_encode_factor(ctx):
    data, geno, ph, _, _ = ctx
    predictors = [spec.factor(spec.phenotypes.var3),
                  spec.factor(spec.phenotypes.var4),
                  spec.factor(spec.phenotypes.var5)]
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno, predictors=predictors,
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 8)
    for pr, (name, levels) in zip(predictors, FACTORS):
        for level in levels:
            np.testing.assert_array_equal(
                matrix.loc[data.index, "{}:level.{}".format(pr.id, level)].values,
                (data[name] == level).astype(float).values)


@pytest.mark.parametrize("extra", [0, 1])
def test_simple_and_multiple_interaction(ctx, extra):
    # (one interaction per model: its id derives from its first term)
    data, geno, ph, _, _ = ctx
    predictors = [spec.genotypes.snp, spec.phenotypes.var1, spec.phenotypes.var2]
    terms = predictors[:2 + extra]
    inter = spec.interaction(*terms)
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=terms + [inter], test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 5 + extra)
    want = data.snp * data.var1 * (data.var2 if extra else 1)
    np.testing.assert_array_equal(matrix.loc[data.index, inter.id].values,
                                  want.values)


def test_factor_interaction(ctx):
    data, geno, ph, _, _ = ctx
    factor = spec.factor(spec.phenotypes.var5)
    inter = spec.interaction(spec.genotypes.snp, factor)
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=[spec.genotypes.snp, factor, inter],
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 9)
    np.testing.assert_array_equal(
        matrix.loc[data.index, spec.genotypes.snp.id].values, data.snp.values)
    for level in (1, 2, 3):
        ind = (data.var5 == level).astype(float)
        np.testing.assert_array_equal(
            matrix.loc[data.index, "{}:level.{}".format(factor.id, level)].values,
            ind.values)
        np.testing.assert_array_equal(
            matrix.loc[data.index, "{}:level.{}".format(inter.id, level)].values,
            (ind * data.snp).values)


def test_complex_interaction(ctx):
    data, geno, ph, _, _ = ctx
    f3, f4, f5 = (spec.factor(spec.phenotypes.var3),
                  spec.factor(spec.phenotypes.var4),
                  spec.factor(spec.phenotypes.var5))
    predictors = [spec.genotypes.snp, spec.phenotypes.var2, f3, f4, f5]
    inter = spec.interaction(spec.genotypes.snp, spec.phenotypes.var2, f3, f4, f5)
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=predictors + [inter], test="linear")
    matrix = m.create_data_matrix(ph, geno)
    _common(matrix, data, 16)
    for l3 in ("x1", "x2"):
        for l4 in ("y1",):
            for l5 in (1, 2, 3):
                want = (data.snp * data.var2 *
                        (data.var3 == l3).astype(float) *
                        (data.var4 == l4).astype(float) *
                        (data.var5 == l5).astype(float))
                col = "{}:level.{}:level.{}:level.{}".format(inter.id, l3, l4, l5)
                np.testing.assert_array_equal(matrix.loc[data.index, col].values,
                                              want.values)


def test_gwas_interaction(ctx):
    data, geno, ph, _, _ = ctx
    inter = spec.gwas_interaction(spec.phenotypes.var1)
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=[spec.SNPs, spec.phenotypes.var1, inter],
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    assert m.has_gwas_interaction
    _common(matrix, data, 3)
    np.testing.assert_array_equal(
        matrix.loc[data.index, spec.phenotypes.var1.id].values, data.var1.values)
    table = m.gwas_interaction
    assert list(table) == [inter.id]
    assert list(table[inter.id]) == [spec.phenotypes.var1.id]


def test_gwas_interaction_with_factors(ctx):
    data, geno, ph, _, _ = ctx
    f3, f5 = spec.factor(spec.phenotypes.var3), spec.factor(spec.phenotypes.var5)
    inter = spec.gwas_interaction(spec.phenotypes.var1, f3, f5)
    m = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                       predictors=[spec.SNPs, spec.phenotypes.var1, f3, f5, inter],
                       test="linear")
    matrix = m.create_data_matrix(ph, geno)
    assert m.has_gwas_interaction
    _common(matrix, data, 8)
    table = m.gwas_interaction
    assert len(table) == 6
    for l3 in ("x1", "x2"):
        for l5 in (1, 2, 3):
            cols = table["{}:level.{}:level.{}".format(inter.id, l3, l5)]
            assert len(cols) == 3
            assert set(cols) == {spec.phenotypes.var1.id,
                                 "{}:level.{}".format(f3.id, l3),
                                 "{}:level.{}".format(f5.id, l5)}


# --------------------------------------------------------------------------
# the same matrices through the formula grammar (genetest/tests/test_grammar.py)
# --------------------------------------------------------------------------
def _from_formula(formula, ph, geno):
    from genetest_b200.statistics.models.linear import StatsLinear
    model = spec.parse_formula(formula)
    model["test"] = StatsLinear
    del model["conditions"]
    m = spec.ModelSpec(**model)
    return m, m.create_data_matrix(ph, geno)


def _col(m, matrix, data, name):
    """Column of the matrix whose entity translates to ``name``."""
    ids = [i for i, n in m.get_translations().items() if n == name]
    assert len(ids) == 1, (name, m.get_translations())
    return matrix.loc[data.index, ids[0]].values


def _outcome(m, matrix, data):
    assert matrix.intercept.unique().tolist() == [1]
    np.testing.assert_array_equal(_col(m, matrix, data, "pheno"),
                                  data.pheno.values)


def test_formula_simple(ctx):
    data, geno, ph, _, _ = ctx
    m, matrix = _from_formula("pheno ~ g(snp) + var1 + var2", ph, geno)
    assert matrix.shape == (data.shape[0], 5)
    _outcome(m, matrix, data)
    for name in ("snp", "var1", "var2"):
        np.testing.assert_array_equal(_col(m, matrix, data, name),
                                      data[name].values)


def test_formula_pow_and_log(ctx):
    data, geno, ph, _, _ = ctx
    m, matrix = _from_formula("pheno ~ g(snp) + pow(var1, 4)", ph, geno)
    assert matrix.shape == (data.shape[0], 4)
    _outcome(m, matrix, data)
    got = sorted(matrix.drop(columns=["intercept"]).sum().tolist())
    assert float((data.var1.values.astype(float) ** 4).sum()) in got
    spec._reset()
    m, matrix = _from_formula(
        "pheno ~ g(snp) + var1 + log10(var1) + ln(var1)", ph, geno)
    assert matrix.shape == (data.shape[0], 6)
    _outcome(m, matrix, data)
    v = _col(m, matrix, data, "var1")
    np.testing.assert_array_equal(v, data.var1.values)
    others = matrix.loc[data.index].drop(
        columns=["intercept"] + [i for i, n in m.get_translations().items()
                                 if n in ("pheno", "snp", "var1")])
    assert others.shape[1] == 2
    want = {tuple(np.log10(data.var1.values)), tuple(np.log(data.var1.values))}
    assert {tuple(others[c].values) for c in others.columns} == want


def test_formula_factors_and_interactions(ctx):
    data, geno, ph, _, _ = ctx
    m, matrix = _from_formula(
        "pheno ~ factor(var3) + factor(var4) + factor(var5)", ph, geno)
    assert matrix.shape == (data.shape[0], 8)
    _outcome(m, matrix, data)
    for name, levels in FACTORS:
        for level in levels:
            hits = [c for c in matrix.columns
                    if c.endswith(":level.{}".format(level))]
            assert len(hits) == 1
            np.testing.assert_array_equal(
                matrix.loc[data.index, hits[0]].values,
                (data[name] == level).astype(float).values)
    spec._reset()
    m, matrix = _from_formula("pheno ~ g(snp) + var1 + g(snp)*var1", ph, geno)
    assert matrix.shape == (data.shape[0], 5)
    inter = [c for c in matrix.columns if c.startswith("TRANSFORM:INTERACTION")]
    assert len(inter) == 1
    np.testing.assert_array_equal(matrix.loc[data.index, inter[0]].values,
                                  (data.snp * data.var1).values)
    spec._reset()
    m, matrix = _from_formula(
        "pheno ~ g(snp) + var1 + var2 + g(snp)*var1*var2", ph, geno)
    assert matrix.shape == (data.shape[0], 6)
    inter = [c for c in matrix.columns if c.startswith("TRANSFORM:INTERACTION")]
    np.testing.assert_array_equal(
        matrix.loc[data.index, inter[0]].values,
        (data.snp * data.var1 * data.var2).values)
    spec._reset()
    m, matrix = _from_formula(
        "pheno ~ g(snp) + factor(var5) + g(snp)*factor(var5)", ph, geno)
    assert matrix.shape == (data.shape[0], 9)
    for level in (1, 2, 3):
        col = [c for c in matrix.columns
               if c.startswith("TRANSFORM:INTERACTION")
               and c.endswith(":level.{}".format(level))]
        assert len(col) == 1
        np.testing.assert_array_equal(
            matrix.loc[data.index, col[0]].values,
            ((data.var5 == level).astype(float) * data.snp).values)
    spec._reset()
    m, matrix = _from_formula(
        "pheno ~ g(snp) + var2 + factor(var3) + factor(var4) + "
        "        factor(var5) + "
        "        g(snp)*var2*factor(var3)*factor(var4)*factor(var5)", ph, geno)
    assert matrix.shape == (data.shape[0], 16)
    for l3 in ("x1", "x2"):
        for l5 in (1, 2, 3):
            suffix = ":level.{}:level.y1:level.{}".format(l3, l5)
            col = [c for c in matrix.columns if c.endswith(suffix)]
            assert len(col) == 1
            want = (data.snp * data.var2 * (data.var3 == l3).astype(float) *
                    (data.var4 == "y1").astype(float) *
                    (data.var5 == l5).astype(float))
            np.testing.assert_array_equal(matrix.loc[data.index, col[0]].values,
                                          want.values)
