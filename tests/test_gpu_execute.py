"""Replay the reference's known-answer tests through ``analysis.execute``.

Same fixtures and assertions as genetest/tests/test_linear.py,
test_logistic.py, test_coxph.py and test_gwaswriter.py (numbers from
``tests/golden/expected.json``): PLINK files written with a permuted sample
order, phenotype rows AND columns re-permuted before every case, so that
alignment is proven to be by label; results come out of the CUDA engine.
"""

import os

import numpy as np
import pandas as pd
import pytest

from genetest_b200 import analysis, subscribers
from genetest_b200 import modelspec as spec
from genetest_b200.genotypes import DataFrameReader, PlinkReader, PlinkWriter
from genetest_b200.phenotypes import _DummyPhenotypes
from genetest_b200.statistics.core import StatsError
from genetest_b200.statistics.models.linear import StatsLinear
from genetest_b200.subscribers import GWASWriter

from golden_support import (CASES, EXPECTED, MARKERS, NSNPS, OUTCOME,
                            case_ids, check_record, load_table)

pytestmark = pytest.mark.gpu


def _phenotypes(df, model, rng):
    ph = _DummyPhenotypes()
    data = df.drop([c for c in df.columns if c.startswith("snp")], axis=1)
    ph.data = data.iloc[rng.permutation(data.shape[0]),
                        rng.permutation(data.shape[1])]
    return ph


def _plink(df, prefix, order, snps):
    with PlinkWriter(prefix) as w:
        for s in snps:
            w.write_genotypes(df.loc[order, s].fillna(-1).values)
        w.write_bim([(MARKERS[s][0], s, MARKERS[s][1], MARKERS[s][2],
                      MARKERS[s][3]) for s in snps])
        w.write_fam(order)
    return PlinkReader(prefix)


def _dataframe_reader(df, order, snps):
    info = pd.DataFrame(
        {"chrom": [MARKERS[s][0] for s in snps],
         "pos": [MARKERS[s][1] for s in snps],
         "a1": [MARKERS[s][2] for s in snps],
         "a2": [MARKERS[s][3] for s in snps]}, index=snps)
    return DataFrameReader(df.loc[order, snps].copy(), info)


def _build_modelspec(case):
    spec._reset()
    model = case["model"]
    gender = spec.factor(spec.phenotypes.gender)
    if model == "coxph":
        outcome = dict(tte=spec.phenotypes.tte, event=spec.phenotypes.event)
    else:
        outcome = spec.phenotypes[OUTCOME[model][0]]
    inter = None
    if case["mode"] == "gwas":
        first = spec.SNPs
        if case["inter"] == "var1":
            inter = spec.gwas_interaction(spec.phenotypes.var1)
        elif case["inter"] == "gender":
            inter = spec.gwas_interaction(gender)
    else:
        first = spec.genotypes[case["snp"]]
        other = spec.genotypes[case.get("inter_snp", case["snp"])]
        if case["inter"] == "var1":
            inter = spec.interaction(spec.phenotypes.var1, other)
        elif case["inter"] == "gender":
            inter = spec.interaction(gender, other)
    predictors = [first, spec.phenotypes.age, spec.phenotypes.var1, gender]
    if inter is not None:
        predictors.append(inter)
    if model == "linear" and case["kwargs"]:
        test = lambda: StatsLinear(**case["kwargs"])    # noqa: E731
    else:
        test = model
    return spec.ModelSpec(outcome=outcome, predictors=predictors,
                          test=test), inter


@pytest.mark.parametrize("model", ["linear", "logistic", "coxph"])
def test_reference_goldens_through_execute(model, tmp_path):
    rng = np.random.default_rng(12345)
    snps = ["snp{}".format(i + 1) for i in range(NSNPS[model])]
    readers = {}
    for name in case_ids(model):
        case = CASES[model][name]
        rec = EXPECTED[model][name]
        df = load_table(case["data"])
        if case["data"] not in readers:
            order = rng.permutation(df.index.values)
            if model == "linear":
                readers[case["data"]] = _plink(
                    df, str(tmp_path / case["data"]), order, snps)
            else:
                readers[case["data"]] = _dataframe_reader(df, order, snps)
        genotypes = readers[case["data"]]
        phenotypes = _phenotypes(df, model, rng)
        modelspec, inter = _build_modelspec(case)
        sub = subscribers.ResultsMemory()
        prefix = str(tmp_path / (name + "_results"))

        def key(entity):
            return entity.replace("INTER", inter.id) if inter else entity

        if "raises" in rec:
            with pytest.raises(StatsError) as err:
                analysis.execute(phenotypes, genotypes, modelspec,
                                 subscribers=[sub])
            assert str(err.value) == rec["raises"], name
            continue
        analysis.execute(phenotypes, genotypes, modelspec, subscribers=[sub],
                         output_prefix=prefix)
        if case["mode"] == "gwas":
            results = sub._get_gwas_results()
            assert len(results) == rec["n_results"], name
            with open(prefix + "_failed_snps.txt") as f:
                failed = [line.split("\t") for line in f.read().splitlines()]
            assert failed == rec["failed"], name
            problems = check_record(
                rec, lambda snp, ent, fld: results[snp][key(ent)][fld])
        else:
            res = sub.results[0]
            problems = check_record(
                rec, lambda snp, ent, fld: res[key(ent)][fld])
        assert not problems, (name, problems)


@pytest.fixture()
def factors(tmp_path):
    df = load_table("factors")
    rng = np.random.default_rng(99)
    order = rng.permutation(df.index.values)
    prefix = str(tmp_path / "input")
    with PlinkWriter(prefix) as w:
        for s in ("snp1", "snp2", "snp3"):
            w.write_genotypes(df.loc[order, s].values)
        w.write_bim([(1, "snp{}".format(i), i, "B", "A") for i in (1, 2, 3)])
        w.write_fam(order)
    ph = _DummyPhenotypes()
    data = df.drop(["snp1", "snp2", "snp3"], axis=1)
    ph.data = data.iloc[rng.permutation(data.shape[0]),
                        rng.permutation(data.shape[1])]
    spec._reset()
    out = str(tmp_path / "output.txt")
    writer = GWASWriter(out, "linear")
    yield ph, PlinkReader(prefix), writer, out
    writer.close()


COMMON = {"snp", "chr", "pos", "major", "minor", "maf", "n", "ll", "adj_r2"}
STATS = ("coef", "se", "lower", "upper", "t", "p")


def _table(path):
    df = pd.read_csv(path, sep="\t")
    return df.set_index("snp", drop=False).loc[["snp1", "snp2", "snp3"], :]


def test_gwaswriter_normal(factors):
    """test_gwaswriter.py:101-134."""
    ph, geno, writer, out = factors
    analysis.execute_formula(
        ph, geno, "pheno ~ SNPs + var1 + var2 + factor(var3) + factor(var4) "
        "+ factor(var5)", "linear", subscribers=[writer], output_prefix="")
    df = _table(out)
    assert set(df.columns) == COMMON | set(STATS)
    np.testing.assert_array_almost_equal(
        df.coef.values,
        [6.48616317675085, 4.29701960321627, -1.44891512389655])


def test_gwaswriter_subgroups(factors):
    """test_gwaswriter.py:136-177."""
    ph, geno, writer, out = factors
    analysis.execute_formula(
        ph, geno, "pheno | var4 ~ SNPs + var1 + var2 + factor(var3) + "
        "factor(var5)", "linear", subscribers=[writer], output_prefix="")
    df = _table(out)
    assert set(df.columns) == COMMON | set(STATS) | {"subgroup"}
    np.testing.assert_array_almost_equal(
        df.loc[df.subgroup == "var4:y0", "coef"].values,
        [-0.6798617801022, 5.9538885434287, -15.1758958532932])
    np.testing.assert_array_almost_equal(
        df.loc[df.subgroup == "var4:y1", "coef"].values,
        [12.69000471889982, 4.2496912087917, 7.5291802846083])


def test_gwaswriter_interactions(factors, tmp_path):
    """test_gwaswriter.py:179-267."""
    ph, geno, writer, out = factors
    analysis.execute_formula(
        ph, geno, "pheno ~ SNPs + var1 + var2 + factor(var3) + factor(var4) "
        "+ factor(var5) + SNPs*var1", "linear", subscribers=[writer],
        output_prefix="")
    df = _table(out)
    inter = "TRANSFORM:GWAS_INTER"
    assert set(df.columns) == COMMON | {inter + ":" + s for s in STATS}
    np.testing.assert_array_almost_equal(
        df[inter + ":coef"].values,
        [0.4073975765317, -0.04253919820585, -0.05814970176832])

    spec._reset()
    out2 = str(tmp_path / "output2.txt")
    writer2 = GWASWriter(out2, "linear")
    analysis.execute_formula(
        ph, geno, "pheno ~ SNPs + var1 + var2 + factor(var3) + factor(var4) "
        "+ factor(var5) + SNPs*factor(var3)", "linear", subscribers=[writer2],
        output_prefix=out2)
    df = _table(out2)
    cols = {"{}:level.{}:{}".format(inter, lv, s)
            for lv in ("x1", "x2") for s in STATS}
    assert set(df.columns) == COMMON | cols
    np.testing.assert_array_almost_equal(
        df[inter + ":level.x1:coef"].values,
        [6.15574045825434, 2.2178331615269, -1.94057337856961])
    np.testing.assert_array_almost_equal(
        df[inter + ":level.x2:coef"].values,
        [21.47033225862596, 2.7136560668229, -0.05622018358578])


def test_gwaswriter_interaction_subgroups(factors):
    """test_gwaswriter.py:269-334."""
    ph, geno, writer, out = factors
    analysis.execute_formula(
        ph, geno, "pheno | var4 ~ SNPs + var1 + var2 + factor(var3) + "
        "factor(var5) + SNPs*factor(var3)", "linear", subscribers=[writer],
        output_prefix=out)
    df = _table(out)
    inter = "TRANSFORM:GWAS_INTER"
    y0, y1 = df.subgroup == "var4:y0", df.subgroup == "var4:y1"
    np.testing.assert_array_almost_equal(
        df.loc[y0, inter + ":level.x1:coef"].values,
        [7.9838584795821, 0.8787655737690, -5.4257289463845])
    np.testing.assert_array_almost_equal(
        df.loc[y0, inter + ":level.x2:coef"].values,
        [9.5318845882662, 2.1792184593254, -3.8901765902700])
    np.testing.assert_array_almost_equal(
        df.loc[y1, inter + ":level.x1:coef"].values,
        [-7.33292032058131, -2.53148534004492, -14.560258817242])
    np.testing.assert_array_almost_equal(
        df.loc[y1, inter + ":level.x2:coef"].values,
        [9.91504894770588, -1.35580107828104, -3.1717946298856])


def test_name_filter_and_maf_threshold(tmp_path):
    """variant predicates and the MAF failure line (analysis.py:151-153,
    322-325)."""
    from genetest_b200.modelspec.predicates import NameFilter
    df = load_table("linear")
    rng = np.random.default_rng(5)
    snps = ["snp{}".format(i + 1) for i in range(5)]
    geno = _plink(df, str(tmp_path / "in"), rng.permutation(df.index.values),
                  snps)
    ph = _phenotypes(df, "linear", rng)
    modelspec, _ = _build_modelspec(CASES["linear"]["test_linear_gwas"])
    sub = subscribers.ResultsMemory()
    prefix = str(tmp_path / "out")
    analysis.execute(ph, geno, modelspec, subscribers=[sub],
                     variant_predicates=[NameFilter({"snp1", "snp2", "snp4"})],
                     output_prefix=prefix, maf_t=0.05)
    assert sorted(sub._get_gwas_results()) == ["snp2", "snp4"]
    assert open(prefix + "_not_analyzed_snps.txt").read().split() == \
        ["snp3", "snp5"]
    assert open(prefix + "_failed_snps.txt").read().splitlines() == \
        ["snp1\tMAF: 0.016666666666666666 < 0.05"]


@pytest.mark.parametrize("reader", ["plink", "dataframe"])
def test_sexual_chromosome_maf_through_execute(tmp_path, reader):
    """execute(sexual_chromosome=True): MAF, minor / major and the flip of
    every marker follow get_sex_maf (analysis.py:141-148); checked against the
    oracle's loop body for a .bed source and a decoded source whose samples
    are a permuted superset of the phenotypes'."""
    import oracle.gwas as og
    from genetest_b200.phenotypes import DataFramePhenotypes
    rng = np.random.default_rng(17)
    n_file, n, n_snps = 260, 200, 24
    ids = np.array(["s{:03d}".format(i) for i in range(n_file)])
    pheno = pd.DataFrame({
        "y": rng.normal(size=n), "age": rng.normal(50, 10, n),
        "sex": rng.integers(0, 2, n).astype(float)},
        index=rng.permutation(ids)[:n])
    pheno.loc[pheno.index[:7], "sex"] = np.nan
    G = rng.binomial(2, rng.uniform(0.1, 0.9, n_snps)[:, None],
                     (n_snps, n_file)).astype(float)
    G[rng.random(G.shape) < 0.03] = np.nan
    names = ["rs{}".format(j) for j in range(n_snps)]
    if reader == "plink":
        with PlinkWriter(str(tmp_path / "x")) as w:
            for j in range(n_snps):
                w.write_genotypes(np.nan_to_num(G[j], nan=-1))
            w.write_bim([(23, nm, 100 + j, "A", "G")
                         for j, nm in enumerate(names)])
            w.write_fam(list(ids))
        geno = PlinkReader(str(tmp_path / "x"))
    else:
        info = pd.DataFrame({"chrom": 23, "pos": np.arange(n_snps) + 100,
                             "a1": "A", "a2": "G"}, index=names)
        geno = DataFrameReader(pd.DataFrame(G.T, index=ids, columns=names),
                               info)
    spec._reset()
    sub = subscribers.ResultsMemory()
    analysis.execute_formula(
        DataFramePhenotypes(pheno, sex_column="sex"), geno,
        "y ~ SNPs + age", "linear", subscribers=[sub],
        output_prefix=str(tmp_path / "out"), sexual_chromosome=True)
    got = sub._get_gwas_results()
    # the oracle's loop body on the complete cases, design rows in file order
    keep = pheno.dropna(subset=["y", "age"])
    order = [s for s in ids if s in set(keep.index)]
    cols = np.array([np.flatnonzero(ids == s)[0] for s in order])
    C = np.column_stack((keep.loc[order, "age"].values, np.ones(len(order))))
    sexes = keep.loc[order, "sex"].values
    assert len(got) == n_snps
    flips = 0
    for j, nm in enumerate(names):
        st, o = og.gwas_one("linear", keep.loc[order, ["y"]].values, C,
                            ["age", "intercept"], G[j, cols], sexes=sexes,
                            coded="A", reference="G")
        assert st == "ok"
        r = got[nm]["SNPs"]
        assert r["maf"] == o["SNPs"]["maf"]
        assert (r["minor"], r["major"]) == (o["SNPs"]["minor"],
                                            o["SNPs"]["major"])
        assert abs(r["coef"] - o["SNPs"]["coef"]) <= 1e-8 * abs(r["coef"])
        flips += o["_flip"]
    assert 3 < flips < n_snps - 3


@pytest.mark.parametrize("test", ["linear", "logistic"])
def test_impute2_dosages_through_execute(tmp_path, test):
    """Imputed (real-valued) dosages from an IMPUTE2 file: the decoded-block
    path aligns the samples on the host and runs the FP64 dosage kernels;
    checked against the oracle's loop body."""
    import oracle.gwas as og
    from genetest_b200.genotypes import Impute2Reader
    from genetest_b200.phenotypes import DataFramePhenotypes
    rng = np.random.default_rng(23)
    n_file, n, n_snps = 340, 300, 12
    ids = ["s{:03d}".format(i) for i in range(n_file)]
    (tmp_path / "d.sample").write_text(
        "ID_1 ID_2 missing\n0 0 0\n" +
        "".join("f{0} {0} 0\n".format(s) for s in ids))
    lines, D = [], np.zeros((n_snps, n_file))
    for j in range(n_snps):
        maf = rng.uniform(0.1, 0.9)
        g = rng.binomial(2, maf, n_file)
        probs = np.full((n_file, 3), 0.01)
        probs[np.arange(n_file), g] = 0.98
        blur = rng.random(n_file) < 0.15                 # uncertain calls
        probs[blur] = rng.dirichlet([2, 2, 2], blur.sum())
        probs = np.round(probs, 4)
        d = probs[:, 1] + 2 * probs[:, 2]
        d[probs.max(axis=1) < 0.9] = np.nan
        D[j] = d
        lines.append("1 rs{} {} A G ".format(j, 100 + j) +
                     " ".join("{:.4f}".format(v) for v in probs.ravel()))
    (tmp_path / "d.impute2").write_text("\n".join(lines) + "\n")
    keep = rng.permutation(n_file)[:n]
    age = rng.normal(50, 10, n)
    if test == "linear":
        y = 0.02 * age + rng.normal(size=n)
    else:
        y = (rng.random(n) < 0.45).astype(float)
    pheno = pd.DataFrame({"y": y, "age": age}, index=[ids[i] for i in keep])
    spec._reset()
    sub = subscribers.ResultsMemory()
    analysis.execute_formula(
        DataFramePhenotypes(pheno),
        Impute2Reader(str(tmp_path / "d.impute2"), str(tmp_path / "d.sample")),
        "y ~ SNPs + age", test, subscribers=[sub],
        output_prefix=str(tmp_path / "out"))
    got = sub._get_gwas_results()
    order = np.sort(keep)                                # design rows in file order
    rows = pheno.loc[[ids[i] for i in order]]
    C = np.column_stack((rows["age"].values, np.ones(n)))
    tol = 1e-8 if test == "linear" else 1e-6
    n_ok = 0
    for j in range(n_snps):
        st, o = og.gwas_one(test, rows[["y"]].values, C, ["age", "intercept"],
                            D[j, order], coded="G", reference="A")
        name = "rs{}".format(j)
        if st != "ok":
            assert name not in got
            continue
        n_ok += 1
        r = got[name]
        assert abs(r["SNPs"]["maf"] - o["SNPs"]["maf"]) < 1e-14
        assert (r["SNPs"]["minor"], r["SNPs"]["major"]) == \
            (o["SNPs"]["minor"], o["SNPs"]["major"])
        for key in ("SNPs", "age", "intercept"):
            for f in ("coef", "std_err", "p_value"):
                assert abs(r[key][f] - o[key][f]) <= tol * abs(o[key][f]), \
                    (name, key, f)
        assert r["MODEL"]["nobs"] == o["MODEL"]["nobs"]
    assert n_ok >= n_snps - 2


def test_maf_filter_runs_on_the_packed_path(tmp_path):
    """``MAFFilter`` (predicates.py:33-57: MAF over ALL samples of the file,
    phenotyped or not) is evaluated on the device for .bed sources; a subclass
    -- any predicate the engine does not know -- sends the same analysis through
    the reference's per-marker loop.  Both give the same rows, the same skipped
    list and the same failure lines, and a subscriber that overrides ``handle``
    (the reference's extension point) is called once per marker on either path."""
    from genetest_b200.modelspec.predicates import MAFFilter, NameFilter
    from genetest_b200.phenotypes import DataFramePhenotypes
    rng = np.random.default_rng(77)
    n_file, n, n_snps = 500, 420, 96
    ids = np.array(["s{:03d}".format(i) for i in range(n_file)])
    maf = np.concatenate([rng.uniform(0.001, 0.08, n_snps // 2),
                          rng.uniform(0.08, 0.5, n_snps - n_snps // 2)])
    G = rng.binomial(2, maf[:, None], size=(n_snps, n_file)).astype(float)
    G[rng.random(G.shape) < 0.03] = np.nan
    G[5] = np.nan                                        # nothing genotyped: always skipped
    names = ["rs{}".format(j) for j in range(n_snps)]
    with PlinkWriter(str(tmp_path / "g")) as w:
        for j in range(n_snps):
            w.write_genotypes(np.nan_to_num(G[j], nan=-1))
        w.write_bim([(1, nm, 100 + j, "A", "G") for j, nm in enumerate(names)])
        w.write_fam(list(ids))
    keep = rng.permutation(n_file)[:n]
    pheno = pd.DataFrame({"y": rng.normal(size=n), "age": rng.normal(50, 10, n)},
                         index=ids[keep])

    class CountingWriter(GWASWriter):
        calls = 0

        def handle(self, results):
            type(self).calls += 1
            return super().handle(results)

    class OtherMAF(MAFFilter):
        pass

    outs = {}
    for tag, pred in (("packed", MAFFilter(0.05)), ("decoded", OtherMAF(0.05))):
        spec._reset()
        CountingWriter.calls = 0
        prefix = str(tmp_path / tag)
        sub = CountingWriter(prefix + ".tsv", "linear")
        analysis.execute_formula(
            DataFramePhenotypes(pheno), PlinkReader(str(tmp_path / "g")),
            "y ~ SNPs + age", "linear", subscribers=[sub],
            variant_predicates=[pred, NameFilter(set(names[3:]))],
            output_prefix=prefix, maf_t=0.06)
        outs[tag] = (open(prefix + ".tsv").read(),
                     sorted(open(prefix + "_not_analyzed_snps.txt").read().split()),
                     sorted(open(prefix + "_failed_snps.txt").read().splitlines()),
                     CountingWriter.calls)
    assert outs["packed"][1] == outs["decoded"][1]
    assert outs["packed"][2] == outs["decoded"][2]
    # the same rows (the two paths sum in different orders: equal to rounding, not to the bit)
    import io
    ta = pd.read_csv(io.StringIO(outs["packed"][0]), sep="\t")
    tb = pd.read_csv(io.StringIO(outs["decoded"][0]), sep="\t")
    assert list(ta.columns) == list(tb.columns) and len(ta) == len(tb) > 10
    for col in ta.columns:
        if ta[col].dtype.kind == "f":
            np.testing.assert_allclose(ta[col].values, tb[col].values, rtol=1e-9)
        else:
            assert (ta[col] == tb[col]).all(), col
    assert outs["packed"][3] == outs["decoded"][3] == len(ta)
    # the filter saw the whole file: its MAF differs from the analysed samples' MAF
    file_maf = np.nanmean(G, axis=1) / 2
    file_maf = np.minimum(file_maf, 1 - file_maf)
    want_skipped = sorted(set(names[:3]) | {nm for nm, m in zip(names, file_maf)
                                            if not m >= 0.05})
    assert outs["packed"][1] == want_skipped
    assert "rs5" in want_skipped and len(outs["packed"][2]) > 0


def test_stratified_cox_through_execute(tmp_path):
    """``outcome=dict(tte=..., event=..., strata=...)``: the strata column of y
    reaches PHReg(strata=...) in the reference (survival.py:54-56); here it
    reaches the engine, for a GWAS and for a single fit."""
    import oracle.gwas as og
    from genetest_b200.phenotypes import DataFramePhenotypes
    from genetest_b200.statistics.models.survival import StatsCoxPH
    rng = np.random.default_rng(31)
    n, n_snps = 900, 10
    ids = np.array(["s{:03d}".format(i) for i in range(n)])
    age = rng.normal(50, 10, n)
    t_event = rng.exponential(1.0 / np.exp(0.02 * (age - 50)))
    t_cens = rng.exponential(2.0, n)
    pheno = pd.DataFrame({
        "tte": np.ceil(np.minimum(t_event, t_cens) * 20) / 20,
        "event": (t_event <= t_cens).astype(float), "age": age,
        "center": rng.integers(0, 3, n).astype(float)}, index=ids)
    G = rng.binomial(2, 0.3, size=(n_snps, n)).astype(float)
    names = ["rs{}".format(j) for j in range(n_snps)]
    info = pd.DataFrame({"chrom": 1, "pos": np.arange(n_snps) + 1, "a1": "A", "a2": "G"},
                        index=names)
    spec._reset()
    modelspec = spec.ModelSpec(
        outcome=dict(tte=spec.phenotypes.tte, event=spec.phenotypes.event,
                     strata=spec.phenotypes.center),
        predictors=[spec.SNPs, spec.phenotypes.age], test="coxph")
    sub = subscribers.ResultsMemory()
    analysis.execute(DataFramePhenotypes(pheno),
                     DataFrameReader(pd.DataFrame(G.T, index=ids, columns=names), info),
                     modelspec, subscribers=[sub], output_prefix=str(tmp_path / "o"))
    got = sub._get_gwas_results()
    Y = pheno[["tte", "event", "center"]].values
    C = pheno[["age"]].values
    for j, nm in enumerate(names):
        st, o = og.gwas_one("coxph", Y, C, ["age"], G[j], coded="G", reference="A")
        assert st == "ok"
        for key in ("SNPs", "age"):
            for f in ("coef", "std_err", "p_value"):
                assert abs(got[nm][key][f] - o[key][f]) <= 1e-6 * abs(o[key][f]), (nm, key, f)
    # single fit: Stats*.fit(y, X) with a strata column in y
    y = pheno[["tte", "event"]].copy()
    y["strata"] = pheno["center"]
    X = pd.DataFrame({"age": age, "SNPs": G[0]}, index=ids)
    one = StatsCoxPH().fit(y, X)
    st, o = og.gwas_one("coxph", Y, C, ["age"], G[0])
    # (gwas_one flips to the minor allele; G[0] has MAF 0.3: no flip)
    assert abs(one["SNPs"]["coef"] - o["SNPs"]["coef"]) <= 1e-6 * abs(o["SNPs"]["coef"])
    assert abs(one["MODEL"]["log_likelihood"] - o["MODEL"]["log_likelihood"]) <= 1e-8 * abs(o["MODEL"]["log_likelihood"])


def test_mixedlm_two_step(tmp_path):
    """The "optimized" MixedLM analysis (cli.py:183-310): mixed model without the SNPs on the
    host, linear GWAS of the predicted random effects on the GPU, the real mixed model only
    for the markers under the p-value threshold.  Data of the reference's test_mixedlm.py;
    the markers kept by a threshold of 1 get the R / lme4 numbers of test_mixedlm_gwas, and
    the approximation file holds the OLS fit of the random effects on each marker."""
    from genetest_b200.phenotypes import DataFramePhenotypes
    df = load_table("mixedlm")
    df.index = df["sampleid"].values
    snps = ["snp1", "snp2", "snp3", "snp4"]
    markers = {"snp1": ("3", 1234, "T", "G"), "snp2": ("3", 2345, "C", "A"),
               "snp3": ("2", 3456, "G", "T"), "snp4": ("22", 4567, "A", "C")}
    rng = np.random.default_rng(4)
    pheno = df.drop(snps, axis=1).iloc[rng.permutation(df.shape[0])]
    geno = df.loc[~df.index.duplicated(), snps]
    geno = geno.loc[rng.permutation(geno.index.values)]
    info = pd.DataFrame({"chrom": [markers[s][0] for s in snps], "pos": [markers[s][1] for s in snps],
                         "a1": [markers[s][2] for s in snps], "a2": [markers[s][3] for s in snps]},
                        index=snps)
    formula = ("[outcome=pheno3, groups=sampleid] ~ SNPs + age + var1 + factor(gender) + "
               "factor(visit)")
    prefix = str(tmp_path / "two_step")
    spec._reset()
    kept = analysis.execute_mixedlm_two_step(
        DataFramePhenotypes(pheno, repeated_measurements=True), DataFrameReader(geno, info),
        formula, p_threshold=1.0, output_prefix=prefix)
    assert kept == set(snps)
    approx = pd.read_csv(prefix + ".mixedlm_approximation.txt", sep="\t").set_index("snp")
    assert list(approx.index) == snps and (approx.approximation == "Yes").all()
    assert (approx.n == 60).all()
    final = pd.read_csv(prefix + ".txt", sep="\t").set_index("snp")
    rec = EXPECTED["mixedlm"]["test_mixedlm_gwas"]
    for c in rec["checks"]:
        if c["entity"] != "SNPs" or c["field"] not in ("coef", "std_err", "z_value"):
            continue
        col = {"coef": "coef", "std_err": "se", "z_value": "z"}[c["field"]]
        assert abs(final.loc[c["snp"], col] - c["expected"]) < 10.0 ** -min(c["places"], 5), c
    # a strict threshold keeps only the strongest markers
    spec._reset()
    p_sorted = approx.p.sort_values()
    kept2 = analysis.execute_mixedlm_two_step(
        DataFramePhenotypes(pheno, repeated_measurements=True), DataFrameReader(geno, info),
        formula, p_threshold=float(p_sorted.iloc[1]) * 1.0001, output_prefix=prefix + "_b")
    assert kept2 == set(p_sorted.index[:2])


def test_phewas_matches_single_fits(tmp_path):
    """Phenome-wide scan (analysis.py:328-452): the same predictors against every phenotype
    that is not one of them; each result equals the single fit of that outcome."""
    df = load_table("linear")
    rng = np.random.default_rng(8)
    order = rng.permutation(df.index.values)
    genotypes = _dataframe_reader(df, order, ["snp1", "snp2"])
    cols = [c for c in df.columns if not c.startswith("snp")]

    def phenotypes():
        ph = _DummyPhenotypes()
        ph.data = df[cols].iloc[rng.permutation(df.shape[0])]
        return ph

    spec._reset()
    sub = subscribers.ResultsMemory()
    modelspec = spec.ModelSpec(outcome=spec.PheWAS(),
                               predictors=[spec.genotypes.snp1, spec.phenotypes.age],
                               test="linear")
    analysis.execute(phenotypes(), genotypes, modelspec, subscribers=[sub])
    outcomes = sorted(r["outcome"] for r in sub.results)
    assert outcomes == sorted(c for c in cols if c != "age") and len(outcomes) >= 3
    for r in sub.results:
        spec._reset()
        one = subscribers.ResultsMemory()
        single = spec.ModelSpec(outcome=spec.phenotypes[r["outcome"]],
                                predictors=[spec.genotypes.snp1, spec.phenotypes.age],
                                test="linear")
        analysis.execute(phenotypes(), genotypes, single, subscribers=[one])
        want = one.results[0]
        for col in ("snp1", "age", "intercept"):
            for fld in ("coef", "std_err", "p_value"):
                assert r[col][fld] == pytest.approx(want[col][fld], rel=1e-10), (r["outcome"], col, fld)
        assert r["MODEL"]["nobs"] == want["MODEL"]["nobs"]
    # GWAS-only options are refused as in the reference
    spec._reset()
    with pytest.raises(ValueError):
        analysis.execute(phenotypes(), genotypes,
                         spec.ModelSpec(outcome=spec.PheWAS(), predictors=[spec.phenotypes.age],
                                        test="linear"),
                         subscribers=[subscribers.ResultsMemory()],
                         variant_predicates=[lambda g: True])
