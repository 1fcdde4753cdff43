"""CPU tests of the host-side mirror of the reference API: formula grammar,
design-matrix construction, subscribers, PLINK I/O, sample alignment, the
C-ABI export table and the multi-rank gather (gloo, world_size 2)."""

import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from genetest_b200 import modelspec as spec
from genetest_b200 import sharding
from genetest_b200.genotypes import (DataFrameReader, PlinkReader,
                                     PlinkWriter, pack_rows, unpack_rows)
from genetest_b200.modelspec.predicates import MAFFilter, NameFilter
from genetest_b200.phenotypes import _DummyPhenotypes
from genetest_b200.statistics.descriptive import get_maf, get_sex_maf
from genetest_b200.subscribers import GWASWriter, ResultsMemory, RowWriter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _reset_spec():
    spec._reset()
    yield
    spec._reset()


@pytest.fixture()
def containers():
    rng = np.random.default_rng(3)
    n = 100
    data = pd.DataFrame({
        "pheno": rng.normal(size=n), "var1": rng.normal(2, 1, n) ** 2 + 1,
        "var2": rng.integers(0, 3, n).astype(float),
        "var3": rng.choice(["x0", "x1", "x2"], n),
        "snp": rng.integers(0, 3, n).astype(float),
    }, index=["s{}".format(i) for i in range(n)])
    data.loc["s5", "var1"] = np.nan
    ph = _DummyPhenotypes()
    ph.data = data.drop("snp", axis=1).iloc[rng.permutation(n)]
    info = pd.DataFrame({"chrom": ["1"], "pos": [10], "a1": ["B"],
                         "a2": ["A"]}, index=["snp"])
    geno = DataFrameReader(data[["snp"]].iloc[rng.permutation(n)], info)
    return data, ph, geno


# ---- grammar (modelspec.ebnf) ---------------------------------------------
def test_formula_simple_and_transformations(containers):
    data, ph, geno = containers
    model = spec.parse_formula(
        "pheno ~ g(snp) + var1 + pow(var1, 4) + log10(var1) + ln(var1)")
    assert model["conditions"] is None
    del model["conditions"]
    model["test"] = "linear"
    ms = spec.ModelSpec(**model)
    matrix = ms.create_data_matrix(ph, geno)
    assert matrix.shape == (100, 7)
    assert matrix.intercept.unique().tolist() == [1]
    m = matrix.loc[data.index]
    # the marker column is oriented on the minor allele
    g = data.snp.values
    if np.nanmean(g) / 2 > 0.5:
        g = 2 - g
    np.testing.assert_array_equal(m[spec.genotypes.snp.id].values, g)
    np.testing.assert_array_equal(
        m[spec.pow(spec.phenotypes.var1, 4).id].values, data.var1.values ** 4)
    np.testing.assert_array_equal(
        m[spec.log10(spec.phenotypes.var1).id].values,
        np.log10(data.var1.values))
    np.testing.assert_array_equal(
        m[spec.ln(spec.phenotypes.var1).id].values, np.log(data.var1.values))
    assert ms.get_translations()[spec.phenotypes.var1.id] == "var1"
    meta = ms.variant_metadata[spec.genotypes.snp.id]
    assert meta["name"] == "snp" and meta["chrom"] == "1"


def test_formula_factor_interaction_and_gwas(containers):
    data, ph, geno = containers
    model = spec.parse_formula(
        "pheno ~ SNPs + var1 + factor(var3) as f3 + var1*var2 as v12 + "
        "SNPs*factor(var3)")
    assert model["predictors"][0] == "SNPs"
    del model["conditions"]
    model["test"] = "linear"
    ms = spec.ModelSpec(**model)
    matrix = ms.create_data_matrix(ph, geno).loc[data.index]
    f3 = spec.factor(spec.phenotypes.var3, name="f3")
    assert f3.id == "TRANSFORM:f3"
    for level in ("x1", "x2"):
        np.testing.assert_array_equal(
            matrix["TRANSFORM:f3:level." + level].values,
            (data.var3 == level).astype(float).values)
    assert "TRANSFORM:f3:level.x0" not in matrix.columns   # reference level
    got = matrix["TRANSFORM:v12"].values
    np.testing.assert_array_equal(np.isnan(got), np.isnan(data.var1.values))
    np.testing.assert_allclose(got[~np.isnan(got)],
                               (data.var1 * data.var2).dropna().values)
    assert ms.has_gwas_interaction
    assert sorted(ms.gwas_interaction) == [
        "TRANSFORM:GWAS_INTER:level.x1", "TRANSFORM:GWAS_INTER:level.x2"]
    fac = "TRANSFORM:ENCODE_FACTOR:" + spec.phenotypes.var3.id
    assert ms.gwas_interaction["TRANSFORM:GWAS_INTER:level.x1"] == \
        (fac + ":level.x1",)
    # the GWAS interaction adds no column by itself
    assert not any(c.startswith("TRANSFORM:GWAS_INTER")
                   for c in matrix.columns)


def test_formula_conditions_and_labelled_outcomes():
    ms, subgroups = spec.modelspec_from_formula(
        "[tte=t, event=e] | sex = 1, g(rs1) = \"AT\", grp ~ SNPs + x",
        "coxph", None)
    assert set(ms.outcome) == {"tte", "event"}
    assert ms.outcome["tte"] is spec.phenotypes.t
    assert subgroups == [1, "AT", None]
    assert ms.stratify_by == [spec.phenotypes.sex, spec.genotypes.rs1,
                              spec.phenotypes.grp]
    with pytest.raises(RuntimeError):
        spec.parse_formula("y ~ ")
    with pytest.raises(RuntimeError):
        spec.parse_formula("y ~ x +")
    with pytest.raises(ValueError):
        spec.ModelSpec(outcome=spec.phenotypes.y, predictors=[], test="nope")
    # names may contain ':' and '_' and "as" is only a keyword on its own
    model = spec.parse_formula("y_1 ~ assay + a:b")
    assert model["predictors"] == [spec.phenotypes.assay,
                                   spec.phenotypes["a:b"]]


def test_modelspec_sample_intersection_and_missing_marker(containers):
    data, ph, geno = containers
    geno.df = geno.df.iloc[:60]
    ms = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                        predictors=[spec.genotypes.snp, spec.phenotypes.var2],
                        test="linear")
    matrix = ms.create_data_matrix(ph, geno)
    assert sorted(matrix.index) == sorted(geno.df.index)
    spec._reset()
    ms = spec.ModelSpec(outcome=spec.phenotypes.pheno,
                        predictors=[spec.genotypes.unknown], test="linear")
    with pytest.raises(ValueError):
        ms.create_data_matrix(ph, geno)


# ---- descriptive ------------------------------------------------------------
def test_maf_helpers():
    g = np.array([0, 0, 0, 1, 0, 0, 1, 2, 0, 0], dtype=float)
    assert get_maf(g, "B", "A") == (0.2, "B", "A", False)
    maf, minor, major, flip = get_maf(2 - g, "B", "A")
    assert (minor, major, flip) == ("A", "B", True) and abs(maf - 0.2) < 1e-15
    g[[1, 5]] = np.nan
    assert get_maf(g, "B", "A")[0] == 4 / 16
    assert np.isnan(get_maf(np.full(3, np.nan), "B", "A")[0])
    sex = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0], dtype=float)
    g = np.array([0, 0, 0, 2, 0, 0, 1, 2, 0, 0], dtype=float)
    assert abs(get_sex_maf(g, sex, "B", "A")[0] - 3 / 15) < 1e-15
    assert np.isnan(get_sex_maf(np.full(3, np.nan), np.ones(3), "B", "A")[0])


# ---- PLINK I/O -----------------------------------------------------------------
def test_plink_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    n, m = 37, 9                       # 37 samples: last byte is padded
    G = rng.integers(-1, 3, (m, n)).astype(float)
    G[G < 0] = np.nan
    packed = pack_rows(G)
    assert packed.shape == (m, 10)
    np.testing.assert_array_equal(unpack_rows(packed, n), G)
    # PLINK coding: 00 hom A1 (=2), 01 missing, 10 het, 11 hom A2 (=0)
    assert pack_rows(np.array([[2, np.nan, 1, 0]]))[0, 0] == 0b11100100
    prefix = str(tmp_path / "t")
    samples = ["id{}".format(i) for i in range(n)]
    with PlinkWriter(prefix) as w:
        for row in G:
            w.write_genotypes(np.where(np.isnan(row), -1, row))
        w.write_bim([(1, "rs{}".format(j), 100 + j, "T", "C")
                     for j in range(m)])
        w.write_fam(samples)
    assert open(prefix + ".bed", "rb").read(3) == bytes([0x6C, 0x1B, 0x01])
    with PlinkReader(prefix) as r:
        assert r.get_samples() == samples
        assert r.get_number_variants() == m and r.get_number_samples() == n
        rows = list(r.iter_genotypes())
        np.testing.assert_array_equal(np.array([x.genotypes for x in rows]), G)
        assert rows[3].variant.name == "rs3" and rows[3].variant.pos == 103
        assert (rows[3].coded, rows[3].reference) == ("T", "C")
        one = r.get_variant_by_name("rs7")
        assert len(one) == 1
        np.testing.assert_array_equal(one[0].genotypes, G[7])
        assert r.get_variant_by_name("nope") == []
        np.testing.assert_array_equal(np.asarray(r.packed()), packed)
        assert MAFFilter(0.0)(rows[0]) and NameFilter({"rs1"})(rows[1])
        assert not NameFilter({"rs1"})(rows[2])


# ---- sample alignment ------------------------------------------------------------
def test_generate_sample_order():
    from genetest_b200.analysis import _generate_sample_order
    x = pd.Index(["c", "a", "z", "b"])
    geno = ["q", "b", "a", "r", "c"]
    order, labels = _generate_sample_order(x, geno)
    assert order.tolist() == [1, 2, 4]
    assert labels.tolist() == ["b", "a", "c"]
    order, labels = _generate_sample_order(pd.Index(["x"]), geno)
    assert order.shape[0] == 0


# ---- subscribers ---------------------------------------------------------------------
def _fake_result(name, coef):
    stats = {"coef": coef, "std_err": 0.5, "lower_ci": coef - 1,
             "upper_ci": coef + 1, "t_value": 2 * coef, "p_value": 0.01}
    snp = dict(stats, chrom="3", pos=12, major="C", minor="T", name=name,
               maf=0.25)
    return {"MODEL": {"nobs": 60.0, "log_likelihood": -10.5,
                      "r_squared_adj": 0.3}, "SNPs": snp,
            "INT:level.2": dict(stats)}


def test_gwas_writer_columns(tmp_path):
    path = str(tmp_path / "out.txt")
    w = GWASWriter(path, "linear")
    w.init(None)
    w.handle(_fake_result("rs1", 1.5))
    w.handle(_fake_result("rs2", -0.25))
    w.close()
    lines = open(path).read().splitlines()
    assert lines[0].split("\t") == [
        "snp", "chr", "pos", "major", "minor", "maf", "n", "ll", "adj_r2",
        "coef", "se", "lower", "upper", "t", "p"]
    assert lines[1].split("\t") == [
        "rs1", "3", "12", "C", "T", "0.25", "60.0", "-10.5", "0.3", "1.5",
        "0.5", "0.5", "2.5", "3.0", "0.01"]
    # interaction + subgroup headers are rewritten in place
    path = str(tmp_path / "out2.txt")
    w = GWASWriter(path, "linear")
    w.init(None)
    w._update_gwas_interaction(["INT:level.2"])
    w._update_current_subset({"sex": 1})
    w.handle(_fake_result("rs1", 1.0))
    w.close()
    header = open(path).read().splitlines()[0].split("\t")
    assert header[:9] == ["snp", "chr", "pos", "major", "minor", "maf", "n",
                          "ll", "subgroup"]
    assert header[9:] == ["adj_r2"] + ["INT:level.2:" + s for s in
                                       ("coef", "se", "lower", "upper", "t",
                                        "p")]
    row = open(path).read().splitlines()[1].split("\t")
    assert row[8] == "sex:1" and row[10] == "1.0"
    for test, col in (("logistic", "n_events"), ("coxph", "n_events")):
        w = GWASWriter(str(tmp_path / (test + ".txt")), test)
        assert col in [c[0] for c in w.columns]
        w.close()
    cox = GWASWriter(str(tmp_path / "c.txt"), "coxph")
    assert [c[0] for c in cox.columns][9:] == [
        "coef", "se", "hr", "hr_lower", "hr_upper", "z", "p"]
    cox.close()


def test_results_memory_and_row_writer(tmp_path, capsys):
    class _Spec(object):
        def get_translations(self):
            return {"uuid-1": "age"}
    mem = ResultsMemory()
    mem.init(_Spec())
    res = _fake_result("rs9", 2.0)
    res["uuid-1"] = {"coef": 1.0}
    mem.handle(res)
    assert "age" in mem.results[0] and "uuid-1" not in mem.results[0]
    assert list(mem._get_gwas_results()) == ["rs9"]
    rw = RowWriter(columns=[("name", spec.result["SNPs"]["name"]),
                            ("tag", "const")], header=True, sep=",")
    rw.init(None)
    rw.handle(res)
    assert capsys.readouterr().out.splitlines() == ["name,tag", "rs9,const"]
    # factor levels are gathered per level
    fac = spec.phenotypes.var3
    got = spec.result[fac]["coef"]  # noqa: F841 (path building only)
    per_level = spec.core.Result(fac).get(
        {fac.id + ":level.a": 1, fac.id + ":level.b": 2, "other": 3})
    assert per_level == {"level.a": 1, "level.b": 2}


# ---- C ABI ---------------------------------------------------------------------------------
def test_shared_library_exports_every_declared_symbol():
    """The library loads on a CPU box and exports what include/*.h declares
    (no compute call is made here)."""
    header = open(os.path.join(ROOT, "include", "genetest_b200.h")).read()
    declared = set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    from genetest_b200 import engine
    assert declared == set(engine.ABI_SYMBOLS)
    lib = ctypes.CDLL(engine.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), name
    lib.gt_abi_version.restype = ctypes.c_int
    assert lib.gt_abi_version() == 2
    lib.gt_packed_pitch.restype = ctypes.c_int64
    assert lib.gt_packed_pitch(100000) == 25088


def test_engine_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genetest_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


# ---- sharding (gloo, world_size 2) -------------------------------------------------------
def test_shard_range_covers_everything():
    for n in (0, 1, 7, 62500, 500000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
    assert sharding.shard_range(500000, 8, 3) == (187500, 250000)


def _gather_worker(rank, world, port, tmp):
    import torch.distributed as dist
    from genetest_b200 import sharding as sh
    from genetest_b200.engine import Results
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sh.rank_world() == (rank, world)
        n = 500
        lo, hi = sh.shard_range(n, world, rank)
        if rank == 1:
            hi -= 3                       # ragged shard
        idx = np.arange(lo, hi)
        out = np.vstack([idx * 1.5, -idx.astype(float)])
        iout = np.vstack([idx, idx % 7]).astype(np.int32)
        res = Results(out, iout, 1)
        meta = [("rs{}".format(i), "1", int(i), "A", "B") for i in idx]
        got, got_meta = sh.gather_results(res, meta)
        if rank == 0:
            want = np.concatenate([np.arange(0, 250), np.arange(250, 497)])
            np.testing.assert_array_equal(got.iout[0], want)
            np.testing.assert_array_equal(got.out[0], want * 1.5)
            assert [m[2] for m in got_meta] == want.tolist()
            open(os.path.join(tmp, "ok"), "w").write("ok")
        else:
            assert got is None and got_meta is None
    finally:
        dist.destroy_process_group()


def test_gather_results_two_ranks_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2,
             join=True)
    assert os.path.exists(str(tmp_path / "ok"))


def _round_worker(rank, world, port, tmp):
    """The packed path of a multi-rank analysis without a GPU: the batch plan, the
    dealing of batches to ranks and the per-round gather (gloo), with a fake
    engine that derives every number from the marker's row index."""
    import torch.distributed as dist
    from genetest_b200 import analysis, sharding as sh
    from genetest_b200.engine import Results
    from genetest_b200.modelspec.predicates import MAFFilter, NameFilter
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world),
                      RANK=str(rank), LOCAL_RANK=str(rank))
    assert sh.ensure_initialized()
    try:
        assert sh.rank_world() == (rank, world)
        n_markers = 1000
        table = [("rs{}".format(i), "1", i, "A", "G") for i in range(n_markers)]

        class Reader(object):
            def variant_table(self):
                return table

            def packed(self):
                return np.arange(n_markers, dtype=np.int64)[:, None].view(np.uint8)

        class Eng(object):
            n_fields, n_params = 10, 1

            def run_packed_host(self, rows, file_maf=None):
                idx = np.ascontiguousarray(rows).view(np.int64)[:, 0]
                if file_maf is not None:
                    file_maf[:] = (idx % 10) / 20.0
                out = np.outer(np.arange(1, 11), idx).astype(float)
                iout = np.vstack([idx % 3 == 0, idx, idx % 2, idx * 0]).astype(np.int32)
                return Results(out, iout, 1)

        old = analysis.BATCH_SNPS
        analysis.BATCH_SNPS = 64
        try:
            messages = {"skipped": [], "failed": []}
            plan = analysis._PackedPlan(
                Reader(), [NameFilter({t[0] for t in table[5:]}), MAFFilter(0.1)],
                messages, rank, world)
            plan.batches = [b[:40] for b in plan.batches]      # small batches, several rounds
            got = []
            n_rounds = -(-len(plan.batches) // world)
            for i in range(n_rounds):
                j = i * world + rank
                mine = plan.run(Eng(), j) if j < len(plan.batches) else None
                sizes = [len(plan.batches[i * world + r]) if i * world + r < len(plan.batches)
                         else 0 for r in range(world)]
                blocks = sh.gather_round(mine, sizes, 10, 1)
                if rank == 0:
                    for r, blk in enumerate(blocks):
                        if blk is not None:
                            got.append(plan.finish(i * world + r, blk))
                else:
                    assert blocks is None
        finally:
            analysis.BATCH_SNPS = old
        if rank == 0:
            names = [m[0] for _res, meta in got for m in meta]
            want = [t[0] for b in plan.batches for t in (table[i] for i in b)
                    if (t[2] % 10) / 20.0 >= 0.1]
            assert names == want                           # marker order, MAF filter applied
            nobs = np.concatenate([res.nobs for res, _m in got])
            assert nobs.tolist() == [int(nm[2:]) for nm in names]
            out0 = np.concatenate([res.out[9] for res, _m in got])
            np.testing.assert_array_equal(out0, 10.0 * nobs)
            assert set(messages["skipped"]) >= {"rs0", "rs4", "rs10", "rs11"}
            open(os.path.join(tmp, "ok2"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_multi_rank_packed_rounds_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_round_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(str(tmp_path / "ok2"))


# ---- streaming result sink ---------------------------------------------------------------
def test_native_formatter_matches_python_repr():
    from genetest_b200.engine import format_rows
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.normal(size=20000) * 10.0 ** rng.integers(-12, 12, 20000),
        [0.1, 1e-5, 1e16, 1e-4, 9.999e-5, 123456789.123, 5e-324, 1e22, np.nan,
         np.inf, -np.inf, -0.0, 0.0, 60.0, 1234567890123456.0,
         12345678901234567.0, 1.7976931348623157e308, 2.5, 1e15, 1e17]])
    got = format_rows([x]).decode().split("\n")[:-1]
    assert got == [repr(v) for v in x.tolist()]
    keep = np.arange(x.shape[0]) % 3 != 0
    names = ["rs{}".format(i) for i in range(x.shape[0])]
    out = format_rows([names, "chr1", np.arange(x.shape[0]), x], keep=keep,
                      sep=",").decode().splitlines()
    assert len(out) == keep.sum()
    assert out[0] == "rs1,chr1,1,{!r}".format(float(x[1]))


@pytest.mark.parametrize("test", ["linear", "logistic", "coxph"])
def test_block_sink_writes_the_same_bytes_as_the_row_sink(test, tmp_path):
    """GWASWriter.handle_block (vectorised, native formatter) vs the
    reference protocol handle(dict) once per marker."""
    from genetest_b200.analysis import ResultBlock
    from genetest_b200.engine import Results
    rng = np.random.default_rng(1)
    n, cols = 64, ["age", "intercept", "SNPs", "I:x"]
    if test == "coxph":
        cols.remove("intercept")
    nf = 4 + 6 * len(cols)
    out = rng.normal(size=(nf, n)) * 10.0 ** rng.integers(-5, 3, (nf, n))
    out[0] = rng.random(n) / 2
    iout = np.zeros((4, n), dtype=np.int32)
    iout[0, ::7] = 3                      # failed markers are not written
    iout[1] = rng.integers(50, 60, n)
    iout[2] = rng.integers(0, 2, n)
    meta = [("rs{}".format(i), i % 22 + 1, 1000 + i, "A", "G")
            for i in range(n)]
    block = ResultBlock(test, cols, Results(out, iout, len(cols)), meta, None)
    paths = []
    for name in ("a.txt", "b.txt"):
        w = GWASWriter(str(tmp_path / name), test)
        w.init(None)
        w._update_gwas_interaction(["I:x"])
        w._update_current_subset({"sex": 1})
        if name == "a.txt":
            w.handle_block(block)
        else:
            for i in range(n):
                if block.ok(i):
                    w.handle(block.to_dict(i))
        w.close()
        paths.append(str(tmp_path / name))
    a, b = open(paths[0]).read(), open(paths[1]).read()
    assert a == b and len(a.splitlines()) == 1 + int((iout[0] == 0).sum())
    with pytest.raises(KeyError):
        block.column(["SNPs", "no_such_field"])


def test_native_formatter_thread_slices_keep_row_order():
    """More than 4,096 rows are formatted by several host threads, each into
    its own slice of the output buffer: the text must not depend on that."""
    from genetest_b200.engine import format_rows
    rng = np.random.default_rng(3)
    for n in (4095, 4096, 40001):
        x = rng.normal(size=n) * 10.0 ** rng.integers(-30, 30, n)
        k = rng.integers(-5, 1000, n)
        names = ["rs{}".format(i) * (1 + i % 3) for i in range(n)]
        tail = ["A" * (i % 5) for i in range(n)]          # empty strings too
        keep = rng.random(n) < 0.6
        got = format_rows([names, "c", x, k, tail], keep=keep) \
            .decode().split("\n")[:-1]
        want = ["{}\tc\t{!r}\t{}\t{}".format(names[i], float(x[i]), k[i],
                                            tail[i])
                for i in range(n) if keep[i]]
        assert got == want
        assert format_rows([x, names]).decode().split("\n")[:-1] == \
            ["{!r}\t{}".format(float(x[i]), names[i]) for i in range(n)]


def test_plink_reader_bim_variants(tmp_path):
    """.bim files separated by tabs or runs of spaces, non-numeric
    chromosomes, duplicated marker names."""
    prefix = str(tmp_path / "v")
    G = np.array([[0, 1, 2, -1, 0], [2, 2, 1, 0, -1], [1, 0, 0, 2, 2]])
    with PlinkWriter(prefix) as w:
        for row in G:
            w.write_genotypes(row)
        w.write_fam(["a", "b", "c", "d", "e"])
    with open(prefix + ".bim", "w") as f:
        f.write("X\trs1\t0\t1000\tA\tG\n")
        f.write("MT   rs2  0.5   2000 AT  G\n")
        f.write("23\trs1\t0\t3000\tC\tT\n")
    with PlinkReader(prefix) as r:
        assert r.variant_table() == [("rs1", "X", 1000, "A", "G"),
                                     ("rs2", "MT", 2000, "AT", "G"),
                                     ("rs1", "23", 3000, "C", "T")]
        both = r.get_variant_by_name("rs1")
        assert [v.variant.pos for v in both] == [1000, 3000]
        assert r.get_variant_by_name("nope") == []
        np.testing.assert_array_equal(
            np.nan_to_num(both[1].genotypes, nan=-1), G[2])
        assert r.packed().shape == (3, 2)


def test_impute2_reader(tmp_path):
    """Dosage of a2, probability threshold, gz files, lookups by name."""
    import gzip
    from genetest_b200.genotypes import Impute2Reader, parsers
    assert parsers["impute2"] is Impute2Reader
    sample = tmp_path / "x.sample"
    sample.write_text("ID_1 ID_2 missing\n0 0 0\nf1 s1 0\nf2 s2 0\nf3 s3 0\n")
    lines = ["1 rs1 100 A G 1 0 0 0 1 0 0 0 1",
             "1 rs2 200 C T 0.95 0.05 0 0.5 0.4 0.1 0 0 0",
             "X rs3 300 G GA 0.02 0.08 0.9 0 0.1 0.9 0.05 0.9 0.05"]
    gen = tmp_path / "x.impute2.gz"
    with gzip.open(gen, "wt") as f:
        f.write("\n".join(lines) + "\n")
    with Impute2Reader(str(gen), str(sample)) as r:
        assert r.get_samples() == ["s1", "s2", "s3"]
        assert r.get_number_variants() == 3
        got = list(r.iter_genotypes())
        np.testing.assert_array_equal(got[0].genotypes, [0.0, 1.0, 2.0])
        assert (got[0].coded, got[0].reference) == ("G", "A")
        np.testing.assert_allclose(got[1].genotypes, [0.05, np.nan, np.nan])
        np.testing.assert_allclose(got[2].genotypes, [1.88, 1.9, 1.0])
        assert got[2].variant.chrom == "X" and got[2].variant.pos == 300
        assert [v.name for v in r.iter_variants()] == ["rs1", "rs2", "rs3"]
        one = r.get_variant_by_name("rs2")
        assert len(one) == 1 and one[0].variant.name == "rs2"
        assert r.get_variant_by_name("nope") == []
    loose = Impute2Reader(str(gen), str(sample), probability_threshold=0.0)
    np.testing.assert_allclose(list(loose.iter_genotypes())[1].genotypes,
                               [0.05, 0.6, np.nan])


def test_p_value_formulas_match_scipy():
    """The Student-t / normal tails and the t quantile the kernels use
    (gt_common.cuh, compiled for the host behind gt_host_*): relative error
    against scipy over the ranges a GWAS meets, tiny p-values included."""
    import scipy.stats as st
    from genetest_b200.engine import load_library
    lib = load_library()
    rng = np.random.default_rng(0)
    dfs = [1, 2, 3, 5, 8, 10, 11, 20, 31, 57, 58, 100, 500, 1000, 5000, 19988,
           49988, 99988, 500000, 1e6]
    ts = np.concatenate([
        [0, 1e-12, 1e-6, 1e-3, 0.01, 0.1, 0.5, 1, 1.5, 2, 2.5, 2.9, 3, 3.1, 4,
         5, 7, 10, 15, 20, 30, 38, 50, 100, 300, 1000],
        rng.uniform(0, 6, 200), 10 ** rng.uniform(-3, 2.5, 200)])
    worst = 0.0
    for df in dfs:
        want = 2 * st.t.sf(ts, df)
        for t, w in zip(ts, want):
            got = lib.gt_host_t_sf2(float(t), float(df))
            assert lib.gt_host_t_sf2(float(-t), float(df)) == got
            if w < 1e-300:
                assert got < 1e-299
                continue
            worst = max(worst, abs(got - w) / w)
    assert worst < 1e-9, worst
    # Large degrees of freedom (biobank sizes): the asymptotic expansion the
    # kernels switch to at df >= 500 against 40-digit references (scipy's own
    # tail loses digits to cancellation there), df up to 1e7.
    import mpmath
    mpmath.mp.dps = 40
    worst_big = 0.0
    for df in (499, 500, 501, 750, 1000, 3000, 19988, 99988, 499988, 1e6, 1e7):
        for t in (1e-9, 1e-4, 0.01, 0.3, 1, 1.9, 2.5, 3, 3.5, 4.5, 6, 9, 14, 22, 30, 36):
            x = mpmath.mpf(df) / (mpmath.mpf(df) + mpmath.mpf(t) ** 2)
            w = float(mpmath.betainc(mpmath.mpf(df) / 2, mpmath.mpf(1) / 2, 0, x,
                                     regularized=True))
            got = lib.gt_host_t_sf2(float(t), float(df))
            worst_big = max(worst_big, abs(got - w) / w)
    assert worst_big < 5e-13, worst_big
    assert np.isnan(lib.gt_host_t_sf2(float("nan"), 10.0))
    assert lib.gt_host_t_sf2(float("inf"), 10.0) == 0.0
    for df in dfs + [4, 6, 7, 9, 15, 40]:
        w = st.t.ppf(0.975, df)
        assert abs(lib.gt_host_t_ppf975(float(df)) - w) <= 1e-12 * w
    for z in np.concatenate([[0, 1e-9, 0.5, 1, 2, 5, 10, 20, 37],
                             rng.uniform(0, 37, 200)]):
        w = 2 * st.norm.sf(z)
        assert abs(lib.gt_host_norm_sf2(float(z)) - w) <= 1e-12 * w
        assert lib.gt_host_norm_sf2(float(-z)) == lib.gt_host_norm_sf2(float(z))


def test_logistic_thread_exp_log_formulas():
    """exp(-a) and log(1 + t) of the thread-per-SNP logistic kernel (host build of the same
    source) against numpy: the deviance stop rule compares sums of these to 1e-8."""
    from genetest_b200 import engine
    lib = engine.load_library()
    rng = np.random.default_rng(11)
    a = np.concatenate([[0.0, 1e-300, 1e-17, 1e-9, 0.3, 0.34657, 0.6931471805599453, 1.0, 36.0,
                         37.5, 699.9, 700.0, 800.0], rng.uniform(0, 40, 4000),
                        rng.exponential(1.0, 4000)])
    got = np.array([lib.gt_host_lt_exp_neg(float(v)) for v in a])
    want = np.exp(-np.minimum(a, 700.0))
    assert np.max(np.abs(got - want) / want) < 6e-16
    t = np.concatenate([[0.0, 1e-300, 1e-20, 1e-9, 0.4142135623730950, 0.4142135623730951,
                         0.41421356237309515, 0.5, 1.0], rng.uniform(0, 1, 4000),
                        np.exp(-rng.uniform(0, 40, 4000))])
    got = np.array([lib.gt_host_lt_log1p01(float(v)) for v in t])
    want = np.log1p(t)
    # log(1 + t) through d = 1 + t: absolute accuracy of one ulp of d (the deviance needs no more)
    assert np.max(np.abs(got - want)) < 2.3e-16
    big = t > 1e-3
    assert np.max(np.abs(got[big] - want[big]) / want[big]) < 3e-13


def test_properties_pack_roundtrip_and_float_repr():
    """Property tests (hypothesis): 2-bit packing round-trips for any shape,
    and the native formatter reproduces ``repr(float)`` for any double."""
    from hypothesis import given, settings, strategies as hs
    from genetest_b200.engine import format_rows

    @settings(max_examples=60, deadline=None)
    @given(hs.integers(1, 9), hs.integers(1, 70), hs.integers(0, 2 ** 32 - 1))
    def roundtrip(m, n, seed):
        rng = np.random.default_rng(seed)
        G = rng.integers(-1, 3, (m, n)).astype(float)
        G[G < 0] = np.nan
        packed = pack_rows(G)
        assert packed.shape == (m, (n + 3) // 4) and packed.dtype == np.uint8
        np.testing.assert_array_equal(unpack_rows(packed, n), G)

    @settings(max_examples=40, deadline=None)
    @given(hs.lists(hs.floats(allow_nan=True, allow_infinity=True, width=64),
                    min_size=1, max_size=200))
    def float_repr(values):
        x = np.array(values, dtype=np.float64)
        got = format_rows([x]).decode().split("\n")[:-1]
        assert got == [repr(float(v)) for v in x]

    roundtrip()
    float_repr()


def _write_bgen(path, samples, variants, bits=8, compress=True, layout=2,
                embed_samples=True):
    """Minimal BGEN writer for the tests (unphased diploid bi-allelic).
    ``variants``: [(rsid, chrom, pos, a1, a2, probs[n, 3], missing[n])]."""
    import struct
    import zlib

    def lstr(s, fmt):
        b = s.encode()
        return struct.pack(fmt, len(b)) + b

    n = len(samples)
    flags = (1 if compress else 0) | (layout << 2) | ((1 << 31) if embed_samples else 0)
    header = struct.pack("<III", 20, len(variants), n) + b"bgen" + struct.pack("<I", flags)
    sample_block = b""
    if embed_samples:
        ids = b"".join(lstr(s, "<H") for s in samples)
        sample_block = struct.pack("<II", 8 + len(ids), n) + ids
    body = b""
    for rsid, chrom, pos, a1, a2, probs, missing in variants:
        v = b""
        if layout == 1:
            v += struct.pack("<I", n)
        v += lstr("id_" + rsid, "<H") + lstr(rsid, "<H") + lstr(chrom, "<H")
        v += struct.pack("<I", pos)
        if layout == 2:
            v += struct.pack("<H", 2)
        v += lstr(a1, "<I") + lstr(a2, "<I")
        if layout == 1:
            q = np.where(missing[:, None], 0.0, probs)
            data = np.round(q * 32768).astype("<u2").tobytes()
            v += (struct.pack("<I", len(zlib.compress(data))) + zlib.compress(data)
                  if compress else data)
        else:
            scale = 2 ** bits - 1
            ints = np.round(probs[:, :2] * scale).astype(np.uint64)
            if bits in (8, 16, 32):
                payload = ints.astype({8: "<u1", 16: "<u2", 32: "<u4"}[bits]).tobytes()
            else:
                flat = ((ints.reshape(-1, 1) >> np.arange(bits, dtype=np.uint64)) & 1) \
                    .astype(np.uint8).ravel()
                payload = np.packbits(flat, bitorder="little").tobytes()
            ploidy = (np.full(n, 2, dtype=np.uint8) | (missing.astype(np.uint8) << 7))
            data = struct.pack("<IHBB", n, 2, 2, 2) + ploidy.tobytes() + \
                struct.pack("<BB", 0, bits) + payload
            if compress:
                z = zlib.compress(data)
                v += struct.pack("<II", len(z) + 4, len(data)) + z
            else:
                v += struct.pack("<I", len(data)) + data
        body += v
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(header) + len(sample_block)))
        f.write(header + sample_block + body)


@pytest.mark.parametrize("layout, bits, compress", [
    (2, 8, True), (2, 16, True), (2, 8, False), (2, 11, True), (1, 16, True),
    (1, 16, False)])
def test_bgen_reader(tmp_path, layout, bits, compress):
    from genetest_b200.genotypes import BgenReader, parsers
    assert parsers["bgen"] is BgenReader
    rng = np.random.default_rng(layout * 100 + bits)
    n, m = 37, 5
    samples = ["s{}".format(i) for i in range(n)]
    variants = []
    for j in range(m):
        g = rng.integers(0, 3, n)
        probs = np.full((n, 3), 0.02)
        probs[np.arange(n), g] = 0.96
        blur = rng.random(n) < 0.2
        probs[blur] = rng.dirichlet([1, 1, 1], int(blur.sum()))
        missing = rng.random(n) < 0.1
        variants.append(("rs{}".format(j), str(1 + j % 2), 1000 + j, "A", "GT",
                         probs, missing))
    path = str(tmp_path / "t.bgen")
    _write_bgen(path, samples, variants, bits=bits, compress=compress,
                layout=layout, embed_samples=(layout == 2))
    sample_file = None
    if layout == 1:
        sample_file = str(tmp_path / "t.sample")
        with open(sample_file, "w") as f:
            f.write("ID_1 ID_2 missing\n0 0 0\n")
            f.writelines("f {} 0\n".format(s) for s in samples)
    scale = 32768.0 if layout == 1 else float(2 ** bits - 1)
    with BgenReader(path, sample_file) as r:
        assert r.get_samples() == samples and r.get_number_variants() == m
        got = list(r.iter_genotypes())
        assert [g.variant.name for g in got] == [v[0] for v in variants]
        assert [v.name for v in r.iter_variants()] == [v[0] for v in variants]
        for g, (rsid, chrom, pos, a1, a2, probs, missing) in zip(got, variants):
            assert (g.variant.chrom, g.variant.pos) == (chrom, pos)
            assert (g.reference, g.coded) == ("A", "GT")
            if layout == 1:
                q = np.round(probs * scale) / scale
            else:
                q2 = np.round(probs[:, :2] * scale) / scale
                q = np.column_stack((q2, 1 - q2.sum(axis=1)))
            want = q[:, 1] + 2 * q[:, 2]
            want[missing | (q.max(axis=1) < 0.9)] = np.nan
            np.testing.assert_allclose(g.genotypes, want, rtol=0, atol=1e-12)
        one = r.get_variant_by_name("rs3")
        assert len(one) == 1 and one[0].variant.pos == 1003
        np.testing.assert_array_equal(np.isnan(one[0].genotypes),
                                      np.isnan(got[3].genotypes))
        # iteration restarts cleanly after a lookup
        assert len(list(r.iter_genotypes())) == m


def test_plink_reader_text_tables(tmp_path):
    """.fam / .bim parsing: any whitespace, blank lines; a line without six fields is an error."""
    from genetest_b200.genotypes import PlinkReader, PlinkWriter
    p = str(tmp_path / "c")
    with PlinkWriter(p) as w:
        w.write_genotypes(np.array([0, 1, 2, -1, 1]))
        w.write_genotypes(np.array([2, 2, 0, 0, 1]))
        w.write_bim([(1, "rs1", 10, "A", "G"), (2, "rs2", 20, "C", "T")])
        w.write_fam(["a", "b", "c", "d", "e"])
    want = [("rs1", "1", 10, "A", "G"), ("rs2", "2", 20, "C", "T")]
    assert PlinkReader(p).variant_table() == want
    with open(p + ".bim", "w") as f:
        f.write("1  rs1\t0 10 A G\n\n2\trs2   0\t20\tC\tT\n   \n")
    r = PlinkReader(p)
    assert r.variant_table() == want and r.get_samples() == ["a", "b", "c", "d", "e"]
    with open(p + ".bim", "w") as f:
        f.write("1 rs1 0 10 A G\n2 rs2 0 20 C\n")
    with pytest.raises(ValueError):
        PlinkReader(p)
