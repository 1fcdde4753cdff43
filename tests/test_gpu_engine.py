"""GPU parity tests of the C-ABI engine against the CPU oracle.

Everything here calls ``libgenetest_b200.so`` through ctypes (the drop-in
boundary) and compares with ``oracle/`` on the same seeded inputs.
Tolerances: linear 1e-8, logistic 1e-6 (BASELINE.json north_star);
indices / nobs / flip / MAF / status bit-exact.
"""

import numpy as np
import pytest

from genetest_b200.genotypes import pack_rows, unpack_rows

from gpu_support import (compare, run_oracle, synth_design, synth_genotypes)

pytestmark = pytest.mark.gpu

LIN_TOL = 1e-8
LOGIT_TOL = 1e-6


@pytest.fixture(scope="module")
def engine():
    from genetest_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def _linear_y(rng, C, G=None):
    y = 0.1 * C[:, 0] + rng.normal(size=C.shape[0]) + 3.0
    if G is not None:
        g = np.nan_to_num(G[0])
        y = y + 0.05 * g
    return y


def test_linear_config1_packed(engine):
    """BASELINE config 1: 1,000 SNPs x 1,000 samples, age + sex."""
    rng = np.random.default_rng(20261019)
    n, n_snps = 1000, 1000
    C, names = synth_design(rng, n)
    G = synth_genotypes(rng, n_snps, n)
    y = _linear_y(rng, C, G)
    engine.set_design("linear", y, C)
    res = engine.run_packed_host(pack_rows(G))
    oracle = run_oracle("linear", y[:, None], C, names, G)
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert n_ok == n_snps
    print("linear config1 worst rel err", worst)


def test_linear_missing_subset_permuted(engine):
    """1 % missing, samples a permuted subset of the genotype file, edge
    markers: monomorphic, all-missing, MAF > 0.5 (flip), MAF threshold."""
    rng = np.random.default_rng(7)
    n_file, n, n_snps = 1003, 901, 300
    C, names = synth_design(rng, n, k_extra=3)
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.02, maf_lo=0.02,
                            maf_hi=0.98)
    Gfile[0] = 0.0                      # monomorphic -> cond inf
    Gfile[1] = np.nan                   # all missing
    Gfile[2] = 2.0                      # monomorphic after flip
    Gfile[3] = 1.0                      # all heterozygous: collinear with 1
    Gfile[4, pos] = np.nan              # missing for every design sample
    Gfile[5] = 0.0
    Gfile[5, pos[:3]] = 1.0             # MAF below threshold
    y = _linear_y(rng, C)
    G = Gfile[:, pos]
    engine.set_design("linear", y, C, sample_pos=pos, n_file=n_file)
    res = engine.run_packed_host(pack_rows(Gfile))
    oracle = run_oracle("linear", y[:, None], C, names, G)
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert res.status[0] == 3 and np.isinf(res.aux[0])
    assert res.status[1] == 1 and res.status[4] == 1
    assert res.status[2] == 3 and np.isinf(res.aux[2])
    assert res.status[3] == 3
    assert n_ok > 250
    print("linear missing/subset worst rel err", worst)
    # same batch with a MAF threshold (analysis.py:151-153)
    engine.set_design("linear", y, C, sample_pos=pos, n_file=n_file,
                      maf_t=0.01)
    res = engine.run_packed_host(pack_rows(Gfile))
    oracle = run_oracle("linear", y[:, None], C, names, G, maf_t=0.01)
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert res.status[5] == 2 and res.status[0] == 2


def test_linear_device_resident_matches_host_path(engine):
    import torch
    rng = np.random.default_rng(11)
    n, n_snps = 777, 130
    C, names = synth_design(rng, n, k_extra=8)
    G = synth_genotypes(rng, n_snps, n, missing=0.01)
    y = _linear_y(rng, C)
    engine.set_design("linear", y, C)
    host = engine.run_packed_host(pack_rows(G))
    pitch = engine.packed_pitch()
    buf = np.zeros((n_snps, pitch), dtype=np.uint8)
    packed = pack_rows(G)
    buf[:, :packed.shape[1]] = packed
    dev = torch.from_numpy(buf).cuda()
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    np.testing.assert_array_equal(res.iout, host.iout)
    np.testing.assert_array_equal(np.nan_to_num(res.out),
                                  np.nan_to_num(host.out))
    oracle = run_oracle("linear", y[:, None], C, names, G)
    problems, worst, _ = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]


@pytest.mark.parametrize("n_inter", [1, 2])
def test_linear_interaction(engine, n_inter):
    """SNP x covariate interaction columns (analysis.py:159-166)."""
    rng = np.random.default_rng(100 + n_inter)
    n, n_snps = 1500, 120
    C, names = synth_design(rng, n, k_extra=2)
    G = synth_genotypes(rng, n_snps, n, missing=0.01, maf_lo=0.05,
                        maf_hi=0.95)
    y = _linear_y(rng, C)
    inter = {"I:age": ["age"]}
    W = [C[:, 0]]
    if n_inter == 2:
        inter["I:sexc0"] = ["sex", "c0"]
        W.append(C[:, 1] * C[:, 2])
    engine.set_design("linear", y, C, W=np.column_stack(W),
                      condition_value_t=1e6)
    res = engine.run_packed_host(pack_rows(G))
    oracle = run_oracle("linear", y[:, None], C, names, G, interaction=inter,
                        model_kwargs={"condition_value_t": 1e6})
    problems, worst, n_ok = compare("linear", res, oracle, names,
                                    inter_keys=list(inter), rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert n_ok == n_snps
    print("linear interaction", n_inter, "worst rel err", worst)


def test_linear_dosage(engine):
    """Real-valued dosages with NaN (impute2 / bgen / dataframe readers)."""
    rng = np.random.default_rng(5)
    n, n_snps = 400, 64
    C, names = synth_design(rng, n, k_extra=1)
    G = np.clip(synth_genotypes(rng, n_snps, n) +
                rng.normal(0, 0.05, (n_snps, n)), 0, 2)
    G[rng.random(G.shape) < 0.03] = np.nan
    y = _linear_y(rng, C)
    engine.set_design("linear", y, C)
    res = engine.run_dosage_host(G)
    oracle = run_oracle("linear", y[:, None], C, names, G)
    # dosage MAF is a float sum: compare to 1e-14, not bit-exact
    for j, (st, o) in enumerate(oracle):
        assert st == "ok"
        assert abs(res.maf[j] - o["SNPs"]["maf"]) < 1e-14
        o["SNPs"]["maf"] = res.maf[j]
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]


def test_linear_condition_gate_uses_jacobi(engine):
    """Raw age gives cond ~ 700.  A gate just above the largest condition
    number cannot be decided by the trace bound -> Jacobi eigenvalues, every
    marker passes; a gate just below the smallest makes every marker fail
    with the reference's message and value (linear.py:50-53)."""
    from oracle import sm_port
    rng = np.random.default_rng(3)
    n, n_snps = 600, 40
    C, names = synth_design(rng, n, standardise=False)
    G = synth_genotypes(rng, n_snps, n, missing=0.01)
    y = _linear_y(rng, C)
    conds = []
    for j in range(n_snps):
        ok = ~np.isnan(G[j])
        conds.append(sm_port.ols_pinv(
            y[ok], np.column_stack((C[ok], G[j, ok])))["condition_number"])
    conds = np.array(conds)
    hi = float(conds.max() * 1.01)
    engine.set_design("linear", y, C, condition_value_t=hi)
    res = engine.run_packed_host(pack_rows(G))
    oracle = run_oracle("linear", y[:, None], C, names, G,
                        model_kwargs={"condition_value_t": hi})
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert n_ok == n_snps and res.niter.min() == 1    # Jacobi path taken
    lo = float(conds.min() * 0.99)
    engine.set_design("linear", y, C, condition_value_t=lo)
    res = engine.run_packed_host(pack_rows(G))
    assert (res.status == 3).all()
    np.testing.assert_allclose(res.aux, conds, rtol=1e-9)


def test_logistic_packed(engine):
    rng = np.random.default_rng(21)
    n, n_snps = 2000, 200
    C, names = synth_design(rng, n, k_extra=4)
    G = synth_genotypes(rng, n_snps, n, missing=0.01, maf_lo=0.05,
                        maf_hi=0.95)
    eta = -0.5 + 0.2 * C[:, 0] + 0.3 * np.nan_to_num(G[0])
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    engine.set_design("logistic", y, C)
    res = engine.run_packed_host(pack_rows(G))
    oracle = run_oracle("logistic", y[:, None], C, names, G)
    problems, worst, n_ok = compare("logistic", res, oracle, names,
                                    rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    assert n_ok == n_snps
    print("logistic worst rel err", worst)


def test_logistic_thread_kernel_matches_oracle_and_warp_kernel(engine):
    """The thread-per-SNP IRLS kernel (binary outcome, no interaction): every marker against
    the oracle, and statuses / iteration counts / values against the warp-per-SNP kernel.
    Edge markers: perfect separation (deferred to the warp kernel), monomorphic, all
    missing, MAF above one half (flip), MAF under the threshold; samples are a permuted
    subset of the file."""
    rng = np.random.default_rng(77)
    n_file, n, n_snps = 1500, 1300, 500
    C, names = synth_design(rng, n, k_extra=3)
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.03, maf_lo=0.02, maf_hi=0.98)
    eta = -0.3 + 0.4 * C[:, 0] + 0.25 * np.nan_to_num(Gfile[7, pos])
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    Gfile[0, pos] = 2.0 * y            # separates the outcome
    Gfile[1] = 0.0                     # monomorphic
    Gfile[2] = np.nan                  # all missing
    Gfile[3] = (rng.random(n_file) < 0.004) * 1.0   # MAF below the threshold
    engine.set_design("logistic", y, C, sample_pos=pos, n_file=n_file, maf_t=0.01)
    packed = pack_rows(Gfile)
    try:
        engine.set_option("thread_iter_min", 64)
        res = engine.run_packed_host(packed)
        engine.set_option("thread_iter", 0)
        ref = engine.run_packed_host(packed)
    finally:
        engine.set_option("thread_iter", 1)
        engine.set_option("thread_iter_min", 4096)
    G = Gfile[:, pos]
    oracle = run_oracle("logistic", y[:, None], C, names, G, maf_t=0.01)
    problems, worst, n_ok = compare("logistic", res, oracle, names, rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    assert res.status[0] == 6 and res.status[2] == 1 and res.status[3] == 2
    assert n_ok >= n_snps - 5
    assert (res.status == ref.status).all() and (res.niter == ref.niter).all()
    assert (res.nobs == ref.nobs).all() and (res.flip == ref.flip).all()
    assert res.flip.sum() > 100
    ok = res.status == 0
    np.testing.assert_allclose(res.out[:, ok], ref.out[:, ok], rtol=1e-9, atol=0, equal_nan=True)
    print("logistic thread kernel worst rel err", worst, "ok", n_ok)


def test_logistic_thread_kernel_interaction(engine):
    """Thread-per-SNP IRLS with one SNP x covariate interaction column (design row
    [q | g | g w], analysis.py:159-166): against the oracle and the warp-per-SNP kernel, on a
    permuted subset of the file with missing genotypes and flipped markers."""
    rng = np.random.default_rng(79)
    n_file, n, n_snps = 1300, 1100, 240
    C, names = synth_design(rng, n, k_extra=2)
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.02, maf_lo=0.05, maf_hi=0.95)
    eta = -0.2 + 0.3 * C[:, 0] + 0.2 * np.nan_to_num(Gfile[5, pos]) * C[:, 0]
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    inter = {"I:age": ["age"]}
    engine.set_design("logistic", y, C, W=C[:, [0]], sample_pos=pos, n_file=n_file)
    assert engine.dominant_kernel() == "logistic_thread_kernel"
    packed = pack_rows(Gfile)
    try:
        engine.set_option("thread_iter_min", 64)
        res = engine.run_packed_host(packed)
        engine.set_option("thread_iter", 0)
        ref = engine.run_packed_host(packed)
    finally:
        engine.set_option("thread_iter", 1)
        engine.set_option("thread_iter_min", 4096)
    oracle = run_oracle("logistic", y[:, None], C, names, Gfile[:, pos], interaction=inter)
    problems, worst, n_ok = compare("logistic", res, oracle, names, inter_keys=list(inter),
                                    rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    assert n_ok == n_snps
    assert (res.status == ref.status).all() and (res.niter == ref.niter).all()
    assert (res.flip == ref.flip).all() and res.flip.sum() > 60
    np.testing.assert_allclose(res.out, ref.out, rtol=1e-9, atol=0, equal_nan=True)
    print("logistic thread kernel with interaction: worst rel err", worst)


def test_logistic_interaction_subset_and_separation(engine):
    rng = np.random.default_rng(22)
    n_file, n, n_snps = 1200, 1000, 60
    C, names = synth_design(rng, n, k_extra=1)
    pos = np.sort(rng.permutation(n_file)[:n]).astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.02)
    eta = -0.3 + 0.4 * C[:, 0]
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    # marker 0 separates the outcome perfectly; marker 1 is monomorphic
    Gfile[0, pos] = 2.0 * y
    Gfile[1] = 0.0
    inter = {"I:age": ["age"]}
    engine.set_design("logistic", y, C, W=C[:, [0]], sample_pos=pos,
                      n_file=n_file)
    res = engine.run_packed_host(pack_rows(Gfile))
    G = Gfile[:, pos]
    oracle = run_oracle("logistic", y[:, None], C, names, G, interaction=inter)
    problems, worst, n_ok = compare("logistic", res, oracle, names,
                                    inter_keys=list(inter), rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    assert res.status[0] == 6
    print("logistic interaction worst rel err", worst, "ok", n_ok)


def test_synthetic_generator_roundtrip(engine):
    """The device-side .bed generator: decoded rows feed the oracle."""
    rng = np.random.default_rng(9)
    n, n_snps = 515, 96
    C, names = synth_design(rng, n, k_extra=2)
    y = _linear_y(rng, C)
    engine.set_design("linear", y, C)
    dev = engine.synth_packed(n, n_snps, seed=123, missing_rate=0.01)
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    G = unpack_rows(dev.cpu().numpy()[:, :(n + 3) // 4], n)
    assert 0.002 < np.isnan(G).mean() < 0.03
    oracle = run_oracle("linear", y[:, None], C, names, G)
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]


def _cox_data(rng, n, C, ties=False):
    """tte ~ Exp(rate = exp(0.1 z(age))), censoring Exp(0.5)."""
    t_event = rng.exponential(1.0 / np.exp(0.1 * C[:, 0]))
    t_cens = rng.exponential(2.0, n)
    tte = np.minimum(t_event, t_cens)
    event = (t_event <= t_cens).astype(float)
    if ties:
        tte = np.ceil(tte * 120) / 120        # heavy ties (Efron correction)
    return tte, event


@pytest.mark.parametrize("ties", [False, True, "heavy"])
def test_coxph_packed(engine, ties):
    rng = np.random.default_rng(31 + bool(ties) + 2 * (ties == "heavy"))
    n, n_snps = 1500, 120
    C, names = synth_design(rng, n, k_extra=3)
    C, names = C[:, :-1], names[:-1]          # StatsCoxPH drops the intercept
    G = synth_genotypes(rng, n_snps, n, missing=0.01, maf_lo=0.05,
                        maf_hi=0.95)
    G[0] = 0.0                                # monomorphic -> Singular matrix
    tte, event = _cox_data(rng, n, C, bool(ties))
    if ties == "heavy":
        # a few distinct times with hundreds of tied events each (Efron groups
        # spanning several 32-slot blocks) and administrative censoring
        tte = np.ceil(tte * 3) / 3
        late = tte > 1.0
        tte[late], event[late] = 1.0, 0.0
        _, counts = np.unique(tte[event == 1], return_counts=True)
        assert counts.max() > 200 and late.sum() > 100
    elif ties:
        _, counts = np.unique(tte, return_counts=True)
        assert counts.max() > 3
    engine.set_design("coxph", tte, C, event=event)
    res = engine.run_packed_host(pack_rows(G))
    oracle = run_oracle("coxph", np.column_stack((tte, event)), C, names, G)
    problems, worst, n_ok = compare("coxph", res, oracle, names,
                                    rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    assert res.status[0] == 7 and n_ok == n_snps - 1
    print("coxph ties={} worst rel err {}".format(ties, worst))


@pytest.mark.parametrize("mode", ["untied", "tied", "heavy", "strata"])
def test_coxph_thread_kernel_matches_oracle_and_warp_kernel(engine, mode):
    """The thread-per-SNP Newton kernel: every marker against the oracle, and statuses / step
    counts / values against the warp-per-SNP kernel.  untied: censored samples share some
    event times; tied: Efron groups of a few events; heavy: a few time points with hundreds of
    tied events each and administrative censoring; strata: seven strata of uneven sizes (one
    without events) on top of tied times.  Edge markers: monomorphic (singular), all missing,
    MAF above one half (flip), MAF under the threshold; the samples are a permuted subset of
    the file."""
    rng = np.random.default_rng(78 + len(mode))
    n_file, n, n_snps = 1400, 1200, 300
    C, names = synth_design(rng, n, k_extra=3)
    C, names = C[:, :-1], names[:-1]
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.03, maf_lo=0.02, maf_hi=0.98)
    Gfile[1] = 0.0
    Gfile[2] = np.nan
    Gfile[3] = (rng.random(n_file) < 0.014) * 1.0      # MAF ~ 0.007: under the threshold, ~17 carriers
    tte, event = _cox_data(rng, n, C)
    strata = None
    if mode == "untied":
        # censored samples at the time of an event (they belong to its risk set)
        ev_idx = np.flatnonzero(event == 1)[:40]
        ce_idx = np.flatnonzero(event == 0)[:40]
        tte[ce_idx] = tte[ev_idx]
        assert len(np.unique(tte[event == 1])) == int(event.sum())
    elif mode == "heavy":
        tte = np.ceil(tte * 3) / 3
        late = tte > 1.0
        tte[late], event[late] = 1.0, 0.0
        _, counts = np.unique(tte[event == 1], return_counts=True)
        assert counts.max() > 150
    else:
        tte = np.ceil(tte * 60) / 60
        _, counts = np.unique(tte[event == 1], return_counts=True)
        assert counts.max() > 3
        if mode == "strata":
            strata = rng.choice(7, n, p=np.arange(1, 8) / 28.0)
            event[strata == 0] = 0.0
    packed = pack_rows(Gfile)
    G = Gfile[:, pos]
    Y = np.column_stack((tte, event) if strata is None else (tte, event, strata.astype(float)))
    for maf_t in (None, 0.01):
        engine.set_design("coxph", tte, C, event=event, sample_pos=pos, n_file=n_file, maf_t=maf_t,
                          strata=None if strata is None else strata * 5 + 2)
        assert engine.dominant_kernel() == "cox_thread_kernel"
        try:
            engine.set_option("thread_iter_min", 64)
            res = engine.run_packed_host(packed)
            engine.set_option("thread_iter", 0)
            ref = engine.run_packed_host(packed)
        finally:
            engine.set_option("thread_iter", 1)
            engine.set_option("thread_iter_min", 4096)
        oracle = run_oracle("coxph", Y, C, names, G, maf_t=maf_t)
        problems, worst, n_ok = compare("coxph", res, oracle, names, rtol=LOGIT_TOL)
        assert not problems, problems[:10]
        if maf_t is None:
            assert res.status[1] == 7 and res.status[2] == 1
        else:
            assert res.status[1] == 2 and res.status[2] == 1 and res.status[3] == 2
        assert n_ok >= n_snps - 8
        assert (res.status == ref.status).all() and (res.niter == ref.niter).all()
        assert (res.nobs == ref.nobs).all() and (res.flip == ref.flip).all()
        assert res.flip.sum() > 60
        ok = res.status == 0
        np.testing.assert_allclose(res.out[:, ok], ref.out[:, ok], rtol=1e-7, atol=0, equal_nan=True)
        print("coxph thread kernel,", mode, "maf_t", maf_t, "worst rel err", worst, "ok", n_ok)


def test_coxph_interaction_subset_dosage(engine):
    rng = np.random.default_rng(41)
    n_file, n, n_snps = 900, 800, 40
    C, names = synth_design(rng, n, k_extra=1)
    C, names = C[:, :-1], names[:-1]
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.02)
    tte, event = _cox_data(rng, n, C)
    inter = {"I:age": ["age"]}
    engine.set_design("coxph", tte, C, event=event, W=C[:, [0]],
                      sample_pos=pos, n_file=n_file)
    res = engine.run_packed_host(pack_rows(Gfile))
    G = Gfile[:, pos]
    oracle = run_oracle("coxph", np.column_stack((tte, event)), C, names, G,
                        interaction=inter)
    problems, worst, n_ok = compare("coxph", res, oracle, names,
                                    inter_keys=list(inter), rtol=LOGIT_TOL)
    assert not problems, problems[:10]
    # real-valued dosages, aligned to the design rows
    D = np.clip(G + rng.normal(0, 0.03, G.shape), 0, 2)
    engine.set_design("coxph", tte, C, event=event, W=C[:, [0]])
    res = engine.run_dosage_host(D)
    oracle = run_oracle("coxph", np.column_stack((tte, event)), C, names, D,
                        interaction=inter)
    for j, (st, o) in enumerate(oracle):
        if st == "ok":
            assert abs(res.maf[j] - o["SNPs"]["maf"]) < 1e-14
            o["SNPs"]["maf"] = res.maf[j]
    problems, worst, n_ok = compare("coxph", res, oracle, names,
                                    inter_keys=list(inter), rtol=LOGIT_TOL)
    assert not problems, problems[:10]


def test_linear_config2_shape_sampled_parity(engine):
    """BASELINE config 2 shape: 100,000 samples, 10 covariates + intercept,
    1 % missing.  4,096 markers generated on the device; a stratified sample
    (lowest MAF, most missing, plus a random draw) is checked against the
    oracle, and size-independent properties on all of them: nobs + missing =
    n, MAF in (0, 0.5], recoding A1 <-> A2 gives the same fit with the flip
    flag inverted."""
    rng = np.random.default_rng(20261019)
    n, n_snps = 100000, 4096
    C, names = synth_design(rng, n, k_extra=8)
    y = 0.1 * C[:, 0] + rng.normal(size=n) + 10.0
    pos = rng.permutation(n).astype(np.int32)
    engine.set_design("linear", y, C, sample_pos=pos, n_file=n)
    dev = engine.synth_packed(n, n_snps, seed=7, missing_rate=0.01)
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    assert (res.status == 0).all()
    packed = dev.cpu().numpy()[:, :(n + 3) // 4]
    G = unpack_rows(packed, n)
    np.testing.assert_array_equal(res.nobs, n - np.isnan(G).sum(axis=1))
    assert ((res.maf > 0) & (res.maf <= 0.5)).all()
    pick = np.unique(np.concatenate([
        np.argsort(res.maf)[:8], np.argsort(res.nobs)[:8],
        rng.choice(n_snps, 16, replace=False)]))
    sub = type(res)(res.out[:, pick], res.iout[:, pick], res.n_params)
    oracle = run_oracle("linear", y[:, None], C, names, G[pick][:, pos])
    problems, worst, n_ok = compare("linear", sub, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    print("config-2 shape: {} markers vs oracle, worst rel err {:.2e}".format(
        n_ok, worst))
    # recoding A1 <-> A2 (00 <-> 11): the engine flips it back -> same fit
    swapped = packed.copy()
    hom_a1 = np.zeros_like(swapped)
    for s in (0, 2, 4, 6):
        code = (swapped >> s) & 3
        hom_a1 |= (np.where(code == 0, 3, np.where(code == 3, 0, code))
                   << s).astype(np.uint8)
    res2 = engine.run_packed_host(hom_a1[:256])
    k = len(names)
    np.testing.assert_array_equal(res2.flip, 1 - res.flip[:256])
    np.testing.assert_allclose(res2.maf, res.maf[:256], rtol=1e-14)
    np.testing.assert_allclose(res2.param(k)[0], res.param(k)[0, :256],
                               rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(res2.param(k)[1], res.param(k)[1, :256],
                               rtol=1e-10)


@pytest.mark.parametrize("shape", [(1000, 1000, 3, 0.0), (1003, 901, 5, 0.02),
                                   (5000, 5000, 11, 0.01)])
def test_linear_tcgen05_contraction_matches_fp64_and_oracle(engine, shape):
    """The int8 tensor-core contraction (tcgen05.mma kind::i8, exact fixed
    point) against the FP64 kernel and the oracle."""
    n_file, n, k_extra, miss = shape
    rng = np.random.default_rng(n_file + k_extra)
    n_snps = 300
    C, names = synth_design(rng, n, k_extra=k_extra)
    pos = np.sort(rng.permutation(n_file)[:n]).astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=miss, maf_lo=0.02,
                            maf_hi=0.98)
    Gfile[0] = 0.0
    Gfile[1] = 2.0
    y = _linear_y(rng, C) + 50.0
    packed = pack_rows(Gfile)
    engine.set_design("linear", y, C, sample_pos=pos, n_file=n_file)
    try:
        engine.set_option("contract_tc", 0)
        ref = engine.run_packed_host(packed)
        engine.set_option("contract_tc", 1)
        res = engine.run_packed_host(packed)
    finally:
        engine.set_option("contract_tc", 0)
    np.testing.assert_array_equal(res.iout, ref.iout)
    ok = res.status == 0
    np.testing.assert_array_equal(res.maf[ok], ref.maf[ok])
    k = len(names)
    for j in range(k + 1):
        a, b = res.param(j)[:, ok], ref.param(j)[:, ok]
        scale = np.maximum(np.abs(b[0]), np.abs(b[1]))
        assert np.max(np.abs(a[0] - b[0]) / scale) < 1e-10
        assert np.max(np.abs(a[1] - b[1]) / b[1]) < 1e-10
    oracle = run_oracle("linear", y[:, None], C, names, Gfile[:, pos])
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    print("tcgen05 contraction worst rel err vs oracle", worst)


@pytest.mark.parametrize("model", ["linear", "logistic", "coxph"])
def test_host_pipeline_chunks_are_bit_identical(engine, model):
    """The *_host entry points pipeline copy-in / compute / copy-back in
    chunks; any chunking must give the bits of one whole-batch call."""
    import torch
    rng = np.random.default_rng(77)
    n, n_snps = 1203, 2777
    C, names = synth_design(rng, n, k_extra=4)
    G = synth_genotypes(rng, n_snps, n, missing=0.02)
    event = None
    if model == "linear":
        y = _linear_y(rng, C)
    elif model == "logistic":
        y = (rng.random(n) < 0.4).astype(float)
    else:
        C = C[:, :-1]
        y, event = _cox_data(rng, n, C, False)
    engine.set_design(model, y, C, event=event)
    packed = pack_rows(G)
    pitch = engine.packed_pitch()
    buf = np.zeros((n_snps, pitch), dtype=np.uint8)
    buf[:, :packed.shape[1]] = packed
    out, iout = engine.run_packed_dev(torch.from_numpy(buf).cuda(), n_snps)
    ref = engine.fetch(out, iout)
    pinned = torch.from_numpy(packed).pin_memory().numpy()
    try:
        for chunk, src in ((0, packed), (500, packed), (1024, pinned),
                           (2777, packed), (1, packed[:40])):
            engine.set_option("host_chunk", chunk)
            got = engine.run_packed_host(src)
            m = src.shape[0]
            np.testing.assert_array_equal(got.iout, ref.iout[:, :m])
            np.testing.assert_array_equal(got.out.view(np.int64),
                                          ref.out[:, :m].view(np.int64))
        engine.set_option("host_chunk", 700)
        dos = G[:1500].copy()
        a = engine.run_dosage_host(dos)
        engine.set_option("host_chunk", 0)
        b = engine.run_dosage_host(dos)
        np.testing.assert_array_equal(a.iout, b.iout)
        np.testing.assert_array_equal(a.out.view(np.int64),
                                      b.out.view(np.int64))
    finally:
        engine.set_option("host_chunk", 0)


@pytest.mark.parametrize("config", ["3-logistic", "4-linear-gxe", "5-coxph"])
def test_baseline_config_shapes_sampled_parity(engine, config):
    """BASELINE configs 3-5 at their full sample sizes: markers
    generated on the device, a stratified sample (lowest MAF, fewest
    observations, random draw) checked against the oracle; nobs is checked
    on every marker."""
    rng = np.random.default_rng(20261019)
    inter, W, event, kwargs = None, None, None, {}
    if config == "3-logistic":
        model, n, n_snps, miss, n_pick, tol = "logistic", 50000, 1024, 0.0, 4, LOGIT_TOL
        C, names = synth_design(rng, n, k_extra=8)
        y = (rng.random(n) < 1 / (1 + np.exp(0.5 - 0.2 * C[:, 0]))).astype(float)
        Y = y[:, None]
    elif config == "4-linear-gxe":
        model, n, n_snps, miss, n_pick, tol = "linear", 100000, 2048, 0.01, 6, LIN_TOL
        C, names = synth_design(rng, n, k_extra=8)
        y = 0.1 * C[:, 0] + rng.normal(size=n)
        Y = y[:, None]
        inter, W = {"I:age": ["age"]}, C[:, [0]]
    else:
        model, n, n_snps, miss, n_pick, tol = "coxph", 20000, 512, 0.0, 3, LOGIT_TOL
        C, names = synth_design(rng, n, k_extra=8)
        C, names = C[:, :-1], names[:-1]
        y, event = _cox_data(rng, n, C)
        Y = np.column_stack((y, event))
    pos = rng.permutation(n).astype(np.int32)
    engine.set_design(model, y, C, W=W, event=event, sample_pos=pos, n_file=n)
    dev = engine.synth_packed(n, n_snps, seed=11, missing_rate=miss)
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    assert (res.status == 0).all()
    G = unpack_rows(dev.cpu().numpy()[:, :(n + 3) // 4], n)
    np.testing.assert_array_equal(res.nobs, n - np.isnan(G).sum(axis=1))
    pick = np.unique(np.concatenate([
        np.argsort(res.maf)[:n_pick], np.argsort(res.nobs)[:n_pick],
        rng.choice(n_snps, n_pick, replace=False)]))
    sub = type(res)(res.out[:, pick], res.iout[:, pick], res.n_params)
    oracle = run_oracle(model, Y, C, names, G[pick][:, pos], interaction=inter)
    problems, worst, n_ok = compare(model, sub, oracle, names,
                                    inter_keys=list(inter or ()), rtol=tol)
    assert not problems, problems[:10]
    assert n_ok == len(pick)
    print("config {}: {} markers vs oracle, worst rel err {:.2e}".format(
        config, n_ok, worst))


@pytest.mark.parametrize("model", ["linear", "logistic", "coxph"])
def test_sex_aware_maf(engine, model):
    """sexual_chromosome=True: MAF / flip follow get_sex_maf
    (descriptive.py:51-93) -- bit-exact for .bed rows, including unknown
    sexes, a subset of permuted samples, the MAF threshold and a marker whose
    genotyped samples all have an unknown sex (MAF = NaN, no flip)."""
    rng = np.random.default_rng(91)
    n_file, n, n_snps = 1100, 1000, 96
    C, names = synth_design(rng, n, k_extra=2)
    event = None
    if model == "linear":
        y = _linear_y(rng, C)
        Y = y[:, None]
    elif model == "logistic":
        y = (rng.random(n) < 0.45).astype(float)
        Y = y[:, None]
    else:
        C, names = C[:, :-1], names[:-1]
        y, event = _cox_data(rng, n, C)
        Y = np.column_stack((y, event))
    sex = rng.integers(0, 2, n).astype(float)
    sex[rng.random(n) < 0.05] = np.nan
    pos = rng.permutation(n_file)[:n].astype(np.int32)
    Gfile = synth_genotypes(rng, n_snps, n_file, missing=0.02, maf_lo=0.02,
                            maf_hi=0.98)
    Gfile[0, pos[~np.isnan(sex)]] = np.nan       # only unknown sexes genotyped
    males = pos[sex == 1]
    Gfile[1:40][:, males] = np.where(Gfile[1:40][:, males] == 1, 2,
                                     Gfile[1:40][:, males])   # hemizygous males
    engine.set_design(model, y, C, event=event, sample_pos=pos, n_file=n_file,
                      maf_t=0.05)
    engine.set_sample_sex(sex)
    res = engine.run_packed_host(pack_rows(Gfile))
    G = Gfile[:, pos]
    oracle = run_oracle(model, Y, C, names, G, maf_t=0.05, sexes=sex)
    tol = LIN_TOL if model == "linear" else LOGIT_TOL
    problems, worst, n_ok = compare(model, res, oracle, names, rtol=tol)
    assert not problems, problems[:10]
    assert np.isnan(res.maf[0]) and res.flip[0] == 0
    assert (res.status == 2).sum() >= 1 and res.flip.sum() > 5
    # autosomal MAF differs on most markers: the override really ran
    engine.set_sample_sex(None)
    plain = engine.run_packed_host(pack_rows(Gfile))
    assert (plain.maf[1:] != res.maf[1:]).mean() > 0.5
    # real-valued dosages, aligned to the design rows
    D = np.clip(G[:32] + rng.normal(0, 0.03, G[:32].shape), 0, 2)
    engine.set_design(model, y, C, event=event)
    engine.set_sample_sex(sex)
    res = engine.run_dosage_host(D)
    oracle = run_oracle(model, Y, C, names, D, sexes=sex)
    for j, (st, o) in enumerate(oracle):
        if st == "ok" and not np.isnan(o["SNPs"]["maf"]):
            assert abs(res.maf[j] - o["SNPs"]["maf"]) < 1e-14
            o["SNPs"]["maf"] = res.maf[j]
    problems, worst, n_ok = compare(model, res, oracle, names, rtol=tol)
    assert not problems, problems[:10]
    engine.set_sample_sex(None)


def test_dosage_rows_need_an_identity_alignment(engine):
    from genetest_b200.engine import EngineError
    rng = np.random.default_rng(3)
    C, names = synth_design(rng, 50)
    engine.set_design("linear", _linear_y(rng, C), C,
                      sample_pos=np.arange(50, dtype=np.int32)[::-1].copy(),
                      n_file=50)
    with pytest.raises(EngineError, match="aligned to the design rows"):
        engine.run_dosage_host(np.zeros((2, 50)))


def test_linear_config2_full_size_properties(engine):
    """BASELINE config 2 at its full size: 500,000 markers x 100,000 samples
    (12.5 GB of 2-bit genotypes in HBM, seven batches of the contraction
    kernel, the last one partial).  Size-independent properties on every
    marker, bit-identical results when a sub-range is run on its own (batch
    offsets), and oracle parity on markers drawn from the batch edges."""
    import torch
    rng = np.random.default_rng(20261019)
    n, n_snps = 100000, 500000
    C, names = synth_design(rng, n, k_extra=8)
    y = 0.1 * C[:, 0] + rng.normal(size=n)
    pos = rng.permutation(n).astype(np.int32)
    engine.set_design("linear", y, C, sample_pos=pos, n_file=n)
    dev = engine.synth_packed(n, n_snps, seed=3, missing_rate=0.01)
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    assert (res.status == 0).all()
    assert ((res.maf > 0) & (res.maf <= 0.5)).all()
    assert (res.nobs <= n).all() and (res.nobs > 0.985 * n).all()
    k = len(names)
    coef, se, lo, hi, t, p = res.param(k)
    assert np.isfinite(np.delete(res.out, 1, axis=0)).all()      # row 1 = aux, NaN when ok
    assert (se > 0).all() and (lo < coef).all() and (coef < hi).all()
    np.testing.assert_allclose(t, coef / se, rtol=1e-12)
    assert ((p > 0) & (p <= 1)).all()
    # null markers: p-values are uniform (Kolmogorov distance, 500k draws)
    ps = np.sort(p)
    assert np.abs(ps - (np.arange(n_snps) + 0.5) / n_snps).max() < 0.01
    # a sub-range that straddles batch edges, run on its own: same bits
    batch = 148 * 2 * 128 * 2
    a, b = 3 * batch - 700, 3 * batch + 900
    out2, iout2 = engine.run_packed_dev(dev[a:b], b - a)
    sub = engine.fetch(out2, iout2)
    np.testing.assert_array_equal(sub.iout, res.iout[:, a:b])
    np.testing.assert_array_equal(sub.out.view(np.int64),
                                  np.ascontiguousarray(res.out[:, a:b]).view(np.int64))
    # oracle parity at the batch edges and at both ends
    pick = np.array([0, batch - 1, batch, 3 * batch + 1, 6 * batch - 1,
                     6 * batch, n_snps - 1])
    packed = dev[torch.from_numpy(pick).cuda()].cpu().numpy()[:, :(n + 3) // 4]
    G = unpack_rows(packed, n)
    picked = type(res)(res.out[:, pick], res.iout[:, pick], res.n_params)
    oracle = run_oracle("linear", y[:, None], C, names, G[:, pos])
    problems, worst, n_ok = compare("linear", picked, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]
    assert n_ok == len(pick)
    del dev, out, iout
    torch.cuda.empty_cache()


@pytest.mark.parametrize("config", ["3-logistic", "4-linear-gxe", "5-coxph"])
def test_baseline_configs_full_size_properties(engine, config):
    """BASELINE configs 3-5 at their full marker counts: every marker gets a
    finite result, a sub-range run on its own reproduces the same bits, and
    markers from both ends agree with the oracle."""
    import torch
    rng = np.random.default_rng(20261020)
    inter, W, event = None, None, None
    if config == "3-logistic":
        model, n, n_snps, miss, tol = "logistic", 50000, 500000, 0.0, LOGIT_TOL
        C, names = synth_design(rng, n, k_extra=8)
        y = (rng.random(n) < 1 / (1 + np.exp(0.5 - 0.2 * C[:, 0]))).astype(float)
        Y = y[:, None]
    elif config == "4-linear-gxe":
        model, n, n_snps, miss, tol = "linear", 100000, 200000, 0.01, LIN_TOL
        C, names = synth_design(rng, n, k_extra=8)
        y = 0.1 * C[:, 0] + rng.normal(size=n)
        Y = y[:, None]
        inter, W = {"I:age": ["age"]}, C[:, [0]]
    else:
        model, n, n_snps, miss, tol = "coxph", 20000, 100000, 0.0, LOGIT_TOL
        C, names = synth_design(rng, n, k_extra=8)
        C, names = C[:, :-1], names[:-1]
        y, event = _cox_data(rng, n, C)
        Y = np.column_stack((y, event))
    pos = rng.permutation(n).astype(np.int32)
    engine.set_design(model, y, C, W=W, event=event, sample_pos=pos, n_file=n)
    dev = engine.synth_packed(n, n_snps, seed=5, missing_rate=miss)
    out, iout = engine.run_packed_dev(dev, n_snps)
    res = engine.fetch(out, iout)
    assert (res.status == 0).all()
    assert np.isfinite(np.delete(res.out, 1, axis=0)).all()
    assert ((res.maf > 0) & (res.maf <= 0.5)).all()
    n_cols = len(names) + 1 + (1 if inter else 0)
    for j in range(n_cols):
        coef, se, lo, hi, t, p = res.param(j)
        assert (se > 0).all() and (lo < coef).all() and (coef < hi).all()
        assert ((p >= 0) & (p <= 1)).all()
    p_snp = np.sort(res.param(len(names))[5])
    assert np.abs(p_snp - (np.arange(n_snps) + 0.5) / n_snps).max() < 0.02
    # (logistic / Cox: calls of at least 4,096 markers run on the thread-per-SNP kernels, smaller
    # ones on the warp-per-SNP kernel, which sums in another order: keep the sub-range on
    # the same kernel)
    a, b = n_snps // 2 - 777, n_snps // 2 + (4321 if model in ("logistic", "coxph") else 1234)
    out2, iout2 = engine.run_packed_dev(dev[a:b], b - a)
    sub = engine.fetch(out2, iout2)
    np.testing.assert_array_equal(sub.iout, res.iout[:, a:b])
    np.testing.assert_array_equal(
        sub.out.view(np.int64),
        np.ascontiguousarray(res.out[:, a:b]).view(np.int64))
    pick = np.array([0, n_snps // 2, n_snps - 1])
    packed = dev[torch.from_numpy(pick).cuda()].cpu().numpy()[:, :(n + 3) // 4]
    G = unpack_rows(packed, n)
    picked = type(res)(res.out[:, pick], res.iout[:, pick], res.n_params)
    oracle = run_oracle(model, Y, C, names, G[:, pos], interaction=inter)
    problems, worst, n_ok = compare(model, picked, oracle, names,
                                    inter_keys=list(inter or ()), rtol=tol)
    assert not problems, problems[:10]
    assert n_ok == len(pick)
    del dev, out, iout
    torch.cuda.empty_cache()


@pytest.mark.parametrize("k_extra", [14, 19, 27])
def test_linear_wide_designs(engine, k_extra):
    """Designs wider than 16 columns: the solve kernel runs one SNP per warp
    (32 lanes), the corrections use 3-4 MMA tiles per side; with and without
    the tcgen05 contraction (N = 8 (k + 1) + 1 <= 192 only)."""
    rng = np.random.default_rng(300 + k_extra)
    n, n_snps = 3000, 150
    C, names = synth_design(rng, n, k_extra=k_extra)
    G = synth_genotypes(rng, n_snps, n, missing=0.02, maf_lo=0.05, maf_hi=0.95)
    y = _linear_y(rng, C, G)
    oracle = run_oracle("linear", y[:, None], C, names, G)
    engine.set_design("linear", y, C)
    for tc in (1, 0):
        engine.set_option("contract_tc", tc)
        res = engine.run_packed_host(pack_rows(G))
        problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
        assert not problems, (tc, problems[:10])
        assert n_ok == n_snps
    engine.set_option("contract_tc", 1)
    res = engine.run_dosage_host(G)
    problems, worst, n_ok = compare("linear", res, oracle, names, rtol=LIN_TOL)
    assert not problems, problems[:10]


@pytest.mark.parametrize("model", ["linear", "logistic", "coxph"])
def test_device_rows_with_a_tight_pitch(engine, model):
    """gt_run_packed_dev only asks for a pitch that is a multiple of 16 B and
    covers the row: kernels must not depend on the 128 B padding of
    gt_packed_pitch (nor read past such rows)."""
    import torch
    rng = np.random.default_rng(55)
    n, n_snps = 1030, 300                       # 258 B per row -> pitch 272
    C, names = synth_design(rng, n, k_extra=2)
    G = synth_genotypes(rng, n_snps, n, missing=0.02)
    event = None
    if model == "linear":
        y = _linear_y(rng, C)
    elif model == "logistic":
        y = (rng.random(n) < 0.4).astype(float)
    else:
        C = C[:, :-1]
        y, event = _cox_data(rng, n, C)
    sex = rng.integers(0, 2, n).astype(float)
    packed = pack_rows(G)
    tight = np.zeros((n_snps, 272), dtype=np.uint8)
    tight[:, :packed.shape[1]] = packed
    dev = torch.from_numpy(tight).cuda()
    for inter in (False, True):
        engine.set_design(model, y, C, event=event,
                          W=C[:, [0]] if inter else None)
        engine.set_sample_sex(sex)
        ref = engine.run_packed_host(packed)
        for tc in (1, 0):
            engine.set_option("contract_tc", tc)
            out, iout = engine.run_packed_dev(dev, n_snps, pitch=272)
            got = engine.fetch(out, iout)
            np.testing.assert_array_equal(got.iout, ref.iout)
            np.testing.assert_allclose(np.nan_to_num(got.out),
                                       np.nan_to_num(ref.out), rtol=1e-9,
                                       atol=1e-12)
        engine.set_option("contract_tc", 1)
    engine.set_sample_sex(None)
