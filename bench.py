#!/usr/bin/env python
"""Benchmark of the per-SNP association hot path (SNP-tests/s).

    python bench.py --gpus N --steps K --warmup W [--model linear|logistic]
    python bench.py --impl reference ...      # CPU restatement of the reference

One step = one pass of the hot path over one shard of synthetic PLINK rows
that are already resident in HBM (BASELINE config 2 for
``linear`` -- 100,000 samples, 10 covariates + intercept, 1 % missing
genotypes, 500,000 SNPs per GPU by default; config 3 for ``logistic``).
Under torchrun every rank owns its own SNP shard (weak scaling, no data-path
collective) and rank 0 gathers the fixed-size result blocks over NCCL.
Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261019
WORKLOADS = {
    # model: (samples, covariates w/o intercept, SNPs per GPU per step, missing)
    "linear": dict(n=100000, k_cov=10, snps=500000, missing=0.01,
                   name="config2: linear GWAS 500k SNPs x 100k samples, 10 "
                        "covariates + intercept, 1% missing genotypes"),
    # (75,776 markers = two waves of 148 CTAs x 256 markers of the thread-per-SNP kernels)
    "logistic": dict(n=50000, k_cov=10, snps=75776, missing=0.0,
                     name="config3: logistic GWAS 50k samples, 10 covariates "
                          "+ intercept (batched IRLS), SNP shard per step"),
    "linear_gxe": dict(n=100000, k_cov=10, snps=200000, missing=0.01,
                       name="config4: linear GWAS with a SNP x age interaction, "
                            "200k SNPs x 100k samples, 1% missing genotypes"),
    "coxph": dict(n=20000, k_cov=10, snps=75776, missing=0.0,
                  name="config5: Cox PH (Efron) 20k samples, 10 covariates, "
                       "time/event phenotypes, SNP shard per step"),
}


def base_model(model):
    """Engine / oracle model of a workload (config 4 is the linear model with
    one SNP x covariate interaction column)."""
    return "linear" if model == "linear_gxe" else model


def interaction_of(model):
    return {"I:age": ["age"]} if model == "linear_gxe" else None


def make_design(model, n, k_cov, rng):
    model = base_model(model)
    """SURVEY.md 8(d): age ~ N(50,10) standardised, sex ~ Bernoulli(.5),
    remaining covariates iid N(0,1), intercept last."""
    age = rng.normal(50, 10, n)
    age = (age - age.mean()) / age.std()
    sex = rng.integers(0, 2, n).astype(float)
    cols = [age, sex] + [rng.normal(size=n) for _ in range(k_cov - 2)]
    names = ["age", "sex"] + ["c{}".format(i) for i in range(k_cov - 2)]
    if model == "coxph":
        # tte ~ Exp(rate = exp(0.1 z(age))), independent censoring Exp(0.5);
        # StatsCoxPH drops the intercept (survival.py:51-52)
        C = np.column_stack(cols)
        t_event = rng.exponential(1.0 / np.exp(0.1 * age))
        t_cens = rng.exponential(2.0, n)
        y = np.column_stack((np.minimum(t_event, t_cens),
                             (t_event <= t_cens).astype(float)))
        return y, C, names
    C = np.column_stack(cols + [np.ones(n)])
    names.append("intercept")
    if model == "linear":
        y = 0.1 * age + rng.normal(size=n)
    else:
        y = (rng.random(n) < 1 / (1 + np.exp(0.5 - 0.2 * age))).astype(float)
    return y, C, names


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        # samples under load: the upper half of the observed clocks
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------
# CPU arm: the oracle (restated reference path) on the host cores
# --------------------------------------------------------------------------
_CPU = {}


def _cpu_init(model, y, C, names):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass
    from oracle import gwas as og
    _CPU.update(model=base_model(model), y=y, C=C, names=names, og=og,
                interaction=interaction_of(model),
                missing=WORKLOADS[model]["missing"])


def _cpu_work(args):
    """Returns (markers done, seconds spent in the fits): the synthetic
    genotype rows are drawn outside the timed part."""
    seed, count = args
    og = _CPU["og"]
    rng = np.random.default_rng(seed)
    n = _CPU["C"].shape[0]
    y = _CPU["y"]
    y = y if y.ndim == 2 else y[:, None]
    done, spent = 0, 0.0
    chunk = 8
    while done < count:
        rows = []
        for _ in range(min(chunk, count - done)):
            maf = rng.uniform(0.05, 0.5)
            g = rng.binomial(2, maf, n).astype(float)
            g[rng.random(n) < _CPU["missing"]] = np.nan
            rows.append(g)
        t0 = time.perf_counter()
        for g in rows:
            og.gwas_one(_CPU["model"], y, _CPU["C"], _CPU["names"], g,
                        interaction=_CPU["interaction"])
        spent += time.perf_counter() - t0
        done += len(rows)
    return done, spent


def cpu_throughput(model, y, C, names, n_snps, cores, pool=None):
    """SNP-tests/s of the restated reference path on ``cores`` processes."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores, _cpu_init,
                                           (model, y, C, names))
    per = max(1, n_snps // cores)
    jobs = [(SEED + 17 * i, per) for i in range(cores)]
    res = pool.map(_cpu_work, jobs)
    done = sum(r[0] for r in res)
    # every process fits its share concurrently: the job takes as long as
    # the slowest one spends fitting (drawing the rows is not counted)
    dt = max(r[1] for r in res)
    if own:
        pool.close()
    return done / dt, done, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.model]
    rng = np.random.default_rng(SEED)
    y, C, names = make_design(args.model, w["n"], w["k_cov"], rng)
    cores = len(os.sched_getaffinity(0)) or 1
    per_step = args.ref_snps or (cores * {"linear": 48, "logistic": 12,
                                            "coxph": 1}[base_model(args.model)])
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(cores, _cpu_init,
                                       (args.model, y, C, names))
    for _ in range(args.warmup):
        cpu_throughput(args.model, y, C, names, max(cores, per_step // 4),
                       cores, pool)
    total, dt = 0, 0.0
    for _ in range(args.steps):
        _, done, step_dt = cpu_throughput(args.model, y, C, names, per_step,
                                          cores, pool)
        total += done
        dt += step_dt
    pool.close()
    value = total / dt
    sample = ("{} SNPs per step of the same synthetic workload, oracle port "
              "(statsmodels restated in numpy), {} processes x 1 BLAS thread"
              .format(per_step, cores))
    line = {
        "impl": "reference", "metric": "SNP-tests/sec", "value": value,
        "unit": "SNP-tests/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "model": args.model,
                   "samples": w["n"], "covariates": w["k_cov"] + 1},
        "cpu_baseline": {"value": value, "unit": "SNP-tests/s",
                         "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "SNP-tests/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



def tc_shape(kernel_name):
    """(N, MMA rows per marker) of the tcgen05 contraction the engine picked, from the
    kernel name it reports (contract_tc2_kernel<NT=..,TILES=..,HOM=..>), else None."""
    import re
    m = re.match(r"contract_tc2_kernel<NT=(\d+),TILES=(\d+),HOM=(\d+)>", kernel_name or "")
    return (16 * int(m.group(1)), 2 if int(m.group(3)) else 1) if m else None


def tc_roofline(kernel_name, n, per_launch_snps, kernel_ms, peaks):
    """int8 tensor-core figures of one contraction launch: 2 n N operations per MMA row;
    peak = twice the measured dense bf16 rate (kind::i8 runs at twice the bf16 rate)."""
    N, rows = tc_shape(kernel_name)
    ops = 2.0 * n * N * rows
    peak = 2.0 * float(peaks.get("bf16_tflops", 2250.0))
    ach = per_launch_snps * ops / (kernel_ms * 1e-3) / 1e12
    return {"pipe": "tcgen05 kind::i8", "achieved_tops": ach, "peak_tops": peak,
            "frac": ach / peak,
            "peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json (burst)" if "bf16_tflops" in peaks
                           else "2 x nominal dense bf16", "ops_per_snp": ops, "N": N,
            "mma_rows_per_marker": rows}


def flops_per_snp(workload, n, p, mean_iter):
    """FP64 work of one marker."""
    model = base_model(workload)
    if workload == "linear_gxe":
        # (g, g w) against [q, y, 1, w] and (g^2) against [1, w, w^2]
        return 2.0 * n * (2 * (p + 1) + 3)
    if model == "linear":
        return 2.0 * n * (p + 2)              # contraction columns (k+2) + g^2
    if model == "logistic":
        # n (p(p+1) + 4p + 40) per IRLS pass, start pass included
        return n * (p * (p + 1) + 4 * p + 40.0) * (mean_iter + 1.0)
    # n (p(p+1) + 4p + 30) per pass over the samples; two passes per Newton
    # step (risk-set totals, then score + Hessian) + the final one
    return n * (p * (p + 1) + 4 * p + 30.0) * (2.0 * mean_iter + 2.0)


def set_workload_design(eng, workload, n, rng_seed):
    w = WORKLOADS[workload]
    y, C, names = make_design(workload, n, w["k_cov"], np.random.default_rng(rng_seed))
    pos = np.random.default_rng(rng_seed + 1).permutation(n).astype(np.int32)
    model = base_model(workload)
    if model == "coxph":
        eng.set_design(model, y[:, 0], C, event=y[:, 1], sample_pos=pos, n_file=n)
    else:
        eng.set_design(model, y, C, sample_pos=pos, n_file=n,
                       W=C[:, [0]] if workload == "linear_gxe" else None)
    return y, C, names


def measure_side_workload(eng, torch, dist, rank, world, workload, steps,
                          warmup, peaks, snps=None, cpu=True):
    """One of the other BASELINE configs (or the strong-scaling shard of config
    2) on this rank's GPU: device-resident throughput (CUDA events, max over
    ranks, every rank its own shard, no gather), the FP64 / HBM roofline of its
    dominant kernel, the end-to-end figure through gt_run_packed_host and a
    small CPU sample."""
    w = WORKLOADS[workload]
    n = w["n"]
    snps = snps or w["snps"]
    y, C, names = set_workload_design(eng, workload, n, SEED)
    packed = eng.synth_packed(n, snps, seed=SEED + 1000 * rank + 7,
                              missing_rate=w["missing"])
    pitch = packed.stride(0)
    out, iout = eng.alloc_results(snps)
    for _ in range(warmup):
        eng.run_packed_dev(packed, snps, pitch, out, iout)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.enable_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        eng.run_packed_dev(packed, snps, pitch, out, iout)
    ev1.record()
    torch.cuda.synchronize()
    eng.synchronize()
    ms = ev0.elapsed_time(ev1)
    kms, klaunches, all_launches = eng.kernel_time()
    breakdown = eng.kernel_breakdown()
    eng.enable_timing(False)
    # end to end with host buffers
    # (the iterative models: whole waves of their thread-per-SNP kernels, 148 CTAs x 256 markers)
    e2e_snps = min(snps, 65536 if base_model(workload) == "linear" else 75776)
    hp = torch.empty((e2e_snps, (n + 3) // 4), dtype=torch.uint8).pin_memory()
    hp.copy_(packed[:e2e_snps, :(n + 3) // 4].cpu())
    ho = torch.empty((eng.n_fields, e2e_snps), dtype=torch.float64).pin_memory()
    hi = torch.empty((4, e2e_snps), dtype=torch.int32).pin_memory()
    eng.run_packed_host(hp.numpy(), ho.numpy(), hi.numpy())
    t0 = time.perf_counter()
    for _ in range(2):
        eng.run_packed_host(hp.numpy(), ho.numpy(), hi.numpy())
    e2e_dt = (time.perf_counter() - t0) / 2
    if world > 1:
        t = torch.tensor([ms, e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_dt = float(t[0].item()), float(t[1].item())
    p = eng.n_params
    mean_iter = float(iout[3].double().mean().item())
    n_ok = int((iout[0] == 0).sum().item())
    kernel_ms = kms / max(1, klaunches)
    per_launch = snps * steps / max(1, klaunches)
    fl = flops_per_snp(workload, n, p, mean_iter)
    achieved_tf = per_launch * fl / (kernel_ms * 1e-3) / 1e12
    bytes_per_snp = (n + 3) // 4 + 8 * (6 * p + 5) + 8
    achieved_gbs = per_launch * bytes_per_snp / (kernel_ms * 1e-3) / 1e9
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    model = base_model(workload)
    if model == "linear" and tc_shape(eng.dominant_kernel()):
        roof = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "kernel": eng.dominant_kernel(),
                "compute": tc_roofline(eng.dominant_kernel(), n, per_launch, kernel_ms, peaks)}
    else:
        # interaction designs run the FP64 FMA contraction, the iterative models
        # accumulate with FP64 tensor-core MMAs: each against its own measured peak
        fma = workload == "linear_gxe" or eng.dominant_kernel() in ("logistic_thread_kernel", "cox_thread_kernel")
        peak = eng.fp64_peak_tflops() if fma else eng.dmma_peak_tflops()
        roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved_tf / peak if peak else None,
                "peak_source": "measured FP64 {} rate (gt_measure_{}_peak)".format(
                    "FMA" if fma else "m8n8k4 MMA", "fp64" if fma else "dmma"),
                "kernel": eng.dominant_kernel(), "flop_per_snp": fl,
                "mean_iterations": mean_iter,
                "hbm": {"achieved_gbs": achieved_gbs, "frac": achieved_gbs / hbm_peak}}
    roof.update({"kernel_ms_per_launch": kernel_ms, "kernel_share_of_step": kms / ms,
                 "step_breakdown_ms": {k: v / steps for k, v in breakdown.items()}})
    res = {
        "workload": w["name"], "model": workload, "samples": n, "params": p,
        "snps_per_gpu_per_step": snps, "steps": steps,
        "value": world * snps * steps / (ms * 1e-3), "unit": "SNP-tests/s",
        "ms_per_step": ms / steps, "ok_markers_last_step": n_ok,
        "roofline": roof,
        "e2e": {"value": world * e2e_snps / e2e_dt, "unit": "SNP-tests/s",
                "h2d_bytes_per_step": int(hp.numel()),
                "d2h_bytes_per_step": int(ho.numel() * 8 + hi.numel() * 4),
                "snps_per_step": e2e_snps,
                "call": "gt_run_packed_host (pinned host buffers)"},
        "gpu_launches": int(all_launches),
    }
    if cpu and rank == 0:
        cores = len(os.sched_getaffinity(0)) or 1
        cpu_snps = cores * {"linear": 32, "logistic": 12, "coxph": 1}[model]
        v, done, dt = cpu_throughput(workload, y, C, names, cpu_snps, cores)
        res["cpu_baseline"] = {"value": v, "unit": "SNP-tests/s", "cores": cores, "kind": "port",
                               "sample": "{} SNPs in {:.1f} s of fitting, oracle port, {} processes x 1 "
                                         "BLAS thread".format(done, dt, cores)}
    del packed, out, iout
    torch.cuda.empty_cache()
    return res


def measure_h2d_ceiling(torch, dist, world, mib=512, reps=16):
    """What the box can move host -> device when every rank copies at once: each
    rank streams ``reps`` x ``mib`` MiB from pinned memory to its GPU (one
    cudaMemcpyAsync per copy); aggregate GB/s over the slowest rank.  The
    end-to-end figure is bounded by it."""
    src = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(mib << 20, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return world * reps * (mib << 20) / dt / 1e9


def measure_execute(torch, eng, n, snps, k_cov):
    """End to end through the reference's entry point: ``analysis.execute_formula``
    + ``GWASWriter`` on a real PLINK .bed / .bim / .fam (tmpfs), phenotype table
    in a random order, rows formatted to a TSV.  One warm-up analysis on a slice
    of the file (engine / workspace creation), then the whole file is timed."""
    import shutil
    import tempfile
    import pandas as pd
    from genetest_b200 import analysis
    from genetest_b200 import modelspec as spec
    from genetest_b200.genotypes import PlinkReader
    from genetest_b200.modelspec.predicates import NameFilter
    from genetest_b200.phenotypes import DataFramePhenotypes
    from genetest_b200.subscribers import GWASWriter
    bps = (n + 3) // 4
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    try:
        free = shutil.disk_usage(base or tempfile.gettempdir()).free
    except OSError:
        free = 0
    while snps > 16384 and free < 1.3 * snps * bps + (1 << 30):
        snps //= 2
    tmp = tempfile.mkdtemp(prefix="gt_bench_", dir=base)
    try:
        prefix = os.path.join(tmp, "cohort")
        packed = eng.synth_packed(n, snps, seed=SEED + 5, missing_rate=0.01)
        with open(prefix + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            step = 32768
            for s0 in range(0, snps, step):
                f.write(packed[s0:s0 + step, :bps].cpu().numpy().tobytes())
        del packed
        torch.cuda.empty_cache()
        with open(prefix + ".bim", "w") as f:
            f.write("".join("1\trs{}\t0\t{}\tA\tG\n".format(i, i + 1) for i in range(snps)))
        ids = ["s{:06d}".format(i) for i in range(n)]
        with open(prefix + ".fam", "w") as f:
            f.write("".join("f{0} {0} 0 0 0 -9\n".format(s) for s in ids))
        rng = np.random.default_rng(SEED + 6)
        y, C, names = make_design("linear", n, k_cov, rng)
        df = pd.DataFrame(C[:, :-1], columns=names[:-1], index=ids)
        df["y"] = y
        df = df.iloc[rng.permutation(n)]
        formula = "y ~ SNPs + " + " + ".join(names[:-1])

        names_q = {"rs{}".format(i) for i in range(snps // 4)}

        def run(tag, predicates):
            spec._reset()
            out = os.path.join(tmp, tag)
            t0 = time.perf_counter()
            analysis.execute_formula(
                DataFramePhenotypes(df), PlinkReader(prefix), formula, "linear",
                subscribers=[GWASWriter(out + ".tsv", "linear")],
                variant_predicates=predicates, output_prefix=out)
            dt = time.perf_counter() - t0
            rows = sum(1 for _ in open(out + ".tsv")) - 1
            return dt, rows, os.path.getsize(out + ".tsv")
        # warm-up on the whole file: the engine's pinned staging buffers and device
        # workspaces are sized by the first call of that size (cudaHostAlloc / cudaMalloc of
        # several GB, ~1 s, once per process) -- context creation, not the path
        run("warm", None)
        dt_q, rows_q, _ = run("quarter", [NameFilter(names_q)])
        dt, rows, tsv_bytes = run("full", None)
        # the analysis has a fixed part (phenotype alignment, QR and digit planes of the
        # design, .bim parsing) and a per-marker part: the rate of the latter from the
        # two sizes
        marginal = (rows - rows_q) / max(dt - dt_q, 1e-9)
        return {"value": rows / dt, "unit": "SNP-tests/s", "seconds": dt,
                "snps": snps, "rows_written": rows,
                "quarter_file_seconds": dt_q, "marginal_snps_per_s": marginal,
                "bed_bytes": snps * bps + 3, "tsv_bytes": tsv_bytes,
                "call": "analysis.execute_formula + GWASWriter on a PLINK .bed "
                        "(tmpfs, mmap -> pinned staging -> H2D), TSV written; "
                        ".fam / .bim parsing, phenotype alignment and design set-up "
                        "(a fixed ~0.5 s) inside the timed region; marginal_snps_per_s "
                        "is the per-marker rate from the quarter-file and whole-file runs"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from genetest_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # host buffers next to the GPU (NUMA): matters for the end-to-end leg
        from genetest_b200 import sharding
        sharding.bind_to_device_cpus(local_rank)
    w = WORKLOADS[args.model]
    n = args.samples or w["n"]
    snps = args.snps or w["snps"]
    if args.scaling == "strong":
        # the configuration as named: the SNPs of ONE job split over the GPUs
        snps = (snps + world - 1) // world
    rng = np.random.default_rng(SEED)
    y, C, names = make_design(args.model, n, w["k_cov"], rng)

    eng = Engine(local_rank)
    eng.use_torch_stream()
    # sample ids in a random permutation relative to the phenotype table
    pos = np.random.default_rng(SEED + 1).permutation(n).astype(np.int32)
    workload = args.model
    args.model = base_model(workload)
    if args.model == "coxph":
        eng.set_design(args.model, y[:, 0], C, event=y[:, 1], sample_pos=pos,
                       n_file=n)
    else:
        eng.set_design(args.model, y, C, sample_pos=pos, n_file=n,
                       W=C[:, [0]] if workload == "linear_gxe" else None)
    if args.model == "linear":
        eng.set_option("contract_tc", int(args.contract == "tc"))
    packed = eng.synth_packed(n, snps, seed=SEED + 1000 * rank,
                              missing_rate=w["missing"])
    pitch = packed.stride(0)
    out, iout = eng.alloc_results(snps)
    # Result blocks are double-buffered: the gather of step i (the only
    # collective: fixed-size blocks -> rank 0, on NCCL's stream) overlaps the
    # kernels of step i + 1, as it does when an analysis streams its batches.
    gather_on = world > 1 and not os.environ.get("GT_BENCH_NO_GATHER")
    bufs = [(out, iout)]
    if gather_on:
        bufs.append(eng.alloc_results(snps))
    gathered = [None, None]
    if gather_on and rank == 0:
        gathered = [([torch.empty_like(o) for _ in range(world)],
                     [torch.empty_like(io) for _ in range(world)])
                    for o, io in bufs]
    pending = [None, None]
    torch.cuda.synchronize()
    step_no = [0]

    def drain(slot):
        if pending[slot] is not None:
            for work in pending[slot]:
                work.wait()                 # the compute stream waits, not the host
            pending[slot] = None

    def step():
        slot = step_no[0] % len(bufs)
        step_no[0] += 1
        o, io = bufs[slot]
        drain(slot)                         # this block's previous gather is done
        eng.run_packed_dev(packed, snps, pitch, o, io)
        if gather_on:
            g = gathered[slot] if rank == 0 else (None, None)
            pending[slot] = (dist.gather(o, g[0], dst=0, async_op=True),
                             dist.gather(io, g[1], dst=0, async_op=True))

    for _ in range(args.warmup):
        step()
    drain(0), drain(1)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # every rank enters the timed region together (rank 0 just slept)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.enable_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    drain(0), drain(1)                      # every gather inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.synchronize()
    ms = ev0.elapsed_time(ev1)
    if os.environ.get("GT_BENCH_DEBUG"):
        print("rank {} local ms/step {:.2f}".format(rank, ms / args.steps),
              file=sys.stderr, flush=True)
    kms, klaunches, all_launches = eng.kernel_time()
    breakdown = eng.kernel_breakdown()
    eng.enable_timing(False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out, iout = bufs[(step_no[0] - 1) % len(bufs)]
    status = iout[0]
    n_tested = int(status.numel())
    n_ok = int((status == 0).sum().item())

    # ---- end to end through the C ABI with HOST buffers (H2D + D2H inside) ----
    e2e_snps = min(snps, args.e2e_snps if base_model(workload) == "linear" else max(args.e2e_snps, 75776))
    host_packed = torch.empty((e2e_snps, (n + 3) // 4), dtype=torch.uint8).pin_memory()
    host_packed.copy_(packed[:e2e_snps, :(n + 3) // 4].cpu())
    host_out = torch.empty((eng.n_fields, e2e_snps), dtype=torch.float64).pin_memory()
    host_iout = torch.empty((4, e2e_snps), dtype=torch.int32).pin_memory()
    hp, ho, hi = host_packed.numpy(), host_out.numpy(), host_iout.numpy()
    eng.run_packed_host(hp, ho, hi)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        eng.run_packed_host(hp, ho, hi)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = world * e2e_snps * e2e_steps / e2e_dt
    h2d_ceiling = measure_h2d_ceiling(torch, dist, world)
    del host_packed, host_out, host_iout

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # everything the main line needs from the engine, before other designs replace this one
    p = eng.n_params
    mean_iter = float(iout[3].double().mean().item())
    dominant = eng.dominant_kernel()
    fp64_peak = eng.fp64_peak_tflops()
    dmma_peak = eng.dmma_peak_tflops()
    del packed, bufs, out, iout, gathered
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and the strong-scaling shard, same run ----
    others, strong = [], None
    if not args.no_side:
        side_steps = max(2, min(args.steps, 3))
        if workload == "linear" and args.scaling == "weak":
            strong = measure_side_workload(
                eng, torch, dist, rank, world, "linear", side_steps, 2, peaks,
                snps=(w["snps"] + world - 1) // world, cpu=False)
            strong["scaling"] = "strong"
            strong["snps_total_per_step"] = w["snps"]
        for wl in ("logistic", "linear_gxe", "coxph"):
            if wl != workload:
                others.append(measure_side_workload(
                    eng, torch, dist, rank, world, wl, side_steps, 2, peaks))

    e2e_execute = None
    if world == 1 and args.execute_snps > 0 and workload == "linear":
        try:
            e2e_execute = measure_execute(torch, eng, n, args.execute_snps, w["k_cov"])
        except Exception as exc:                     # e.g. no room on tmpfs
            e2e_execute = {"unavailable": "{}: {}".format(type(exc).__name__, exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    bytes_per_snp = (n + 3) // 4 + 8 * (6 * p + 5) + 8
    per_launch_snps = snps * args.steps / max(1, klaunches)
    kernel_ms = kms / max(1, klaunches)
    achieved = per_launch_snps * bytes_per_snp / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:        # dram bytes per launch of the dominant kernel, from the ncu capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = prof.get(workload, {})
        traffic = rec.get("dram_bytes_per_launch")
        if traffic and rec.get("snps_per_launch"):
            # the capture is one launch of `snps_per_launch` markers: scale to this run's launches
            traffic = traffic * per_launch_snps / rec["snps_per_launch"]
    except Exception:
        pass
    flop_per_snp = flops_per_snp(workload, n, p, mean_iter)
    fp64_achieved = per_launch_snps * flop_per_snp / (kernel_ms * 1e-3) / 1e12
    on_tc = base_model(workload) == "linear" and tc_shape(dominant) is not None
    if on_tc:
        # exact fixed point on the int8 tensor cores: 7 digit planes for each column
        # that is not derived from the constant + the plane of ones, rounded up to the
        # N = 16 * ceil(.) columns of the MMA (interaction designs: a second MMA row per
        # marker carries the homozygote indicator for the g^2 sums)
        compute = tc_roofline(dominant, n, per_launch_snps, kernel_ms, peaks)
        compute.update({"fp64_fma_peak_tflops_measured": fp64_peak,
                        "fp64_dmma_peak_tflops_measured": dmma_peak})
    else:
        compute = {"pipe": "fp64", "achieved_tflops": fp64_achieved,
                   "peak_tflops_measured": fp64_peak,
                   "dmma_peak_tflops_measured": dmma_peak,
                   "frac": fp64_achieved / fp64_peak if fp64_peak else None,
                   "flop_per_snp": flop_per_snp,
                   "mean_iterations": mean_iter}

    # ---- CPU baseline: restated reference path on the host cores ----
    cores = len(os.sched_getaffinity(0)) or 1
    cpu_snps = args.cpu_snps or (cores * {"linear": 200, "logistic": 50,
                                          "coxph": 2}[args.model])
    cpu_value, cpu_done, cpu_dt = cpu_throughput(workload, y, C, names,
                                                 cpu_snps, cores)
    value = world * snps * args.steps / (ms * 1e-3)
    line = {
        "metric": "SNP-tests/sec", "value": value, "unit": "SNP-tests/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["name"], "model": workload, "samples": n,
                   "covariates": w["k_cov"] + 1, "params": p,
                   "snps_per_gpu_per_step": snps,
                   "l2": "inputs ({:.1f} GB packed per GPU) larger than L2"
                         .format(snps * pitch / 1e9),
                   "parallelism": "snp-shard x{}".format(world),
                   "contraction": args.contract if args.model == "linear"
                   else None,
                   "ok_markers_last_step": n_ok, "markers_last_step": n_tested},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": dominant,
                     "kernel_ms_per_launch": kernel_ms,
                     "kernel_share_of_step": kms / ms,
                     "step_breakdown_ms": {k: v / args.steps
                                           for k, v in breakdown.items()},
                     "algorithmic_bytes_per_snp": bytes_per_snp,
                     "compute": compute},
        "cpu_baseline": {"value": cpu_value, "unit": "SNP-tests/s",
                         "cores": cores, "kind": "port",
                         "sample": "{} SNPs of the same synthetic workload in "
                                   "{:.1f} s, oracle port, {} processes x 1 "
                                   "BLAS thread".format(cpu_done, cpu_dt, cores)},
        "e2e": {"value": e2e_value, "unit": "SNP-tests/s",
                "h2d_bytes_per_step": int(hp.nbytes),
                "d2h_bytes_per_step": int(ho.nbytes + hi.nbytes),
                "snps_per_step": e2e_snps, "steps": e2e_steps,
                "call": "gt_run_packed_host (pinned host buffers)",
                # the copy-in alone, all ranks at once, against what it achieved
                "h2d_ceiling_gbs": h2d_ceiling,
                "h2d_achieved_gbs": e2e_value * ((n + 3) // 4) / 1e9,
                "frac_of_h2d_ceiling": e2e_value * ((n + 3) // 4) / 1e9 / h2d_ceiling},
        "gpu_launches": int(all_launches),
        "clocks": clocks,
    }
    if workload != "linear" and not on_tc:
        # the iterative models are bound by the FP64 tensor pipe, the FP64
        # contraction of interaction designs by the FP64 FMA pipe, not by HBM:
        # report that roofline, keep the HBM figures beside it
        r = line["roofline"]
        fma = workload == "linear_gxe" or dominant in ("logistic_thread_kernel", "cox_thread_kernel")
        pk = fp64_peak if fma else dmma_peak
        r.update({"bound": "fp64" if fma else "tensor", "unit": "TFLOP/s",
                  "achieved": fp64_achieved, "peak": pk,
                  "frac": fp64_achieved / pk if pk else None,
                  "peak_source": "measured FP64 {} rate (gt_measure_{}_peak; "
                                 "not in MEASURED_PEAKS.json)".format(
                                     "FMA" if fma else "m8n8k4 MMA",
                                     "fp64" if fma else "dmma"),
                  "hbm": {"achieved_gbs": achieved, "peak_gbs": hbm_peak,
                          "frac": achieved / hbm_peak}})
    if e2e_execute is not None:
        line["e2e_execute"] = e2e_execute
    if strong is not None:
        line["strong_scaling"] = strong
    if others:
        line["other_workloads"] = others
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="linear", choices=sorted(WORKLOADS))
    ap.add_argument("--snps", type=int, default=0, help="SNPs per GPU per step")
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--contract", default="tc", choices=["tc", "fp64"],
                    help="linear contraction kernel: tcgen05 int8 or FP64")
    ap.add_argument("--e2e-snps", type=int, default=65536)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU its own full shard; strong: the SNPs "
                         "of one job (500k for config 2) split over the GPUs")
    ap.add_argument("--execute-snps", type=int, default=262144,
                    help="markers of the PLINK file of the analysis.execute leg "
                         "(0 = skip; one GPU only)")
    ap.add_argument("--no-side", action="store_true",
                    help="skip the other configs / the strong-scaling shard")
    ap.add_argument("--cpu-snps", type=int, default=0)
    ap.add_argument("--ref-snps", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
