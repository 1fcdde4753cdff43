"""Where the time of analysis.execute_formula + GWASWriter goes (PLINK files on tmpfs):
wall-clock of the phases of the second (warm) analysis.

    python scripts/execute_phases.py [n_samples] [n_snps]
"""
import os, sys, tempfile, time, shutil
import numpy as np, pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genetest_b200 import analysis, engine as eng_mod
from genetest_b200.genotypes import PlinkReader
from genetest_b200.phenotypes import DataFramePhenotypes
from genetest_b200.subscribers import GWASWriter
from genetest_b200.modelspec.predicates import NameFilter
from genetest_b200 import modelspec as spec
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
eng = eng_mod.get_engine(0)
tmp = tempfile.mkdtemp(prefix="gt_ph_", dir="/dev/shm")
prefix = os.path.join(tmp, "cohort")
bps = (n + 3) // 4
packed = eng.synth_packed(n, m, seed=5, missing_rate=0.01)
with open(prefix + ".bed", "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01]))
    for s0 in range(0, m, 32768):
        f.write(packed[s0:s0 + 32768, :bps].cpu().numpy().tobytes())
del packed
with open(prefix + ".bim", "w") as f:
    f.write("".join("1\trs{}\t0\t{}\tA\tG\n".format(i, i + 1) for i in range(m)))
ids = ["s{:06d}".format(i) for i in range(n)]
with open(prefix + ".fam", "w") as f:
    f.write("".join("f{0} {0} 0 0 0 -9\n".format(s) for s in ids))
rng = np.random.default_rng(6)
y, C, names = bench.make_design("linear", n, 10, rng)
df = pd.DataFrame(C[:, :-1], columns=names[:-1], index=ids)
df["y"] = y
df = df.iloc[rng.permutation(n)]
formula = "y ~ SNPs + " + " + ".join(names[:-1])

T = {}
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T.setdefault(label, []).append(time.perf_counter() - t0)
    setattr(obj, name, w)
wrap(analysis._PackedPlan, "__init__", "plan init (variant table, filters)")
wrap(analysis._PackedPlan, "run", "plan.run (gt_run_packed_host)")
wrap(analysis._PackedPlan, "finish", "plan.finish")
wrap(analysis, "_dispatch_block", "sinks (helper thread)")
wrap(analysis, "_generate_sample_order", "sample order")
wrap(eng_mod.Engine, "set_design", "set_design")
wrap(analysis, "_execute_gwas", "_execute_gwas total")
_PR = PlinkReader.__init__
def pr_init(self, *a, **k):
    t0 = time.perf_counter(); _PR(self, *a, **k); T.setdefault("PlinkReader()", []).append(time.perf_counter() - t0)
PlinkReader.__init__ = pr_init

def run(tag, pred):
    spec._reset(); T.clear()
    out = os.path.join(tmp, tag)
    t0 = time.perf_counter()
    analysis.execute_formula(DataFramePhenotypes(df), PlinkReader(prefix), formula, "linear",
                             subscribers=[GWASWriter(out + ".tsv", "linear")],
                             variant_predicates=pred, output_prefix=out)
    dt = time.perf_counter() - t0
    print("%s: %.3f s" % (tag, dt))
    for k, v in T.items():
        print("   %-40s n=%3d total %.3f s  (%s)" % (k, len(v), sum(v), " ".join("%.3f" % x for x in v[:12])))
run("warm", [NameFilter({"rs{}".format(i) for i in range(4096)})])
run("full", None)
run("full2", None)
shutil.rmtree(tmp, ignore_errors=True)
