"""End-to-end throughput of analysis.execute_formula on PLINK files with a
GWASWriter sink (the reference's own user-facing call), on the GPU box.

    python scripts/e2e_execute.py [n_samples] [n_snps]
"""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from genetest_b200 import analysis                      # noqa: E402
from genetest_b200.genotypes import PlinkReader, PlinkWriter, pack_rows  # noqa: E402
from genetest_b200.phenotypes import DataFramePhenotypes  # noqa: E402
from genetest_b200.subscribers import GWASWriter        # noqa: E402
from genetest_b200 import modelspec as spec             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(20261019)
tmp = tempfile.mkdtemp(prefix="gt_e2e_")
prefix = os.path.join(tmp, "geno")
samples = ["s{:06d}".format(i) for i in range(n)]
order = rng.permutation(n)
t0 = time.perf_counter()
with PlinkWriter(prefix) as w:
    # a pool of 4096 distinct markers, written in a fresh order per block
    maf = rng.uniform(0.05, 0.5, 4096)
    g = rng.binomial(2, maf[:, None], size=(4096, n)).astype(np.int8)
    g[rng.random((4096, n)) < 0.01] = -1
    pool = pack_rows(g)
    for s in range(0, m, 4096):
        k = min(4096, m - s)
        w.write_packed(pool[rng.permutation(4096)[:k]])
    w.write_bim([(1 + j % 22, "rs{}".format(j), 1000 + j, "A", "G")
                 for j in range(m)])
    w.write_fam([samples[i] for i in order])
print("wrote {} SNPs x {} samples in {:.1f} s".format(
    m, n, time.perf_counter() - t0), flush=True)

age = rng.normal(50, 10, n)
pheno = pd.DataFrame({
    "y": 0.1 * (age - 50) / 10 + rng.normal(size=n), "age": age,
    "sex": rng.integers(0, 2, n).astype(float),
    "pc1": rng.normal(size=n), "pc2": rng.normal(size=n)}, index=samples)
for test, formula in (("linear", "y ~ SNPs + age + sex + pc1 + pc2"),):
    spec._reset()
    out = os.path.join(tmp, test + ".txt")
    writer = GWASWriter(out, test)
    import cProfile
    import pstats
    prof = cProfile.Profile() if os.environ.get("GT_PROFILE") else None
    t0 = time.perf_counter()
    if prof:
        prof.enable()
    analysis.execute_formula(DataFramePhenotypes(pheno), PlinkReader(prefix),
                             formula, test, subscribers=[writer],
                             output_prefix=os.path.join(tmp, test))
    if prof:
        prof.disable()
    dt = time.perf_counter() - t0
    if prof:
        pstats.Stats(prof).sort_stats("cumulative").print_stats(35)
    rows = sum(1 for _ in open(out)) - 1
    print("{}: {} rows in {:.2f} s -> {:.0f} SNP-tests/s end to end "
          "(file -> GPU -> TSV)".format(test, rows, dt, m / dt), flush=True)
    print(open(out).read(300))
