"""analysis.execute_formula + GWASWriter on a PLINK file, on however many ranks
torchrun started (one GPU each):

    python scripts/execute_multi.py <plink prefix> <phenotypes.csv> <out prefix>
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/execute_multi.py <plink prefix> <phenotypes.csv> <out prefix>

Rank 0 writes <out prefix>.tsv (and the failed / not-analysed lists); the other
ranks only compute.  Used by tests/test_gpu_multi.py: the file of a 2-rank run
must equal the file of a 1-rank run byte for byte."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pandas as pd                                           # noqa: E402

from genetest_b200 import analysis, sharding                  # noqa: E402
from genetest_b200 import modelspec as spec                   # noqa: E402
from genetest_b200.genotypes import PlinkReader               # noqa: E402
from genetest_b200.modelspec.predicates import MAFFilter      # noqa: E402
from genetest_b200.phenotypes import DataFramePhenotypes      # noqa: E402
from genetest_b200.subscribers import GWASWriter              # noqa: E402

prefix, pheno_csv, out = sys.argv[1:4]
pheno = pd.read_csv(pheno_csv, index_col=0)
pheno.index = pheno.index.astype(str)
covs = [c for c in pheno.columns if c != "y"]
spec._reset()
t0 = time.perf_counter()
analysis.execute_formula(
    DataFramePhenotypes(pheno), PlinkReader(prefix), "y ~ SNPs + " + " + ".join(covs),
    "linear", subscribers=[GWASWriter(out + ".tsv", "linear")],
    variant_predicates=[MAFFilter(0.02)], output_prefix=out, maf_t=0.03)
rank, world = sharding.rank_world()
print("rank {} of {}: done in {:.2f} s".format(rank, world, time.perf_counter() - t0), flush=True)
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
