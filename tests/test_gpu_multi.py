"""Multi-GPU ``analysis.execute`` on hardware: one process per GPU under torchrun,
NCCL gather of the result blocks round by round, sinks on rank 0 only.  Needs two
visible GPUs (skipped otherwise); the CPU suite covers the same logic with gloo
(tests/test_host_api.py::test_multi_rank_packed_rounds_gloo)."""

import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from genetest_b200.genotypes import PlinkWriter

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_execute_writes_the_same_bytes(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(12)
    n_file, n, n_snps = 3000, 2700, 60000
    ids = np.array(["s{:05d}".format(i) for i in range(n_file)])
    prefix = str(tmp_path / "cohort")
    with PlinkWriter(prefix) as w:
        maf = rng.uniform(0.005, 0.5, 2048)
        pool = rng.binomial(2, maf[:, None], size=(2048, n_file)).astype(np.int8)
        pool[rng.random(pool.shape) < 0.01] = -1
        for j in range(n_snps):
            w.write_genotypes(pool[(j * 7919) % 2048])
        w.write_bim([(1 + j % 22, "rs{}".format(j), 100 + j, "A", "G") for j in range(n_snps)])
        w.write_fam(list(ids))
    keep = rng.permutation(n_file)[:n]
    pheno = pd.DataFrame({"y": rng.normal(size=n), "age": rng.normal(50, 10, n),
                          "pc1": rng.normal(size=n)}, index=ids[keep])
    pheno.to_csv(str(tmp_path / "pheno.csv"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    script = os.path.join(ROOT, "scripts", "execute_multi.py")
    one = str(tmp_path / "one")
    two = str(tmp_path / "two")
    subprocess.run([sys.executable, script, prefix, str(tmp_path / "pheno.csv"), one],
                   check=True, env=env, timeout=600)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                    "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29517", script, prefix,
                    str(tmp_path / "pheno.csv"), two], check=True, env=env, timeout=600)
    for suffix in (".tsv", "_failed_snps.txt", "_not_analyzed_snps.txt"):
        a, b = open(one + suffix, "rb").read(), open(two + suffix, "rb").read()
        assert len(a) > 0 and a == b, suffix
    assert open(one + ".tsv").read().count("\n") > 30000
