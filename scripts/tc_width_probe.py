"""Contraction time against the number of digit planes N (tuning probe): designs of k covariates."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
n = 100000
snps = 151552
packed = eng.synth_packed(n, snps, seed=1, missing_rate=0.01)
for k in [int(a) for a in sys.argv[1:]] or (4, 6, 7, 8, 9, 10, 11, 12, 13, 15, 17):
    rng = np.random.default_rng(0)
    y, C, names = bench.make_design("linear", n, k, rng)
    eng.set_design("linear", y, C)
    out, iout = eng.alloc_results(snps)
    for dbg in (0,):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
        eng.synchronize()
        eng.enable_timing(True)
        for _ in range(3):
            eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
        eng.synchronize()
        ms, kk, _ = eng.kernel_time()
        br = eng.kernel_breakdown()
        eng.enable_timing(False)
        print("k=%2d %s: contraction %.3f ms per launch (%d launches / 3 runs); per 3 runs: %s"
              % (k, eng.dominant_kernel(), ms / kk, kk, {a: round(b, 3) for a, b in br.items()}), flush=True)
