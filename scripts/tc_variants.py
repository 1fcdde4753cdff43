"""Time the tcgen05 contraction with parts disabled (tuning probe)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
rng = np.random.default_rng(0)
n, snps = 100000, 75776
y, C, names = bench.make_design("linear", n, 10, rng)
eng.set_design("linear", y, C)
eng.set_option("contract_tc", 1)
packed = eng.synth_packed(n, snps, seed=1, missing_rate=0.01)
out, iout = eng.alloc_results(snps)
for dbg in (0, 1, 16, 17, 2, 4, 8, 1 | 2 | 4 | 16, 1 | 2 | 4 | 8 | 16):
    eng.set_option("tc_debug", dbg)
    for _ in range(2):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    eng.enable_timing(True)
    for _ in range(3):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    ms, k, _ = eng.kernel_time()
    eng.enable_timing(False)
    print("dbg %2d: %.3f ms per launch" % (dbg, ms / k), flush=True)
