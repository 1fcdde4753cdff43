"""Time the linear finish kernel against the missing rate (tuning probe)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
rng = np.random.default_rng(0)
n, snps = 100000, 75776
y, C, names = bench.make_design("linear", n, 10, rng)
eng.set_design("linear", y, C)
out, iout = eng.alloc_results(snps)
for miss in (0.0, 0.0025, 0.005, 0.01, 0.02):
    packed = eng.synth_packed(n, snps, seed=1, missing_rate=miss)
    for _ in range(2):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    eng.enable_timing(True)
    for _ in range(3):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    print("missing %.4f:" % miss, {k: round(v / 3, 3) for k, v in
                                   eng.kernel_breakdown().items()}, flush=True)
    eng.enable_timing(False)
