"""One linear batch (75,776 markers x 100,000 samples, 1 % missing): the
profiling target for the contraction / finish kernels."""
import sys
sys.path.insert(0, ".")
import numpy as np
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
rng = np.random.default_rng(0)
n, snps = 100000, 75776
miss = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
y, C, names = bench.make_design("linear", n, 10, rng)
eng.set_design("linear", y, C)
out, iout = eng.alloc_results(snps)
packed = eng.synth_packed(n, snps, seed=1, missing_rate=miss)
for _ in range(3):
    eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
eng.synchronize()
print("ok")
