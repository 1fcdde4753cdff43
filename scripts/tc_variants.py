"""Time the tcgen05 contraction with parts compiled out (tuning probe):
option "tc_debug" (variants compiled in: 1 no records, 2 no MMA, 6 no MMA / TMEM store,
14 + genotypes from L2, 61 MMA only, 63 synchronisation skeleton, 128 wait cycles per role)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
rng = np.random.default_rng(0)
n = 100000
snps = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
y, C, names = bench.make_design("linear", n, 10, rng)
eng.set_design("linear", y, C)
eng.set_option("contract_tc", 1)
packed = eng.synth_packed(n, snps, seed=1, missing_rate=0.01)
out, iout = eng.alloc_results(snps)
for dbg in [int(a) for a in sys.argv[2:]] or (0, 1, 2, 6, 14, 61, 63, 128):
    eng.set_option("tc_debug", dbg)
    for _ in range(1):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    eng.enable_timing(True)
    for _ in range(1 if dbg & 128 else 3):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    ms, k, _ = eng.kernel_time()
    br = eng.kernel_breakdown()
    eng.enable_timing(False)
    gbs = snps * 25000.0 / (ms / k * 1e-3) / 1e9
    print("dbg %2d: contraction %.3f ms per launch of %d SNPs = %.0f GB/s packed; per 3 runs: %s"
          % (dbg, ms / k, snps, gbs, {a: round(b, 3) for a, b in br.items()}), flush=True)
