"""Thread-per-SNP logistic kernel against the warp-per-SNP one (same markers): agreement and time."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
import os
eng.set_option("thread_fold", int(os.environ.get("GT_FOLD", "1")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
snps = int(sys.argv[1]) if len(sys.argv) > 1 else 33152
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
bench.set_workload_design(eng, "logistic", n, bench.SEED) if n == 50000 else None
if n != 50000:
    y, C, names = bench.make_design("logistic", n, 10, np.random.default_rng(1))
    eng.set_design("logistic", y, C)
packed = eng.synth_packed(n, snps, seed=5, missing_rate=miss)
res = {}
for mode in (1, 0):
    eng.set_option("thread_iter", mode)
    out, iout = eng.alloc_results(snps)
    eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    dt = (time.perf_counter() - t0) / 2
    res[mode] = (out.cpu().numpy().copy(), iout.cpu().numpy().copy())
    print("thread_iter=%d: %.1f ms per %d markers = %.0f SNP-tests/s; status counts %s; mean iterations %.3f"
          % (mode, dt * 1e3, snps, snps / dt, np.unique(res[mode][1][0], return_counts=True),
             res[mode][1][3].mean()), flush=True)
a, b = res[1], res[0]
print("iout equal:", [bool((a[1][f] == b[1][f]).all()) for f in range(4)])
ok = (a[1][0] == 0) & (b[1][0] == 0)
with np.errstate(all="ignore"):
    rel = np.abs(a[0] - b[0]) / np.maximum(np.abs(b[0]), 1e-300)
rel[:, ~ok] = 0
rel[np.isnan(a[0]) & np.isnan(b[0])] = 0
print("worst relative difference per field:", np.nanmax(rel, axis=1)[:12], "overall", np.nanmax(rel))
