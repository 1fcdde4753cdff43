"""Thread-per-SNP Cox kernel with tied event times (times rounded to a grid): agreement with the
warp-per-SNP kernel and time."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from genetest_b200.engine import Engine
import bench
eng = Engine(0)
n, snps = 20000, int(sys.argv[1]) if len(sys.argv) > 1 else 37888
grid = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
y, C, names = bench.make_design("coxph", n, 10, np.random.default_rng(bench.SEED))
tte = np.ceil(y[:, 0] * grid) / grid
_, counts = np.unique(tte[y[:, 1] == 1], return_counts=True)
print("distinct event times", len(counts), "largest group", counts.max(), "events in groups", int(counts[counts > 1].sum()))
eng.set_design("coxph", tte, C, event=y[:, 1])
print(eng.dominant_kernel())
packed = eng.synth_packed(n, snps, seed=5, missing_rate=0.0)
res = {}
for mode in (1, 0):
    eng.set_option("thread_iter", mode)
    out, iout = eng.alloc_results(snps)
    eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    t0 = time.perf_counter()
    eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    dt = time.perf_counter() - t0
    res[mode] = (out.cpu().numpy().copy(), iout.cpu().numpy().copy())
    print("thread_iter=%d: %.1f ms per %d markers = %.0f SNP-tests/s; status %s; mean steps %.3f"
          % (mode, dt * 1e3, snps, snps / dt, np.unique(res[mode][1][0], return_counts=True), res[mode][1][3].mean()), flush=True)
a, b = res[1], res[0]
print("iout equal:", [bool((a[1][f] == b[1][f]).all()) for f in range(4)])
with np.errstate(all="ignore"):
    rel = np.abs(a[0] - b[0]) / np.maximum(np.abs(b[0]), 1e-300)
rel[np.isnan(a[0]) & np.isnan(b[0])] = 0
print("worst relative difference", np.nanmax(rel))
