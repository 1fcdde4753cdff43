"""Time the two correction kernels (streaming thread-per-SNP vs gather warp-per-SNP)."""
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
packed = eng.synth_packed(n, snps, seed=1, missing_rate=0.01)
out, iout = eng.alloc_results(snps)
ref = None
for stream in (1, 0, 1):
    eng.set_option("stream_corrections", stream)
    eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    eng.enable_timing(True)
    for _ in range(3):
        eng.run_packed_dev(packed, snps, packed.stride(0), out, iout)
    eng.synchronize()
    br = eng.kernel_breakdown()
    eng.enable_timing(False)
    o = out.clone()
    if ref is None:
        ref = o
    d = (o - ref).abs().max().item()
    print("stream %d: per launch %s  max |diff| vs first %.3e" % (
        stream, {a: round(b / 3, 3) for a, b in br.items()}, d), flush=True)
