"""Bandwidth of the SNP-row streaming pattern (tuning probe)."""
import sys
sys.path.insert(0, ".")
import torch
from genetest_b200.engine import Engine
eng = Engine(0)
n = 100000
for rows in (37888, 75776, 303104):
    packed = eng.synth_packed(n, rows, seed=1)
    torch.cuda.synchronize()
    for tpr, vec in ((1, 1), (1, 2), (1, 8), (2, 1), (2, 4), (4, 1), (4, 2), (1, 32), (2, 16), (4, 8)):
        print(rows, "rows tpr", tpr, "vec", vec, "-> %.0f GB/s" % eng.read_pattern_gbs(packed, tpr, vec), flush=True)
    del packed
