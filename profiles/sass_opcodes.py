#!/usr/bin/env python
"""SASS opcode counts of the built library -> profiles/sass_opcodes.txt

    python profiles/sass_opcodes.py [genetest_b200/libgenetest_b200.so]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "genetest_b200/libgenetest_b200.so"
OPS = ["UTCIMMA", "UTCBAR", "STTM", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "DMMA", "DFMA",
       "ELECT", "UTCATOMSWS", "PRMT", "POPC"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
names = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
out = ["# SASS opcode counts of {} (cuobjdump -sass), built for sm_100a".format(lib),
       "# what proves a Blackwell-native kernel: UTCIMMA = tcgen05.mma kind::i8, STTM / LDTM = tcgen05.st / .ld,",
       "# UTMALDG = cp.async.bulk.tensor (TMA tiles), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,",
       "# LDGSTS = cp.async, DMMA = FP64 mma.sync (warp-per-SNP logistic / Cox, fallback corrections)", "",
       "whole library: " + ", ".join("{} {}".format(o, tot[o]) for o in OPS), ""]
rows = []
for (mangled, c), name in zip(per.items(), names):
    short = name.replace("(int)", "").replace("(bool)", "").split("(")[0]
    hits = ", ".join("{} {}".format(o, c[o]) for o in OPS if c[o] and o not in ("SYNCS", "ELECT", "PRMT"))
    if hits:
        rows.append("{:<70s} {}".format(short[:70], hits))
out += sorted(rows)
open("profiles/sass_opcodes.txt", "w").write("\n".join(out) + "\n")
print(out[5])
