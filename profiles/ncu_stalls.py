#!/usr/bin/env python
"""Per-instruction stall summary of one kernel of an ncu report:

    python profiles/ncu_stalls.py <file.ncu-rep> <kernel regex> [top_n] > out.txt

reads `ncu -i ... --page source --print-source sass --csv` and prints the share of
the warp-stall samples by opcode and by stall reason, the instruction mix
(instructions executed by opcode) and the hottest SASS instructions with their
three main stall reasons."""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv",
                      "-k", "regex:" + kern], capture_output=True, text=True).stdout
hdr, data, seen = None, [], set()
for r in csv.reader(io.StringIO(raw)):
    if "Source" in r and "# Samples" in r:
        hdr = r
        isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
        iex, iaddr = hdr.index("Instructions Executed"), hdr.index("Address")
        stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[iaddr] in seen:
        continue
    seen.add(r[iaddr])
    try:
        data.append((float(r[isamp] or 0), r[isrc].strip(), float(r[iex] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1.0
by_op, by_stall, mix = collections.Counter(), collections.Counter(), collections.Counter()
for s, src, ex, r in data:
    toks = src.split()
    op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
    by_op[op] += s
    mix[op] += ex
    for j in stall_cols:
        try:
            by_stall[hdr[j][6:]] += float(r[j] or 0)
        except ValueError:
            pass
print("kernel regex: {}   stall samples: {:.0f}   SASS instructions: {}".format(kern, tot, len(data)))
print("samples by opcode : " + ", ".join("{} {:.1f}%".format(k, 100 * v / tot) for k, v in by_op.most_common(12)))
print("samples by stall  : " + ", ".join("{} {:.1f}%".format(k, 100 * v / tot) for k, v in by_stall.most_common(10)))
tm = sum(mix.values()) or 1.0
print("instructions executed: {:.3e} warp-level; mix: ".format(tm) +
      ", ".join("{} {:.1f}%".format(k, 100 * v / tm) for k, v in mix.most_common(14)))
print()
for s, src, ex, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((hdr[j][6:], float(r[j])) for j in stall_cols if r[j] not in ("", "0")), key=lambda kv: -kv[1])[:3]
    print("{:5.1f}%  {:<60s} executed {:>9.0f}  {}".format(
        100 * s / tot, src[:60], ex, ", ".join("{} {:.0f}".format(a, b) for a, b in st)))
