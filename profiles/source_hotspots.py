"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`
into per-source-line totals (stall samples, instructions executed).

    python profiles/source_hotspots.py finish_cuda.csv [top_n]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None
hdr = None
agg = defaultdict(lambda: [0.0, 0.0, ""])
cur_line = None
for row in csv.reader(open(path, errors="replace")):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]
        hdr = None
        continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None:
        continue
    if row[0].strip():                     # a CUDA source line
        cur_line = (fname, int(row[0]))
        agg[cur_line][2] = row[1].strip()[:90]
        try:
            agg[cur_line][0] += float(row[i_samp] or 0)
            agg[cur_line][1] += float(row[i_inst] or 0)
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()) or 1.0
tot_i = sum(v[1] for v in agg.values()) or 1.0
print("total samples {:.0f}, instructions {:.0f}".format(tot_s, tot_i))
byfile = defaultdict(lambda: [0.0, 0.0])
for (f, _l), v in agg.items():
    byfile[f][0] += v[0]
    byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("{:24s} samples {:5.1f}%  inst {:5.1f}%".format(
        f, 100 * v[0] / tot_s, 100 * v[1] / tot_i))
print()
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("{:5.1f}% {:5.1f}%  {}:{}  {}".format(
        100 * v[0] / tot_s, 100 * v[1] / tot_i, f, l, v[2]))


def sass_hotspots(path, top=30):
    """Top SASS instructions of `ncu --page source --csv` (SASS view) with
    their three largest stall reasons."""
    rows = list(csv.reader(open(path, errors="replace")))
    out = []
    hdr = None
    for r in rows:
        if "Address" in r and "Source" in r and "# Samples" in r:
            hdr = r
            isrc = len(hdr) - 1 - hdr[::-1].index("Source")      # the SASS text
            isamp = hdr.index("# Samples")
            iline = hdr.index("Line No") if "Line No" in hdr else None
            stall = [(i, h) for i, h in enumerate(hdr)
                     if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if iline is not None and r[iline].strip():
            continue                        # a CUDA source line, not an instruction
        try:
            s = float(r[isamp] or 0)
        except ValueError:
            continue
        st = sorted(((float(r[i] or 0), h) for i, h in stall), reverse=True)[:3]
        out.append((s, r[isrc][:64], " ".join("%s=%d" % (h[6:], v) for v, h in st if v > 0)))
    tot = sum(s for s, _a, _b in out) or 1.0
    for s, a, b in sorted(out, reverse=True)[:top]:
        print("%5.2f%%  %-64s %s" % (100 * s / tot, a, b))


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "sass":
    print()
    sass_hotspots(sys.argv[1], top)
