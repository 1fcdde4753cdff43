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
