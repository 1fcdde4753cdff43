#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep) and launch lists into the small
text summaries committed under profiles/.

    python profiles/summarise_ncu.py rep  <file.ncu-rep> <out.md>
    python profiles/summarise_ncu.py list <launches.csv> <out.csv>
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full summary of `{}`\n\n".format(path))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## {}\n\n| metric | value | unit |\n|---|---|---|\n".format(
                d["Kernel Name"]))
            for k in KEYS:
                if k in d:
                    f.write("| {} | {} | {} |\n".format(
                        k, d[k], units[hdr.index(k)]))
            f.write("\n")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if "Kernel Name" in r][0]
    i0 = rows.index(hdr)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[i0 + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ns,mean_ns,share_of_gpu_time\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("{},{},{:.0f},{:.0f},{:.4f}\n".format(
                k, v[0], v[1], v[1] / v[0], v[1] / tot))


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
