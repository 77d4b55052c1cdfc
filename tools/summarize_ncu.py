#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  launch list : python tools/summarize_ncu.py launches gpurun_out/rN_launches.csv > profiles/rNN_launches.txt
  full capture: python tools/summarize_ncu.py full gpurun_out/rN_prof.ncu-rep  > profiles/rNN_top_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[kn].split("(")[0].replace("void ", ""), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / 1e3 / v[0]:10.1f} {v[1] / tot:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(m) for m in ["Kernel Name"] + METRICS if m in hdr]
    print(f"# ncu --set full --clock-control none ({path}); one line per captured launch")
    for r in rows[2:]:
        print("; ".join(f"{hdr[i]}={r[i]}{(' ' + units[i]) if units[i] else ''}" for i in idx))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
