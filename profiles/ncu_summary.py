#!/usr/bin/env python
"""Condense ncu output into the small tracked files under profiles/.

  python profiles/ncu_summary.py full  gpurun_out/prof.ncu-rep     profiles/rNN_ncu_full.csv
  python profiles/ncu_summary.py list  gpurun_out/launches.csv     profiles/rNN_launches.csv

`full`: one row per profiled launch of an `ncu --set full` report (read with `ncu -i ... --page raw --csv`).
`list`: per-kernel totals of a `--metrics gpu__time_duration.sum` launch list (ours and ATen's), with
each kernel's share of the listed device time.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:60]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["%s [%s]" % (k, units[i]) if units[i] else k for k, i in cols])
        for r in rows[2:]:
            w.writerow([short(r[i]) if k == "Kernel Name" else r[i] for k, i in cols])
    print("wrote", out, len(rows) - 2, "launches")


def launch_list(src, out):
    lines = [l for l in open(src, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_listed_time", "grid(last)", "block(last)"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], "%.1f" % (a[1] / 1e3), "%.1f" % (a[1] / 1e3 / a[0]), "%.4f" % (a[1] / tot), a[2], a[3]])
    print("wrote", out, len(agg), "kernels,", "%.2f ms listed" % (tot / 1e6))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
