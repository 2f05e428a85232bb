#!/usr/bin/env python
"""Summarise ncu output for profiles/.

  python tools/ncu_summary.py launches <launches.csv> [header text]   -> per-kernel totals / shares of a launch list
  python tools/ncu_summary.py full <prof.ncu-rep> [header text]       -> key metrics of every captured launch (--set full)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").strip()


def launches(path, header):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r["Metric Unit"] in ("us", "usecond"):
                v *= 1e3
            elif r["Metric Unit"] in ("ms", "msecond"):
                v *= 1e6
            rows.append((short(r["Kernel Name"]), v))
    tot = sum(v for _, v in rows)
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    if header:
        print("# " + header)
    print(f"# total {tot / 1e3:.1f} us over {len(rows)} launches (cold-cache, serialised: compare shares, not absolutes)")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} n={n:4d} tot={v / 1e3:9.1f} us avg={v / 1e3 / n:7.1f} us share={v / tot:.3f}")


def full(path, header):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines(True) if l.startswith('"')]
    rd = csv.reader(io.StringIO("".join(lines)))
    head = next(rd)
    units = next(rd)
    if header:
        print("# " + header)
    idx = {h: i for i, h in enumerate(head)}
    stall_cols = [h for h in head if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in head if h.startswith("smsp__average_warp_latency_issue_stalled_") or h.startswith("smsp__average_warps_issue_stalled_")]
    for r in rd:
        print("--- " + short(r[idx["Kernel Name"]]) + "  grid " + r[idx.get("Grid Size", 0)] + " block " + r[idx.get("Block Size", 0)])
        for k in KEYS:
            if k in idx:
                print(f"  {k:66s} {r[idx[k]]} {units[idx[k]]}")
        st = []
        for h in stall_cols:
            try:
                st.append((float(r[idx[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        print("  stalls: " + " | ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:8]))


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    header = sys.argv[3] if len(sys.argv) > 3 else ""
    (launches if mode == "launches" else full)(path, header)
