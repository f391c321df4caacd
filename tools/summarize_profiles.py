#!/usr/bin/env python
"""Turns the raw ncu output in gpurun_out/ into the small, committed summaries under profiles/.

    python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches_X.csv] [--rep name=gpurun_out/prof.ncu-rep ...]

Writes profiles/<tag>_launches.csv (per-kernel totals and shares of the profiled command),
profiles/<tag>_<name>_metrics.csv (selected raw metrics of each captured launch) and updates
profiles/traffic.json (dram bytes per launch of each captured kernel, read by bench.py's roofline.traffic).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
)
STALL = "smsp__average_warps_issue_stalled_"


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        if "cub::" in name:
            name = "cub::" + name.split("cub::")[1].split("<")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    tot = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "mean_ms", "share_of_profiled_command"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1]:.3f}", f"{v[1] / v[0]:.4f}", f"{v[1] / tot:.4f}"])
    return agg


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches_ = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio"))]
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{j}:{r[ni].split('(')[0]}" for j, r in enumerate(launches_)])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in launches_])
    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in launches_:
        name = r[ni].split("(")[0].split("::")[-1]
        b = float(r[ri].replace(",", "")) * mult[units[ri]] + float(r[wi].replace(",", "")) * mult[units[wi]]
        traffic.setdefault(name, []).append(b)
    return {k: sum(v) / len(v) for k, v in traffic.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    a = ap.parse_args()
    pdir = os.path.join(ROOT, "profiles")
    os.makedirs(pdir, exist_ok=True)
    if a.launches:
        agg = launches(a.launches, os.path.join(pdir, f"{a.tag}_launches.csv"))
        print({k: round(v[1], 2) for k, v in agg.items()})
    tpath = os.path.join(pdir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for spec in a.rep:
        name, path = spec.split("=", 1)
        t = rep(path, os.path.join(pdir, f"{a.tag}_{name}_metrics.csv"))
        traffic.update(t)
        print(name, t)
    if a.rep:
        json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
