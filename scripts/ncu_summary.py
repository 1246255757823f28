#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
  python scripts/ncu_summary.py <tag> [launches.csv] [prof.ncu-rep]
Writes profiles/<tag>_launches.md (per-kernel totals and SHARES from the gpu__time_duration launch list - cold-cache,
serialised, so only the shares are meaningful) and profiles/<tag>_full.md (key metrics of the `--set full` capture)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][:60]
        v = float(row["Metric Value"].replace(",", ""))
        if row.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache + serialised: read the SHARES)", "",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.1f} | {100 * t / tot:.1f}% |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    # one control step in launch order: from one physics launch to the next
    rows = list(csv.DictReader(lines))
    ph = [i for i, r in enumerate(rows) if "physics_soa_kernel" in r["Kernel Name"]]
    if len(ph) >= 3:
        a, b = ph[-3], ph[-2]
        o2 = [f"# {tag}: one control step in launch order (eager, serialised; ncu gpu__time_duration, cold cache)", "",
              "| # | kernel | grid | block | us |", "|---|---|---|---|---|"]
        tot2 = 0.0
        for i, r in enumerate(rows[a:b]):
            v = float(r["Metric Value"].replace(",", "")) * (1e3 if r.get("Metric Unit", "ns") in ("us", "usecond") else 1)
            tot2 += v
            o2.append(f"| {i} | `{r['Kernel Name'].split('(')[0][:70]}` | {r.get('Grid Size', '')} | {r.get('Block Size', '')} | {v / 1e3:.1f} |")
        o2.append(f"\nsum {tot2 / 1e3:.1f} us")
        open(os.path.join(ROOT, "profiles", f"{tag}_step.md"), "w").write("\n".join(o2) + "\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def full(tag, rep):
    """rep: a .ncu-rep file, or the CSV that `ncu -i rep --page raw --csv` wrote on the GPU box."""
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {w: hdr.index(w) for w in WANT if w in hdr}
    kn = hdr.index("Kernel Name")
    best = {}
    gs = hdr.index("Grid Size") if "Grid Size" in hdr else None
    for r in rows[2:]:
        name = r[kn].split("(")[0][:60]
        d = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        if name not in best or d > best[name][0]:
            best[name] = (d, r)
    out = [f"# {tag}: ncu --set full (longest captured launch of each kernel; --clock-control none)", ""]
    for name, (d, r) in best.items():
        out += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for w, i in idx.items():
            out.append(f"| {w} | {r[i]} | {units[i]} |")
        rd, wr = idx.get("dram__bytes_read.sum"), idx.get("dram__bytes_write.sum")
        out.append("")
    open(os.path.join(ROOT, "profiles", f"{tag}_full.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
        launches(tag, sys.argv[2])
    if len(sys.argv) > 3 and os.path.exists(sys.argv[3]):
        full(tag, sys.argv[3])
