#!/usr/bin/env python
"""Condense ncu output into the markdown kept under profiles/.

    python scripts/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep --out profiles/X.md
"""
import argparse
import collections
import csv
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/tex throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def launches_table(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f"| `{name[:80]}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {t / total:.3f} |")
    return "\n".join(out)


def rep_table(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    seen, out = set(), []
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---:|---|")
        for key, label in KEYS:
            if key in h:
                i = h.index(key)
                out.append(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
    return "\n".join(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--out", required=True)
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    parts = [f"# {a.title}\n", a.note, ""]
    if a.launches:
        parts += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
                  "compare SHARES)\n", launches_table(a.launches), ""]
    if a.rep:
        parts += ["## Per-kernel detail (`ncu --set full --clock-control none`, first captured launch of each kernel)",
                  rep_table(a.rep)]
    open(a.out, "w").write("\n".join(parts) + "\n")
    print("wrote", a.out)
