#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report: DRAM bytes (read + write) of the first captured
launch of every kernel, with the workload, command and commit of the capture -- the ONLY source of the
`roofline.traffic` fields of bench.py (bench.ncu_traffic refuses a capture of another workload).

    python scripts/ncu_traffic.py --rep gpurun_out/prof.ncu-rep --workload "plane_stress 4096x2048 x1" \
        --capture profiles/r02_x.md --command "ncu --set full ... python bench.py ..."
"""
import argparse
import csv
import json
import re
import subprocess

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    """`void fe::k_spmv_stream<1, 0>(int, ...)` -> `k_spmv_stream<1,0>`"""
    name = name.split("(")[0].replace("void ", "").replace("fe::", "").strip()
    return re.sub(r"\s+", "", name)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", required=True, nargs="+", help="one or more .ncu-rep files of the same workload")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--capture", default="")
    ap.add_argument("--command", default="")
    ap.add_argument("--out", default="profiles/ncu_traffic.json")
    a = ap.parse_args()
    kernels = {}
    allrows = []
    for rep in a.rep:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        h, units = rows[0], rows[1]
        allrows += [(h, units, r) for r in rows[2:]]
    for h, units, r in allrows:
        ir, iw, it = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
        k = short(r[h.index("Kernel Name")])
        if k in kernels:
            continue
        rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
        wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
        kernels[k] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                      "duration_under_ncu": r[it] + " " + units[it]}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"workload": a.workload, "capture": a.capture, "command": a.command, "commit": commit, "kernels": kernels},
              open(a.out, "w"), indent=1)
    print("wrote", a.out, sorted(kernels))
