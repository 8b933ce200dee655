#!/usr/bin/env python3
"""Per-kernel DRAM traffic from an `ncu --set full` capture -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic).

usage: ncu_traffic.py <workload> <report.ncu-rep> [more reports ...] [-o profiles/ncu_traffic.json]

For every kernel class bench.py times (the KK_* classes of csrc/gbin_internal.h) the launches found in the reports are averaged:
dram__bytes_read.sum + dram__bytes_write.sum per launch, plus duration, issue utilisation and warp-instruction count so that the
summary in profiles/ can be regenerated from the same file.  The file records the git commit the capture was taken at; bench.py
leaves roofline.traffic null when the workload does not match.
"""
import csv
import json
import os
import subprocess
import sys

CLASSES = [  # (substring of the kernel name, class name in bench.py's per_kernel tables)
    ("group3_kernel", "skr_group"),
    ("skr_group_kernel", "skr_group"),
    ("skr_scan2_kernel", "skr_scan"),
    ("skr_scan_kernel", "skr_scan"),
    ("radix_scatter_kernel", "radix_scatter"),
    ("radix_hist_kernel", "radix_hist"),
    ("finalize3_kernel", "v3_span"),
    ("make_entries_kernel", "v3_entries"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6,
        "second": 1e3}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield {h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)}


def val(d, key):
    if key not in d or d[key][0] in ("", "n/a"):
        return None
    v, u = d[key]
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main():
    args = sys.argv[1:]
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ncu_traffic.json")
    if "-o" in args:
        i = args.index("-o")
        out_path = args[i + 1]
        del args[i:i + 2]
    workload, reports = args[0], args[1:]
    acc = {}
    for rep in reports:
        for d in rows_of(rep):
            name = d["Kernel Name"][0]
            cls = next((c for s, c in CLASSES if s in name), None)
            if cls is None:
                continue
            rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
            if rd is None or wr is None:
                continue
            a = acc.setdefault(cls, {"kernel": name.split("(")[0], "launches": 0, "dram": 0.0, "read": 0.0, "write": 0.0, "ms": 0.0, "inst": 0.0,
                                     "issue": 0.0, "report": os.path.basename(rep)})
            a["launches"] += 1
            a["dram"] += rd + wr
            a["read"] += rd
            a["write"] += wr
            a["ms"] += val(d, "gpu__time_duration.sum") or 0.0
            a["inst"] += val(d, "smsp__inst_executed.sum") or 0.0
            a["issue"] += val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0
    kernels = {}
    for cls, a in acc.items():
        n = a["launches"]
        kernels[cls] = {"kernel": a["kernel"], "launches_in_capture": n, "dram_bytes_per_launch": a["dram"] / n, "dram_read_bytes_per_launch": a["read"] / n,
                        "dram_write_bytes_per_launch": a["write"] / n, "ms_per_launch_under_ncu": a["ms"] / n, "warp_instructions_per_launch": a["inst"] / n,
                        "issue_active_pct": a["issue"] / n, "report": a["report"]}
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(out_path)).stdout.strip()
    except OSError:
        commit = ""
    json.dump({"workload": workload, "commit": commit, "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum",
               "kernels": kernels}, open(out_path, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
