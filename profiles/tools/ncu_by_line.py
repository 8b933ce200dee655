#!/usr/bin/env python
"""Aggregate an ncu `--page source --csv` SASS export per CUDA source line.

usage: ncu_by_line.py <source.csv> <nvdisasm -g -c output> <mangled kernel substring> [top]
The SASS export carries no line numbers in CSV form, so instructions are aligned by index with the
nvdisasm listing of the same build (opcodes are cross-checked)."""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = rows[hi + 1:]
col = {n: i for i, n in enumerate(hdr)}

lines = open(sass_path).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
ins = []
cur = None
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2).strip()))
assert len(ins) >= len(body), (len(ins), len(body))
agg = defaultdict(lambda: defaultdict(float))
keys = ["Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_barrier", "stall_long_sb", "stall_short_sb",
        "stall_mio", "stall_lg", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_membar", "stall_sleep"]
bad = 0
for (loc, op), r in zip(ins, body):
    sop = r[col["Source"]].strip().rstrip(";").strip()
    if sop.split()[0:1] != op.split()[0:1] and not (sop.startswith("@") and op.startswith("@")):
        bad += 1
    for k in keys:
        if k in col:
            try:
                agg[loc][k] += float(r[col[k]])
            except ValueError:
                pass
tot = defaultdict(float)
for loc in agg:
    for k in keys:
        tot[k] += agg[loc][k]
print(f"# {len(body)} instructions, opcode mismatches {bad}")
print("# totals:", {k: int(v) for k, v in tot.items() if v})
print(f"{'line':>22} {'inst%':>6} {'samp%':>6} {'shWave':>9} {'shIdeal':>9}  top stalls")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = sorted(((k, v) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print(f"{str(loc[0])[:14] + ':' + str(loc[1]) if loc else '?':>22} {100 * a['Instructions Executed'] / max(tot['Instructions Executed'], 1):6.2f} "
          f"{100 * a['# Samples'] / max(tot['# Samples'], 1):6.2f} {int(a['L1 Wavefronts Shared']):9d} {int(a['L1 Wavefronts Shared Ideal']):9d}  "
          + ", ".join(f"{k[6:]}={int(v)}" for k, v in st if v))
