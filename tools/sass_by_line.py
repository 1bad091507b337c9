#!/usr/bin/env python
"""Attribute executed warp instructions of one kernel to CUDA source lines.
  ncu -i X.ncu-rep --page source --print-source sass --csv --kernel-name regex:K --launch-count 1 > k_sass.csv
  python tools/sass_by_line.py k_sass.csv <cubin> <mangled-kernel-substring> [top]
Joins ncu's per-SASS-instruction 'Instructions Executed' with `nvdisasm -g` line info by instruction order."""
import csv
import re
import subprocess
import sys

sass_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ie, src = hdr.index("Instructions Executed"), hdr.index("Source")
execd = [(r[src].strip(), float(r[ie] or 0)) for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") and ".text." in l:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(1).strip()))
print(f"ncu SASS instructions: {len(execd)}, nvdisasm instructions: {len(lines)}")
n = min(len(execd), len(lines))
agg = {}
for (loc, _), (_, c) in zip(lines[:n], execd[:n]):
    agg[loc] = agg.get(loc, 0.0) + c
tot = sum(agg.values())
srcs = {}
for (f, ln), c in sorted(agg.items(), key=lambda x: -x[1])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(f"primal_ppo_b200/csrc/{f}").read().splitlines()
        except Exception:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100 * c / tot:5.1f}%  {f}:{ln:<4d} {text}")
print(f"total warp instructions {tot:.3g}")
