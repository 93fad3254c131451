#!/usr/bin/env python
"""Aggregate warp-stall samples of one kernel per CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{rx}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
lines = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] and r[0].isdigit():  # a source line row (its SASS rows follow and are already summed into it)
        i_s = hdr.index("# Samples")
        i_e = hdr.index("Instructions Executed")
        try:
            ns, ne = int(r[i_s]), int(r[i_e])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        a = lines.setdefault(key, [0, 0])
        a[0] += ns
        a[1] += ne
tot = sum(v[0] for v in lines.values())
print("total samples", tot)
for (f, ln, src), (ns, ne) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{ns:6d} {100.0 * ns / max(tot, 1):5.1f}%  ex={ne:>9}  {f}:{ln}  {src}")
