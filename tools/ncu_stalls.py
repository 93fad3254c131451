#!/usr/bin/env python
"""Aggregate the per-instruction stall sampling of one kernel in an .ncu-rep.
usage: python tools/ncu_stalls.py report.ncu-rep [kernel-regex] [launch-skip] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip",
                      skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:140] if rows and len(rows[0]) > 1 else "")
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
idx = {h: i for i, h in enumerate(hdr)}


def iv(x):
    try:
        return int(x)
    except ValueError:
        return 0


seen, uniq = set(), []
for r in data:
    if r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    uniq.append(r)
n = lambda r: iv(r[idx["# Samples"]])
tot = sum(n(r) for r in uniq)
print("total samples", tot, "instructions", len(uniq), "executed", sum(iv(r[idx["Instructions Executed"]]) for r in uniq))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in uniq:
    for s_ in stalls:
        agg[s_] = agg.get(s_, 0) + iv(r[idx[s_]])
print(", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for r in sorted(uniq, key=lambda r: -n(r))[:topn]:
    st = sorted([(iv(r[idx[s]]), s[6:]) for s in stalls], reverse=True)[:2]
    print(f"{n(r):6d} ex={r[idx['Instructions Executed']]:>9} {r[idx['Source']].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
