#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_gemm_r1.ncu-rep profiles/r1_gemm_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # the tensor-pipe figure /opt/skills/guides/B200_PROFILING.md greps; the TPC.TriageCompute "realtime" variant
        # below it reads about 0.4-0.8x of it on the same launch (it was the one round 1's summaries quoted)
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += t
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {n} launches, {tot:.1f} us total\n")
        f.write("# per-launch times are cold-cache/serialised: compare SHARES, not absolutes\n")
        f.write(f"{'total_us':>12} {'n':>4} {'avg_us':>10} {'share':>7}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:12.1f} {v[0]:4d} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%  {k}\n")
    print(open(dst).read())


EXTRA = ("pipe_tensor", "xbar2l1tex", "lts__t_bytes", "lts__t_sectors_srcunit_tex", "dram__throughput",
         "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled")


def full(src, dst):
    """Summary of the judged metrics + the raw per-kernel CSV next to it (dst with .csv), so the summary can be re-cut."""
    if src.endswith(".csv"):      # re-cut a summary from a raw CSV kept under profiles/
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        with open(re.sub(r"\.txt$", "", dst) + ".raw.csv", "w") as f:
            f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = list(KEYS) + sorted(h for h in hdr if h not in KEYS and any(e in h for e in EXTRA))
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on : {src}\n")
        for r in rows[2:]:
            f.write("\n" + r[idx["Kernel Name"]][:150] + "\n")
            for k in keys:
                if k in idx:
                    f.write(f"  {k:95s} {r[idx[k]]:>16} {units[idx[k]]}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
