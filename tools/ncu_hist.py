#!/usr/bin/env python
"""Executed-instruction histogram and key metrics of one kernel from an .ncu-rep (no GPU needed)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2]
for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]:
    if m in hdr:
        i = hdr.index(m)
        print(f"{m} [{units[i]}] = {data[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
hist = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    try:
        n = int(r[ie])
    except ValueError:
        continue
    op = re.sub(r"^@!?U?P\w+\s+", "", r[ia].strip()).split()[0]
    o2 = op.split(".")[0]
    if op.startswith("IMAD.MOV") or op.startswith("MOV"):
        o2 = "MOV*"
    hist[o2] += n
    tot += n
print("executed warp instructions:", tot)
for k, v in hist.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f}%")
