#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep taken with --import-source on:
  python tools/ncu_lines.py <report.ncu-rep> <kernel-name> [min_share_percent]"""
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                              text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
ci = {}
for k, n in enumerate(H):
    ci.setdefault(n, k)
lines = []
for r in rows[hdr + 1:]:
    if len(r) < len(H) or not r[0].isdigit():
        continue
    def g(n):
        m = re.match(r"-?\d+", r[ci[n]] or "")
        return int(m.group()) if m else 0
    if len(r) != len(H):                                 # a source line with quotes in it (inline asm) splits badly: skip
        continue
    lines.append((int(r[0]), r[1], g("# Samples"), g("Instructions Executed"), g("stall_long_sb"), g("stall_short_sb"), g("stall_wait"),
                  g("stall_math"), g("stall_barrier"), g("stall_mio"), g("stall_lg")))
ts = sum(l[2] for l in lines) or 1
ti = sum(l[3] for l in lines) or 1
print("%s: %d stall samples, %d warp instructions" % (kern, ts, ti))
print(" line  samples  instr | long_sb short_sb wait math barrier mio lg | source")
for l in lines:
    if 100.0 * l[2] / ts >= thr or 100.0 * l[3] / ti >= thr:
        print("%5d  %5.1f%%  %5.1f%% | %5d %5d %5d %5d %5d %5d %5d | %s" % (l[0], 100.0 * l[2] / ts, 100.0 * l[3] / ti, l[4], l[5], l[6], l[7], l[8], l[9], l[10], l[1].strip()[:140]))
