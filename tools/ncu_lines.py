#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: samples, executed warp instructions, avg active threads.
usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr, r))
        # duplicated header name "Source": first = cuda line text
        lines.append((int(r[0]), r[1], int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), int(d["Thread Instructions Executed"] or 0),
                      int(d.get("L2 Theoretical Sectors Local") or 0), int(d.get("stall_long_sb") or 0), int(d.get("stall_lg") or 0)))
ts = sum(l[2] for l in lines) or 1
ti = sum(l[3] for l in lines) or 1
print("total samples %d, warp instr %d, thread instr %d, avg active %.1f" % (ts, ti, sum(l[4] for l in lines), sum(l[4] for l in lines) / ti))
print("%5s %6s %6s %5s %8s %6s %6s  %s" % ("line", "samp%", "inst%", "act", "localsec", "longsb", "lg", "source"))
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    print("%5d %6.2f %6.2f %5.1f %8d %6d %6d  %s" % (l[0], 100 * l[2] / ts, 100 * l[3] / ti, l[4] / max(l[3], 1), l[5], l[6], l[7], l[1].strip()[:110]))
