#!/usr/bin/env python
"""Dynamic instruction counts of one kernel per SECTION of its top-level source file: joins the per-SASS-instruction counters of
an ncu report (--page source, sass view) with nvdisasm's inline-aware line table (-gi), so that inlined helpers (rand01,
next_scatter, ...) are charged to the line of the kernel body that called them.
usage: ncu_sections.py report.ncu-rep object.o kernel-mangled-substring topfile 'name:lo-hi,name:lo-hi,...'"""
import csv, os, re, subprocess, sys, tempfile
rep, obj, kname, topfile, spec = sys.argv[1:6]
secs = []
for s in spec.split(","):
    n, r = s.split(":")
    lo, hi = r.split("-")
    secs.append((n, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# address -> top-level line
addr2line, cur, infn = {}, None, False
for l in dis:
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        top = [c for c in chain if c[0].endswith(topfile)]
        cur = top[-1][1] if top else None
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
tot = {}
base = None
for r in rows:
    if r and "Address" in r and "Source" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        a = int(d["Address"], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    line = addr2line.get(a - base)
    name = "other"
    if line is not None:
        for n, lo, hi in secs:
            if lo <= line <= hi:
                name = n
                break
    t = tot.setdefault(name, [0, 0, 0])
    t[0] += int(d.get("# Samples") or 0)
    t[1] += int(d.get("Instructions Executed") or 0)
    t[2] += int(d.get("Thread Instructions Executed") or 0)
S = sum(t[0] for t in tot.values()) or 1
I = sum(t[1] for t in tot.values()) or 1
print("%-12s %8s %8s %6s %16s" % ("section", "samp%", "inst%", "act", "warp instr"))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-12s %8.2f %8.2f %6.1f %16d" % (n, 100 * t[0] / S, 100 * t[1] / I, t[2] / max(t[1], 1), t[1]))
