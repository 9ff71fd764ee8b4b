#!/usr/bin/env python
"""Fill profiles/ncu_traffic.json from one `ncu --set full` report of a bench run: per launch DRAM bytes, global reductions (thread
level: Thread Instructions Executed of the RED.* SASS rows), issue-slot utilisation, IPC, active threads per warp instruction.
usage: ncu_summary.py report.ncu-rep workload:method photons [note]"""
import csv, json, os, subprocess, sys
rep, key, photons = sys.argv[1], sys.argv[2], float(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
m, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
def num(k):      # bytes in bytes, times in ms
    if k not in m or m[k] in ("", "n/a"):
        return None
    return float(m[k].replace(",", "")) * SCALE.get(u.get(k, ""), 1.0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
r2 = list(csv.reader(src.splitlines()))
h2 = None
reds = atoms = 0
for r in r2:
    if r and "Address" in r and "Source" in r:
        h2 = r
        continue
    if h2 and len(r) == len(h2):
        d = dict(zip(h2, r))
        op = d["Source"].strip().split()
        op = [o for o in op if not o.startswith("@")]
        if op and op[0].startswith("REDG."):
            reds += int(d.get("Thread Instructions Executed") or 0)
        if op and op[0].startswith("ATOMS"):
            atoms += int(d.get("Thread Instructions Executed") or 0)
out = {"dram_bytes_per_launch": int((num("dram__bytes_read.sum") or 0) + (num("dram__bytes_write.sum") or 0)),
       "photons": int(photons), "kernel_ms": num("gpu__time_duration.sum"),
       "global_reds_per_photon": reds / photons, "shared_atomics_per_photon": atoms / photons,
       "issue_slots_busy": (num("sm__inst_issued.avg.pct_of_peak_sustained_active") or 0) / 100.0,
       "ipc": num("sm__inst_executed.avg.per_cycle_active"),
       "active_threads": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
       "l1_hit": (num("l1tex__t_sector_hit_rate.pct") or 0) / 100.0,
       "report": "%s (%s)" % (os.path.basename(rep), note or "ncu --set full --clock-control none")}
p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
j = json.load(open(p)) if os.path.exists(p) else {}
j[key] = out
json.dump(j, open(p, "w"), indent=1)
print(key, json.dumps(out))
