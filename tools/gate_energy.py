#!/usr/bin/env python
"""Absorbed energy per time gate (-O E, summed over the volume) on the C2 workload: this engine (element and dual-grid kernels) against the
reference CUDA kernel in both modes.  A ratio that drifts with the gate index is a cumulative error in the walk (clock, weight, loss of photons).
usage: python tools/gate_energy.py [nphoton]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import cases, orc
import mmc_b200 as mmc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
m = np.load(os.path.join(ROOT, "tests", "golden", "sphshells_mesh.npz"))
kw = dict(nphoton=n, srcpos=(30.0, 30.1, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=5e-10, isreflect=1, basisorder=0, e0=4916)
res = {}
extra = {}
if os.environ.get("SCHEDULE"):
    extra["schedule"] = int(os.environ["SCHEDULE"])
for mode in os.environ.get("MODES", "elem grid").split():
    g = mmc.run(dict(node=m["node"], elem=m["elem"], elemprop=m["etype"], prop=np.vstack([[0, 0, 1, 1], m["prop"]]), evol=m["evol"], method=mode,
                     steps=(1.0, 1.0, 1.0), seed=29012392, outputtype="energy", isnormalized=0, **kw, **extra))
    res["ours " + mode] = g["raw"][..., 0].sum(axis=1)
    if mode not in ("elem", "grid") or os.environ.get("NOREF"):
        continue
    r = orc.run_ref(m["node"], m["elem"], m["etype"], m["prop"], cuda=True, timeout=900, evol=m["evol"], seed=1648335518, steps=1.0, outputtype=cases.ENERGY,
                    isnormalized=0, method=cases.GRID if mode == "grid" else cases.BLBADOUEL, **kw)
    res["refcuda " + mode] = r["field_flat"].reshape(10, -1).sum(axis=1)
base = res.get("refcuda elem")
if base is None:
    r = orc.run_ref(m["node"], m["elem"], m["etype"], m["prop"], cuda=True, timeout=900, evol=m["evol"], seed=1648335518, steps=1.0, outputtype=cases.ENERGY,
                    isnormalized=0, method=cases.BLBADOUEL, **kw)
    base = r["field_flat"].reshape(10, -1).sum(axis=1)
for k, v in res.items():
    print("%-14s" % k, "sum %.6e" % v.sum(), "ratio to %s per gate:" % "refcuda", np.round(v / base, 5).tolist(), flush=True)
