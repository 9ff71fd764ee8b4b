#!/usr/bin/env python
"""Per-gate sums of the C2 workload (1e8 photons) against the reference-CPU fixture tests/golden/ref_c2_1e8.npz: ratio ours / reference per gate.
usage: [MMCB_LIB=...] python tools/gate_trend.py [nphoton]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmc_b200 as mmc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
z = np.load(os.path.join(ROOT, "tests", "golden", "ref_c2_1e8.npz"))
m = np.load(os.path.join(ROOT, "tests", "golden", "sphshells_mesh.npz"))
g = mmc.run(dict(node=m["node"], elem=m["elem"], elemprop=m["etype"], prop=np.vstack([[0, 0, 1, 1], m["prop"]]), evol=m["evol"],
                 method=os.environ.get("METHOD", "grid"), e0=4916, steps=(1.0, 1.0, 1.0), nphoton=n, seed=29012392, srcpos=(30.0, 30.1, 0.0), srcdir=(0, 0, 1),
                 tstart=0.0, tend=5e-9, tstep=5e-10, isreflect=1, basisorder=0))
gs = g["raw"][..., 0].sum(axis=1) * (1e8 / n)
print(os.environ.get("MMCB_LIB", "product"), "kernel_ms %.1f" % g["kernel_ms"], "ratio per gate:", np.round(gs / z["gatesum"], 5).tolist())
if os.environ.get("REFCUDA"):
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import cases, orc
    r = orc.run_ref(m["node"], m["elem"], m["etype"], m["prop"], cuda=True, timeout=900, e0=4916, evol=m["evol"], nphoton=n, seed=1648335518,
                    srcpos=(30.0, 30.1, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=5e-10, isreflect=1, method=cases.GRID if os.environ.get("METHOD", "grid") == "grid" else cases.BLBADOUEL, basisorder=0, steps=1.0)
    rs = r["field_flat"].reshape(10, -1).sum(axis=1) * (1e8 / n)
    if os.environ.get("METHOD", "grid") == "grid":
        print("reference CUDA kernel, ratio per gate:", np.round(rs / z["gatesum"], 5).tolist(), "absorbed", r["absorbed_frac"])
    print("ours / reference CUDA:", np.round(gs / rs, 5).tolist())
