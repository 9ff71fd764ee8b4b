#!/usr/bin/env python
"""Reference-CPU fixtures at north_star's photon count for the tracers that exist only in the reference's CPU file:
BASELINE config C1 (cube60, mua 0.005 mus 1 g 0.01 n 1.37, pencil, 50 gates, -b 0) with Havel (-M h) and Plucker (-M p),
nodal output (-C 1), 1e8 photons, through the UNMODIFIED reference binary oracle/_ref/mmc_ref with all host threads
(about 13 minutes per tracer on 8 cores).  Kept per tracer: the CW nodal fluence (sum over the gates, float32), the per-gate sums,
the absorbed fraction and the ray-tet count -> tests/golden/ref_c1_1e8.npz (about 250 KB).  /root/reference is only needed to
build the binary; the fixture travels.      usage: python tools/make_golden_1e8.py [nphoton]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import orc  # noqa: E402
from mmc_b200 import meshgen  # noqa: E402


def c2(nph):
    """BASELINE config C2 (the headline workload): shipped dmmc_sphshells mesh, index mismatch + reflection, dual-grid output (61^3 voxels of
    1 mm), 10 gates, through the reference CPU binary (-M g) -> tests/golden/ref_c2_1e8.npz: CW fluence per voxel as float16 of the value
    normalised to its maximum where it exceeds 1e-4 of it (the test reads 1e-3 and above), per-gate sums, absorbed fraction.
    usage: python tools/make_golden_1e8.py c2 [nphoton]"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "sphshells_mesh.npz"))
    t0 = time.time()
    r = orc.run_ref(z["node"], z["elem"], z["etype"], z["prop"], nthread=os.cpu_count() or 1, timeout=4 * 3600, nphoton=nph, seed=1648335518,
                    srcpos=(30.0, 30.1, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=5e-10, isreflect=1, method=cases.GRID, basisorder=0,
                    steps=1.0, e0=4916, evol=z["evol"])
    f = r["field_flat"].reshape(10, -1)
    cw = f.sum(axis=0)
    keep = np.flatnonzero(cw > 1e-4 * cw.max()).astype(np.uint32)
    meta = dict(nphoton=nph, absorbed_frac=r["absorbed_frac"], raytet=r["raytet"], speed=r.get("speed"), wall_s=time.time() - t0,
                threads=os.cpu_count(), nvox=int(cw.size), cwmax=float(cw.max()))
    print("c2", meta, flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_c2_1e8.npz"), idx=keep, cw=cw[keep].astype(np.float32),
                        gatesum=f.sum(axis=1), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        return c2(int(float(sys.argv[2])) if len(sys.argv) > 2 else 100000000)
    nph = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
    node, elem, et = meshgen.cube60()
    med = [(0.005, 1.0, 0.01, 1.37)]
    out, meta = {}, {}
    for name, method in (("havel", cases.HAVEL), ("plucker", cases.PLUCKER)):
        t0 = time.time()
        r = orc.run_ref(node, elem, et, med, nthread=os.cpu_count() or 1, timeout=7200, nphoton=nph, seed=1648335518,
                        srcpos=(30.1, 30.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=1e-10, isreflect=0,
                        method=method, basisorder=1)
        f = r["field_flat"].reshape(50, -1)
        assert f.shape[1] == len(node)
        out[name + "/cw"] = f.sum(axis=0).astype(np.float32)
        out[name + "/gatesum"] = f.sum(axis=1)
        meta[name] = dict(nphoton=nph, absorbed_frac=r["absorbed_frac"], raytet=r["raytet"], normalizer=r.get("normalizer"),
                          speed=r.get("speed"), wall_s=time.time() - t0, threads=os.cpu_count())
        print(name, meta[name], flush=True)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_c1_1e8.npz"), **out)


if __name__ == "__main__":
    main()
