#!/usr/bin/env python
"""Times the reference's OWN CUDA path (oracle/_ref/mmc_refcuda, the unmodified mmc_core.cu compiled for sm_100) on
the bench workloads -- the competitor named by BASELINE.json ("ref CUDA").  Prints one JSON line per workload."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import bench  # noqa: E402
import orc  # noqa: E402
from mmc_b200 import api  # noqa: E402


def main():
    names = sys.argv[1:] or ["sphshells:grid", "sphshells:elem", "cube60:elem"]
    for spec in names:
        name, method = spec.split(":")
        nph = float(os.environ.get("REF_PHOTONS", "1e7"))
        cfg, desc = bench.workload(name, method)
        st = cfg.get("srctype", 0)
        kw = dict(nphoton=int(nph), seed=cfg["seed"], srcpos=cfg["srcpos"], srcdir=cfg["srcdir"],
                  srctype=api.SRCTYPES.index(st) if isinstance(st, str) else st,
                  srcparam1=cfg.get("srcparam1", (0, 0, 0, 0)), srcparam2=cfg.get("srcparam2", (0, 0, 0, 0)),
                  tstart=cfg["tstart"], tend=cfg["tend"], tstep=cfg["tstep"], e0=cfg.get("e0", 0),
                  isreflect=cfg["isreflect"], method=api.METHODS[cfg["method"]], basisorder=0,
                  steps=cfg.get("steps", (1.0,))[0], evol=cfg.get("evol"))
        t0 = time.time()
        try:
            r = orc.run_ref(np.asarray(cfg["node"], np.float32), np.asarray(cfg["elem"], np.int32), np.asarray(cfg["elemprop"], np.int32),
                            np.asarray(cfg["prop"], np.float32)[1:], cuda=True, timeout=600, **kw)
            out = dict(workload=spec, nphoton=nph, kernel_ms=r.get("kernel_ms"), speed=r.get("speed"),
                       absorbed=r.get("absorbed_frac"), wall_s=time.time() - t0)
            if r.get("kernel_ms"):
                out["photons_per_ms_kernel"] = nph / r["kernel_ms"]
            tail = [l for l in r["log"].splitlines() if "threadph" in l or "kernel complete" in l or "speed" in l]
            out["log"] = tail[-4:]
        except Exception as e:  # noqa: BLE001
            out = dict(workload=spec, error=str(e)[-800:])
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
