#!/usr/bin/env python
"""Where the end-to-end time of the one-call path goes: Python marshalling vs the phases inside the C-ABI (MMCB_TRACE)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MMCB_TRACE"] = "1"
import ctypes as C  # noqa: E402
import bench  # noqa: E402
import mmc_b200  # noqa: E402
from mmc_b200 import api  # noqa: E402
name, method = (sys.argv[1:] + ["sphshells", "grid"])[:2]
cfg, desc = bench.workload(name, method)
cfg["nphoton"] = int(float(os.environ.get("PHOTONS", "1e7")))
mmc_b200.run(dict(cfg, nphoton=100000))          # context + module load
for rep in range(2):
    sys.stderr.write("== run %d\n" % rep)
    t0 = time.perf_counter()
    prob = api.Problem(cfg)
    t1 = time.perf_counter()
    sz = prob.sizes()
    t2 = time.perf_counter()
    buf = api._OutBuffers(prob, sz)
    t3 = time.perf_counter()
    api._check(api.lib().mmcb_run_simulation(C.byref(prob.cfg), C.byref(prob.mesh), prob.device, C.byref(buf.out)))
    t4 = time.perf_counter()
    r = buf.result(prob)
    t5 = time.perf_counter()
    sys.stderr.write("[py] Problem %.1f ms, sizes %.1f ms, buffers %.1f ms, run_simulation %.1f ms, result %.1f ms, total %.1f ms (kernel %.1f ms)\n" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3, r["kernel_ms"]))
