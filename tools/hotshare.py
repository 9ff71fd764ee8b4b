#!/usr/bin/env python
"""Prints the hot-line statistics of the pilot batch (MMCB_TRACE line) for the bench workloads."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MMCB_TRACE"] = "1"
import bench  # noqa: E402
import mmc_b200  # noqa: E402
for name, method in (("cube60", "elem"), ("cube60", "grid"), ("sphshells", "elem"), ("sphshells", "grid"), ("skinvessel", "grid"), ("headlike", "elem")):
    cfg, desc = bench.workload(name, method)
    cfg.update(nphoton=1000000)
    sys.stderr.write("== %s:%s\n" % (name, method))
    sys.stderr.flush()
    r = mmc_b200.run(cfg)
    sys.stderr.write("   kernel %.2f ms, %.1f steps/photon\n" % (r["kernel_ms"], r["raytet"] / 1e6))
