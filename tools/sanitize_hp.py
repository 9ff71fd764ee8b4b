#!/usr/bin/env python
"""Small Havel / Plucker runs (element-wise, nodal, with detector records, planar source) for `compute-sanitizer --tool memcheck`:
usage: compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_hp.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import cases  # noqa: E402
import mmc_b200  # noqa: E402
from test_gpu_parity import _cfg  # noqa: E402

for name in ("havel_elem", "havel_nodal", "plucker_elem", "plucker_nodal", "planar_havel_nodal", "havel_elem_det", "plucker_nodal_det", "blb_elem_reflect", "grid_1mm"):
    node, elem, et, med = cases.case_mesh(name)
    kw = cases.case_kwargs(name)
    kw["nphoton"] = 600000          # above the scout threshold: the count-mode scout and the hot-line cache run as well
    g = mmc_b200.run(_cfg(node, elem, et, med, **kw))
    print(name, "absorbed %.4f" % (g["energyabs"][0] / g["energytot"][0]), "raytet %.0f" % g["raytet"], flush=True)
