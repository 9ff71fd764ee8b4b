#!/usr/bin/env python
"""Exercises the streaming kernels around the photon kernel (mesh preparation, elem->node spreading, device normalisation, adjoint post-
kernels) on realistic sizes, for an ncu launch list: head atlas (335 713 tets) nodal one-call run, cube adjoint run (grid + mesh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench, cases, mmc_b200 as mmc
cfg, _ = bench.workload("headatlas", "elem")
mmc.run(dict(cfg, nphoton=200000, basisorder=1, issavedet=0, issaveexit=0))
cfg, _ = bench.workload("skinvessel", "grid")
mmc.run(dict(cfg, nphoton=200000))
node, elem, et, med = cases.two_media_cube(n=40, step=1)
base = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), nphoton=200000, srcpos=(20.1, 20.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9,
            tstep=5e-9, isreflect=1, detpos=[(20.3, 16.4, 0.0, 1.0), (21.7, 24.4, 40.0, 1.0)], detdir=[(0, 0, 1, 0), (0, 0, -1, 0)], outputtype="adjointmuad")
mmc.run(dict(base, method="grid", steps=(1.0, 1.0, 1.0), basisorder=0))
mmc.run(dict(base, method="elem", basisorder=1))
print("ok")
