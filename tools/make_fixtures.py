#!/usr/bin/env python
"""Convert the two tetrahedral meshes the reference SHIPS as example data (text .dat files) into compact
npz fixtures so that the BASELINE configs C2 (sphshells) and C3 (skinvessel) can run on the GPU box,
where /root/reference does not exist.  Mesh data, not source code.

    examples/sphshells/{node,elem,prop}_dmmc_sphshells.dat  + dmmc_sphshells.json  -> tests/golden/sphshells_mesh.npz
    examples/skinvessel/{node,elem,prop}_dmmc_skinvessel.dat + dmmc_skinvessel.json -> tests/golden/skinvessel_mesh.npz
"""
import json
import os

import numpy as np

REF = os.environ.get("MMC_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(dirname, tag):
    d = os.path.join(REF, "examples", dirname)
    node = np.loadtxt(os.path.join(d, "node_%s.dat" % tag), skiprows=1)[:, 1:4].astype(np.float32)
    el = np.loadtxt(os.path.join(d, "elem_%s.dat" % tag), skiprows=1, dtype=np.int64)
    elem, etype = el[:, 1:5].astype(np.int32), el[:, 5].astype(np.int32)
    prop = np.loadtxt(os.path.join(d, "prop_%s.dat" % tag), skiprows=1)[:, 1:5].astype(np.float32)
    cfg = json.load(open(os.path.join(d, "%s.json" % tag)))
    evol = np.loadtxt(os.path.join(d, "velem_%s.dat" % tag), skiprows=1)[:, 1].astype(np.float32)
    return node, elem, etype, prop, cfg, evol


for dirname, tag in (("sphshells", "dmmc_sphshells"), ("skinvessel", "dmmc_skinvessel")):
    node, elem, etype, prop, cfg, evol = load(dirname, tag)
    out = os.path.join(ROOT, "tests", "golden", "%s_mesh.npz" % dirname)
    np.savez_compressed(out, node=node, elem=elem, etype=etype, prop=prop, evol=evol,
                        cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8))
    print(out, node.shape, elem.shape, np.unique(etype), prop, os.path.getsize(out))
