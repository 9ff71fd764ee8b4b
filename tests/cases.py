"""Shared small parity cases (mesh + config) used by the oracle pin tests, the golden-vector
generator (tools/make_golden.py) and the GPU parity tests."""
import numpy as np

from mmc_b200 import meshgen

PLUCKER, HAVEL, BADOUEL, BLBADOUEL, GRID = 0, 1, 2, 3, 4
FLUX, FLUENCE, ENERGY, JACOBIAN, WL, WP = 0, 1, 2, 3, 4, 5


def two_media_cube(n=20, step=2):
    g = np.arange(0, n + 1, step)
    node, elem = meshgen.gen_t5_mesh(g, g, g)
    c = meshgen.centroids(node, elem)
    et = np.ones(len(elem), np.int32)
    et[np.linalg.norm(c - n / 2.0, axis=1) < n / 4.0] = 2
    med = [(0.005, 1.0, 0.01, 1.37), (0.02, 2.0, 0.9, 1.5)]
    return node, elem, et, med


def wide_slab():
    """Small replaywide-style slab (examples/replaywide/createmesh.m:3-29): 20x20x10 mm of medium 1, air layers below
    (tets labelled -1: wide-field source candidates) and above (-2: wide-field detector)."""
    node, elem, et = meshgen.slab_with_wide_src_det(nx=20, ny=20, nz=10, step=2, gap=2.0)
    med = [(0.01, 1.0, 0.01, 1.37)]
    return node, elem, et, med


MESHES = {"cube": two_media_cube, "slab": wide_slab}


def case_mesh(name):
    return MESHES[CASES[name].get("mesh", "cube")]()


def _pattern(nx, ny, srcnum):
    rs = np.random.RandomState(3)
    return (rs.rand(ny * nx, srcnum) > 0.4).astype(np.float32).reshape(-1)


BASE = dict(nphoton=3000, seed=1648335518, srcpos=(10.1, 10.2, 0.0), srcdir=(0.0, 0.0, 1.0),
            tstart=0.0, tend=5e-9, tstep=5e-10)

# name -> overrides.  "exact": the oracle restates the arithmetic 1:1 (bit-exact vs the reference binary at
# one thread); Havel uses a true division where the reference uses rcpps+Newton (CPU-model dependent).
CASES = {
    "blb_elem_raw": dict(method=BLBADOUEL, isreflect=0, isnormalized=0, exact=True),
    "blb_elem_reflect": dict(method=BLBADOUEL, isreflect=1, exact=True),
    "blb_nodal_reflect": dict(method=BLBADOUEL, isreflect=1, basisorder=1, exact=True),
    "grid_1mm": dict(method=GRID, isreflect=1, steps=1.0, exact=True),
    "grid_halfmm": dict(method=GRID, isreflect=1, steps=0.5, exact=True),
    "havel_elem": dict(method=HAVEL, isreflect=1, exact=False),
    "havel_nodal": dict(method=HAVEL, isreflect=1, basisorder=1, exact=False),
    "plucker_elem": dict(method=PLUCKER, isreflect=1, exact=True),
    "plucker_nodal": dict(method=PLUCKER, isreflect=1, basisorder=1, exact=True),
    "blb_onegate": dict(method=BLBADOUEL, isreflect=1, tstep=5e-9, minenergy=0.3, exact=True),
    "blb_fluence": dict(method=BLBADOUEL, outputtype=FLUENCE, exact=True),
    "blb_energy": dict(method=BLBADOUEL, outputtype=ENERGY, exact=True),
    "blb_detectors": dict(method=BLBADOUEL, issavedet=1, issaveexit=1, ismomentum=1,
                          detpos=[(10, 8, 0, 1.5), (10, 12, 0, 1.5)], exact=True),
    "blb_isotropic": dict(method=BLBADOUEL, srctype=1, srcpos=(10.1, 10.2, 6.3), exact=False),
    "blb_mirror": dict(method=BLBADOUEL, isreflect=3, exact=True),
    # wide-field sources / detector on the slab (BASELINE config C5 family)
    "planar_widedet": dict(mesh="slab", method=BLBADOUEL, isreflect=1, srctype=4, srcpos=(5.0, 5.0, -1.0),
                           srcparam1=(10.0, 0, 0, 0), srcparam2=(0, 10.0, 0, 0), issavedet=1, issaveexit=1, exact=True),
    "pattern_share2": dict(mesh="slab", method=BLBADOUEL, isreflect=1, srctype=5, srcpos=(4.0, 4.0, -1.0),
                           srcparam1=(12.0, 0, 0, 4), srcparam2=(0, 12.0, 0, 4), srcnum=2, srcpattern=_pattern(4, 4, 2), exact=True),
    "disk_grid": dict(mesh="slab", method=GRID, steps=1.0, isreflect=1, srctype=8, srcpos=(10.0, 10.0, -1.0),
                      srcparam1=(3.0, 0, 0, 0), exact=True),
    "planar_havel_nodal": dict(mesh="slab", method=HAVEL, basisorder=1, isreflect=1, srctype=4, srcpos=(5.0, 5.0, -1.0),
                               srcparam1=(10.0, 0, 0, 0), srcparam2=(0, 10.0, 0, 0), exact=False),
    "blb_dref": dict(method=BLBADOUEL, isreflect=1, issaveref=1, exact=True),
    # detected-photon records from the Havel / Plucker kernels (element-wise and nodal kernel variants with detector columns)
    "havel_elem_det": dict(method=HAVEL, isreflect=1, issavedet=1, issaveexit=1, ismomentum=1,
                           detpos=[(10, 8, 0, 1.5), (10, 12, 0, 1.5)], exact=False),
    "plucker_nodal_det": dict(method=PLUCKER, isreflect=1, basisorder=1, issavedet=1, issaveexit=1,
                              detpos=[(10, 8, 0, 1.5), (10, 12, 0, 1.5)], exact=True),
}


def case_kwargs(name):
    kw = dict(BASE)
    kw.update(CASES[name])
    kw.pop("exact")
    kw.pop("mesh", None)
    return kw
