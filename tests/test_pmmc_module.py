"""The reference's Python front end on this engine: `import pmmc` (integration/pmmc, _pmmc built from integration/pmmc_b200.cpp by
pybind11 over include/mmc_b200.h) with the module functions of src/pmmc.cpp:1447-1462 and its output dictionary (:1085-1340).
The GPU tests run the reference's own regression script, pmmc/example/test_mesh_adjoint.py, restated here line for line in what it
configures and asserts (the GPU box has no /root/reference), and compare pmmc.run with the ctypes host on the same problems."""
import os
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))
pmmc = pytest.importorskip("pmmc")


def test_module_surface_and_error_convention():
    assert callable(pmmc.run) and callable(pmmc.gpuinfo) and "mmc_b200" in pmmc.version()
    assert isinstance(pmmc.gpuinfo(), list)
    node, elem, et, med = cases.two_media_cube()
    cfg = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), nphoton=100, srcpos=[10.1, 10.2, 0], srcdir=[0, 0, 1],
               tstart=0, tend=5e-9, tstep=5e-10, method="elem")
    with pytest.raises(ValueError, match="3 columns"):
        pmmc.run(dict(cfg, node=node[:, :2]))
    with pytest.raises(ValueError, match="source type"):
        pmmc.run(dict(cfg, srctype="laser"))
    with pytest.raises(RuntimeError, match=r"MMC ERROR\(-?\d+\):.*unitary"):                 # mcx_error -> exception, src/mmc_utils.c:1426-1442
        pmmc.run(dict(cfg, srcdir=[0, 0, 2]))
    with pytest.raises(RuntimeError, match="only valid in the reply mode"):
        pmmc.run(dict(cfg, outputtype="jacobian"))
    if not pmmc.gpuinfo():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pmmc.run(cfg)


def _example_mesh(side=10.0, n_per_side=4):        # pmmc/example/test_mesh_adjoint.py:20-56
    lin = np.linspace(0, side, n_per_side)
    X, Y, Z = np.meshgrid(lin, lin, lin, indexing="ij")
    node = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    n = n_per_side
    idx = lambda i, j, k: i * n * n + j * n + k      # noqa: E731
    elems = []
    for i in range(n - 1):
        for j in range(n - 1):
            for k in range(n - 1):
                v = [idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k), idx(i, j, k + 1), idx(i + 1, j, k + 1),
                     idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)]
                elems += [[v[0], v[1], v[3], v[4]], [v[1], v[3], v[4], v[6]], [v[1], v[2], v[3], v[6]], [v[3], v[4], v[6], v[7]], [v[1], v[4], v[5], v[6]]]
    return node, np.asarray(elems, dtype=np.int32) + 1


@pytest.mark.gpu
@pytest.mark.parametrize("adjointmode", [0, 1])
def test_reference_example_mesh_adjoint(adjointmode):
    """pmmc/example/test_mesh_adjoint.py:59-108: cfg of base_cfg(), assertions of run_and_check()."""
    node, elem = _example_mesh()
    cfg = {"nphoton": 50000, "seed": 17182818, "node": node, "elem": elem, "elemprop": np.ones(elem.shape[0], dtype=np.int32),
           "srcpos": np.array([5.0, 5.0, 0.0], dtype=np.float32), "srcdir": np.array([0.0, 0.0, 1.0], dtype=np.float32),
           "tstart": 0.0, "tstep": 5e-9, "tend": 5e-9, "prop": np.array([[0.0, 0.0, 1.0, 1.0], [0.005, 1.0, 0.01, 1.37]], dtype=np.float32),
           "detpos": np.array([[2.5, 5.0, 0.0, 1.0], [7.5, 5.0, 0.0, 1.0]], dtype=np.float32),
           "detdir": np.array([[0.0, 0.0, -1.0, 0.0], [0.0, 0.0, -1.0, 0.0]], dtype=np.float32),
           "outputtype": "adjoint", "method": "elem", "basisorder": 1, "isnormalized": 1, "isreflect": 0, "e0": 1, "adjointmode": adjointmode}
    out = pmmc.run(cfg)
    assert "flux" in out and "jmua" in out
    Ns, Nd, nn = 1, 2, node.shape[0]
    assert out["jmua"].shape == (nn, Ns * Nd)
    assert out["flux"].shape == (nn, 1, Ns + Nd) and out["flux"].flags.f_contiguous         # [datalen, maxgate, nsrcslots], src/pmmc.cpp:1203-1207
    # The script's detdir (0, 0, -1) is the OUTWARD normal of the z = 0 face its detectors sit on, and a detector slot is launched along
    # detdir as given (src/pmmc.cpp:1009-1012, src/mmc_core.cl:1467-1470): every adjoint photon leaves at once, the detector slots of
    # 'flux' stay empty and J_mua is identically zero -- here as in the reference's kernel.  Its last assertion (nonzero > 0) holds with
    # the inward normal, which is what the rest of this test uses.
    assert np.count_nonzero(out["flux"][:, :, Ns:]) == 0 and np.count_nonzero(out["jmua"]) == 0
    assert np.count_nonzero(out["flux"][:, :, 0]) > 0
    cfg["detdir"] = np.array([[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 1.0, 0.0]], dtype=np.float32)
    out = pmmc.run(cfg)
    assert np.count_nonzero(out["jmua"]) > 0
    assert np.all(out["jmua"] <= 0) and out["jmua"].min() < 0                              # J_mua = -phi_s phi_d (x volume weights)
    assert out["jmua"][:, 0].min() < 0 and out["jmua"][:, 1].min() < 0                      # both source-detector pairs


@pytest.mark.gpu
def test_pmmc_run_equals_the_ctypes_host():
    """Same engine behind both front ends: identical problem, static schedule and seeds => identical raw numbers, pmmc's layouts."""
    import mmc_b200 as mmc
    from test_gpu_parity import _cfg
    node, elem, et, med = cases.two_media_cube()
    kw = cases.case_kwargs("blb_detectors")
    kw.update(nphoton=50000, issaveseed=1, schedule=1, hotcache=-1)
    a = mmc.run(_cfg(node, elem, et, med, **kw))
    cfg = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), method="elem", basisorder=0,
               **{k: v for k, v in kw.items() if k not in ("method", "schedule", "hotcache")})
    cfg["detpos"] = np.asarray(cfg["detpos"], np.float32)
    b = pmmc.run(cfg)
    flux = b["flux"]
    assert flux.shape == (len(elem), 10) and flux.flags.f_contiguous
    assert b["detp"].shape[0] == a["detp"].shape[1] and b["seeds"].shape[0] == 16
    assert b["stat"]["energytot"] == a["energytot"][0] == kw["nphoton"]
    assert abs(b["stat"]["energyabs"] / a["energyabs"][0] - 1) < 0.02
    assert abs(b["detp"].shape[1] - len(a["detp"])) < 6 * np.sqrt(len(a["detp"])) + 5
    lit = a["raw"][..., 0].sum(axis=0) > 0.05 * a["raw"][..., 0].sum(axis=0).max()
    rel = np.abs(flux.sum(axis=1)[lit] - a["raw"][..., 0].sum(axis=0)[lit]) / a["raw"][..., 0].sum(axis=0)[lit]
    assert np.median(rel) < 0.08
    # grid output: [nx, ny, nz, maxgate]
    g = pmmc.run(dict(cfg, method="grid", steps=[1.0, 1.0, 1.0], issavedet=0, issaveseed=0, nphoton=20000))
    assert g["flux"].shape == (21, 21, 21, 10)
    r = mmc.run(dict(_cfg(node, elem, et, med, **dict(kw, nphoton=20000, issavedet=0, issaveseed=0, ismomentum=0, issaveexit=0)), method="grid", steps=(1.0, 1.0, 1.0)))
    assert abs(g["flux"].sum() / r["flux"].sum() - 1) < 0.05
    # the voxel under the source is the brightest in both
    assert np.unravel_index(np.argmax(g["flux"].sum(axis=3)), (21, 21, 21)) == np.unravel_index(np.argmax(r["flux"].sum(axis=3)), (21, 21, 21))
