"""Mesh pre-processing on the device (mmc_b200/csrc/mmcb_prep.cu) against the host restatement of
mesh_getfacenb (src/mmc_highorder.cpp:124-159) and tracer_build (src/mmc_mesh.c:1572-1600) that stays in mmcb_host.cu as the
checker: face-neighbour tables, 96-byte tetrahedron records and centroids must be BIT-IDENTICAL."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
mmc = pytest.importorskip("mmc_b200")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _meshes():
    node, elem, et, med = cases.two_media_cube()
    yield "two_media_cube", dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), srcpos=(10.1, 10.2, 0.0))
    node, elem, et, med = cases.wide_slab()
    yield "slab_void_layers", dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), srcpos=(5.0, 5.0, -1.0),
                                   srctype="planar", srcparam1=(10.0, 0, 0, 0), srcparam2=(0, 10.0, 0, 0))
    z = np.load(os.path.join(GOLD, "sphshells_mesh.npz"))
    yield "sphshells", dict(node=z["node"], elem=z["elem"], elemprop=z["etype"], prop=np.vstack([[0, 0, 1, 1], z["prop"]]),
                            srcpos=(30.0, 30.1, 0.0), e0=4916)
    node, elem, et = mmc.meshgen.head_like()
    prop = [[0, 0, 1, 1], [0.019, 7.8, 0.89, 1.37], [0.019, 7.8, 0.89, 1.37], [0.004, 0.009, 0.89, 1.37], [0.02, 9.0, 0.89, 1.37],
            [0.08, 40.9, 0.84, 1.37]]
    yield "head_like_263k", dict(node=node, elem=elem, elemprop=et, prop=prop, srcpos=(42.0, 52.0, 91.0), srcdir=(0, 0, -1))


@pytest.mark.parametrize("name,cfg", list(_meshes()), ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("isreflect", [1, 0])
def test_device_tables_equal_host_tables(name, cfg, isreflect):
    cfg = dict(cfg, nphoton=1000, tstart=0.0, tend=5e-9, tstep=5e-9, method="elem", basisorder=0, isreflect=isreflect, srcdir=cfg.get("srcdir", (0, 0, 1)))
    with mmc.Session(cfg) as s:
        rec_d, cent_d, fnb_d = s.tables()
    os.environ["MMCB_HOST_PREP"] = "1"
    try:
        with mmc.Session(cfg) as s:
            rec_h, cent_h, fnb_h = s.tables()
    finally:
        del os.environ["MMCB_HOST_PREP"]
    assert np.array_equal(fnb_d, fnb_h), "face neighbours differ at %d entries" % (fnb_d != fnb_h).sum()
    assert (fnb_h < 0).sum() > 0 and (fnb_h == 0).sum() == 0            # exterior faces numbered -1..-nf
    assert np.array_equal(cent_d.view(np.uint32), cent_h.view(np.uint32))
    bad = np.nonzero((rec_d != rec_h).any(axis=1))[0]
    assert len(bad) == 0, "records differ for %d elements, first %d: %s vs %s" % (len(bad), bad[0], rec_d[bad[0]], rec_h[bad[0]])


def test_device_prep_gives_the_same_simulation():
    """same seeds, static schedule: the photon kernel sees identical tables, so the results are identical to the last bit of the
    energy tallies"""
    node, elem, et, med = cases.two_media_cube()
    cfg = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), srcpos=(10.1, 10.2, 0.0), srcdir=(0, 0, 1),
               nphoton=50000, tstart=0.0, tend=5e-9, tstep=5e-10, method="elem", basisorder=0, isreflect=1, schedule=1, hotcache=-1,
               isnormalized=0)
    a = mmc.run(cfg)
    os.environ["MMCB_HOST_PREP"] = "1"
    try:
        b = mmc.run(cfg)
    finally:
        del os.environ["MMCB_HOST_PREP"]
    assert a["raytet"] == b["raytet"]
    assert a["energyesc"][0] == b["energyesc"][0]
    np.testing.assert_allclose(a["raw"], b["raw"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", ["blb_elem_reflect", "blb_nodal_reflect", "grid_1mm", "havel_nodal", "plucker_elem", "blb_energy", "blb_fluence",
                                  "pattern_share2", "blb_dref"])
def test_device_normalisation_equals_the_numpy_restatement(name):
    """mesh_normalize (src/mmc_mesh.c:2154-2279) on the device (mmcb_post.cu: mmcb_norm_*) against the checker's numpy restatement
    (oracle/normalize_np.py) on identical raw volumes: two runs with the static schedule and the same seeds deposit the same sums, one
    is normalised by the library, the other by the checker with the oracle's mesh tables.  The reductions run in a different order:
    rtol 1e-9."""
    import normalize_np
    import orc
    import test_gpu_parity as tp
    node, elem, et, med = cases.case_mesh(name)
    kw = cases.case_kwargs(name)
    kw.update(nphoton=20000, schedule=1, hotcache=-1)
    a = mmc.run(tp._cfg(node, elem, et, med, **dict(kw, isnormalized=1)))
    b = mmc.run(tp._cfg(node, elem, et, med, **dict(kw, isnormalized=0)))
    assert a["raytet"] == b["raytet"] and np.array_equal(a["energyesc"], b["energyesc"])
    okw = {k: v for k, v in kw.items() if k not in ("schedule", "hotcache")}
    o = orc.run(node, elem, et, med, nthread=1, gpu_semantics=1, **dict(okw, nphoton=10))      # mesh tables of the oracle (tracer_prep)
    mua = np.concatenate([[0.0], np.asarray(med, np.float32)[:, 0] * np.float32(kw.get("unitinmm", 1.0))]).astype(np.float32)
    if (o["type"] > len(med)).any():                         # wide-field detector layer: medium prop + 1 is the background
        mua = np.concatenate([mua, [0.0]]).astype(np.float32)
    ref, nz = normalize_np.mesh_normalize(b["raw"], outputtype=kw.get("outputtype", cases.FLUX), method=kw["method"], basisorder=kw.get("basisorder", 0),
                                          energytot=b["energytot"], energyesc=b["energyesc"], tstep=kw["tstep"], elem=o["elem"], etype=o["type"],
                                          evol=o["evol"], nvol=o["nvol"], mua=mua)
    np.testing.assert_allclose(a["normalizer"], nz, rtol=1e-9)
    fa = a["raw"]
    assert fa.shape == ref.shape and np.array_equal(np.isfinite(fa), np.isfinite(ref))
    ok = np.isfinite(ref)
    np.testing.assert_allclose(fa[ok], ref[ok], rtol=1e-9, atol=0)
    assert np.abs(ref[ok]).max() > 0
    if "dref" in b:
        np.testing.assert_allclose(a["dref"], b["dref"] / np.float32(b["energytot"][0]), rtol=1e-6)
