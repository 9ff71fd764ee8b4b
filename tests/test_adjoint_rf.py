"""Multi-slot sources (adjoint mode), RF (frequency-domain) forward runs and the adjoint-Jacobian post-kernels:
src/mmc_core.cl:1431-1515 (slot launch), :872-896 / :1043-1078 (complex deposit), :2218-2649 (Jacobian kernels), driven like
src/mmc_cu_host.cu:229-233,997-1395.  The reference implements all of this on the GPU only, so the checks are

  * multi-slot forward run  == the same slots run one by one as ordinary single sources through the already-validated path
  * RF fluence at omega     == Fourier transform at omega of the time-resolved (50 gates) CW run of the same problem
  * Jacobian post-kernels   == oracle/adjoint_np.py (numpy restatement of the four kernels) on the SAME fluence, rtol 2e-4
                               (float32 products and sums in a different order)
All GPU tests call through the C-ABI (mmc_b200.run)."""
import numpy as np
import pytest

import adjoint_np
import cases

mmc = pytest.importorskip("mmc_b200")

DETS = [(10.3, 8.4, 0.0, 0.2), (11.7, 12.4, 20.0, 0.2)]
DETDIR = [(0, 0, 1, 0), (0, 0, -1, 0)]


def _cfg(**kw):
    node, elem, et, med = cases.two_media_cube()
    c = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), nphoton=300000, seed=1648335518,
             srcpos=(10.1, 10.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=1e-9, isreflect=1, method="elem", basisorder=0)
    c.update(kw)
    return c


# ----------------------------------------------------------------------------------------------------------------------
# host logic (no GPU): sizes and validation of the new configuration fields
# ----------------------------------------------------------------------------------------------------------------------
def test_adjoint_sizes_and_validation_cpu():
    from mmc_b200 import api
    p = api.Problem(_cfg(method="grid", outputtype="adjointmuad", detpos=DETS, detdir=DETDIR, omega=1e9))
    sz = p.sizes()
    assert (sz.nslots, sz.adj_ns, sz.adj_nd) == (3, 1, 2)
    assert sz.fieldlen == sz.datalen * sz.maxgate * 3
    assert sz.jacoblen == sz.datalen * 2 * 2 * 2                 # Ns*Nd pairs x (Re, Im) x (J_mua, J_D)
    p = api.Problem(_cfg(srcid=-2, detpos=DETS, detdir=DETDIR))   # detectors appended as sources, no Jacobian
    sz = p.sizes()
    assert (sz.nslots, sz.jacoblen) == (3, 0)
    sd = np.zeros((2, 16), np.float32)
    sd[:, 0:3] = [(10.1, 10.2, 0.0), (5.5, 5.5, 0.0)]
    sd[:, 3] = 0.5
    sd[:, 6] = 1
    assert api.Problem(_cfg(srcdata=sd)).sizes().nslots == 2       # srcdata without a selector: all slots
    assert api.Problem(_cfg(srcdata=sd, srcid=2)).sizes().nslots == 1
    with pytest.raises(api.MMCError, match="srcid exceeds"):
        api.Problem(_cfg(srcdata=sd, srcid=3)).sizes()
    with pytest.raises(api.MMCError, match="basisorder=1"):
        api.Problem(_cfg(outputtype="adjoint", detpos=DETS, detdir=DETDIR)).sizes()
    with pytest.raises(api.MMCError, match="branch-less Badouel"):
        api.Problem(_cfg(omega=1e9, method="havel")).sizes()


def test_deldotdel_oracle_is_consistent():
    """sum_j grad N_i . grad N_j = 0 (the shape functions sum to one) and the diagonal is positive."""
    node, elem, et, med = cases.two_media_cube()
    e = np.asarray(elem) - 1
    p = np.asarray(node, np.float64)
    vol = np.abs(np.einsum("ij,ij->i", np.cross(p[e[:, 1]] - p[e[:, 0]], p[e[:, 2]] - p[e[:, 0]]), p[e[:, 3]] - p[e[:, 0]])) / 6
    d = adjoint_np.deldotdel(node, elem, vol)
    full = np.zeros((len(e), 4, 4))
    k = 0
    for i in range(4):
        for j in range(i, 4):
            full[:, i, j] = full[:, j, i] = d[:, k]
            k += 1
    assert np.abs(full.sum(axis=2)).max() < 1e-9 * np.abs(full).max()
    assert (np.diagonal(full, axis1=1, axis2=2) > 0).all()


# ----------------------------------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_multislot_blocks_match_single_source_runs():
    N = 900000
    g = mmc.run(_cfg(nphoton=N, srcid=-2, detpos=DETS, detdir=DETDIR, isnormalized=0, issavedet=1))
    assert g["nslots"] == 3 and g["raw"].shape[0] == 3
    slots = [((10.1, 10.2, 0.0), (0, 0, 1), 0.0, 1.0)] + [(d[:3], dd[:3], d[3], 0.5) for d, dd in zip(DETS, DETDIR)]
    # launched weight: every photon picks one of 3 slots uniformly, weights 1, 1/2, 1/2
    assert abs(g["energytot"][0] / (N * 2.0 / 3.0) - 1) < 5e-3
    for k, (pos, dr, radius, w) in enumerate(slots):
        c1 = _cfg(nphoton=N // 3, srcpos=pos, srcdir=dr, isnormalized=0)
        if radius > 0:      # the slot launches every photon in the element that encloses the slot centre (srcparam2.w): same here
            c1.update(srctype="disk", srcparam1=(radius, 0, 0, 0), e0=mmc.mesh_initelem(c1["node"], c1["elem"], pos)[0])
        one = mmc.run(c1)
        blk = g["raw"][k].sum(axis=0) / w                # per-element CW sums of ~N/3 unit-weight photons
        ref = one["raw"][..., 0].sum(axis=0)
        assert abs(blk.sum() / ref.sum() - 1) < 0.02, (k, blk.sum(), ref.sum())
        lit = ref > 0.02 * ref.max()
        rel = np.abs(blk[lit] - ref[lit]) / ref[lit]
        assert lit.sum() > 30 and np.median(rel) < 0.06, (k, lit.sum(), np.median(rel))
    # detected-photon rows carry the launch slot (1-based) in the upper 16 bits of the detector id (src/mmc_core.cl:652-658)
    ids = g["detp"][:, 0].astype(np.uint32)
    assert len(ids) > 100
    assert set(np.unique(ids >> 16)) <= {1, 2, 3} and set(np.unique(ids & 0xFFFF)) <= {1, 2}
    # a photon launched from a detector-source mostly comes back to that detector
    assert ((ids >> 16) == 1).sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["elem", "grid"])
def test_rf_fluence_is_the_fourier_transform_of_the_time_resolved_run(method):
    omega = 2 * np.pi * 2e8                               # 200 MHz
    kw = dict(nphoton=1000000, tstep=1e-10, tend=5e-9, isnormalized=0, method=method, steps=(1.0, 1.0, 1.0))
    cw = mmc.run(_cfg(**kw))
    rf = mmc.run(_cfg(omega=omega, **kw))
    raw = cw["raw"][..., 0]                               # [gate, i]
    tg = (np.arange(raw.shape[0]) + 0.5) * 1e-10
    ft = (raw * np.exp(-1j * omega * tg)[:, None]).sum(axis=0)
    z = rf["raw"][..., 0].sum(axis=0) + 1j * rf["raw_im"][..., 0].sum(axis=0)
    lit = np.abs(ft) > 0.02 * np.abs(ft).max()
    rel = np.abs(z[lit] - ft[lit]) / np.abs(ft[lit])
    assert lit.sum() > 100
    assert np.median(rel) < 0.04, np.median(rel)
    assert abs(np.abs(z[lit]).sum() / np.abs(ft[lit]).sum() - 1) < 0.01
    # the phase lag grows away from the source: it is not a real field in disguise
    assert np.abs(z.imag[lit]).sum() > 0.05 * np.abs(z.real[lit]).sum()
    # same photons (static schedule: thread i runs the same photon ids with the same stream), omega -> 0: the real part is the
    # CW field and the imaginary part vanishes
    kw1 = dict(kw, nphoton=200000, schedule=1, hotcache=-1)
    c0, r0 = mmc.run(_cfg(**kw1)), mmc.run(_cfg(omega=1.0, **kw1))
    a, b = r0["raw"][..., 0].sum(axis=0), c0["raw"][..., 0].sum(axis=0)
    big = b > 0.02 * b.max()
    np.testing.assert_allclose(a[big], b[big], rtol=2e-4)
    assert np.abs(r0["raw_im"]).sum() < 1e-6 * np.abs(r0["raw"]).sum()
    # launched energy is unchanged, the escaped energy is |w|
    assert abs(rf["energytot"][0] - cw["energytot"][0]) < 1e-6 * cw["energytot"][0]
    assert abs(rf["energyesc"][0] / cw["energyesc"][0] - 1) < 5e-3
    assert abs(r0["energyesc"][0] / c0["energyesc"][0] - 1) < 1e-5


def _jac_inputs(g):
    f = g["raw"].astype(np.float32)                       # what the library uploads: float32 of the normalised volumes
    fi = g["raw_im"].astype(np.float32) if "raw_im" in g else None
    return adjoint_np.cw_sum(f), (adjoint_np.cw_sum(fi) if fi is not None else None)


def _close(a, b, what):
    tol = 2e-4 * np.abs(b).max()
    assert np.abs(a - b).max() <= tol + 2e-4 * np.abs(b).max(), (what, np.abs(a - b).max(), np.abs(b).max())
    np.testing.assert_allclose(a, b, rtol=2e-3, atol=tol, err_msg=what)


@pytest.mark.gpu
@pytest.mark.parametrize("omega", [0.0, 2 * np.pi * 1e8])
def test_grid_adjoint_jacobians_match_the_numpy_oracle(omega):
    g = mmc.run(_cfg(method="grid", steps=(1.0, 1.0, 1.0), outputtype="adjointmuad", detpos=DETS, detdir=DETDIR, omega=omega,
                     nphoton=600000, unitinmm=1.0))
    Ns, Nd = g["adj_ns"], g["adj_nd"]
    assert (Ns, Nd) == (1, 2)
    cwr, cwi = _jac_inputs(g)
    mre, mim = adjoint_np.jmua_grid(cwr, cwi, Ns, Nd, -1.0)
    dre, dim_ = adjoint_np.jd_grid(cwr, cwi, Ns, Nd, g["dim"], -1.0)
    J = g["jacob"]                                        # [component, pair, voxel]: CW [Jmua, JD]; RF [Re Jmua, Re JD, Im Jmua, Im JD]
    assert J.shape[0] == (4 if omega else 2)
    assert np.abs(J[0]).max() > 0 and np.abs(J[1]).max() > 0
    _close(J[0], mre, "J_mua re")
    _close(J[1], dre, "J_D re")
    if omega:
        _close(J[2], mim, "J_mua im")
        _close(J[3], dim_, "J_D im")
    # single-component types reuse the same kernels
    g1 = mmc.run(_cfg(method="grid", steps=(1.0, 1.0, 1.0), outputtype="adjoint", detpos=DETS, detdir=DETDIR, omega=omega, nphoton=100000))
    cwr, cwi = _jac_inputs(g1)
    mre, mim = adjoint_np.jmua_grid(cwr, cwi, Ns, Nd, -1.0)
    _close(g1["jacob"][0], mre, "J_mua (A) re")
    if omega:
        _close(g1["jacob"][1], mim, "J_mua (A) im")


@pytest.mark.gpu
@pytest.mark.parametrize("omega", [0.0, 2 * np.pi * 1e8])
def test_mesh_adjoint_jacobians_match_the_numpy_oracle(omega):
    from mmc_b200 import api
    c = _cfg(method="elem", basisorder=1, outputtype="adjointmuad", detpos=DETS, detdir=DETDIR, omega=omega, nphoton=600000)
    g = mmc.run(c)
    Ns, Nd = g["adj_ns"], g["adj_nd"]
    elem, evol, nvol = api.mesh_volumes(c["node"], c["elem"], c["elemprop"])
    cwr, cwi = _jac_inputs(g)
    ddd = adjoint_np.deldotdel(c["node"], elem, evol)
    ref = adjoint_np.jac_mesh_full(cwr, cwi, elem, evol, ddd, Ns, Nd)
    J = g["jacob"]
    assert J.shape == ((4 if omega else 2), Ns * Nd, len(c["node"]))
    _close(J[0], ref["jmua"][0], "mesh J_mua re")
    _close(J[1], ref["jd"][0], "mesh J_D re")
    if omega:
        _close(J[2], ref["jmua"][1], "mesh J_mua im")
        _close(J[3], ref["jd"][1], "mesh J_D im")
    # nodal approximation of J_mua (adjointmode = 1): -nvol phi_s phi_d.  Interior nodes carry the plain nodal volume of
    # mesh_getvolume; nodes on the exterior surface carry tracer_prep's correction nvol *= 2/(1+Reff) (src/mmc_mesh.c:1344-1386), so
    # there the check is that both detector pairs see the same volume
    g1 = mmc.run(dict(c, outputtype="adjoint", adjointmode=1, nphoton=100000))
    cwr, cwi = _jac_inputs(g1)
    re, im = adjoint_np.jmua_mesh_nodal(cwr, cwi, nvol, Ns, Nd)
    xyz = np.asarray(c["node"])
    inner = np.all((xyz > 1e-3) & (xyz < 20 - 1e-3), axis=1)
    _close(g1["jacob"][0][:, inner], re[:, inner], "nodal J_mua re")
    if omega:
        _close(g1["jacob"][1][:, inner], im[:, inner], "nodal J_mua im")
    else:
        ratio = g1["jacob"][0] / np.where(re != 0, re, 1)
        ok = (np.abs(re) > 1e-6 * np.abs(re).max()).all(axis=0) & ~inner
        assert ok.sum() > 20
        np.testing.assert_allclose(ratio[0][ok], ratio[1][ok], rtol=1e-4)
        assert (ratio[0][ok] > 0.5).all() and (ratio[0][ok] < 2.0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["elem", "grid"])
def test_per_node_optical_properties(method):
    """cfg.nodemua / cfg.nodemusp (src/mmc_core.cl:776-793): an element takes the mean of its four nodal values.  Nodal arrays that
    repeat the media table must reproduce the ordinary run exactly (same photons: static schedule); raising the nodal absorption
    raises the absorbed fraction; mus from nodemusp changes the step count."""
    c = _cfg(method=method, steps=(1.0, 1.0, 1.0), nphoton=100000, schedule=1, hotcache=-1, isnormalized=0, tstep=5e-10)
    node, elem, et = np.asarray(c["node"]), np.asarray(c["elem"]), np.asarray(c["elemprop"])
    prop = np.asarray(c["prop"], np.float32)
    base = mmc.run(c)
    # uniform medium 1 everywhere (labels kept): nodal values = medium 1
    c1 = dict(c, prop=np.vstack([prop[0], prop[1], prop[1]]))
    ref1 = mmc.run(c1)
    nod = mmc.run(dict(c1, nodemua=np.full(len(node), prop[1, 0], np.float32), nodemusp=np.full(len(node), prop[1, 1], np.float32)))
    assert nod["raytet"] == ref1["raytet"]
    np.testing.assert_allclose(nod["energyesc"], ref1["energyesc"], rtol=1e-6)
    np.testing.assert_allclose(nod["raw"].sum(), ref1["raw"].sum(), rtol=1e-6)
    # mua only: trajectories are unchanged (same scattering), absorption follows the nodal field
    hi = mmc.run(dict(c1, nodemua=np.full(len(node), 4 * prop[1, 0], np.float32)))
    assert hi["raytet"] <= ref1["raytet"] * 1.0001            # photons are weighted, not killed, but time gates are the same
    fa = lambda r: r["energyabs"][0] / r["energytot"][0]
    assert fa(hi) > 1.5 * fa(ref1)
    # a nodal field that varies in space: absorbed fraction lies between the two uniform cases
    z = node[:, 2]
    mid = mmc.run(dict(c1, nodemua=np.where(z > 10, 4 * prop[1, 0], prop[1, 0]).astype(np.float32)))
    assert fa(ref1) < fa(mid) < fa(hi)
    # musp: doubling the scattering coefficient raises the number of ray-tet steps per photon
    more = mmc.run(dict(c1, nodemua=np.full(len(node), prop[1, 0], np.float32), nodemusp=np.full(len(node), 2 * prop[1, 1], np.float32)))
    assert more["raytet"] > 1.2 * ref1["raytet"]
    assert base["raytet"] > 0
    with pytest.raises(mmc.MMCError, match="branch-less Badouel"):
        mmc.run(dict(c1, method="havel", nodemua=np.full(len(node), 0.01, np.float32)))


def test_fd_gradient_oracle_matches_numpy_second_order():
    """oracle/adjoint_np.fd_grad restates mmc_fd_grad (src/mmc_core.cl:2247-2285): central differences inside, second-order one-sided
    differences at the ends -- the scheme of numpy.gradient(edge_order=2) for unit spacing; exact on quadratics."""
    rs = np.random.RandomState(5)
    vol = rs.rand(7, 6, 5).astype(np.float32)
    for axis in range(3):
        np.testing.assert_allclose(adjoint_np.fd_grad(vol, axis), np.gradient(vol.astype(np.float64), axis=axis, edge_order=2), rtol=2e-5, atol=2e-6)
    z, y, x = np.meshgrid(np.arange(7.0), np.arange(6.0), np.arange(5.0), indexing="ij")
    q = (0.5 * x * x - 2 * y * y + 0.25 * z * z + x * y).astype(np.float32)
    np.testing.assert_allclose(adjoint_np.fd_grad(q, 2), x + y, atol=1e-4)          # d/dx, x fastest
    np.testing.assert_allclose(adjoint_np.fd_grad(q, 1), -4 * y + x, atol=1e-4)
    np.testing.assert_allclose(adjoint_np.fd_grad(q, 0), 0.5 * z, atol=1e-4)
    two = rs.rand(2, 3, 4).astype(np.float32)                                        # N == 2: plain difference on both ends
    np.testing.assert_allclose(adjoint_np.fd_grad(two, 0), np.stack([two[1] - two[0]] * 2))
    # Jacobian products on a tiny volume: J_mua = -phi_s phi_d, J_D = -grad phi_s . grad phi_d (scale -1)
    cw = rs.rand(3, 7 * 6 * 5).astype(np.float32)
    jm, _ = adjoint_np.jmua_grid(cw, None, 1, 2, -1.0)
    np.testing.assert_allclose(jm[1], -cw[0] * cw[2], rtol=1e-6)
    jd, _ = adjoint_np.jd_grid(cw, None, 1, 2, (5, 6, 7), -1.0)
    g = lambda k: np.stack(np.gradient(cw[k].astype(np.float64).reshape(7, 6, 5), edge_order=2))
    np.testing.assert_allclose(jd[0], -(g(0) * g(1)).sum(axis=0).ravel(), rtol=2e-4, atol=2e-5)
