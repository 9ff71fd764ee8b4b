"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test goes through the C-ABI library
(mmc_b200/libmmc_b200.so) via the ctypes host layer; the CPU oracle (oracle/) is only the checker.

Tolerances (north_star): RNG bit-exact; absorbed/escaped energy fractions within 0.1 % at >=1e8 photons --
here, at 2e5..1e6 photons, within 4 sigma of the Monte Carlo noise (stated per test); fluence compared where
it exceeds a fraction of the maximum."""
import numpy as np
import pytest

import cases
import orc

pytestmark = pytest.mark.gpu

mmc = pytest.importorskip("mmc_b200")


def _cfg(node, elem, et, med, **kw):
    names = {cases.PLUCKER: "plucker", cases.HAVEL: "havel", cases.BLBADOUEL: "elem", cases.GRID: "grid"}
    c = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, kw.get("nout", 1.0)], med]))
    for k, v in kw.items():
        if k == "method":
            c["method"] = names[v]
        elif k == "steps":
            c["steps"] = (v, v, v)
        elif k == "srctype":
            c["srctype"] = int(v)
        elif k in ("nthread",):
            continue
        else:
            c[k] = v
    c.setdefault("basisorder", 0)
    return c


@pytest.fixture(scope="module")
def mesh():
    return cases.two_media_cube()


def test_gpu_present():
    info = mmc.gpuinfo()
    assert len(info) >= 1, "no CUDA device visible: the CUDA path cannot be validated"
    assert info[0]["major"] >= 10, info[0]


def test_rng_bit_exact_on_device():
    """Device xorshift128+ == reference generator (src/mmc_core.cl:517-532) for the seeds the host would
    hand out (srand(seed); rand() x4 per thread, src/mmc_cu_host.cu:438,532-534)."""
    seeds = mmc.host_seeds(1648335518, 4 * 64).reshape(64, 4)
    assert np.array_equal(seeds.ravel(), orc.host_seeds(1648335518, 4 * 64))
    dev, st = mmc.rng_selftest(seeds, 257)
    for i in (0, 1, 31, 63):
        ref, refst = orc.rng_floats(seeds[i], 257)
        assert np.array_equal(dev[i].view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(st[i], refst[-1])


GPU_CASES = ["blb_elem_raw", "blb_elem_reflect", "grid_1mm", "grid_halfmm", "blb_onegate", "blb_fluence",
             "blb_energy", "blb_detectors", "blb_isotropic", "blb_mirror", "planar_widedet", "pattern_share2", "disk_grid",
             "blb_dref"]


def _finite(a):
    return np.where(np.isfinite(a), a, 0.0)        # void (mua = 0) elements normalise to inf/nan in the reference as well


@pytest.mark.parametrize("name", GPU_CASES)
def test_statistical_parity_vs_oracle(name):
    node, elem, et, med = cases.case_mesh(name)
    kw = cases.case_kwargs(name)
    N = {"blb_mirror": 20000, "pattern_share2": 500000}.get(name, 200000)    # half-dark patterns: fewer photons per lit element
    kw["nphoton"] = N
    o = orc.run(node, elem, et, med, nthread=8, gpu_semantics=1, **kw)
    g = mmc.run(_cfg(node, elem, et, med, **kw))
    srcnum = kw.get("srcnum", 1)
    for k in range(srcnum):
        # energy bookkeeping: launched, absorbed = tot - esc (src/mmc_cu_host.cu:988)
        assert abs(g["energytot"][k] - o["launchweight"][k]) <= 2e-3 * N
        fo = (o["launchweight"][k] - o["escweight"][k]) / o["launchweight"][k]
        fg = g["energyabs"][k] / g["energytot"][k]
        sigma = np.sqrt(max(fo * (1 - fo), 1e-4) / N)
        assert abs(fg - fo) < 6 * sigma + 3e-4, (k, fg, fo, sigma)
        # output volume: per-gate sums and the well-lit entries
        fo_, fg_ = _finite(o["field"][..., k]), _finite(g["raw"][..., k])
        assert fo_.shape == fg_.shape
        assert np.array_equal(np.isfinite(o["field"][..., k]), np.isfinite(g["raw"][..., k]))
        tot = fo_.sum()
        go, gg = fo_.sum(axis=1), fg_.sum(axis=1)
        big = go > 0.02 * tot
        np.testing.assert_allclose(gg[big], go[big], rtol=0.03)
        assert abs(fg_.sum() / tot - 1) < 0.02
        cw_o, cw_g = fo_.sum(axis=0), fg_.sum(axis=0)
        lit = cw_o > 0.02 * cw_o.max()
        rel = np.abs(cw_g[lit] - cw_o[lit]) / cw_o[lit]
        assert np.median(rel) < 0.05, np.median(rel)
        assert np.mean(rel) < 0.08, np.mean(rel)
    # work per photon
    assert abs(g["raytet"] / o["raytet"] - 1) < 0.02
    if kw.get("issaveref"):                         # diffuse reflectance per exterior face and gate (-X 1)
        do, dg = o["dref"], g["dref"]
        assert do.shape == dg.shape
        assert abs(dg.sum() / do.sum() - 1) < 0.02
        hot = do.sum(axis=0) > 0.05 * do.sum(axis=0).max()
        np.testing.assert_allclose(dg.sum(axis=0)[hot], do.sum(axis=0)[hot], rtol=0.15)
    if kw.get("issavedet"):
        no, ng = o["detectedcount"], len(g["detp"])
        assert abs(no - ng) < 6 * np.sqrt(max(no, 1)) + 5, (no, ng)
        assert g["detp"].shape[1] == o["reclen"]
        # columns: detid, nscat[M], ppath[M], mom[M], p[3], v[3], w0
        do, dg = o["detected"][:no], g["detp"]
        M = len(med)
        for col in (1, 1 + M):                      # scattering counts and partial paths of medium 1
            assert abs(do[:, col].mean() - dg[:, col].mean()) < 0.1 * max(abs(do[:, col].mean()), 0.05)
        vn = np.linalg.norm(dg[:, -4:-1], axis=1)
        assert np.allclose(vn, 1.0, atol=1e-4)     # exit directions are unit vectors (examples/regression/exitangle)
        if name == "blb_detectors":
            assert abs(do[:, 0].mean() - dg[:, 0].mean()) < 0.1            # detector id mix
            assert np.all(dg[:, -5] <= 1e-3)       # exit z on the z=0 face where the detectors sit
        else:                                      # wide-field detector: rows are taken where photons leave the detector layer
            assert abs(do[:, -5].mean() - dg[:, -5].mean()) < 0.1
            assert (dg[:, -5] > 10.0 - 1e-3).mean() > 0.95


def test_per_photon_seed_parity(mesh):
    """Deterministic pairing: replay-style per-photon seeds (SEED_FROM_FILE, src/mmc_core.cl:2191-2194) make every
    photon's stream independent of scheduling, so detected-photon rows can be matched one-to-one by their saved seed."""
    node, elem, et, med = mesh
    N = 20000
    rs = np.random.RandomState(7).randint(1, 2**62, size=(N, 2)).astype(np.uint64)
    kw = cases.case_kwargs("blb_detectors")
    kw.update(nphoton=N, issaveseed=1)
    o = orc.run(node, elem, et, med, nthread=8, gpu_semantics=1, seed=orc.SEED_FROM_FILE, photonseed=rs,
                replayweight=np.ones(N, np.float32), replaytime=np.zeros(N, np.float32), **{k: v for k, v in kw.items() if k != "seed"})
    cfg = _cfg(node, elem, et, med, **{k: v for k, v in kw.items() if k != "seed"})
    cfg.update(replayseed=rs, replayweight=np.ones(N, np.float32), replaytime=np.zeros(N, np.float32))
    g = mmc.run(cfg)
    key_o = {tuple(s): i for i, s in enumerate(o["detseed"])}
    both = [(key_o[tuple(s)], j) for j, s in enumerate(g["seeds"]) if tuple(s) in key_o]
    assert len(both) > 0.97 * max(len(key_o), len(g["seeds"])), (len(both), len(key_o), len(g["seeds"]))
    io, ig = np.array(both).T
    do, dg = o["detected"][io], g["detp"][ig]
    same = np.all(np.abs(do - dg) <= 2e-3 * np.maximum(1.0, np.abs(do)), axis=1)
    assert same.mean() > 0.9, same.mean()        # fp rounding (fast-math vs libm) flips a few trajectories
    assert abs(g["raw"].sum() / o["field"].sum() - 1) < 5e-3


@pytest.mark.parametrize("otype", ["wl", "wp", "jacobian", "wl-havel"])
def test_replay_outputs_vs_oracle(otype, tmp_path):
    """BASELINE config C5 flow (examples/replaywide): run 1 saves detected photons + seeds (.mch), run 2 replays them
    (-E file.mch -O L|P|J) and accumulates path lengths / scattering counts weighted by the detected weight.  Replayed
    photons carry their own seeds, so GPU and oracle follow the same trajectories up to fp rounding."""
    from mmc_b200 import mch
    node, elem, et, med = cases.case_mesh("planar_widedet")
    kw = cases.case_kwargs("planar_widedet")
    kw.update(nphoton=60000, issaveseed=1)
    first = mmc.run(_cfg(node, elem, et, med, **kw))
    assert len(first["detp"]) > 3000
    f = str(tmp_path / "init.mch")
    mch.savemch(f, first["detp"], first["seeds"], maxmedia=len(med), totalphoton=kw["nphoton"], normalizer=first["normalizer"])
    rp = mch.replay_inputs(mch.loadmch(f), np.vstack([[0, 0, 1, 1], med]))
    n = rp["nphoton"]
    kw2 = {k: v for k, v in kw.items() if k not in ("seed", "nphoton", "issaveseed")}
    kw2.update(outputtype={"wl": cases.WL, "wp": cases.WP, "jacobian": cases.JACOBIAN, "wl-havel": cases.WL}[otype], minenergy=0.0)
    if otype == "wl-havel":
        kw2["method"] = cases.HAVEL
    o = orc.run(node, elem, et, med, nthread=8, gpu_semantics=0 if otype == "wl-havel" else 1, seed=orc.SEED_FROM_FILE, nphoton=n, photonseed=rp["replayseed"],
                replayweight=rp["replayweight"], replaytime=rp["replaytime"], **kw2)
    cfg = _cfg(node, elem, et, med, **kw2)
    cfg.update(replayseed=rp["replayseed"], replayweight=rp["replayweight"], replaytime=rp["replaytime"])
    g = mmc.run(cfg)
    # replay invariant of the reference (matlab/mmcjmua.m:55-60): the replayed photons are detected again, same rows
    assert abs(len(g["detp"]) - n) <= 0.02 * n
    fo, fg = np.where(np.isfinite(o["field"][..., 0]), o["field"][..., 0], 0), np.where(np.isfinite(g["raw"][..., 0]), g["raw"][..., 0], 0)
    assert abs(fg.sum() / fo.sum() - 1) < 5e-3
    cw_o, cw_g = fo.sum(axis=0), fg.sum(axis=0)
    lit = cw_o > 0.05 * cw_o.max()
    rel = np.abs(cw_g[lit] - cw_o[lit]) / cw_o[lit]
    assert np.median(rel) < 0.02 and np.percentile(rel, 95) < 0.1, (np.median(rel), np.percentile(rel, 95))


def test_energy_conservation_and_deposit_completeness(mesh):
    """sum(raw energy deposits) == launched - escaped: no deposit is lost between the merged-run flushes."""
    node, elem, et, med = mesh
    kw = cases.case_kwargs("blb_energy")
    kw.update(nphoton=300000, isnormalized=0)
    g = mmc.run(_cfg(node, elem, et, med, **kw))
    assert abs(g["raw"].sum() / g["energyabs"][0] - 1) < 2e-4


HP_CASES = ["havel_elem", "havel_nodal", "plucker_elem", "plucker_nodal", "planar_havel_nodal", "havel_elem_det", "plucker_nodal_det"]


@pytest.mark.parametrize("name", HP_CASES)
def test_havel_plucker_parity_vs_oracle(name):
    """Havel and Plucker exist only in the reference's CPU file (src/mmc_raytrace.c:227-508,531-800): the oracle runs with
    CPU semantics (gpu_semantics=0) and the CUDA kernels must reproduce its absorbed fraction, work per photon and
    per-gate / per-node (or per-element) fluence within Monte Carlo noise."""
    node, elem, et, med = cases.case_mesh(name)
    kw = cases.case_kwargs(name)
    N = 200000
    kw["nphoton"] = N
    o = orc.run(node, elem, et, med, nthread=8, gpu_semantics=0, **kw)
    g = mmc.run(_cfg(node, elem, et, med, **kw))
    fo = (o["absorbweight"] / o["launchweight"])[0]
    fg = g["energyabs"][0] / g["energytot"][0]
    sigma = np.sqrt(max(fo * (1 - fo), 1e-4) / N)
    assert abs(fg - fo) < 6 * sigma + 3e-4, (fg, fo, sigma)
    assert abs(g["raytet"] / o["raytet"] - 1) < 0.02
    fo_, fg_ = o["field"][..., 0], g["raw"][..., 0]
    assert fo_.shape == fg_.shape
    go, gg = fo_.sum(axis=1), fg_.sum(axis=1)
    big = go > 0.02 * go.sum()
    np.testing.assert_allclose(gg[big], go[big], rtol=0.03)
    cw_o, cw_g = fo_.sum(axis=0), fg_.sum(axis=0)
    lit = cw_o > 0.02 * cw_o.max()
    rel = np.abs(cw_g[lit] - cw_o[lit]) / cw_o[lit]
    assert np.median(rel) < 0.05, np.median(rel)
    assert np.mean(rel) < 0.08, np.mean(rel)
    if kw.get("issavedet"):                         # detected-photon records of the detector kernel variants (CPU-file columns)
        no, ng = o["detectedcount"], len(g["detp"])
        assert abs(no - ng) < 6 * np.sqrt(max(no, 1)) + 5, (no, ng)
        assert g["detp"].shape[1] == o["reclen"]
        do, dg = o["detected"][:no], g["detp"]
        M = len(med)
        for col in (1, 1 + M):                      # scattering counts and partial paths of medium 1
            assert abs(do[:, col].mean() - dg[:, col].mean()) < 0.1 * max(abs(do[:, col].mean()), 0.05)
        assert abs(do[:, 0].mean() - dg[:, 0].mean()) < 0.1            # detector id mix
        assert np.allclose(np.linalg.norm(dg[:, -4:-1], axis=1), 1.0, atol=1e-4)


def test_three_tracers_agree_on_cube60():
    """BASELINE config C1 with the tracer it names (Havel) and the two others: same absorbed fraction (reference CPU
    anchors: Havel 17.70356 %, Plucker 17.69245 %, BL-Badouel 17.70358 % at 1e6 photons)."""
    node, elem, et = mmc.meshgen.cube60()
    med = [(0.005, 1.0, 0.01, 1.37)]
    base = dict(nphoton=1000000, seed=1648335518, srcpos=(30.1, 30.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9,
                tstep=1e-10, isreflect=0)
    for method, anchor, steps in ((cases.HAVEL, 0.1770356, 206.9), (cases.PLUCKER, 0.1769245, 206.8), (cases.BLBADOUEL, 0.1770358, 206.9)):
        g = mmc.run(_cfg(node, elem, et, med, method=method, **base))
        assert abs(g["energyabs"][0] / g["energytot"][0] - anchor) < 2.5e-3, method
        assert abs(g["raytet"] / 1e6 - steps) < 3.0, method


@pytest.mark.parametrize("name", ["blb_energy", "grid_halfmm"])
def test_hot_line_cache_loses_nothing(name, mesh):
    """The CTA-private sums for the hottest lines (pilot batch -> key table -> shared-memory accumulation -> flush) must
    deposit exactly what the direct red.global path deposits: energy conservation to fp32 round-off, and the same
    volume as a run with the cache switched off (same seeds => same trajectories up to scheduling)."""
    node, elem, et, med = mesh
    kw = cases.case_kwargs(name)
    kw.update(nphoton=400000, isnormalized=0, outputtype=cases.ENERGY)
    on = mmc.run(_cfg(node, elem, et, med, hotcache=1, **kw))
    off = mmc.run(_cfg(node, elem, et, med, hotcache=-1, **kw))
    assert abs(on["raw"].sum() / on["energyabs"][0] - 1) < 2e-4
    assert abs(off["raw"].sum() / off["energyabs"][0] - 1) < 2e-4
    assert abs(on["energyabs"][0] / off["energyabs"][0] - 1) < 5e-3
    a, b = on["raw"].reshape(-1), off["raw"].reshape(-1)
    top = np.argsort(b)[-200:]                    # the hottest accumulators are the privatised ones
    np.testing.assert_allclose(a[top], b[top], rtol=0.05)
    assert abs(a[top].sum() / b[top].sum() - 1) < 0.01


def test_dynamic_and_static_schedules_agree(mesh):
    node, elem, et, med = mesh
    kw = cases.case_kwargs("blb_elem_reflect")
    kw["nphoton"] = 300000
    a = mmc.run(_cfg(node, elem, et, med, schedule=0, **kw))
    b = mmc.run(_cfg(node, elem, et, med, schedule=1, **kw))
    fa, fb = a["energyabs"][0] / a["energytot"][0], b["energyabs"][0] / b["energytot"][0]
    assert a["energytot"][0] == b["energytot"][0] == kw["nphoton"]
    assert abs(fa - fb) < 5e-3


def test_full_size_cube60_anchor():
    """BASELINE config C1 at full size (29 791 nodes / 135 000 tets, 50 gates, 1e6 photons): absorbed fraction vs the
    reference CPU anchors recorded in BASELINE.md (17.70 % at -b 0, 27.26 % at -b 1; sigma ~0.04 %)."""
    node, elem, et = mmc.meshgen.cube60()
    med = [(0.005, 1.0, 0.01, 1.37)]
    base = dict(nphoton=1000000, seed=1648335518, srcpos=(30.1, 30.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9,
                tstep=1e-10, method=cases.BLBADOUEL)
    g0 = mmc.run(_cfg(node, elem, et, med, isreflect=0, **base))
    assert g0["e0"] == 4497                                    # examples/validation/cube.inp:7
    assert abs(g0["energyabs"][0] / g0["energytot"][0] - 0.17704) < 2.5e-3
    assert abs(g0["raytet"] / 1e6 - 206.9) < 3.0
    g1 = mmc.run(_cfg(node, elem, et, med, isreflect=1, **base))
    assert abs(g1["energyabs"][0] / g1["energytot"][0] - 0.27259) < 2.5e-3
    assert abs(g1["raytet"] / 1e6 - 338.9) < 4.0


def test_onecall_entry_point_matches_session_path(mesh):
    """mmcb_run_simulation (the entry the reference-side stub calls) and create/launch/fetch/destroy are the same engine."""
    node, elem, et, med = mesh
    kw = cases.case_kwargs("blb_detectors")
    kw["nphoton"] = 100000
    a = mmc.run(_cfg(node, elem, et, med, **kw))
    b = mmc.run_onecall(_cfg(node, elem, et, med, **kw))
    assert a["energytot"][0] == b["energytot"][0] == kw["nphoton"]
    assert abs(a["energyabs"][0] / b["energyabs"][0] - 1) < 0.01
    assert abs(len(a["detp"]) - len(b["detp"])) < 6 * np.sqrt(len(a["detp"])) + 5
    assert a["raw"].shape == b["raw"].shape


def test_errors_match_reference_convention(mesh):
    node, elem, et, med = mesh
    kw = cases.case_kwargs("blb_elem_reflect")
    bad = _cfg(node, elem, et, med, **kw)
    bad["srcpos"] = (100.0, 100.0, 100.0)
    with pytest.raises(mmc.MMCError, match="initial element does not enclose the source"):
        mmc.run(bad)
    bad = _cfg(node, elem, et, med, **kw)
    bad["srcdir"] = (0, 0, 2.0)
    with pytest.raises(mmc.MMCError, match="unitary"):
        mmc.run(bad)
