"""BASELINE configs C3 and C5 at the sizes BASELINE.json states, against the reference's own CUDA kernel on the same GPU
(oracle/_ref/mmc_refcuda, the unmodified src/mmc_core.cl + src/mmc_cu_host.cu compiled for sm_100).

C3  examples/skinvessel/dmmc_skinvessel.json: shipped dual-grid mesh, disk source, `--gridsize 0.005` (200 x 229 x 202 voxels ... the mesh's
    bounding box in 5 um voxels), 10 time gates, 1e8 photons.
C5  examples/replaywide (createmesh.m:3-29, createpattern.m:1-24, run_test.sh): 60 x 60 x 20 mm slab with air layers, 40 x 40 pattern
    source over the tets labelled -1, wide-field detector layer labelled -2, run 1 with 1e8 photons saving seeds and exit positions,
    run 2 replays the detected photons (-E init.mch -P 0 -O L)."""
import os
import sys

import numpy as np
import pytest

import cases
import orc
from test_gpu_parity import _cfg

pytestmark = pytest.mark.gpu
mmc = pytest.importorskip("mmc_b200")
needs_refcuda = pytest.mark.skipif(not orc.ref_available(cuda=True), reason="oracle/_ref/mmc_refcuda not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _blocks(v, dims, b):
    """sum a flat x-fastest volume over b x b x b blocks"""
    nx, ny, nz = dims
    a = v.reshape(nz, ny, nx)
    a = a[:nz // b * b, :ny // b * b, :nx // b * b]
    return a.reshape(nz // b, b, ny // b, b, nx // b, b).sum(axis=(1, 3, 5))


@needs_refcuda
def test_c3_skinvessel_dual_grid_1e8_vs_reference_cuda():
    z = np.load(os.path.join(GOLD, "skinvessel_mesh.npz"))
    N = 100000000
    kw = dict(nphoton=N, seed=1648335518, srcpos=(0.5, 0.5, -0.005), srcdir=(0, 0, 1), srctype=8, srcparam1=(0.3, 0, 0, 0),
              tstart=0.0, tend=5e-8, tstep=5e-9, isreflect=0, method=cases.GRID, basisorder=0, steps=0.005)
    g = mmc.run(dict(node=z["node"], elem=z["elem"], elemprop=z["etype"], prop=np.vstack([[0, 0, 1, 1], z["prop"]]), evol=z["evol"], e0=6178,
                     method="grid", srctype="disk", steps=(0.005, 0.005, 0.005),
                     **{k: v for k, v in kw.items() if k not in ("method", "steps", "srctype")}))
    dims = tuple(int(v) for v in g["flux"].shape[:3])
    ours = g["raw"][..., 0].sum(axis=0)
    gate_o = g["raw"][..., 0].sum(axis=1)
    fg = g["energyabs"][0] / g["energytot"][0]
    kms = g["kernel_ms"]
    del g
    r = orc.run_ref(z["node"], z["elem"], z["etype"], z["prop"], cuda=True, timeout=1800, e0=6178, evol=z["evol"], **kw)
    fr = r["absorbed_frac"]
    ref4 = r["field_flat"].reshape(10, -1)
    assert ref4.shape[1] == ours.shape[0] == dims[0] * dims[1] * dims[2], (ref4.shape, ours.shape, dims)
    ref, gate_r = ref4.sum(axis=0), ref4.sum(axis=1)
    del ref4, r["field_flat"]
    print("skinvessel %s voxels: absorbed %.6f vs %.6f; kernel %.0f ms vs reference %.0f ms" % (dims, fg, fr, kms, r.get("kernel_ms", float("nan"))))
    assert abs(fg - fr) < 1e-3 * fr, (fg, fr)                                   # energy fractions within 0.1 %
    big = gate_r > 1e-3 * gate_r.max()
    np.testing.assert_allclose(gate_o[big], gate_r[big], rtol=5e-3)             # time-resolved totals
    # depth profile: every 5 um layer that holds more than 1e-3 of the brightest layer within 1 %
    lo, lr = ours.reshape(dims[2], -1).sum(axis=1), ref.reshape(dims[2], -1).sum(axis=1)
    lay = lr > 1e-3 * lr.max()
    np.testing.assert_allclose(lo[lay], lr[lay], rtol=0.01)
    # 5 um voxels hold a few thousand photon visits each at 1e8 photons, so the voxel-level agreement is bounded by Poisson noise
    # (median), while 4^3-voxel blocks (20 um) resolve north_star's 2 % on every block above 1e-3 of the maximum
    lit = ref > 1e-3 * ref.max()
    rel = np.abs(ours[lit] - ref[lit]) / ref[lit]
    print("  voxels above 1e-3 of the maximum: %d, median %.4f, p99 %.4f" % (lit.sum(), np.median(rel), np.percentile(rel, 99)))
    assert lit.sum() > 10000 and np.median(rel) < 0.03
    stats = {}
    for b in (4, 8, 16):
        bo, br = _blocks(ours, dims, b), _blocks(ref, dims, b)
        blit = br > 1e-3 * br.max()
        brel = np.abs(bo[blit] - br[blit]) / br[blit]
        stats[b] = (int(blit.sum()), float(np.median(brel)), float(np.percentile(brel, 99)), float(brel.max()))
        print("  %2d^3-voxel blocks above 1e-3 of the maximum: %d, median %.4f, p99 %.4f, worst %.4f" % ((b,) + stats[b]))
    assert stats[4][0] > 2000 and stats[4][1] < 0.01 and stats[4][2] < 0.05, stats[4]
    assert stats[16][1] < 0.003 and stats[16][2] < 0.02, stats[16]


def _pattern_half_dark():
    pat = np.ones((40, 40), np.float32)          # createpattern.m:5-6: pat1(1:20, :) = 0, written transposed
    pat[:20, :] = 0
    return np.ascontiguousarray(pat.T).reshape(-1)


@needs_refcuda
def test_c5_replaywide_full_size_vs_reference_cuda(tmp_path):
    from mmc_b200 import mch, meshgen
    node, elem, et = meshgen.slab_with_wide_src_det()             # 60 x 60 x 20 mm + 2 mm air layers: -1 below, -2 above
    med = [(0.01, 1.0, 0.01, 1.37)]                                  # prop_replaywide.dat
    N = 100000000
    kw = dict(nphoton=N, seed=12345678, srcpos=(10.0, 10.0, -1.0), srcdir=(0, 0, 1), srctype=5, srcparam1=(40.0, 0, 0, 40), srcparam2=(0, 40.0, 0, 40),
              srcpattern=_pattern_half_dark(), srcnum=1, tstart=0.0, tend=2e-9, tstep=2e-9, isreflect=1, method=cases.BLBADOUEL, basisorder=0,
              issavedet=1, issaveexit=1, issaveseed=1, maxdetphoton=8000000)
    first = mmc.run(_cfg(node, elem, et, med, **kw))
    nd = len(first["detp"])
    f1 = first["energyabs"][0] / first["energytot"][0]
    print("run 1: %d of %d photons detected on the wide-field layer, launched weight %.0f, absorbed %.5f, kernel %.0f ms"
          % (nd, N, first["energytot"][0], f1, first["kernel_ms"]))
    assert nd > 100000 and first["detectedtotal"] == nd
    assert abs(first["energytot"][0] / N - 0.5) < 2e-3                          # half of the pattern is dark
    # run 1 in the reference's kernel: the same statistics (its program fails at 1e8 photons with detectors on, like on the head mesh;
    # 1e7 photons give the detected fraction to 0.3 %)
    try:
        r1 = orc.run_ref(node, elem, et, med, cuda=True, timeout=900, **dict(kw, issaveseed=0))
        nref = N
    except RuntimeError as e:
        assert "illegal memory access" in str(e), str(e)[-500:]
        r1 = orc.run_ref(node, elem, et, med, cuda=True, timeout=900, **dict(kw, issaveseed=0, nphoton=10000000, maxdetphoton=1000000))
        nref = 10000000
    print("reference run 1 (%g photons): absorbed %.5f, detected %d, kernel %.0f ms" % (nref, r1["absorbed_frac"], r1["detectedcount"], r1.get("kernel_ms", float("nan"))))
    assert abs(f1 - r1["absorbed_frac"]) < (1e-3 if nref == N else 2e-3) * f1
    sc = nref / N
    assert abs(r1["detectedcount"] - nd * sc) < 5 * np.sqrt(r1["detectedcount"] + nd * sc * sc)
    # run 2: both programs replay THIS engine's detected photons from the same .mch (-E init.mch -P 0 -O L)
    f = str(tmp_path / "init.mch")
    mch.savemch(f, first["detp"], first["seeds"], maxmedia=1, totalphoton=N, normalizer=first["normalizer"])
    rp = mch.replay_inputs(mch.loadmch(f), np.vstack([[0, 0, 1, 1], med]))
    n = rp["nphoton"]
    del first
    kw2 = {k: v for k, v in kw.items() if k not in ("seed", "nphoton", "issaveseed", "maxdetphoton")}
    kw2.update(outputtype=cases.WL, minenergy=0.0, isnormalized=0, maxdetphoton=8000000)
    r = orc.run_ref(node, elem, et, med, cuda=True, timeout=900, keep_dir=str(tmp_path), nphoton=n, seed=1, extra_args=["-E", "init.mch", "-P", "0"], **kw2)
    cfg = _cfg(node, elem, et, med, **kw2)
    cfg.update(replayseed=rp["replayseed"], replayweight=rp["replayweight"], replaytime=rp["replaytime"])
    g = mmc.run(cfg)
    ours = np.where(np.isfinite(g["raw"][..., 0]), g["raw"][..., 0], 0).sum(axis=0)
    ref = np.where(np.isfinite(r["field_flat"]), r["field_flat"], 0).reshape(1, -1).sum(axis=0)
    assert ours.shape == ref.shape
    lit = ref > 1e-3 * ref.max()
    rel = np.abs(ours[lit] - ref[lit]) / ref[lit]
    print("replay -O L of %d photons: total ratio %.6f, %d lit elements, median %.2e, p99 %.2e, max %.2e; kernel %.0f ms vs reference %.0f ms"
          % (n, ours.sum() / ref.sum(), lit.sum(), np.median(rel), np.percentile(rel, 99), rel.max(), g["kernel_ms"], r.get("kernel_ms", float("nan"))))
    assert lit.sum() > 20000
    assert abs(ours.sum() / ref.sum() - 1) < 1e-3
    assert np.median(rel) < 2e-3 and np.percentile(rel, 99) < 0.02, (np.median(rel), np.percentile(rel, 99))
    assert abs(len(g["detp"]) - n) <= 0.02 * n                                  # the replayed photons are detected again
