"""Adjoint Jacobians against the reference's OWN CUDA path on the same B200.

The reference computes the adjoint Jacobians on the GPU only (src/mmc_cu_host.cu:997-1395, kernels src/mmc_core.cl:2218-2649) and its
command-line program writes them with mesh_savejacob (src/mmc_mesh.c:1817-1960, `<session>_jmua.jnii`, `<session>_jd.jnii`) when the
output format is JNIfTI.  The stock program cannot run this mode: it sizes its result volume before mcx_prep appends the detector slots and dies of
heap corruption (measured here: SIGSEGV after the first slot's summary line).  oracle/_ref/mmc_refcuda_ms is the same unmodified
host + kernel objects (sm_100, oracle/Makefile.ref) behind oracle/ref_multislot_main.c, a main() that re-sizes that one buffer the
way the mmclab/pmmc containers get it (src/mmc_mesh.c:2389-2394) and changes nothing else.  It is run with `-O w` (J_mua + J_D) on the two-media cube with two detectors, whose directions reach the program through the root-level key
"Optode.Detector.Dir" of a `-j` overlay (src/mmc_utils.c:2347-2372; mcx_prep turns the detectors into disk sources, :3760-3797).
Both sides are Monte Carlo estimates with independent seed-to-photon mappings, so the comparison is statistical: a Jacobian entry is
(minus) a product of two fluences, each carrying its own noise.  One time gate: mesh_savejacob announces a gate axis but the buffer it
writes holds the gate-summed Jacobian only."""
import json
import os

import numpy as np
import pytest

import cases
import orc

pytestmark = pytest.mark.gpu

mmc = pytest.importorskip("mmc_b200")
needs_refcuda = pytest.mark.skipif(not orc.ref_available(multislot=True), reason="oracle/_ref/mmc_refcuda_ms not built")

DETS = [(10.3, 8.4, 0.0, 1.0), (11.7, 12.4, 20.0, 1.0)]
DETDIR = [(0, 0, 1, 0), (0, 0, -1, 0)]
OVERLAY = json.dumps({"Optode.Detector": 1, "Optode.Detector.Dir": [list(d) for d in DETDIR]})


def _stats(ours, ref, lit_frac):
    lit = np.abs(ref) > lit_frac * np.abs(ref).max()
    rel = np.abs(ours[lit] - ref[lit]) / np.abs(ref[lit])
    cc = float(np.corrcoef(ours.ravel(), ref.ravel())[0, 1])
    return int(lit.sum()), float(np.median(rel)), float(np.percentile(rel, 90)), float(ours[lit].sum() / ref[lit].sum()), cc


def _ref_jacobians(method, basisorder, N, tmp, flag="w", omega=0.0):
    node, elem, et, med = cases.two_media_cube()
    kw = dict(nphoton=N, seed=1648335518, srcpos=(10.1, 10.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=5e-9, isreflect=1,
              method=method, basisorder=basisorder, steps=1.0, detpos=DETS,
              e0=int(mmc.mesh_initelem(node, elem, (10.1, 10.2, 0.0))[0]))
    overlay = OVERLAY
    if omega:       # one -j fragment only (the last one wins, src/mmc_utils.c:4088-4097); it resets the gates it does not name
        overlay = json.dumps(dict(json.loads(OVERLAY), Forward={"T0": 0, "T1": 5e-9, "Dt": 5e-9, "N0": 1, "Omega": omega}))
    r = orc.run_ref(node, elem, et, med, cuda=True, multislot=True, timeout=600, check=False, keep_dir=str(tmp), expect="out_jmua.jnii",
                    extra_args=["-O", flag, "-F", "jnii", "-j", overlay], **kw)
    from mmc_b200 import volio
    jm = volio.loadjnii(os.path.join(str(tmp), "out_jmua.jnii"))["vol"]
    jd = volio.loadjnii(os.path.join(str(tmp), "out_jd.jnii"))["vol"] if flag == "w" else np.zeros(0)
    return (node, elem, et, med, kw), r, np.asarray(jm, np.float64), np.asarray(jd, np.float64)


@needs_refcuda
def test_grid_adjoint_jacobians_vs_reference_cuda(tmp_path):
    N = 20000000
    (node, elem, et, med, kw), r, jm, jd = _ref_jacobians(cases.GRID, 0, N, tmp_path)
    g = mmc.run(dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), method="grid", steps=(1.0, 1.0, 1.0),
                     outputtype="adjointmuad", detdir=DETDIR, **{k: v for k, v in kw.items() if k not in ("method", "steps")}))
    J = g["jacob"]                              # [J_mua, J_D][pair][voxel]
    assert (g["adj_ns"], g["adj_nd"]) == (1, 2)
    jm, jd = jm.reshape(2, -1), jd.reshape(2, -1)
    assert jm.shape == J[0].shape, (jm.shape, J[0].shape)
    for pair in range(2):
        n, med_, p90, ratio, cc = _stats(J[0][pair], jm[pair], 1e-2)
        print("grid J_mua pair %d: %d lit voxels, median %.4f, p90 %.4f, sum ratio %.5f, corr %.5f" % (pair, n, med_, p90, ratio, cc))
        assert n > 50
        assert med_ < 0.03 and p90 < 0.10 and abs(ratio - 1) < 0.01 and cc > 0.999, (med_, p90, ratio, cc)
        if not jd[pair].any():
            # measured on this build (B200, sm_100): the reference program writes an all-zero grid J_D, alone (-O d) as well as in the
            # dual types, while its mesh-mode J_D (next test) is fine; the grid J_D of this engine is held to the numpy restatement of
            # mmc_adjoint_dcoeff_kernel instead (tests/test_adjoint_rf.py)
            print("grid J_D   pair %d: the reference wrote zeros; ours spans [%.4g, %.4g]" % (pair, J[1][pair].min(), J[1][pair].max()))
            assert np.abs(J[1][pair]).max() > 0
            continue
        # J_D is a product of finite-difference gradients of two noisy fluences: compared through its strong entries
        n, med_, p90, ratio, cc = _stats(J[1][pair], jd[pair], 5e-2)
        print("grid J_D   pair %d: %d strong voxels, median %.4f, p90 %.4f, sum ratio %.5f, corr %.5f" % (pair, n, med_, p90, ratio, cc))
        assert n > 10
        assert med_ < 0.10 and abs(ratio - 1) < 0.05 and cc > 0.98, (med_, p90, ratio, cc)


@needs_refcuda
def test_mesh_adjoint_jacobians_vs_reference_cuda(tmp_path):
    """Mesh mode (nodal fluence, basisorder 1): full-FEM J_mua and J_D per node (rb_femjacobian formulas, src/mmc_core.cl:2440-2649)."""
    N = 20000000
    (node, elem, et, med, kw), r, jm, jd = _ref_jacobians(cases.BLBADOUEL, 1, N, tmp_path)
    g = mmc.run(dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), method="elem",
                     outputtype="adjointmuad", detdir=DETDIR, **{k: v for k, v in kw.items() if k not in ("method", "steps")}))
    J = g["jacob"]
    jm, jd = jm.reshape(2, -1), jd.reshape(2, -1)
    assert jm.shape == J[0].shape == (2, len(node)), (jm.shape, J[0].shape)
    for pair in range(2):
        n, med_, p90, ratio, cc = _stats(J[0][pair], jm[pair], 1e-2)
        print("mesh J_mua pair %d: %d lit nodes, median %.4f, p90 %.4f, sum ratio %.5f, corr %.5f" % (pair, n, med_, p90, ratio, cc))
        assert n > 20
        assert med_ < 0.03 and p90 < 0.10 and abs(ratio - 1) < 0.01 and cc > 0.999, (med_, p90, ratio, cc)
        n, med_, p90, ratio, cc = _stats(J[1][pair], jd[pair], 5e-2)
        print("mesh J_D   pair %d: %d strong nodes, median %.4f, p90 %.4f, sum ratio %.5f, corr %.5f" % (pair, n, med_, p90, ratio, cc))
        assert n > 5
        assert med_ < 0.10 and abs(ratio - 1) < 0.05 and cc > 0.98, (med_, p90, ratio, cc)


@needs_refcuda
def test_grid_rf_adjoint_jmua_vs_reference_cuda(tmp_path):
    """RF run (omega = 2 pi 200 MHz): the complex J_mua = -V phi_s phi_d is the one output of the reference program that carries the
    IMAGINARY fluence (mesh_saveweight drops it, mesh_savejacob writes [Re, Im] on a trailing axis, src/mmc_mesh.c:1886-1925), so
    this pins the imaginary part of the complex deposit (src/mmc_core.cl:1043-1078) against the reference's own kernel."""
    N = 20000000
    omega = 2 * np.pi * 2e8
    (node, elem, et, med, kw), r, jm, _ = _ref_jacobians(cases.GRID, 0, N, tmp_path, flag="a", omega=omega)
    g = mmc.run(dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), method="grid", steps=(1.0, 1.0, 1.0),
                     outputtype="adjoint", detdir=DETDIR, omega=omega, **{k: v for k, v in kw.items() if k not in ("method", "steps")}))
    J = g["jacob"]                              # [Re J_mua, Im J_mua][pair][voxel]
    jm = jm.reshape(2, 2, -1)                   # [Re, Im][pair][voxel]
    assert J.shape == jm.shape, (J.shape, jm.shape)
    for part, name in ((0, "Re"), (1, "Im")):
        for pair in range(2):
            n, med_, p90, ratio, cc = _stats(J[part][pair], jm[part][pair], 1e-2)
            print("RF grid %s J_mua pair %d: %d lit voxels, median %.4f, p90 %.4f, sum ratio %.5f, corr %.5f" % (name, pair, n, med_, p90, ratio, cc))
            assert n > 50
            assert med_ < 0.05 and p90 < 0.15 and abs(ratio - 1) < 0.02 and cc > 0.995, (name, pair, med_, p90, ratio, cc)
    # the imaginary part is a real signal, not noise around zero
    assert np.abs(jm[1]).max() > 0.05 * np.abs(jm[0]).max()
