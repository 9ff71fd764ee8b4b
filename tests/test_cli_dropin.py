"""Drop-in boundary: the reference's UNMODIFIED host and command line (mmc.c, mmc_host.c, mmc_mesh.c, mmc_utils.c, ...)
linked against integration/mmc_cu_host_b200.cpp + libmmc_b200.so instead of src/mmc_cu_host.cu
(oracle/Makefile.ref target `b200cli` -> oracle/_ref/mmc_b200cli).  `-c cuda` then runs the B200 engine behind
mmc_run_cu(); `-c sse` in the same binary is the reference's own CPU path."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "mmc_b200cli")
needs_cli = pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/mmc_b200cli not built (needs /root/reference at build time)")


def _run(args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd="/tmp")


def _absorbed(out):
    m = re.findall(r"absorbed:\s*(?:\x1b\[[0-9;]*m)*\s*([0-9.]+)%", out)
    assert m, out[-2000:]
    return float(m[-1]) / 100.0


@needs_cli
def test_cli_fails_with_the_reference_error_convention_without_a_gpu():
    import mmc_b200
    if mmc_b200.gpuinfo():
        pytest.skip("a GPU is present")
    r = _run(["--bench", "dmmc-cube60", "-c", "cuda", "-n", "1000", "-D", "T", "-S", "0"])
    assert r.returncode != 0
    assert "MMC ERROR(-1):No GPU device found" in (r.stdout + r.stderr)


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("bench", ["dmmc-cube60", "dmmc-cube60b"])
def test_reference_cli_drives_the_b200_engine(bench):
    """Built-in benchmarks of the reference binary (src/mmc_bench.c:41-110; the CI smoke test of the reference,
    .github/workflows/build_all.yml:151-160): same command line, `-c cuda` (our engine) vs `-c sse` (reference CPU)."""
    gpu = _run(["--bench", bench, "-c", "cuda", "-n", "1e6", "-D", "T", "-S", "0"])
    assert gpu.returncode == 0, gpu.stdout[-2000:] + gpu.stderr[-2000:]
    assert "MMC-B200" in gpu.stdout and "MCX simulation speed" in gpu.stdout
    cpu = _run(["--bench", bench, "-c", "sse", "-n", "1e6", "-D", "T", "-S", "0"])
    assert cpu.returncode == 0, cpu.stdout[-2000:] + cpu.stderr[-2000:]
    fg, fc = _absorbed(gpu.stdout), _absorbed(cpu.stdout)
    assert abs(fg - fc) < 2.5e-3, (fg, fc)          # 1e6 photons: sigma ~ 4e-4


@needs_cli
@pytest.mark.gpu
def test_reference_cli_writes_adjoint_jacobians_through_the_b200_engine(tmp_path):
    """`-O w` (J_mua + J_D) through the unmodified reference command line: input parsing, mcx_prep's detector slots, mesh_savejacob's
    JNIfTI files are the reference's; the engine and the post-kernels are ours.  The stock program cannot finish this run (it sizes its
    volume before the slots exist, see tests/test_gpu_vs_reference_adjoint.py); the stub re-sizes it.  The files must hold what the
    Python host returns for the same problem and seed: same engine, same host seed stream, so the agreement is to rounding."""
    import json
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import mmc_b200 as mmc
    import orc
    from mmc_b200 import volio
    node, elem, et, med = cases.two_media_cube()
    dets = [(10.3, 8.4, 0.0, 1.0), (11.7, 12.4, 20.0, 1.0)]
    detdir = [(0, 0, 1, 0), (0, 0, -1, 0)]
    e0 = int(mmc.mesh_initelem(node, elem, (10.1, 10.2, 0.0))[0])
    kw = dict(nphoton=1000000, seed=1648335518, srcpos=(10.1, 10.2, 0.0), srcdir=(0, 0, 1), tstart=0.0, tend=5e-9, tstep=5e-9, isreflect=1,
              basisorder=0, detpos=dets, e0=e0)
    orc.write_mesh_files(str(tmp_path), "t", node, elem, et, med, None)
    with open(os.path.join(str(tmp_path), "in.inp"), "w") as f:
        f.write("1000000\n1648335518\n10.1 10.2 0\n0 0 1 0\n0 5e-9 5e-9\nt\n%d\n2 1\n" % e0)
        for d in dets:
            f.write("%.9g %.9g %.9g %.9g\n" % d)
    overlay = json.dumps({"Optode.Detector": 1, "Optode.Detector.Dir": [list(d) for d in detdir]})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([CLI, "-f", "in.inp", "-s", "out", "-M", "g", "--gridsize", "1", "-b", "1", "-C", "0", "-U", "1", "-O", "w", "-F", "jnii",
                        "-D", "T", "-S", "1", "-n", "1000000", "-E", "1648335518", "-c", "cuda", "-G", "1", "-j", overlay],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MMC-B200" in r.stdout
    jm = volio.loadjnii(os.path.join(str(tmp_path), "out_jmua.jnii"))["vol"].reshape(2, -1)
    jd = volio.loadjnii(os.path.join(str(tmp_path), "out_jd.jnii"))["vol"].reshape(2, -1)
    g = mmc.run(dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), method="grid", steps=(1.0, 1.0, 1.0),
                     outputtype="adjointmuad", detdir=detdir, **kw))
    J = g["jacob"]
    assert jm.shape == J[0].shape and np.abs(jm).max() > 0 and np.abs(jd).max() > 0
    # both are Monte Carlo runs of the same engine; the thread-to-photon assignment of the dynamic pool is not reproducible, so the
    # comparison is statistical at 1e6 photons (bright voxels)
    for ours, ref, name in ((J[0], jm, "J_mua"), (J[1], jd, "J_D")):
        for pair in range(2):
            lit = np.abs(ref[pair]) > 0.05 * np.abs(ref[pair]).max()
            rel = np.abs(ours[pair][lit] - ref[pair][lit]) / np.abs(ref[pair][lit])
            cc = np.corrcoef(ours[pair], ref[pair])[0, 1]
            print("%s pair %d: %d voxels, median %.4f, corr %.5f" % (name, pair, lit.sum(), np.median(rel), cc))
            assert lit.sum() > 5 and np.median(rel) < 0.15 and cc > 0.97, (name, pair, np.median(rel), cc)


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("flag", ["h", "p"])
def test_cli_offers_havel_and_plucker_on_the_gpu(flag):
    """`-M h` / `-M p` with `-c cuda`: the reference host coerces the tracer to a branch-less Badouel one for every GPU run
    (mcx_validatecfg, src/mmc_utils.c:3542-3544) before mmc_run_cu sees cfg->method, so the user's choice reaches the stub through
    MMC_B200_METHOD (or through the one-line host patch of INTEGRATION.md, which makes the variable unnecessary).  Same binary:
    `-c cuda` runs this engine's Havel / Plucker kernels, `-c sse` the reference's CPU tracers (their only implementation)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1), MMC_B200_METHOD=flag)
    args = ["--bench", "dmmc-cube60", "-n", "1e6", "-D", "T", "-S", "0", "-M", flag, "-C", "0"]
    gpu = subprocess.run([CLI] + args + ["-c", "cuda"], capture_output=True, text=True, timeout=600, env=env, cwd="/tmp")
    assert gpu.returncode == 0, gpu.stdout[-2000:] + gpu.stderr[-2000:]
    assert "MMC-B200" in gpu.stdout and ("tracer: %s" % {"h": "Havel", "p": "Plucker"}[flag]) in gpu.stdout
    cpu = subprocess.run([CLI] + args + ["-c", "sse"], capture_output=True, text=True, timeout=600, env=env, cwd="/tmp")
    assert cpu.returncode == 0, cpu.stdout[-2000:] + cpu.stderr[-2000:]
    assert abs(_absorbed(gpu.stdout) - _absorbed(cpu.stdout)) < 2.5e-3
    m = re.search(r"ray-tet (\d+)", gpu.stdout)
    assert m and abs(int(m.group(1)) / 1e6 - 83.8) < 2.0            # Havel on the 6-tet benchmark mesh: 83.8 tests per photon


@needs_cli
@pytest.mark.gpu
def test_cli_length_unit_is_applied_once():
    """`-u 0.5`: the reference host multiplies mua/mus by the unit before mmc_run_cu (src/mmc_mesh.c:542-546); the stub must hand the
    engine media that end up scaled ONCE.  Same binary, same command line: `-c cuda` (this engine) against `-c sse` (reference CPU)
    and against the CPU oracle."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import orc
    node, elem, et, med = cases.two_media_cube()
    kw = cases.case_kwargs("blb_elem_reflect")
    kw.update(nphoton=300000, unitinmm=0.5)
    gpu = orc.run_ref(node, elem, et, med, binary=CLI, cuda=True, **kw)
    cpu = orc.run_ref(node, elem, et, med, binary=CLI, nthread=os.cpu_count() or 1, **kw)
    assert "MMC-B200" in gpu["log"]
    one = orc.run_ref(node, elem, et, med, binary=CLI, cuda=True, **dict(kw, unitinmm=1.0))
    assert abs(gpu["absorbed_frac"] - cpu["absorbed_frac"]) < 4e-3, (gpu["absorbed_frac"], cpu["absorbed_frac"])
    assert abs(one["absorbed_frac"] - gpu["absorbed_frac"]) > 0.02          # the unit matters on this problem
    fg, fc = gpu["field_flat"], cpu["field_flat"]
    fg, fc = np.where(np.isfinite(fg), fg, 0), np.where(np.isfinite(fc), fc, 0)
    assert abs(fg.sum() / fc.sum() - 1) < 0.02                              # normalisation carries unitinmm^3 (src/mmc_mesh.c:2206)


@needs_cli
@pytest.mark.gpu
def test_cli_widefield_source_and_detector_labels_survive_the_host():
    """Planar source over tets labelled -1 and a wide-field detector layer labelled -2 (examples/replaywide/createmesh.m:11-29): the
    reference host moves the -1 labels into mesh->srcelem and turns -2 into prop+1 before mmc_run_cu (src/mmc_mesh.c:390-427); the stub
    restores both for the engine.  Checked against the reference CPU path of the same binary: absorbed fraction, detected count."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import orc
    node, elem, et, med = cases.case_mesh("planar_widedet")
    kw = cases.case_kwargs("planar_widedet")
    kw.update(nphoton=300000)
    gpu = orc.run_ref(node, elem, et, med, binary=CLI, cuda=True, **kw)
    cpu = orc.run_ref(node, elem, et, med, binary=CLI, nthread=os.cpu_count() or 1, **kw)
    assert "MMC-B200" in gpu["log"]
    assert abs(gpu["absorbed_frac"] - cpu["absorbed_frac"]) < 4e-3, (gpu["absorbed_frac"], cpu["absorbed_frac"])
    ng, nc = gpu["detectedcount"], cpu["detectedcount"]
    assert nc > 1000 and abs(ng - nc) < 6 * nc ** 0.5 + 5, (ng, nc)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference headers (this container only)")
@pytest.mark.parametrize("flavour", [[], ["-DMCX_CONTAINER"]])
def test_stub_compiles_against_the_reference_headers(flavour, tmp_path):
    """integration/mmc_cu_host_b200.cpp against the reference's own headers, as the stand-alone program builds it and as the
    mmclab / pmmc containers do (MCX_CONTAINER: no file output, errors as exceptions).  It must export exactly the two symbols the
    reference host links against (src/mmc_cu_host.h:46-62) and pull nothing but the C-ABI from this repository."""
    ref = "/root/reference/src"
    clh = tmp_path / "mmc_core.clh"
    clh.write_text("unsigned char mmc_core_cl[] = {0};\n")
    obj = str(tmp_path / "stub.o")
    cmd = ["g++", "-c", "-std=c++11", "-O1", "-w", "-fopenmp", "-msse4.1", "-DUSE_OS_TIMER", "-DMMC_XORSHIFT", "-DMCX_EMBED_CL", "-DMMC_USE_SSE",
           "-DHAVE_SSE2", "-DUSE_CUDA"] + flavour + ["-I", str(tmp_path), "-I", ref, "-I", ref + "/ubj", "-I", ref + "/zmat",
           "-I", os.path.join(ROOT, "include"), "-o", obj, os.path.join(ROOT, "integration", "mmc_cu_host_b200.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    nm = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    assert {"mmc_run_cu", "mcx_list_cu_gpu"} <= exported, exported
    ours = {l.split()[-1] for l in nm.splitlines() if " U " in l and l.split()[-1].startswith("mmcb_")}
    hdr = open(os.path.join(ROOT, "include", "mmc_b200.h")).read()
    assert ours and all(re.search(r"\b%s\s*\(" % s, hdr) for s in ours), ours
