"""CPU-side tests of the host layer and the C-ABI library (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
import orc

import mmc_b200
from mmc_b200 import api, meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mmc_b200.h")).read()
    declared = set(re.findall(r"\b(mmcb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"mmcb_last_error"} - {"mmcb_last_error"}
    assert declared, "no declarations found"
    L = ctypes.CDLL(api.LIBPATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(api.EXPORTS) <= declared


def test_version_and_gpu_listing_do_not_need_a_gpu():
    assert mmc_b200.version() == 0x00010000
    assert isinstance(mmc_b200.gpuinfo(), list)


def test_host_seeds_restatement_matches_glibc_rand():
    for seed in (0, 1, 1648335518, 0x623F9A9E & 0x7FFFFFFF, 12345):
        a = mmc_b200.host_seeds(seed, 500)
        assert np.array_equal(a, orc.host_seeds(seed, 500))
    assert np.array_equal(mmc_b200.host_seeds(99, 100, skip=1000), orc.host_seeds(99, 1100)[1000:])


def test_mesh_helpers_match_oracle():
    node, elem, et, med = cases.two_media_cube()
    o = orc.run(node, elem, et, med, **cases.case_kwargs("blb_elem_reflect"))
    e2, evol, nvol = mmc_b200.mesh_volumes(node, elem, et)
    assert np.array_equal(e2, o["elem"])
    assert np.array_equal(evol, o["evol"])
    fnb = mmc_b200.mesh_facenb(e2)
    assert np.array_equal(fnb, np.where(o["facenb"] > 0, o["facenb"], 0))
    assert (fnb == 0).sum() == o["nf"]
    e0, bary = mmc_b200.mesh_initelem(node, e2, (10.1, 10.2, 0.0))
    assert e0 == o["e0"]
    assert abs(bary.sum() - 1) < 1e-5


def test_cube60_mesh_matches_reference_sizes():
    node, elem, et = meshgen.cube60()
    assert node.shape == (29791, 3) and elem.shape == (135000, 4)      # SURVEY.md section 8 sizes
    assert (meshgen.tet_volume6(node, elem) > 0).all()
    assert abs(meshgen.tet_volume6(node, elem).sum() / 6 - 60 ** 3) < 1e-3
    fnb = mmc_b200.mesh_facenb(elem)
    assert (fnb == 0).sum() == 2 * 6 * 30 * 30                         # two triangles per boundary lattice cell face


def test_compute_fails_loudly_without_gpu():
    if len(mmc_b200.gpuinfo()) > 0:
        pytest.skip("a GPU is present")
    node, elem, et, med = cases.two_media_cube()
    cfg = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), nphoton=100,
               srcpos=(10.1, 10.2, 0), srcdir=(0, 0, 1), tstart=0, tend=5e-9, tstep=5e-10, basisorder=0)
    with pytest.raises(mmc_b200.MMCError, match="no CPU fallback"):
        mmc_b200.run(cfg)


def test_config_validation_mirrors_reference_messages():
    node, elem, et, med = cases.two_media_cube()
    cfg = dict(node=node, elem=elem, elemprop=et, prop=np.vstack([[0, 0, 1, 1], med]), nphoton=100,
               srcpos=(10.1, 10.2, 0), srcdir=(0, 0, 1), tstart=0, tend=5e-9, tstep=5e-10, basisorder=0)
    p = api.Problem(cfg)
    sz = p.sizes()
    assert (sz.maxgate, sz.datalen, sz.reclen) == (10, len(elem), 2 * 2 + 2)
    bad = dict(cfg, tstep=0.0)
    with pytest.raises(mmc_b200.MMCError, match="time gate"):
        api.Problem(bad).sizes()
    bad = dict(cfg, srcdir=(0, 0, 3))
    with pytest.raises(mmc_b200.MMCError, match="unitary"):
        api.Problem(bad).sizes()
    bad = dict(cfg, srcpos=(-5, 0, 0))
    with pytest.raises(mmc_b200.MMCError, match="does not enclose"):
        api.Problem(bad).sizes()
    for ot in ("jacobian", "wl", "wp"):          # src/mmc_utils.c:4266-4268: replay outputs need the seeds of an .mch file
        with pytest.raises(mmc_b200.MMCError, match="only valid in the reply mode"):
            api.Problem(dict(cfg, outputtype=ot)).sizes()
    g = api.Problem(dict(cfg, method="grid", steps=(0.5, 0.5, 0.5))).sizes()
    assert tuple(g.dim) == (41, 41, 41) and g.datalen == 41 ** 3        # mesh_createdualmesh, src/mmc_mesh.c:374-380


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/mmc_b200.h must be consumable from C (the reference host is C): compile a C99 program against it with gcc -pedantic,
    link it to the library, call the entry points that need no GPU, and compare sizeof() of every struct with the ctypes mirrors of
    mmc_b200/api.py (a field added on one side only would shift everything behind it)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "mmc_b200.h"
int main(void) {
    mmcb_config c; mmcb_mesh m; mmcb_output o; mmcb_sizes s; mmcb_devptrs d; mmcb_gpuinfo g;
    uint32_t seeds[8];
    (void)c; (void)m; (void)o; (void)s; (void)d; (void)g;
    mmcb_host_seeds(1648335518, 0, 8, seeds);
    printf("%d %zu %zu %zu %zu %zu %zu %u %d\n", mmcb_version(), sizeof(mmcb_config), sizeof(mmcb_mesh), sizeof(mmcb_output),
           sizeof(mmcb_sizes), sizeof(mmcb_devptrs), sizeof(mmcb_gpuinfo), seeds[0], mmcb_query_sizes(NULL, NULL, &s));
    printf("%s\n", mmcb_last_error());
    return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.dirname(api.LIBPATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-lmmc_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    v = out[0].split()
    assert int(v[0]) == 0x00010000
    sizes = [int(x) for x in v[1:7]]
    assert sizes == [ctypes.sizeof(t) for t in (api.Config, api.Mesh, api.Output, api.Sizes, api.DevPtrs, api.GpuInfo)]
    assert int(v[7]) == int(orc.host_seeds(1648335518, 1)[0])
    assert int(v[8]) < 0 and "null" in out[1]                          # error convention: negative id + message
