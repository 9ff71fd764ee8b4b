"""The high-count reference-CPU fixtures are what the 1e8-photon GPU tests stand on (tests/test_gpu_c1_1e8.py, tests/test_gpu_c2_1e8.py): they must
be the files tools/make_golden_1e8.py writes at north_star's photon count, not a trial run of it."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_c1_fixture_is_the_1e8_run():
    z = np.load(os.path.join(GOLD, "ref_c1_1e8.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    for tracer in ("havel", "plucker"):
        assert meta[tracer]["nphoton"] == 100000000
        assert 0.17 < meta[tracer]["absorbed_frac"] < 0.18          # cube60, mua 0.005: reference CPU anchors 17.70 % / 17.69 %
        cw = z[tracer + "/cw"]
        assert cw.shape == (29791,) and np.isfinite(cw).all() and cw.max() > 0
        assert z[tracer + "/gatesum"].shape == (50,)


def test_c2_fixture_is_the_1e8_run():
    z = np.load(os.path.join(GOLD, "ref_c2_1e8.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    assert meta["nphoton"] == 100000000 and meta["nvox"] == 61 * 61 * 61
    assert abs(meta["absorbed_frac"] - 0.5162) < 5e-4                # reference CUDA kernel on the same workload: 0.51619
    idx, cw = z["idx"], z["cw"]
    assert idx.shape == cw.shape and idx.max() < meta["nvox"] and np.all(np.diff(idx.astype(np.int64)) > 0)
    assert (cw > 1e-3 * cw.max()).sum() > 5000                       # the voxels north_star's tolerance is about
    g = z["gatesum"]
    assert g.shape == (10,) and np.all(np.diff(g) < 0)               # the fluence decays from gate to gate
