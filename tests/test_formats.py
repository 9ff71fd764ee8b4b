"""Result-side file formats (.mch detected-photon history, .bin volumes) and the replay preparation of
mesh_loadseedfile (src/mmc_mesh.c:815-898)."""
import os

import numpy as np

from mmc_b200 import mch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_written_header_parses():
    """tests/golden/ref_blb_detectors.mch was written by the reference CPU binary (case blb_detectors).  In this version
    of the reference the CPU driver never fills his.colcount/savedphoton (src/mmc_host.c:413-421 passes the seed size as
    `doappend`), so the file is the 64-byte header only -- enough to pin the header layout."""
    h = mch.loadmch(os.path.join(GOLD, "ref_blb_detectors.mch"))
    assert h["version"] == 1 and h["maxmedia"] == 2 and h["srcnum"] == 1 and h["respin"] == 1
    assert abs(h["normalizer"] - 4365.69) < 0.01 and h["unitinmm"] == 1.0
    assert h["savedphoton"] == 0 and h["detp"].shape == (0, 0)


def test_mch_round_trip_and_replay_inputs(tmp_path):
    rs = np.random.RandomState(1)
    M, n = 2, 50
    detp = np.zeros((n, 2 + 2 * M), dtype=np.float32)            # detid, nscat[M], ppath[M], w0
    detp[:, 0] = rs.randint(1, 3, n)
    detp[:, 1:1 + M] = rs.randint(0, 50, (n, M))
    detp[:, 1 + M:1 + 2 * M] = rs.rand(n, M) * 30
    detp[:, -1] = 1.0
    seeds = rs.randint(1, 2**62, size=(n, 2)).astype(np.uint64)
    f = str(tmp_path / "s.mch")
    mch.savemch(f, detp, seeds, maxmedia=M, totalphoton=1000, detnum=2, normalizer=3.5)
    assert os.path.getsize(f) == 64 + detp.nbytes + seeds.nbytes
    h = mch.loadmch(f)
    assert h["colcount"] == detp.shape[1] and h["savedphoton"] == n and h["seedbyte"] == 16 and h["detnum"] == 2
    assert np.array_equal(h["detp"], detp) and np.array_equal(h["seeds"], seeds)
    prop = np.array([[0, 0, 1, 1], [0.01, 1, 0.9, 1.37], [0.02, 2, 0.9, 1.5]], dtype=np.float32)
    r = mch.replay_inputs(h, prop, replaydet=2)
    sel = detp[:, 0] == 2
    assert r["nphoton"] == sel.sum() and np.array_equal(r["replayseed"], seeds[sel])
    w = np.exp(-(0.01 * detp[sel, 3] + 0.02 * detp[sel, 4]))
    t = (1.37 * detp[sel, 3] + 1.5 * detp[sel, 4]) * 3.335640951981520e-12
    np.testing.assert_allclose(r["replayweight"], w, rtol=1e-5)
    np.testing.assert_allclose(r["replaytime"], t, rtol=1e-5)

    # legacy_columns: the columns mesh_loadseedfile itself reads, j = 2 .. maxmedia+1 (src/mmc_mesh.c:874-882); for maxmedia = 2 that is
    # (nscat_2, ppath_1) -- the reference's own (mis-)weights, needed to replay a multi-media file in both programs on equal terms
    rl = mch.replay_inputs(h, prop, replaydet=2, legacy_columns=True)
    wl = (detp[sel, -1] * np.exp(-prop[1, 0] * detp[sel, 2]) * np.exp(-prop[2, 0] * detp[sel, 3])).astype(np.float32)
    np.testing.assert_allclose(rl["replayweight"], wl, rtol=1e-6)
    np.testing.assert_allclose(rl["replaytime"], (prop[1, 3] * detp[sel, 2] + prop[2, 3] * detp[sel, 3]) * mch.R_C0, rtol=1e-6)
    one = mch.replay_inputs(dict(h, maxmedia=1, detp=detp[:, [0, 1, 3, 4]]), prop[:2])      # maxmedia == 1: both readings coincide
    leg = mch.replay_inputs(dict(h, maxmedia=1, detp=detp[:, [0, 1, 3, 4]]), prop[:2], legacy_columns=True)
    assert np.array_equal(one["replayweight"], leg["replayweight"]) and np.array_equal(one["replaytime"], leg["replaytime"])


def test_bin_round_trip(tmp_path):
    a = np.arange(24, dtype=np.float64).reshape(3, 8)
    f = str(tmp_path / "v.bin")
    mch.savebin(f, a)
    assert np.array_equal(mch.loadbin(f, 3), a)


def test_nii_and_jnii_volume_round_trip(tmp_path):
    """`-F nii` / `-F jnii` volumes (mcx_savenii src/mmc_utils.c:515-611, mcx_savejnii :787-905): header fields the reference
    writes (dims x,y,z,t; voxel size; gate width in microseconds; float64 payload at offset 352) and a loss-less round trip."""
    import json
    import struct

    from mmc_b200 import volio
    rs = np.random.RandomState(2)
    vol = rs.rand(3, 4, 5, 6)                                  # [gate, z, y, x]
    p = str(tmp_path / "v.nii")
    volio.savenii(p, vol, steps=(0.5, 0.5, 0.5), tstep=1e-10)
    raw = open(p, "rb").read()
    assert len(raw) == 352 + vol.size * 8
    assert struct.unpack_from("<i", raw, 0)[0] == 348 and raw[344:348] == b"n+1\0"
    assert struct.unpack_from("<8h", raw, 40) == (4, 6, 5, 4, 3, 0, 0, 0)
    assert struct.unpack_from("<2h", raw, 70) == (64, 64)       # NIFTI_TYPE_FLOAT64, bitpix
    assert struct.unpack_from("<f", raw, 108)[0] == 352.0
    assert abs(struct.unpack_from("<8f", raw, 76)[4] - 1e-4) < 1e-9   # tstep in microseconds
    r = volio.loadnii(p)
    assert np.array_equal(r["vol"], vol) and r["steps"] == (0.5, 0.5, 0.5)
    p = str(tmp_path / "v.jnii")
    volio.savejnii(p, vol.astype(np.float32), steps=(0.5, 0.5, 0.5), tstep=1e-10)
    j = json.load(open(p))
    assert j["NIFTIHeader"]["Dim"] == [6, 5, 4, 3] and j["NIFTIHeader"]["DataType"] == "single"
    assert j["NIFTIData"]["_ArrayZipType_"] == "zlib" and j["NIFTIData"]["_ArraySize_"] == [6, 5, 4, 3]
    assert np.array_equal(volio.loadjnii(p)["vol"], vol.astype(np.float32))
