"""Multi-GPU behind the C-ABI: mmcb_run_multi (one host thread + one session per device, NCCL reduce of the volume, gather of the
detected-photon rows behind the first device's) and the drop-in stub's fan-out over cfg->deviceid / cfg->workload
(src/mmc_cu_host.cu:403-429,1538-1553).  The GPU tests need two devices (gpurun --gpus 2); on a one-GPU box they are skipped and the
host-side split rule is all that runs."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import orc
from test_gpu_parity import _cfg

mmc = pytest.importorskip("mmc_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "mmc_b200cli")


def test_photon_shares_follow_the_workload_rule():
    assert mmc.photon_shares(10, 3).tolist() == [3, 3, 4]                       # equal weights, the last device takes the remainder
    assert mmc.photon_shares(10, 2, [1, 3]).tolist() == [2, 8]                  # nphoton * w / sum(w), truncated
    assert mmc.photon_shares(10 ** 9, 8).sum() == 10 ** 9
    assert mmc.photon_shares(7, 4, [0, 0, 0, 0]).tolist() == [1, 1, 1, 4]       # unspecified (<= 0) weights count as 1


def _two_gpus():
    try:
        return len(mmc.gpuinfo()) >= 2
    except Exception:       # noqa: BLE001
        return False


needs2 = pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs (gpurun --gpus 2)")


@pytest.mark.gpu
@needs2
def test_run_multi_matches_single_gpu_and_oracle():
    node, elem, et, med = cases.two_media_cube()
    kw = cases.case_kwargs("blb_detectors")
    N = 400000
    kw.update(nphoton=N, issaveseed=1)
    cfg = _cfg(node, elem, et, med, **kw)
    one = mmc.run(cfg)
    two = mmc.run_multi(cfg, gpuids=[1, 2], workload=[1, 3])
    o = orc.run(node, elem, et, med, nthread=8, gpu_semantics=1, **kw)
    assert two["energytot"][0] == N == one["energytot"][0]
    fo = (o["launchweight"][0] - o["escweight"][0]) / o["launchweight"][0]
    f1, f2 = one["energyabs"][0] / N, two["energyabs"][0] / N
    sigma = np.sqrt(fo * (1 - fo) / N)
    assert abs(f2 - fo) < 6 * sigma + 3e-4 and abs(f1 - f2) < 8 * sigma + 3e-4, (f1, f2, fo)
    assert abs(two["raytet"] / one["raytet"] - 1) < 0.01
    assert abs(two["raw"].sum() / one["raw"].sum() - 1) < 0.01                  # normalised volume: reduced over both devices
    np.testing.assert_allclose(two["raw"].sum(axis=(1, 2))[:2], one["raw"].sum(axis=(1, 2))[:2], rtol=0.02)      # the gates that hold the light
    n1, n2 = len(one["detp"]), len(two["detp"])
    assert abs(n1 - n2) < 6 * np.sqrt(n1) + 5 and two["detectedtotal"] == n2
    assert set(np.unique(two["detp"][:, 0]).tolist()) <= {1.0, 2.0}
    seeds = np.ascontiguousarray(two["seeds"]).view(np.uint64).reshape(-1, 2)
    assert len(np.unique(seeds, axis=0)) == n2                                  # rows of both devices, no duplicates, no holes
    assert np.all(two["detp"][:, -1] > 0)


@pytest.mark.gpu
@needs2
def test_run_multi_truncates_detected_rows_like_the_reference():
    node, elem, et, med = cases.two_media_cube()
    kw = cases.case_kwargs("blb_detectors")
    kw.update(nphoton=200000, maxdetphoton=300)
    g = mmc.run_multi(_cfg(node, elem, et, med, **kw), gpuids=[1, 2])
    assert len(g["detp"]) == 300 and g["detectedtotal"] > 300                   # src/mmc_cu_host.cu:823-834: warning + truncation
    assert np.all(g["detp"][:, 0] >= 1)


@pytest.mark.gpu
@needs2
def test_run_multi_shards_a_replay_by_photon_index(tmp_path):
    """Replayed photons carry their own seeds: the two-GPU replay must reproduce the one-GPU replay photon for photon."""
    from mmc_b200 import mch
    node, elem, et, med = cases.case_mesh("planar_widedet")
    kw = cases.case_kwargs("planar_widedet")
    kw.update(nphoton=60000, issaveseed=1)
    first = mmc.run(_cfg(node, elem, et, med, **kw))
    f = str(tmp_path / "init.mch")
    mch.savemch(f, first["detp"], first["seeds"], maxmedia=len(med), totalphoton=kw["nphoton"], normalizer=first["normalizer"])
    rp = mch.replay_inputs(mch.loadmch(f), np.vstack([[0, 0, 1, 1], med]))
    kw2 = {k: v for k, v in kw.items() if k not in ("seed", "nphoton", "issaveseed")}
    kw2.update(outputtype=cases.WL, minenergy=0.0)
    cfg = _cfg(node, elem, et, med, **kw2)
    cfg.update(replayseed=rp["replayseed"], replayweight=rp["replayweight"], replaytime=rp["replaytime"])
    a = mmc.run(cfg)
    b = mmc.run_multi(cfg, gpuids=[1, 2])
    assert len(a["detp"]) == len(b["detp"])
    fa, fb = np.where(np.isfinite(a["raw"]), a["raw"], 0), np.where(np.isfinite(b["raw"]), b["raw"], 0)
    assert abs(fb.sum() / fa.sum() - 1) < 1e-6
    lit = fa > 1e-3 * fa.max()
    np.testing.assert_allclose(fb[lit], fa[lit], rtol=1e-5)
    ra, rb = a["detp"][np.lexsort(a["detp"].T[::-1])], b["detp"][np.lexsort(b["detp"].T[::-1])]
    np.testing.assert_allclose(ra, rb, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
@needs2
@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/mmc_b200cli not built")
def test_reference_cli_fans_out_over_two_gpus():
    """`-G 11 -W 1,2` through the unmodified reference command line: the stub hands both devices and the workload to mmcb_run_multi."""
    env = dict(os.environ, OMP_NUM_THREADS="2")

    def run(extra):
        r = subprocess.run([CLI, "--bench", "dmmc-cube60", "-c", "cuda", "-n", "1e6", "-D", "T", "-S", "0"] + extra, capture_output=True, text=True,
                           timeout=600, env=env, cwd="/tmp")
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        m = re.findall(r"absorbed:\s*(?:\x1b\[[0-9;]*m)*\s*([0-9.]+)%", r.stdout)
        return r.stdout, float(m[-1]) / 100.0

    out2, f2 = run(["-G", "11", "-W", "1,2"])
    out1, f1 = run(["-G", "1"])
    assert len(re.findall(r"- \[device \d+\(\d+\)", out2)) == 2 and "np=333333.0" in out2 and "np=666667.0" in out2
    assert abs(f1 - f2) < 2.5e-3, (f1, f2)
