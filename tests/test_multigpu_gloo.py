"""World-size-2 test of the multi-GPU epilogue on CPU (gloo): photon split, sum-reduce of the volume and energy
tallies, counts-then-payload gather of detected-photon rows with truncation at maxdetphoton.  The per-rank
"simulation" here is the CPU oracle (test infrastructure) -- the point is the exchange logic of
mmc_b200/multigpu.py, which on the GPU box runs unchanged over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import orc
from mmc_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_run(rank, share, world):
    node, elem, et, med = cases.two_media_cube()
    kw = cases.case_kwargs("blb_detectors")
    kw.update(nphoton=int(share[rank]), issaveseed=1)
    # rank r draws seed slice r of the host stream: emulate with a per-rank seed
    kw["seed"] = kw["seed"] + rank
    o = orc.run(node, elem, et, med, nthread=1, gpu_semantics=1, **kw)
    return o


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        share, offs = multigpu.split_photons(3001, [1.0] * world)
        o = _local_run(rank, share, world)
        n = o["detectedcount"]
        local = dict(field=o["field"], energytot=o["launchweight"], energyesc=o["escweight"], raytet=o["raytet"],
                     detp=o["detected"][:n], seeds=o["detseed"][:n], maxdetphoton=40, reclen=o["reclen"])
        out = multigpu.reduce_results(local, dist)
        if rank == 0:
            q.put(dict(field=out["field"].numpy(), energytot=out["energytot"], energyesc=out["energyesc"], raytet=out["raytet"],
                       detp=out["detp"], seeds=out["seeds"], detectedtotal=out["detectedtotal"], share=share, offs=offs))
    finally:
        dist.destroy_process_group()


def test_split_photons_follows_reference_rule():
    share, offs = multigpu.split_photons(10, [1, 1, 1])
    assert share.sum() == 10 and list(offs) == [0, share[0], share[0] + share[1]]
    share, _ = multigpu.split_photons(1000, [3, 1])
    assert list(share) == [750, 250]
    with pytest.raises(ValueError):
        multigpu.split_photons(10, [1, 0])


@pytest.mark.timeout(300)
def test_reduce_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    share = got["share"]
    assert share.sum() == 3001
    parts = [_local_run(r, share, world) for r in range(world)]
    np.testing.assert_allclose(got["field"], parts[0]["field"] + parts[1]["field"], rtol=1e-12)
    np.testing.assert_allclose(got["energytot"], parts[0]["launchweight"] + parts[1]["launchweight"])
    np.testing.assert_allclose(got["energyesc"], parts[0]["escweight"] + parts[1]["escweight"])
    assert got["raytet"] == parts[0]["raytet"] + parts[1]["raytet"]
    n0, n1 = parts[0]["detectedcount"], parts[1]["detectedcount"]
    assert got["detectedtotal"] == n0 + n1
    rows = np.concatenate([parts[0]["detected"][:n0], parts[1]["detected"][:n1]])[:40]
    assert got["detp"].shape == rows.shape          # truncated at maxdetphoton like src/mmc_cu_host.cu:823-834
    np.testing.assert_array_equal(got["detp"], rows)
    seeds = np.concatenate([parts[0]["detseed"][:n0], parts[1]["detseed"][:n1]])[:40]
    np.testing.assert_array_equal(got["seeds"].reshape(-1, 2), seeds.reshape(-1, 2))
