#!/usr/bin/env python
"""bench.py -- photons/ms of the photon random-walk hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload sphshells|cube60|...]

One "step" = one pass of the hot path over one batch of synthetic input: `--photons` photons (default 1e7, the
BASELINE config C2 count) launched from the device-resident session.  `value` is device-timed (CUDA events, max over
ranks) with every input already in HBM; `e2e` is the same metric through the public one-call API with HOST buffers
(mesh upload, kernel, result download and normalisation inside the timed region).  N>1: one process per GPU under
torchrun, photons shard (weak scaling: every rank simulates `--photons`), the fluence volume is sum-reduced to rank 0
inside every timed step.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/mmc_ref when it was built, else the
oracle port) on the same workload, on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_STEP = {"elem": 92, "grid": 92}        # 84 B gathered + one fp32 atomic payload (RMW)


# --------------------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# --------------------------------------------------------------------------------------------------------------------
def workload(name, method=None):
    from mmc_b200 import meshgen
    gold = os.path.join(ROOT, "tests", "golden")
    if name == "sphshells":          # configs[1]: shipped dmmc_sphshells mesh + json (reflection on, 10 gates)
        z = np.load(os.path.join(gold, "sphshells_mesh.npz"))
        cfg = dict(node=z["node"], elem=z["elem"], elemprop=z["etype"], prop=np.vstack([[0, 0, 1, 1], z["prop"]]), evol=z["evol"],
                   srcpos=(30.0, 30.1, 0.0), srcdir=(0, 0, 1), e0=4916, tstart=0.0, tend=5e-9, tstep=5e-10,
                   isreflect=1, seed=1648335518, method=method or "grid", steps=(1.0, 1.0, 1.0), basisorder=0)
        desc = "examples/sphshells dmmc_sphshells mesh (3723 nodes/21256 tets, 4 media, n-mismatch, reflection on), pencil, 10 gates, RayTracer=%s" % cfg["method"]
    elif name == "cube60":           # configs[0]
        node, elem, et = meshgen.cube60()
        cfg = dict(node=node, elem=elem, elemprop=et, prop=[[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37]],
                   srcpos=(30.1, 30.2, 0.0), srcdir=(0, 0, 1), e0=4497, tstart=0.0, tend=5e-9, tstep=1e-10,
                   isreflect=0, seed=1648335518, method=method or "elem", basisorder=0)
        desc = "examples/validation cube60 (29791 nodes/135000 tets), mua=0.005 mus=1 g=0.01 n=1.37, pencil, 50 gates, RayTracer=%s" % cfg["method"]
    elif name == "skinvessel":       # configs[2]
        z = np.load(os.path.join(gold, "skinvessel_mesh.npz"))
        cfg = dict(node=z["node"], elem=z["elem"], elemprop=z["etype"], prop=np.vstack([[0, 0, 1, 1], z["prop"]]), evol=z["evol"],
                   srcpos=(0.5, 0.5, -0.005), srcdir=(0, 0, 1), srctype="disk", srcparam1=(0.3, 0, 0, 0), e0=6178,
                   tstart=0.0, tend=5e-8, tstep=5e-9, isreflect=0, seed=1648335518, method=method or "grid",
                   steps=(0.005, 0.005, 0.005), basisorder=0)
        desc = "examples/skinvessel dmmc mesh (1142 nodes/6394 tets), disk source, 10 gates, 200^3 dual grid"
    elif name == "headlike":         # configs[3] stand-in (colin27 is not shipped)
        node, elem, et = meshgen.head_like()
        prop = [[0, 0, 1, 1], [0.019, 7.8, 0.89, 1.37], [0.019, 7.8, 0.89, 1.37], [0.004, 0.009, 0.89, 1.37],
                [0.02, 9.0, 0.89, 1.37], [0.08, 40.9, 0.84, 1.37]]
        cfg = dict(node=node, elem=elem, elemprop=et, prop=prop, srcpos=(42.0, 52.0, 91.0), srcdir=(0, 0, -1),
                   tstart=0.0, tend=5e-9, tstep=5e-10, isreflect=1, seed=1648335518, method=method or "elem", basisorder=0,
                   issavedet=1, detpos=[(52.0, 52.0, 90.0, 3.0)], maxdetphoton=3000000)
        desc = "synthetic colin27-scale head (5 tissue ellipsoids on a T5 lattice), detectors + partial paths"
    elif name == "headatlas":        # configs[3] on the reference's own head mesh (colin27 itself is not shipped): mmclab/example/head_atlas.mat,
        z = np.load(os.path.join(gold, "head_atlas_mesh.npz"))     # media / source of demo_head_atlas.m:32-38 (tools/make_head_atlas.py)
        cfg = dict(node=z["node"], elem=z["elem"].astype(np.int32), elemprop=z["etype"].astype(np.int32), prop=z["prop"],
                   srcpos=tuple(float(v) for v in z["srcpos"]), srcdir=tuple(float(v) for v in z["srcdir"]), tstart=0.0, tend=5e-9, tstep=5e-10,
                   isreflect=1, seed=1648335518, method=method or "elem", basisorder=0, issavedet=1, issaveexit=1,
                   detpos=[tuple(float(v) for v in q) for q in z["detpos"]], maxdetphoton=3000000)
        desc = "mmclab/example/head_atlas.mat (59225 nodes/335713 tets, 5 tissues), pencil at C4h, detectors r=2 mm at 25 and 35 mm + partial paths, 10 gates, RayTracer=%s" % cfg["method"]
    else:
        raise SystemExit("unknown workload " + name)
    return cfg, desc


# --------------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def l2_peaks(ne, nvol, device=0):
    """The two ceilings SURVEY.md section 8(d) names, measured on THIS GPU right before the timed region with tools/microbench
    (MEASURED_PEAKS.json holds neither): dependent random 96-byte record gathers (three 256-bit loads, the photon kernel's access)
    from a table of the mesh's size, and random fire-and-forget f64 reductions into a volume of the accumulator's size.  The run
    takes about half a second.  Falls back to the round-1 figures (profiles/r1_microbench_l2gather_atomics.jsonl) when the binary
    is missing."""
    mb = os.path.join(ROOT, "tools", "microbench")
    out = {"gather_gsteps": 150.0, "red_gatomics": 196.0, "source": "fallback: profiles/r1_microbench_l2gather_atomics.jsonl (B200, round 1)"}
    if os.path.exists(mb):
        try:
            r = subprocess.run([mb, "quick", str(int(ne)), str(int(max(nvol, 1024))), str(device)], capture_output=True, text=True, timeout=120)
            for line in r.stdout.splitlines():
                j = json.loads(line)
                if j.get("bench") == "gather":
                    out["gather_gsteps"] = j["Gsteps_s"]
                elif j.get("bench") == "red":
                    out["red_gatomics"] = j["Gatomics_s"]
            out["source"] = "tools/microbench quick %d %d, run inside bench.py before the timed region" % (ne, nvol)
        except Exception as e:      # noqa: BLE001
            out["source"] += " (live run failed: %s)" % str(e)[:80]
    return out


def ref_cuda_run(cfg, nphoton):
    """The reference's own CUDA kernel (oracle/_ref/mmc_refcuda: the unmodified src/mmc_cu_host.cu + mmc_core.cl compiled for sm_100 by
    oracle/Makefile.ref) on the same workload and photon count, on this GPU.  Time = its own `kernel complete: N ms` line (host clock
    around launch + synchronise, whole milliseconds, src/mmc_cu_host.cu:640,748-753)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from mmc_b200 import api
    if not orc.ref_available(cuda=True):
        return {"unavailable": "oracle/_ref/mmc_refcuda was not built (needs /root/reference at build time)"}
    st = cfg.get("srctype", 0)
    kw = dict(nphoton=int(nphoton), seed=cfg["seed"], srcpos=cfg["srcpos"], srcdir=cfg["srcdir"],
              srctype=api.SRCTYPES.index(st) if isinstance(st, str) else st,
              srcparam1=cfg.get("srcparam1", (0, 0, 0, 0)), srcparam2=cfg.get("srcparam2", (0, 0, 0, 0)),
              tstart=cfg["tstart"], tend=cfg["tend"], tstep=cfg["tstep"], e0=cfg.get("e0", 0), isreflect=cfg["isreflect"],
              method=api.METHODS[cfg["method"]], basisorder=0, steps=cfg.get("steps", (1.0,))[0], evol=cfg.get("evol"),
              issavedet=cfg.get("issavedet", 0), detpos=cfg.get("detpos"), maxdetphoton=cfg.get("maxdetphoton", 1000000))
    try:
        runs = []
        for _ in range(2):          # the first run of the process pays context creation outside the kernel line; keep the faster kernel time
            r = orc.run_ref(np.asarray(cfg["node"], np.float32), np.asarray(cfg["elem"], np.int32), np.asarray(cfg["elemprop"], np.int32),
                            np.asarray(cfg["prop"], np.float32)[1:], cuda=True, timeout=900, **kw)
            if r.get("kernel_ms"):
                runs.append(r)
        if not runs:
            return {"unavailable": "no `kernel complete` line in the reference's output"}
        best = min(runs, key=lambda q: q["kernel_ms"])
        return {"value": nphoton / best["kernel_ms"], "unit": "photons/ms", "kernel_ms": best["kernel_ms"], "photons": int(nphoton),
                "absorbed_fraction": best.get("absorbed_frac"), "runs_ms": [q["kernel_ms"] for q in runs],
                "what": "unmodified reference CUDA kernel (src/mmc_core.cl via src/mmc_cu_host.cu, -gencode arch=compute_100,code=sm_100), same mesh / optics / "
                        "photon count on this GPU, its own `kernel complete` time"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": str(e)[-300:]}


# --------------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, nphoton, threads):
    """One bounded sample of the workload through the reference's own CPU implementation; returns photons/ms."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from mmc_b200 import api
    node, elem = np.asarray(cfg["node"], np.float32), np.asarray(cfg["elem"], np.int32)
    et, med = np.asarray(cfg["elemprop"], np.int32), np.asarray(cfg["prop"], np.float32)[1:]
    st = cfg.get("srctype", 0)
    kw = dict(nphoton=int(nphoton), seed=cfg["seed"], srcpos=cfg["srcpos"], srcdir=cfg["srcdir"],
              srctype=api.SRCTYPES.index(st) if isinstance(st, str) else st,
              srcparam1=cfg.get("srcparam1", (0, 0, 0, 0)), srcparam2=cfg.get("srcparam2", (0, 0, 0, 0)),
              tstart=cfg["tstart"], tend=cfg["tend"], tstep=cfg["tstep"], e0=cfg.get("e0", 0), isreflect=cfg["isreflect"],
              method=api.METHODS[cfg["method"]], basisorder=cfg.get("basisorder", 0), steps=cfg.get("steps", (1.0,))[0],
              issavedet=cfg.get("issavedet", 0), detpos=cfg.get("detpos"), evol=cfg.get("evol"))
    if orc.ref_available():
        t0 = time.time()
        r = orc.run_ref(node, elem, et, med, nthread=threads, **kw)
        wall = (time.time() - t0) * 1e3
        return dict(value=r.get("speed", nphoton / wall), kind="reference", wall_ms=wall, raytet=r.get("raytet"))
    t0 = time.time()
    o = orc.run(node, elem, et, med, nthread=threads, **kw)
    wall = (time.time() - t0) * 1e3
    return dict(value=nphoton / wall, kind="port", wall_ms=wall, raytet=o["raytet"])


def reference_arm(args, cfg, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = int(args.ref_photons)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(cfg, sample, threads)
        if i >= args.warmup:
            vals.append(r)
    wall = sum(v["wall_ms"] for v in vals)
    value = float(np.mean([v["value"] for v in vals]))
    line = {"impl": "reference", "metric": "photons/ms", "value": value, "unit": "photons/ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(1, len(vals)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "photons_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "photons/ms", "cores": threads, "kind": vals[0]["kind"],
                             "sample": "%d photons per step of the same workload, reference CPU (SSE4 BLB tracer, OpenMP, all host threads)" % sample},
            "e2e": {"value": value, "unit": "photons/ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="sphshells")
    ap.add_argument("--method", default=None)
    ap.add_argument("--basisorder", type=int, default=None)        # 1: nodal output (Havel / Plucker deposit into nodes inside the kernel)
    ap.add_argument("--photons", type=float, default=1e7)
    ap.add_argument("--ref-photons", type=float, default=1e6)      # ~6 s of CPU work per sample at 16 host threads (the reference CPU path runs this workload at 0.15-0.27 k photons/ms)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])   # strong: --photons is the WHOLE job, split over the ranks (BASELINE C4)
    args = ap.parse_args()
    cfg, desc = workload(args.workload, args.method)
    if args.basisorder is not None:
        cfg["basisorder"] = args.basisorder
        desc += ", basisorder=%d" % args.basisorder
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    total_photons = int(args.photons) * (world_env if args.scaling == "weak" else 1)
    if args.scaling == "strong":            # reference rule for the split (src/mmc_cu_host.cu:425-429), equal workloads
        from mmc_b200 import multigpu
        nphoton = int(multigpu.split_photons(int(args.photons), [1.0] * world_env)[0][int(os.environ.get("RANK", "0"))])
    else:
        nphoton = int(args.photons)
    cfg["nphoton"] = nphoton
    if os.environ.get("MMCB_HOTCACHE"):
        cfg["hotcache"] = int(os.environ["MMCB_HOTCACHE"])
    if os.environ.get("MMCB_BLOCK"):               # tuning runs (tools/tune.py)
        cfg["nblocksize"] = int(os.environ["MMCB_BLOCK"])

    if args.impl == "reference":
        reference_arm(args, cfg, desc)
        return

    import torch
    import mmc_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg["gpuid"] = local + 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    sess = mmc_b200.Session(cfg)
    dp = sess.devptrs()
    # the accumulator volume lives in a torch tensor so that NCCL can reduce it in place
    field = torch.zeros(dp.fieldlen, dtype=torch.float64 if dp.field_is_double else torch.float32, device=dev)
    sess.set_field_buffer(field.data_ptr())
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    class _DevView:                         # zero-copy torch view of session memory (detected-photon rows, their counter)
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}

    isdet = bool(cfg.get("issavedet")) and bool(dp.detected)
    reclen = int(dp.reclen)
    if isdet:
        det_rows = torch.as_tensor(_DevView(dp.detected, (int(cfg.get("maxdetphoton", 1000000)), reclen), "<f4"), device=dev)
        det_count = torch.as_tensor(_DevView(dp.detcount, (1,), "<i4"), device=dev)
    gathered = {"rows": 0, "prev": 0, "bytes": 0}

    def gather_detected():
        """north_star's "gather of detector records over NVLink": the rows this step appended on every rank go to rank 0 through NCCL --
        counts first (all_gather), then the payload padded to the largest count (gather)."""
        n_now = min(int(det_count.item()), det_rows.shape[0])
        mine = det_rows[gathered["prev"]:n_now]
        gathered["prev"] = n_now
        n = torch.tensor([mine.shape[0]], device=dev, dtype=torch.int64)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c) for c in counts]
        nmax = max(max(counts), 1)
        pad = torch.zeros((nmax, reclen), device=dev, dtype=torch.float32)
        pad[:mine.shape[0]] = mine
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if rank == 0:
            gathered["rows"] += sum(counts)
            gathered["bytes"] += sum(counts[1:]) * reclen * 4
    # atomic ceiling: random f64 reductions into an L2-RESIDENT volume (at most 54 MB of it): the deposits of a run concentrate around
    # the beam, so even the 862 MB volume of config C3 is hit where L2 holds it (ncu: DRAM throughput 1 % of peak); uniformly random
    # reductions over a volume that spills to HBM (23 G/s) would understate the ceiling
    l2pk = l2_peaks(len(cfg["elem"]), min(dp.fieldlen, 6750000), local) if rank == 0 else None

    # everything of a step (L2 flush, photon kernel, NCCL reduce, timing events) is enqueued on ONE explicit non-default
    # stream: torch.cuda.Event only sees the stream it is recorded on, and a NULL stream handle would make the C-ABI fall
    # back to the session's private stream
    stream = torch.cuda.Stream(device=dev)

    def step(i, timed):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xFF)                  # evict the mesh tables and the volume from L2 between steps
            stream.synchronize()
            if dist is not None and timed:
                dist.barrier()
                stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            # every rank draws consecutive slices of ITS OWN host stream (srand(seed + 7919 rank)): interleaving the ranks' slices in one
            # stream makes each rank generate and discard the other ranks' words (17 ms per step at 8 ranks)
            sess.launch(nphoton, photon_offset=0, seed=cfg["seed"] + 7919 * rank, seed_offset=i, stream=stream.cuda_stream)
            if dist is not None:
                # every step reduces what THIS step deposited: the ranks' volumes are summed into rank 0's and zeroed on the others, so
                # nothing is added twice and rank 0 ends with the sum over all ranks and steps
                dist.reduce(field, dst=0, op=dist.ReduceOp.SUM)
                if rank != 0:
                    field.zero_()
                if isdet:
                    gather_detected()
            e1.record(stream)
            stream.synchronize()
        return e0.elapsed_time(e1), sess.sync()

    for i in range(args.warmup):
        step(i, False)
    sess.reset()
    gathered.update(rows=0, prev=0, bytes=0)
    with torch.cuda.stream(stream):
        field.zero_()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kern_ms, step_ms = [], []
    for i in range(args.steps):
        sm, km = step(args.warmup + i, True)
        step_ms.append(sm)
        kern_ms.append(km)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, total_kern_ms = float(t[0]), float(t[1])
    res = sess.fetch()
    raytet = res["raytet"]
    absorbed = float(res["energyabs"][0] / max(res["energytot"][0], 1e-30))
    sess.close()

    # ---- end to end through the public one-call API (what a pmmc/mmclab user calls): host arrays in, host arrays out.  Inside the
    # timed region: mesh preparation, H2D of the tables and seeds, pilot + photon kernels, D2H of the volume, normalisation and,
    # for N>1, the NCCL reduce of the ranks' volumes to rank 0.
    e2e = None
    if not args.no_e2e:
        from mmc_b200 import multigpu
        reps, e2e_ms, e2e_kern = max(1, min(3, args.steps)), [], []
        for i in range(-1, reps):       # i = -1: one untimed warm-up call of this path (first use of the preparation / scout kernels and of
            if dist is not None:        # the stream-ordered pool after the session above was closed: 5-40 ms that belong to the process)
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = mmc_b200.run(dict(cfg, gpuid=local + 1, seed=cfg["seed"] + 7919 * (rank + world * (i + 1))))
            if dist is not None:
                multigpu.reduce_results(dict(field=r["raw"], energytot=r["energytot"], energyesc=r["energyesc"], raytet=r["raytet"]), dist, device=dev)
                torch.cuda.synchronize()
            if i >= 0:
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
                e2e_kern.append(float(r["kernel_ms"]))
        tt = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ne, nn = len(cfg["elem"]), len(cfg["node"])
        nthread = 148 * 8 * 128
        # elem (twice: face-neighbour pass + session table), numbered face neighbours, labels, nodes, seed words; the 96-byte records and
        # the centroids are built on the device.  D2H: the volume + the raw face-neighbour table (numbered on the host)
        h2d = ne * 16 * 2 + ne * 16 + ne * 4 + nn * 12 + 16 * nthread
        e2e = {"value": total_photons * reps / float(tt[0]), "unit": "photons/ms", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(r["raw"].size * 8 + ne * 16), "ms": float(tt[0]) / reps, "kernel_ms": float(np.mean(e2e_kern)), "runs": reps,
               "ms_runs": [round(x, 2) for x in e2e_ms], "warmup_calls": 1}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = total_photons * args.steps / total_ms
    hbm, peak_src = peaks()
    bps = BYTES_PER_STEP.get(cfg["method"], 92)
    steps_per_launch = raytet / args.steps
    kernel_ms = total_kern_ms / args.steps
    achieved = steps_per_launch * bps / (kernel_ms * 1e-3) / 1e9
    # from the last `ncu --set full` capture of this workload (profiles/ncu_traffic.json): DRAM bytes, global reductions and issue-slot
    # utilisation per launch
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        ncu = json.load(open(tpath)).get("%s:%s" % (args.workload, cfg["method"])) or {}
    traffic = ncu.get("dram_bytes_per_launch")
    gsteps = steps_per_launch / (kernel_ms * 1e-3) / 1e9
    # the mesh tables (and here the volume) are L2-resident, so the memory-system ceiling is the L2 GATHER rate
    # (one 96-byte record per ray-tet step), next to it the L2 ATOMIC rate for the deposits; HBM is the ceiling only for volumes that
    # spill.  All three are reported; `bound` names the one with the largest fraction, in the algorithmic bytes of section 8(d)
    # (92 B per step) so that achieved / peak is the ratio of step rates.
    fr_gather = gsteps / l2pk["gather_gsteps"]
    reds = ncu.get("global_reds_per_photon")
    gred = (reds * nphoton / (kernel_ms * 1e-3) / 1e9) if reds else None
    fr_atomic = (gred / l2pk["red_gatomics"]) if gred else None
    fr_hbm = achieved / hbm
    # HBM can only bound the kernel when tables + volume do not fit the 126 MB L2 (C3's 640 MB volume); otherwise the algorithmic bytes
    # never reach DRAM (`traffic`) and the HBM-equivalent figure is kept for reference only
    working_set = len(cfg["elem"]) * 96 + dp.fieldlen * (8 if dp.field_is_double else 4)
    spills = working_set > 120e6
    if spills and traffic:
        fr_hbm = traffic / (kernel_ms * 1e-3) / 1e9 / hbm
    cands = [("l2_gather", fr_gather)] + ([("atomic", fr_atomic)] if fr_atomic else []) + ([("hbm", fr_hbm)] if spills else [])
    bound = max(cands, key=lambda c: c[1])[0]
    roof = {"bound": bound, "unit": "GB/s", "traffic": traffic, "kernel_ms": kernel_ms, "gsteps_per_s": gsteps,
            "l2_gather": {"achieved_gsteps_per_s": gsteps, "peak_gsteps_per_s": l2pk["gather_gsteps"], "frac": fr_gather},
            "atomic": {"achieved_gred_per_s": gred, "peak_gred_per_s": l2pk["red_gatomics"], "frac": fr_atomic,
                       "reds_per_photon": reds, "reds_source": ncu.get("report")},
            "hbm_equivalent": {"achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "peak_source": peak_src,
                               "working_set_bytes": int(working_set), "spills_l2": bool(spills)},
            "issue": {"issue_slots_busy": ncu.get("issue_slots_busy"), "ipc": ncu.get("ipc"), "active_threads_per_warp_inst": ncu.get("active_threads"),
                      "source": ncu.get("report")},
            "peak_source": l2pk["source"],
            "note": "algorithmic bytes = %d B per ray-tet step (84 B record gather + 8 B atomic payload) x %.3g steps per launch; achieved and peak of the "
                    "named bound are both in those bytes.  DRAM traffic per launch (`traffic`) is a few MB: nothing here is HBM-bound.  The unit that "
                    "actually limits the kernel is instruction issue (`issue`, from ncu)." % (bps, steps_per_launch)}
    if bound == "atomic":
        roof.update(achieved=gred * 8, peak=l2pk["red_gatomics"] * 8, frac=fr_atomic)
    elif bound == "hbm":
        roof.update(achieved=fr_hbm * hbm, peak=hbm, frac=fr_hbm)
    else:
        roof.update(achieved=achieved, peak=l2pk["gather_gsteps"] * bps, frac=fr_gather)
    line = {"metric": "photons/ms", "value": value, "unit": "photons/ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "photons_per_step_per_gpu": nphoton, "photons_per_step": total_photons, "l2": "flushed between timed steps (192 MiB fill)",
                       "accumulator": "f64 red.global.add" if dp.field_is_double else "f32 red.global.add",
                       "raytet_steps_per_photon": steps_per_launch / nphoton, "absorbed_fraction": absorbed,
                       "e2e_warmup_calls": 1, "reference_arm_photons_per_step": int(args.ref_photons),
                       "reference_arm_note": "the CPU arm runs %d photons per step of the same mesh / optics (a rate metric; %d photons would take "
                                             "about a minute per step on the host cores)" % (int(args.ref_photons), nphoton)},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": roof}
    if isdet:
        line["detected"] = {"rows_gathered_on_rank0": gathered["rows"], "per_step": gathered["rows"] / max(1, args.steps),
                            "nccl_payload_bytes_per_step": gathered["bytes"] / max(1, args.steps), "reclen": reclen,
                            "how": "single GPU: rows stay in the session" if world == 1 else "counts all_gather + padded gather to rank 0 over NCCL inside every timed step"}
    if e2e is not None:
        line["e2e"] = e2e

    if not args.no_cpu_baseline and world == 1:      # the CPU leg is timed at N=1 only (rank 0); larger runs carry the GPU numbers alone
        threads = os.cpu_count() or 1
        c = cpu_reference_run(cfg, int(args.ref_photons), threads)
        line["cpu_baseline"] = {"value": c["value"], "unit": "photons/ms", "cores": threads, "kind": c["kind"],
                                "sample": "%d photons of the same workload, one run, reference CPU path with all host threads" % int(args.ref_photons)}
    if not args.no_ref_cuda and world == 1:          # the competitor north_star names: the reference's own CUDA kernel on this GPU
        rc = ref_cuda_run(cfg, nphoton)
        line["ref_cuda"] = rc
        line["vs_ref_cuda"] = (value / rc["value"]) if rc.get("value") else None
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
