"""Multi-GPU driver: photons shard across ranks (one process per GPU), every rank holds the whole mesh and its own
slice of the host seed stream; the only exchange is the epilogue -- a sum-reduce of the fluence volume and energy
tallies to rank 0 and a gather of detected-photon records -- through torch.distributed (NCCL over NVLink on GPUs,
gloo in the CPU tests).  The reference does this with one OpenMP host thread per device and host `+=` loops
(src/mmc_cu_host.cu:403-429 workload split, :755-783,832-853,916-928 merges).
"""
from __future__ import annotations

import numpy as np


def split_photons(nphoton, workload):
    """Reference rule (src/mmc_cu_host.cu:425-429): rank g simulates nphoton*w_g/sum(w) photons; the remainder goes to
    the last ranks so that the total is exact."""
    w = np.asarray(workload, dtype=np.float64)
    if (w <= 0).any():
        raise ValueError("workload was unspecified for an active device")
    share = np.floor(nphoton * w / w.sum()).astype(np.int64)
    rem = int(nphoton - share.sum())
    for i in range(rem):
        share[len(share) - 1 - (i % len(share))] += 1
    offs = np.concatenate([[0], np.cumsum(share)[:-1]])
    return share, offs


def reduce_results(local, dist, device="cpu", dst=0):
    """Sum-reduce field/energy/raytet to rank `dst`, gather detected-photon rows (counts first, then padded payload)
    and truncate at maxdetphoton like the reference (src/mmc_cu_host.cu:823-834).  `local` is a dict with numpy (or
    torch) entries: field, energytot, energyesc, raytet, field_im (optional, RF), detp [n,reclen] (optional), seeds [n,2] (optional).
    Adjoint Jacobians are products of slot fluences: reduce the raw volumes first (in place, through mmcb_set_field_buffer /
    mmcb_get_devptrs), then let rank `dst` fetch -- mmcb_fetch normalises and runs the post-kernels on the reduced volumes."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()

    def T(x, dtype=torch.float64):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
        return t.to(device=device, dtype=dtype).contiguous()

    out = {}
    field = T(local["field"])
    dist.reduce(field, dst=dst, op=dist.ReduceOp.SUM)
    scal = T(np.concatenate([np.atleast_1d(local["energytot"]), np.atleast_1d(local["energyesc"]), [local["raytet"]]]))
    dist.reduce(scal, dst=dst, op=dist.ReduceOp.SUM)
    ns = (len(scal) - 1) // 2
    field_im = None
    if local.get("field_im") is not None:       # RF runs: the imaginary volume is reduced like the real one
        field_im = T(local["field_im"])
        dist.reduce(field_im, dst=dst, op=dist.ReduceOp.SUM)
    if rank == dst:
        out["field"] = field
        if field_im is not None:
            out["field_im"] = field_im
        out["energytot"] = scal[:ns].cpu().numpy()
        out["energyesc"] = scal[ns:2 * ns].cpu().numpy()
        out["raytet"] = float(scal[-1])
    if local.get("detp") is not None:
        detp = T(local["detp"], torch.float32)
        n = torch.tensor([detp.shape[0]], device=device, dtype=torch.int64)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c) for c in counts]
        nmax = max(max(counts), 1)
        reclen = detp.shape[1] if detp.ndim == 2 else int(local.get("reclen", 1))
        pad = torch.zeros((nmax, reclen), device=device, dtype=torch.float32)
        pad[:detp.shape[0]] = detp.reshape(-1, reclen)
        bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
        seeds_b = None
        if local.get("seeds") is not None:
            sd = torch.as_tensor(np.ascontiguousarray(local["seeds"]).view(np.int64).reshape(-1, 2)).to(device)
            spad = torch.zeros((nmax, 2), device=device, dtype=torch.int64)
            spad[:sd.shape[0]] = sd
            seeds_b = [torch.zeros_like(spad) for _ in range(world)] if rank == dst else None
            dist.gather(spad, seeds_b, dst=dst)
        if rank == dst:
            rows = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
            maxdet = int(local.get("maxdetphoton", rows.shape[0]))
            out["detectedtotal"] = int(sum(counts))
            out["detp"] = rows[:maxdet].cpu().numpy()
            if seeds_b is not None:
                s = torch.cat([b[:c] for b, c in zip(seeds_b, counts)], dim=0)[:maxdet]
                out["seeds"] = s.cpu().numpy().view(np.uint64)
    return out
