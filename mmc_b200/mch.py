"""On-disk result formats of the reference, written/read from this engine's buffers:

  <session>.mch   64-byte `history` header (src/mmc_utils.h:168-184) + float32 rows
                  [detid, nscat[M], ppath[M], (mom[M]), (p[3], v[3]), w0] + optional 16-byte xorshift128+ seeds
                  (writers: mesh_savedetphoton src/mmc_mesh.c:1978-2010, mcx_savedetphoton src/mmc_utils.c:4309-4343)
  <session>.bin   raw float64 volume, gate-major (src/mmc_mesh.c:1745-1756)

and the replay preparation of mesh_loadseedfile (src/mmc_mesh.c:815-898): detected-photon weights and arrival times
for `-E file.mch -O L|P|J`.
"""
from __future__ import annotations

import struct

import numpy as np

R_C0 = np.float32(3.335640951981520e-12)        # src/mmc_const.h:45, s/mm
# magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton, unitinmm, seedbyte, normalizer,
# srcnum, respin, savedetflag, reserved[2]   (note: srcnum BEFORE respin -- matlab/loadmch.m:70-72 reads them swapped)
_HDR = struct.Struct("<4s7IfIfIiI2i")
assert _HDR.size == 64


def savemch(path, detp, seeds=None, *, maxmedia, totalphoton, detnum=0, detected=None, unitinmm=1.0, normalizer=1.0,
            srcnum=1, respin=1, savedetflag=0):
    """Write detected-photon rows (float32 [n, colcount]) and optional seeds (uint64 [n, 2]) as the reference does."""
    detp = np.ascontiguousarray(detp, dtype=np.float32)
    if detp.ndim != 2:
        raise ValueError("detp must be [n, colcount]")
    n, col = detp.shape
    seedbyte = 0
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds).view(np.uint64).reshape(-1, 2)
        if len(seeds) != n:
            raise ValueError("one seed per detected photon is required")
        seedbyte = 16                               # sizeof(RandType)*RAND_BUF_LEN for xorshift128+
    hdr = _HDR.pack(b"MCXH", 1, int(maxmedia), int(detnum), col, int(totalphoton) & 0xFFFFFFFF,
                    int(n if detected is None else detected), n, float(unitinmm), seedbyte, float(normalizer),
                    int(srcnum), int(respin), int(savedetflag), 0, 0)
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(detp.tobytes())
        if seedbyte:
            f.write(seeds.tobytes())


def loadmch(path):
    """Read an .mch file: returns dict(header fields..., detp float32 [n, colcount], seeds uint64 [n, 2] or None)."""
    with open(path, "rb") as f:
        raw = f.read()
    (magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton, unitinmm, seedbyte, normalizer,
     srcnum, respin, savedetflag, _, _) = _HDR.unpack_from(raw, 0)
    if magic != b"MCXH":
        raise ValueError("not an MCX history file")
    off = 64
    nb = savedphoton * colcount * 4
    detp = np.frombuffer(raw, dtype=np.float32, count=savedphoton * colcount, offset=off).reshape(savedphoton, colcount).copy()
    off += nb
    seeds = None
    if seedbyte:
        if seedbyte != 16:
            raise ValueError("only 16-byte xorshift128+ seeds are supported")
        seeds = np.frombuffer(raw, dtype=np.uint64, count=savedphoton * 2, offset=off).reshape(savedphoton, 2).copy()
    return dict(version=version, maxmedia=maxmedia, detnum=detnum, colcount=colcount, totalphoton=totalphoton,
                detected=detected, savedphoton=savedphoton, unitinmm=unitinmm, seedbyte=seedbyte, normalizer=normalizer,
                srcnum=srcnum, respin=respin, savedetflag=savedetflag, detp=detp, seeds=seeds)


def replay_inputs(mch, prop, replaydet=0, legacy_columns=False):
    """mesh_loadseedfile (src/mmc_mesh.c:855-891): select the photons of detector `replaydet` (0 = all) and compute
    replayweight = w0 * prod_j exp(-mua_j * ppath_j * unitinmm) and replaytime = sum_j n_j * ppath_j * R_C0.

    prop: [[mua, mus, g, n]] with row 0 = background (like cfg.prop).  The reference loops j = 2 .. maxmedia+1 over the
    row, i.e. it assumes ONE scattering-count column before the partial paths (the legacy MCX layout); rows written by
    MMC carry maxmedia scattering-count columns, so for maxmedia > 1 the reference reads the wrong columns.  This
    helper reads the partial-path columns where MMC writes them (1+M .. 2M), which coincides with the reference for
    maxmedia == 1.  legacy_columns=True reads columns 2 .. maxmedia+1 exactly like the reference does (used by the parity test that
    replays a two-media run in both programs: the weights are then equally "wrong" on both sides and the trajectories must coincide)."""
    if mch["seeds"] is None:
        raise ValueError("the history file carries no seeds (run with issaveseed=1)")
    prop = np.asarray(prop, dtype=np.float32).reshape(-1, 4)
    M = int(mch["maxmedia"])
    d = mch["detp"]
    sel = np.ones(len(d), bool) if replaydet == 0 else (d[:, 0].astype(np.int64) == int(replaydet))
    pp = d[sel, 2:2 + M] if legacy_columns else d[sel, 1 + M:1 + 2 * M]
    w = d[sel, -1].astype(np.float32).copy()
    t = np.zeros(len(w), dtype=np.float32)
    for j in range(M):
        w *= np.exp(-prop[j + 1, 0] * pp[:, j] * np.float32(mch["unitinmm"]), dtype=np.float32)
        t += prop[j + 1, 3] * pp[:, j] * R_C0
    return dict(replayseed=mch["seeds"][sel].copy(), replayweight=w, replaytime=t, nphoton=int(sel.sum()))


def savebin(path, field):
    """-F bin: raw float64, gate-major (src/mmc_mesh.c:1745-1756)."""
    np.ascontiguousarray(field, dtype=np.float64).tofile(path)


def loadbin(path, maxgate):
    a = np.fromfile(path, dtype=np.float64)
    return a.reshape(int(maxgate), -1)
