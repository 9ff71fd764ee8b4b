"""TEST INFRASTRUCTURE (checker only; nothing under mmc_b200/ imports this).

numpy restatement of the reference's output normalisation, mesh_normalize (src/mmc_mesh.c:2154-2279), applied to a RAW volume
of deposits [maxgate][datalen][srcnum].  It checks the device normalisation kernels of the product (mmc_b200/csrc/mmcb_post.cu:
mmcb_norm_*) on identical raw sums; the mesh quantities it needs (evol, corrected nvol, elem, labels) come from the C oracle
(oracle/mmc_oracle.c via orc.run), which is pinned bit for bit against the reference binary."""
import numpy as np

FLUX, FLUENCE, ENERGY, JACOBIAN, WL, WP = 0, 1, 2, 3, 4, 5
GRID = 4


def mesh_normalize(raw, *, outputtype, method, basisorder, energytot, energyesc, tstep, unitinmm=1.0, nphoton=0, replay=False,
                   elem=None, etype=None, evol=None, nvol=None, mua=None):
    """raw: float64 [maxgate, datalen, srcnum]; energytot/energyesc: per pattern; mua: per medium label (already times unitinmm).
    Returns (normalised copy, mean normaliser)."""
    W = np.array(raw, dtype=np.float64, copy=True)
    maxgate, datalen, srcnum = W.shape
    facs = []
    for p in range(srcnum):
        etot = np.float32(energytot[p])
        eabs = np.float32(energytot[p] - energyesc[p])                       # float arguments, src/mmc_cu_host.cu:988
        if replay and outputtype in (JACOBIAN, WL, WP):                     # :2169-2181
            nz = np.float32(1.0) / (np.float32(1e-4) * np.float32(nphoton)) if outputtype == JACOBIAN else np.float32(1.0) / etot
            W[:, :, p] *= np.float64(nz)
            facs.append(float(nz))
            continue
        if outputtype == ENERGY:                                            # :2183-2191
            nz = np.float64(np.float32(1.0) / etot)
            W[:, :, p] *= nz
            facs.append(float(nz))
            continue
        if method == GRID:                                                  # :2205-2207
            nz = 1.0 / (np.float64(etot) * np.float64(np.float32(unitinmm)) ** 3)
        elif basisorder:                                                    # :2208-2246
            pos = nvol > 0
            W[:, pos, p] /= nvol[pos].astype(np.float64)
            wf = W[:, :, p].astype(np.float32)                              # `float re_val = ...`
            esum = wf[:, elem - 1].astype(np.float64).sum(axis=(0, 2))      # per element: gates x 4 nodes, accumulated in double
            dep = float((esum * evol.astype(np.float64) * mua[etype].astype(np.float64)).sum())
            nz = np.float64(eabs) / (np.float64(etot) * dep * np.float64(np.float32(0.25)))
        else:                                                               # :2247-2260
            dep = float(W[:, :, p].sum())
            emua = (evol * mua[etype]).astype(np.float32).astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                W[:, :, p] /= emua[None, :]
            nz = np.float64(eabs) / (np.float64(etot) * dep)
        if outputtype == FLUX:
            nz = nz / np.float64(np.float32(tstep))
        with np.errstate(invalid="ignore"):
            W[:, :, p] *= nz
        facs.append(float(nz))
    return W, float(np.mean(facs))
