"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's adjoint-Jacobian post-processing -- the four kernels of
src/mmc_core.cl:2218-2649 as driven by src/mmc_cu_host.cu:1063-1395 -- used by tests/ to check the CUDA post-kernels of
mmc_b200/csrc/mmcb_post.cu.  Only tests/ may import this; nothing under mmc_b200/ does.

Parity status: the reference implements these kernels on the GPU only (the CPU program has no adjoint path) and ships no
golden vectors for them, so this restatement is pinned by construction (formula by formula, float32 like the kernels), not by a
reference run: "parity unpinned" for the adjoint rows.

Field layout (all functions): field[slot, gate, i] float32 -- the reference's field[i + (gate + slot*maxgate)*N]."""
import numpy as np

F = np.float32


def cw_sum(field):
    """mmc_cw_sum, src/mmc_core.cl:2227-2242: sequential float32 sum over the time gates -> [slot, i]."""
    field = np.asarray(field, dtype=F)
    out = np.zeros((field.shape[0], field.shape[2]), dtype=F)
    for t in range(field.shape[1]):
        out = (out + field[:, t, :]).astype(F)
    return out


def _cplx(re, im):
    return (re, im if im is not None else np.zeros_like(re))


def jmua_grid(cw_re, cw_im, Ns, Nd, scale):
    """mmc_adjoint_kernel (:2298-2333) + host scaling by -Vvox (src/mmc_cu_host.cu:1357-1380): returns (re[Ns*Nd, N], im or None)."""
    N = cw_re.shape[1]
    re = np.zeros((Ns * Nd, N), dtype=F)
    im = np.zeros((Ns * Nd, N), dtype=F) if cw_im is not None else None
    for s in range(Ns):
        for d in range(Nd):
            sr, dr = cw_re[s], cw_re[Ns + d]
            r = sr * dr
            if cw_im is not None:
                si, di = cw_im[s], cw_im[Ns + d]
                r = r - si * di
                im[s * Nd + d] = F(scale) * (sr * di + si * dr)
            re[s * Nd + d] = F(scale) * r
    return re, im


def fd_grad(vol, axis):
    """mmc_fd_grad (:2247-2285): 2nd-order one-sided differences at the ends, central inside, unit spacing; vol[z, y, x]."""
    n = vol.shape[axis]
    g = np.zeros_like(vol)
    if n <= 1:
        return g
    v = np.moveaxis(vol, axis, 0)
    o = np.moveaxis(g, axis, 0)
    if n == 2:
        o[0] = v[1] - v[0]
        o[1] = v[1] - v[0]
        return g
    o[0] = (F(-3) * v[0] + F(4) * v[1] - v[2]) * F(0.5)
    o[-1] = (v[-3] - F(4) * v[-2] + F(3) * v[-1]) * F(0.5)
    o[1:-1] = (v[2:] - v[:-2]) * F(0.5)
    return g


def jd_grid(cw_re, cw_im, Ns, Nd, dim, scale):
    """mmc_adjoint_dcoeff_kernel (:2343-2401) + host scaling by -unitinmm; dim = (Nx, Ny, Nz), voxel index = iz*Ny*Nx + iy*Nx + ix."""
    Nx, Ny, Nz = dim
    N = cw_re.shape[1]

    def grads(a):
        vol = np.asarray(a, dtype=F).reshape(Nz, Ny, Nx)
        return [fd_grad(vol, 2).reshape(N), fd_grad(vol, 1).reshape(N), fd_grad(vol, 0).reshape(N)]      # x, y, z

    gr = [grads(cw_re[k]) for k in range(cw_re.shape[0])]
    gi = [grads(cw_im[k]) for k in range(cw_re.shape[0])] if cw_im is not None else None
    re = np.zeros((Ns * Nd, N), dtype=F)
    im = np.zeros((Ns * Nd, N), dtype=F) if cw_im is not None else None
    for s in range(Ns):
        for d in range(Nd):
            a, b = gr[s], gr[Ns + d]
            r = a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
            if gi is not None:
                ai, bi = gi[s], gi[Ns + d]
                r = r - (ai[0] * bi[0] + ai[1] * bi[1] + ai[2] * bi[2])
                im[s * Nd + d] = F(scale) * (a[0] * bi[0] + a[1] * bi[1] + a[2] * bi[2] + ai[0] * b[0] + ai[1] * b[1] + ai[2] * b[2])
            re[s * Nd + d] = F(scale) * r
    return re, im


def deldotdel(node, elem, evol):
    """mesh_deldotdel, src/mmc_mesh.c:990-1049: <grad N_i . grad N_j> * Ve, packed upper triangle [00,01,02,03,11,12,13,22,23,33]
    (double precision like the reference; elem 1-based [ne, 4])."""
    N = np.asarray(node, dtype=np.float64)
    e = np.asarray(elem, dtype=np.int64) - 1
    p1, p2, p3, p4 = N[e[:, 0]], N[e[:, 1]], N[e[:, 2]], N[e[:, 3]]
    Ve = np.asarray(evol, dtype=np.float64)
    R = 1.0 / (Ve * 6.0)
    x, y, z = 0, 1, 2

    def der(a, b, c, sx, sy, sz):
        # the three components follow the reference's expressions with (a, b, c) = the three other nodes
        dx = sx * ((b[:, y] * c[:, z] - b[:, z] * c[:, y]) - a[:, y] * (c[:, z] - b[:, z]) + a[:, z] * (c[:, y] - b[:, y])) * R
        dy = sy * ((b[:, x] * c[:, z] - c[:, x] * b[:, z]) - a[:, x] * (c[:, z] - b[:, z]) + a[:, z] * (c[:, x] - b[:, x])) * R
        dz = sz * ((b[:, x] * c[:, y] - b[:, y] * c[:, x]) - a[:, x] * (c[:, y] - b[:, y]) + a[:, y] * (c[:, x] - b[:, x])) * R
        return np.stack([dx, dy, dz], axis=1)

    g = [der(p2, p3, p4, -1, 1, -1), der(p1, p3, p4, 1, -1, 1), der(p1, p2, p4, -1, 1, -1), der(p1, p2, p3, 1, -1, 1)]
    out = np.zeros((len(e), 10))
    k = 0
    for i in range(4):
        for j in range(i, 4):
            out[:, k] = (g[i] * g[j]).sum(axis=1) * Ve
            k += 1
    return out


def jac_mesh_full(cw_re, cw_im, elem, evol, ddd, Ns, Nd, want_mua=True, want_d=True):
    """mmc_adjoint_mesh_full_kernel (:2438-2587) with nodal output (0.25-weighted scatter to the 4 nodes): returns dict of
    (re[Ns*Nd, nn], im or None) for 'jmua' and 'jd'."""
    e = np.asarray(elem, dtype=np.int64) - 1
    nn = cw_re.shape[1]
    Ve = np.asarray(evol, dtype=F)
    ddd = np.asarray(ddd, dtype=F)
    diag, off, pa, pb = [0, 4, 7, 9], [1, 2, 3, 5, 6, 8], [0, 0, 0, 1, 1, 2], [1, 2, 3, 2, 3, 3]
    rf = cw_im is not None
    res = {k: (np.zeros((Ns * Nd, nn), dtype=np.float64), np.zeros((Ns * Nd, nn), dtype=np.float64) if rf else None) for k in ("jmua", "jd")}
    for s in range(Ns):
        psr = cw_re[s][e]
        psi = cw_im[s][e] if rf else np.zeros_like(psr)
        for d in range(Nd):
            pdr = cw_re[Ns + d][e]
            pdi = cw_im[Ns + d][e] if rf else np.zeros_like(pdr)
            mr = np.zeros(len(e), dtype=F)
            mi = np.zeros(len(e), dtype=F)
            dr = np.zeros(len(e), dtype=F)
            di = np.zeros(len(e), dtype=F)
            for i in range(4):
                pre = psr[:, i] * pdr[:, i] - psi[:, i] * pdi[:, i]
                pim = psr[:, i] * pdi[:, i] + psi[:, i] * pdr[:, i]
                mr += pre
                mi += pim
                dr += ddd[:, diag[i]] * pre
                di += ddd[:, diag[i]] * pim
            for p in range(6):
                a, b = pa[p], pb[p]
                pre = psr[:, a] * pdr[:, b] + psr[:, b] * pdr[:, a] - psi[:, a] * pdi[:, b] - psi[:, b] * pdi[:, a]
                pim = psr[:, a] * pdi[:, b] + psr[:, b] * pdi[:, a] + psi[:, a] * pdr[:, b] + psi[:, b] * pdr[:, a]
                mr += F(0.5) * pre
                mi += F(0.5) * pim
                dr += ddd[:, off[p]] * pre
                di += ddd[:, off[p]] * pim
            mr = mr * F(-0.1) * Ve * F(0.25)
            mi = mi * F(-0.1) * Ve * F(0.25)
            dr = dr * F(-0.25)
            di = di * F(-0.25)
            sd = s * Nd + d
            for k in range(4):
                if want_mua:
                    np.add.at(res["jmua"][0][sd], e[:, k], mr)
                    if rf:
                        np.add.at(res["jmua"][1][sd], e[:, k], mi)
                if want_d:
                    np.add.at(res["jd"][0][sd], e[:, k], dr)
                    if rf:
                        np.add.at(res["jd"][1][sd], e[:, k], di)
    return res


def jmua_mesh_nodal(cw_re, cw_im, nvol, Ns, Nd):
    """mmc_adjoint_mesh_nodal_kernel (:2609-2649): J_mua[n] = -nvol[n] phi_s[n] phi_d[n]."""
    vol = np.asarray(nvol, dtype=F)
    nn = cw_re.shape[1]
    re = np.zeros((Ns * Nd, nn), dtype=F)
    im = np.zeros((Ns * Nd, nn), dtype=F) if cw_im is not None else None
    for s in range(Ns):
        for d in range(Nd):
            sr, dr = cw_re[s], cw_re[Ns + d]
            if cw_im is not None:
                si, di = cw_im[s], cw_im[Ns + d]
                re[s * Nd + d] = -vol * (sr * dr - si * di)
                im[s * Nd + d] = -vol * (sr * di + si * dr)
            else:
                re[s * Nd + d] = -vol * sr * dr
    return re, im
