"""Synthetic tetrahedral meshes for the BASELINE configs.

`gen_t5_mesh` restates the reference's lattice tessellation (matlab/genT5mesh.m:17-60:
40 tetrahedra per 2x2x2 block of lattice cells) so that the C1 benchmark mesh
(`examples/validation/createmesh.m:4`, genT5mesh(0:2:60,0:2:60,0:2:60) -> 29791 nodes,
135000 elements) can be produced without MATLAB.  Node indices in `elem` are 1-based like
every mesh array of the reference.
"""
from __future__ import annotations

import numpy as np

# matlab/genT5mesh.m:32-41 -- node numbers (1..27) inside one 3x3x3 node block
_CUBE8 = np.array([
    1, 4, 5, 13, 1, 2, 5, 11, 1, 10, 11, 13, 11, 13, 14, 5, 11, 13, 1, 5,
    2, 3, 5, 11, 3, 5, 6, 15, 15, 11, 12, 3, 15, 11, 14, 5, 11, 15, 3, 5,
    4, 5, 7, 13, 5, 7, 8, 17, 16, 17, 13, 7, 13, 17, 14, 5, 5, 7, 17, 13,
    5, 6, 9, 15, 5, 8, 9, 17, 17, 18, 15, 9, 17, 15, 14, 5, 17, 15, 5, 9,
    10, 13, 11, 19, 13, 11, 14, 23, 22, 19, 23, 13, 19, 23, 20, 11, 13, 11, 19, 23,
    11, 12, 15, 21, 11, 15, 14, 23, 23, 21, 20, 11, 23, 24, 21, 15, 23, 21, 11, 15,
    16, 13, 17, 25, 13, 17, 14, 23, 25, 26, 23, 17, 25, 22, 23, 13, 13, 17, 25, 23,
    17, 18, 15, 27, 17, 15, 14, 23, 26, 27, 23, 17, 27, 23, 24, 15, 23, 27, 17, 15,
], dtype=np.int64).reshape(40, 4)


def gen_t5_mesh(xs, ys, zs):
    """Tessellate the lattice xs x ys x zs (each with an odd number >= 3 of points).

    Returns (node float32 [nn,3], elem int32 [ne,4] 1-based).  Node (i,j,k) has the linear
    index i + nx*(j + ny*k) (ndgrid order, genT5mesh.m:63-72)."""
    vs = []
    for v in (xs, ys, zs):
        v = np.asarray(v, dtype=np.float64)
        if len(v) % 2 == 0:  # genT5mesh.m:23-28
            v = np.linspace(v[0], v[-1], len(v) + 1)
        if len(v) < 3:
            raise ValueError("each dimension needs at least 3 lattice points")
        vs.append(v)
    nx, ny, nz = (len(v) for v in vs)
    gx, gy, gz = np.meshgrid(vs[0], vs[1], vs[2], indexing="ij")
    node = np.stack([gx.ravel(order="F"), gy.ravel(order="F"), gz.ravel(order="F")], axis=1).astype(np.float32)

    base = np.array([0, 1, 2, nx, nx + 1, nx + 2, 2 * nx, 2 * nx + 1, 2 * nx + 2], dtype=np.int64)
    shift = np.concatenate([base, base + nx * ny, base + 2 * nx * ny])  # genT5mesh.m:52-54
    local = shift[_CUBE8 - 1]                                         # [40,4] offsets from the block origin
    # block origins, ordered like ind=sub2ind(...,ix(:),iy(:),iz(:)) of a meshgrid (iy fastest, then ix, then iz)
    ix = np.arange(0, nx - 2, 2)
    iy = np.arange(0, ny - 2, 2)
    iz = np.arange(0, nz - 2, 2)
    IZ, IX, IY = np.meshgrid(iz, ix, iy, indexing="ij")
    ind = (IX + nx * (IY + ny * IZ)).ravel() + 1                       # 1-based
    elem = (ind[:, None, None] + local[None, :, :]).reshape(-1, 4)
    return node, reorient(node, elem.astype(np.int32))


def tet_volume6(node, elem):
    """6 x signed volume, positive for the orientation the reference expects
    (mesh_getvolume, src/mmc_mesh.c:920-937 flips nodes 3,4 when this is negative)."""
    p = node.astype(np.float64)[elem - 1]
    a, b, c = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    return -np.einsum("ij,ij->i", np.cross(a, b), c)


def reorient(node, elem):
    """meshreorient equivalent: swap the last two nodes of negatively oriented tets."""
    elem = elem.copy()
    neg = tet_volume6(node, elem) < 0
    elem[neg, 2], elem[neg, 3] = elem[neg, 3].copy(), elem[neg, 2].copy()
    return elem


def centroids(node, elem):
    return node.astype(np.float64)[elem - 1].mean(axis=1)


def cube60():
    """BASELINE config C1: 60 mm cube, 2 mm lattice, one medium."""
    g = np.arange(0, 61, 2)
    node, elem = gen_t5_mesh(g, g, g)
    return node, elem, np.ones(len(elem), dtype=np.int32)


def layered_sphere_cube(n=60, step=2, radii=(25.0, 23.0, 10.0), center=None):
    """A cube lattice whose tets are labelled by concentric spherical shells (a lattice stand-in
    for the sphshells geometry; labels: 1 outside all spheres, 2.. going inwards)."""
    g = np.arange(0, n + 1, step)
    node, elem = gen_t5_mesh(g, g, g)
    c = centroids(node, elem)
    ctr = np.full(3, n / 2.0) if center is None else np.asarray(center, dtype=np.float64)
    r = np.linalg.norm(c - ctr, axis=1)
    etype = np.ones(len(elem), dtype=np.int32)
    for k, rad in enumerate(radii):
        etype[r < rad] = k + 2
    return node, elem, etype


def head_like(n=(84, 104, 92), step=2):
    """A colin27-scale (~420k tets) synthetic head: nested ellipsoids scalp/skull/CSF/gray/white
    on a T5 lattice.  Not anatomical -- a same-size, same-media-count workload for BASELINE C4."""
    gx, gy, gz = (np.arange(0, m + 1, step) for m in n)
    node, elem = gen_t5_mesh(gx, gy, gz)
    c = centroids(node, elem)
    ctr = np.array(n, dtype=np.float64) / 2.0
    half = np.array(n, dtype=np.float64) / 2.0
    q = np.sqrt((((c - ctr) / half) ** 2).sum(axis=1))   # 1 on the outer ellipsoid
    etype = np.zeros(len(elem), dtype=np.int32)
    for lab, lim in ((1, 1.0), (2, 0.92), (3, 0.84), (4, 0.78), (5, 0.60)):
        etype[q < lim] = lab
    keep = etype > 0
    elem, etype = elem[keep], etype[keep]
    used = np.unique(elem)
    remap = np.zeros(len(node) + 1, dtype=np.int32)
    remap[used] = np.arange(1, len(used) + 1, dtype=np.int32)
    return node[used - 1], remap[elem], etype


def slab_with_wide_src_det(nx=60, ny=60, nz=20, step=2, gap=2.0,
                           src_rect=((10.0, 10.0), (50.0, 50.0)), det_z=None):
    """BASELINE config C5 stand-in (examples/replaywide/createmesh.m:3-29): a slab [0,nx]x[0,ny]x[0,nz]
    of medium 1 with an air layer of thickness `gap` below (z<0) and above (z>nz).  Tets of the lower
    air layer are labelled -1 (wide-field source candidates), those of the upper layer -2
    (wide-field detector); src/mmc_mesh.c:390-427 describes how the labels are consumed."""
    gx = np.arange(0, nx + 1, step)
    gy = np.arange(0, ny + 1, step)
    gz = np.concatenate([[-gap, -gap / 2.0], np.arange(0, nz + 1, step), [nz + gap / 2.0, nz + gap]])
    node, elem = gen_t5_mesh(gx, gy, gz)
    c = centroids(node, elem)
    etype = np.ones(len(elem), dtype=np.int32)
    etype[c[:, 2] < 0] = -1
    etype[c[:, 2] > nz] = -2
    return node, elem, etype
