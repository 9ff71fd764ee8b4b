#!/usr/bin/env python
"""Convert the reference's own head mesh, mmclab/example/head_atlas.mat (59 225 nodes, 335 713 tetrahedra, 7 labels: the mesh of
mmclab/example/demo_head_atlas.m and BASELINE config C4's stand-in for colin27, which the repository does not ship), into
tests/golden/head_atlas_mesh.npz so that it travels to the GPU box.  Nothing is computed here except two detector positions: the scalp
nodes closest to 25 mm and 35 mm from the demo's source (demo_head_atlas.m:36-38), for the detected-photon / partial-path runs of C4.
Run where /root/reference exists:   python tools/make_head_atlas.py"""
import os
import sys

import numpy as np
import scipy.io as sio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/mmclab/example/head_atlas.mat"


def main():
    m = sio.loadmat(SRC)
    node = np.ascontiguousarray(m["node"], dtype=np.float32)
    elem = np.ascontiguousarray(m["elem"], dtype=np.uint16)            # 59 225 nodes: the ids fit 16 bits
    etype = np.ascontiguousarray(m["prop"].ravel(), dtype=np.uint8)
    assert elem.max() <= len(node) and elem.min() >= 1
    # media of demo_head_atlas.m:32 (row 0 = ambient); labels 1..6: air cavities, scalp, skull, CSF, gray matter, white matter
    prop = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0.019, 7.8, 0.89, 1.37], [0.019, 7.8, 0.89, 1.37], [0.0004, 0.009, 0.89, 1.37],
                     [0.02, 9.0, 0.89, 1.37], [0.08, 40.9, 0.84, 1.37]], dtype=np.float32)
    srcdir = np.array([-0.5086, -0.1822, -0.8415], dtype=np.float64)
    srcdir /= np.linalg.norm(srcdir)
    srcpos = np.array([133.5370, 90.1988, 200.0700]) + 0.001 * srcdir  # :36-38
    # scalp surface = faces that belong to exactly one non-air element
    e = elem.astype(np.int64)
    solid = etype >= 2
    faces = np.concatenate([e[solid][:, [0, 1, 2]], e[solid][:, [0, 1, 3]], e[solid][:, [0, 2, 3]], e[solid][:, [1, 2, 3]]])
    faces.sort(axis=1)
    key = (faces[:, 0] << 42) | (faces[:, 1] << 21) | faces[:, 2]
    uniq, cnt = np.unique(key, return_counts=True)
    surf = uniq[cnt == 1]
    ids = np.unique(np.concatenate([surf >> 42, (surf >> 21) & ((1 << 21) - 1), surf & ((1 << 21) - 1)]))
    pts = node[ids - 1].astype(np.float64)
    d = np.linalg.norm(pts - srcpos, axis=1)
    outer = pts[:, 2] > srcpos[2] - 40                                  # stay on the outer scalp near the source, not in a cavity
    cand = np.where(outer)[0]
    # two detectors, 25 mm and 35 mm from the source (two, because the reference's CUDA host rejects this mesh with ONE detector: its
    # constant-memory budget check `>= MAX_PROP` trips when 4000 - (media + detectors) is a multiple of 4, src/mmc_cu_host.cu:294,497)
    dets = [pts[cand[np.argmin(np.abs(d[cand] - r))]] for r in (25.0, 35.0)]
    print("labels", np.bincount(etype), "surface nodes", len(ids), "detectors", dets, "distances", [np.linalg.norm(q - srcpos) for q in dets])
    out = os.path.join(ROOT, "tests", "golden", "head_atlas_mesh.npz")
    np.savez_compressed(out, node=node, elem=elem, etype=etype, prop=prop, srcpos=srcpos.astype(np.float32), srcdir=srcdir.astype(np.float32),
                        detpos=np.array([[q[0], q[1], q[2], 2.0] for q in dets], dtype=np.float32))
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
