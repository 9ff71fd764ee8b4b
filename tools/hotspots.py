#!/usr/bin/env python
"""Atomic hot-spot analysis: with a library built with -DMMCB_COUNT_DEPOSITS the output volume holds the NUMBER of
red.global.add operations per accumulator; this prints how concentrated they are per 128-byte L2 line.
    python tools/hotspots.py build | run"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "build", "variants", "libmmc_b200_count.so")

if sys.argv[1] == "build":
    from mmc_b200 import build
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    build.build(force=True, extra=["-DMMCB_COUNT_DEPOSITS"], out=LIB, tag="_count")
    sys.exit(0)

os.environ["MMCB_LIB"] = LIB
import bench  # noqa: E402
import mmc_b200  # noqa: E402

for name, method in (("cube60", "elem"), ("sphshells", "grid"), ("sphshells", "elem"), ("cube60", "grid")):
    cfg, desc = bench.workload(name, method)
    cfg.update(nphoton=1000000, isnormalized=0)
    r = mmc_b200.run(cfg)
    cnt = np.asarray(r["raw"], np.float64).reshape(-1)
    total = cnt.sum()
    n = (len(cnt) + 15) // 16 * 16
    lines = np.zeros(n)
    lines[:len(cnt)] = cnt
    lines = np.sort(lines.reshape(-1, 16).sum(axis=1))[::-1]
    addr = np.sort(cnt)[::-1]
    out = dict(workload="%s:%s" % (name, method), photons=1000000, atomics=total, raytet=r["raytet"], atomics_per_step=total / r["raytet"],
               top_addr=[int(x) for x in addr[:12]], top_lines=[int(x) for x in lines[:12]],
               lines_over_1pct_of_photons=int((lines > 1e4).sum()), share_top32_addr=float(addr[:32].sum() / total),
               share_top256_addr=float(addr[:256].sum() / total), share_top4096_addr=float(addr[:4096].sum() / total),
               share_top64_lines=float(lines[:64].sum() / total), share_top1024_lines=float(lines[:1024].sum() / total))
    print(json.dumps(out), flush=True)
