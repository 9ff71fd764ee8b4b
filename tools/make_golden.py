#!/usr/bin/env python
"""Generate tests/golden/ref_cases.npz by running the UNMODIFIED reference binary
(oracle/_ref/mmc_ref, built by `make -C oracle -f Makefile.ref`) single-threaded on the shared
small cases of tests/cases.py.  /root/reference is only needed to (re)build that binary; the
fixtures travel to the GPU box where it is absent.

For each case we keep: ray-tet count, absorbed fraction, normaliser, detected count, the sum of the
output volume, its per-gate sums and 64 sampled entries (fixed indices) -- enough to pin the oracle
bit-for-bit without committing megabytes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import orc  # noqa: E402


def sample_idx(n, k=64):
    rng = np.random.RandomState(12345)
    return np.sort(rng.choice(n, size=min(k, n), replace=False))


def main():
    out = {}
    meta = {}
    for name in cases.CASES:
        node, elem, et, med = cases.case_mesh(name)
        kw = cases.case_kwargs(name)
        r = orc.run_ref(node, elem, et, med, nthread=1, **kw)
        f = r["field_flat"]
        nonfinite = int((~np.isfinite(f)).sum())       # void (mua=0) elements normalise to inf/nan in the reference itself
        f = np.where(np.isfinite(f), f, 0.0)
        idx = np.argsort(-f)[:32]                      # the 32 largest entries ...
        idx = np.unique(np.concatenate([idx, sample_idx(len(f))]))   # ... plus 64 random ones
        out[name + "/idx"] = idx.astype(np.int64)
        out[name + "/val"] = f[idx]
        ng = int(round((kw["tend"] - kw["tstart"]) / kw["tstep"]))
        out[name + "/gatesum"] = f.reshape(ng, -1).sum(axis=1)
        meta[name] = dict(raytet=r.get("raytet"), absorbed_frac=r.get("absorbed_frac"),
                          normalizer=r.get("normalizer"), detectedcount=r.get("detectedcount"),
                          total=float(f.sum()), size=int(len(f)), nonfinite=nonfinite)
        print(name, meta[name])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_cases.npz"), **out)


if __name__ == "__main__":
    main()
