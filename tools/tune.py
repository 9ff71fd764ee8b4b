#!/usr/bin/env python
"""Build tuning variants of the library (launch bounds, accumulator type) into build/variants/ here, then time them on the
GPU box:   python tools/tune.py build      (CPU box)
           python tools/tune.py run [workload method photons]   (GPU box; prints one line per variant)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")
VARIANTS = {
    "b128_m8": ["-DMMCB_MAXTHREADS=128", "-DMMCB_MINBLOCKS=8"],
    "b256_m4": ["-DMMCB_MAXTHREADS=256", "-DMMCB_MINBLOCKS=4", "-DMMCB_MINBLOCKS_HP=2"],
    "b64_m16": ["-DMMCB_MAXTHREADS=64", "-DMMCB_MINBLOCKS=16", "-DMMCB_MINBLOCKS_HP=10"],
}


EXTRA_ENV = {}


def main():
    if sys.argv[1] == "build":
        from mmc_b200 import build
        os.makedirs(VDIR, exist_ok=True)
        for tag, flags in VARIANTS.items():
            out = os.path.join(VDIR, "libmmc_b200_%s.so" % tag)
            build.build(force=True, extra=flags, out=out, tag="_" + tag)
            print("built", out)
        return
    wl = sys.argv[2:] or ["sphshells:grid:1e7", "cube60:elem:1e7"]
    for tag in VARIANTS:
        lib = os.path.join(VDIR, "libmmc_b200_%s.so" % tag)
        if not os.path.exists(lib):
            continue
        block = tag.split("_")[0][1:]
        for w in wl:
            for hot in (0,):
                name, method, nph = w.split(":")
                env = dict(os.environ, MMCB_LIB=lib, MMCB_BLOCK=block, MMCB_HOTCACHE=str(hot))
                r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", name, "--method", method, "--photons", nph,
                                    "--steps", "2", "--warmup", "1", "--no-e2e", "--no-cpu-baseline"], env=env, capture_output=True, text=True)
                try:
                    j = json.loads(r.stdout.strip().splitlines()[-1])
                    print(json.dumps(dict(variant=tag, hot=hot, workload=w, photons_per_ms=round(j["value"]),
                                          kernel_ms=round(j["roofline"]["kernel_ms"], 2))), flush=True)
                except Exception:
                    print(json.dumps(dict(variant=tag, workload=w, error=(r.stderr or r.stdout)[-300:])), flush=True)


if __name__ == "__main__":
    main()
