"""Builds libmmc_b200.so (the C-ABI library: CUDA kernels for sm_100a + host layer) in-tree with nvcc.

    python -m mmc_b200.build            # build if stale
    python -m mmc_b200.build --force
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["mmcb_kernel.cu", "mmcb_post.cu", "mmcb_prep.cu", "mmcb_host.cu"]
HEADERS = ["mmcb_types.h", "mmcb_kernel_rp.cuh", os.path.join("..", "..", "include", "mmc_b200.h")]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-use_fast_math", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--cudart", "static"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=(), out=None, tag=""):
    """out/tag: build a tuning variant (extra -D flags) into another file without touching the product library"""
    if out is None and not force and not stale():
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [NVCC] + FLAGS + list(extra) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([NVCC, "-shared", "--cudart", "static", "-o", out or LIB] + objs)
    return out or LIB


def build_pmmc(force=False):
    """integration/pmmc/_pmmc*.so: the reference's Python module name and functions (src/pmmc.cpp:1447-1462) on this engine
    (integration/pmmc_b200.cpp, pybind11, against include/mmc_b200.h only)."""
    import sysconfig
    import pybind11
    root = os.path.dirname(HERE)
    src = os.path.join(root, "integration", "pmmc_b200.cpp")
    out = os.path.join(root, "integration", "pmmc", "_pmmc" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-fvisibility=hidden", "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           "-I", os.path.join(root, "include"), src, "-o", out, "-L", HERE, "-lmmc_b200", "-Wl,-rpath,$ORIGIN/../../mmc_b200"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_pmmc(force="--force" in sys.argv))
