"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/mmc_oracle.c) and a runner
for the unmodified reference binary (oracle/_ref/mmc_ref).  Imported by tests/, bench.py's
cpu_baseline/--impl reference legs and __graft_entry__.smoke() only -- never by mmc_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "mmc_ref")
REF_CUDA_BIN = os.path.join(HERE, "_ref", "mmc_refcuda")
REF_CUDA_MS_BIN = os.path.join(HERE, "_ref", "mmc_refcuda_ms")    # reference CUDA objects behind ref_multislot_main.c

PLUCKER, HAVEL, BADOUEL, BLBADOUEL, GRID = 0, 1, 2, 3, 4
FLUX, FLUENCE, ENERGY, JACOBIAN, WL, WP = 0, 1, 2, 3, 4, 5
SEED_FROM_FILE = -999
METHOD_FLAG = {PLUCKER: "p", HAVEL: "h", BADOUEL: "b", BLBADOUEL: "s", GRID: "g"}
OUTPUT_FLAG = {FLUX: "X", FLUENCE: "F", ENERGY: "E", JACOBIAN: "J", WL: "L", WP: "P"}


def build(force=False):
    src = os.path.join(HERE, "mmc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "mmc_oracle.h"))):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    return LIB


class Mesh(C.Structure):
    _fields_ = [("nn", C.c_int), ("ne", C.c_int), ("nf", C.c_int), ("prop", C.c_int), ("isextdet", C.c_int),
                ("node", C.POINTER(C.c_float)), ("elem", C.POINTER(C.c_int)), ("type", C.POINTER(C.c_int)),
                ("med", C.POINTER(C.c_float)), ("facenb", C.POINTER(C.c_int)), ("evol", C.POINTER(C.c_float)),
                ("nvol", C.POINTER(C.c_float)), ("srcelem", C.POINTER(C.c_int)), ("srcelemlen", C.c_int),
                ("detelem", C.POINTER(C.c_int)), ("detelemlen", C.c_int), ("n", C.POINTER(C.c_float)),
                ("m", C.POINTER(C.c_float)), ("pd", C.POINTER(C.c_float)), ("pm", C.POINTER(C.c_float)),
                ("nmin", C.c_float * 3), ("nmax", C.c_float * 3), ("e0_from_src", C.c_int)]


class Config(C.Structure):
    _fields_ = [("nphoton", C.c_uint64), ("seed", C.c_int), ("nthread", C.c_int),
                ("srcpos", C.c_float * 4), ("srcdir", C.c_float * 4), ("srctype", C.c_int),
                ("srcparam1", C.c_float * 4), ("srcparam2", C.c_float * 4),
                ("srcpattern", C.POINTER(C.c_float)), ("srcnum", C.c_int),
                ("tstart", C.c_float), ("tstep", C.c_float), ("tend", C.c_float), ("e0", C.c_int),
                ("isreflect", C.c_int), ("isnormalized", C.c_int), ("issavedet", C.c_int), ("ismomentum", C.c_int),
                ("issaveexit", C.c_int), ("isspecular", C.c_int), ("issaveseed", C.c_int), ("issaveref", C.c_int),
                ("method", C.c_int), ("basisorder", C.c_int), ("outputtype", C.c_int),
                ("roulettesize", C.c_float), ("minenergy", C.c_float), ("nout", C.c_float),
                ("voidtime", C.c_int), ("unitinmm", C.c_float), ("steps", C.c_float),
                ("detnum", C.c_int), ("detpos", C.POINTER(C.c_float)), ("maxdetphoton", C.c_uint),
                ("photonseed", C.POINTER(C.c_uint64)), ("replayweight", C.POINTER(C.c_float)),
                ("replaytime", C.POINTER(C.c_float)), ("savetraj", C.c_int), ("maxjumpdebug", C.c_uint),
                ("gpu_semantics", C.c_int)]


class Result(C.Structure):
    _fields_ = [("field", C.POINTER(C.c_double)), ("dref", C.POINTER(C.c_double)),
                ("detected", C.POINTER(C.c_float)), ("detseed", C.POINTER(C.c_uint64)),
                ("detectedcount", C.c_uint), ("traj", C.POINTER(C.c_float)), ("trajcount", C.c_uint),
                ("launchweight", C.c_double * 16), ("absorbweight", C.c_double * 16), ("escweight", C.c_double * 16),
                ("raytet", C.c_double), ("normalizer", C.c_double),
                ("maxgate", C.c_int), ("datalen", C.c_int), ("reclen", C.c_int), ("e0", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_mesh_create.restype = C.POINTER(Mesh)
        L.orc_mesh_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_mesh_free.argtypes = [C.POINTER(Mesh)]
        L.orc_mesh_build_tracer.argtypes = [C.POINTER(Mesh), C.c_int]
        L.orc_run.argtypes = [C.POINTER(Mesh), C.POINTER(Config), C.POINTER(Result)]
        L.orc_run.restype = C.c_int
        L.orc_last_error.restype = C.c_char_p
        L.orc_rng_nextf.restype = C.c_float
        L.orc_rng_nextf.argtypes = [C.c_void_p]
        L.orc_rng_seed.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_host_seeds.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_maxgate.argtypes = [C.POINTER(Config)]
        L.orc_datalen.argtypes = [C.POINTER(Mesh), C.POINTER(Config)]
        L.orc_reclen.argtypes = [C.POINTER(Mesh), C.POINTER(Config)]
        L.orc_mesh_initelem.argtypes = [C.POINTER(Mesh), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


DEFAULTS = dict(nphoton=1000, seed=1648335518, nthread=1, srcpos=(0, 0, 0, 0), srcdir=(0, 0, 1, 0), srctype=0,
                srcparam1=(0, 0, 0, 0), srcparam2=(0, 0, 0, 0), srcpattern=None, srcnum=1,
                tstart=0.0, tstep=5e-9, tend=5e-9, e0=0, isreflect=1, isnormalized=1, issavedet=0, ismomentum=0,
                issaveexit=0, isspecular=0, issaveseed=0, issaveref=0, method=BLBADOUEL, basisorder=0,
                outputtype=FLUX, roulettesize=10.0, minenergy=1e-6, nout=1.0, voidtime=1, unitinmm=1.0, steps=1.0,
                detpos=None, maxdetphoton=1000000, photonseed=None, replayweight=None, replaytime=None,
                savetraj=0, maxjumpdebug=100000, gpu_semantics=0)


def host_seeds(seed, count):
    out = np.zeros(count, dtype=np.uint32)
    lib().orc_host_seeds(int(seed), int(count), out.ctypes.data)
    return out


def rng_floats(seed4, n):
    st = np.zeros(2, dtype=np.uint64)
    s4 = np.asarray(seed4, dtype=np.uint32)
    lib().orc_rng_seed(s4.ctypes.data, st.ctypes.data)
    out = np.zeros(n, dtype=np.float32)
    states = np.zeros((n, 2), dtype=np.uint64)
    for i in range(n):
        out[i] = lib().orc_rng_nextf(st.ctypes.data)
        states[i] = st
    return out, states


def _f4(v):
    v = list(v) + [0.0] * (4 - len(v))
    return (C.c_float * 4)(*[float(x) for x in v[:4]])


def run(node, elem, etype, med, facenb=None, evol=None, **kw):
    """Run the oracle.  med: [(mua,mus,g,n)] for media 1..prop (medium 0 is added like mesh_loadmedia).
    Returns a dict with field [maxgate, datalen, srcnum] (float64) and tallies."""
    L = lib()
    p = dict(DEFAULTS)
    unknown = set(kw) - set(p)
    if unknown:
        raise TypeError("unknown oracle options: %s" % sorted(unknown))
    p.update(kw)
    node = np.ascontiguousarray(node, dtype=np.float32)
    elem = np.ascontiguousarray(elem, dtype=np.int32)
    etype = np.ascontiguousarray(etype, dtype=np.int32)
    med = np.asarray(med, dtype=np.float32).reshape(-1, 4)
    prop = len(med)
    medfull = np.ascontiguousarray(np.vstack([[0, 0, 1, p["nout"]], med]), dtype=np.float32)
    fnb = None if facenb is None else np.ascontiguousarray(facenb, dtype=np.int32)
    ev = None if evol is None else np.ascontiguousarray(evol, dtype=np.float32)
    mesh = L.orc_mesh_create(len(node), node.ctypes.data, len(elem), elem.ctypes.data, etype.ctypes.data, prop,
                             medfull.ctypes.data, C.c_float(p["nout"]), C.c_float(p["unitinmm"]),
                             None if fnb is None else fnb.ctypes.data, None if ev is None else ev.ctypes.data)
    try:
        cfg = Config()
        keep = []
        for k in ("nphoton", "seed", "nthread", "srctype", "srcnum", "tstart", "tstep", "tend", "e0", "isreflect",
                  "isnormalized", "issavedet", "ismomentum", "issaveexit", "isspecular", "issaveseed", "issaveref",
                  "method", "basisorder", "outputtype", "roulettesize", "minenergy", "nout", "voidtime", "unitinmm",
                  "steps", "maxdetphoton", "savetraj", "maxjumpdebug", "gpu_semantics"):
            setattr(cfg, k, p[k])
        cfg.srcpos, cfg.srcdir = _f4(p["srcpos"]), _f4(p["srcdir"])
        cfg.srcparam1, cfg.srcparam2 = _f4(p["srcparam1"]), _f4(p["srcparam2"])
        if p["srcpattern"] is not None:
            pat = np.ascontiguousarray(p["srcpattern"], dtype=np.float32)
            keep.append(pat)
            cfg.srcpattern = pat.ctypes.data_as(C.POINTER(C.c_float))
        det = None
        if p["detpos"] is not None and len(p["detpos"]):
            det = np.ascontiguousarray(p["detpos"], dtype=np.float32).reshape(-1, 4)
            cfg.detnum = len(det)
            cfg.detpos = det.ctypes.data_as(C.POINTER(C.c_float))
        if p["photonseed"] is not None:
            ps = np.ascontiguousarray(p["photonseed"]).view(np.uint64).reshape(-1, 2)
            keep.append(ps)
            cfg.photonseed = ps.ctypes.data_as(C.POINTER(C.c_uint64))
            rw = np.ascontiguousarray(p["replayweight"], dtype=np.float32).copy()
            rt = np.ascontiguousarray(p["replaytime"], dtype=np.float32)
            keep += [rw, rt]
            cfg.replayweight = rw.ctypes.data_as(C.POINTER(C.c_float))
            cfg.replaytime = rt.ctypes.data_as(C.POINTER(C.c_float))
        maxgate = L.orc_maxgate(C.byref(cfg))
        if p["method"] == GRID:
            cfg.basisorder = 0
        datalen = L.orc_datalen(mesh, C.byref(cfg))
        res = Result()
        field = np.zeros((maxgate, datalen, p["srcnum"]), dtype=np.float64)
        res.field = field.ctypes.data_as(C.POINTER(C.c_double))
        # exterior faces <= 4*ne
        dref = None
        if p["issaveref"]:
            dref = np.zeros((maxgate, 4 * len(elem)), dtype=np.float64)
        reclen_max = 3 * prop + 8
        detected = np.zeros((int(p["maxdetphoton"]) if p["issavedet"] else 1, reclen_max), dtype=np.float32)
        detseed = np.zeros((len(detected), 2), dtype=np.uint64)
        res.detected = detected.ctypes.data_as(C.POINTER(C.c_float))
        res.detseed = detseed.ctypes.data_as(C.POINTER(C.c_uint64))
        traj = np.zeros((int(p["maxjumpdebug"]) if p["savetraj"] else 1, 6), dtype=np.float32)
        res.traj = traj.ctypes.data_as(C.POINTER(C.c_float))
        if dref is not None:
            # nf is only known after prep; give the oracle a big enough flat buffer and reshape after
            res.dref = dref.ctypes.data_as(C.POINTER(C.c_double))
        rc = L.orc_run(mesh, C.byref(cfg), C.byref(res))
        if rc != 0:
            raise RuntimeError("oracle: " + L.orc_last_error().decode())
        M = mesh.contents
        nd = min(res.detectedcount, len(detected))
        out = dict(field=field, maxgate=res.maxgate, datalen=res.datalen, reclen=res.reclen, e0=res.e0,
                   launchweight=np.array(res.launchweight[:p["srcnum"]]),
                   absorbweight=np.array(res.absorbweight[:p["srcnum"]]),
                   escweight=np.array(res.escweight[:p["srcnum"]]), raytet=res.raytet, normalizer=res.normalizer,
                   detectedcount=res.detectedcount,
                   detected=detected.reshape(-1)[:nd * res.reclen].reshape(nd, res.reclen).copy(),
                   detseed=detseed[:nd].copy(), traj=traj[:res.trajcount].copy(), nf=M.nf,
                   elem=np.ctypeslib.as_array(M.elem, shape=(M.ne, 4)).copy(),
                   facenb=np.ctypeslib.as_array(M.facenb, shape=(M.ne, 4)).copy(),
                   evol=np.ctypeslib.as_array(M.evol, shape=(M.ne,)).copy(),
                   nvol=np.ctypeslib.as_array(M.nvol, shape=(M.nn,)).copy(),
                   type=np.ctypeslib.as_array(M.type, shape=(M.ne,)).copy())
        if dref is not None:
            out["dref"] = dref.reshape(-1)[:res.maxgate * M.nf].reshape(res.maxgate, M.nf).copy()
        if M.n:
            out["normals"] = np.ctypeslib.as_array(M.n, shape=(M.ne, 16)).copy()
        if M.m:
            out["havel"] = np.ctypeslib.as_array(M.m, shape=(M.ne, 48)).copy()
        return out
    finally:
        L.orc_mesh_free(mesh)


# ---------------------------------------------------------------------------------------------------
# the unmodified reference binary
# ---------------------------------------------------------------------------------------------------
def ref_available(cuda=False, multislot=False):
    return os.path.exists(REF_CUDA_MS_BIN if multislot else (REF_CUDA_BIN if cuda else REF_BIN))


def write_mesh_files(dirname, tag, node, elem, etype, med, evol=None):
    """node_/elem_/prop_ text files (formats: SURVEY.md Appendix D; src/mmc_mesh.c:455-550,668-713)."""
    node = np.asarray(node)
    elem = np.asarray(elem)
    with open(os.path.join(dirname, "node_%s.dat" % tag), "w") as f:
        f.write("1 %d\n" % len(node))
        for i, p in enumerate(node):
            f.write("%d %.9g %.9g %.9g\n" % (i + 1, p[0], p[1], p[2]))
    with open(os.path.join(dirname, "elem_%s.dat" % tag), "w") as f:
        f.write("1 %d\n" % len(elem))
        rows = np.column_stack([np.arange(1, len(elem) + 1), elem, etype])
        np.savetxt(f, rows, fmt="%d")
    if evol is not None:   # velem_<id>.dat: given volumes => the loader does not re-orient elements (src/mmc_mesh.c:723-761)
        with open(os.path.join(dirname, "velem_%s.dat" % tag), "w") as f:
            f.write("1 %d\n" % len(elem))
            for i, v in enumerate(np.asarray(evol, dtype=np.float64)):
                f.write("%d %.9e\n" % (i + 1, v))
    med = np.asarray(med, dtype=np.float64).reshape(-1, 4)
    with open(os.path.join(dirname, "prop_%s.dat" % tag), "w") as f:
        f.write("1 %d\n" % len(med))
        for i, m in enumerate(med):
            f.write("%d %.9g %.9g %.9g %.9g\n" % (i + 1, m[0], m[1], m[2], m[3]))


def run_ref(node, elem, etype, med, *, nthread=1, cuda=False, extra_args=(), timeout=3600, keep_dir=None, evol=None, check=True,
            expect="out.bin", multislot=False, binary=None, **kw):
    """Run oracle/_ref/mmc_ref (or mmc_refcuda, or `binary`: e.g. the drop-in CLI oracle/_ref/mmc_b200cli) on the same inputs;
    returns dict(field, absorbed_frac, speed, ...)."""
    p = dict(DEFAULTS)
    p.update(kw)
    binp = binary or (REF_CUDA_MS_BIN if multislot else (REF_CUDA_BIN if cuda else REF_BIN))
    tmp = keep_dir or tempfile.mkdtemp(prefix="mmcref_")
    os.makedirs(tmp, exist_ok=True)
    tag = "t"
    write_mesh_files(tmp, tag, node, elem, etype, med, evol)
    det = np.zeros((0, 4)) if p["detpos"] is None else np.asarray(p["detpos"], dtype=np.float64).reshape(-1, 4)
    with open(os.path.join(tmp, "in.inp"), "w") as f:
        f.write("%d\n%d\n" % (p["nphoton"], p["seed"]))
        f.write("%.9g %.9g %.9g\n" % tuple(p["srcpos"][:3]))
        sd = list(p["srcdir"]) + [0.0] * (4 - len(p["srcdir"]))
        f.write("%.9g %.9g %.9g %.9g\n" % tuple(sd[:4]))
        f.write("%.9g %.9g %.9g\n" % (p["tstart"], p["tend"], p["tstep"]))
        f.write("%s\n%d\n" % (tag, p["e0"]))
        f.write("%d %.9g\n" % (len(det), det[0, 3] if len(det) else 1.0))
        for d in det:
            f.write("%.9g %.9g %.9g %.9g\n" % tuple(d))
        if p["srctype"] != 0:
            names = ["pencil", "isotropic", "cone", "gaussian", "planar", "pattern", "fourier", "arcsine", "disk",
                     "fourierx", "fourierx2d", "zgaussian", "line", "slit"]
            f.write("%s\n" % names[p["srctype"]])
            f.write("%.9g %.9g %.9g %.9g\n" % tuple(list(p["srcparam1"]) + [0] * (4 - len(p["srcparam1"]))))
            f.write("%.9g %.9g %.9g %.9g\n" % tuple(list(p["srcparam2"]) + [0] * (4 - len(p["srcparam2"]))))
            if p["srctype"] == 5:
                pat = np.asarray(p["srcpattern"], dtype=np.float32)
                pat.tofile(os.path.join(tmp, "pattern.bin"))
                f.write("pattern.bin %d\n" % p["srcnum"])
    args = [binp, "-f", "in.inp", "-s", "out", "-M", METHOD_FLAG[p["method"]], "-b", str(p["isreflect"]),
            "-C", str(p["basisorder"]), "-U", str(p["isnormalized"]), "-O", OUTPUT_FLAG[p["outputtype"]],
            "-F", "bin", "-D", "T", "-S", "1", "-e", "%.9g" % p["minenergy"], "-d", str(p["issavedet"]),
            "-x", str(p["issaveexit"]), "--momentum", str(p["ismomentum"]), "-q", str(p["issaveseed"]),
            "-u", "%.9g" % p["unitinmm"], "-H", str(p["maxdetphoton"]), "-V", str(p["isspecular"]),
            "-X", str(p["issaveref"]), "-n", str(p["nphoton"]), "-E", str(p["seed"])]
    if p["method"] == GRID:
        args += ["--gridsize", "%.9g" % p["steps"]]
    if p["nout"] != 1.0:
        args += ["-j", '{"Forward":{"N0":%.9g}}' % p["nout"]]
    args += ["-c", "cuda", "-G", "1"] if cuda else ["-c", "sse"]
    args += list(extra_args)
    env = dict(os.environ, OMP_NUM_THREADS=str(nthread))
    r = subprocess.run(args, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    log = r.stdout.decode(errors="replace")
    if r.returncode != 0 and (check or not os.path.exists(os.path.join(tmp, expect))):     # check=False: results (`expect`) were saved before a crash at exit
        raise RuntimeError("reference failed (%d): %s\n%s" % (r.returncode, " ".join(args), log[-3000:]))
    out = dict(log=log, dir=tmp)
    m = re.search(r"total simulated energy:\s*([0-9.eE+-]+)\s*absorbed:\s*(?:\x1b\[[0-9;]*m)*([0-9.eE+-]+)%", log)
    if m:
        out["launched"] = float(m.group(1))
        out["absorbed_frac"] = float(m.group(2)) / 100.0
    m = re.search(r"normalizor=([0-9.eE+-]+)", log)
    if m:
        out["normalizer"] = float(m.group(1))
    m = re.search(r"([0-9.]+) photon/ms(?:\x1b\[[0-9;]*m)*, ([0-9.]+) ray-tetrahedron", log)
    if m:
        out["speed"] = float(m.group(1))
        out["raytet"] = float(m.group(2))
    m = re.search(r"MCX simulation speed: ([0-9.]+) photon/ms", log)
    if m:
        out["speed"] = float(m.group(1))
    m = re.search(r"kernel complete:\s*([0-9]+) ms", log)
    if m:
        out["kernel_ms"] = float(m.group(1))
    m = re.search(r"detected (\d+) photons", log)
    if m:
        out["detectedcount"] = int(m.group(1))
    fb = os.path.join(tmp, "out.bin")
    if os.path.exists(fb):
        out["field_flat"] = np.fromfile(fb, dtype=np.float64)
    mch = os.path.join(tmp, "out.mch")
    if os.path.exists(mch):
        out["mch"] = mch
    return out
