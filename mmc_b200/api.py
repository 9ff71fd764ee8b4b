"""Python host side of the mmc_b200 engine: a ctypes binding of the C-ABI (include/mmc_b200.h) and a
`run(cfg)` entry point that mirrors the reference's Python binding `pmmc.run(cfg)` (src/pmmc.cpp:903-1289):
same dict keys (`nphoton`, `node`, `elem`, `elemprop`, `prop`, `srcpos`, `srcdir`, `tstart`, ... ) and the
same result keys (`flux`, `detp`, `seeds`, `traj`, `dref`).

There is no CPU path: every call that computes goes through libmmc_b200.so and raises MMCError when the
library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("MMCB_LIB") or os.path.join(HERE, "libmmc_b200.so")   # MMCB_LIB: tuning builds (tools/tune.py)

SEED_FROM_FILE = -999
MAX_SRCNUM = 16

# string tables of the reference front-ends (src/mmc_utils.c:139-175)
SRCTYPES = ["pencil", "isotropic", "cone", "gaussian", "planar", "pattern", "fourier", "arcsine", "disk",
            "fourierx", "fourierx2d", "zgaussian", "line", "slit"]
METHODS = {"plucker": 0, "p": 0, "havel": 1, "h": 1, "badouel": 2, "b": 2, "elem": 3, "s": 3, "blbadouel": 3,
           "grid": 4, "g": 4}
OUTPUTTYPES = {"flux": 0, "x": 0, "fluence": 1, "f": 1, "energy": 2, "e": 2, "jacobian": 3, "j": 3,
               "wl": 4, "l": 4, "wp": 5, "p": 5,
               # RF forward and adjoint Jacobians (src/mmc_utils.c:4444-4449)
               "rf": 6, "r": 6, "adjoint": 8, "a": 8, "adjointd": 9, "d": 9, "adjointmus": 10, "u": 10, "adjointmusp": 11, "v": 11,
               "adjointmuad": 12, "w": 12, "adjointmuamusp": 13, "q": 13}


class MMCError(RuntimeError):
    """Raised where the reference would call mcx_error(id, msg, file, line) (src/mmc_utils.c:1426-1442)."""

    def __init__(self, code, msg):
        super().__init__("MMC ERROR(%d):%s" % (code, msg))
        self.code = code


class Medium(C.Structure):
    _fields_ = [("mua", C.c_float), ("mus", C.c_float), ("g", C.c_float), ("n", C.c_float)]


class Mesh(C.Structure):
    _fields_ = [("nn", C.c_int), ("ne", C.c_int), ("prop", C.c_int),
                ("node", C.c_void_p), ("elem", C.c_void_p), ("type", C.c_void_p), ("med", C.c_void_p),
                ("facenb", C.c_void_p), ("evol", C.c_void_p), ("nvol", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("nphoton", C.c_uint64), ("seed", C.c_int),
                ("srcpos", C.c_float * 4), ("srcdir", C.c_float * 4), ("srctype", C.c_int),
                ("srcparam1", C.c_float * 4), ("srcparam2", C.c_float * 4),
                ("srcpattern", C.c_void_p), ("srcnum", C.c_int),
                ("tstart", C.c_float), ("tstep", C.c_float), ("tend", C.c_float), ("e0", C.c_int),
                ("isreflect", C.c_int), ("isnormalized", C.c_int),
                ("issavedet", C.c_int), ("ismomentum", C.c_int), ("issaveexit", C.c_int), ("issaveseed", C.c_int),
                ("isspecular", C.c_int), ("issaveref", C.c_int),
                ("method", C.c_int), ("basisorder", C.c_int), ("outputtype", C.c_int),
                ("roulettesize", C.c_float), ("minenergy", C.c_float), ("nout", C.c_float),
                ("voidtime", C.c_int), ("unitinmm", C.c_float), ("steps", C.c_float),
                ("detnum", C.c_int), ("detpos", C.c_void_p), ("maxdetphoton", C.c_uint),
                ("photonseed", C.c_void_p), ("replayweight", C.c_void_p), ("replaytime", C.c_void_p),
                ("savetraj", C.c_int), ("maxjumpdebug", C.c_uint),
                ("nthread", C.c_int), ("nblocksize", C.c_int), ("schedule", C.c_int), ("respin", C.c_int), ("hotcache", C.c_int),
                ("omega", C.c_float), ("srcid", C.c_int), ("extrasrclen", C.c_int), ("srcdata", C.c_void_p), ("detdir", C.c_void_p),
                ("adjointmode", C.c_int), ("nodemua", C.c_void_p), ("nodemusp", C.c_void_p)]


class GpuInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("id", C.c_int), ("devcount", C.c_int), ("major", C.c_int), ("minor", C.c_int),
                ("globalmem", C.c_size_t), ("constmem", C.c_size_t), ("sharedmem", C.c_size_t),
                ("regcount", C.c_int), ("clock", C.c_int), ("sm", C.c_int), ("core", C.c_int),
                ("autoblock", C.c_size_t), ("autothread", C.c_size_t), ("maxgate", C.c_int), ("maxmpthread", C.c_int)]


class Output(C.Structure):
    _fields_ = [("field", C.c_void_p), ("dref", C.c_void_p), ("detected", C.c_void_p), ("detseed", C.c_void_p),
                ("traj", C.c_void_p),
                ("detectedcount", C.c_uint), ("detectedtotal", C.c_uint), ("trajcount", C.c_uint),
                ("energytot", C.c_double * 16), ("energyesc", C.c_double * 16),
                ("raytet", C.c_double), ("normalizer", C.c_double), ("kernel_ms", C.c_float), ("e0", C.c_int),
                ("field_im", C.c_void_p), ("jacob", C.c_void_p), ("overwrite", C.c_int)]


class Sizes(C.Structure):
    _fields_ = [("maxgate", C.c_int), ("datalen", C.c_int), ("reclen", C.c_int), ("nf", C.c_int), ("srcnum", C.c_int),
                ("dim", C.c_int * 3), ("fieldlen", C.c_size_t), ("nslots", C.c_int), ("adj_ns", C.c_int), ("adj_nd", C.c_int),
                ("jacoblen", C.c_size_t)]


class DevPtrs(C.Structure):
    _fields_ = [("field", C.c_void_p), ("fieldlen", C.c_size_t), ("field_is_double", C.c_int),
                ("energy", C.c_void_p), ("raytet", C.c_void_p),
                ("detected", C.c_void_p), ("detcount", C.c_void_p), ("reclen", C.c_int),
                ("detseed", C.c_void_p), ("dref", C.c_void_p), ("dreflen", C.c_size_t), ("field_im", C.c_void_p)]


EXPORTS = ["mmcb_version", "mmcb_last_error", "mmcb_list_gpu", "mmcb_query_sizes", "mmcb_run_simulation",
           "mmcb_create", "mmcb_set_field_buffer", "mmcb_launch", "mmcb_sync", "mmcb_last_kernel_ms",
           "mmcb_get_devptrs", "mmcb_get_sizes", "mmcb_fetch", "mmcb_reset", "mmcb_destroy", "mmcb_get_tables", "mmcb_run_session",
           "mmcb_mesh_volumes", "mmcb_mesh_facenb", "mmcb_mesh_initelem", "mmcb_host_seeds", "mmcb_rng_selftest",
           "mmcb_run_multi", "mmcb_photon_shares"]

_lib = None


def lib():
    """Load libmmc_b200.so; fails loudly when it has not been built (python -m mmc_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise MMCError(-1, "libmmc_b200.so is missing -- build it with `python -m mmc_b200.build` "
                               "(the engine has no CPU fallback)")
        L = C.CDLL(LIBPATH)
        L.mmcb_last_error.restype = C.c_char_p
        L.mmcb_create.restype = C.c_void_p
        L.mmcb_create.argtypes = [C.POINTER(Config), C.POINTER(Mesh), C.c_int]
        L.mmcb_destroy.argtypes = [C.c_void_p]
        L.mmcb_destroy.restype = None
        L.mmcb_query_sizes.argtypes = [C.POINTER(Config), C.POINTER(Mesh), C.POINTER(Sizes)]
        L.mmcb_run_simulation.argtypes = [C.POINTER(Config), C.POINTER(Mesh), C.c_int, C.POINTER(Output)]
        L.mmcb_set_field_buffer.argtypes = [C.c_void_p, C.c_void_p]
        L.mmcb_launch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.mmcb_sync.argtypes = [C.c_void_p]
        L.mmcb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.mmcb_get_devptrs.argtypes = [C.c_void_p, C.POINTER(DevPtrs)]
        L.mmcb_get_sizes.argtypes = [C.c_void_p, C.POINTER(Sizes)]
        L.mmcb_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Output)]
        L.mmcb_reset.argtypes = [C.c_void_p]
        L.mmcb_run_session.argtypes = [C.c_void_p, C.POINTER(Output)]
        L.mmcb_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mmcb_list_gpu.argtypes = [C.POINTER(GpuInfo), C.c_int]
        L.mmcb_mesh_volumes.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mmcb_mesh_facenb.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.mmcb_mesh_initelem.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mmcb_host_seeds.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
        L.mmcb_host_seeds.restype = None
        L.mmcb_rng_selftest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mmcb_run_multi.argtypes = [C.POINTER(Config), C.POINTER(Mesh), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Output)]
        L.mmcb_photon_shares.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.mmcb_photon_shares.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise MMCError(rc, lib().mmcb_last_error().decode(errors="replace"))


def version():
    return lib().mmcb_version()


def gpuinfo():
    """pmmc.gpuinfo(): list of dicts, one per CUDA device (mcx_list_cu_gpu)."""
    arr = (GpuInfo * 16)()
    n = lib().mmcb_list_gpu(arr, 16)
    if n < 0:
        _check(n)
    keys = ("id", "devcount", "major", "minor", "globalmem", "constmem", "sharedmem", "regcount", "clock", "sm",
            "core", "autoblock", "autothread", "maxgate", "maxmpthread")
    return [dict(name=arr[i].name.decode(), **{k: getattr(arr[i], k) for k in keys}) for i in range(min(n, 16))]


def host_seeds(seed, count, skip=0):
    out = np.zeros(count, dtype=np.uint32)
    lib().mmcb_host_seeds(int(seed), int(skip), int(count), out.ctypes.data)
    return out


def rng_selftest(seeds4, ndraw):
    """Device draws of the kernel's xorshift128+ generator: returns (floats [nstream, ndraw], final states [nstream, 2])."""
    s4 = np.ascontiguousarray(seeds4, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros((len(s4), ndraw), dtype=np.float32)
    st = np.zeros((len(s4), 2), dtype=np.uint64)
    _check(lib().mmcb_rng_selftest(s4.ctypes.data, len(s4), int(ndraw), out.ctypes.data, st.ctypes.data))
    return out, st


def mesh_facenb(elem):
    elem = np.ascontiguousarray(elem, dtype=np.int32)
    out = np.zeros_like(elem)
    _check(lib().mmcb_mesh_facenb(len(elem), elem.ctypes.data, out.ctypes.data))
    return out


def mesh_volumes(node, elem, etype=None):
    node = np.ascontiguousarray(node, dtype=np.float32)
    elem = np.ascontiguousarray(elem, dtype=np.int32).copy()
    et = None if etype is None else np.ascontiguousarray(etype, dtype=np.int32)
    evol = np.zeros(len(elem), dtype=np.float32)
    nvol = np.zeros(len(node), dtype=np.float32)
    _check(lib().mmcb_mesh_volumes(len(node), node.ctypes.data, len(elem), elem.ctypes.data,
                                   None if et is None else et.ctypes.data, evol.ctypes.data, nvol.ctypes.data))
    return elem, evol, nvol


def mesh_initelem(node, elem, srcpos):
    node = np.ascontiguousarray(node, dtype=np.float32)
    elem = np.ascontiguousarray(elem, dtype=np.int32)
    sp = np.ascontiguousarray(srcpos, dtype=np.float32)
    bary = np.zeros(4, dtype=np.float32)
    e0 = lib().mmcb_mesh_initelem(len(node), node.ctypes.data, len(elem), elem.ctypes.data, sp.ctypes.data, bary.ctypes.data)
    return e0, bary


def _vec4(v):
    v = [float(x) for x in np.asarray(v, dtype=np.float64).ravel()]
    v = (v + [0.0] * 4)[:4]
    return (C.c_float * 4)(*v)


DEFAULTS = dict(nphoton=0, seed=0x623F9A9E, srcpos=(0, 0, 0), srcdir=(0, 0, 1, 0), srctype="pencil",
                srcparam1=(0, 0, 0, 0), srcparam2=(0, 0, 0, 0), srcpattern=None, srcnum=1,
                tstart=0.0, tstep=0.0, tend=0.0, e0=0, isreflect=1, isnormalized=1, issavedet=0, ismomentum=0,
                issaveexit=0, issaveseed=0, isspecular=0, issaveref=0, method="elem", basisorder=1,
                outputtype="flux", roulettesize=10.0, minenergy=1e-6, nout=1.0, voidtime=1, unitinmm=1.0,
                steps=(1.0, 1.0, 1.0), detpos=None, maxdetphoton=1000000, maxjumpdebug=10000000,
                debuglevel="", nthread=0, nblocksize=0, schedule=0, respin=1, hotcache=0, gpuid=1,
                replayseed=None, replayweight=None, replaytime=None,
                omega=0.0, srcid=0, srcdata=None, detdir=None, adjointmode=0, nodemua=None, nodemusp=None)


class Problem:
    """Marshals a pmmc-style cfg dict into the C-ABI structs and keeps the numpy buffers alive."""

    def __init__(self, cfg):
        p = dict(DEFAULTS)
        unknown = set(cfg) - set(p) - {"node", "elem", "elemprop", "prop", "facenb", "evol", "nvol", "session",
                                       "compute", "isatomic", "issave2pt", "flog", "workload", "isref3", "optlevel"}
        if unknown:
            raise MMCError(-2, "unknown cfg fields: %s" % ", ".join(sorted(unknown)))
        p.update({k: v for k, v in cfg.items() if k in p})
        for k in ("node", "elem", "prop"):
            if k not in cfg:
                raise MMCError(-2, "cfg.%s is required" % k)
        self.keep = []
        node = np.ascontiguousarray(cfg["node"], dtype=np.float32)
        elem = np.asarray(cfg["elem"])
        if elem.ndim != 2 or elem.shape[1] < 4:
            raise MMCError(-2, "the 'elem' field must have 4 or 5 columns")
        if "elemprop" in cfg:
            etype = np.ascontiguousarray(cfg["elemprop"], dtype=np.int32).ravel()
        elif elem.shape[1] >= 5:
            etype = np.ascontiguousarray(elem[:, 4], dtype=np.int32)     # pmmc: 5th column = label
        else:
            raise MMCError(-2, "cfg.elemprop is missing")
        elem = np.ascontiguousarray(elem[:, :4], dtype=np.int32)
        if node.ndim != 2 or node.shape[1] != 3:
            raise MMCError(-2, "the 'node' field must have 3 columns (x,y,z)")
        if len(etype) != len(elem):
            raise MMCError(-2, "elemprop and elem differ in length")
        prop = np.ascontiguousarray(cfg["prop"], dtype=np.float32).reshape(-1, 4)   # row 0 = background, like pmmc
        if len(prop) < 2:
            raise MMCError(-2, "cfg.prop needs the background row plus at least one medium")
        self.node, self.elem, self.etype, self.prop = node, elem, etype, prop
        m = Mesh()
        m.nn, m.ne, m.prop = len(node), len(elem), len(prop) - 1
        m.node, m.elem, m.type, m.med = node.ctypes.data, elem.ctypes.data, etype.ctypes.data, prop.ctypes.data
        for k, dt in (("facenb", np.int32), ("evol", np.float32), ("nvol", np.float32)):
            if cfg.get(k) is not None:
                a = np.ascontiguousarray(cfg[k], dtype=dt)
                self.keep.append(a)
                setattr(m, k, a.ctypes.data)
        self.mesh = m
        c = Config()
        c.nphoton = int(p["nphoton"])
        c.seed = int(p["seed"]) if not isinstance(p["seed"], (np.ndarray, list, tuple)) else SEED_FROM_FILE
        c.srcpos, c.srcdir = _vec4(p["srcpos"]), _vec4(p["srcdir"])
        st = p["srctype"]
        c.srctype = SRCTYPES.index(st.lower()) if isinstance(st, str) else int(st)
        c.srcparam1, c.srcparam2 = _vec4(p["srcparam1"]), _vec4(p["srcparam2"])
        c.srcnum = int(p["srcnum"])
        if p["srcpattern"] is not None:
            pat = np.ascontiguousarray(p["srcpattern"], dtype=np.float32)
            self.keep.append(pat)
            c.srcpattern = pat.ctypes.data
        for k in ("tstart", "tstep", "tend", "roulettesize", "minenergy", "nout", "unitinmm"):
            setattr(c, k, float(p[k]))
        for k in ("e0", "isreflect", "isnormalized", "issavedet", "ismomentum", "issaveexit", "issaveseed", "isspecular",
                  "issaveref", "basisorder", "voidtime", "maxdetphoton", "maxjumpdebug", "nthread", "nblocksize",
                  "schedule", "respin", "hotcache"):
            setattr(c, k, int(p[k]))
        me, ot = p["method"], p["outputtype"]
        c.method = METHODS[me.lower()] if isinstance(me, str) else int(me)
        c.outputtype = OUTPUTTYPES[ot.lower()] if isinstance(ot, str) else int(ot)
        steps = np.atleast_1d(np.asarray(p["steps"], dtype=np.float64))
        if len(steps) > 1 and not (steps[0] == steps[1] == steps[2]):
            raise MMCError(-2, "MMC dual-grid algorithm currently does not support anisotropic voxels")
        c.steps = float(steps[0])
        if p["detpos"] is not None and len(p["detpos"]):
            det = np.ascontiguousarray(p["detpos"], dtype=np.float32).reshape(-1, 4)
            self.keep.append(det)
            c.detnum, c.detpos = len(det), det.ctypes.data
        c.savetraj = 1 if ("M" in str(p["debuglevel"]).upper()) else 0
        c.omega, c.srcid, c.adjointmode = float(p["omega"]), int(p["srcid"]), int(p["adjointmode"])
        if p["srcdata"] is not None:      # rows of 16 floats: srcpos(4) srcdir(4) srcparam1(4) srcparam2(4) (ExtraSrc, src/mmc_utils.h:147-152)
            sd = np.ascontiguousarray(p["srcdata"], dtype=np.float32).reshape(-1, 16)
            self.keep.append(sd)
            c.extrasrclen, c.srcdata = len(sd), sd.ctypes.data
        for k in ("nodemua", "nodemusp"):       # per-node optical properties (cfg.nodemua / cfg.nodemusp of pmmc)
            if p[k] is not None:
                a = np.ascontiguousarray(p[k], dtype=np.float32).ravel()
                if len(a) != len(node):
                    raise MMCError(-2, "%s needs one value per node" % k)
                self.keep.append(a)
                setattr(c, k, a.ctypes.data)
        if p["detdir"] is not None:
            dd = np.ascontiguousarray(p["detdir"], dtype=np.float32).reshape(-1, 4)
            if dd.shape[0] != c.detnum:
                raise MMCError(-2, "detdir needs one row (nx, ny, nz, focal length) per detector")
            self.keep.append(dd)
            c.detdir = dd.ctypes.data
        if p["replayseed"] is not None:
            rs = np.ascontiguousarray(p["replayseed"]).view(np.uint64).reshape(-1, 2)
            rw = np.ascontiguousarray(p["replayweight"], dtype=np.float32)
            rt = np.ascontiguousarray(p["replaytime"], dtype=np.float32)
            self.keep += [rs, rw, rt]
            c.seed = SEED_FROM_FILE
            c.nphoton = len(rs)
            c.photonseed, c.replayweight, c.replaytime = rs.ctypes.data, rw.ctypes.data, rt.ctypes.data
        self.cfg = c
        self.device = int(p["gpuid"]) - 1
        self.params = p

    def sizes(self):
        s = Sizes()
        _check(lib().mmcb_query_sizes(C.byref(self.cfg), C.byref(self.mesh), C.byref(s)))
        return s


class _OutBuffers:
    def __init__(self, prob, sz):
        c = prob.cfg
        self.field = np.empty(sz.fieldlen, dtype=np.float64)      # overwrite=1: filled by the library
        self.dref = np.zeros(max(1, sz.nf * sz.maxgate), dtype=np.float64) if c.issaveref else None
        nd = int(c.maxdetphoton) if c.issavedet else 0
        self.detected = np.zeros((max(nd, 1), sz.reclen), dtype=np.float32)
        self.detseed = np.zeros((max(nd, 1), 2), dtype=np.uint64)
        self.traj = np.zeros((int(c.maxjumpdebug) if c.savetraj else 1, 6), dtype=np.float32)
        self.field_im = np.zeros(sz.fieldlen, dtype=np.float64) if (c.omega > 0 and c.seed != SEED_FROM_FILE) else None
        self.jacob = np.zeros(sz.jacoblen, dtype=np.float32) if sz.jacoblen else None
        o = Output()
        o.field = self.field.ctypes.data
        o.dref = self.dref.ctypes.data if self.dref is not None else None
        o.detected, o.detseed, o.traj = self.detected.ctypes.data, self.detseed.ctypes.data, self.traj.ctypes.data
        o.field_im = self.field_im.ctypes.data if self.field_im is not None else None
        o.jacob = self.jacob.ctypes.data if self.jacob is not None else None
        o.overwrite = 1          # fresh buffers: the library stores instead of accumulating (no host pass over the volume)
        self.out = o
        self.sz = sz

    def result(self, prob):
        o, sz, c = self.out, self.sz, prob.cfg
        srcnum = sz.srcnum
        if sz.nslots > 1:
            return self._result_slots(prob)
        flux = self.field.reshape(sz.maxgate, sz.datalen, srcnum)
        if prob.cfg.method == 4:
            # grid output: x fastest (idx = iz*dim.y*dim.x + iy*dim.x + ix), gates last like pmmc's 'flux'
            vol = flux.reshape(sz.maxgate, sz.dim[2], sz.dim[1], sz.dim[0], srcnum)
            fl = np.transpose(vol, (3, 2, 1, 0, 4))
        else:
            fl = np.transpose(flux, (1, 0, 2))
        if srcnum == 1:
            fl = fl[..., 0]
        res = dict(flux=fl, raw=flux, energytot=np.array(o.energytot[:srcnum]), energyesc=np.array(o.energyesc[:srcnum]),
                   raytet=o.raytet, normalizer=o.normalizer, kernel_ms=o.kernel_ms, e0=o.e0,
                   detectedtotal=o.detectedtotal, maxgate=sz.maxgate, datalen=sz.datalen, reclen=sz.reclen, nf=sz.nf)
        res["energyabs"] = res["energytot"] - res["energyesc"]
        if c.issavedet:
            res["detp"] = self.detected[:o.detectedcount].copy()
            if c.issaveseed:
                res["seeds"] = self.detseed[:o.detectedcount].copy()
        if c.savetraj:
            res["traj"] = self.traj[:o.trajcount].copy()
        if self.dref is not None:
            res["dref"] = self.dref.reshape(sz.maxgate, sz.nf).copy()
        if self.field_im is not None:
            res["raw_im"] = self.field_im.reshape(sz.maxgate, sz.datalen, srcnum)
        return res

    def _result_slots(self, prob):
        """Multi-slot (adjoint) run: one [maxgate, datalen] block per source slot (src/mmc_cu_host.cu:930-938); 'jacob' holds the
        adjoint Jacobian components as [component, Ns*Nd, datalen] (component order: see mmcb_output.jacob)."""
        o, sz, c = self.out, self.sz, prob.cfg
        raw = self.field.reshape(sz.nslots, sz.maxgate, sz.datalen)
        res = dict(raw=raw, flux=raw, nslots=sz.nslots, energytot=np.array(o.energytot[:1]), energyesc=np.array(o.energyesc[:1]),
                   raytet=o.raytet, normalizer=o.normalizer, kernel_ms=o.kernel_ms, e0=o.e0, detectedtotal=o.detectedtotal,
                   maxgate=sz.maxgate, datalen=sz.datalen, reclen=sz.reclen, nf=sz.nf, dim=tuple(sz.dim))
        res["energyabs"] = res["energytot"] - res["energyesc"]
        if self.field_im is not None:
            res["raw_im"] = self.field_im.reshape(sz.nslots, sz.maxgate, sz.datalen)
        if self.jacob is not None:
            res["jacob"] = self.jacob.reshape(-1, sz.adj_ns * sz.adj_nd, sz.datalen)
            res["adj_ns"], res["adj_nd"] = sz.adj_ns, sz.adj_nd
        if c.issavedet:
            res["detp"] = self.detected[:o.detectedcount].copy()
        return res


def run(cfg):
    """Run one simulation, host buffers in, host buffers out -- the equivalent of pmmc.run(cfg) with cfg['compute']='cuda'
    (src/pmmc.cpp:1060-1075).  Same work as the one-call mmcb_run_simulation (which the reference-side stub uses), spelled
    with the session calls so that the output buffers are sized from the session instead of a second mesh preparation."""
    prob = Problem(cfg)
    L = lib()
    h = L.mmcb_create(C.byref(prob.cfg), C.byref(prob.mesh), prob.device)
    if not h:
        raise MMCError(-1, L.mmcb_last_error().decode(errors="replace"))
    try:
        sz = Sizes()
        _check(L.mmcb_get_sizes(h, C.byref(sz)))
        buf = _OutBuffers(prob, sz)
        # all respins (src/mmc_cu_host.cu:656,893-906) + fetch; the library pre-faults a large result array while the last launch runs
        _check(L.mmcb_run_session(h, C.byref(buf.out)))
    finally:
        L.mmcb_destroy(h)
    return buf.result(prob)


def run_onecall(cfg):
    """The same through the single C entry point mmcb_run_simulation (what integration/mmc_cu_host_b200.cpp calls)."""
    prob = Problem(cfg)
    sz = prob.sizes()
    buf = _OutBuffers(prob, sz)
    _check(lib().mmcb_run_simulation(C.byref(prob.cfg), C.byref(prob.mesh), prob.device, C.byref(buf.out)))
    return buf.result(prob)


def photon_shares(nphoton, ndev, workload=None):
    """Photons per device as mmcb_run_multi splits them (reference rule: src/mmc_cu_host.cu:403-429)."""
    out = np.zeros(ndev, dtype=np.uint64)
    w = None if workload is None else np.ascontiguousarray(workload, dtype=np.float32)
    lib().mmcb_photon_shares(int(nphoton), int(ndev), None if w is None else w.ctypes.data, out.ctypes.data)
    return out


def run_multi(cfg, gpuids=None, workload=None):
    """One simulation sharded over several GPUs of this box inside the library (mmcb_run_multi): `gpuids` are 1-based like cfg['gpuid']
    (default: every GPU), `workload` the relative shares (cfg['workload'] of pmmc, `-W` of the command line)."""
    prob = Problem(cfg)
    if gpuids is None:
        gpuids = [g["id"] for g in gpuinfo()]
    dev = np.ascontiguousarray([int(g) - 1 for g in gpuids], dtype=np.int32)
    w = None if workload is None else np.ascontiguousarray(workload, dtype=np.float32)
    if w is not None and len(w) != len(dev):
        raise MMCError(-2, "workload needs one entry per GPU")
    sz = prob.sizes()
    buf = _OutBuffers(prob, sz)
    _check(lib().mmcb_run_multi(C.byref(prob.cfg), C.byref(prob.mesh), len(dev), dev.ctypes.data, None if w is None else w.ctypes.data,
                                C.byref(buf.out)))
    return buf.result(prob)


class Session:
    """Upload once, launch many times, fetch once: the device-resident path used by bench.py and by the
    multi-GPU driver (mmc_b200/multigpu.py)."""

    def __init__(self, cfg):
        self.prob = Problem(cfg)
        self.h = lib().mmcb_create(C.byref(self.prob.cfg), C.byref(self.prob.mesh), self.prob.device)
        if not self.h:
            raise MMCError(-1, lib().mmcb_last_error().decode(errors="replace"))
        self.sz = Sizes()
        _check(lib().mmcb_get_sizes(self.h, C.byref(self.sz)))

    def launch(self, nphoton=None, photon_offset=0, seed=None, seed_offset=0, stream=None):
        n = int(self.prob.cfg.nphoton if nphoton is None else nphoton)
        sd = int(self.prob.cfg.seed if seed is None else seed)
        _check(lib().mmcb_launch(self.h, n, int(photon_offset), sd, int(seed_offset), stream))

    def sync(self):
        _check(lib().mmcb_sync(self.h))
        ms = C.c_float()
        _check(lib().mmcb_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def reset(self):
        _check(lib().mmcb_reset(self.h))

    def devptrs(self):
        d = DevPtrs()
        _check(lib().mmcb_get_devptrs(self.h, C.byref(d)))
        return d

    def tables(self):
        """Diagnostics: (records uint32 [ne, 24], centroids float32 [ne, 4], facenb int32 [ne, 4]) as the session holds them."""
        ne = len(self.prob.elem)
        rec = np.zeros((ne, 24), dtype=np.uint32)
        cent = np.zeros((ne, 4), dtype=np.float32)
        fnb = np.zeros((ne, 4), dtype=np.int32)
        _check(lib().mmcb_get_tables(self.h, rec.ctypes.data, cent.ctypes.data, fnb.ctypes.data))
        return rec, cent, fnb

    def set_field_buffer(self, device_ptr):
        _check(lib().mmcb_set_field_buffer(self.h, C.c_void_p(device_ptr)))

    def fetch(self, energytot=None, energyesc=None):
        buf = _OutBuffers(self.prob, self.sz)
        et = ee = None
        if energytot is not None:
            et = np.zeros(MAX_SRCNUM, dtype=np.float64)
            ee = np.zeros(MAX_SRCNUM, dtype=np.float64)
            et[:len(energytot)] = energytot
            ee[:len(energyesc)] = energyesc
        _check(lib().mmcb_fetch(self.h, None if et is None else et.ctypes.data, None if ee is None else ee.ctypes.data,
                                C.byref(buf.out)))
        return buf.result(self.prob)

    def close(self):
        if self.h:
            lib().mmcb_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
