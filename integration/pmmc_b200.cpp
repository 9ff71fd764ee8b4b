// _pmmc: the Python module of the reference (src/pmmc.cpp:1447-1462 -- run, gpuinfo, version) bound to the mmc_b200 engine.
//
// The reference's own src/pmmc.cpp cannot be linked against a CUDA-only engine as it stands: its USE_CUDA branches call
// mcx_list_cu_gpu with four arguments (src/pmmc.cpp:925,1402) where src/mmc_cu_host.h:58 declares two, so it only builds with
// OpenCL.  This file is the replacement a maintainer would ship: the same module name, the same three functions, the same cfg keys
// (parse_config, src/pmmc.cpp:226-900) and the same output dictionary (src/pmmc.cpp:1085-1340: 'flux', 'fluximag', 'detp', 'seeds',
// 'traj', 'dref', 'jmua' / 'jd' / 'jmus' / 'jmusp' [+ '_re' / '_im' for RF], 'stat'), Fortran-ordered like pmmc returns them --
// written against include/mmc_b200.h alone (no reference header, no torch).  `import pmmc` (integration/pmmc/__init__.py) then runs
// pmmc/example/test_mesh_adjoint.py unchanged (tests/test_pmmc_module.py).
//
// Build (mmc_b200/build.py: build_pmmc):
//   g++ -O2 -shared -fPIC -std=c++17 $(python -m pybind11 --includes) -I include integration/pmmc_b200.cpp
//       -o integration/pmmc/_pmmc$(python3-config --extension-suffix) -L mmc_b200 -lmmc_b200 -Wl,-rpath,'$ORIGIN/../../mmc_b200'
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mmc_b200.h"

namespace py = pybind11;

namespace {

[[noreturn]] void engine_error(int rc) {       // mcx_error -> mmc_throw_exception in the containers (src/mmc_utils.c:1426-1442)
    throw std::runtime_error("MMC ERROR(" + std::to_string(rc) + "):" + mmcb_last_error());
}

std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}

// C-ordered copy of a 2-D (or 1-D) array field
template <typename T>
std::vector<T> rows(const py::dict& cfg, const char* key, int ncol, size_t* nrow, const char* msg) {
    auto a = py::array_t<T, py::array::c_style | py::array::forcecast>::ensure(cfg[key]);

    if (!a) {
        throw py::value_error(std::string("Invalid ") + key + " field value");
    }

    const py::buffer_info b = a.request();
    size_t r = 1, c = 1;

    if (b.ndim == 1) {
        c = (size_t)b.shape[0];
    } else if (b.ndim == 2) {
        r = (size_t)b.shape[0];
        c = (size_t)b.shape[1];
    } else {
        throw py::value_error(msg);
    }

    if (ncol > 0 && c != (size_t)ncol && !(ncol == 4 && b.ndim == 2 && c >= 4)) {
        throw py::value_error(msg);
    }

    *nrow = r;
    const T* p = static_cast<const T*>(b.ptr);
    std::vector<T> out;

    if (ncol > 0 && c > (size_t)ncol) {         // 'elem' with a label column: keep the first ncol
        out.resize(r * ncol);

        for (size_t i = 0; i < r; i++) {
            std::copy(p + i * c, p + i * c + ncol, out.begin() + i * ncol);
        }
    } else {
        out.assign(p, p + r * c);
    }

    return out;
}

void vec4(const py::dict& cfg, const char* key, float* dst) {
    if (!cfg.contains(key)) {
        return;
    }

    auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(cfg[key]);

    if (!a || a.size() < 3 || a.size() > 4) {
        throw py::value_error(std::string("the '") + key + "' field must have 3 or 4 elements");
    }

    const float* p = a.data();

    for (ssize_t i = 0; i < a.size(); i++) {
        dst[i] = p[i];
    }
}

template <typename T>
void scalar(const py::dict& cfg, const char* key, T* dst) {
    if (cfg.contains(key)) {
        *dst = (T)py::float_(cfg[key]).cast<double>();
    }
}

const std::vector<std::string> SRCTYPES = {"pencil", "isotropic", "cone", "gaussian", "planar", "pattern", "fourier", "arcsine", "disk",
                                           "fourierx", "fourierx2d", "zgaussian", "line", "slit"
                                          };
const std::map<std::string, int> METHODS = {{"plucker", 0}, {"p", 0}, {"havel", 1}, {"h", 1}, {"badouel", 2}, {"b", 2}, {"elem", 3}, {"s", 3},
    {"blbadouel", 3}, {"grid", 4}, {"g", 4}
};
const std::map<std::string, int> OUTPUTS = {{"flux", 0}, {"x", 0}, {"fluence", 1}, {"f", 1}, {"energy", 2}, {"e", 2}, {"jacobian", 3}, {"j", 3},
    {"wl", 4}, {"l", 4}, {"wp", 5}, {"p", 5}, {"rf", 6}, {"r", 6}, {"adjoint", 8}, {"a", 8}, {"adjointd", 9}, {"d", 9},
    {"adjointmus", 10}, {"u", 10}, {"adjointmusp", 11}, {"v", 11}, {"adjointmuad", 12}, {"w", 12}, {"adjointmuamusp", 13}, {"q", 13}
};

struct Problem {            // owns every array the C structs point into
    mmcb_config c;
    mmcb_mesh m;
    std::vector<float> node, prop, det, detdir, pattern, evol, nvol, nodemua, nodemusp, srcdata, replayweight, replaytime;
    std::vector<int> elem, type, facenb;
    std::vector<uint64_t> seeds;
    std::vector<int> devices;
    std::vector<float> workload;
};

void parse(const py::dict& cfg, Problem& P) {
    memset(&P.c, 0, sizeof(P.c));
    memset(&P.m, 0, sizeof(P.m));
    mmcb_config& c = P.c;
    // defaults of mcx_initcfg (src/mmc_utils.c:200-330)
    c.seed = 0x623F9A9E;
    c.srcdir[2] = 1.f;
    c.srcnum = 1;
    c.isreflect = 1;
    c.isnormalized = 1;
    c.basisorder = 1;
    c.method = MMCB_RT_BLBADOUEL_GRID;
    c.roulettesize = 10.f;
    c.minenergy = 1e-6f;
    c.nout = 1.f;
    c.voidtime = 1;
    c.unitinmm = 1.f;
    c.steps = 1.f;
    c.maxdetphoton = 1000000;
    c.maxjumpdebug = 10000000;
    c.respin = 1;

    for (const char* k : {"node", "elem", "prop"}) {
        if (!cfg.contains(k)) {
            throw py::value_error(std::string("the '") + k + "' field is required");
        }
    }

    size_t nn = 0, ne = 0, np = 0, n = 0;
    P.node = rows<float>(cfg, "node", 3, &nn, "the 'node' field must have 3 columns (x,y,z)");
    {
        auto a = py::array_t<int, py::array::c_style | py::array::forcecast>::ensure(cfg["elem"]);

        if (!a || a.ndim() != 2 || a.shape(1) < 4) {
            throw py::value_error("the 'elem' field must have 4 or 5 columns");
        }

        ne = (size_t)a.shape(0);
        const size_t w = (size_t)a.shape(1);
        P.elem.resize(4 * ne);
        P.type.assign(ne, 1);

        for (size_t i = 0; i < ne; i++) {
            std::copy(a.data() + i * w, a.data() + i * w + 4, P.elem.begin() + 4 * i);

            if (w >= 5) {
                P.type[i] = a.data()[i * w + 4];
            }
        }
    }

    if (cfg.contains("elemprop")) {
        auto a = py::array_t<int, py::array::c_style | py::array::forcecast>::ensure(cfg["elemprop"]);

        if (!a || (size_t)a.size() != ne) {
            throw py::value_error("the 'elemprop' field must have 1 row or 1 column");
        }

        P.type.assign(a.data(), a.data() + ne);
    }

    P.prop = rows<float>(cfg, "prop", 4, &np, "the 'prop' field must have 4 columns (mua,mus,g,n)");

    if (np < 2) {
        throw py::value_error("the 'prop' field needs the background row and at least one medium");
    }

    P.m.nn = (int)nn;
    P.m.ne = (int)ne;
    P.m.prop = (int)np - 1;
    P.m.node = P.node.data();
    P.m.elem = P.elem.data();
    P.m.type = P.type.data();
    P.m.med = (const mmcb_medium*)P.prop.data();

    if (cfg.contains("facenb")) {
        P.facenb = rows<int>(cfg, "facenb", 4, &n, "the 'facenb' field must have 4 or 10 columns");
        P.m.facenb = P.facenb.data();
    }

    if (cfg.contains("evol")) {
        P.evol = rows<float>(cfg, "evol", 0, &n, "evol");
        P.m.evol = P.evol.data();
    }

    if (cfg.contains("nvol")) {
        P.nvol = rows<float>(cfg, "nvol", 0, &n, "nvol");
        P.m.nvol = P.nvol.data();
    }

    double nphoton = 0;
    scalar(cfg, "nphoton", &nphoton);
    c.nphoton = (uint64_t)nphoton;
    scalar(cfg, "tstart", &c.tstart);
    scalar(cfg, "tstep", &c.tstep);
    scalar(cfg, "tend", &c.tend);

    for (auto kv : std::map<const char*, int*> {{"isreflect", &c.isreflect}, {"isspecular", &c.isspecular}, {"ismomentum", &c.ismomentum},
    {"issaveexit", &c.issaveexit}, {"issavedet", &c.issavedet}, {"issaveseed", &c.issaveseed}, {"basisorder", &c.basisorder},
    {"isnormalized", &c.isnormalized}, {"issaveref", &c.issaveref}, {"voidtime", &c.voidtime}, {"e0", &c.e0}, {"srcid", &c.srcid},
    {"adjointmode", &c.adjointmode}, {"nthread", &c.nthread}, {"nblocksize", &c.nblocksize}, {"respin", &c.respin}, {"srcnum", &c.srcnum}
}) {
        scalar(cfg, kv.first, kv.second);
    }

    scalar(cfg, "roulettesize", &c.roulettesize);
    scalar(cfg, "nout", &c.nout);
    scalar(cfg, "minenergy", &c.minenergy);
    scalar(cfg, "unitinmm", &c.unitinmm);
    scalar(cfg, "omega", &c.omega);
    scalar(cfg, "maxdetphoton", &c.maxdetphoton);
    scalar(cfg, "maxjumpdebug", &c.maxjumpdebug);
    vec4(cfg, "srcpos", c.srcpos);
    vec4(cfg, "srcdir", c.srcdir);
    vec4(cfg, "srcparam1", c.srcparam1);
    vec4(cfg, "srcparam2", c.srcparam2);

    if (cfg.contains("steps")) {
        auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(cfg["steps"]);

        if (!a || a.size() < 1) {
            throw py::value_error("Invalid steps field value");
        }

        if (a.size() >= 3 && !(a.data()[0] == a.data()[1] && a.data()[1] == a.data()[2])) {
            throw py::value_error("MMC dual-grid algorithm currently does not support anisotropic voxels");
        }

        c.steps = a.data()[0];
    }

    if (cfg.contains("srctype")) {
        const std::string s = lower(py::str(cfg["srctype"]));
        const auto it = std::find(SRCTYPES.begin(), SRCTYPES.end(), s);

        if (it == SRCTYPES.end()) {
            throw py::value_error("the specified source type is not supported");
        }

        c.srctype = (int)(it - SRCTYPES.begin());
    }

    if (cfg.contains("method")) {
        const auto it = METHODS.find(lower(py::str(cfg["method"])));

        if (it == METHODS.end()) {
            throw py::value_error("the specified ray-tracing method is not supported");
        }

        c.method = it->second;
    }

    if (cfg.contains("outputtype")) {
        const auto it = OUTPUTS.find(lower(py::str(cfg["outputtype"])));

        if (it == OUTPUTS.end()) {
            throw py::value_error("the specified output type is not supported");
        }

        c.outputtype = it->second;
    }

    if (cfg.contains("debuglevel")) {
        const std::string d = lower(py::str(cfg["debuglevel"]));
        c.savetraj = (d.find('m') != std::string::npos) ? 1 : 0;
    }

    if (cfg.contains("detpos")) {
        P.det = rows<float>(cfg, "detpos", 4, &n, "the 'detpos' field must have 4 columns (x,y,z,radius)");
        c.detnum = (int)n;
        c.detpos = P.det.data();
    }

    if (cfg.contains("detdir")) {
        P.detdir = rows<float>(cfg, "detdir", 4, &n, "the 'detdir' field must have 4 columns (nx,ny,nz,focal length)");

        if ((int)n != c.detnum) {
            throw py::value_error("detdir needs one row per detector");
        }

        c.detdir = P.detdir.data();
    }

    if (cfg.contains("srcpattern")) {
        auto a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(cfg["srcpattern"]);

        if (!a) {
            throw py::value_error("Invalid srcpattern field value");
        }

        P.pattern.assign(a.data(), a.data() + a.size());
        c.srcpattern = P.pattern.data();

        if (a.ndim() == 3 && !cfg.contains("srcnum")) {     // [srcnum, nx, ny] stack of patterns (photon sharing)
            c.srcnum = (int)a.shape(0);
        }
    }

    if (cfg.contains("nodemua") && (!cfg.contains("isnodalmua") || py::int_(cfg["isnodalmua"]).cast<int>())) {
        P.nodemua = rows<float>(cfg, "nodemua", 0, &n, "nodemua");
        c.nodemua = P.nodemua.data();
    }

    if (c.nodemua && cfg.contains("nodemusp") && (!cfg.contains("isnodalmusp") || py::int_(cfg["isnodalmusp"]).cast<int>())) {
        P.nodemusp = rows<float>(cfg, "nodemusp", 0, &n, "nodemusp");
        c.nodemusp = P.nodemusp.data();
    }

    if (cfg.contains("srcdata")) {
        P.srcdata = rows<float>(cfg, "srcdata", 16, &n, "the 'srcdata' field must have 16 columns");
        c.extrasrclen = (int)n;
        c.srcdata = P.srcdata.data();
    }

    // cfg.seed: an integer, or the detected seeds of an earlier run (uint8 [16, n]: replay, src/pmmc.cpp:700-760) with
    // cfg.replayweight / cfg.replaytime next to it
    if (cfg.contains("seed")) {
        py::object s = cfg["seed"];

        if (py::isinstance<py::int_>(s) || py::isinstance<py::float_>(s)) {
            c.seed = (int)py::float_(s).cast<double>();
        } else {
            auto a = py::array_t<uint8_t, py::array::f_style | py::array::forcecast>::ensure(s);

            if (!a || a.ndim() != 2 || a.shape(0) != 16) {
                throw py::value_error("the 'seed' field must be an integer or a uint8 array of 16 rows (one column per photon)");
            }

            const size_t np_ = (size_t)a.shape(1);
            P.seeds.resize(2 * np_);
            memcpy(P.seeds.data(), a.data(), 16 * np_);
            c.seed = MMCB_SEED_FROM_FILE;
            c.nphoton = np_;
            c.photonseed = P.seeds.data();
            P.replayweight = cfg.contains("replayweight") ? rows<float>(cfg, "replayweight", 0, &n, "replayweight") : std::vector<float>(np_, 1.f);
            P.replaytime = cfg.contains("replaytime") ? rows<float>(cfg, "replaytime", 0, &n, "replaytime") : std::vector<float>(np_, 0.f);

            if (P.replayweight.size() != np_ || P.replaytime.size() != np_) {
                throw py::value_error("replayweight / replaytime need one entry per seed");
            }

            c.replayweight = P.replayweight.data();
            c.replaytime = P.replaytime.data();
        }
    }

    // devices: cfg.gpuid is a 1-based index or a '1101'-style mask (src/pmmc.cpp:820-850), cfg.workload the shares
    P.devices.assign(1, 0);

    if (cfg.contains("gpuid")) {
        py::object g = cfg["gpuid"];

        if (py::isinstance<py::str>(g)) {
            const std::string mask = py::str(g);
            P.devices.clear();

            for (size_t i = 0; i < mask.size(); i++) {
                if (mask[i] == '1') {
                    P.devices.push_back((int)i);
                }
            }

            if (P.devices.empty()) {
                throw py::value_error("the 'gpuid' mask enables no device");
            }
        } else {
            P.devices[0] = std::max(1, py::int_(g).cast<int>()) - 1;
        }
    }

    if (cfg.contains("workload")) {
        P.workload = rows<float>(cfg, "workload", 0, &n, "workload");
        P.workload.resize(P.devices.size(), 0.f);
    }
}

template <typename T>
py::array_t<T, py::array::f_style> fortran(const std::vector<size_t>& dims, const T* src) {
    py::array_t<T, py::array::f_style> a(dims);
    memcpy(a.mutable_data(), src, sizeof(T) * (size_t)a.size());
    return a;
}

py::dict run(const py::dict& user_cfg) {
    Problem P;
    parse(user_cfg, P);
    mmcb_sizes sz;
    int rc = mmcb_query_sizes(&P.c, &P.m, &sz);

    if (rc) {
        engine_error(rc);
    }

    const mmcb_config& c = P.c;
    const bool isrf = (c.omega > 0.f && c.seed != MMCB_SEED_FROM_FILE);
    std::vector<double> field(sz.fieldlen, 0.0), field_im(isrf ? sz.fieldlen : 0, 0.0), dref(c.issaveref ? (size_t)sz.nf * sz.maxgate : 0, 0.0);
    std::vector<float> det(c.issavedet ? (size_t)c.maxdetphoton * sz.reclen : 0), traj(c.savetraj ? (size_t)c.maxjumpdebug * 6 : 0), jac(sz.jacoblen);
    std::vector<uint64_t> seeds((c.issavedet && c.issaveseed) ? (size_t)c.maxdetphoton * 2 : 0);
    mmcb_output out;
    memset(&out, 0, sizeof(out));
    out.field = field.data();
    out.field_im = isrf ? field_im.data() : NULL;
    out.dref = dref.empty() ? NULL : dref.data();
    out.detected = det.empty() ? NULL : det.data();
    out.detseed = seeds.empty() ? NULL : seeds.data();
    out.traj = traj.empty() ? NULL : traj.data();
    out.jacob = jac.empty() ? NULL : jac.data();
    out.overwrite = 1;
    {
        py::gil_scoped_release nogil;
        rc = (P.devices.size() > 1) ? mmcb_run_multi(&P.c, &P.m, (int)P.devices.size(), P.devices.data(), P.workload.empty() ? NULL : P.workload.data(), &out)
             : mmcb_run_simulation(&P.c, &P.m, P.devices[0], &out);
    }

    if (rc) {
        engine_error(rc);
    }

    // ---- the output dictionary, src/pmmc.cpp:1085-1340
    py::dict res;
    const size_t datalen = (size_t)sz.datalen, maxgate = (size_t)sz.maxgate, srcnum = (size_t)sz.srcnum, nslots = (size_t)sz.nslots;
    std::vector<size_t> dims;

    if (c.method == MMCB_RT_BLBADOUEL_GRID) {
        const size_t nx = (size_t)sz.dim[0], ny = (size_t)sz.dim[1], nz = (size_t)sz.dim[2];
        dims = (nslots > 1) ? std::vector<size_t> {nx, ny, nz, maxgate, nslots} : (srcnum > 1 ? std::vector<size_t> {srcnum, nx, ny, nz, maxgate} : std::vector<size_t> {nx, ny, nz, maxgate});
    } else {
        dims = (nslots > 1) ? std::vector<size_t> {datalen, maxgate, nslots} : (srcnum > 1 ? std::vector<size_t> {srcnum, datalen, maxgate} : std::vector<size_t> {datalen, maxgate});
    }

    res["flux"] = fortran<double>(dims, field.data());

    if (isrf) {
        std::vector<float> im(field_im.begin(), field_im.end());
        res["fluximag"] = fortran<float>(dims, im.data());
    }

    if (c.issaveref) {
        res["dref"] = fortran<double>({(size_t)sz.nf, maxgate}, dref.data());
    }

    if (c.issavedet && out.detectedcount > 0) {
        res["detp"] = fortran<float>({(size_t)sz.reclen, (size_t)out.detectedcount}, det.data());

        if (c.issaveseed) {
            res["seeds"] = fortran<uint8_t>({16, (size_t)out.detectedcount}, (const uint8_t*)seeds.data());
        }
    }

    if (c.savetraj) {
        res["traj"] = fortran<float>({6, (size_t)out.trajcount}, traj.data());
    }

    if (sz.jacoblen) {
        const size_t pairs = (size_t)sz.adj_ns * sz.adj_nd;
        const bool dual = (c.outputtype >= MMCB_OT_ADJOINT_MUAD);
        std::vector<size_t> jd = (c.method == MMCB_RT_BLBADOUEL_GRID) ? std::vector<size_t> {(size_t)sz.dim[0], (size_t)sz.dim[1], (size_t)sz.dim[2], 1, pairs}
                                 : std::vector<size_t> {datalen, pairs};
        const size_t adjlen = datalen * pairs;
        const char* n1 = (c.outputtype == MMCB_OT_ADJOINT_DCOEFF) ? "jd" : (c.outputtype == MMCB_OT_ADJOINT_MUS) ? "jmus" : (c.outputtype == MMCB_OT_ADJOINT_MUSP) ? "jmusp" : "jmua";
        const char* n2 = (c.outputtype == MMCB_OT_ADJOINT_MUAMUSP) ? "jmusp" : "jd";
        // engine layout per component: [datalen][pairs] with the pair index fastest; pmmc returns [datalen..., pairs] Fortran-ordered
        auto component = [&](const float* src) {
            std::vector<float> t(adjlen);

            for (size_t i = 0; i < datalen; i++)
                for (size_t p = 0; p < pairs; p++) {
                    t[p * datalen + i] = src[i * pairs + p];
                }

            return fortran<float>(jd, t.data());
        };
        // packing: CW [J1] | CW dual [J1, J2] | RF [Re J1, Im J1] | RF dual [Re J1, Re J2, Im J1, Im J2]
        const float* re1 = jac.data(), *re2 = dual ? jac.data() + adjlen : NULL;
        const float* im1 = isrf ? jac.data() + (dual ? 2 : 1) * adjlen : NULL, *im2 = (isrf && dual) ? jac.data() + 3 * adjlen : NULL;

        if (isrf) {
            res[(std::string(n1) + "_re").c_str()] = component(re1);
            res[(std::string(n1) + "_im").c_str()] = component(im1);

            if (dual) {
                res[(std::string(n2) + "_re").c_str()] = component(re2);
                res[(std::string(n2) + "_im").c_str()] = component(im2);
            }
        } else {
            res[n1] = component(re1);

            if (dual) {
                res[n2] = component(re2);
            }
        }
    }

    py::dict stat;
    double etot = 0, eesc = 0;

    for (int j = 0; j < sz.srcnum; j++) {
        etot += out.energytot[j];
        eesc += out.energyesc[j];
    }

    stat["runtime"] = out.kernel_ms;
    stat["nphoton"] = (double)c.nphoton;
    stat["energytot"] = etot;
    stat["energyabs"] = etot - eesc;
    stat["normalizer"] = out.normalizer;
    stat["unitinmm"] = c.unitinmm;
    stat["raytet"] = out.raytet;
    stat["detected"] = out.detectedtotal;
    stat["e0"] = out.e0;
    res["stat"] = stat;
    return res;
}

py::dict run_kwargs(py::kwargs kw) {
    return run(py::dict(kw));
}

py::list gpuinfo() {        // src/pmmc.cpp:1390-1445
    mmcb_gpuinfo info[64];
    const int n = mmcb_list_gpu(info, 64);
    py::list out;

    for (int i = 0; i < std::min(n, 64); i++) {
        py::dict g;
        g["name"] = std::string(info[i].name);
        g["id"] = info[i].id;
        g["devcount"] = info[i].devcount;
        g["major"] = info[i].major;
        g["minor"] = info[i].minor;
        g["globalmem"] = info[i].globalmem;
        g["constmem"] = info[i].constmem;
        g["sharedmem"] = info[i].sharedmem;
        g["regcount"] = info[i].regcount;
        g["clock"] = info[i].clock;
        g["sm"] = info[i].sm;
        g["core"] = info[i].core;
        g["autoblock"] = info[i].autoblock;
        g["autothread"] = info[i].autothread;
        g["maxgate"] = info[i].maxgate;
        out.append(g);
    }

    return out;
}

}   // namespace

PYBIND11_MODULE(_pmmc, m) {
    m.doc() = "PMMC: Python bindings for Mesh-based Monte Carlo, driving the mmc_b200 engine (CUDA, sm_100a)";
    m.def("run", &run, "Runs MMC with the given config.");
    m.def("run", &run_kwargs, "Runs MMC with the given config.");
    m.def("gpuinfo", &gpuinfo, "Prints out the list of CUDA-capable devices attached to this system.");
    m.def("version", []() {
        return std::string("v2025.10 (mmc_b200 ") + std::to_string(mmcb_version()) + ")";
    }, "Prints mmc version information.");
}
