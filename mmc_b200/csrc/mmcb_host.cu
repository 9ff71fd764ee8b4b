// mmc_b200 host layer: the C-ABI of include/mmc_b200.h.
//
// Replaces the reference's per-device CUDA driver (src/mmc_cu_host.cu:204-1528 mmc_run_simulation) and the host-side
// precompute its callers run in mmc_prep (src/mmc_host.c:136-165): volumes, face neighbours, BLB face planes,
// initial-element search, exterior-face numbering, surface nodal-volume correction, normalisation.
// The tables are repacked into the structure-of-records layout of mmcb_types.h before upload.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>           // types only: the library is resolved with dlopen at the first multi-GPU run

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mmc_b200.h"
#include "mmcb_types.h"

// kernel-side launchers (mmcb_kernel.cu)
extern "C" int mmcb_k_upload_param(const mmcb_kparam* hp, const float* det4, int detnum, cudaStream_t st);
extern "C" int mmcb_k_launch_photons(const mmcb_kargs* a, int grid, int block, size_t smem, int method, int isdet, int isgeneral, int isrf, int carveout, int repack, int iscap,
                                     cudaStream_t st);
extern "C" int mmcb_k_max_block(int method, int repack);
extern "C" size_t mmcb_k_rp_smem(int block, int isdet, int devreclen);
extern "C" int mmcb_k_occupancy(int block, size_t smem, int method, int isdet, int isgeneral, int isrf, int repack, int iscap, int* blocks_per_sm);
// mesh pre-processing on the device (mmcb_prep.cu)
extern "C" int mmcb_k_facenb(const int* d_elem, int ne, int* d_facenb, cudaStream_t st);
extern "C" int mmcb_k_build_hpaux(const float* d_node, const int* d_elem, const mmcb_tetrec* d_rec, int ne, float* d_aux, cudaStream_t st);
extern "C" int mmcb_k_build_records(const float* d_node, const int* d_elem, const int* d_facenb, const int* d_type, const float* d_med_n, int ne,
                                    float nout, int isreflect, mmcb_tetrec* d_rec, float4* d_cent, cudaStream_t st);
// mesh_normalize on the device (mmcb_post.cu)
extern "C" int mmcb_k_norm_sum(const double* W, size_t nentry, int srcnum, double* dep, cudaStream_t st);
extern "C" int mmcb_k_norm_nvol(double* W, size_t n, int nn, int srcnum, const float* nvol, cudaStream_t st);
extern "C" int mmcb_k_double_to_float(const double* in, float* out, size_t n, cudaStream_t st);
extern "C" int mmcb_k_norm_elemdep(const double* W, const double* Wim, const int* elem, const float* evol, const float* emua, int ne, int nn, int maxgate, int srcnum,
                                   double* dep, cudaStream_t st);
extern "C" int mmcb_k_norm_scale(const double* in, double* out, size_t n, int datalen, int srcnum, const float* evol, const float* emua,
                                 const double* fac16, cudaStream_t st);
// adjoint-Jacobian post-kernels (mmcb_post.cu)
extern "C" int mmcb_k_adj_cw(const float* field, float* cw, size_t N, int maxgate, int nslots, cudaStream_t st);
extern "C" int mmcb_k_adj_mua(const float* cw_re, const float* cw_im, float* out, size_t N, int Ns, int Nd, float scale, cudaStream_t st);
extern "C" int mmcb_k_adj_dcoeff(const float* cw_re, const float* cw_im, float* out, size_t N, int Ns, int Nd, unsigned int Nx, unsigned int Ny,
                                 float scale, cudaStream_t st);
extern "C" int mmcb_k_adj_mesh_full(const float* cw_re, const float* cw_im, const int* elem, const float* node, const float* evol, float* jmua,
                                    float* jd, int ne, int nn, int Ns, int Nd, cudaStream_t st);
extern "C" int mmcb_k_adj_mesh_nodal(const float* cw_re, const float* cw_im, const float* nvol, float* jmua, int nn, int Ns, int Nd,
                                     cudaStream_t st);
extern "C" int mmcb_k_spread_nodes(const void* efield, double* nfield, const int* elem, int ne, int nn, int maxgate, int srcnum, cudaStream_t st);
extern "C" int mmcb_k_acc_to_double(const void* in, double* out, size_t n, cudaStream_t st);
extern "C" int mmcb_k_acc_is_double(void);
extern "C" int mmcb_k_hot_select(const void* field, size_t fieldlen, unsigned int* stat, void* cand, unsigned int cap, unsigned int* keys,
                                 float minshare, const double* raytet, const unsigned int* alive, float n0, int maxgate, cudaStream_t st);
extern "C" int mmcb_k_rng(const uint32_t* dseeds, int nstream, int ndraw, float* dout, unsigned long long* dstate, cudaStream_t st);

#define MMCB_MAX_DEVICES_RING 16

namespace {

thread_local std::string g_err;
thread_local int g_code = 0;

int fail(int code, const char* fmt, ...) {
    g_code = code;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(MMCB_ERR_CUDA, "CUDA error %d (%s) at %s:%d", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CUK(call) do { int e_ = (call); if (e_ != 0) return fail(MMCB_ERR_CUDA, "CUDA error %d (%s) at %s:%d", e_, cudaGetErrorString((cudaError_t)e_), __FILE__, __LINE__); } while (0)

// MMCB_TRACE=1 prints the wall-clock of every host phase to stderr (the reference prints "init complete / kernel complete /
// transfer complete" lines with StartTimer/GetTimeMillis, src/mmc_cu_host.cu:640,749-753,875)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t;
    Trace() : on(getenv("MMCB_TRACE") != NULL), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (on) {
            auto n = std::chrono::steady_clock::now();
            fprintf(stderr, "[mmcb] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
            t = n;
        }
    }
};

// the hot-line cache pays for its lookups when the hottest 128-byte line draws more than this share of the deposits
// (a flat pilot volume, e.g. a wide disk source on a fine grid, has no line worth privatising; profiles/hotline_r1.md)
const float MMCB_HOT_MINSHARE = 0.005f;
const float EPSF = 1e-6f;
const float VERY_BIG = 1e30f;
// index tables, src/mmc_mesh.c:59-103 and src/mmc_highorder.cpp:50
const int OUT[4][3] = {{0, 3, 1}, {3, 2, 1}, {0, 2, 3}, {0, 1, 2}};
const int FACEMAP[4] = {2, 0, 1, 3};
const int FACEORDER[4] = {1, 3, 2, 0};
const int IFACEORDER[4] = {3, 0, 2, 1};
const int FACELIST[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};

inline const float* nd(const float* node, int id1) {
    return node + 3 * (size_t)(id1 - 1);
}

// ---------------------------------------------------------------------------------------------------
// glibc rand() (TYPE_3 additive feedback generator) restated so that seeds do not depend on libc state
// ---------------------------------------------------------------------------------------------------
struct GlibcRand {
    int32_t r[34];
    uint32_t ring[31];
    int pos;
    explicit GlibcRand(unsigned int seed) {
        if (seed == 0) {
            seed = 1;
        }

        std::vector<uint32_t> v(344);
        int32_t word = (int32_t)seed;
        v[0] = (uint32_t)word;

        for (int i = 1; i < 31; i++) {
            long hi = word / 127773, lo = word % 127773;
            word = (int32_t)(16807 * lo - 2836 * hi);

            if (word < 0) {
                word += 2147483647;
            }

            v[i] = (uint32_t)word;
        }

        for (int i = 31; i < 34; i++) {
            v[i] = v[i - 31];
        }

        for (int i = 34; i < 344; i++) {
            v[i] = v[i - 31] + v[i - 3];
        }

        for (int i = 0; i < 31; i++) {
            ring[i] = v[344 - 31 + i];
        }

        pos = 0;
    }
    uint32_t next() {
        // o_k = o_{k-31} + o_{k-3}
        uint32_t val = ring[pos] + ring[pos >= 3 ? pos - 3 : pos + 28];
        ring[pos] = val;
        pos = (pos == 30) ? 0 : pos + 1;
        return val >> 1;
    }
    // n consecutive outputs; whole turns of the ring run without index arithmetic (604 k words per launch: 2.5 ms -> 0.5 ms)
    void fill(uint32_t* out, size_t n) {
        size_t i = 0;

        for (; i < n && pos != 0; i++) {
            out[i] = next();
        }

        for (; i + 31 <= n; i += 31) {
            for (int k = 0; k < 3; k++) {
                ring[k] += ring[k + 28];
                out[i + k] = ring[k] >> 1;
            }

            for (int k = 3; k < 31; k++) {
                ring[k] += ring[k - 3];
                out[i + k] = ring[k] >> 1;
            }
        }

        for (; i < n; i++) {
            out[i] = next();
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// prepared mesh (host side)
// ---------------------------------------------------------------------------------------------------
struct PrepMesh {
    int nn = 0, ne = 0, nf = 0, prop = 0, isextdet = 0, e0_from_src = 0;
    std::vector<float> node;
    std::vector<int> elem, type, facenb, srcelem, detelem;
    std::vector<mmcb_medium> med;
    std::vector<float> evol, nvol;
    float nmin[3], nmax[3];
};

// ---- element volumes.  The arithmetic of the determinant is the reference's (src/mmc_mesh.c:916-925: the volumes are inputs of the
// normalisation and are compared bit for bit with the reference's tables); everything around it is written for this engine: one pass
// that measures and orients the elements, one pass that hands a quarter of each volume to its nodes.
inline float tet_volume6(const float* a, const float* b, const float* c, const float* d) {
    const float ex = c[0] - d[0], ey = c[1] - d[1], ez = c[2] - d[2];
    const float yz = c[1] * d[2] - c[2] * d[1], xz = c[0] * d[2] - c[2] * d[0], xy = c[0] * d[1] - c[1] * d[0];
    float v = b[0] * yz - b[1] * xz + b[2] * xy;
    v += -a[0] * (yz + b[1] * ez - b[2] * ey);
    v += +a[1] * (xz + b[0] * ez - b[2] * ex);
    v += -a[2] * (xy + b[0] * ey - b[1] * ex);
    return -v;
}

void volumes(int nn, const float* node, int ne, int* elem, const int* type, float* evol, float* nvol) {
    for (int e = 0; e < ne; e++) {          // inverted elements get their last two nodes swapped (mesh_getvolume does the same in place)
        int* q = elem + 4 * (size_t)e;
        float v6 = tet_volume6(nd(node, q[0]), nd(node, q[1]), nd(node, q[2]), nd(node, q[3]));

        if (v6 < 0.f) {
            std::swap(q[2], q[3]);
            v6 = -v6;
        }

        evol[e] = v6 * (1.f / 6.f);
    }

    std::fill(nvol, nvol + nn, 0.f);

    for (int e = 0; e < ne; e++) {          // void elements (label 0) carry no nodal volume
        if (type && type[e] == 0) {
            continue;
        }

        const float quarter = evol[e] * 0.25f;

        for (int k = 0; k < 4; k++) {
            nvol[elem[4 * (size_t)e + k] - 1] += quarter;
        }
    }
}

void facenb_build(int ne, const int* elem, int* facenb) {
    // mesh_getfacenb, src/mmc_highorder.cpp:124-159: match faces through their sorted node triples
    struct Key {
        int a, b, c, slot;
    };
    std::vector<Key> k((size_t)ne * 4);

    for (int i = 0; i < ne; i++) {
        const int* ee = elem + 4 * (size_t)i;

        for (int j = 0; j < 4; j++) {
            int v[3] = {ee[FACELIST[j][0]], ee[FACELIST[j][1]], ee[FACELIST[j][2]]};
            std::sort(v, v + 3);
            k[(size_t)i * 4 + j] = {v[0], v[1], v[2], i * 4 + j};
        }
    }

    std::sort(k.begin(), k.end(), [](const Key & x, const Key & y) {
        if (x.a != y.a) {
            return x.a < y.a;
        }

        if (x.b != y.b) {
            return x.b < y.b;
        }

        if (x.c != y.c) {
            return x.c < y.c;
        }

        return x.slot < y.slot;
    });
    std::fill(facenb, facenb + (size_t)ne * 4, 0);

    for (size_t i = 0; i + 1 < k.size(); i++) {
        if (k[i].a == k[i + 1].a && k[i].b == k[i + 1].b && k[i].c == k[i + 1].c) {
            facenb[k[i].slot] = (k[i + 1].slot >> 2) + 1;
            facenb[k[i + 1].slot] = (k[i].slot >> 2) + 1;
            i++;
        }
    }
}

// ---- point location.  inside_weights: un-normalised barycentric weights of `pt` in element `e` from the four face planes, weight of
// local node FACEMAP[f] = -(pt - A_f) . ((B_f - A_f) x (C_f - A_f)) with the face nodes in the reference's order (src/mmc_mesh.c:59,1180-1190:
// the enclosing element must come out the same for sources that sit on a face).  Returns false when the point is outside.
bool inside_weights(const float* node, const int* quad, const float* pt, float w[4]) {
    bool inside = true;

    for (int f = 0; f < 4; f++) {
        const float* A = nd(node, quad[OUT[f][0]]), *B = nd(node, quad[OUT[f][1]]), *C = nd(node, quad[OUT[f][2]]);
        const float u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, v[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        const float r[3] = {pt[0] - A[0], pt[1] - A[1], pt[2] - A[2]};
        const float n[3] = {u[1]* v[2] - u[2]* v[1], u[2]* v[0] - u[0]* v[2], u[0]* v[1] - u[1]* v[0]};
        w[FACEMAP[f]] = -(r[0] * n[0] + r[1] * n[1] + r[2] * n[2]);
    }

    for (int k = 0; k < 4; k++) {
        inside = inside && !(w[k] < 0.f);
    }

    return inside;
}

// barycentric coordinates of srcpos in element e0 (1-based); non-zero when e0 is invalid or does not enclose the point (mesh_barycentric)
int barycentric(const float* node, const int* elem, int ne, int e0, float* bary, const float* srcpos) {
    if (e0 < 1 || e0 > ne || !inside_weights(node, elem + 4 * (size_t)(e0 - 1), srcpos, bary)) {
        return 1;
    }

    float total = 0.f;

    for (int k = 0; k < 4; k++) {
        total += bary[k];
    }

    for (int k = 0; k < 4; k++) {
        bary[k] /= total;
    }

    return 0;
}

// first element (lowest index, like mesh_initelem's scan) whose bounding box and four planes enclose srcpos; 0 when there is none
int initelem(const float* node, const int* elem, int ne, const float* srcpos, float* bary) {
    for (int e = 0; e < ne; e++) {
        const int* q = elem + 4 * (size_t)e;
        bool inbox = true;

        for (int k = 0; k < 3 && inbox; k++) {
            const float c0 = nd(node, q[0])[k], c1 = nd(node, q[1])[k], c2 = nd(node, q[2])[k], c3 = nd(node, q[3])[k];
            const float lo = std::min(std::min(c0, c1), std::min(c2, c3)), hi = std::max(std::max(c0, c1), std::max(c2, c3));
            inbox = (srcpos[k] >= lo && srcpos[k] <= hi);
        }

        if (inbox && barycentric(node, elem, ne, e + 1, bary, srcpos) == 0) {
            return e + 1;
        }
    }

    return 0;
}

// Effective reflection coefficient of a diffusing medium against the outside (the nodal-volume correction of surface nodes,
// src/mmc_mesh.c:1344-1386): the Fresnel reflectance of unpolarised light, R(t), weighted over the half space,
//     R_phi = int 2 sin t cos t R dt,   R_j = int 3 sin t cos^2 t R dt,   Reff = (R_phi + R_j) / (2 - R_phi + R_j),
// by the rectangle rule on 1000 angles, which is what mesh_getreff (:2305-2334) integrates with -- the same nodes and weights, so
// the corrected volumes agree with the reference's.
double getreff(double n_in, double n_out) {
    const int nangle = 1000;
    const double dt = M_PI / (2.0 * nangle), critical = asin(1.0 / n_in);
    auto fresnel = [&](double t) {
        if (!(t < critical)) {
            return 1.0;
        }

        const double ci = cos(t), si = n_in * sin(t), ct = sqrt(1. - si * si);
        const double rs = (n_in * ct - n_out * ci) / (n_in * ct + n_out * ci), rp = (n_in * ci - n_out * ct) / (n_in * ci + n_out * ct);
        return 0.5 * rs * rs + 0.5 * rp * rp;
    };
    double r_phi = 0.0, r_j = 0.0;

    for (int k = 0; k < nangle; k++) {
        const double t = k * dt, R = fresnel(t);
        r_phi += 2.0 * sin(t) * cos(t) * R;
        r_j += 3.0 * sin(t) * cos(t) * cos(t) * R;
    }

    r_phi *= dt;
    r_j *= dt;
    return (r_phi + r_j) / (2.0 - r_phi + r_j);
}

struct Cfg {              // validated copy of mmcb_config
    mmcb_config c;
    int maxgate = 1;
    int datalen = 0, reclen = 0;
    int dim[3] = {0, 0, 0};
    unsigned int crop0[3] = {0, 0, 0};
    std::vector<float> pattern, detpos;
    float bary0[4] = {0.f, 0.f, 0.f, 0.f};     // cfg->bary0: barycentric coordinates of srcpos in e0
    // multi-slot sources / RF / adjoint output
    std::vector<float> srcdata;                // extrasrclen x 16 (ExtraSrc)
    int nslots = 1, multisrc = 0, isrf = 0;
    int adj_ns = 0, adj_nd = 0, adj_dual = 0;  // adjoint output: source slots, detector slots, two components
    size_t jacoblen = 0;
};

int validate(const mmcb_config* in, const mmcb_mesh* mesh, Cfg& o) {
    // mcx_validatecfg (src/mmc_utils.c:3500-3566) + mcx_prep (:3724-3736)
    if (!in || !mesh) {
        return fail(MMCB_ERR_INPUT, "null config or mesh");
    }

    o.c = *in;
    mmcb_config& c = o.c;

    if (c.nphoton == 0 && c.seed != MMCB_SEED_FROM_FILE) {
        return fail(MMCB_ERR_INPUT, "cfg.nphoton must be a positive number");
    }

    if (c.tstart > c.tend || c.tstep == 0.f) {
        return fail(MMCB_ERR_INPUT, "incorrect time gate settings or missing tstart/tend/tstep fields");
    }

    if (c.tstep > c.tend - c.tstart) {
        c.tstep = c.tend - c.tstart;
    }

    if (fabs(c.srcdir[0] * c.srcdir[0] + c.srcdir[1] * c.srcdir[1] + c.srcdir[2] * c.srcdir[2] - 1.f) > 1e-4) {
        return fail(MMCB_ERR_INPUT, "field 'srcdir' must be a unitary vector (tolerance is 1e-4)");
    }

    if (c.tend <= c.tstart) {
        return fail(MMCB_ERR_INPUT, "field 'tend' must be greater than field 'tstart'");
    }

    o.maxgate = (int)((c.tend - c.tstart) / c.tstep + 0.5);
    c.tend = c.tstart + c.tstep * o.maxgate;

    if (c.srcnum < 1) {
        c.srcnum = 1;
    }

    if (c.srcnum > MMCB_MAX_SRCNUM) {
        return fail(MMCB_ERR_LIMIT, "at most %d simultaneous patterns are supported", MMCB_MAX_SRCNUM);
    }

    if (c.srctype == MMCB_SRC_PATTERN && c.srcpattern == NULL) {
        return fail(MMCB_ERR_INPUT, "the 'srcpattern' field can not be empty when your 'srctype' is 'pattern'");
    }

    if (c.srctype < 0 || c.srctype > MMCB_SRC_SLIT) {
        return fail(MMCB_ERR_INPUT, "source type %d is not supported on the GPU path", c.srctype);
    }

    if (c.srcnum > 1 && c.seed == MMCB_SEED_FROM_FILE) {
        return fail(MMCB_ERR_INPUT, "multiple source simulation is currently not supported under replay mode");
    }

    if (c.seed == MMCB_SEED_FROM_FILE && (!c.photonseed || !c.replayweight || !c.replaytime)) {
        return fail(MMCB_ERR_INPUT, "replay needs photonseed, replayweight and replaytime");
    }

    if (c.method == MMCB_RT_BADOUEL) {
        c.method = MMCB_RT_BLBADOUEL;      // the GPU path offers the branch-less variant (src/mmc_utils.c:3542-3544)
    }

    if (c.method < MMCB_RT_PLUCKER || c.method > MMCB_RT_BLBADOUEL_GRID) {
        return fail(MMCB_ERR_INPUT, "unknown ray tracer %d (p, h, b, s or g)", c.method);
    }

    if ((c.method == MMCB_RT_PLUCKER || c.method == MMCB_RT_HAVEL) && c.srcnum > 1 && c.basisorder) {
        return fail(MMCB_ERR_INPUT, "photon sharing with nodal Havel/Plucker output is not supported; use basisorder 0 or tracer 's'");
    }

    if (c.method == MMCB_RT_BLBADOUEL_GRID) {
        c.basisorder = 0;

        if (!(c.steps > 0.f)) {
            return fail(MMCB_ERR_INPUT, "dual-grid voxel size must be positive");
        }
    }

    if (c.detnum > MMCB_MAX_DET) {
        return fail(MMCB_ERR_LIMIT, "at most %d point detectors are supported", MMCB_MAX_DET);
    }

    if (c.detnum > 0 && !c.detpos) {
        return fail(MMCB_ERR_INPUT, "detnum>0 but detpos is NULL");
    }

    if (c.respin < 1) {
        c.respin = 1;
    }

    if (c.roulettesize <= 0.f) {
        c.roulettesize = 10.f;
    }

    if (c.unitinmm <= 0.f) {
        c.unitinmm = 1.f;
    }

    if (c.srctype == MMCB_SRC_PATTERN) {
        size_t n = (size_t)((int)c.srcparam1[3]) * (size_t)((int)c.srcparam2[3]) * c.srcnum;
        o.pattern.assign(c.srcpattern, c.srcpattern + n);
    }

    if (c.detnum > 0) {
        o.detpos.assign(c.detpos, c.detpos + 4 * (size_t)c.detnum);
    }

    // ---- RF forward and multi-slot sources (mcx_prep, src/mmc_utils.c:3724-3800; mmc_run_simulation, src/mmc_cu_host.cu:229-233,262-264)
    o.isrf = (c.omega > 0.f && c.seed != MMCB_SEED_FROM_FILE);

    if (o.isrf && c.method != MMCB_RT_BLBADOUEL && c.method != MMCB_RT_BLBADOUEL_GRID) {
        return fail(MMCB_ERR_INPUT, "RF (omega > 0) forward runs need the branch-less Badouel tracer (s or g), like the reference GPU path");
    }

    if (o.isrf && c.srcnum > 1) {
        return fail(MMCB_ERR_INPUT, "RF forward runs do not support photon sharing (srcnum > 1)");
    }

    if (c.nodemusp && !c.nodemua) {
        return fail(MMCB_ERR_INPUT, "nodemusp needs nodemua (isnodalmusp is read only under isnodalmua, src/mmc_core.cl:785-793)");
    }

    if (c.nodemua && c.method != MMCB_RT_BLBADOUEL && c.method != MMCB_RT_BLBADOUEL_GRID) {
        return fail(MMCB_ERR_INPUT, "per-node optical properties need the branch-less Badouel tracer (s or g), like the reference GPU path");
    }

    const bool isadjoint = (c.outputtype >= MMCB_OT_ADJOINT);

    if ((c.outputtype == MMCB_OT_JACOBIAN || c.outputtype == MMCB_OT_WL || c.outputtype == MMCB_OT_WP) && c.seed != MMCB_SEED_FROM_FILE) {
        return fail(MMCB_ERR_INPUT, "Jacobian output is only valid in the reply mode. Please give an mch file after '-E'.");    // src/mmc_utils.c:4266-4268
    }

    if (isadjoint && c.seed == MMCB_SEED_FROM_FILE) {
        return fail(MMCB_ERR_INPUT, "Adjoint Jacobian output is not valid in replay mode.");       // src/mmc_utils.c:4270-4272
    }

    if (c.extrasrclen < 0 || (c.extrasrclen > 0 && !c.srcdata)) {
        return fail(MMCB_ERR_INPUT, "extrasrclen > 0 needs srcdata");
    }

    if (c.extrasrclen > 0) {
        o.srcdata.assign(c.srcdata, c.srcdata + 16 * (size_t)c.extrasrclen);
    } else if ((isadjoint || c.srcid == -2) && c.seed != MMCB_SEED_FROM_FILE && c.detnum > 0 && c.detdir) {
        // detectors become reversed (adjoint) sources behind the forward source(s): srcdata[0..Ns-1] forward, [Ns..Ns+Nd-1] detectors
        const int Ns = c.srcnum, Nd = c.detnum;
        c.extrasrclen = Ns + Nd;
        o.srcdata.assign(16 * (size_t)c.extrasrclen, 0.f);

        for (int is = 0; is < Ns; is++) {
            float* q = &o.srcdata[16 * (size_t)is];
            memcpy(q, c.srcpos, 12);
            q[3] = 1.f / Ns;
            memcpy(q + 4, c.srcdir, 12);
            memcpy(q + 8, c.srcparam1, 16);
            memcpy(q + 12, c.srcparam2, 16);
        }

        for (int id = 0; id < Nd; id++) {
            float* q = &o.srcdata[16 * (size_t)(Ns + id)];
            memcpy(q, c.detpos + 4 * (size_t)id, 12);
            q[3] = 1.f / Nd;
            memcpy(q + 4, c.detdir + 4 * (size_t)id, 16);
            q[8] = c.detpos[4 * (size_t)id + 3];      // detector radius: disk source
        }

        if (isadjoint) {
            c.srcid = -1;
        }
    }

    if (c.extrasrclen > 0 && c.srcid > c.extrasrclen) {
        return fail(MMCB_ERR_INPUT, "srcid exceeds total defined source count");
    }

    if (c.extrasrclen > 0 && c.srcid == 0) {
        c.srcid = -1;          // srcdata without a selector: all slots (src/mmc_cu_host.cu:262)
    }

    o.multisrc = (c.extrasrclen > 0 && c.srcid != 0);
    o.nslots = (c.extrasrclen > 0 && c.srcid < 0) ? c.extrasrclen : 1;

    if (o.multisrc && c.srcnum > 1) {
        return fail(MMCB_ERR_INPUT, "multi-slot sources (srcdata) cannot be combined with photon sharing (srcnum > 1)");
    }

    if (o.multisrc && c.seed == MMCB_SEED_FROM_FILE) {
        return fail(MMCB_ERR_INPUT, "multi-slot sources are not supported under replay mode");
    }

    if (isadjoint) {
        if (c.method != MMCB_RT_BLBADOUEL && c.method != MMCB_RT_BLBADOUEL_GRID) {
            return fail(MMCB_ERR_INPUT, "adjoint Jacobian output needs the branch-less Badouel tracer (s or g)");
        }

        if (c.method != MMCB_RT_BLBADOUEL_GRID && !c.basisorder) {      // src/mmc_utils.c:3685-3690
            return fail(MMCB_ERR_INPUT, "mesh-mode adjoint Jacobian requires basisorder=1 (nodal fluence)");
        }

        if (o.nslots < 2 || c.detnum < 1 || !c.detdir || c.extrasrclen <= c.detnum) {
            return fail(MMCB_ERR_INPUT, "adjoint Jacobian output needs detectors with detdir and at least one source slot");
        }

        o.adj_nd = c.detnum;
        o.adj_ns = c.extrasrclen - c.detnum;
        o.adj_dual = (c.outputtype >= MMCB_OT_ADJOINT_MUAD);
    }

    return 0;
}

// gpu_stream: when not NULL (session path, device selected) the face-neighbour table is built on the device (mmcb_prep.cu);
// the host sort remains for mmcb_query_sizes / mmcb_mesh_facenb, which must work without a GPU
// sizes_only: mmcb_query_sizes needs the counts, not the tables -- the face-neighbour sort (the one expensive host step) runs only when
// nf (the exterior-face count behind the diffuse-reflectance length) is asked for
int prepare_mesh(const mmcb_mesh* in, Cfg& cfg, PrepMesh& m, cudaStream_t gpu_stream = NULL, bool use_gpu = false, bool sizes_only = false) {
    mmcb_config& c = cfg.c;

    if (in->nn <= 0 || in->ne <= 0 || !in->node || !in->elem || !in->type || !in->med || in->prop < 1) {
        return fail(MMCB_ERR_MESH, "mesh is missing");
    }

    m.nn = in->nn;
    m.ne = in->ne;
    m.prop = in->prop;
    m.node.assign(in->node, in->node + 3 * (size_t)in->nn);
    m.elem.assign(in->elem, in->elem + 4 * (size_t)in->ne);
    m.type.assign(in->type, in->type + in->ne);

    for (size_t i = 0; i < m.elem.size(); i++)
        if (m.elem[i] < 1 || m.elem[i] > m.nn) {
            return fail(MMCB_ERR_MESH, "element %zu references node %d outside 1..%d", i / 4 + 1, m.elem[i], m.nn);
        }

    // mesh_srcdetelem, src/mmc_mesh.c:390-427
    for (int i = 0; i < m.ne; i++) {
        if (m.type[i] == -1) {
            m.srcelem.push_back(i + 1);

            if (!m.e0_from_src) {
                m.e0_from_src = i + 1;
            }

            m.type[i] = 0;
        } else if (m.type[i] == -2) {
            m.detelem.push_back(i + 1);
            m.isextdet = 1;
        }
    }

    if (m.isextdet) {
        c.detnum = 0;      // wide-field detectors suppress point detectors (:406)
    }

    // mesh_loadmedia, src/mmc_mesh.c:511-547
    m.med.assign(in->med, in->med + in->prop + 1);
    m.med[0] = {0.f, 0.f, 1.f, c.nout};

    if (m.isextdet) {
        m.med.push_back(m.med[0]);

        for (int i = 0; i < m.ne; i++)
            if (m.type[i] == -2) {
                m.type[i] = m.prop + 1;
            }
    }

    for (int i = 0; i < m.ne; i++)
        if (m.type[i] < 0 || m.type[i] > m.prop + m.isextdet) {
            return fail(MMCB_ERR_MESH, "element %d has medium label %d outside 0..%d", i + 1, m.type[i], m.prop);
        }

    if (c.unitinmm != 1.f)
        for (int i = 1; i <= m.prop; i++) {
            m.med[i].mus *= c.unitinmm;
            m.med[i].mua *= c.unitinmm;
        }

    m.evol.resize(m.ne);
    m.nvol.assign(m.nn, 0.f);

    if (in->evol) {
        // mesh_loadelemvol (src/mmc_mesh.c:723-761): volumes given => elements are used as they are (no node swap)
        std::copy(in->evol, in->evol + m.ne, m.evol.begin());

        for (int i = 0; i < m.ne; i++) {
            if (m.type[i] == 0) {
                continue;
            }

            for (int j = 0; j < 4; j++) {
                m.nvol[m.elem[4 * (size_t)i + j] - 1] += m.evol[i] * 0.25f;
            }
        }
    } else {
        volumes(m.nn, m.node.data(), m.ne, m.elem.data(), m.type.data(), m.evol.data(), m.nvol.data());
    }

    if (in->nvol) {
        std::copy(in->nvol, in->nvol + m.nn, m.nvol.begin());
    }

    m.facenb.resize(4 * (size_t)m.ne);

    const bool skip_facenb = sizes_only && !c.issaveref;

    if (skip_facenb) {
        std::fill(m.facenb.begin(), m.facenb.end(), 1);        // placeholder: no exterior faces are counted, nf stays 0
    } else if (in->facenb) {
        for (size_t i = 0; i < m.facenb.size(); i++) {
            m.facenb[i] = in->facenb[i] > 0 ? in->facenb[i] : 0;
        }
    } else if (use_gpu && m.nn < (1 << 21) && !getenv("MMCB_HOST_PREP")) {
        int* d_e = NULL, *d_f = NULL;
        CU(cudaMallocAsync(&d_e, sizeof(int) * m.elem.size(), gpu_stream));
        CU(cudaMallocAsync(&d_f, sizeof(int) * m.elem.size(), gpu_stream));
        CU(cudaMemcpyAsync(d_e, m.elem.data(), sizeof(int) * m.elem.size(), cudaMemcpyHostToDevice, gpu_stream));
        CUK(mmcb_k_facenb(d_e, m.ne, d_f, gpu_stream));
        CU(cudaMemcpyAsync(m.facenb.data(), d_f, sizeof(int) * m.elem.size(), cudaMemcpyDeviceToHost, gpu_stream));
        CU(cudaStreamSynchronize(gpu_stream));
        cudaFreeAsync(d_e, gpu_stream);
        cudaFreeAsync(d_f, gpu_stream);
    } else {
        facenb_build(m.ne, m.elem.data(), m.facenb.data());
    }

    // mcx_prep: detector bookkeeping (src/mmc_utils.c:3729-3736)
    if (c.issavedet && c.detnum == 0 && m.isextdet == 0) {
        c.issavedet = 0;
    }

    if (!c.issavedet) {
        c.ismomentum = 0;
        c.issaveexit = 0;
        c.issaveseed = 0;
    }

    if (c.e0 == 0 && m.e0_from_src) {
        c.e0 = m.e0_from_src;
    }

    // tracer_prep: initial element for point-like sources, src/mmc_mesh.c:1324-1329
    if (c.srctype == MMCB_SRC_PENCIL || c.srctype == MMCB_SRC_ISOTROPIC || c.srctype == MMCB_SRC_CONE || c.srctype == MMCB_SRC_ARCSINE) {
        float* bary = cfg.bary0;

        if (c.e0 <= 0 || barycentric(m.node.data(), m.elem.data(), m.ne, c.e0, bary, c.srcpos)) {
            c.e0 = initelem(m.node.data(), m.elem.data(), m.ne, c.srcpos, bary);

            if (c.e0 == 0) {
                return fail(MMCB_ERR_MESH, "initial element does not enclose the source!");
            }
        }
    } else if (c.e0 <= 0 && m.srcelem.empty()) {
        return fail(MMCB_ERR_MESH, "area sources need an initial element or elements labelled -1");
    }

    if (c.e0 > m.ne) {
        return fail(MMCB_ERR_MESH, "initial element index exceeds total element count");
    }

    // surface nodal volumes x 2/(1+Reff), src/mmc_mesh.c:1344-1386
    if (c.isnormalized == 1 && c.method != MMCB_RT_BLBADOUEL_GRID && c.basisorder) {
        std::vector<float> Reff(m.prop + 2, 0.f);

        if (c.isreflect) {
            for (int i = 1; i <= m.prop; i++) {
                for (int j = 1; j < i; j++)
                    if (m.med[j].n == m.med[i].n) {
                        Reff[i] = Reff[j];
                        break;
                    }

                if (Reff[i] == 0.f) {
                    Reff[i] = (float)getreff(m.med[i].n, m.med[0].n);
                }
            }
        }

        for (int i = 0; i < m.ne; i++) {
            const int* ee = &m.elem[4 * (size_t)i], *enb = &m.facenb[4 * (size_t)i];

            for (int j = 0; j < 4; j++)
                if (enb[j] == 0)
                    for (int k = 0; k < 3; k++) {
                        int nid = ee[OUT[IFACEORDER[j]][k]] - 1;

                        if (m.nvol[nid] > 0.f && m.type[i] >= 0) {
                            m.nvol[nid] *= -(2.f / (1.0 + Reff[m.type[i] > m.prop ? 0 : m.type[i]]));
                        }
                    }
        }

        for (int i = 0; i < m.nn; i++)
            if (m.nvol[i] < 0.f) {
                m.nvol[i] = -m.nvol[i];
            }
    }

    // exterior faces numbered -1..-nf, src/mmc_mesh.c:1466-1474
    m.nf = 0;

    for (size_t i = 0; i < m.facenb.size(); i++)
        if (m.facenb[i] == 0) {
            m.facenb[i] = -(++m.nf);
        }

    // dual grid, src/mmc_mesh.c:349-381
    for (int k = 0; k < 3; k++) {
        m.nmin[k] = VERY_BIG;
        m.nmax[k] = -VERY_BIG;
    }

    for (int i = 0; i < m.nn; i++)
        for (int k = 0; k < 3; k++) {
            m.nmin[k] = std::min(m.nmin[k], m.node[3 * (size_t)i + k]);
            m.nmax[k] = std::max(m.nmax[k], m.node[3 * (size_t)i + k]);
        }

    for (int k = 0; k < 3; k++) {
        m.nmin[k] -= EPSF;
        m.nmax[k] += EPSF;
    }

    if (c.method == MMCB_RT_BLBADOUEL_GRID) {
        for (int k = 0; k < 3; k++) {
            cfg.dim[k] = (int)((m.nmax[k] - m.nmin[k]) / c.steps) + 1;
        }

        cfg.crop0[0] = cfg.dim[0];
        cfg.crop0[1] = cfg.dim[1] * cfg.dim[0];
        cfg.crop0[2] = cfg.dim[1] * cfg.dim[0] * cfg.dim[2];
        cfg.datalen = (int)cfg.crop0[2];
    } else {
        cfg.datalen = c.basisorder ? m.nn : m.ne;
    }

    cfg.reclen = (2 + (c.ismomentum > 0)) * m.prop + (c.issaveexit > 0) * 6 + 2;   // host record length (src/mmc_host.c:248)
    return 0;
}

// BLB face planes (tracer_build, src/mmc_mesh.c:1572-1600) + neighbours/type/flags -> 96-byte records
void build_records(const PrepMesh& m, const Cfg& cfg, std::vector<mmcb_tetrec>& rec, std::vector<float>& cent) {
    const mmcb_config& c = cfg.c;
    rec.resize(m.ne);
    cent.resize(4 * (size_t)m.ne);

    for (int i = 0; i < m.ne; i++) {
        mmcb_tetrec& r = rec[i];
        memset(&r, 0, sizeof(r));
        const int* ee = &m.elem[4 * (size_t)i];
        const float n_here = m.med[m.type[i]].n;

        for (int j = 0; j < 4; j++) {
            const float* a = nd(m.node.data(), ee[OUT[j][0]]), *b = nd(m.node.data(), ee[OUT[j][1]]), *cc = nd(m.node.data(), ee[OUT[j][2]]);
            float AB[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, AC[3] = {cc[0] - a[0], cc[1] - a[1], cc[2] - a[2]};
            float N[3] = {AB[1]* AC[2] - AB[2]* AC[1], AB[2]* AC[0] - AB[0]* AC[2], AB[0]* AC[1] - AB[1]* AC[0]};
            float Rn2 = 1.f / sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
            N[0] *= Rn2;
            N[1] *= Rn2;
            N[2] *= Rn2;
            r.nx[j] = N[0];
            r.ny[j] = N[1];
            r.nz[j] = N[2];
            r.d[j] = N[0] * a[0] + N[1] * a[1] + N[2] * a[2];
            int nb = m.facenb[4 * (size_t)i + FACEORDER[j]];
            r.nb[j] = nb;
            // when does crossing face j call reflectray?  src/mmc_core.cl:1957-1958
            bool refl;

            if (nb <= 0) {
                refl = !((n_here == c.nout && c.isreflect != MMCB_BC_MIRROR) || c.isreflect == MMCB_BC_ABSORB_EXTERIOR);
            } else {
                refl = (m.med[m.type[nb - 1]].n != n_here);
            }

            if (refl) {
                r.flags |= MMCB_F_REFLECT(j);
            }

            if (nb > 0) {
                if (m.type[i] != 0 && m.type[nb - 1] == 0) {
                    r.flags |= MMCB_F_TO_VOID(j);
                }

                if (m.type[i] == 0 && m.type[nb - 1] != 0) {
                    r.flags |= MMCB_F_FROM_VOID(j);
                }
            }
        }

        r.type = m.type[i];
        float cx = 0.f, cy = 0.f, cz = 0.f;

        for (int j = 0; j < 4; j++) {
            const float* q = nd(m.node.data(), ee[j]);
            cx += q[0];
            cy += q[1];
            cz += q[2];
        }

        cent[4 * (size_t)i] = cx * 0.25f;
        cent[4 * (size_t)i + 1] = cy * 0.25f;
        cent[4 * (size_t)i + 2] = cz * 0.25f;
        cent[4 * (size_t)i + 3] = 0.f;
    }
}


// Device memory comes from the stream-ordered pool (cudaMallocAsync): a session is ~25 buffers, and plain cudaFree costs up
// to ~20 ms each next to another allocator's arena (measured: 390 ms to tear a session down inside the bench process).
// The pool keeps the memory for the next session of the process.
thread_local cudaStream_t g_stream = NULL;

int pool_setup(int device) {
    static bool done[64] = {false};

    if (device >= 0 && device < 64 && !done[device]) {
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        done[device] = true;
    }

    return 0;
}

template <typename T>
int dev_alloc(T** dptr, size_t n) {
    CU(cudaMallocAsync((void**)dptr, n * sizeof(T), g_stream));
    return 0;
}

void dev_free(void* p) {
    if (p) {
        cudaFreeAsync(p, g_stream);
    }
}

template <typename T>
int dev_alloc_copy(T** dptr, const T* host, size_t n) {
    *dptr = NULL;

    if (n == 0) {
        return 0;
    }

    CU(cudaMallocAsync((void**)dptr, n * sizeof(T), g_stream));

    if (host) {     // pageable source: the runtime stages the bytes before returning, the vector may die afterwards
        CU(cudaMemcpyAsync(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice, g_stream));
    } else {
        CU(cudaMemsetAsync(*dptr, 0, n * sizeof(T), g_stream));
    }

    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------------------
struct mmcb_session {
    int device = 0;
    Cfg cfg;
    PrepMesh mesh;
    mmcb_kparam kp, kp_pilot;
    size_t smem_scout = 0;         // shared memory of the scout launch (no detector columns, no cache)
    int carveout = -1;             // preferred shared-memory carve-out (percent) of the photon kernel
    mmcb_kargs ka;
    cudaStream_t stream = NULL;
    cudaEvent_t ev0 = NULL, ev1 = NULL;
    float last_ms = 0.f, total_ms = 0.f;
    int grid = 0, block = 128, nthread = 0;
    size_t smem = 0;
    bool isgrid = false, ishp = false, isdet = false, isgeneral = false;
    float lcap_vox = 4.f;
    bool iscap = false;            // dual grid: steps longer than kp.lcap are walked in pieces (kernel variant CAP)
    bool repack = false;           // lane re-packing kernel (mmcb_kernel_rp.cuh): two walkers (RNG streams) per thread
    size_t fieldlen = 0, efieldlen = 0;     // output volume / kernel accumulator volume (differ for nodal BLB)
    bool acc_double = true, field_external = false;
    // device allocations
    mmcb_tetrec* d_tet = NULL;
    float* d_tetaux = NULL;         // Havel / Plucker nodal output: companion records (mmcb_types.h)
    float4* d_cent = NULL;
    float* d_node = NULL;
    int* d_elem = NULL;
    int* d_srcelem = NULL, *d_srccell = NULL, *d_srcitem = NULL;      // candidate elements of wide-field sources and their search grid
    float4* d_med = NULL;
    float* d_pattern = NULL;
    uint32_t* d_seeds = NULL;
    unsigned long long* d_replayseed = NULL;
    float* d_replayweight = NULL, *d_replaytime = NULL;
    void* d_field = NULL;
    void* d_field_im = NULL;       // RF: imaginary part
    float4* d_srcdata = NULL;      // multi-slot sources
    float2* d_eprop = NULL;        // per-element means of the nodal optical properties
    double* d_dref = NULL;
    float* d_detected = NULL;
    unsigned int* d_detcount = NULL;
    unsigned long long* d_detseed = NULL;
    float* d_traj = NULL;
    unsigned int* d_trajcount = NULL;
    double* d_energy = NULL, *d_raytet = NULL;
    unsigned long long* d_counter = NULL;
    // hot-line cache (mmcb_types.h): keys picked once per session from a pilot batch
    unsigned int* d_hotkeys = NULL, *d_hotstat = NULL;
    uint2* d_hotcand = NULL;
    bool hot_allowed = false, hot_ready = false;
    size_t smem_base = 0;
    std::vector<uint32_t> hseeds, hseeds_next;     // this launch's seed words / the next slice, drawn ahead
    bool next_valid = false;
    int next_seed = 0;
    size_t next_start = 0;
    // host seed stream position: consecutive slices (ranks/respins/bench steps) continue instead of replaying rand() from 0
    GlibcRand* seedgen = NULL;
    int seedgen_seed = 0;
    size_t seedgen_pos = 0;
    uint64_t launched = 0;
};

// kernel variant bits of mmcb_k_launch_photons / mmcb_k_occupancy: 1 = dual grid with the capped segment loop, 2 = Havel / Plucker
// with the nodal deposit inside the kernel
static inline int kvariant(const mmcb_session* s) {
    return (s->iscap ? 1 : 0) | ((s->ishp && s->cfg.c.basisorder) ? 2 : 0);
}

static int session_free(mmcb_session* s) {
    if (!s) {
        return 0;
    }

    cudaSetDevice(s->device);
    g_stream = s->stream;
    dev_free(s->d_tet);
    dev_free(s->d_tetaux);
    dev_free(s->d_cent);
    dev_free(s->d_node);
    dev_free(s->d_elem);
    dev_free(s->d_srcelem);
    dev_free(s->d_srccell);
    dev_free(s->d_srcitem);
    dev_free(s->d_med);
    dev_free(s->d_pattern);
    dev_free(s->d_seeds);
    dev_free(s->d_replayseed);
    dev_free(s->d_replayweight);
    dev_free(s->d_replaytime);

    if (!s->field_external) {
        dev_free(s->d_field);
    }

    dev_free(s->d_field_im);
    dev_free(s->d_srcdata);
    dev_free(s->d_eprop);
    dev_free(s->d_dref);
    dev_free(s->d_detected);
    dev_free(s->d_detcount);
    dev_free(s->d_detseed);
    dev_free(s->d_traj);
    dev_free(s->d_trajcount);
    dev_free(s->d_energy);
    dev_free(s->d_raytet);
    dev_free(s->d_counter);
    dev_free(s->d_hotkeys);
    dev_free(s->d_hotstat);
    dev_free(s->d_hotcand);
    delete s->seedgen;

    if (s->stream) {
        cudaStreamSynchronize(s->stream);
    }

    if (s->ev0) {
        cudaEventDestroy(s->ev0);
    }

    if (s->ev1) {
        cudaEventDestroy(s->ev1);
    }

    if (s->stream) {
        cudaStreamDestroy(s->stream);
    }

    delete s;
    return 0;
}

static int session_build(mmcb_session* s, const mmcb_config* cfgin, const mmcb_mesh* meshin, int device) {
    Trace tr;
    int rc = validate(cfgin, meshin, s->cfg);

    if (rc) {
        return rc;
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);

    if (e != cudaSuccess || ndev == 0) {
        return fail(MMCB_ERR_CUDA, "no CUDA device is available (%s); mmc_b200 has no CPU fallback", cudaGetErrorString(e));
    }

    if (device < 0 || device >= ndev) {
        return fail(MMCB_ERR_CUDA, "GPU ID must be within 0..%d", ndev - 1);
    }

    s->device = device;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    g_stream = s->stream;

    if (pool_setup(device)) {
        return g_code;
    }

    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    tr.mark("device+stream");
    rc = prepare_mesh(meshin, s->cfg, s->mesh, s->stream, true);

    if (rc) {
        return rc;
    }

    tr.mark("validate+prepare_mesh");
    const mmcb_config& c = s->cfg.c;
    const PrepMesh& m = s->mesh;
    s->acc_double = mmcb_k_acc_is_double() != 0;
    s->isgrid = (c.method == MMCB_RT_BLBADOUEL_GRID);
    s->ishp = (c.method == MMCB_RT_PLUCKER || c.method == MMCB_RT_HAVEL);
    s->isdet = c.issavedet != 0;
    s->isgeneral = !(c.srctype == MMCB_SRC_PENCIL || c.srctype == MMCB_SRC_ISOTROPIC) || c.srcnum > 1 ||
                   c.seed == MMCB_SEED_FROM_FILE || c.savetraj || c.issaveref || s->cfg.multisrc || s->cfg.isrf || c.nodemua != NULL;
    // Dual grid: a warp runs the segment loop of a step as long as its longest lane needs.  Where steps span many voxels (skinvessel: 5 um
    // voxels, 60 um elements: 7.7 of 32 lanes active in the loop) a lane deposits at most 2 * lcap_vox segments per iteration and carries
    // the rest of the step over (kernel variant CAP, mmcb_kernel.cu).  Where they do not (sphshells, cube60: the typical step is one or
    // two voxels) the variant would only cost its bookkeeping (+3 % measured), so it is chosen from the typical step length in voxels:
    // min(mean element edge, longest mean free path) / voxel >= 4.  MMCB_SEGCAP=<voxels> forces it, 0 disables
    // (profiles/r2c_segcap.jsonl).
    s->lcap_vox = 4.f;

    if (s->isgrid && !s->cfg.isrf) {
        double vol = 0, mfp = 0;

        for (int i = 0; i < m.ne; i++) {
            vol += m.evol[i];
        }

        for (int i = 1; i <= m.prop; i++) {
            mfp = std::max(mfp, m.med[i].mus > 1e-6f ? 1.0 / m.med[i].mus : 1e30);
        }

        const double typical = std::min(std::cbrt(vol / std::max(m.ne, 1)), mfp) / c.steps;
        s->iscap = typical >= 4.0;

        if (const char* e = getenv("MMCB_SEGCAP")) {
            s->iscap = atof(e) > 0.0;
            s->lcap_vox = s->iscap ? (float)atof(e) : 4.f;
        }
    }

    // EXPERIMENTAL lane re-packing kernel (mmcb_kernel_rp.cuh), opt-in with MMCB_REPACK=1: measured SLOWER than the flattened kernel on
    // every workload (profiles/r2a_repack_*), kept for the record and for its parity test.  It serves every single-pattern source;
    // photon sharing, replay, trajectories, diffuse reflectance, multi-slot sources, RF and per-node optical properties change the
    // step itself and stay on the flattened kernel, as do Havel / Plucker and the static (reference) schedule
    s->repack = !s->ishp && c.schedule != 1 && getenv("MMCB_REPACK") && atoi(getenv("MMCB_REPACK")) > 0 &&
                !(c.srcnum > 1 || c.seed == MMCB_SEED_FROM_FILE || c.savetraj || c.issaveref || s->cfg.multisrc || s->cfg.isrf || c.nodemua != NULL);

    // per-slot launch element (mesh_init_srcdata_eid, src/mmc_mesh.c:1114-1156): srcparam2.w of every slot that has none
    for (int slot = 0; slot < c.extrasrclen; slot++) {
        float* q = &s->cfg.srcdata[16 * (size_t)slot];

        if (!(q[15] > 0.f)) {
            float bary[4];
            const int eid = initelem(m.node.data(), m.elem.data(), m.ne, q, bary);

            if (eid <= 0 && c.srcid <= 0) {
                return fail(MMCB_ERR_MESH, "source slot %d at (%g, %g, %g) is not enclosed by any element", slot + 1, q[0], q[1], q[2]);
            }

            q[15] = (float)std::max(eid, 0);
        }
    }

    // tables
    if ((rc = dev_alloc_copy(&s->d_node, m.node.data(), m.node.size()))) {
        return rc;
    }

    if ((rc = dev_alloc_copy(&s->d_elem, m.elem.data(), m.elem.size()))) {
        return rc;
    }

    if (!getenv("MMCB_HOST_PREP")) {
        // branch-less Badouel records and centroids are built on the device (mmcb_prep.cu): 20 bytes per element go up
        // (neighbours + label) instead of 112 (records + centroids)
        int* d_fnb = NULL, *d_type = NULL;
        float* d_medn = NULL;
        std::vector<float> medn(m.med.size());

        for (size_t i = 0; i < m.med.size(); i++) {
            medn[i] = m.med[i].n;
        }

        if ((rc = dev_alloc_copy(&d_fnb, m.facenb.data(), m.facenb.size())) || (rc = dev_alloc_copy(&d_type, m.type.data(), m.type.size())) ||
                (rc = dev_alloc_copy(&d_medn, medn.data(), medn.size()))) {
            return rc;
        }

        CU(cudaMallocAsync(&s->d_tet, sizeof(mmcb_tetrec) * (size_t)m.ne, s->stream));
        CU(cudaMallocAsync(&s->d_cent, sizeof(float4) * (size_t)m.ne, s->stream));
        CUK(mmcb_k_build_records(s->d_node, s->d_elem, d_fnb, d_type, d_medn, m.ne, c.nout, c.isreflect, s->d_tet, s->d_cent, s->stream));
        dev_free(d_fnb);
        dev_free(d_type);
        dev_free(d_medn);
        tr.mark("build_records (device)");
    } else {
        std::vector<mmcb_tetrec> rec;
        std::vector<float> cent;
        build_records(m, s->cfg, rec, cent);
        tr.mark("build_records");
        rc = dev_alloc_copy(&s->d_tet, rec.data(), rec.size());

        if (rc) {
            return rc;
        }

        if ((rc = dev_alloc_copy((float**)&s->d_cent, cent.data(), cent.size()))) {
            return rc;
        }
    }

    if (s->ishp && c.basisorder) {      // Havel / Plucker nodal deposit: reciprocal node heights + opposite node ids per face
        CU(cudaMallocAsync(&s->d_tetaux, sizeof(float) * MMCB_HPAUX_FLOATS * (size_t)m.ne, s->stream));
        CUK(mmcb_k_build_hpaux(s->d_node, s->d_elem, s->d_tet, m.ne, s->d_tetaux, s->stream));
    }

    if ((rc = dev_alloc_copy(&s->d_srcelem, m.srcelem.data(), m.srcelem.size()))) {
        return rc;
    }

    // search grid over the candidate elements (see find_launch_elem): about two candidates per cell on average, every candidate entered
    // in all cells its bounding box (grown by 1e-3, far more than the -1e-4 slack of the enclosure test) touches, in list order
    int gdim[3] = {0, 0, 0};
    float glo[3] = {0.f, 0.f, 0.f}, ginv[3] = {0.f, 0.f, 0.f};

    if (m.srcelem.size() > 16 && !getenv("MMCB_NO_SRCGRID")) {
        const float grow = 1e-3f;
        float hi[3] = { -VERY_BIG, -VERY_BIG, -VERY_BIG};
        glo[0] = glo[1] = glo[2] = VERY_BIG;
        std::vector<float> box(6 * m.srcelem.size());

        for (size_t i = 0; i < m.srcelem.size(); i++) {
            const int* q = &m.elem[4 * (size_t)(m.srcelem[i] - 1)];

            for (int k = 0; k < 3; k++) {
                const float c0 = nd(m.node.data(), q[0])[k], c1 = nd(m.node.data(), q[1])[k], c2 = nd(m.node.data(), q[2])[k], c3 = nd(m.node.data(), q[3])[k];
                box[6 * i + k] = std::min(std::min(c0, c1), std::min(c2, c3)) - grow;
                box[6 * i + 3 + k] = std::max(std::max(c0, c1), std::max(c2, c3)) + grow;
                glo[k] = std::min(glo[k], box[6 * i + k]);
                hi[k] = std::max(hi[k], box[6 * i + 3 + k]);
            }
        }

        const double ext[3] = {std::max(hi[0] - glo[0], 1e-6f), std::max(hi[1] - glo[1], 1e-6f), std::max(hi[2] - glo[2], 1e-6f)};
        const double h = std::cbrt(ext[0] * ext[1] * ext[2] / (0.5 * m.srcelem.size()));

        for (int k = 0; k < 3; k++) {
            gdim[k] = std::min(256, std::max(1, (int)std::ceil(ext[k] / h)));
            ginv[k] = (float)(gdim[k] / ext[k]);
        }

        const size_t ncell = (size_t)gdim[0] * gdim[1] * gdim[2];
        std::vector<int> count(ncell + 1, 0), item;
        auto cells = [&](size_t i, int k, int& c0, int& c1) {
            c0 = std::min(std::max((int)((box[6 * i + k] - glo[k]) * ginv[k]), 0), gdim[k] - 1);
            c1 = std::min(std::max((int)((box[6 * i + 3 + k] - glo[k]) * ginv[k]), 0), gdim[k] - 1);
        };

        for (int pass = 0; pass < 2; pass++) {      // count, then fill (candidates in list order within every cell)
            for (size_t i = 0; i < m.srcelem.size(); i++) {
                int x0, x1, y0, y1, z0, z1;
                cells(i, 0, x0, x1);
                cells(i, 1, y0, y1);
                cells(i, 2, z0, z1);

                for (int z = z0; z <= z1; z++)
                    for (int y = y0; y <= y1; y++)
                        for (int x = x0; x <= x1; x++) {
                            const size_t c = ((size_t)z * gdim[1] + y) * gdim[0] + x;

                            if (pass == 0) {
                                count[c + 1]++;
                            } else {
                                item[count[c]++] = m.srcelem[i];
                            }
                        }
            }

            if (pass == 0) {
                for (size_t c = 0; c < ncell; c++) {
                    count[c + 1] += count[c];
                }

                item.resize(count[ncell]);
            } else {                                // the fill advanced every start to its end: shift back
                for (size_t c = ncell; c > 0; c--) {
                    count[c] = count[c - 1];
                }

                count[0] = 0;
            }
        }

        if ((rc = dev_alloc_copy(&s->d_srccell, count.data(), count.size())) || (rc = dev_alloc_copy(&s->d_srcitem, item.data(), item.size()))) {
            return rc;
        }
    }

    {
        // device media table: two float4 per medium -- {mua, mus, g, n} and the per-step derived constants
        // {1/mus (0: no scattering), n/c0, 1/mua (0: mua < EPS), c0/n} so that the step issues no divisions for them
        std::vector<float4> dm(2 * m.med.size());

        for (size_t i = 0; i < m.med.size(); i++) {
            const mmcb_medium& q = m.med[i];
            const float rc = q.n * 3.335640951981520e-12f;
            dm[2 * i] = make_float4(q.mua, q.mus, q.g, q.n);
            dm[2 * i + 1] = make_float4(q.mus <= 1e-6f ? 0.f : 1.f / q.mus, rc, q.mua < 1e-6f ? 0.f : 1.f / q.mua, 1.f / rc);
        }

        if ((rc = dev_alloc_copy(&s->d_med, dm.data(), dm.size()))) {
            return rc;
        }
    }

    if ((rc = dev_alloc_copy(&s->d_pattern, s->cfg.pattern.data(), s->cfg.pattern.size()))) {
        return rc;
    }

    // accumulators
    const int srcnum = c.srcnum;
    const int nslots = s->cfg.nslots;
    s->fieldlen = (size_t)s->cfg.datalen * s->cfg.maxgate * srcnum * nslots;
    // BLB deposits per element (nodal output is spread on fetch); Havel/Plucker deposit straight into nodes for basisorder=1
    size_t framelen = s->isgrid ? (size_t)s->cfg.crop0[2] : ((s->ishp && c.basisorder) ? (size_t)m.nn : (size_t)m.ne);
    s->efieldlen = framelen * s->cfg.maxgate * srcnum * nslots;

    if (s->efieldlen >= 0xFFFFFFFFull) {
        return fail(MMCB_ERR_LIMIT, "output volume of %zu entries exceeds the 32-bit index range of the kernel", s->efieldlen);
    }

    CU(cudaMallocAsync(&s->d_field, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));
    CU(cudaMemsetAsync(s->d_field, 0, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));

    if (s->cfg.isrf) {
        CU(cudaMallocAsync(&s->d_field_im, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));
        CU(cudaMemsetAsync(s->d_field_im, 0, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));
    }

    if (c.extrasrclen > 0 && (rc = dev_alloc_copy((float**)&s->d_srcdata, s->cfg.srcdata.data(), s->cfg.srcdata.size()))) {
        return rc;
    }

    if (c.nodemua) {        // 0.25 * (sum of the four nodal values), in the reference's order of additions
        std::vector<float2> ep(m.ne);

        for (int i = 0; i < m.ne; i++) {
            const int* ee = &m.elem[4 * (size_t)i];
            ep[i].x = 0.25f * (c.nodemua[ee[0] - 1] + c.nodemua[ee[1] - 1] + c.nodemua[ee[2] - 1] + c.nodemua[ee[3] - 1]);     // used as given: the reference does not scale them by unitinmm
            ep[i].y = c.nodemusp ? 0.25f * (c.nodemusp[ee[0] - 1] + c.nodemusp[ee[1] - 1] + c.nodemusp[ee[2] - 1] + c.nodemusp[ee[3] - 1]) : 0.f;
        }

        if ((rc = dev_alloc_copy(&s->d_eprop, ep.data(), ep.size()))) {
            return rc;
        }
    }

    if (c.issaveref) {
        if ((rc = dev_alloc_copy(&s->d_dref, (const double*)NULL, (size_t)m.nf * s->cfg.maxgate))) {
            return rc;
        }
    }

    const int devreclen = s->cfg.reclen - 1;      // kernel-side record (without detid)

    if (s->isdet) {
        if ((rc = dev_alloc_copy(&s->d_detected, (const float*)NULL, (size_t)c.maxdetphoton * s->cfg.reclen))) {
            return rc;
        }

        if (c.issaveseed && (rc = dev_alloc_copy(&s->d_detseed, (const unsigned long long*)NULL, (size_t)c.maxdetphoton * 2))) {
            return rc;
        }
    }

    if ((rc = dev_alloc_copy(&s->d_detcount, (const unsigned int*)NULL, 1))) {
        return rc;
    }

    if (c.savetraj) {
        if ((rc = dev_alloc_copy(&s->d_traj, (const float*)NULL, (size_t)c.maxjumpdebug * MMCB_DEBUG_REC))) {
            return rc;
        }
    }

    if ((rc = dev_alloc_copy(&s->d_trajcount, (const unsigned int*)NULL, 1))) {
        return rc;
    }

    if ((rc = dev_alloc_copy(&s->d_energy, (const double*)NULL, 2 * MMCB_MAX_SRCNUM))) {
        return rc;
    }

    if ((rc = dev_alloc_copy(&s->d_raytet, (const double*)NULL, 1))) {
        return rc;
    }

    if ((rc = dev_alloc_copy(&s->d_counter, (const unsigned long long*)NULL, 1))) {
        return rc;
    }

    if (c.seed == MMCB_SEED_FROM_FILE) {
        if ((rc = dev_alloc_copy(&s->d_replayseed, (const unsigned long long*)c.photonseed, (size_t)c.nphoton * 2))) {
            return rc;
        }

        if ((rc = dev_alloc_copy(&s->d_replayweight, c.replayweight, (size_t)c.nphoton))) {
            return rc;
        }

        if ((rc = dev_alloc_copy(&s->d_replaytime, c.replaytime, (size_t)c.nphoton))) {
            return rc;
        }
    }

    tr.mark("alloc+upload");
    // launch shape: persistent grid = resident CTAs per SM x SM count (the reference sizes to 64 thr x 32 x #SM, src/mmc_cu_host.cu:159-163)
    int smcount = 0, smemoptin = 0;       // two attributes instead of cudaGetDeviceProperties (several ms per call)
    CU(cudaDeviceGetAttribute(&smcount, cudaDevAttrMultiProcessorCount, device));
    CU(cudaDeviceGetAttribute(&smemoptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    s->block = (c.nblocksize > 0) ? std::min(c.nblocksize, mmcb_k_max_block(c.method, s->repack)) : mmcb_k_max_block(c.method, s->repack);
    s->block = std::max(32, (s->block / 32) * 32);
    s->smem_base = 2 * sizeof(float4) * m.med.size() + (s->isdet ? sizeof(float) * (size_t)devreclen * s->block : 0);
    s->hot_allowed = (c.hotcache >= 0 && srcnum == 1);
    s->smem_scout = 2 * sizeof(float4) * m.med.size();

    if (s->repack) {
        s->smem_base = 2 * sizeof(float4) * m.med.size() + mmcb_k_rp_smem(s->block, s->isdet, devreclen);
        s->smem_scout = 2 * sizeof(float4) * m.med.size() + mmcb_k_rp_smem(s->block, 0, 0);
    }
    // the grid (= number of RNG streams) is sized for the larger footprint so that pilot and main launch share it
    s->smem = s->smem_base + (s->hot_allowed ? sizeof(unsigned int) * MMCB_HOT_SLOTS + sizeof(float) * MMCB_HOT_SLOTS * MMCB_HOT_GROUP : 0);

    if (s->hot_allowed) {
        CU(cudaMallocAsync(&s->d_hotkeys, sizeof(unsigned int) * MMCB_HOT_SLOTS, s->stream));
        CU(cudaMemsetAsync(s->d_hotkeys, 0xFF, sizeof(unsigned int) * MMCB_HOT_SLOTS, s->stream));
        CU(cudaMallocAsync(&s->d_hotstat, sizeof(unsigned int) * MMCB_HOT_STAT_WORDS, s->stream));
        CU(cudaMemsetAsync(s->d_hotstat, 0, sizeof(unsigned int) * MMCB_HOT_STAT_WORDS, s->stream));
        CU(cudaMallocAsync(&s->d_hotcand, sizeof(uint2) * 2 * MMCB_HOT_SLOTS, s->stream));
    }

    if (s->smem > (size_t)smemoptin) {
        return fail(MMCB_ERR_LIMIT, "media table and detector records need %zu bytes of shared memory, device offers %zu", s->smem, (size_t)smemoptin);
    }

    int bps = 0;
    CUK(mmcb_k_occupancy(s->block, s->smem, c.method, s->isdet, s->isgeneral, s->cfg.isrf, s->repack, kvariant(s), &bps));

    if (bps < 1) {
        return fail(MMCB_ERR_CUDA, "kernel cannot be resident with block=%d smem=%zu", s->block, s->smem);
    }

    // resident CTAs x (dynamic + 1 KB reserved shared memory) of the 228 KB an SM can carve out
    s->carveout = (int)std::min<size_t>(100, ((size_t)bps * (s->smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));

    if (c.nthread > 0) {
        s->grid = std::max(1, c.nthread / s->block);
    } else {
        s->grid = bps * smcount;
    }

    s->nthread = s->grid * s->block * (s->repack ? 2 : 1);      // RNG streams: one per thread, two walkers per thread when re-packing
    CU(cudaMallocAsync(&s->d_seeds, sizeof(uint32_t) * 4 * (size_t)s->nthread, s->stream));
    // kernel parameters
    mmcb_kparam& k = s->kp;
    memset(&k, 0, sizeof(k));
    memcpy(k.srcpos, c.srcpos, sizeof(k.srcpos));
    memcpy(k.srcdir, c.srcdir, sizeof(k.srcdir));
    memcpy(k.srcparam1, c.srcparam1, sizeof(k.srcparam1));
    memcpy(k.srcparam2, c.srcparam2, sizeof(k.srcparam2));
    k.srctype = c.srctype;
    k.srcnum = srcnum;
    k.srcelemlen = (int)m.srcelem.size();
    k.e0 = c.e0;
    memcpy(k.bary0, s->cfg.bary0, sizeof(k.bary0));
    k.focus = c.srcdir[3];
    k.tstart = c.tstart;
    k.tend = c.tend;
    k.Rtstep = 1.f / c.tstep;
    k.maxgate = s->cfg.maxgate;
    k.isreflect = c.isreflect;
    k.isspecular = c.isspecular;
    k.voidtime = c.voidtime;
    k.isextdet = m.isextdet;
    k.outputtype = c.outputtype;
    k.method = c.method;
    k.basisorder = c.basisorder;
    k.minenergy = c.minenergy;
    k.roulettesize = c.roulettesize;
    k.nout = c.nout;
    k.doroulette = ((c.tend - c.tstart) * k.Rtstep <= 1.f);
    k.roulette_w = (k.doroulette && c.minenergy > 0.f) ? c.minenergy : -1.f;
    k.nn = m.nn;
    k.ne = m.ne;
    k.nf = m.nf;
    k.maxmedia = m.prop;
    k.framelen = (unsigned int)framelen;
    memcpy(k.nmin, m.nmin, sizeof(k.nmin));
    k.dstep = s->isgrid ? 1.f / c.steps : 1.f;
    k.segcap = s->iscap ? std::max(2, (int)(2.f * s->lcap_vox)) : 0x7FFFFFFF;
    memcpy(k.crop0, s->cfg.crop0, sizeof(k.crop0));
    k.issavedet = c.issavedet;
    k.ismomentum = c.ismomentum;
    k.issaveexit = c.issaveexit;
    k.issaveseed = c.issaveseed;
    k.issaveref = c.issaveref;
    k.detnum = c.detnum;
    k.reclen = devreclen;
    k.maxdetphoton = c.maxdetphoton;
    k.isreplay = (c.seed == MMCB_SEED_FROM_FILE);
    k.savetraj = c.savetraj;
    k.maxjumpdebug = c.maxjumpdebug;
    k.schedule = c.schedule;
    k.nmedia = (int)m.med.size();
    k.fieldlen = (unsigned int)s->efieldlen;
    k.multisrc = s->cfg.multisrc;
    k.srcid = c.srcid;
    k.extrasrclen = c.extrasrclen;
    k.slotstride = (unsigned int)(framelen * s->cfg.maxgate);
    k.omega = s->cfg.isrf ? c.omega : 0.f;
    k.isnodalprop = c.nodemua ? (c.nodemusp ? 2 : 1) : 0;
    memcpy(k.srcgrid_lo, glo, sizeof(glo));
    memcpy(k.srcgrid_inv, ginv, sizeof(ginv));
    memcpy(k.srcgrid_dim, gdim, sizeof(gdim));
    mmcb_kargs& a = s->ka;
    memset(&a, 0, sizeof(a));
    a.tet = s->d_tet;
    a.tetaux = s->d_tetaux;
    a.cent = s->d_cent;
    a.node = s->d_node;
    a.elem = s->d_elem;
    a.srcelem = s->d_srcelem;
    a.srccell = s->d_srccell;
    a.srcitem = s->d_srcitem;
    a.med = s->d_med;
    a.srcpattern = s->d_pattern;
    a.seeds = s->d_seeds;
    a.hotkeys = s->d_hotkeys;
    a.hotstat = s->d_hotstat;
    a.replayseed = s->d_replayseed;
    a.replayweight = s->d_replayweight;
    a.replaytime = s->d_replaytime;
    a.field = s->d_field;
    a.field_im = s->d_field_im;
    a.srcdata = s->d_srcdata;
    a.eprop = s->d_eprop;
    a.dref = s->d_dref;
    a.detected = s->d_detected;
    a.detcount = s->d_detcount;
    a.detseed = s->d_detseed;
    a.traj = s->d_traj;
    a.trajcount = s->d_trajcount;
    a.energy = s->d_energy;
    a.raytet = s->d_raytet;
    a.photon_counter = s->d_counter;
    CU(cudaStreamSynchronize(s->stream));      // uploads done: launches may come on any stream
    tr.mark("occupancy+params");
    return 0;
}

extern "C" {

int mmcb_version(void) {
    return MMCB_VERSION;
}

const char* mmcb_last_error(void) {
    return g_err.c_str();
}

void mmcb_host_seeds(int seed, size_t skip, size_t count, uint32_t* out) {
    GlibcRand g((unsigned int)seed);

    for (size_t i = 0; i < skip; i++) {
        g.next();
    }

    g.fill(out, count);
}

int mmcb_rng_selftest(const uint32_t* seeds4, int nstream, int ndraw, float* out, uint64_t* state_out) {
    if (!seeds4 || !out || nstream <= 0 || ndraw <= 0) {
        return fail(MMCB_ERR_INPUT, "bad argument");
    }

    uint32_t* ds = NULL;
    float* dout = NULL;
    unsigned long long* dst = NULL;
    CU(cudaMalloc(&ds, sizeof(uint32_t) * 4 * (size_t)nstream));
    CU(cudaMalloc(&dout, sizeof(float) * (size_t)nstream * ndraw));
    CU(cudaMalloc(&dst, sizeof(unsigned long long) * 2 * (size_t)nstream));
    CU(cudaMemcpy(ds, seeds4, sizeof(uint32_t) * 4 * (size_t)nstream, cudaMemcpyHostToDevice));
    CUK(mmcb_k_rng(ds, nstream, ndraw, dout, dst, 0));
    CU(cudaMemcpy(out, dout, sizeof(float) * (size_t)nstream * ndraw, cudaMemcpyDeviceToHost));

    if (state_out) {
        CU(cudaMemcpy(state_out, dst, sizeof(uint64_t) * 2 * (size_t)nstream, cudaMemcpyDeviceToHost));
    }

    cudaFree(ds);
    cudaFree(dout);
    cudaFree(dst);
    return 0;
}

int mmcb_list_gpu(mmcb_gpuinfo* info, int maxcount) {
    // mcx_list_cu_gpu, src/mmc_cu_host.cu:108-198
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);

    if (e != cudaSuccess) {
        g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return 0;
    }

    for (int i = 0; i < ndev && i < maxcount; i++) {
        cudaDeviceProp dp;

        if (cudaGetDeviceProperties(&dp, i) != cudaSuccess) {
            return fail(MMCB_ERR_CUDA, "cudaGetDeviceProperties failed");
        }

        mmcb_gpuinfo& g = info[i];
        memset(&g, 0, sizeof(g));
        snprintf(g.name, sizeof(g.name), "%s", dp.name);
        g.id = i + 1;
        g.devcount = ndev;
        g.major = dp.major;
        g.minor = dp.minor;
        g.globalmem = dp.totalGlobalMem;
        g.constmem = dp.totalConstMem;
        g.sharedmem = dp.sharedMemPerBlock;
        g.regcount = dp.regsPerBlock;
        g.clock = dp.clockRate;
        g.sm = dp.multiProcessorCount;
        g.core = dp.multiProcessorCount * 128;
        g.maxmpthread = dp.maxThreadsPerMultiProcessor;
        g.autoblock = 128;
        g.autothread = (size_t)g.autoblock * (dp.maxThreadsPerMultiProcessor / 2 / g.autoblock) * g.sm;
        g.maxgate = 0;
    }

    return ndev;
}

int mmcb_mesh_volumes(int nn, const float* node, int ne, int* elem_inout, const int* type, float* evol, float* nvol) {
    if (!node || !elem_inout || !evol || !nvol) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    volumes(nn, node, ne, elem_inout, type, evol, nvol);
    return 0;
}

int mmcb_mesh_facenb(int ne, const int* elem, int* facenb) {
    if (!elem || !facenb) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    facenb_build(ne, elem, facenb);
    return 0;
}

int mmcb_mesh_initelem(int nn, const float* node, int ne, const int* elem, const float* srcpos, float* bary4) {
    (void)nn;
    float b[4];
    int e0 = initelem(node, elem, ne, srcpos, bary4 ? bary4 : b);
    return e0;
}

static void fill_sizes(const Cfg& c, int nf, mmcb_sizes* sz) {
    sz->maxgate = c.maxgate;
    sz->datalen = c.datalen;
    sz->reclen = c.reclen;
    sz->nf = nf;
    sz->srcnum = c.c.srcnum;
    memcpy(sz->dim, c.dim, sizeof(sz->dim));
    sz->nslots = c.nslots;
    sz->fieldlen = (size_t)c.datalen * c.maxgate * c.c.srcnum * c.nslots;
    sz->adj_ns = c.adj_ns;
    sz->adj_nd = c.adj_nd;
    // [datalen][Ns*Nd] per component; RF doubles (Re, Im), dual types double again (src/mmc_cu_host.cu:1085-1094,1256-1263)
    sz->jacoblen = (size_t)c.datalen * c.adj_ns * c.adj_nd * (c.isrf ? 2 : 1) * (c.adj_dual ? 2 : 1);
}

int mmcb_query_sizes(const mmcb_config* cfg, const mmcb_mesh* mesh, mmcb_sizes* sz) {
    Cfg c;
    PrepMesh m;
    int rc = validate(cfg, mesh, c);

    if (rc) {
        return rc;
    }

    if ((rc = prepare_mesh(mesh, c, m, NULL, false, true))) {
        return rc;
    }

    fill_sizes(c, m.nf, sz);
    return 0;
}

mmcb_session* mmcb_create(const mmcb_config* cfg, const mmcb_mesh* mesh, int device) {
    mmcb_session* s = new mmcb_session();
    int rc = session_build(s, cfg, mesh, device);

    if (rc) {
        std::string keep = g_err;
        session_free(s);
        g_err = keep;
        return NULL;
    }

    return s;
}

void mmcb_destroy(mmcb_session* s) {
    session_free(s);
}

int mmcb_get_sizes(mmcb_session* s, mmcb_sizes* sz) {
    if (!s || !sz) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    fill_sizes(s->cfg, s->mesh.nf, sz);
    return 0;
}

int mmcb_set_field_buffer(mmcb_session* s, void* device_ptr) {
    if (!s || !device_ptr) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    CU(cudaSetDevice(s->device));

    if (!s->field_external && s->d_field) {      // on the session's own stream: the thread-local default may belong to another session
        CU(cudaFreeAsync(s->d_field, s->stream));
    }

    s->d_field = device_ptr;
    s->ka.field = device_ptr;
    s->field_external = true;
    return 0;
}

int mmcb_launch(mmcb_session* s, uint64_t nphoton, uint64_t photon_offset, int seed, int seed_offset, void* cuda_stream) {
    if (!s) {
        return fail(MMCB_ERR_INPUT, "null session");
    }

    if (nphoton >= 0xFFF00000ull) {
        return fail(MMCB_ERR_LIMIT, "one launch takes fewer than 2^32 photons; use respin (-r) for %llu", (unsigned long long)nphoton);
    }

    CU(cudaSetDevice(s->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->stream;
    // per-thread seeds: srand(seed); Pseed[j]=rand() (src/mmc_cu_host.cu:438,529-540); slices for ranks/respins
    s->hseeds.resize(4 * (size_t)s->nthread);
    {
        const size_t skip = (size_t)seed_offset * 4 * (size_t)s->nthread;

        if (s->next_valid && s->next_seed == seed && s->next_start == skip && s->hseeds_next.size() == s->hseeds.size()) {
            s->hseeds.swap(s->hseeds_next);     // generated while the previous launch was running (below)
        } else {
            if (!s->seedgen || s->seedgen_seed != seed || s->seedgen_pos > skip) {
                delete s->seedgen;
                s->seedgen = new GlibcRand((unsigned int)seed);
                s->seedgen_seed = seed;
                s->seedgen_pos = 0;
            }

            for (; s->seedgen_pos < skip; s->seedgen_pos++) {
                s->seedgen->next();
            }

            s->seedgen->fill(s->hseeds.data(), s->hseeds.size());
            s->seedgen_pos += s->hseeds.size();
        }

        s->next_valid = false;
    }

    CU(cudaMemcpyAsync(s->d_seeds, s->hseeds.data(), s->hseeds.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const mmcb_config& c = s->cfg.c;
    // first big launch of a session: a pilot batch shows where the deposits pile up.  Time-resolved single-slot runs use a SCOUT: extra
    // photons that live for the first time gate only and deposit into a scratch copy of that gate's block; they are discarded, so
    // the scout costs ~0.1 % of the run and has no long-lived stragglers (the hottest lines are the early-time lines next to the
    // source).  Other runs (one gate, multi-slot) take the pilot out of the requested photons and keep its results.
    bool pilot = s->hot_allowed && !s->hot_ready && (c.hotcache > 0 ? nphoton >= 4096 : nphoton >= 500000);
    uint64_t n0 = pilot ? std::min<uint64_t>(std::max<uint64_t>(nphoton / 64, 16384), 262144) : 0;
    n0 = std::min(n0, nphoton / 2);
    CU(cudaEventRecord(s->ev0, st));

    const bool scout = pilot && s->cfg.maxgate > 1 && s->cfg.nslots == 1 && !getenv("MMCB_NO_SCOUT");

    // count-mode scout: deposits counted per line and set against the estimated cost of the walk (mmcb_hot_floor_kernel).  Forced-on caches
    // (hotcache > 0: tests, tuning) and the re-packing kernels keep the weight-ranked selection.  MMCB_HOT_BYWEIGHT=1 restores it for A/B runs.
    const bool countscout = scout && c.hotcache == 0 && !s->repack && !getenv("MMCB_HOT_BYWEIGHT");

    if (scout) {            // a few ten thousand photons rank the lines near the source well enough (the selection works on ratios)
        n0 = std::min<uint64_t>(std::max<uint64_t>(nphoton / 256, 16384), 65536);
    }

    if (scout) {
        const size_t flen = (size_t)s->kp.framelen * c.srcnum, accsize = s->acc_double ? 8 : 4;
        char* scr = NULL;           // [volume of gate 0][energy tot/esc][raytet][detcount, trajcount]
        const size_t tail = sizeof(double) * (2 * MMCB_MAX_SRCNUM + 1) + 2 * sizeof(unsigned int), voff = (flen * accsize + 15) / 16 * 16;
        CU(cudaMallocAsync(&scr, voff + tail, st));
        CU(cudaMemsetAsync(scr, 0, voff + tail, st));
        mmcb_kparam& kp = s->kp_pilot;
        kp = s->kp;
        kp.maxgate = 1;
        kp.tend = c.tstart + c.tstep;
        kp.nphoton = n0;
        kp.photon_offset = photon_offset;
        kp.threadphoton = (int)(n0 / (uint64_t)s->nthread);
        kp.oddphotons = (int)(n0 - (uint64_t)kp.threadphoton * s->nthread);
        kp.hotcache = 0;
        kp.countmode = countscout ? 1 : 0;
        kp.savetraj = 0;
        kp.issaveref = 0;
        kp.issavedet = 0;
        kp.fieldlen = (unsigned int)flen;
        kp.omega = 0.f;
        mmcb_kargs ka = s->ka;
        ka.field = scr;
        ka.field_im = NULL;
        ka.dref = NULL;
        ka.energy = (double*)(scr + voff);
        ka.raytet = ka.energy + 2 * MMCB_MAX_SRCNUM;
        ka.detcount = (unsigned int*)(ka.raytet + 1);
        ka.trajcount = ka.detcount + 1;
        CU(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), st));
        CUK(mmcb_k_upload_param(&kp, s->cfg.detpos.data(), c.detnum, st));
        // count mode lives in the general kernels only (no instruction in the plain ones): the scout of a plain run uses the general
        // instantiation of the same tracer, which walks a pencil / isotropic source identically
        CUK(mmcb_k_launch_photons(&ka, s->grid, s->block, s->smem_scout, c.method, 0, countscout ? 1 : s->isgeneral, 0, s->carveout, s->repack, kvariant(s), st));
        CUK(mmcb_k_hot_select(scr, flen, s->d_hotstat, s->d_hotcand, 2 * MMCB_HOT_SLOTS, s->d_hotkeys, c.hotcache > 0 ? 0.f : MMCB_HOT_MINSHARE,
                              countscout ? ka.raytet : NULL, countscout ? ka.trajcount : NULL, (float)n0, s->cfg.maxgate, st));
        CU(cudaFreeAsync(scr, st));
        s->hot_ready = true;
        pilot = false;
        n0 = 0;
    }

    for (int part = pilot ? 0 : 1; part < 2; part++) {
        const uint64_t n = (part == 0) ? n0 : nphoton - n0, off = (part == 0) ? photon_offset : photon_offset + n0;
        mmcb_kparam& kp = (part == 0) ? s->kp_pilot : s->kp;

        if (part == 0) {
            kp = s->kp;
        }

        kp.nphoton = n;
        kp.photon_offset = off;
        kp.threadphoton = (int)(n / (uint64_t)s->nthread);                 // src/mmc_cu_host.cu:425-429
        kp.oddphotons = (int)(n - (uint64_t)kp.threadphoton * s->nthread);
        kp.hotcache = (part == 1 && s->hot_ready) ? 1 : 0;
        CU(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), st));
        CUK(mmcb_k_upload_param(&kp, s->cfg.detpos.data(), c.detnum, st));
        CUK(mmcb_k_launch_photons(&s->ka, s->grid, s->block, kp.hotcache ? s->smem : s->smem_base, c.method, s->isdet, s->isgeneral, s->cfg.isrf, s->carveout, s->repack, kvariant(s), st));

        if (part == 0) {     // the streams continue from the states the pilot wrote back (no reseeding)
            CUK(mmcb_k_hot_select(s->d_field, s->efieldlen, s->d_hotstat, s->d_hotcand, 2 * MMCB_HOT_SLOTS, s->d_hotkeys,
                                    c.hotcache > 0 ? 0.f : MMCB_HOT_MINSHARE, NULL, NULL, 0.f, 0, st));
            s->hot_ready = true;
        }
    }

    CU(cudaEventRecord(s->ev1, st));
    s->launched += nphoton;

    // the kernel is running: draw the next slice of the host stream now (respin, bench steps and ranks with their own seed ask for
    // consecutive slices), so that the next launch does not start with ~2 ms of rand() on the critical path
    if (s->seedgen && s->seedgen_seed == seed) {
        s->hseeds_next.resize(s->hseeds.size());
        s->next_start = s->seedgen_pos;

        s->seedgen->fill(s->hseeds_next.data(), s->hseeds_next.size());

        s->seedgen_pos += s->hseeds_next.size();
        s->next_seed = seed;
        s->next_valid = true;
    }

    return 0;
}

int mmcb_sync(mmcb_session* s) {
    if (!s) {
        return fail(MMCB_ERR_INPUT, "null session");
    }

    CU(cudaSetDevice(s->device));
    CU(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_ms = ms;
    CU(cudaGetLastError());
    return 0;
}

int mmcb_last_kernel_ms(mmcb_session* s, float* ms) {
    if (!s || !ms) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    *ms = s->last_ms;
    return 0;
}

int mmcb_get_devptrs(mmcb_session* s, mmcb_devptrs* p) {
    if (!s || !p) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    p->field = s->d_field;
    p->fieldlen = s->efieldlen;
    p->field_is_double = s->acc_double;
    p->energy = s->d_energy;
    p->raytet = s->d_raytet;
    p->detected = s->d_detected;
    p->detcount = s->d_detcount;
    p->reclen = s->cfg.reclen;
    p->detseed = (uint64_t*)s->d_detseed;
    p->dref = s->d_dref;
    p->dreflen = (size_t)s->mesh.nf * s->cfg.maxgate;
    p->field_im = s->d_field_im;
    return 0;
}

int mmcb_get_tables(mmcb_session* s, void* tetrec_out, float* cent_out, int* facenb_out) {
    if (!s) {
        return fail(MMCB_ERR_INPUT, "null session");
    }

    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));

    if (tetrec_out) {
        CU(cudaMemcpy(tetrec_out, s->d_tet, sizeof(mmcb_tetrec) * (size_t)s->mesh.ne, cudaMemcpyDeviceToHost));
    }

    if (cent_out) {
        CU(cudaMemcpy(cent_out, s->d_cent, sizeof(float4) * (size_t)s->mesh.ne, cudaMemcpyDeviceToHost));
    }

    if (facenb_out) {
        memcpy(facenb_out, s->mesh.facenb.data(), sizeof(int) * s->mesh.facenb.size());
    }

    return 0;
}

int mmcb_reset(mmcb_session* s) {
    if (!s) {
        return fail(MMCB_ERR_INPUT, "null session");
    }

    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->d_field, 0, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));

    if (s->d_field_im) {
        CU(cudaMemsetAsync(s->d_field_im, 0, s->efieldlen * (s->acc_double ? 8 : 4), s->stream));
    }

    CU(cudaMemsetAsync(s->d_energy, 0, sizeof(double) * 2 * MMCB_MAX_SRCNUM, s->stream));
    CU(cudaMemsetAsync(s->d_raytet, 0, sizeof(double), s->stream));
    CU(cudaMemsetAsync(s->d_detcount, 0, sizeof(unsigned int), s->stream));
    CU(cudaMemsetAsync(s->d_trajcount, 0, sizeof(unsigned int), s->stream));

    if (s->d_dref) {
        CU(cudaMemsetAsync(s->d_dref, 0, sizeof(double) * (size_t)s->mesh.nf * s->cfg.maxgate, s->stream));
    }

    CU(cudaStreamSynchronize(s->stream));
    s->launched = 0;
    return 0;
}

static int rc_dev_alloc_float(float** d, const std::vector<float>& h, cudaStream_t st) {
    cudaError_t e = cudaMallocAsync(d, sizeof(float) * h.size(), st);

    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(*d, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice, st);
    }

    if (e != cudaSuccess) {
        return fail(MMCB_ERR_CUDA, "%s", cudaGetErrorString(e));
    }

    return 0;
}

// ---- device volume -> caller's (pageable) array.  A plain cudaMemcpy into pageable memory runs at ~4 GB/s on this box (the driver stages
// it through one small pinned buffer): 225 ms for the 862 MB volume of BASELINE config C3, six times its kernel.  Here eight host threads
// each take an eighth of the volume: asynchronous copies into their own two pinned 4 MB buffers on their own stream, and while one
// buffer is in flight the other is copied (or added) into the caller's array -- PCIe and the host-side copy overlap, and the first-touch
// page faults of a fresh result array are spread over the threads.  The pinned ring (64 MB per device) is allocated on first use and kept.
namespace {

struct PinnedRing {
    static const int T = 8, B = 2;
    static const size_t CHUNK = (size_t)4 << 20;
    void* buf[T][B];
    cudaStream_t st[T];
    cudaEvent_t ev[T][B];
    bool ready = false;
};

PinnedRing g_ring[MMCB_MAX_DEVICES_RING];
std::mutex g_ring_mutex;

int volume_to_host(int device, const double* d_src, double* host, size_t n, bool add) {
    const size_t bytes = n * sizeof(double);

    if (bytes < ((size_t)4 << 20) || device < 0 || device >= MMCB_MAX_DEVICES_RING || getenv("MMCB_NO_PINNED_RING")) {
        if (!add) {
            CU(cudaMemcpy(host, d_src, bytes, cudaMemcpyDeviceToHost));
        } else {
            std::vector<double> W(n);
            CU(cudaMemcpy(W.data(), d_src, bytes, cudaMemcpyDeviceToHost));

            for (size_t i = 0; i < n; i++) {
                host[i] += W[i];
            }
        }

        return 0;
    }

    std::lock_guard<std::mutex> lock(g_ring_mutex);
    PinnedRing& R = g_ring[device];

    if (!R.ready) {
        for (int t = 0; t < PinnedRing::T; t++) {
            CU(cudaStreamCreateWithFlags(&R.st[t], cudaStreamNonBlocking));

            for (int b = 0; b < PinnedRing::B; b++) {
                CU(cudaHostAlloc(&R.buf[t][b], PinnedRing::CHUNK, cudaHostAllocDefault));
                CU(cudaEventCreateWithFlags(&R.ev[t][b], cudaEventDisableTiming));
            }
        }

        R.ready = true;
    }

    const size_t per = (n + PinnedRing::T - 1) / PinnedRing::T, cn = PinnedRing::CHUNK / sizeof(double);
    int rc[PinnedRing::T] = {0};
    std::vector<std::thread> workers;

    for (int t = 0; t < PinnedRing::T; t++) {
        workers.emplace_back([&, t]() {
            if (cudaSetDevice(device) != cudaSuccess) {
                rc[t] = 1;
                return;
            }

            const size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
            const size_t nchunk = (hi - lo + cn - 1) / cn;

            for (size_t c = 0; c <= nchunk; c++) {       // chunk c is issued, then chunk c - 1 is consumed
                if (c < nchunk) {
                    const size_t off = lo + c * cn, len = std::min(cn, hi - off);

                    if (cudaMemcpyAsync(R.buf[t][c & 1], d_src + off, len * sizeof(double), cudaMemcpyDeviceToHost, R.st[t]) != cudaSuccess ||
                            cudaEventRecord(R.ev[t][c & 1], R.st[t]) != cudaSuccess) {
                        rc[t] = 1;
                        return;
                    }
                }

                if (c > 0) {
                    const size_t off = lo + (c - 1) * cn, len = std::min(cn, hi - off);
                    const double* src = (const double*)R.buf[t][(c - 1) & 1];

                    if (cudaEventSynchronize(R.ev[t][(c - 1) & 1]) != cudaSuccess) {
                        rc[t] = 1;
                        return;
                    }

                    if (!add) {
                        memcpy(host + off, src, len * sizeof(double));
                    } else {
                        for (size_t i = 0; i < len; i++) {
                            host[off + i] += src[i];
                        }
                    }
                }
            }
        });
    }

    for (std::thread& w : workers) {
        w.join();
    }

    for (int t = 0; t < PinnedRing::T; t++) {
        if (rc[t]) {
            return fail(MMCB_ERR_CUDA, "volume download failed (%s)", cudaGetErrorString(cudaGetLastError()));
        }
    }

    return 0;
}

}   // namespace

int mmcb_fetch(mmcb_session* s, const double* energytot, const double* energyesc, mmcb_output* out) {
    if (!s || !out) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    Trace tr;
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    const mmcb_config& c = s->cfg.c;
    const PrepMesh& m = s->mesh;
    double en[2 * MMCB_MAX_SRCNUM];
    CU(cudaMemcpy(en, s->d_energy, sizeof(en), cudaMemcpyDeviceToHost));

    for (int j = 0; j < MMCB_MAX_SRCNUM; j++) {
        out->energytot[j] = energytot ? energytot[j] : en[j];
        out->energyesc[j] = energyesc ? energyesc[j] : en[MMCB_MAX_SRCNUM + j];
    }

    CU(cudaMemcpy(&out->raytet, s->d_raytet, sizeof(double), cudaMemcpyDeviceToHost));
    out->kernel_ms = s->last_ms;
    out->e0 = c.e0;
    out->normalizer = 1.0;
    unsigned int det = 0, traj = 0;
    CU(cudaMemcpy(&det, s->d_detcount, sizeof(det), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&traj, s->d_trajcount, sizeof(traj), cudaMemcpyDeviceToHost));
    out->detectedtotal = det;
    out->detectedcount = s->isdet ? std::min(det, c.maxdetphoton) : 0;   // overflow is a warning, data truncated (src/mmc_cu_host.cu:823-834)
    out->trajcount = std::min(traj, c.maxjumpdebug);

    if (getenv("MMCB_COUNT_FIX")) {     // analysis builds (-DMMCB_COUNT_FIX): see the kernel's not-found path
        fprintf(stderr, "[mmcb] not-found events %u, photons dropped after %d retries %u\n", traj & 0xFFFFFu, MMCB_MAX_TRIAL, traj >> 20);
    }

    if (out->detected && out->detectedcount) {
        CU(cudaMemcpy(out->detected, s->d_detected, sizeof(float) * (size_t)out->detectedcount * s->cfg.reclen, cudaMemcpyDeviceToHost));
    }

    if (out->detseed && out->detectedcount && s->d_detseed) {
        CU(cudaMemcpy(out->detseed, s->d_detseed, sizeof(uint64_t) * 2 * (size_t)out->detectedcount, cudaMemcpyDeviceToHost));
    }

    if (out->traj && out->trajcount && s->d_traj) {
        CU(cudaMemcpy(out->traj, s->d_traj, sizeof(float) * MMCB_DEBUG_REC * (size_t)out->trajcount, cudaMemcpyDeviceToHost));
    }

    std::vector<double> dref;

    if (c.issaveref && s->d_dref) {
        dref.resize((size_t)m.nf * s->cfg.maxgate);
        CU(cudaMemcpy(dref.data(), s->d_dref, sizeof(double) * dref.size(), cudaMemcpyDeviceToHost));
    }

    tr.mark("fetch: scalars+records");

    if (tr.on && s->d_hotstat) {
        unsigned int st[MMCB_HOT_STAT_WORDS];
        CU(cudaMemcpy(st, s->d_hotstat, sizeof(st), cudaMemcpyDeviceToHost));
        float mx, tot;
        memcpy(&mx, &st[0], 4);
        memcpy(&tot, &st[MMCB_HOT_STAT_TOTAL], 4);
        fprintf(stderr, "[mmcb] hot-line cache: ready=%d useful=%u candidates=%u hottest line holds %.4f of the pilot %s\n", (int)s->hot_ready,
                st[MMCB_HOT_STAT_USEFUL], st[1], tot > 0.f ? mx / tot : 0.f, st[MMCB_HOT_STAT_FLOOR] ? "deposits" : "weight");

        if (st[MMCB_HOT_STAT_FLOOR]) {
            float fl, se;
            memcpy(&fl, &st[MMCB_HOT_STAT_FLOOR], 4);
            memcpy(&se, &st[MMCB_HOT_STAT_STEPS], 4);
            fprintf(stderr, "[mmcb] count-mode scout: estimated %.0f ray-tet steps per photon, hottest line %.0f deposits, candidate floor %.0f deposits\n", se, mx, fl);
        }
    }

    // ---- the volume: elem -> node spreading, mesh_normalize (src/mmc_mesh.c:2154-2279) and the slot broadcast of the normaliser
    //      (src/mmc_cu_host.cu:997-1060) all run on the device (mmcb_post.cu: mmcb_norm_*); the volume crosses PCIe once, in its final
    //      form.  Layout: patterns interleaved [(gate*datalen + i)*srcnum + p]; multi-slot runs (srcnum == 1) one [maxgate][datalen]
    //      block per slot; RF runs a second volume (imaginary part) of the same shape.
    double* d_re = NULL, *d_im = NULL;              // working copies: the accumulators stay untouched, the session may keep adding to them

    if (out->field) {
        const bool nodal = (!s->isgrid && !s->ishp && c.basisorder);    // BLB deposits per element; nodal output is spread here
        const bool rf = s->cfg.isrf && s->d_field_im;
        const int nslots = s->cfg.nslots, datalen = s->cfg.datalen, maxgate = s->cfg.maxgate, srcnum = c.srcnum;
        const size_t n = s->fieldlen, n1 = (size_t)datalen * maxgate * srcnum;      // whole volume / block of the first slot
        const double* src_re = (const double*)s->d_field, *src_im = rf ? (const double*)s->d_field_im : NULL;
        auto working_copy = [&](const void* acc, double** d, const double** src, bool force) -> int {
            if (nodal) {            // slot blocks are consecutive gate blocks of the same stride: maxgate * nslots "gates"
                CU(cudaMallocAsync(d, sizeof(double) * n, s->stream));
                CU(cudaMemsetAsync(*d, 0, sizeof(double) * n, s->stream));
                CUK(mmcb_k_spread_nodes(acc, *d, s->d_elem, m.ne, m.nn, maxgate * nslots, srcnum, s->stream));
            } else if (!s->acc_double) {
                CU(cudaMallocAsync(d, sizeof(double) * n, s->stream));
                CUK(mmcb_k_acc_to_double(acc, *d, n, s->stream));
            } else if (force) {
                CU(cudaMallocAsync(d, sizeof(double) * n, s->stream));
                CU(cudaMemcpyAsync(*d, acc, sizeof(double) * n, cudaMemcpyDeviceToDevice, s->stream));
            } else {
                return 0;
            }

            *src = *d;
            return 0;
        };

        if (working_copy(s->d_field, &d_re, &src_re, c.isnormalized != 0) || (rf && working_copy(s->d_field_im, &d_im, &src_im, c.isnormalized != 0))) {
            return g_code;
        }

        if (c.isnormalized) {
            const bool replay = (c.seed == MMCB_SEED_FROM_FILE && (c.outputtype == MMCB_OT_JACOBIAN || c.outputtype == MMCB_OT_WL || c.outputtype == MMCB_OT_WP));
            const bool basis1 = (!s->isgrid && c.basisorder), basis0 = (!s->isgrid && !c.basisorder);
            const bool needdep = !replay && c.outputtype != MMCB_OT_ENERGY && !s->isgrid;       // the energy-deposit sum of :2213-2260
            const bool im1 = rf && srcnum == 1;         // mesh_normalize's imag_slot
            double fac[16], dep[MMCB_MAX_SRCNUM];
            double* d_dep = NULL;
            float* d_evol = NULL, *d_emua = NULL, *d_nvol = NULL;

            if (basis1 && (needdep || nslots > srcnum) && rc_dev_alloc_float(&d_nvol, m.nvol, s->stream)) {
                return g_code;
            }

            if (needdep) {
                std::vector<float> emua(m.ne);

                for (int i = 0; i < m.ne; i++) {
                    emua[i] = m.med[m.type[i]].mua;
                }

                if (rc_dev_alloc_float(&d_evol, m.evol, s->stream) || rc_dev_alloc_float(&d_emua, emua, s->stream)) {
                    return g_code;
                }

                CU(cudaMallocAsync(&d_dep, sizeof(double) * MMCB_MAX_SRCNUM, s->stream));
                CU(cudaMemsetAsync(d_dep, 0, sizeof(double) * MMCB_MAX_SRCNUM, s->stream));

                if (basis1) {       // W /= nvol, then sum_e (sum over gates and the 4 nodes of |phi|) evol mua
                    CUK(mmcb_k_norm_nvol(d_re, n1, m.nn, srcnum, d_nvol, s->stream));

                    if (im1) {
                        CUK(mmcb_k_norm_nvol(d_im, n1, m.nn, srcnum, d_nvol, s->stream));
                    }

                    CUK(mmcb_k_norm_elemdep(d_re, im1 ? d_im : NULL, s->d_elem, d_evol, d_emua, m.ne, m.nn, maxgate, srcnum, d_dep, s->stream));
                } else {
                    CUK(mmcb_k_norm_sum(d_re, n1 / srcnum, srcnum, d_dep, s->stream));
                }

                CU(cudaMemcpyAsync(dep, d_dep, sizeof(double) * MMCB_MAX_SRCNUM, cudaMemcpyDeviceToHost, s->stream));
                CU(cudaStreamSynchronize(s->stream));
            }

            double sum = 0;

            for (int j = 0; j < 16; j++) {
                fac[j] = 1.0;
            }

            for (int j = 0; j < srcnum; j++) {
                const float Eabsorb = (float)(out->energytot[j] - out->energyesc[j]), Etotal = (float)out->energytot[j];   // :988, float args
                double nz;

                if (replay) {
                    nz = (c.outputtype == MMCB_OT_JACOBIAN) ? 1.f / (1e-4f * c.nphoton) : 1.f / Etotal;
                } else if (c.outputtype == MMCB_OT_ENERGY) {
                    nz = 1.f / Etotal;
                } else {
                    if (s->isgrid) {
                        nz = 1.0 / (Etotal * c.unitinmm * c.unitinmm * c.unitinmm);
                    } else if (basis1) {
                        nz = Eabsorb / (Etotal * dep[j] * 0.25f);
                    } else {
                        nz = Eabsorb / (Etotal * dep[j]);
                    }

                    if (c.outputtype == MMCB_OT_FLUX) {
                        nz /= c.tstep;
                    }
                }

                fac[j] = nz;
                sum += nz;
            }

            out->normalizer = sum / srcnum;
            const bool divide = needdep && basis0;      // basisorder 0: every entry is divided by evol * mua first
            CUK(mmcb_k_norm_scale(d_re, d_re, n1, datalen, srcnum, divide ? d_evol : NULL, divide ? d_emua : NULL, fac, s->stream));

            if (im1 && !(replay || c.outputtype == MMCB_OT_ENERGY)) {
                // both parts of one complex fluence get the same factors; the reference leaves the imaginary part of a per-element
                // volume undivided (src/mmc_mesh.c:2250-2256), which cannot be intended.  The factor is rounded to float (:2268-2270)
                double facf[16];

                for (int j = 0; j < 16; j++) {
                    facf[j] = (double)(float)fac[j];
                }

                CUK(mmcb_k_norm_scale(d_im, d_im, n1, datalen, srcnum, divide ? d_evol : NULL, divide ? d_emua : NULL, facf, s->stream));
            }

            if (n > n1) {           // the slots behind the first: nodal-volume division and the average normaliser (src/mmc_cu_host.cu:997-1060)
                double all[16];

                for (int j = 0; j < 16; j++) {
                    all[j] = out->normalizer;
                }

                if (basis1) {
                    CUK(mmcb_k_norm_nvol(d_re + n1, n - n1, m.nn, srcnum, d_nvol, s->stream));
                }

                CUK(mmcb_k_norm_scale(d_re + n1, d_re + n1, n - n1, datalen, srcnum, NULL, NULL, all, s->stream));

                if (rf) {
                    for (int j = 0; j < 16; j++) {
                        all[j] = (double)(float)out->normalizer;
                    }

                    if (basis1) {
                        CUK(mmcb_k_norm_nvol(d_im + n1, n - n1, m.nn, srcnum, d_nvol, s->stream));
                    }

                    CUK(mmcb_k_norm_scale(d_im + n1, d_im + n1, n - n1, datalen, srcnum, NULL, NULL, all, s->stream));
                }
            }

            if (c.issaveref && !dref.empty()) {          // :2160-2167
                const float nz = 1.f / (float)out->energytot[0];

                for (size_t i = 0; i < dref.size(); i++) {
                    dref[i] *= nz;
                }
            }

            for (void* q : {(void*)d_dep, (void*)d_evol, (void*)d_emua, (void*)d_nvol}) {
                if (q) {
                    cudaFreeAsync(q, s->stream);
                }
            }
        }

        tr.mark("fetch: normalise (device)");
        auto download = [&](const double* d, double* host) -> int {
            CU(cudaStreamSynchronize(s->stream));
            return volume_to_host(s->device, d, host, n, out->overwrite == 0);     // cfg->exportfield[i] += field[i]  (src/mmc_cu_host.cu:918-920)
        };

        if (download(src_re, out->field) || (rf && out->field_im && download(src_im, out->field_im))) {
            return g_code;
        }

        tr.mark("fetch: volume D2H");

        // ---- adjoint Jacobians from the slots' normalised fluence (src/mmc_cu_host.cu:1063-1395): float copies of the volumes, one pass
        //      sums the gates per slot, the pair kernels write [datalen][Ns*Nd] per component
        if (out->jacob && s->cfg.adj_ns > 0 && s->cfg.adj_nd > 0) {
            const int Ns = s->cfg.adj_ns, Nd = s->cfg.adj_nd, dual = s->cfg.adj_dual;
            const size_t N = (size_t)s->cfg.datalen, adjlen = N * Ns * Nd, single = adjlen * (rf ? 2 : 1);
            float* d_f = NULL, *d_cwr = NULL, *d_cwi = NULL, *d_j1 = NULL, *d_j2 = NULL, *d_evol = NULL, *d_nvol = NULL;
            CU(cudaMallocAsync(&d_f, sizeof(float) * n, s->stream));
            CU(cudaMallocAsync(&d_cwr, sizeof(float) * N * nslots, s->stream));

            for (int part = 0; part < (rf ? 2 : 1); part++) {
                if (part) {
                    CU(cudaMallocAsync(&d_cwi, sizeof(float) * N * nslots, s->stream));
                }

                CUK(mmcb_k_double_to_float(part ? src_im : src_re, d_f, n, s->stream));
                CUK(mmcb_k_adj_cw(d_f, part ? d_cwi : d_cwr, N, s->cfg.maxgate, nslots, s->stream));
            }

            const bool want_mua = (c.outputtype == MMCB_OT_ADJOINT || dual), want_d = (c.outputtype != MMCB_OT_ADJOINT);
            CU(cudaMallocAsync(&d_j1, sizeof(float) * single, s->stream));
            CU(cudaMemsetAsync(d_j1, 0, sizeof(float) * single, s->stream));

            if (dual) {
                CU(cudaMallocAsync(&d_j2, sizeof(float) * single, s->stream));
                CU(cudaMemsetAsync(d_j2, 0, sizeof(float) * single, s->stream));
            }

            float* d_mua = want_mua ? d_j1 : NULL, *d_d = want_d ? (dual ? d_j2 : d_j1) : NULL;

            if (s->isgrid) {        // :1253-1395, scaled by -Vvox (J_mua) and -unitinmm (J_D)
                const float vvox = c.unitinmm * c.unitinmm * c.unitinmm;

                if (d_mua) {
                    CUK(mmcb_k_adj_mua(d_cwr, d_cwi, d_mua, N, Ns, Nd, -vvox, s->stream));
                }

                if (d_d) {
                    CUK(mmcb_k_adj_dcoeff(d_cwr, d_cwi, d_d, N, Ns, Nd, (unsigned int)s->cfg.dim[0], (unsigned int)s->cfg.dim[1], -c.unitinmm, s->stream));
                }
            } else {                // :1067-1250
                if ((rc_dev_alloc_float(&d_evol, m.evol, s->stream))) {
                    return g_code;
                }

                if (d_mua && c.adjointmode == 1) {
                    if ((rc_dev_alloc_float(&d_nvol, m.nvol, s->stream))) {
                        return g_code;
                    }

                    CUK(mmcb_k_adj_mesh_nodal(d_cwr, d_cwi, d_nvol, d_mua, m.nn, Ns, Nd, s->stream));
                }

                if ((d_mua && c.adjointmode != 1) || d_d) {
                    CUK(mmcb_k_adj_mesh_full(d_cwr, d_cwi, s->d_elem, s->d_node, d_evol, c.adjointmode == 1 ? NULL : d_mua, d_d, m.ne, m.nn, Ns, Nd,
                                             s->stream));
                }
            }

            // pack: CW [J1] | CW dual [J1, J2] | RF [Re J1, Im J1] | RF dual [Re J1, Re J2, Im J1, Im J2]
            if (!dual) {
                CU(cudaMemcpyAsync(out->jacob, d_j1, sizeof(float) * single, cudaMemcpyDeviceToHost, s->stream));
            } else {
                CU(cudaMemcpyAsync(out->jacob, d_j1, sizeof(float) * adjlen, cudaMemcpyDeviceToHost, s->stream));
                CU(cudaMemcpyAsync(out->jacob + adjlen, d_j2, sizeof(float) * adjlen, cudaMemcpyDeviceToHost, s->stream));

                if (rf) {
                    CU(cudaMemcpyAsync(out->jacob + 2 * adjlen, d_j1 + adjlen, sizeof(float) * adjlen, cudaMemcpyDeviceToHost, s->stream));
                    CU(cudaMemcpyAsync(out->jacob + 3 * adjlen, d_j2 + adjlen, sizeof(float) * adjlen, cudaMemcpyDeviceToHost, s->stream));
                }
            }

            CU(cudaStreamSynchronize(s->stream));

            for (float* q : {d_f, d_cwr, d_cwi, d_j1, d_j2, d_evol, d_nvol}) {
                if (q) {
                    cudaFreeAsync(q, s->stream);
                }
            }

            tr.mark("fetch: adjoint Jacobian");
        }

        for (double* q : {d_re, d_im}) {
            if (q) {
                cudaFreeAsync(q, s->stream);
            }
        }
    }

    if (out->dref && !dref.empty()) {
        for (size_t i = 0; i < dref.size(); i++) {
            out->dref[i] = (out->overwrite ? 0.0 : out->dref[i]) + dref[i];
        }
    }

    return 0;
}

// The host has nothing to do while the last photon launch runs.  For a large result volume (BASELINE config C3: 862 MB of doubles)
// the download that follows is bound by the first-touch page faults of the caller's fresh array (~210 000 faults), so they are taken
// here, eight threads wide, while the GPU works: one read-modify-write of the first byte of every 4 KB page leaves the contents as
// they are (accumulate mode adds to them later).  Small volumes skip it (measured neutral for 18 MB, profiles/r1l_negative_results.jsonl).
static void prefault_pages(void* ptr, size_t bytes) {
    static const size_t minbytes = (size_t)(getenv("MMCB_PREFAULT_MIN_MB") ? atoi(getenv("MMCB_PREFAULT_MIN_MB")) : 64) << 20;

    if (!ptr || bytes < minbytes || getenv("MMCB_NO_PREFAULT")) {
        return;
    }

    const int T = 8;
    const size_t page = 4096, npage = (bytes + page - 1) / page, per = (npage + T - 1) / T;
    std::vector<std::thread> workers;

    for (int t = 0; t < T; t++) {
        workers.emplace_back([=]() {
            volatile char* base = (volatile char*)ptr;
            size_t i = per * t, hi = std::min(npage, per * (t + 1));
            // (madvise(MADV_POPULATE_WRITE) over the same ranges was measured slower: 65-97 ms against 22-44 ms for 862 MB)

            for (; i < hi; i++) {
                base[i * page] = base[i * page];
            }
        });
    }

    for (std::thread& w : workers) {
        w.join();
    }
}

int mmcb_run_session(mmcb_session* s, mmcb_output* out) {
    if (!s || !out) {
        return fail(MMCB_ERR_INPUT, "null argument");
    }

    Trace tr;
    int rc = 0;
    float ms = 0.f;
    const int respin = s->cfg.c.respin;
    uint64_t per = s->cfg.c.nphoton / respin;

    for (int it = 0; it < respin && rc == 0; it++) {         // src/mmc_cu_host.cu:656,893-906
        uint64_t n = (it == respin - 1) ? s->cfg.c.nphoton - per * (respin - 1) : per;
        rc = mmcb_launch(s, n, per * it, s->cfg.c.seed, it, NULL);

        if (rc == 0 && it == respin - 1) {
            prefault_pages(out->field, s->fieldlen * sizeof(double));
            prefault_pages(out->field_im, (s->cfg.isrf && out->field_im) ? s->fieldlen * sizeof(double) : 0);
            tr.mark("run: prefault (kernel running)");
        }

        if (rc == 0) {
            rc = mmcb_sync(s);
        }

        ms += s->last_ms;
    }

    tr.mark("run: launch+sync");

    if (rc == 0) {
        rc = mmcb_fetch(s, NULL, NULL, out);
        out->kernel_ms = ms;
    }

    tr.mark("run: fetch");
    return rc;
}

int mmcb_run_simulation(const mmcb_config* cfg, const mmcb_mesh* mesh, int device, mmcb_output* out) {
    if (!out) {
        return fail(MMCB_ERR_INPUT, "null output");
    }

    Trace tr;
    mmcb_session* s = mmcb_create(cfg, mesh, device);

    if (!s) {
        return g_code ? g_code : MMCB_ERR_CUDA;
    }

    tr.mark("run: create");
    int rc = mmcb_run_session(s, out);
    std::string keep = g_err;
    mmcb_destroy(s);
    g_err = keep;
    tr.mark("run: destroy");
    return rc;
}

// -----------------------------------------------------------------------------------------------------------------------------------
// Photon sharding over the GPUs of one box: what mmc_run_cu does with cfg->deviceid / cfg->workload (src/mmc_cu_host.cu:403-429,
// 1538-1553: one host thread per GPU, photons split by workload, results merged under omp critical).  Here: one host thread and one
// session per device for the walk -- no data-path collective --, then an epilogue over NVLink: ncclReduce of the accumulator
// volume(s) and the diffuse reflectance to the first device, the detected-photon rows and seeds gathered behind the first device's
// own rows (counts first, then payload, truncated at maxdetphoton like :823-853), energy tallies summed on the host, and ONE
// normalisation + download from the first device.  NCCL is resolved with dlopen (libnccl.so.2) so that single-GPU users need none.
// -----------------------------------------------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
    void* lib = NULL;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = NULL;
    ncclResult_t (*CommDestroy)(ncclComm_t) = NULL;
    ncclResult_t (*GroupStart)() = NULL;
    ncclResult_t (*GroupEnd)() = NULL;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = NULL;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = NULL;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = NULL;
    const char* (*GetErrorString)(ncclResult_t) = NULL;

    int load() {
        if (lib) {
            return 0;
        }

        const char* names[] = {getenv("MMCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};

        for (const char* n : names) {
            if (n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) {
                break;
            }
        }

        if (!lib) {
            return fail(MMCB_ERR_CUDA, "multi-GPU runs need NCCL: libnccl.so.2 could not be loaded (%s)", dlerror());
        }

#define NCCL_SYM(field, name) do { *(void**)(&field) = dlsym(lib, name); if (!field) return fail(MMCB_ERR_CUDA, "NCCL symbol %s is missing", name); } while (0)
        NCCL_SYM(CommInitAll, "ncclCommInitAll");
        NCCL_SYM(CommDestroy, "ncclCommDestroy");
        NCCL_SYM(GroupStart, "ncclGroupStart");
        NCCL_SYM(GroupEnd, "ncclGroupEnd");
        NCCL_SYM(Reduce, "ncclReduce");
        NCCL_SYM(Send, "ncclSend");
        NCCL_SYM(Recv, "ncclRecv");
        NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
        return 0;
    }
};

NcclApi g_nccl;
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(MMCB_ERR_CUDA, "NCCL error %d (%s) at %s:%d", (int)r_, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)

// photons of device g: nphoton * w_g / sum(w), the last device takes the remainder (the reference truncates every share, :425-429)
std::vector<uint64_t> photon_shares(uint64_t nphoton, int ndev, const float* workload) {
    std::vector<double> w(ndev, 1.0);
    double total = 0;

    for (int g = 0; g < ndev; g++) {
        if (workload && workload[g] > 0.f) {
            w[g] = workload[g];
        }

        total += w[g];
    }

    std::vector<uint64_t> share(ndev);
    uint64_t given = 0;

    for (int g = 0; g < ndev; g++) {
        share[g] = (g == ndev - 1) ? nphoton - given : (uint64_t)((double)nphoton * w[g] / total);
        given += share[g];
    }

    return share;
}

}   // namespace

void mmcb_photon_shares(uint64_t nphoton, int ndev, const float* workload, uint64_t* share) {
    const std::vector<uint64_t> v = photon_shares(nphoton, ndev, workload);
    std::copy(v.begin(), v.end(), share);
}

int mmcb_run_multi(const mmcb_config* cfg, const mmcb_mesh* mesh, int ndev, const int* devices, const float* workload, mmcb_output* out) {
    if (!cfg || !mesh || !out || ndev < 1 || !devices) {
        return fail(MMCB_ERR_INPUT, "null argument or no device");
    }

    if (ndev == 1) {
        return mmcb_run_simulation(cfg, mesh, devices[0], out);
    }

    for (int g = 0; g < ndev; g++)
        for (int h = 0; h < g; h++)
            if (devices[g] == devices[h]) {
                return fail(MMCB_ERR_INPUT, "device %d is listed twice", devices[g]);
            }

    if (workload)
        for (int g = 0; g < ndev; g++)
            if (!(workload[g] >= 0.f)) {
                return fail(MMCB_ERR_INPUT, "workload was unspecified for an active device");       // src/mmc_cu_host.cu:420-422
            }

    if (g_nccl.load()) {
        return g_code;
    }

    Trace tr;
    const std::vector<uint64_t> share = photon_shares(cfg->nphoton, ndev, workload);
    std::vector<uint64_t> first(ndev, 0);

    for (int g = 1; g < ndev; g++) {
        first[g] = first[g - 1] + share[g - 1];
    }

    // ---- the walk: one thread and one session per device; replay runs shard the photon index range, the others draw their thread
    //      seeds from their own host stream srand(seed + 7919 g) (slices of one stream would make every device generate and discard
    //      the other devices' words)
    std::vector<mmcb_session*> sess(ndev, (mmcb_session*)NULL);
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::vector<float> ms(ndev, 0.f);
    std::vector<std::thread> workers;
    const bool replay = (cfg->seed == MMCB_SEED_FROM_FILE);

    for (int g = 0; g < ndev; g++) {
        workers.emplace_back([&, g]() {
            mmcb_config c = *cfg;
            c.nphoton = share[g];

            if (replay) {           // every session holds the whole seed table; the kernel indexes it with photon_offset + id
                c.nphoton = cfg->nphoton;
            }

            mmcb_session* s = mmcb_create(&c, mesh, devices[g]);

            if (!s) {
                rcs[g] = g_code ? g_code : MMCB_ERR_CUDA;
                errs[g] = g_err;
                return;
            }

            sess[g] = s;
            const int respin = s->cfg.c.respin;
            const uint64_t mine = share[g], per = mine / respin;

            for (int it = 0; it < respin && rcs[g] == 0 && mine > 0; it++) {
                const uint64_t n = (it == respin - 1) ? mine - per * (respin - 1) : per;

                if (n == 0) {
                    continue;
                }

                rcs[g] = mmcb_launch(s, n, (replay ? first[g] : 0) + per * it, replay ? cfg->seed : cfg->seed + 7919 * g, it, NULL);

                if (rcs[g] == 0) {
                    rcs[g] = mmcb_sync(s);
                }

                ms[g] += s->last_ms;
            }

            if (rcs[g]) {
                errs[g] = g_err;
            }
        });
    }

    for (std::thread& t : workers) {
        t.join();
    }

    auto cleanup = [&]() {
        for (mmcb_session* s : sess) {
            if (s) {
                session_free(s);
            }
        }
    };

    for (int g = 0; g < ndev; g++) {
        if (rcs[g]) {
            const std::string keep = errs[g];
            const int code = rcs[g];
            cleanup();
            return fail(code, "device %d: %s", devices[g], keep.c_str());
        }
    }

    tr.mark("multi: walk on all devices");
    // ---- epilogue over NVLink
    mmcb_session* s0 = sess[0];
    const size_t accsize = s0->acc_double ? 8 : 4;
    const ncclDataType_t acctype = s0->acc_double ? ncclDouble : ncclFloat;
    std::vector<ncclComm_t> comm(ndev);
    auto body = [&]() -> int {
        NC(g_nccl.CommInitAll(comm.data(), ndev, devices));
        // detector rows: counts first (host), then the payload behind the first device's rows
        std::vector<unsigned int> cnt(ndev, 0);
        std::vector<double> en(2 * MMCB_MAX_SRCNUM, 0.0);
        double raytet = 0;

        for (int g = 0; g < ndev; g++) {
            double e[2 * MMCB_MAX_SRCNUM], r = 0;
            CU(cudaSetDevice(devices[g]));
            CU(cudaMemcpy(e, sess[g]->d_energy, sizeof(e), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&r, sess[g]->d_raytet, sizeof(r), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&cnt[g], sess[g]->d_detcount, sizeof(unsigned int), cudaMemcpyDeviceToHost));

            for (int j = 0; j < 2 * MMCB_MAX_SRCNUM; j++) {
                en[j] += e[j];
            }

            raytet += r;
        }

        std::vector<unsigned int> tcnt(ndev, 0);

        if (s0->d_traj) {
            for (int g = 0; g < ndev; g++) {
                CU(cudaSetDevice(devices[g]));
                CU(cudaMemcpy(&tcnt[g], sess[g]->d_trajcount, sizeof(unsigned int), cudaMemcpyDeviceToHost));
            }
        }

        const unsigned int cap = s0->cfg.c.maxdetphoton;
        const int reclen = s0->cfg.reclen;
        NC(g_nccl.GroupStart());

        for (int g = 0; g < ndev; g++) {        // volume(s) and diffuse reflectance: sum into the first device, in place
            NC(g_nccl.Reduce(sess[g]->d_field, s0->d_field, s0->efieldlen, acctype, ncclSum, 0, comm[g], sess[g]->stream));

            if (s0->d_field_im) {
                NC(g_nccl.Reduce(sess[g]->d_field_im, s0->d_field_im, s0->efieldlen, acctype, ncclSum, 0, comm[g], sess[g]->stream));
            }

            if (s0->d_dref) {
                NC(g_nccl.Reduce(sess[g]->d_dref, s0->d_dref, (size_t)s0->mesh.nf * s0->cfg.maxgate, ncclDouble, ncclSum, 0, comm[g], sess[g]->stream));
            }
        }

        (void)accsize;
        NC(g_nccl.GroupEnd());
        unsigned long long total = std::min(cnt[0], cap), detected_all = cnt[0];

        if (s0->isdet) {
            NC(g_nccl.GroupStart());

            for (int g = 1; g < ndev; g++) {
                detected_all += cnt[g];
                const unsigned int have = std::min(cnt[g], cap);                                     // rows that device stored
                const unsigned int take = (unsigned int)std::min<unsigned long long>(have, cap - total); // rows that still fit (:823-834)

                if (take > 0) {
                    NC(g_nccl.Send(sess[g]->d_detected, (size_t)take * reclen, ncclFloat, 0, comm[g], sess[g]->stream));
                    NC(g_nccl.Recv(s0->d_detected + (size_t)total * reclen, (size_t)take * reclen, ncclFloat, g, comm[0], s0->stream));

                    if (s0->d_detseed) {
                        NC(g_nccl.Send(sess[g]->d_detseed, (size_t)take * 2, ncclUint64, 0, comm[g], sess[g]->stream));
                        NC(g_nccl.Recv(s0->d_detseed + (size_t)total * 2, (size_t)take * 2, ncclUint64, g, comm[0], s0->stream));
                    }
                }

                total += take;
            }

            NC(g_nccl.GroupEnd());
        }

        if (s0->d_traj) {       // trajectory records (-D M), merged like src/mmc_cu_host.cu:790-810 and truncated at maxjumpdebug
            const unsigned int tcap = s0->cfg.c.maxjumpdebug;
            unsigned long long ttotal = std::min(tcnt[0], tcap), tall = tcnt[0];
            NC(g_nccl.GroupStart());

            for (int g = 1; g < ndev; g++) {
                tall += tcnt[g];
                const unsigned int take = (unsigned int)std::min<unsigned long long>(std::min(tcnt[g], tcap), tcap - ttotal);

                if (take > 0) {
                    NC(g_nccl.Send(sess[g]->d_traj, (size_t)take * MMCB_DEBUG_REC, ncclFloat, 0, comm[g], sess[g]->stream));
                    NC(g_nccl.Recv(s0->d_traj + (size_t)ttotal * MMCB_DEBUG_REC, (size_t)take * MMCB_DEBUG_REC, ncclFloat, g, comm[0], s0->stream));
                }

                ttotal += take;
            }

            NC(g_nccl.GroupEnd());
            const unsigned int t32 = (unsigned int)std::min<unsigned long long>(tall, 0xFFFFFFFFull);
            CU(cudaSetDevice(devices[0]));
            CU(cudaMemcpyAsync(s0->d_trajcount, &t32, sizeof(unsigned int), cudaMemcpyHostToDevice, s0->stream));
        }

        for (int g = 0; g < ndev; g++) {
            CU(cudaSetDevice(devices[g]));
            CU(cudaStreamSynchronize(sess[g]->stream));
        }

        // the first session now holds everything; its detected count is the total that hit a detector (may exceed the capacity)
        CU(cudaSetDevice(devices[0]));
        const unsigned int alldet = (unsigned int)std::min<unsigned long long>(detected_all, 0xFFFFFFFFull);
        CU(cudaMemcpy(s0->d_detcount, &alldet, sizeof(unsigned int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s0->d_raytet, &raytet, sizeof(double), cudaMemcpyHostToDevice));
        tr.mark("multi: NCCL reduce + gather");
        // the fetch normalises with cfg.nphoton in replay mode (Jacobian): the first session was created with the whole count there
        s0->cfg.c.nphoton = cfg->nphoton;
        int rc = mmcb_fetch(s0, en.data(), en.data() + MMCB_MAX_SRCNUM, out);
        out->kernel_ms = *std::max_element(ms.begin(), ms.end());
        return rc;
    };
    const int rc = body();
    const std::string keep = g_err;

    for (ncclComm_t c : comm) {
        if (c) {
            g_nccl.CommDestroy(c);
        }
    }

    cleanup();

    if (rc) {
        g_err = keep;
    }

    return rc;
}

}  // extern "C"
