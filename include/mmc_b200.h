/*
 * mmc_b200 -- C-ABI of the B200-native mesh-based Monte Carlo photon engine.
 *
 * This is the drop-in boundary for the ONE hot path of fangq/mmc this repository rebuilds: the
 * per-photon random walk behind
 *     void mmc_run_cu(mcconfig* cfg, tetmesh* mesh, raytracer* tracer)      (src/mmc_cu_host.h:46-62)
 *     void mmc_run_simulation(mcconfig*, tetmesh*, raytracer*, GPUInfo*)    (src/mmc_cu_host.cu:204)
 *     int  mcx_list_cu_gpu(mcconfig*, GPUInfo**)                            (src/mmc_cu_host.cu:108)
 * (citations relative to the reference tree).  The entry points below take plain pointers and sizes
 * only -- no reference structs, no torch types -- so that the reference's own host (mmc.c, mmclab.cpp,
 * pmmc.cpp) can bind them with a ~100-line shim that unpacks mcconfig/tetmesh (INTEGRATION.md shows it;
 * integration/mmc_cu_host_b200.cpp is that shim, compiled against the reference headers).
 *
 * Error convention: every function returns 0 on success and a negative code on failure;
 * mmcb_last_error() returns the message.  The shim forwards failures to mcx_error(id,msg,file,line)
 * exactly like CUDA_ASSERT does in the reference (src/mmc_cu_host.cu:52-53,99-103).
 * There is NO CPU fallback: every entry point that computes fails with MMCB_ERR_CUDA when no CUDA
 * device is usable.
 */
#ifndef MMC_B200_H
#define MMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMCB_VERSION 0x00010000

/* error codes (negative, like the ids the reference passes to mcx_error) */
#define MMCB_OK            0
#define MMCB_ERR_CUDA     (-1)   /* CUDA runtime failure / no device (reference: CUDA_ASSERT -> mcx_error) */
#define MMCB_ERR_INPUT    (-2)   /* invalid configuration (reference: mcx_validatecfg, MMC_ERROR(-2,...)) */
#define MMCB_ERR_MESH     (-3)   /* mesh problem, e.g. "initial element does not enclose the source!" */
#define MMCB_ERR_MEMORY   (-4)
#define MMCB_ERR_LIMIT    (-5)   /* too many media/detectors for the constant tables (reference: MAX_PROP) */

/* ray tracers, src/mmc_utils.h enum TRTMethod ('-M p|h|b|s|g') */
enum { MMCB_RT_PLUCKER = 0, MMCB_RT_HAVEL = 1, MMCB_RT_BADOUEL = 2, MMCB_RT_BLBADOUEL = 3, MMCB_RT_BLBADOUEL_GRID = 4 };
/* output types, enum TOutputType ('-O X|F|E|J|L|P') */
enum { MMCB_OT_FLUX = 0, MMCB_OT_FLUENCE = 1, MMCB_OT_ENERGY = 2, MMCB_OT_JACOBIAN = 3, MMCB_OT_WL = 4, MMCB_OT_WP = 5,
       /* '-O R|A|D|W' (src/mmc_utils.h:106-111): RF forward and the adjoint Jacobians computed from a multi-slot run */
       MMCB_OT_RF = 6, MMCB_OT_RFMUS = 7, MMCB_OT_ADJOINT = 8, MMCB_OT_ADJOINT_DCOEFF = 9, MMCB_OT_ADJOINT_MUS = 10,
       MMCB_OT_ADJOINT_MUSP = 11, MMCB_OT_ADJOINT_MUAD = 12, MMCB_OT_ADJOINT_MUAMUSP = 13
     };
/* boundary conditions, enum TBoundary ('-b 0|1|2|3') */
enum { MMCB_BC_NOREFLECT = 0, MMCB_BC_REFLECT = 1, MMCB_BC_ABSORB_EXTERIOR = 2, MMCB_BC_MIRROR = 3 };
/* source types, src/mmc_const.h:49-62 */
enum { MMCB_SRC_PENCIL = 0, MMCB_SRC_ISOTROPIC = 1, MMCB_SRC_CONE = 2, MMCB_SRC_GAUSSIAN = 3, MMCB_SRC_PLANAR = 4,
       MMCB_SRC_PATTERN = 5, MMCB_SRC_FOURIER = 6, MMCB_SRC_ARCSINE = 7, MMCB_SRC_DISK = 8, MMCB_SRC_FOURIERX = 9,
       MMCB_SRC_FOURIERX2D = 10, MMCB_SRC_ZGAUSSIAN = 11, MMCB_SRC_LINE = 12, MMCB_SRC_SLIT = 13
     };
#define MMCB_SEED_FROM_FILE (-999)     /* src/mmc_utils.h:58: replay, per-photon seeds */

/* photon scheduling */
enum { MMCB_SCHED_DYNAMIC = 0,   /* persistent warps pull photon ids from a global pool (default) */
       MMCB_SCHED_STATIC  = 1    /* the reference's split: thread i runs photons [i*n+min(i,odd), ...) (src/mmc_core.cl:2190-2203) */
     };

typedef struct mmcb_medium {     /* src/mmc_utils.h:155-160 */
    float mua, mus, g, n;
} mmcb_medium;

/* Mesh as the reference's loaders leave it BEFORE mmc_prep (src/mmc_mesh.h:87-122).  All pointers are
 * host memory owned by the caller; arrays are copied.  Node indices are 1-based. */
typedef struct mmcb_mesh {
    int nn, ne;
    int prop;                    /* number of media, NOT counting medium 0 (mesh->prop) */
    const float* node;           /* nn*3 (FLOAT3) */
    const int*   elem;           /* ne*4 */
    const int*   type;           /* ne; -1 = wide-field source candidate, -2 = wide-field detector (src/mmc_mesh.c:390-427) */
    const mmcb_medium* med;      /* prop+1 entries, med[0] = background (mua=mus=0,g=1,n=nout) */
    const int*   facenb;         /* optional ne*4 (0 = exterior); NULL => computed (mesh_getfacenb) */
    const float* evol;           /* optional ne; NULL => computed (mesh_getvolume), incl. the node-3/4 swap */
    const float* nvol;           /* optional nn */
} mmcb_mesh;

/* The subset of mcconfig (src/mmc_utils.h:210-345) the photon path reads; same names, same meaning. */
typedef struct mmcb_config {
    uint64_t nphoton;
    int   seed;                  /* RNG seed; MMCB_SEED_FROM_FILE => replay from photonseed[] */
    float srcpos[4];             /* xyz (mm) */
    float srcdir[4];             /* unit vector, w = focal length (srcdir.w) */
    int   srctype;
    float srcparam1[4], srcparam2[4];
    const float* srcpattern;     /* Nx*Ny*srcnum floats (pattern source) */
    int   srcnum;
    float tstart, tstep, tend;   /* seconds */
    int   e0;                    /* initial element, 1-based; 0 => search (mesh_initelem) */
    int   isreflect;             /* TBoundary */
    int   isnormalized;          /* 1 normalise like mesh_normalize; 0 raw sums */
    int   issavedet, ismomentum, issaveexit, issaveseed, isspecular, issaveref;
    int   method;                /* TRTMethod */
    int   basisorder;            /* 0 per-element, 1 nodal */
    int   outputtype;            /* TOutputType */
    float roulettesize, minenergy, nout;
    int   voidtime;
    float unitinmm;
    float steps;                 /* dual-grid voxel edge (cfg->steps.x) */
    int   detnum;
    const float* detpos;         /* detnum*4 (x,y,z,radius) */
    unsigned int maxdetphoton;
    /* replay (src/mmc_mesh.c:815-898 fills these from an .mch file) */
    const uint64_t* photonseed;  /* nphoton*2 (16 B xorshift128+ state per photon) */
    const float* replayweight;   /* nphoton */
    const float* replaytime;     /* nphoton */
    /* trajectory debug ('-D M') */
    int   savetraj; unsigned int maxjumpdebug;
    /* launch shape; 0 = autopilot */
    int   nthread;               /* total device threads = number of RNG streams */
    int   nblocksize;
    int   schedule;              /* MMCB_SCHED_* */
    int   respin;                /* repeat count, results accumulate (-r) */
    int   hotcache;              /* CTA-private sums for the hottest 128-byte lines of the volume, picked from a pilot batch:
                                    0 = auto (on from 500 000 photons per launch), 1 = always, -1 = never */
    /* multi-slot sources / adjoint mode and RF forward (src/mmc_utils.h:144-152,240,275,303-311; src/mmc_core.cl:1431-1515) */
    float omega;                 /* modulation angular frequency (rad/s); > 0 (and no replay) => complex (RF) fluence */
    int   srcid;                 /* 0 default; -1 all slots of srcdata, one field block per slot; -2 append detectors as sources
                                    without Jacobian output; > 0 only slot srcid-1 (cfg->srcid) */
    int   extrasrclen;           /* entries of srcdata (cfg->extrasrclen) */
    const float* srcdata;        /* extrasrclen*16 floats, ExtraSrc: srcpos(w = launch weight), srcdir(w = focal length), srcparam1
                                    (x = disk radius), srcparam2 (w = enclosing element, filled by the library when 0) */
    const float* detdir;         /* detnum*4 detector normals (w = focal length): needed to turn detectors into adjoint sources */
    int   adjointmode;           /* mesh-mode J_mua: 0 full FEM form, 1 nodal approximation (cfg->adjointmode) */
    /* per-node optical properties (cfg->nodemua / nodemusp with isnodalmua / isnodalmusp, src/mmc_core.cl:776-793): an element uses
     * the mean of its four nodal values instead of its medium's mua (and mus); NULL = off; nodemusp needs nodemua */
    const float* nodemua;        /* nn */
    const float* nodemusp;       /* nn */
} mmcb_config;

typedef struct mmcb_gpuinfo {    /* src/mmc_utils.h:187-201 GPUInfo */
    char   name[256];
    int    id, devcount;
    int    major, minor;
    size_t globalmem, constmem, sharedmem;
    int    regcount, clock, sm, core;
    size_t autoblock, autothread;
    int    maxgate, maxmpthread;
} mmcb_gpuinfo;

/* Host-side results; every array is caller-allocated (sizes from mmcb_query_sizes) and may be NULL. */
typedef struct mmcb_output {
    double* field;               /* datalen*maxgate*srcnum; ACCUMULATED (+=) like cfg->exportfield, then normalised */
    double* dref;                /* nf*maxgate diffuse reflectance (mesh->dref) or NULL */
    float*  detected;            /* maxdetphoton*reclen rows [detid, nscat[M], ppath[M], (mom[M]), (p[3],v[3]), w0] */
    uint64_t* detseed;           /* maxdetphoton*2 */
    float*  traj;                /* maxjumpdebug*6 */
    /* scalars written by the call */
    unsigned int detectedcount;  /* rows stored in `detected` */
    unsigned int detectedtotal;  /* photons that hit a detector (may exceed maxdetphoton; cfg->his.detected) */
    unsigned int trajcount;
    double energytot[16], energyesc[16];   /* per pattern (cfg->energytot / cfg->energyesc) */
    double raytet;               /* ray-tetrahedron tests (reporter.raytet) */
    double normalizer;           /* cfg->his.normalizer */
    float  kernel_ms;            /* CUDA-event time of the photon kernel(s) */
    int    e0;
    double* field_im;            /* RF forward: imaginary part of the fluence, same layout as `field` (cfg->exportadjoint); may be NULL */
    float*  jacob;               /* adjoint output types: sizes.jacoblen floats (cfg->exportjacob), layout [datalen][Ns*Nd] per component:
                                    CW [J1] | CW dual [J1, J2] | RF [Re J1, Im J1] | RF dual [Re J1, Re J2, Im J1, Im J2]; may be NULL */
    int     overwrite;           /* 0: results are ADDED to field/dref/field_im like the reference adds to cfg->exportfield (default);
                                    1: the arrays are stored (=) without being read: the normalised volume is copied from the device
                                    straight into `field`, which saves a host pass for callers that hand in fresh buffers */
} mmcb_output;

typedef struct mmcb_sizes {
    int maxgate, datalen, reclen, nf, srcnum;   /* nf (exterior faces, dref length nf*maxgate): from mmcb_query_sizes only when
                                                   cfg.issaveref is set (it costs the face-neighbour table); always from a session */
    int dim[3];                  /* dual-grid dimensions (cfg->dim) */
    size_t fieldlen;             /* datalen*maxgate*srcnum*nslots */
    int nslots;                  /* field blocks: extrasrclen in multi-slot mode (srcid < 0), else 1; block stride datalen*maxgate */
    int adj_ns, adj_nd;          /* adjoint output: source and detector slots */
    size_t jacoblen;             /* floats in mmcb_output.jacob (0 unless an adjoint output type is requested) */
} mmcb_sizes;

typedef struct mmcb_session mmcb_session;

/* raw device pointers of a session, for NCCL/peer epilogues run by the caller */
typedef struct mmcb_devptrs {
    void*  field;       size_t fieldlen;  int field_is_double;   /* accumulator volume (elements), element mode for nodal BLB */
    double* energy;     /* [2*16]: tot[16], esc[16] */
    double* raytet;     /* [1] */
    float*  detected;   unsigned int* detcount;  int reclen;
    uint64_t* detseed;
    double* dref;       size_t dreflen;
    void*  field_im;    /* RF runs: imaginary accumulator volume, same length and type as `field`; NULL otherwise */
} mmcb_devptrs;

/* ---- library info ------------------------------------------------------------------------------- */
int         mmcb_version(void);
const char* mmcb_last_error(void);
/* mcx_list_cu_gpu (src/mmc_cu_host.cu:108-198): fills up to `maxcount` entries, returns the device count (>=0) or <0 */
int         mmcb_list_gpu(mmcb_gpuinfo* info, int maxcount);

/* ---- one-call path: what mmc_run_simulation does for one device (src/mmc_cu_host.cu:204-1528) ------ */
int mmcb_query_sizes(const mmcb_config* cfg, const mmcb_mesh* mesh, mmcb_sizes* sizes);
int mmcb_run_simulation(const mmcb_config* cfg, const mmcb_mesh* mesh, int device, mmcb_output* out);

/* ---- session path: upload once, launch many, reduce across GPUs, fetch once ----------------------- */
mmcb_session* mmcb_create(const mmcb_config* cfg, const mmcb_mesh* mesh, int device);
/* replace the session's accumulator volume by caller-owned device memory (e.g. a torch tensor that NCCL will
 * reduce); must hold fieldlen elements of the session's accumulator type, zeroed by the caller */
int  mmcb_set_field_buffer(mmcb_session* s, void* device_ptr);
/* asynchronous launch of `nphoton` photons on `cuda_stream` (NULL = the session's stream); RNG streams are seeded from
 * host rand() words [seed_offset*4*nthread ...) of srand(seed) so that ranks/respins draw disjoint slices;
 * photon ids (replay) start at photon_offset */
int  mmcb_launch(mmcb_session* s, uint64_t nphoton, uint64_t photon_offset, int seed, int seed_offset, void* cuda_stream);
int  mmcb_sync(mmcb_session* s);
int  mmcb_last_kernel_ms(mmcb_session* s, float* ms);
int  mmcb_get_devptrs(mmcb_session* s, mmcb_devptrs* p);
int  mmcb_get_sizes(mmcb_session* s, mmcb_sizes* sizes);
/* D2H + elem->node spreading + mesh_normalize (src/mmc_mesh.c:2154-2279).  energytot/energyesc may be NULL to use the
 * session's own tallies, or point to globally reduced values (multi-GPU). */
int  mmcb_fetch(mmcb_session* s, const double* energytot, const double* energyesc, mmcb_output* out);
int  mmcb_reset(mmcb_session* s);            /* zero all accumulators */
/* diagnostics: the session's device tables -- ne 96-byte tetrahedron records (tracer_build planes + neighbours + flags), ne*4 centroid
 * floats -- and the prepared face-neighbour table (ne*4, exterior faces numbered -1..-nf like tracer_prep); any pointer may be NULL */
int  mmcb_get_tables(mmcb_session* s, void* tetrec_out, float* cent_out, int* facenb_out);
void mmcb_destroy(mmcb_session* s);
/* what mmcb_run_simulation does, on an existing session: all `respin` launches of cfg.nphoton photons, then mmcb_fetch.  With mmcb_create + mmcb_get_sizes in front
 * the caller sizes its output arrays from the session and the mesh is prepared once (mmcb_query_sizes prepares it a second time) */
int mmcb_run_session(mmcb_session* s, mmcb_output* out);

/* ---- photon sharding over the GPUs of one box: the fan-out of mmc_run_cu over cfg->deviceid / cfg->workload
 * (src/mmc_cu_host.cu:403-429,1538-1553).  `devices`: ndev distinct CUDA ordinals (0-based); `workload`: ndev relative weights or NULL
 * (equal shares); device g simulates nphoton * w_g / sum(w) photons (the last one takes the remainder) with its own seeds
 * (srand(seed + 7919 g); replay runs shard the photon index range instead).  One host thread and one session per device, no collective in
 * the walk; afterwards NCCL (libnccl.so.2, loaded on first use) sum-reduces the volume(s) and the diffuse reflectance to devices[0] and
 * gathers the detected-photon rows, seeds and trajectory records behind devices[0]'s own, truncated at maxdetphoton / maxjumpdebug like
 * :823-853; the result is normalised and downloaded once.  ndev == 1 is mmcb_run_simulation.  out->kernel_ms = the slowest device. */
int  mmcb_run_multi(const mmcb_config* cfg, const mmcb_mesh* mesh, int ndev, const int* devices, const float* workload, mmcb_output* out);
/* the photon split mmcb_run_multi uses (exposed for callers that drive the devices themselves, and for tests) */
void mmcb_photon_shares(uint64_t nphoton, int ndev, const float* workload, uint64_t* share);

/* ---- host-side mesh helpers (what mmc_prep/tracer_prep compute; exposed for callers and tests) ------- */
int mmcb_mesh_volumes(int nn, const float* node, int ne, int* elem_inout, const int* type, float* evol, float* nvol); /* mesh_getvolume src/mmc_mesh.c:910 */
int mmcb_mesh_facenb(int ne, const int* elem, int* facenb);                                                           /* mesh_getfacenb src/mmc_highorder.cpp:124 */
int mmcb_mesh_initelem(int nn, const float* node, int ne, const int* elem, const float* srcpos, float* bary4);        /* mesh_initelem src/mmc_mesh.c:1060 */
/* host seeds: srand(seed); rand() x count (src/mmc_cu_host.cu:438,532-534) restated without libc state */
void mmcb_host_seeds(int seed, size_t skip, size_t count, uint32_t* out);

/* RNG known-answer test: stream i (seed words seeds4[4i..4i+3], xorshift128p_seed src/mmc_core.cl:545-548) draws ndraw
 * uniform floats on the DEVICE with the kernel's own generator (xorshift128p_nextf, src/mmc_core.cl:517-532);
 * out[i*ndraw+k], optional final states state_out[2i..2i+1].  Bit-exact to the reference by construction. */
int mmcb_rng_selftest(const uint32_t* seeds4, int nstream, int ndraw, float* out, uint64_t* state_out);

#ifdef __cplusplus
}
#endif
#endif /* MMC_B200_H */
