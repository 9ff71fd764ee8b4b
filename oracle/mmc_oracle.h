/*
 * TEST INFRASTRUCTURE ONLY.  This is the CPU oracle: a plain-C restatement of the
 * reference's (fangq/mmc) photon random walk, used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg as the CHECKER for the CUDA path.  It is never
 * linked into, imported by or called from the product library (mmc_b200/).
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py runs this restatement and the
 * unmodified reference binary (oracle/_ref/mmc_ref, built by oracle/Makefile.ref from
 * /root/reference/src) single-threaded on the same mesh/seed and requires the raw
 * fluence, energy tallies and detected-photon rows to agree; the resulting vectors are
 * committed under tests/golden/ (generator: tools/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/) it restates.
 */
#ifndef MMC_ORACLE_H
#define MMC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ray-tracer ids: src/mmc_utils.h enum TRTMethod */
enum { ORC_PLUCKER = 0, ORC_HAVEL = 1, ORC_BADOUEL = 2, ORC_BLBADOUEL = 3, ORC_BLBADOUEL_GRID = 4 };
/* output types: src/mmc_utils.h enum TOutputType */
enum { ORC_FLUX = 0, ORC_FLUENCE = 1, ORC_ENERGY = 2, ORC_JACOBIAN = 3, ORC_WL = 4, ORC_WP = 5 };
/* boundary conditions: src/mmc_utils.h enum TBoundary */
enum { ORC_BC_NOREFLECT = 0, ORC_BC_REFLECT = 1, ORC_BC_ABSORB_EXTERIOR = 2, ORC_BC_MIRROR = 3 };

#define ORC_SEED_FROM_FILE (-999)

typedef struct orc_mesh {
    int nn, ne, nf, prop;      /* nodes, elements, exterior faces, media (excluding medium 0) */
    int isextdet;              /* 1 if any element was labelled -2 (wide-field detector) */
    float* node;               /* nn*3 */
    int*   elem;               /* ne*4, 1-based; nodes 3,4 swapped where volume<0 (mmc_mesh.c:932-937) */
    int*   type;               /* ne */
    float* med;                /* (prop+1+isextdet)*4: mua mus g n */
    int*   facenb;             /* ne*4; exterior faces numbered -1..-nf (mmc_mesh.c:1466-1474) */
    float* evol;               /* ne */
    float* nvol;               /* nn */
    int*   srcelem; int srcelemlen;
    int*   detelem; int detelemlen;
    float* n;                  /* BLB/Plucker normals, ne*16 column-wise (mmc_mesh.c:1572-1600) */
    float* m;                  /* Havel/Badouel table, ne*48 (mmc_mesh.c:1532-1567) */
    float* pd; float* pm;      /* Plucker edge tables, ne*6*4 each (mmc_mesh.c:1518-1531) */
    float nmin[3], nmax[3];    /* dual-grid bounding box (mmc_mesh.c:349-373) */
    int e0_from_src;           /* first element labelled -1, or 0 */
} orc_mesh;

typedef struct orc_config {
    uint64_t nphoton;
    int seed;                  /* RNG seed, or ORC_SEED_FROM_FILE for replay */
    int nthread;               /* number of RNG streams / OpenMP threads (reference: omp_get_num_threads) */
    float srcpos[4];
    float srcdir[4];           /* w = focal length */
    int   srctype;
    float srcparam1[4], srcparam2[4];
    const float* srcpattern;   /* Nx*Ny*srcnum */
    int   srcnum;
    float tstart, tstep, tend;
    int   e0;                  /* initial element (1-based); 0 => search */
    int   isreflect, isnormalized, issavedet, ismomentum, issaveexit, isspecular, issaveseed, issaveref;
    int   method, basisorder, outputtype;
    float roulettesize, minenergy, nout;
    int   voidtime;
    float unitinmm;
    float steps;               /* dual-grid voxel size */
    int   detnum;
    const float* detpos;       /* detnum*4 (x y z r) */
    unsigned int maxdetphoton;
    /* replay */
    const uint64_t* photonseed;   /* nphoton*2 */
    float* replayweight;          /* nphoton (modified for pattern replay like the reference) */
    const float* replaytime;      /* nphoton */
    /* trajectory debug (dlTraj) */
    int   savetraj; unsigned int maxjumpdebug;
    /* semantic switches between the reference's CPU file and its CUDA kernel */
    int   gpu_semantics;       /* 0: mmc_raytrace.c ; 1: mmc_core.cl where they differ */
} orc_config;

typedef struct orc_result {
    double* field;             /* datalen*maxgate*srcnum, caller-allocated, accumulated (+=) */
    double* dref;              /* nf*maxgate or NULL */
    float*  detected;          /* maxdetphoton*reclen, caller-allocated */
    uint64_t* detseed;         /* maxdetphoton*2 or NULL */
    unsigned int detectedcount;
    float*  traj; unsigned int trajcount;   /* maxjumpdebug*6 */
    double  launchweight[16];  /* per pattern */
    double  absorbweight[16];
    double  escweight[16];     /* GPU-style escaped tally */
    double  raytet;
    double  normalizer;
    int     maxgate, datalen, reclen;
    int     e0;
} orc_result;

orc_mesh* orc_mesh_create(int nn, const float* node, int ne, const int* elem, const int* type,
                          int prop, const float* med, float nout, float unitinmm, const int* facenb_or_null,
                          const float* evol_or_null);
void orc_mesh_free(orc_mesh* m);
void orc_mesh_build_tracer(orc_mesh* m, int method);
int  orc_mesh_initelem(const orc_mesh* m, const float* srcpos, float* bary4);
void orc_mesh_dualgrid(orc_mesh* m, float step, int* dim3, unsigned int* crop3);

int  orc_maxgate(const orc_config* cfg);
int  orc_datalen(const orc_mesh* m, const orc_config* cfg);
int  orc_reclen(const orc_mesh* m, const orc_config* cfg);

/* prepare (tracer_prep) + run (mmc_run_mp) + optional normalisation (mesh_normalize) */
int  orc_run(orc_mesh* m, orc_config* cfg, orc_result* res);

/* RNG known-answer helpers (src/mmc_rand_xorshift128p.c:55-76) */
void orc_rng_seed(const uint32_t seed4[4], uint64_t state[2]);
float orc_rng_nextf(uint64_t state[2]);
void orc_host_seeds(int seed, int count, uint32_t* out);  /* srand(seed); rand() x count */

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
