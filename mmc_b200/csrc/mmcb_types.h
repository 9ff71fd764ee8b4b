// Shared host/device plain-data types of the mmc_b200 engine.
#pragma once
#include <stdint.h>

#define MMCB_MAX_DET      256       // point detectors kept in constant memory
#define MMCB_MAX_SRCNUM   16        // patterns simulated together (photon sharing)
#ifndef MMCB_MAX_TRIAL
#define MMCB_MAX_TRIAL    3         // src/mmc_core.cl:355
#endif
#define MMCB_MAX_STALL    1000      // consecutive zero-length steps before a trapped photon is dropped
#define MMCB_MAX_RELOC    16        // relocations of a photon that is not in its element (mmcb_kernel.cu, no-exit-face path) before it is given up
#define MMCB_DEBUG_REC    6         // floats per trajectory record (src/mmc_core.cl:360)
// Hot-line cache: the L2 serialises atomics that hit one 128-byte line (~0.7 G red/s measured, profiles/), and every
// photon deposits next to the source.  The hottest lines of the accumulator volume (16 doubles / 32 floats worth of
// consecutive accumulators: MMCB_HOT_GROUP entries) are privatised per CTA in shared memory and flushed once.
#define MMCB_HOT_GROUP_LOG2 4       // accumulators per cached group
#define MMCB_HOT_GROUP    (1 << MMCB_HOT_GROUP_LOG2)
#ifndef MMCB_HOT_SLOTS_LOG2
#define MMCB_HOT_SLOTS_LOG2 4       // direct-mapped slots per CTA (16 x 16 floats = 1 KB + 64 B of keys): the few hottest lines are what
#endif                              // matters (measured: 4 to 256 slots run within 3 %, profiles/r1h_tune_hotslots.jsonl)
#define MMCB_HOT_SLOTS    (1 << MMCB_HOT_SLOTS_LOG2)
#define MMCB_HOT_EMPTY    0xFFFFFFFFu
// selection scratch (unsigned int words): [0] bits of the largest group sum, [1] candidate count, [2..33] histogram of the
// exponent distance to the maximum, [34] bits of the total deposited weight (float), [35] 1 when the cache is worth its lookups
#define MMCB_HOT_STAT_WORDS 40
#define MMCB_HOT_STAT_TOTAL 34
#define MMCB_HOT_STAT_USEFUL 35
#define MMCB_HOT_STAT_LO   36     // first accumulator index of the window spanned by the cached groups
#define MMCB_HOT_STAT_SPAN 37     // its length (0: cache off)
#define MMCB_HOT_STAT_FLOOR 38    // float bits: smallest group sum that can make a candidate (count-mode scout, see mmcb_hot_floor_kernel)
#define MMCB_HOT_STAT_STEPS 39    // float bits: the scout's estimate of ray-tet steps per photon over the whole time window (trace output)
#define MMCB_HOT_HASH(g)  (((g) * 0x9E3779B1u) >> (32 - MMCB_HOT_SLOTS_LOG2))

// One tetrahedron = one 96-byte record, 32-byte aligned: three 256-bit gathers (LDG.E.256) bring everything a
// branch-less Badouel step needs -- the reference reads the same data from three arrays (normal[4*eid..],
// facenb[eid], type[eid]; src/mmc_core.cl:747-754,1954,1924) with six or more scattered load instructions.
//   sector 0: nx[4] ny[4]     sector 1: nz[4] d[4]     sector 2: nb[4] type flags pad pad
// Faces are in tracer order j=0..3 (nodes out[j], src/mmc_mesh.c:59); nb[] is ALREADY permuted by faceorder[]
// (src/mmc_mesh.c:84) so nb[j] is the element behind tracer face j; exterior faces hold -(1..nf)
// (src/mmc_mesh.c:1466-1474).
struct __attribute__((aligned(32))) mmcb_tetrec {
    float nx[4], ny[4], nz[4], d[4];
    int   nb[4];
    int   type;
    unsigned int flags;
    int   pad[2];
};
// flags bits (pre-computed per session from the media table, nout and the boundary condition)
#define MMCB_F_REFLECT(j)  (1u << (j))        // crossing face j calls reflectray (src/mmc_core.cl:1957-1958)
#define MMCB_F_TO_VOID(j)  (1u << (4 + (j)))  // this tet has type>0, the neighbour has type 0 (src/mmc_core.cl:1981)
#define MMCB_F_FROM_VOID(j) (1u << (8 + (j))) // this tet has type 0, the neighbour type>0 (src/mmc_core.cl:1970)

// Havel / Plucker kernels, nodal output only: 32-byte companion of the plane record, per tracer face j the reciprocal height of the
// opposite node above the face (a barycentric coordinate is a plane distance times this) and that node's 1-based id as int bits:
//   float invh[4]; int oppnode[4];       (mmcb_build_hpaux_kernel, mmcb_prep.cu)
#define MMCB_HPAUX_FLOATS 8

struct mmcb_kparam {
    // source
    float srcpos[4], srcdir[4], srcparam1[4], srcparam2[4];
    int   srctype, srcnum, srcelemlen, e0;
    float bary0[4];              // barycentric coordinates of srcpos in e0 (cfg->bary0; nodal Havel/Plucker deposit)
    float focus;
    // time gates
    float tstart, tend, Rtstep;
    int   maxgate;
    // physics switches
    int   isreflect, isspecular, voidtime, isextdet, outputtype, method, basisorder;
    float minenergy, roulettesize, nout;
    int   doroulette;            // (tend-tstart)*Rtstep <= 1 (src/mmc_core.cl:2101)
    float roulette_w;            // minenergy when doroulette && minenergy > 0, else -1: `w < roulette_w` is the whole roulette test
    // mesh sizes
    int   nn, ne, nf, maxmedia;
    unsigned int framelen;       // per-gate stride of the accumulator volume
    // dual grid
    float nmin[3]; float dstep;  // dstep = 1/steps.x
    int   segcap;                // dual grid, CAP kernels: deposit segments a lane handles per iteration (2 per voxel edge of path)
    unsigned int crop0[3];
    // detection
    int   issavedet, ismomentum, issaveexit, issaveseed, issaveref, detnum, reclen;
    unsigned int maxdetphoton;
    // replay / debug
    int   isreplay, savetraj;
    unsigned int maxjumpdebug;
    // scheduling
    int   schedule;
    unsigned long long nphoton, photon_offset;
    int   threadphoton, oddphotons;
    int   nmedia;                // entries of the media table (prop+1+isextdet)
    int   hotcache;              // 1: kargs.hotkeys holds MMCB_HOT_SLOTS group keys, deposits to those groups go to shared memory
    unsigned int fieldlen;       // accumulator volume entries (guards the flush of the last, partial group)
    float hotshare;              // the cache is used when the hottest line holds more than this share of the deposited weight
    int   countmode;             // scout launch (general kernels only): deposits count 1 each, photons alive at tend are tallied in kargs.trajcount
    // multi-slot sources (adjoint mode; src/mmc_core.cl:1431-1515) and RF (frequency-domain) forward runs (:872-896,1043-1078)
    int   multisrc;              // 1: photons are launched from kargs.srcdata[] slots
    int   srcid;                 // < 0: every photon picks a slot uniformly (field has one block per slot); > 0: only slot srcid-1
    int   extrasrclen;           // slots in kargs.srcdata
    unsigned int slotstride;     // framelen * maxgate: offset between the slots' blocks of the accumulator volume
    float omega;                 // modulation angular frequency (rad/s); > 0 only in the RF kernel variants
    int   isnodalprop;           // 0 off; 1 kargs.eprop[e].x overrides mua; 2 .y overrides mus as well (src/mmc_core.cl:776-793)
    // launch-element search grid over the candidate elements of wide-field sources (kargs.srccell / kargs.srcitem); dim[0] == 0: none
    float srcgrid_lo[3], srcgrid_inv[3];
    int   srcgrid_dim[3];
};

struct mmcb_kargs {
    const mmcb_tetrec* tet;
    const float*  tetaux;        // [ne][MMCB_HPAUX_FLOATS], Havel / Plucker with basisorder = 1, else NULL
    const float4* cent;          // element centroids (fixphoton, src/mmc_core.cl:1390-1402)
    const float*  node;          // nn*3
    const int*    elem;          // ne*4
    const int*    srcelem;
    const int*    srccell;       // search grid: cell c holds the candidates srcitem[srccell[c] .. srccell[c + 1])
    const int*    srcitem;       // candidate element ids per cell, in the order of srcelem[]
    const float4* med;           // media table (copied to shared memory by each CTA)
    const float*  srcpattern;
    uint32_t* seeds;             // nthread*4: seed words in, stream states out (same packing)
    const unsigned int* hotkeys; // MMCB_HOT_SLOTS group ids (idx >> MMCB_HOT_GROUP_LOG2) or MMCB_HOT_EMPTY
    const unsigned int* hotstat; // MMCB_HOT_STAT_WORDS words of the selection pass
    const unsigned long long* replayseed;
    const float* replayweight;
    const float* replaytime;
    void*   field;               // accumulator volume
    void*   field_im;            // RF: imaginary part, same layout
    const float2* eprop;         // per-element {mua, mus}: means of the four nodal values (per-node optical properties)
    const float4* srcdata;       // multi-slot sources: 4 float4 per slot {srcpos(w: weight), srcdir(w: focal length), srcparam1, srcparam2(w: e0)}
    double* dref;
    float*  detected; unsigned int* detcount; unsigned long long* detseed;
    float*  traj;     unsigned int* trajcount;
    double* energy;              // tot[16], esc[16]
    double* raytet;
    unsigned long long* photon_counter;
};
