// mmc_b200 photon-transport kernels for sm_100a (hand-written; no CPU fallback exists).
//
// What the kernel computes is the reference's per-photon random walk (src/mmc_core.cl:1851-2161 with
// mmc_raytrace.c semantics where the CUDA file has none); HOW it is computed is new:
//   * one flattened, warp-converged state machine: every loop iteration is exactly one ray-tetrahedron
//     step for every lane; a lane whose photon ends re-launches in place.  The reference runs
//     `for photon: onephoton()` so a warp waits for its longest photon (src/mmc_core.cl:2190-2203).
//   * one 96-byte record per tetrahedron fetched with three 256-bit gathers (see mmcb_types.h) instead of
//     >=6 scattered loads from normal[], facenb[] and type[]; face flags make the Fresnel / void tests local.
//   * fire-and-forget `red.global.add` deposits (double or float) -- the reference needs the returned old
//     value for its MAX_ACCUM overflow trick (src/mmc_core.cl:904-912).
//   * persistent warps claim photon ids in chunks from a global counter (work stealing); one xorshift128+
//     stream per thread slot, bit-exact to src/mmc_core.cl:517-532.
//   * detected-photon records are appended with warp-ballot compaction (one atomic per warp and iteration).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "mmcb_types.h"

#ifndef MMCB_ACC_T
#define MMCB_ACC_T double
#endif
typedef MMCB_ACC_T acc_t;

#define MMC_UNDEFINED   3.40282347e+38f
#define EPS             1e-6f                  // src/mmc_mesh.h:70 (CUDA/CPU value)
#define R_C0            3.335640951981520e-12f // src/mmc_core.cl:335
#define R_MIN_MUS       1e9f
#define FIX_PHOTON      1e-3f
#define TWO_PI_D        (3.14159265358979323846 * 2.0)   // src/mmc_mesh.h:69: a double expression
#define TWO_PI_F        6.28318530717958647692f
#define JUST_BELOW_ONE  0.9998f
#define DELTA_MUA       1e-4f
#define POOL_CHUNK      32

__constant__ mmcb_kparam gp;
__constant__ float4 gdet[MMCB_MAX_DET];

// ----------------------------------------------------------------------------------------------------
// RNG: xorshift128+, src/mmc_core.cl:517-561 (bit-exact)
// ----------------------------------------------------------------------------------------------------
struct Rng {
    unsigned long long t0, t1;
};
__device__ __forceinline__ float rand01(Rng& r) {
    unsigned long long s1 = r.t0;
    const unsigned long long s0 = r.t1;
    r.t0 = s0;
    s1 ^= s1 << 23;
    r.t1 = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5);
    unsigned int lo = (unsigned int)(r.t1 + s0);
    return __uint_as_float(0x3F800000U | (lo >> 9)) - 1.0f;
}
__device__ __forceinline__ float rand_scatlen(Rng& r) {
    return -logf(rand01(r) + EPS);
}

// ----------------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const void* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ float sel4(const float* a, int j) {   // register-friendly a[j]
    return j == 0 ? a[0] : (j == 1 ? a[1] : (j == 2 ? a[2] : a[3]));
}
// fire-and-forget reductions on explicitly GLOBAL addresses: atomicAdd() on a pointer loaded from a struct is a generic
// atomic (isspacep branch + shared-memory CAS loop + returning ATOM in SASS); red.global never returns a value
__device__ __forceinline__ void red_add(double* p, float v) {
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(__cvta_generic_to_global(p)), "d"((double)v) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_add_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long atom_add_u64(unsigned long long* p, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_d(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

struct Photon {
    float px, py, pz;
    float vx, vy, vz;
    float w, t, slen, slen0;
    int   eid;                 // 1-based current element
    unsigned int oldidx;       // run-length merged deposit (src/mmc_core.cl:862,898-956)
    float oldw;
    unsigned int posidx;
    unsigned int id;
    int   fixcount;
    unsigned int slotoff;      // multi-slot sources: offset of the photon's slot block in the volume
    unsigned int kdone;        // dual grid, CAP kernels: segments of the current step that earlier iterations already deposited
    float w_im, oldw_im;       // RF variants: imaginary weight and pending imaginary deposit (src/mmc_core.cl:377,383)
    int   oldeid;              // Havel / Plucker nodal deposit: element of the pending run (0: none; oldidx holds its gate / slot offset)
    float nw[4];               //   and its four node sums, in the order of the companion record (node opposite tracer face j)
};

// rotatevector, src/mmc_core.cl:1307-1330
__device__ __forceinline__ void rotatevector(float& vx, float& vy, float& vz, float stheta, float ctheta, float sphi, float cphi) {
    float px, py, pz;

    if (vz > -1.f + EPS && vz < 1.f - EPS) {
        float tmp0 = 1.f - vz * vz;
        float tmp1 = stheta * rsqrtf(tmp0);
        px = tmp1 * (vx * vz * cphi - vy * sphi) + vx * ctheta;
        py = tmp1 * (vy * vz * cphi + vx * sphi) + vy * ctheta;
        pz = -tmp1 * tmp0 * cphi + vz * ctheta;
    } else {
        px = stheta * cphi;
        py = stheta * sphi;
        pz = (vz > 0.f) ? ctheta : -ctheta;
    }

    float inv = rsqrtf(px * px + py * py + pz * pz);
    vx = px * inv;
    vy = py * inv;
    vz = pz * inv;
}

// mc_next_scatter, src/mmc_core.cl:1344-1377
__device__ __forceinline__ float next_scatter(float g, Photon& p, Rng& rng, float& mom) {
    float nextslen = rand_scatlen(rng);
    // the reference multiplies by the double expression TWO_PI (src/mmc_mesh.h:69) and rounds to float; the float product
    // differs from that by at most one ulp of the angle and saves two conversions and a DMUL per scattering event
    float tmp0 = TWO_PI_F * rand01(rng);
    float sphi, cphi, stheta, ctheta;
    sincosf(tmp0, &sphi, &cphi);

    if (g > EPS) {
        tmp0 = (1.f - g * g) / (1.f - g + 2.f * g * rand01(rng));
        tmp0 *= tmp0;
        tmp0 = (1.f + g * g - tmp0) / (2.f * g);
        tmp0 = fmaxf(-1.f, fminf(tmp0, 1.f));
        stheta = sqrtf(1.f - tmp0 * tmp0);
        ctheta = tmp0;
    } else {
        float theta = acosf(2.f * rand01(rng) - 1.f);
        sincosf(theta, &stheta, &ctheta);
    }

    rotatevector(p.vx, p.vy, p.vz, stheta, ctheta, sphi, cphi);
    mom = 1.f - ctheta;
    return nextslen;
}

// enclosing-element search for area sources: src/mmc_core.cl:1786-1827 (candidate list) preceded by a test of the
// current element like the CPU path (src/mmc_raytrace.c:2599-2611).  The reference walks the whole candidate list for every
// photon (9 000 tetrahedra under the 40 x 40 mm pattern of examples/replaywide: 72 s for 1e8 photons in its CUDA kernel); here the
// candidates are binned into a uniform grid at session build, in list order, and only the bin of the launch point is walked:
// the FIRST enclosing candidate in list order is found either way (srcelem[first .. last) is the whole list or one bin).
__device__ __noinline__ int find_launch_elem(float px, float py, float pz, int eid, const int* __restrict__ srcelem, int first, int last,
        const int* __restrict__ elem, const float* __restrict__ node) {
    // everything by value: a reference to the photon or to the kernel argument block would force both into local memory
    for (int is = first - 1; is < last; is++) {
        int cand = (is < first) ? eid : srcelem[is];

        if (cand <= 0) {
            continue;
        }

        const int* ee = elem + 4 * (size_t)(cand - 1);
        const int e0 = ee[0], e1 = ee[1], e2 = ee[2], e3 = ee[3];
        bool include = true;
        #pragma unroll

        for (int i = 0; i < 4; i++) {       // faces out[i] = {0,3,1},{3,2,1},{0,2,3},{0,1,2} (src/mmc_mesh.c:59)
            const int ia = (i == 1) ? e3 : e0;
            const int ib = (i == 0) ? e3 : ((i == 3) ? e1 : e2);
            const int ic = (i < 2) ? e1 : ((i == 2) ? e3 : e2);
            const float* na = node + 3 * (size_t)(ia - 1);
            const float* nb = node + 3 * (size_t)(ib - 1);
            const float* nc = node + 3 * (size_t)(ic - 1);
            float abx = nb[0] - na[0], aby = nb[1] - na[1], abz = nb[2] - na[2];
            float acx = nc[0] - na[0], acy = nc[1] - na[1], acz = nc[2] - na[2];
            float sx = px - na[0], sy = py - na[1], sz = pz - na[2];
            float nx = aby * acz - abz * acy, ny = abz * acx - abx * acz, nz = abx * acy - aby * acx;
            float bary = -(sx * nx + sy * ny + sz * nz);

            if (bary < -1e-4f) {
                include = false;
            }
        }

        if (include) {
            return cand;
        }
    }

    return eid;
}

// launchnewphoton, src/mmc_core.cl:1417-1834 (single-slot sources; multi-slot/adjoint srcdata is out of scope)
template <bool GENERAL>
__device__ __forceinline__ void launch_photon(Photon& p, Rng& rng, const mmcb_kargs& a) {
    p.px = gp.srcpos[0];
    p.py = gp.srcpos[1];
    p.pz = gp.srcpos[2];
    p.vx = gp.srcdir[0];
    p.vy = gp.srcdir[1];
    p.vz = gp.srcdir[2];
    p.eid = gp.e0;
    p.w = 1.f;
    p.t = 0.f;
    p.slen0 = 0.f;
    p.oldidx = 0xFFFFFFFFu;
    p.oldw = 0.f;
    p.oldeid = 0;
    p.nw[0] = p.nw[1] = p.nw[2] = p.nw[3] = 0.f;
    p.posidx = 0;
    p.fixcount = 0;
    p.slotoff = 0;
    p.kdone = 0;
    p.w_im = 0.f;
    p.oldw_im = 0.f;

    if (GENERAL && gp.multisrc) {       // multi-slot sources (adjoint mode), src/mmc_core.cl:1431-1515
        unsigned int slot;

        if (gp.srcid < 0) {             // every photon picks a slot; its deposits go to that slot's block of the volume
            slot = min((unsigned int)(rand01(rng) * gp.extrasrclen), (unsigned int)gp.extrasrclen - 1u);
            p.posidx = slot;
            p.slotoff = slot * gp.slotstride;
        } else {
            slot = (unsigned int)(gp.srcid - 1);
        }

        const float4 sp = __ldg(a.srcdata + 4 * slot), sd = __ldg(a.srcdata + 4 * slot + 1);
        const float radius = __ldg(&a.srcdata[4 * slot + 2].x);
        const int slot_eid = (int)__ldg(&a.srcdata[4 * slot + 3].w);
        p.px = sp.x;
        p.py = sp.y;
        p.pz = sp.z;
        p.vx = sd.x;
        p.vy = sd.y;
        p.vz = sd.z;

        if (radius > 0.f) {             // detector-as-source: uniform disk of the detector radius (:1470-1489)
            float sphi, cphi;
            sincosf(TWO_PI_F * rand01(rng), &sphi, &cphi);
            const float r0 = sqrtf(rand01(rng)) * radius;

            if (sd.z > -1.f + EPS && sd.z < 1.f - EPS) {
                const float tmp0 = 1.f - sd.z * sd.z;
                const float tmp1 = r0 * rsqrtf(tmp0);
                p.px += tmp1 * (sd.x * sd.z * cphi - sd.y * sphi);
                p.py += tmp1 * (sd.y * sd.z * cphi + sd.x * sphi);
                p.pz -= tmp1 * tmp0 * cphi;
            } else {
                p.px += r0 * cphi;
                p.py += r0 * sphi;
            }
        }

        p.w = sp.w;                     // importance weight of the slot
        p.eid = (slot_eid > 0) ? slot_eid : gp.e0;
        p.slen = rand_scatlen(rng);
        return;
    }

    p.slen = rand_scatlen(rng);
    const int st = gp.srctype;

    if (st == 0) {           // pencil :1521-1526
        return;
    }

    float ox = p.px, oy = p.py, oz = p.pz;
    bool canfocus = true;
    const float* sp1 = gp.srcparam1;
    const float* sp2 = gp.srcparam2;

    if (st == 4 || st == 5 || st == 6) {        // planar / pattern / fourier :1531-1562
        float rx = rand01(rng), ry = rand01(rng);
        p.px = gp.srcpos[0] + rx * sp1[0] + ry * sp2[0];
        p.py = gp.srcpos[1] + rx * sp1[1] + ry * sp2[1];
        p.pz = gp.srcpos[2] + rx * sp1[2] + ry * sp2[2];
        p.w = 1.f;

        if (st == 5) {
            int xsize = (int)sp1[3], ysize = (int)sp2[3];
            p.posidx = min((int)(ry * JUST_BELOW_ONE * ysize), ysize - 1) * xsize + min((int)(rx * JUST_BELOW_ONE * xsize), xsize - 1);
            p.w = (gp.srcnum > 1) ? 1.f : a.srcpattern[p.posidx];
        } else if (st == 6) {
            p.w = (cosf((floorf(sp1[3]) * rx + floorf(sp2[3]) * ry + sp1[3] - floorf(sp1[3])) * (float)TWO_PI_D) * (1.f - sp2[3] + floorf(sp2[3])) + 1.f) * 0.5f;
        }

        ox += (sp1[0] + sp2[0]) * 0.5f;
        oy += (sp1[1] + sp2[1]) * 0.5f;
        oz += (sp1[2] + sp2[2]) * 0.5f;
    } else if (st == 9 || st == 10) {           // fourierx / fourierx2d :1566-1591
        float rx = rand01(rng), ry = rand01(rng);
        float tmp = sp1[3] * rsqrtf(sp1[0] * sp1[0] + sp1[1] * sp1[1] + sp1[2] * sp1[2]);
        float v2x = tmp * (gp.srcdir[1] * sp1[2] - gp.srcdir[2] * sp1[1]);
        float v2y = tmp * (gp.srcdir[2] * sp1[0] - gp.srcdir[0] * sp1[2]);
        float v2z = tmp * (gp.srcdir[0] * sp1[1] - gp.srcdir[1] * sp1[0]);
        p.px = gp.srcpos[0] + rx * sp1[0] + ry * v2x;
        p.py = gp.srcpos[1] + rx * sp1[1] + ry * v2y;
        p.pz = gp.srcpos[2] + rx * sp1[2] + ry * v2z;

        if (st == 10) {
            p.w = (sinf((sp2[0] * rx + sp2[2]) * (float)TWO_PI_D) * sinf((sp2[1] * ry + sp2[3]) * (float)TWO_PI_D) + 1.f) * 0.5f;
        } else {
            p.w = (cosf((sp2[0] * rx + sp2[1] * ry + sp2[2]) * (float)TWO_PI_D) * (1.f - sp2[3]) + 1.f) * 0.5f;
        }

        ox += (sp1[0] + v2x) * 0.5f;
        oy += (sp1[1] + v2y) * 0.5f;
        oz += (sp1[2] + v2z) * 0.5f;
    } else if (st == 8 || st == 3) {            // disk / gaussian :1595-1636
        float phi = (float)(TWO_PI_D * rand01(rng));
        float sphi = sinf(phi), cphi = cosf(phi), r0;

        if (st == 8) {
            r0 = sqrtf(rand01(rng)) * sp1[0];
        } else if (fabsf(gp.focus) < 1e-5f || fabsf(sp1[1]) < 1e-5f) {
            r0 = sqrtf(-logf(rand01(rng))) * sp1[0];
        } else {
            float z0 = sp1[0] * sp1[0] * 3.14159265358979f / sp1[1];
            r0 = sqrtf(-logf(rand01(rng)) * (1.f + (gp.focus * gp.focus / (z0 * z0)))) * sp1[0];
        }

        if (gp.srcdir[2] > -1.f + EPS && gp.srcdir[2] < 1.f - EPS) {
            float tmp0 = 1.f - gp.srcdir[2] * gp.srcdir[2];
            float tmp1 = r0 * rsqrtf(tmp0);
            p.px = gp.srcpos[0] + tmp1 * (gp.srcdir[0] * gp.srcdir[2] * cphi - gp.srcdir[1] * sphi);
            p.py = gp.srcpos[1] + tmp1 * (gp.srcdir[1] * gp.srcdir[2] * cphi + gp.srcdir[0] * sphi);
            p.pz = gp.srcpos[2] - tmp1 * tmp0 * cphi;
        } else {
            p.px += r0 * cphi;
            p.py += r0 * sphi;
        }
    } else if (st == 2 || st == 1 || st == 7) { // cone / isotropic / arcsine :1641-1684
        float ang = (float)(TWO_PI_D * rand01(rng));
        float sphi = sinf(ang), cphi = cosf(ang);

        if (st == 2) {
            do {
                ang = (sp1[1] > 0) ? (float)(TWO_PI_D * rand01(rng)) : acosf(2.f * rand01(rng) - 1.f);
            } while (ang > sp1[0]);
        } else if (st == 1) {
            ang = acosf(2.f * rand01(rng) - 1.f);
        } else {
            ang = 3.14159265358979f * rand01(rng);
        }

        float stheta = sinf(ang), ctheta = cosf(ang);
        // direction relative to srcdir like the CPU path (src/mmc_raytrace.c:2479-2481); identical to
        // src/mmc_core.cl:1676-1678 when srcdir=(0,0,1)
        rotatevector(p.vx, p.vy, p.vz, stheta, ctheta, sphi, cphi);
        canfocus = false;

        if (p.eid > 0) {
            return;
        }
    } else if (st == 11) {                      // zgaussian :1689-1701
        float ang = (float)(TWO_PI_D * rand01(rng));
        float sphi = sinf(ang), cphi = cosf(ang);
        ang = sqrtf(-2.f * logf(rand01(rng))) * (1.f - 2.f * rand01(rng)) * sp1[0];
        float stheta = sinf(ang), ctheta = cosf(ang);
        rotatevector(p.vx, p.vy, p.vz, stheta, ctheta, sphi, cphi);
        canfocus = false;
    } else if (st == 12 || st == 13) {          // line / slit :1705-1735
        float t = rand01(rng);
        p.px += t * sp1[0];
        p.py += t * sp1[1];
        p.pz += t * sp1[2];

        if (st == 12) {
            float s, q;
            t = 1.f - 2.f * rand01(rng);
            s = 1.f - 2.f * rand01(rng);
            q = sqrtf(1.f - p.vx * p.vx - p.vy * p.vy) * (rand01(rng) > 0.5f ? 1.f : -1.f);
            float nx = p.vy * q - p.vz * s, ny = p.vz * t - p.vx * q, nz = p.vx * s - p.vy * t;
            p.vx = nx;
            p.vy = ny;
            p.vz = nz;
        }

        ox += sp1[0] * 0.5f;
        oy += sp1[1] * 0.5f;
        oz += sp1[2] * 0.5f;
        canfocus = (st == 13);
    }

    if (canfocus) {                             // :1742-1776
        float f = gp.focus;

        if (isnan(f)) {
            float ang = (float)(TWO_PI_D * rand01(rng)), sphi, cphi, stheta, ctheta;
            sincosf(ang, &sphi, &cphi);
            ang = acosf(2.f * rand01(rng) - 1.f);
            sincosf(ang, &stheta, &ctheta);
            rotatevector(p.vx, p.vy, p.vz, stheta, ctheta, sphi, cphi);
        } else if (f < 0.f && isinf(f)) {
            float ang = (float)(TWO_PI_D * rand01(rng)), sphi, cphi;
            sincosf(ang, &sphi, &cphi);
            float stheta = sqrtf(rand01(rng));
            float ctheta = sqrtf(1.f - stheta * stheta);
            rotatevector(p.vx, p.vy, p.vz, stheta, ctheta, sphi, cphi);
        } else if (f != 0.f) {
            ox += f * p.vx;
            oy += f * p.vy;
            oz += f * p.vz;

            if (f < 0.f) {
                p.vx = p.px - ox;
                p.vy = p.py - oy;
                p.vz = p.pz - oz;
            } else {
                p.vx = ox - p.px;
                p.vy = oy - p.py;
                p.vz = oz - p.pz;
            }

            float rn = rsqrtf(p.vx * p.vx + p.vy * p.vy + p.vz * p.vz);
            p.vx *= rn;
            p.vy *= rn;
            p.vz *= rn;
        }
    }

    p.px += p.vx * EPS;                         // :1778
    p.py += p.vy * EPS;
    p.pz += p.vz * EPS;
    if (gp.srcgrid_dim[0] > 0) {
        const int cx = min(max((int)((p.px - gp.srcgrid_lo[0]) * gp.srcgrid_inv[0]), 0), gp.srcgrid_dim[0] - 1);
        const int cy = min(max((int)((p.py - gp.srcgrid_lo[1]) * gp.srcgrid_inv[1]), 0), gp.srcgrid_dim[1] - 1);
        const int cz = min(max((int)((p.pz - gp.srcgrid_lo[2]) * gp.srcgrid_inv[2]), 0), gp.srcgrid_dim[2] - 1);
        const int cell = (cz * gp.srcgrid_dim[1] + cy) * gp.srcgrid_dim[0] + cx;
        p.eid = find_launch_elem(p.px, p.py, p.pz, p.eid, a.srcitem, __ldg(a.srccell + cell), __ldg(a.srccell + cell + 1), a.elem, a.node);
    } else {
        p.eid = find_launch_elem(p.px, p.py, p.pz, p.eid, a.srcelem, 0, gp.srcelemlen, a.elem, a.node);
    }
}

// Fresnel reflection / refraction, src/mmc_core.cl:1247-1303.  (nx,ny,nz): outward normal of the exit face.
__device__ __forceinline__ void reflectray(Photon& p, int& neweid, float nx, float ny, float nz, float n1,
        const float4* smed, const mmcb_kargs& a, Rng& rng) {
    float Icos = fabsf(p.vx * nx + p.vy * ny + p.vz * nz);
    float n2 = gp.nout;

    if (neweid > 0) {
        int t2 = a.tet[neweid - 1].type;         // rare path: one extra 4-byte gather
        n2 = smed[2 * t2].w;
    }

    float tmp0 = n1 * n1, tmp1 = n2 * n2;
    float tmp2 = 1.f - tmp0 / tmp1 * (1.f - Icos * Icos);

    if (tmp2 > 0.f && !(neweid <= 0 && gp.isreflect == 3)) {
        float Re = tmp0 * Icos * Icos + tmp1 * tmp2;
        tmp2 = sqrtf(tmp2);
        float Im = 2.f * n1 * n2 * Icos * tmp2;
        float Rtotal = (Re - Im) / (Re + Im);
        Re = tmp1 * Icos * Icos + tmp0 * tmp2 * tmp2;
        Rtotal = (Rtotal + (Re - Im) / (Re + Im)) * 0.5f;

        if (rand01(rng) <= Rtotal) {
            p.vx += -2.f * Icos * nx;
            p.vy += -2.f * Icos * ny;
            p.vz += -2.f * Icos * nz;
            neweid = p.eid;
        } else if (gp.isspecular == 2 && neweid == 0) {
        } else {
            float r = n1 / n2;
            p.vx = tmp2 * nx + r * (p.vx - Icos * nx);
            p.vy = tmp2 * ny + r * (p.vy - Icos * ny);
            p.vz = tmp2 * nz + r * (p.vz - Icos * nz);
        }
    } else {
        p.vx += -2.f * Icos * nx;
        p.vy += -2.f * Icos * ny;
        p.vz += -2.f * Icos * nz;
        neweid = p.eid;
    }

    float inv = rsqrtf(p.vx * p.vx + p.vy * p.vy + p.vz * p.vz);
    p.vx *= inv;
    p.vy *= inv;
    p.vz *= inv;
}

// deposit of a merged run (single source or photon-sharing patterns), src/mmc_core.cl:902-946
// Shared-memory layout (dynamic): [hot keys: MMCB_HOT_SLOTS u32][hot sums: SLOTS*GROUP f32] when the hot-line cache is on,
// then the media table (float4 each), then the detected-photon columns.  The cache sits at offset 0 so that its addresses
// are immediates.
#define MMCB_HOT_BYTES (MMCB_HOT_SLOTS * 4 + MMCB_HOT_SLOTS * MMCB_HOT_GROUP * 4)
extern __shared__ float4 smem4[];

__device__ __forceinline__ void red_global(unsigned long long gaddr, float v, double*) {
    // the widening is spelled in PTX: under -use_fast_math the C++ cast becomes cvt.ftz, which costs an extra FMUL.FTZ x1 per deposit
    asm volatile("{\n\t.reg .f64 d;\n\tcvt.f64.f32 d, %1;\n\tred.global.add.f64 [%0], d;\n\t}" :: "l"(gaddr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_global(unsigned long long gaddr, float v, float*) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(gaddr), "f"(v) : "memory");
}

// deposit of a merged run (single source or photon-sharing patterns), src/mmc_core.cl:902-946.  gfield: global-space address
// of the accumulator volume.
template <bool GENERAL>
__device__ __forceinline__ void flush_deposit(unsigned long long gfield, unsigned int idx, float w, const Photon& p, const mmcb_kargs& a, const uint2 hot) {
#ifdef MMCB_COUNT_DEPOSITS      // analysis build (tools/hotspots.py): the volume counts the atomics that reach the L2
    w = 1.f;
#endif

    if (GENERAL && gp.countmode) {  // scout launch: the scratch volume counts deposits (what the L2 serialises per 128-byte line)
        w = 1.f;
    }

    if (!GENERAL || gp.srcnum == 1) {
        // CTA-private sums for the hottest 128-byte lines (see mmcb_types.h).  hot = {first index, span} of the cached groups: one
        // unsigned compare sends every deposit outside that window (other gates, other regions) straight to the volume
        if ((idx - hot.x) < hot.y) {
            const unsigned int* hkeys = (const unsigned int*)smem4;
            float* hvals = (float*)smem4 + MMCB_HOT_SLOTS;
            const unsigned int g = idx >> MMCB_HOT_GROUP_LOG2, h = MMCB_HOT_HASH(g);

            if (hkeys[h] == g) {
#ifdef MMCB_COUNT_DEPOSITS
                w = 0.f;
#endif
                atomicAdd(hvals + (h << MMCB_HOT_GROUP_LOG2) + (idx & (MMCB_HOT_GROUP - 1)), w);   // ATOMS CAS loop, CTA-local
                return;
            }
        }

        red_global(gfield + (unsigned long long)idx * sizeof(acc_t), w, (acc_t*)0);
    } else {
        for (int k = 0; k < gp.srcnum; k++) {
            red_global(gfield + ((unsigned long long)idx * gp.srcnum + k) * sizeof(acc_t), w * a.srcpattern[(size_t)p.posidx * gp.srcnum + k], (acc_t*)0);
        }
    }
}

__device__ __forceinline__ void savedebug(const Photon& p, const mmcb_kargs& a) {  // src/mmc_core.cl:692-704
    unsigned int pos = atom_add_u32(a.trajcount, 1u);

    if (pos < gp.maxjumpdebug) {
        float* d = a.traj + (size_t)pos * MMCB_DEBUG_REC;
        d[0] = __uint_as_float(p.id);
        d[1] = p.px;
        d[2] = p.py;
        d[3] = p.pz;
        d[4] = p.w;
        d[5] = __int_as_float(p.eid);
    }
}


// ----------------------------------------------------------------------------------------------------
// Havel and Plucker ray-tetrahedron steps with the semantics of the reference's CPU file, which is their only
// implementation (src/mmc_raytrace.c:531-800 havel_raytet, :227-508 plucker_raytet):
// ">=" time-window test, one deposit per step (no run-length merge), barycentric nodal deposit for basisorder=1.
//
// Both tracers run on the 96-byte plane record of the BLB kernels.  The reference keeps per-face edge vectors (Havel, 192 B per
// element) or per-edge Plucker coordinates (144 B) and gathers them every step; what those tables decide is a statement about
// the four face planes, and that is how it is evaluated here:
//   Havel  (havel_sse4, :531-561): first face with n.v >= 0, 0 <= t <= 1e10 and the hit inside the triangle (0 <= u, 0 <= v,
//          u + v <= 1).  A triangle of a tetrahedron is its plane cut by the three other face planes, so "inside" is
//          d_j - n_j.(p + t v) >= 0 for the three other faces j -- one FMA each on values the plane tests already produced.  Only
//          the nearest admissible face can pass (a farther hit lies behind the nearest plane), so it is the one tested.
//   Plucker (:283-343): the six edge signs say through which face the LINE leaves the element and whether it meets the element
//          at all; for a convex cell that is: exit face = nearest plane among those with n.v > 0 (negative distances allowed),
//          and the line meets the cell iff the farthest entry plane (n.v < 0) is not behind it.
// Exit barycentric coordinates (nodal deposit): the coordinate of node m is the distance to the face opposite m over the node's
// height above it, so b[opp(j)] = (d_j - n_j.q) * invh_j with invh_j and the node ids in a 32-byte companion record
// (kargs.tetaux, built on the device, read only by nodal runs).  The same expression at the step's start point replaces the
// entry coordinates the reference carries from element to element by matching global node ids (:709-720, :456-480): no
// dependent gather of the neighbour's node ids, no per-photon state.
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool samesign(float a, float b) {     // !((a ^ b) & sign bit), src/mmc_raytrace.c:540-555
    return ((__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u) == 0;
}
__device__ __forceinline__ bool signclear(float a) {
    return (__float_as_uint(a) & 0x80000000u) == 0;
}

// nodal deposit of a closed run: the four node sums of element p.oldeid in gate / slot block p.oldidx (the node ids are re-read from
// the companion record of that element, an L1/L2 hit)
template <bool GENERAL>
__device__ __forceinline__ void flush_nodal(Photon& p, const mmcb_kargs& a, unsigned long long gfield, const uint2 hot) {
    if (p.oldeid > 0) {
        const int4 nd = *(const int4*)(a.tetaux + MMCB_HPAUX_FLOATS * (size_t)(p.oldeid - 1) + 4);
        const int id[4] = {nd.x, nd.y, nd.z, nd.w};
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            if (p.nw[j] != 0.f) {
                flush_deposit<GENERAL>(gfield, (unsigned int)(id[j] - 1) + p.oldidx, p.nw[j], p, a, hot);
            }

            p.nw[j] = 0.f;
        }
    }

    p.oldeid = 0;
}

template <int METHOD, bool GENERAL, bool NODAL>
__device__ __forceinline__ void hp_step(Photon& p, const mmcb_kargs& a, const float4* smed, unsigned long long gfield, const uint2 hot,
                                        bool& found, float& Lmove, bool& isend, bool& timeup, int& neweid, float& fnx, float& fny, float& fnz,
                                        int& type, unsigned& flags, float4& prop) {
    const mmcb_tetrec* rec = a.tet + (p.eid - 1);
    constexpr bool nodal = NODAL;       // gp.basisorder != 0 (kernel variant: the element-wise kernels carry no plane values past the face search)
    float r0[8], r1[8], r2[8];          // nx[4] ny[4] | nz[4] d[4] | nb[4] type flags
    float4 ax4 = make_float4(0.f, 0.f, 0.f, 0.f);              // invh[4] (the node ids behind them are read when a run is written out)
    ld256(rec, r0);
    ld256((const char*)rec + 32, r1);
    ld256((const char*)rec + 64, r2);

    if (nodal) {
        ax4 = __ldg((const float4*)(a.tetaux + MMCB_HPAUX_FLOATS * (size_t)(p.eid - 1)));
    }

    const float ax[4] = {ax4.x, ax4.y, ax4.z, ax4.w};

    type = __float_as_int(r2[4]);
    float S[4], Tn[4], T[4];            // n_j.v, d_j - n_j.p (distance to the plane, inward positive), their ratio
    #pragma unroll

    for (int j = 0; j < 4; j++) {
        S[j] = r0[j] * p.vx + r0[4 + j] * p.vy + r1[j] * p.vz;
        Tn[j] = (-r0[j] * p.px + -r0[4 + j] * p.py) + (-r1[j] * p.pz + r1[4 + j]);
        T[j] = __fdividef(Tn[j], S[j]);
    }

    int fi = -1;                // tracer face 0..3
    float Lp0 = 0.f;

    if constexpr (METHOD == 1) {
        // admissible: det = n.v >= 0 and 0 <= t <= 1e10 (the reference's two sign tests on det and on dett, 1e10 det - dett; NaN fails)
        float tt[4];
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            tt[j] = (S[j] >= 0.f && T[j] >= 0.f && T[j] <= 1e10f) ? T[j] : 3.0e38f;
        }

        const float tmin = fminf(fminf(tt[0], tt[1]), fminf(tt[2], tt[3]));
        bool inside = (tmin < 3.0e38f);
        #pragma unroll

        for (int j = 0; j < 4; j++) {       // the hit face itself (and a face hit at the very same distance: a shared edge) is not a test
            inside = inside && (tt[j] == tmin || fmaf(-tmin, S[j], Tn[j]) >= 0.f);
        }

        if (inside) {
            fi = (tt[0] == tmin) ? 0 : ((tt[1] == tmin) ? 1 : ((tt[2] == tmin) ? 2 : 3));
            Lp0 = tmin;
        }
    } else {
        float to[4], ti[4];
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            to[j] = (S[j] > 0.f) ? T[j] : 3.0e38f;
            ti[j] = (S[j] < 0.f) ? T[j] : -3.0e38f;
        }

        const float tout = fminf(fminf(to[0], to[1]), fminf(to[2], to[3]));
        const float tin = fmaxf(fmaxf(ti[0], ti[1]), fmaxf(ti[2], ti[3]));

        if (tout < 3.0e38f && tin <= tout) {
            fi = (to[0] == tout) ? 0 : ((to[1] == tout) ? 1 : ((to[2] == tout) ? 2 : 3));
            Lp0 = fmaxf(tout, 0.f);     // an origin that drifted past the exit face stays where it is and hops on
        }
    }

    found = (fi >= 0);

    if (!found) {
        return;
    }

    fnx = sel4(r0, fi);         // outward normal of the exit face for reflectray (:2272-2276)
    fny = sel4(r0 + 4, fi);
    fnz = sel4(r1, fi);
    flags = __float_as_uint(r2[5]) >> fi;
    prop = smed[2 * type];
    const float4 pd = smed[2 * type + 1];       // 1/mus (0: none), n/c0, 1/mua (0: mua < EPS), c0/n
    const float mus = prop.y;
    const float dlen = (pd.x == 0.f) ? R_MIN_MUS : p.slen * pd.x;
    isend = (Lp0 > dlen);
    Lmove = isend ? dlen : Lp0;
    neweid = __float_as_int(sel4(r2, fi));
    const float rc = pd.y;

    // common step tail, src/mmc_raytrace.c:357-388 / :633-664 (">=" window test)
    if ((int)((p.t + Lmove * rc - gp.tstart) * gp.Rtstep) >= (int)((gp.tend - gp.tstart) * gp.Rtstep)) {
        timeup = true;
        Lmove = (gp.tend - p.t) * pd.w - 1e-4f;
    }

    float currweight = p.w;
    p.w *= __expf(-prop.x * Lmove);

    if (GENERAL && gp.isreplay) {
        if (gp.outputtype == 3) {               // otJacobian: CPU semantics exp(-DELTA_MUA L) (:1523-1526 equivalent lines)
            currweight = __expf(-DELTA_MUA * Lmove) * a.replayweight[p.id] + p.w;
        } else if (gp.outputtype == 4) {
            currweight = Lmove * a.replayweight[p.id] + p.w;
        }
    }

    p.slen -= Lmove * mus;

    if (GENERAL && gp.isreplay && gp.outputtype == 5) {
        currweight = ((p.slen0 < EPS) ? 1.f : (Lmove * mus / p.slen0)) * a.replayweight[p.id] + p.w;
    }

    const bool fluence = (gp.outputtype != 2 && gp.outputtype != 4 && gp.outputtype != 5);

    if constexpr (METHOD == 1) {
        if (Lp0 == 0.f) {       // :666-668: early break -- no deposit, no clock advance; the photon hops on
            Lmove = 0.f;
            return;
        }
    }

    // a crossing photon continues from the exit point (Plucker: the interpolated pout, src/mmc_raytrace.c:1895 -- the same point)
    p.px += Lmove * p.vx;
    p.py += Lmove * p.vy;
    p.pz += Lmove * p.vz;

    if constexpr (METHOD == 0) {
        if (nodal && !(Lp0 > EPS)) {    // :430: nodal Plucker skips degenerate steps entirely (the position still advances)
            return;
        }
    }

    float ww = currweight - p.w;
    p.t += Lmove * rc;

    if (fluence) {
        ww = (pd.z == 0.f) ? (currweight * Lmove) : (ww * pd.z);
    }

    int gate;

    if (GENERAL && gp.isreplay && (gp.outputtype == 4 || gp.outputtype == 5)) {
        gate = min((int)(a.replaytime[p.id] * gp.Rtstep), gp.maxgate - 1);
    } else {
        gate = min((int)((p.t - gp.tstart) * gp.Rtstep), gp.maxgate - 1);
    }

    const unsigned int tshift = (unsigned int)gate * gp.framelen + (GENERAL ? p.slotoff : 0u);

    if (!nodal) {
        // the reference adds every step to the volume (:413, :753); consecutive steps in one element and gate are summed in a
        // register first (same total, one atomic per visit).  The run is closed when the photon leaves the element, runs out of
        // time, or ends (main loop), so nothing is dropped.
        const unsigned int newidx = (unsigned int)(p.eid - 1) + tshift;

        if (newidx != p.oldidx) {
            if (p.oldw > 0.f) {
                flush_deposit<GENERAL>(gfield, p.oldidx, p.oldw, p, a, hot);
            }

            p.oldidx = newidx;
            p.oldw = 0.f;
        }

        p.oldw += ww;
        return;
    }

    // ---- nodal deposit: w/2 (bary_in + bary_end) to the four nodes; bary_end is the exit point or, when the path ends inside,
    //      the point reached (:709-720,763-780 Havel; :456-480 Plucker).  In plane distances: bary_in[opp(j)] = Tn_j invh_j and
    //      bary_end[opp(j)] = (Tn_j - L S_j) invh_j with L = Lmove (path ends inside) or Lp0 (exit point; exactly 0 on the exit face)
    //      The reference adds the four shares every step (:713-720, :470-480); steps that stay in one element and gate are summed in
    //      registers first and written when the photon leaves the element, changes gate or ends (same totals, fewer atomics).
    if (METHOD == 1 || prop.x > 0.f || fluence) {
        if (p.eid != p.oldeid || tshift != p.oldidx) {
            flush_nodal<GENERAL>(p, a, gfield, hot);
            p.oldeid = p.eid;
            p.oldidx = tshift;
        }

        const float h = ww * 0.5f, Lb = isend ? Lmove : Lp0;
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            const float e = (!isend && j == fi) ? 0.f : fmaf(-Lb, S[j], Tn[j]);
            p.nw[j] += (Tn[j] + e) * ax[j] * h;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// the photon kernel.  METHOD: ray tracer (0 Plucker, 1 Havel, 3 branch-less Badouel, 4 BLB with dual-grid (DMMC) deposit;
// enum TRTMethod, src/mmc_utils.h); DET: detected-photon records;
// GENERAL: area sources, photon sharing, replay, trajectories, diffuse reflectance.
// ----------------------------------------------------------------------------------------------------
#ifndef MMCB_EXP
#define MMCB_EXP(x) __expf(x)
#endif
#ifndef MMCB_GRID_UNROLL
#define MMCB_GRID_UNROLL 2      // segment loop of the dual-grid deposit (measured: profiles/)
#endif
#define MMCB_PRAGMA_(x) _Pragma(#x)
#define MMCB_UNROLL(n) MMCB_PRAGMA_(unroll n)
#ifndef MMCB_MAXTHREADS
#define MMCB_MAXTHREADS 256      // BLB kernels: 4 CTAs x 256 threads per SM at 64 registers (measured against 8 x 128 and 16 x 64:
                                 // +0.6 % sphshells grid, +1.2 % cube60, +3.4 % head-like with detector columns; profiles/r1j_tune_block.jsonl)
#endif
#ifndef MMCB_MINBLOCKS
#define MMCB_MINBLOCKS 4
#endif
#ifndef MMCB_MINBLOCKS_DET
#define MMCB_MINBLOCKS_DET 3     // BLB kernels with detected-photon records: at 4 CTAs per SM (64 registers) they spill 36 bytes inside the loop (ncu: 2.6e9
#endif                           // local-memory sectors per 1e7 photons on the head atlas); 3 CTAs (80 registers, no spill): head atlas 325.7 -> 287.8 ms,
                                 // head-like lattice 288.1 -> 268.8 ms (profiles/r2d_det_occupancy.jsonl)
// Havel / Plucker kernels (plane-record formulation, hp_step): the element-wise plain kernels need 70 registers and, like the BLB
// kernels, run best at 64 with 4 x 256 threads per SM (measured against 6/7/8 x 128 and 3 x 256, profiles/r2h_hp_planes.jsonl); the
// plain nodal kernels run 3 CTAs of 256 (80 registers), the detector and general variants (86-115 registers) 2.
#ifndef MMCB_MAXTHREADS_HP
#define MMCB_MAXTHREADS_HP 256
#endif
#ifndef MMCB_MINBLOCKS_HP
#define MMCB_MINBLOCKS_HP 2
#endif
#ifndef MMCB_MINBLOCKS_HAVEL
#define MMCB_MINBLOCKS_HAVEL 4   // element-wise Havel and Plucker without detector records or general sources
#endif
#ifndef MMCB_MINBLOCKS_HPNODAL
#define MMCB_MINBLOCKS_HPNODAL 3 // nodal Havel and Plucker, plain: 80 registers without spills (2 CTAs: cube60 72.7 ms, 3: 60.4 ms, 4 with 20-40 B spilled: 64.7 ms)
#endif
template <int METHOD, bool DET, bool GENERAL, bool RF = false, bool CAP = false, bool NODAL = false>
__global__ void __launch_bounds__((METHOD <= 1) ? MMCB_MAXTHREADS_HP : MMCB_MAXTHREADS, (METHOD <= 1 && !DET && !GENERAL) ? (NODAL ? MMCB_MINBLOCKS_HPNODAL : MMCB_MINBLOCKS_HAVEL) : ((METHOD <= 1) ? MMCB_MINBLOCKS_HP : (DET ? MMCB_MINBLOCKS_DET : MMCB_MINBLOCKS)))
mmcb_photon_kernel(const mmcb_kargs a) {
    static_assert(!NODAL || METHOD <= 1, "nodal deposit inside the kernel: Havel / Plucker only (the BLB kernels spread elements to nodes afterwards)");
    static_assert(!RF || (GENERAL && METHOD >= 3), "RF forward runs use the general branch-less Badouel kernels");
    static_assert(!CAP || (METHOD == 4 && !RF), "long steps are walked in pieces by the dual-grid kernels only");
    constexpr bool GRID = (METHOD == 4);
    constexpr bool HP = (METHOD <= 1);          // Havel / Plucker: 256-byte records, CPU-file semantics (src/mmc_raytrace.c)
    const bool hoton = gp.hotcache != 0 && a.hotstat[MMCB_HOT_STAT_USEFUL] != 0;
    const uint2 hot = hoton ? make_uint2(a.hotstat[MMCB_HOT_STAT_LO], a.hotstat[MMCB_HOT_STAT_SPAN]) : make_uint2(0u, 0u);
    unsigned int* hkeys = (unsigned int*)smem4;
    float* hvals = (float*)smem4 + MMCB_HOT_SLOTS;
    float4* smed = smem4 + (gp.hotcache ? MMCB_HOT_BYTES / 16 : 0);     // media table, gp.nmedia entries
    float* ppath = (float*)(smed + 2 * gp.nmedia);                  // DET: [reclen][blockDim]
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    acc_t* field = (acc_t*)a.field;
    const unsigned long long gfield = (unsigned long long)__cvta_generic_to_global(a.field);
    const unsigned long long gfield_im = RF ? (unsigned long long)__cvta_generic_to_global(a.field_im) : 0ull;

    for (int i = threadIdx.x; i < 2 * gp.nmedia; i += blockDim.x) {      // {mua mus g n}, {1/mus n/c0 1/mua c0/n} per medium
        smed[i] = a.med[i];
    }

    if (hoton) {
        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS; i += blockDim.x) {
            hkeys[i] = a.hotkeys[i];
        }

        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS * MMCB_HOT_GROUP; i += blockDim.x) {
            hvals[i] = 0.f;
        }
    }

    __syncthreads();

    Rng rng;
    {
        const uint4 s = *(const uint4*)(a.seeds + 4 * (size_t)tid);   // xorshift128p_seed, src/mmc_core.cl:545-548
        rng.t0 = ((unsigned long long)s.x << 32) | s.y;
        rng.t1 = ((unsigned long long)s.z << 32) | s.w;
    }
    Rng initseed = rng;

    Photon p;
    p.eid = 0;
    p.w = 0.f;
    int state = 0;                    // 0: needs a photon, 1: in flight, 2: no photons left
    float etot = 0.f, eesc = 0.f;     // per-thread tallies like src/mmc_core.cl:1908,2155
    unsigned int nraytet = 0;
    // photon ids of this launch are 32-bit offsets (the host splits larger launches).  Dynamic schedule: lane 0 owns the warp's
    // pool [pool_next, pool_end); static schedule: every lane owns its own range (src/mmc_core.cl:2190-2203)
    unsigned int pool_next = 0, pool_end = 0;
    const unsigned int nlaunch = (unsigned int)gp.nphoton;

    if (gp.schedule == 1) {
        pool_next = (unsigned int)tid * (unsigned int)gp.threadphoton + (unsigned int)min(tid, gp.oddphotons);
        pool_end = pool_next + (unsigned int)gp.threadphoton + (tid < gp.oddphotons ? 1u : 0u);
    }

    const int reclen = gp.reclen;
    const int M = gp.maxmedia;
#define PPATH(k) ppath[(k) * blockDim.x + threadIdx.x]
    // detected-photon bookkeeping of the medium the photon is in: path length, scattering count and momentum transfer are summed in
    // registers and written to the shared-memory columns when the medium changes or the photon ends (the reference updates
    // ppath[] in memory every step, src/mmc_core.cl:1943-1945,2128-2134)
    int acct = 0;
    float accL = 0.f, accN = 0.f, accM = 0.f;
#define PPATH_FLUSH() do { if (acct > 0 && acct <= M) { PPATH(M + acct - 1) += accL; PPATH(acct - 1) += accN; if (gp.ismomentum) { PPATH(2 * M + acct - 1) += accM; } } accL = accN = accM = 0.f; } while (0)

    while (true) {
        // ------------------------------------------------------------------ photon supply
        if (!__all_sync(FULL, state == 1)) {       // some lane is out of work: one vote per iteration in the steady state
            unsigned need = __ballot_sync(FULL, state == 0);

            if (need) {
                unsigned int myid = 0;
                bool got = false;

                if (gp.schedule == 1) {
                    if (state == 0 && pool_next < pool_end) {
                        myid = pool_next++;
                        got = true;
                    }
                } else {
                    const int n = __popc(need);
                    unsigned int base = 0;
                    int avail = 0;

                    if (lane == 0 && pool_end - pool_next < (unsigned int)n) {
                        // serve what is left of the old range first (base/avail), then open a fresh chunk
                        base = pool_next;
                        avail = (int)(pool_end - pool_next);
                        const unsigned int want = (unsigned int)(POOL_CHUNK + n - avail);
                        const unsigned long long g0 = atom_add_u64(a.photon_counter, want);
                        pool_next = (unsigned int)min(g0, (unsigned long long)nlaunch);
                        pool_end = max((unsigned int)min(g0 + want, (unsigned long long)nlaunch), pool_next);
                    }

                    // broadcast the pool and distribute: first `avail` needy lanes take the leftover range, the rest the pool
                    avail = __shfl_sync(FULL, avail, 0);
                    base = __shfl_sync(FULL, base, 0);
                    const unsigned int pn = __shfl_sync(FULL, pool_next, 0);
                    const unsigned int pe = __shfl_sync(FULL, pool_end, 0);
                    const int rank = __popc(need & ((1u << lane) - 1));

                    if (state == 0) {
                        if (rank < avail) {
                            myid = base + rank;
                            got = true;
                        } else {
                            const unsigned int cand = pn + (unsigned int)(rank - avail);

                            if (cand < pe) {
                                myid = cand;
                                got = true;
                            }
                        }
                    }

                    if (lane == 0) {
                        pool_next = min(pool_next + (unsigned int)max(0, n - avail), pool_end);
                    }
                }

                if (state == 0) {
                    if (got) {
                        p.id = myid + (unsigned int)gp.photon_offset;

                        if (GENERAL && gp.isreplay) {           // src/mmc_core.cl:2191-2194
                            rng.t0 = a.replayseed[2 * (size_t)p.id];
                            rng.t1 = a.replayseed[2 * (size_t)p.id + 1];
                        }

                        if (DET) {
                            initseed = rng;

                            for (int k = 0; k < reclen; k++) {
                                PPATH(k) = 0.f;
                            }
                        }

                        launch_photon<GENERAL>(p, rng, a);

                        if (DET) {
                            if (!GENERAL || gp.srctype != 5 || gp.srcnum == 1) {
                                PPATH(reclen - 1) = p.w;                       // :1894-1898
                            } else {
                                PPATH(reclen - 1) = __uint_as_float(p.posidx);
                            }
                        }

                        if (!GENERAL || gp.srcnum == 1) {
                            if (HP && GENERAL && gp.isreplay && (gp.outputtype == 4 || gp.outputtype == 5)) {
                                etot += a.replayweight[p.id];       // CPU-file semantics: src/mmc_raytrace.c:1811-1816
                            } else {
                                etot += p.w;
                            }
                        } else {
                            for (int k = 0; k < gp.srcnum; k++) {
                                red_add_d(a.energy + k, (double)(p.w * a.srcpattern[(size_t)p.posidx * gp.srcnum + k]));
                            }
                        }

                        if (GENERAL && gp.savetraj) {
                            savedebug(p, a);
                        }

                        state = 1;
                    } else {
                        state = 2;
                    }
                }
            }

            if (!__any_sync(FULL, state == 1)) {
                break;
            }
        }

        int detid = 0;

        if (state == 1) {
        // ------------------------------------------------------------------ one ray-tetrahedron step
        float Lmove = 0.f, fnx = 0.f, fny = 0.f, fnz = 0.f;
        bool capped = false;            // dual grid, CAP kernels: the step has segments left for the next iteration (see gp.segcap)
        float4 prop = make_float4(0.f, 0.f, 0.f, 1.f);
        int neweid = 0, type = 0, faceidx = 0;
        unsigned flags = 0;
        bool found = false, isend = false, timeup = false;
        bool terminate = false, detect = false;
        int exiteid = p.eid;          // value of r.eid at termination (<=0: left the mesh)

        if constexpr (!HP) {
        const mmcb_tetrec* rec = a.tet + (p.eid - 1);
        float r0[8], r1[8], r2[8];
        ld256(rec, r0);                                     // nx[4] ny[4]
        ld256((const char*)rec + 32, r1);                   // nz[4] d[4]
        ld256((const char*)rec + 64, r2);                   // nb[4] type flags
        // src/mmc_core.cl:752-771: T_j = (d_j - N_j.p) / (N_j.v) for faces with N_j.v > 0, else 1e10; the exit face is the first
        // minimum.  Only the neighbour id is selected here; the outward normal is needed by reflectray alone (index-mismatch
        // faces) and is re-read there from the record, which is an L1 hit, so the 24 record registers die before the deposit code.
        float T[4];
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            const float S = p.vx * r0[j] + p.vy * r0[4 + j] + p.vz * r1[j];
            const float Tn = r1[4 + j] - (p.px * r0[j] + p.py * r0[4 + j] + p.pz * r1[j]);
            T[j] = (S > 0.f) ? __fdividef(Tn, S) : 1e10f;
        }

        const float Lmin = fminf(fminf(T[0], T[1]), fminf(T[2], T[3]));
        faceidx = (T[0] == Lmin) ? 0 : ((T[1] == Lmin) ? 1 : ((T[2] == Lmin) ? 2 : 3));
        neweid = __float_as_int((faceidx == 0) ? r2[0] : ((faceidx == 1) ? r2[1] : ((faceidx == 2) ? r2[2] : r2[3])));
        type = __float_as_int(r2[4]);
        flags = __float_as_uint(r2[5]) >> faceidx;      // bit 0: reflect, bit 4: to void, bit 8: from void
        found = (Lmin < 1e10f && Lmin >= 0.f);

        if (found) {
            prop = smed[2 * type];                          // mua mus g n
            float4 pd = smed[2 * type + 1];                 // 1/mus (0: none), n/c0, 1/mua (0: mua < EPS), c0/n

            if (GENERAL && gp.isnodalprop) {                // per-node optical properties: element means (:776-793)
                const float2 ep = __ldg(a.eprop + (p.eid - 1));
                prop.x = ep.x;
                pd.z = (ep.x < EPS) ? 0.f : __frcp_rn(ep.x);

                if (gp.isnodalprop > 1) {
                    prop.y = ep.y;
                    pd.x = (ep.y <= EPS) ? 0.f : __frcp_rn(ep.y);
                }
            }

            Lmove = (pd.x == 0.f) ? R_MIN_MUS : p.slen * pd.x;
            isend = (Lmin > Lmove);
            Lmove = isend ? Lmove : Lmin;

            const float rc = pd.y;
            float tnew = p.t + Lmove * rc;
            int gate = (int)((tnew - gp.tstart) * gp.Rtstep);

            if (gate > gp.maxgate - 1) {                    // :803-807
                timeup = true;
                Lmove = (gp.tend - p.t) * pd.w - 1e-4f;
                tnew = p.t + Lmove * rc;
                gate = min((int)((tnew - gp.tstart) * gp.Rtstep), gp.maxgate - 1);
            }

            float currweight = p.w;
            // CAP kernels walk the segments of a long step over several iterations (below): the step is recomputed from the untouched
            // photon every time and committed when its last segment is in
            const float w_before = p.w, slen_before = p.slen, t_before = p.t;
            float totalloss = MMCB_EXP(-prop.x * Lmove);
            p.w *= totalloss;
            totalloss = 1.f - totalloss;

            if (GENERAL && gp.isreplay) {                   // :814-829
                if (gp.outputtype == 4 || gp.outputtype == 3) {
                    currweight = Lmove * a.replayweight[p.id] + p.w;
                } else if (gp.outputtype == 5) {
                    currweight = ((p.slen0 < EPS) ? 1.f : (Lmove * prop.y / p.slen0)) * a.replayweight[p.id] + p.w;
                }
            }

            p.slen -= Lmove * prop.y;
            float ww = currweight - p.w;
            p.t = tnew;

            if (GENERAL && gp.isreplay && (gp.outputtype == 4 || gp.outputtype == 5)) {
                gate = min((int)(a.replaytime[p.id] * gp.Rtstep), gp.maxgate - 1);
            }

            const unsigned int tshift = (unsigned int)gate * gp.framelen + (GENERAL ? p.slotoff : 0u);

            if (gp.outputtype != 2 && gp.outputtype != 4 && gp.outputtype != 5) {       // :844-851
                ww = (pd.z == 0.f) ? (currweight * Lmove) : (ww * pd.z);
            }

            const bool flushnow = timeup || !isend;
            // RF forward (omega > 0): the weight is complex, w *= exp(-(mua + i omega n/c0) L), and a step deposits
            // (w0 - w1) / (mua + i omega n/c0)  (src/mmc_core.cl:872-896 mesh, :1043-1078 per dual-grid segment)
            const float a_im = RF ? gp.omega * rc : 0.f;
            const float a_mag2 = prop.x * prop.x + a_im * a_im;
            float ww_im = 0.f;

            if (!GRID) {                                    // :856-1010: run-length merge of deposits into one accumulator
                const unsigned int newidx = (unsigned int)(p.eid - 1) + tshift;

                if constexpr (RF) {
                    const float att = (totalloss < 1.f) ? (1.f - totalloss) : 1.f;      // exp(-mua L)
                    float sph, cph;
                    __sincosf(a_im * Lmove, &sph, &cph);
                    const float w0r = currweight, w0i = p.w_im;
                    const float nr = att * (w0r * cph + w0i * sph), ni = att * (-w0r * sph + w0i * cph);
                    const float dr = w0r - nr, di = w0i - ni;
                    ww = (a_mag2 > 0.f) ? __fdividef(dr * prop.x + di * a_im, a_mag2) : (w0r * Lmove);
                    ww_im = (a_mag2 > 0.f) ? __fdividef(di * prop.x - dr * a_im, a_mag2) : (w0i * Lmove);
                    p.w = nr;
                    p.w_im = ni;
                }

                #pragma unroll

                for (int k = 0; k < 2; k++) {               // k == 1 is the closing flush (one deposit site)
                    const unsigned int idx = (k == 0) ? newidx : (flushnow ? 0xFFFFFFFFu : newidx);

                    if (idx != p.oldidx) {
                        // RF: the real part of a merged run may be <= 0; like the reference, a run that ends because the
                        // element changed is dropped then (:907), the closing flush is unconditional (:961-984)
                        if (RF ? (p.oldidx != 0xFFFFFFFFu && (k == 1 || p.oldw > 0.f)) : (p.oldw > 0.f)) {
                            flush_deposit<GENERAL>(gfield, p.oldidx, p.oldw, p, a, hot);

                            if constexpr (RF) {
                                red_global(gfield_im + (unsigned long long)p.oldidx * sizeof(acc_t), p.oldw_im, (acc_t*)0);
                            }
                        }

                        p.oldidx = idx;
                        p.oldw = 0.f;
                        p.oldw_im = 0.f;
                    }

                    if (k == 0) {
                        p.oldw += ww;
                        p.oldw_im += ww_im;
                    }
                }
            } else {                                        // dual-grid deposit :1022-1206
                int seg = ((int)(Lmove * gp.dstep) + 1) << 1;
                float seglen = Lmove / seg;
                float segdecay = MMCB_EXP(-prop.x * seglen);
                // segment midpoints in voxel units: g = ((p - nmin) + (k + 1/2) v seglen) / step, index = max(floor(g), 0) like the
                // reference's `(S.x > 0) ? __float2int_rd(S.x * dstep) : 0` (:1056-1058)
                const float vs = seglen * gp.dstep;
                float dx = p.vx * vs, dy = p.vy * vs, dz = p.vz * vs;
                float sx = (p.px - gp.nmin[0]) * gp.dstep + dx * 0.5f, sy = (p.py - gp.nmin[1]) * gp.dstep + dy * 0.5f,
                      sz = (p.pz - gp.nmin[2]) * gp.dstep + dz * 0.5f;
                float frac = (totalloss == 0.f) ? 0.f : (1.f - segdecay) / totalloss;
                float segw = ww;
                float seg_re = currweight, seg_im = p.w_im, dsn = 0.f, dcs = 1.f;

                if constexpr (RF) {
                    __sincosf(a_im * seglen, &dsn, &dcs);
                }

                // consecutive segments in one voxel are merged before they reach the volume (the reference issues one atomic
                // per segment once the photon is about to leave the element, src/mmc_core.cl:1150-1206): same sums, fewer
                // atomics.  The loop body is one segment; the run that is still open when the photon leaves the element (or runs
                // out of time) is closed after the loop.
                const unsigned int cx = gp.crop0[0], cy = gp.crop0[1];
                int k0 = 0, k1 = seg;

                if constexpr (CAP) {
                    // A warp runs this loop as long as its longest lane needs (skinvessel, 5 um voxels: 7.7 of 32 lanes active).  Here a
                    // lane deposits at most gp.segcap segments per iteration; a step with more of them is NOT committed: the photon keeps
                    // its pre-step state, remembers how many segments are done (p.kdone) and the next iteration recomputes the same step
                    // (the warp executes that code for its other lanes anyway) and continues with segment kdone.  The segments, their
                    // midpoints and weights are exactly the reference's 2 (int(L / voxel) + 1) equal pieces (src/mmc_core.cl:1024-1078).
                    k0 = (int)p.kdone;
                    k1 = min(seg, k0 + gp.segcap);
                    capped = (k1 < seg);

                    if (k0 > 0) {
                        const float fk = (float)k0;
                        sx += fk * dx;
                        sy += fk * dy;
                        sz += fk * dz;
                        segw *= __expf(-prop.x * seglen * fk);
                    }
                }

                MMCB_UNROLL(MMCB_GRID_UNROLL)

                for (int k = k0; k < k1; k++) {
                    const int ix = max(__float2int_rd(sx), 0), iy = max(__float2int_rd(sy), 0), iz = max(__float2int_rd(sz), 0);
                    const unsigned int newidx = (unsigned int)iz * cy + (unsigned int)iy * cx + (unsigned int)ix + tshift;

                    if (newidx != p.oldidx) {
                        if (RF ? (p.oldidx != 0xFFFFFFFFu) : (p.oldw > 0.f)) {      // RF: ungated like :1084-1112
                            flush_deposit<GENERAL>(gfield, p.oldidx, p.oldw, p, a, hot);

                            if constexpr (RF) {
                                red_global(gfield_im + (unsigned long long)p.oldidx * sizeof(acc_t), p.oldw_im, (acc_t*)0);
                            }
                        }

                        p.oldidx = newidx;
                        p.oldw = 0.f;
                        p.oldw_im = 0.f;
                    }

                    if constexpr (RF) {
                        const float w0r = seg_re, w0i = seg_im;
                        seg_re = segdecay * (w0r * dcs + w0i * dsn);
                        seg_im = segdecay * (-w0r * dsn + w0i * dcs);
                        const float dr = w0r - seg_re, di = w0i - seg_im;
                        p.oldw += (a_mag2 < EPS) ? (w0r * segw) : __fdividef(dr * prop.x + di * a_im, a_mag2);
                        p.oldw_im += (a_mag2 < EPS) ? (w0i * segw) : __fdividef(di * prop.x - dr * a_im, a_mag2);
                    } else {
                        p.oldw += segw * frac;
                    }

                    segw *= segdecay;
                    sx += dx;
                    sy += dy;
                    sz += dz;
                }

                if (CAP && capped) {            // more segments to come: the step stays uncommitted
                    p.kdone = (unsigned int)k1;
                    p.w = w_before;
                    p.slen = slen_before;
                    p.t = t_before;
                } else if (flushnow) {
                    if (RF ? (p.oldidx != 0xFFFFFFFFu) : (p.oldw > 0.f)) {
                        flush_deposit<GENERAL>(gfield, p.oldidx, p.oldw, p, a, hot);

                        if constexpr (RF) {
                            red_global(gfield_im + (unsigned long long)p.oldidx * sizeof(acc_t), p.oldw_im, (acc_t*)0);
                        }
                    }

                    p.oldidx = 0xFFFFFFFFu;
                    p.oldw = 0.f;
                    p.oldw_im = 0.f;
                }

                if constexpr (CAP) {
                    if (!capped) {
                        p.kdone = 0;
                    }
                }

                if constexpr (RF) {                         // :1209-1212
                    p.w = seg_re;
                    p.w_im = seg_im;
                }
            }

        }   // found
        } else {
            hp_step<METHOD, GENERAL, NODAL>(p, a, smed, gfield, hot, found, Lmove, isend, timeup, neweid, fnx, fny, fnz, type, flags, prop);
        }

        nraytet += (CAP && capped) ? 0u : 1u;

        if (found && !(CAP && capped)) {
            if constexpr (!HP) {
                p.px += Lmove * p.vx;                       // :1222
                p.py += Lmove * p.vy;
                p.pz += Lmove * p.vz;
            }

            // progress guard (not in the reference): a photon that makes no headway for MMCB_MAX_STALL consecutive steps is
            // trapped between degenerate/inverted tetrahedra (the reference CPU path spins forever there) and is dropped
            // fixcount: bits 0-7 = failed exit-face searches since the last step with headway (the reference's `fixcount`), bits 8-14 =
            // relocations of this photon (below; never cleared), bit 15 = it has made headway at least once, bits 16+ = consecutive
            // steps without headway
            p.fixcount = (Lmove > 0.f) ? ((p.fixcount & 0x7F00) | 0x8000) : (p.fixcount + 0x10000);

            if (DET) {                                      // :1943-1945
                if (type != acct) {
                    PPATH_FLUSH();
                    acct = type;
                }

                accL += Lmove;
            }

            if (timeup || p.fixcount >= (MMCB_MAX_STALL << 16)) {
                terminate = true;                           // :1928-1930 / :2007-2009 (photon stays inside: no detection)
            } else if (!isend) {
                // ---- cross the face: neighbour hop + boundary physics :1950-1990
                // r.p0 = r.pout: already there, Lmove == Lmin on this branch
                if (gp.isreflect && (flags & 1u)) {
                    if constexpr (!HP) {                    // outward normal of the exit face: record sectors 0/1, L1-resident
                        const mmcb_tetrec* rec = a.tet + (p.eid - 1);
                        fnx = __ldg(rec->nx + faceidx);
                        fny = __ldg(rec->ny + faceidx);
                        fnz = __ldg(rec->nz + faceidx);
                    }

                    reflectray(p, neweid, fnx, fny, fnz, prop.w, smed, a, rng);
                }

                if (neweid <= 0) {
                    terminate = true;
                    detect = true;
                    exiteid = neweid;
                } else if (neweid != p.eid) {
                    if ((flags & 0x100u) && !gp.voidtime) {
                        p.t = 0.f;                          // :1970-1978
                    }

                    if ((flags & 0x10u) && !gp.isextdet) {
                        terminate = true;                   // :1981-1990 (r.eid = 0)
                        detect = true;
                        exiteid = 0;
                    } else {
                        p.eid = neweid;     // (pulling the neighbour's record into L1 here -- prefetch.global.L1 in round 1, three LDGSTS.ca copies
                    }                       // into a shared-memory sink in round 2 -- costs more than it hides: profiles/r2d_negative_results.jsonl)
                }
            } else {
                // ---- end of the scattering path: roulette :2101-2114, then a new direction :2117-2135
                bool dead = false;

                if ((RF ? sqrtf(p.w * p.w + p.w_im * p.w_im) : p.w) < gp.roulette_w) {  // roulette_w = minenergy when roulette applies, else -1
                    if (rand01(rng) * gp.roulettesize <= 1.f) {
                        p.w *= gp.roulettesize;

                        if constexpr (RF) {
                            p.w_im *= gp.roulettesize;      // :2105-2107
                        }

                    } else {
                        dead = true;
                    }
                }

                if (dead) {
                    terminate = true;
                } else {
                    float mom;
                    p.slen0 = next_scatter(prop.z, p, rng, mom);
                    p.slen = p.slen0;

                    if (GENERAL && gp.savetraj) {
                        savedebug(p, a);
                    }

                    if (DET) {              // type == acct here: the step above set it
                        accM += mom;
                        accN += 1.f;
                    }
                }
            }
        } else if (!found) {
            // no exit face found: pull the photon towards the centroid and retry (:1932-1935, :2013-2024)
#ifdef MMCB_COUNT_FIX           // analysis build: not-found events in the low 20 bits of kargs.trajcount, dropped photons above
            atom_add_u32(a.trajcount, ((p.fixcount & 0xFF) < MMCB_MAX_TRIAL) ? 1u : (1u << 20));
#endif

            if ((p.fixcount++ & 0xFF) < MMCB_MAX_TRIAL) {
                float4 c = a.cent[p.eid - 1];
                p.px += (c.x - p.px) * FIX_PHOTON;
                p.py += (c.y - p.py) * FIX_PHOTON;
                p.pz += (c.z - p.pz) * FIX_PHOTON;
            } else {
                // The pulls did not help: the photon is not in this element.  It happens behind degenerate elements (the shipped sphshells mesh
                // has one flat tetrahedron of 1e-15 mm^3 and 5 mm^2: its four plane distances are rounding noise, so the exit face -- and with
                // it the neighbour -- is a coin toss, and the wrong neighbour does not contain the photon by 0.1 mm).  The reference gives such
                // a photon up (r.eid = ID_UNDEFINED, :2027-2031); measured on config C2 that rule cost this engine 3.5e-4 of its photons, a
                // deficit that grew to 0.5 % in the last time gate against BOTH reference programs (profiles/r2n_gate_drift.txt).  Instead the
                // photon is relocated: it steps, without moving, through the face it is farthest beyond, until an element holds it -- at most
                // MMCB_MAX_RELOC times in its life (a photon that cannot be placed is given up as before; an unbounded search can cycle
                // between elements with minute steps of headway).  Leaving the mesh this way ends the photon.
                const mmcb_tetrec* rr = a.tet + (p.eid - 1);
                float worst = 0.f;
                int nbw = 0;
                #pragma unroll

                for (int j = 0; j < 4; j++) {
                    const float tn = __ldg(rr->d + j) - (p.px * __ldg(rr->nx + j) + p.py * __ldg(rr->ny + j) + p.pz * __ldg(rr->nz + j));

                    if (tn < worst) {
                        worst = tn;
                        nbw = __ldg(rr->nb + j);
                    }
                }

                // Plain kernels only, and only photons that have moved: area sources and source slots (general kernels) launch photons
                // outside their launch element, and the reference's results (adjoint Jacobians, tests/test_gpu_vs_reference_adjoint.py:
                // +46 % when a build relocated them and cleared the trial counter at zero-length steps) count on those being given up.
                const int nreloc = (p.fixcount >> 8) & 0x7F;

                if (!GENERAL && nbw > 0 && nreloc < MMCB_MAX_RELOC && (p.fixcount & 0x8000)) {
                    p.eid = nbw;
                    p.fixcount = (p.fixcount & ~0x7FFF) | ((nreloc + 1) << 8);      // searches start over in the new element
                } else {
                    terminate = true;                       // r.eid = ID_UNDEFINED: dropped without detection
                }
            }
        }

        // ------------------------------------------------------------------ photon end: tallies + detection
        if (terminate) {
            if (DET) {
                PPATH_FLUSH();
                acct = 0;
            }

            if (detect) {
                if (GENERAL && gp.issaveref && exiteid < 0 && a.dref) {     // src/mmc_raytrace.c:2000-2003
                    int g = min((int)((p.t - gp.tstart) * gp.Rtstep), gp.maxgate - 1);
                    red_add_d(a.dref + ((size_t)g * gp.nf + (-exiteid - 1)), (double)p.w);
                }

                if (DET) {                                  // finddetector :608-621 / wide-field :2072
                    if (gp.isextdet && type == M + 1) {
                        detid = p.eid;
                    } else {
                        for (int i = 0; i < gp.detnum; i++) {
                            float4 dp = gdet[i];
                            float ddx = dp.x - p.px, ddy = dp.y - p.py, ddz = dp.z - p.pz;

                            if (ddx * ddx + ddy * ddy + ddz * ddz < dp.w * dp.w) {
                                detid = i + 1;
                                break;
                            }
                        }
                    }
                }
            }

            if (GENERAL && gp.savetraj) {
                savedebug(p, a);
            }

            if (GENERAL && gp.countmode && timeup) {        // scout: photons that outlive its (first-gate) window
                atom_add_u32(a.trajcount, 1u);
            }

            if (!GENERAL || gp.srcnum == 1) {
                eesc += RF ? sqrtf(p.w * p.w + p.w_im * p.w_im) : p.w;     // RF: |w| (:2150-2152)
            } else {
                for (int k = 0; k < gp.srcnum; k++) {
                    red_add_d(a.energy + MMCB_MAX_SRCNUM + k, (double)(p.w * a.srcpattern[(size_t)p.posidx * gp.srcnum + k]));
                }
            }

            // a merged deposit may still be pending (photon died at a scattering site): the reference GPU kernel drops it only
            // when the run ended on `isend` -- it flushes on the NEXT step, which never comes; we keep that behaviour.  The
            // Havel/Plucker kernels follow the CPU file, which deposits every step: their pending run is written out.
            if constexpr (HP) {
                if constexpr (NODAL) {
                    flush_nodal<GENERAL>(p, a, gfield, hot);
                } else if (p.oldw > 0.f) {
                    flush_deposit<GENERAL>(gfield, p.oldidx, p.oldw, p, a, hot);
                }

                p.oldidx = 0xFFFFFFFFu;
                p.oldw = 0.f;
            }

            state = 0;
        }
        }   // state == 1

        if (DET) {
            unsigned detmask = __ballot_sync(FULL, detid != 0);

            if (detmask) {                                  // warp-ballot compaction of savedetphoton (:624-682)
                unsigned int base = 0;

                if (lane == __ffs(detmask) - 1) {
                    base = atom_add_u32(a.detcount, (unsigned int)__popc(detmask));
                }

                base = __shfl_sync(FULL, base, __ffs(detmask) - 1);

                if (detid) {
                    unsigned int slot = base + __popc(detmask & ((1u << lane) - 1));

                    if (slot < gp.maxdetphoton) {
                        float* out = a.detected + (size_t)slot * (reclen + 1);

                        if (gp.issaveexit) {
                            PPATH(reclen - 7) = p.px;
                            PPATH(reclen - 6) = p.py;
                            PPATH(reclen - 5) = p.pz;
                            PPATH(reclen - 4) = p.vx;
                            PPATH(reclen - 3) = p.vy;
                            PPATH(reclen - 2) = p.vz;
                        }

                        out[0] = (float)((GENERAL && gp.multisrc && gp.srcid <= 0) ? ((unsigned int)detid | ((p.posidx + 1u) << 16)) : (unsigned int)detid);   // :652-658

                        for (int k = 0; k < reclen; k++) {
                            out[1 + k] = PPATH(k);
                        }

                        if (gp.issaveseed) {
                            a.detseed[2 * (size_t)slot] = initseed.t0;
                            a.detseed[2 * (size_t)slot + 1] = initseed.t1;
                        }
                    }
                }
            }
        }
    }

#undef PPATH_FLUSH
#undef PPATH
    // the stream state goes back in the seed-word packing: the next launch of the session may continue the streams
    *(uint4*)(a.seeds + 4 * (size_t)tid) = make_uint4((unsigned int)(rng.t0 >> 32), (unsigned int)rng.t0, (unsigned int)(rng.t1 >> 32), (unsigned int)rng.t1);

    if (hoton) {            // flush the CTA-private sums of the hot lines
        __syncthreads();

        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS * MMCB_HOT_GROUP; i += blockDim.x) {
            const unsigned int g = hkeys[i >> MMCB_HOT_GROUP_LOG2];
            const float v = hvals[i];
            const unsigned int idx = (g << MMCB_HOT_GROUP_LOG2) + (i & (MMCB_HOT_GROUP - 1));

            if (g != MMCB_HOT_EMPTY && v != 0.f && idx < gp.fieldlen) {
                red_add(field + idx, v);
            }
        }
    }

    // ---------------------------------------------------------------------- per-warp reduction of the tallies
    double dt = etot, de = eesc, dr = (double)nraytet;
    #pragma unroll

    for (int o = 16; o > 0; o >>= 1) {
        dt += __shfl_xor_sync(FULL, dt, o);
        de += __shfl_xor_sync(FULL, de, o);
        dr += __shfl_xor_sync(FULL, dr, o);
    }

    if (lane == 0) {
        if (!GENERAL || gp.srcnum == 1) {
            red_add_d(a.energy, dt);
            red_add_d(a.energy + MMCB_MAX_SRCNUM, de);
        }

        red_add_d(a.raytet, dr);
    }
}

#include "mmcb_kernel_rp.cuh"

// elem -> node spreading for nodal output with the BLB tracer (the reference does this on the host,
// src/mmc_cu_host.cu:929-975): node += 0.25 * elem for the 4 nodes of each element, per gate and pattern.
__global__ void mmcb_spread_nodes_kernel(const acc_t* __restrict__ efield, double* __restrict__ nfield, const int* __restrict__ elem,
        int ne, int nn, int maxgate, int srcnum) {
    size_t total = (size_t)ne * maxgate * srcnum;

    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int s = (int)(i % srcnum);
        size_t r = i / srcnum;
        int e = (int)(r % ne);
        int g = (int)(r / ne);
        double w = (double)efield[i] * 0.25;

        if (w != 0.0) {
            const int* ee = elem + 4 * (size_t)e;
            #pragma unroll

            for (int k = 0; k < 4; k++) {
                red_add_d(nfield + ((size_t)g * nn + (ee[k] - 1)) * srcnum + s, w);
            }
        }
    }
}

__global__ void mmcb_acc_to_double_kernel(const acc_t* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = (double)in[i];
    }
}

// ----------------------------------------------------------------------------------------------------
// hot-line selection from the pilot batch (see mmcb_types.h): group = MMCB_HOT_GROUP consecutive accumulators.
// 1) max group sum, 2) histogram of exponent distance to the max, 3) candidates down to the bin that still fits,
// 4) direct-mapped insertion, hottest bins first.  Everything stays on the device: no host synchronisation.
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hot_group_sum(const acc_t* __restrict__ field, size_t g, size_t fieldlen) {
    float s = 0.f;
    const size_t i0 = g << MMCB_HOT_GROUP_LOG2;
    #pragma unroll

    for (int j = 0; j < MMCB_HOT_GROUP; j++) {
        if (i0 + j < fieldlen) {
            s += (float)field[i0 + j];
        }
    }

    return s;
}

__device__ __forceinline__ int hot_bin(float s, unsigned int maxbits) {
    return min(31, (int)((maxbits >> 23) & 0xFF) - (int)((__float_as_uint(s) >> 23) & 0xFF));
}

__device__ __forceinline__ int hot_cutbin(const unsigned int* __restrict__ hist, unsigned int cap) {
    unsigned int cum = 0;
    int cut = -1;

    for (int b = 0; b < 32; b++) {
        cum += hist[b];

        if (cum > cap) {
            break;
        }

        cut = b;
    }

    return cut;
}

// Count-mode scout (time-resolved single-slot runs): the scratch volume holds the number of deposits per accumulator for n0 photons
// followed for the first time gate.  The L2 serialises the atomics of one 128-byte line at ~0.7 G/s (1.43 ns each, tools/microbench),
// so a line that takes c deposits per photon costs the run c * 1.43 ns per photon, in parallel with every other line and with the
// walk itself.  The walk costs steps_per_photon / R with R <= 85 G ray-tet steps/s on this GPU (BLB element kernel; slower kernels
// leave more room).  A line can hold the run up only when c * 1.43 ns is comparable to that, so only groups with
//     c >= max(MMCB_HOT_CMIN, MMCB_HOT_CFRAC * steps_per_photon / (85 * 1.43))
// become candidates; with none, the cache (and its lookups) stays off.  steps_per_photon of the whole window is extrapolated from
// the scout: s0 steps per photon in gate 0 and the share f of photons still alive at its end, s0 * (1 + f + ... + f^(gates-1)) --
// the loss per gate falls with time (survivors sit deeper), so this under-estimates and errs towards caching.
#ifndef MMCB_HOT_CMIN
#define MMCB_HOT_CMIN 0.5f
#endif
#ifndef MMCB_HOT_CFRAC
#define MMCB_HOT_CFRAC 0.5f
#endif
__global__ void mmcb_hot_floor_kernel(unsigned int* __restrict__ stat, const double* __restrict__ raytet, const unsigned int* __restrict__ alive,
                                      float n0, int maxgate) {
    float floorv = 0.f, steps = 0.f;

    if (raytet && n0 > 0.f) {
        const float s0 = (float)(*raytet) / n0, f = fminf((float)(*alive) / n0, 0.999f);
        steps = s0 * (1.f - powf(f, (float)maxgate)) / (1.f - f);
        floorv = n0 * fmaxf(MMCB_HOT_CMIN, MMCB_HOT_CFRAC * steps / (85.f * 1.43f));
    }

    stat[MMCB_HOT_STAT_FLOOR] = __float_as_uint(floorv);
    stat[MMCB_HOT_STAT_STEPS] = __float_as_uint(steps);
}

__global__ void mmcb_hot_max_kernel(const acc_t* __restrict__ field, size_t ngroups, size_t fieldlen, unsigned int* __restrict__ stat) {
    float m = 0.f, tot = 0.f;

    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * blockDim.x) {
        const float v = hot_group_sum(field, g, fieldlen);
        m = fmaxf(m, v);
        tot += v;
    }

    #pragma unroll

    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
    }

    if ((threadIdx.x & 31) == 0 && m > 0.f) {
        atomicMax(stat, __float_as_uint(m));      // non-negative floats order like their bit patterns
        atomicAdd((float*)(stat + MMCB_HOT_STAT_TOTAL), tot);
    }
}

__global__ void mmcb_hot_hist_kernel(const acc_t* __restrict__ field, size_t ngroups, size_t fieldlen, unsigned int* __restrict__ stat) {
    __shared__ unsigned int h[32];

    if (threadIdx.x < 32) {
        h[threadIdx.x] = 0;
    }

    __syncthreads();
    const unsigned int maxbits = stat[0];
    const float floorv = __uint_as_float(stat[MMCB_HOT_STAT_FLOOR]);

    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * blockDim.x) {
        float s = hot_group_sum(field, g, fieldlen);

        if (s > 0.f && s >= floorv) {
            atomicAdd(&h[hot_bin(s, maxbits)], 1u);
        }
    }

    __syncthreads();

    if (threadIdx.x < 32 && h[threadIdx.x]) {
        atomicAdd(stat + 2 + threadIdx.x, h[threadIdx.x]);
    }
}

// stat: [0] max bits, [1] candidate count, [2..33] histogram
__global__ void mmcb_hot_select_kernel(const acc_t* __restrict__ field, size_t ngroups, size_t fieldlen, unsigned int* __restrict__ stat,
                                       uint2* __restrict__ cand, unsigned int cap) {
    const unsigned int maxbits = stat[0];
    const int cut = hot_cutbin(stat + 2, cap);
    const float floorv = __uint_as_float(stat[MMCB_HOT_STAT_FLOOR]);

    if (cut < 0 || maxbits == 0) {
        return;
    }

    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * blockDim.x) {
        float s = hot_group_sum(field, g, fieldlen);

        if (s > 0.f && s >= floorv) {
            int b = hot_bin(s, maxbits);

            if (b <= cut) {
                unsigned int pos = atomicAdd(stat + 1, 1u);

                if (pos < cap) {
                    cand[pos] = make_uint2((unsigned int)g, (unsigned int)b);
                }
            }
        }
    }
}

__global__ void mmcb_hot_build_kernel(const uint2* __restrict__ cand, unsigned int* __restrict__ stat, unsigned int cap,
                                      unsigned int* __restrict__ keys, float minshare) {
    __shared__ unsigned int k[MMCB_HOT_SLOTS];

    for (int i = threadIdx.x; i < MMCB_HOT_SLOTS; i += blockDim.x) {
        k[i] = MMCB_HOT_EMPTY;
    }

    __syncthreads();
    const unsigned int n = min(stat[1], cap);

    for (int b = 0; b < 32; b++) {                  // hottest bins claim their slots first
        for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
            if (cand[i].y == (unsigned int)b) {
                atomicCAS(&k[MMCB_HOT_HASH(cand[i].x)], MMCB_HOT_EMPTY, cand[i].x);
            }
        }

        __syncthreads();
    }

    for (int i = threadIdx.x; i < MMCB_HOT_SLOTS; i += blockDim.x) {
        keys[i] = k[i];
    }

    if (threadIdx.x == 0) {     // one line serialises the kernel only when it draws a sizeable share of all deposits
        const float mx = __uint_as_float(stat[0]), tot = __uint_as_float(stat[MMCB_HOT_STAT_TOTAL]);
        unsigned int lo = 0xFFFFFFFFu, hi = 0;

        for (int i = 0; i < MMCB_HOT_SLOTS; i++) {
            if (k[i] != MMCB_HOT_EMPTY) {
                lo = min(lo, k[i]);
                hi = max(hi, k[i]);
            }
        }

        // weight-mode pilots keep the share rule; a count-mode scout has already filtered its candidates by the floor
        const bool useful = (n > 0 && lo <= hi && tot > 0.f && (stat[MMCB_HOT_STAT_FLOOR] ? true : (mx > minshare * tot)));
        stat[MMCB_HOT_STAT_USEFUL] = useful ? 1u : 0u;
        stat[MMCB_HOT_STAT_LO] = useful ? (lo << MMCB_HOT_GROUP_LOG2) : 0u;             // window of accumulator indices that can hit the cache
        stat[MMCB_HOT_STAT_SPAN] = useful ? ((hi - lo + 1u) << MMCB_HOT_GROUP_LOG2) : 0u;
    }
}

// raytet / alive / n0 / maxgate: tallies of a count-mode scout (NULL / 0: the volume holds weights, selection by share)
extern "C" int mmcb_k_hot_select(const void* field, size_t fieldlen, unsigned int* stat, void* cand, unsigned int cap, unsigned int* keys,
                                 float minshare, const double* raytet, const unsigned int* alive, float n0, int maxgate, cudaStream_t st) {
    const size_t ngroups = (fieldlen + MMCB_HOT_GROUP - 1) >> MMCB_HOT_GROUP_LOG2;
    const int grid = (int)std::min<size_t>(148 * 8, (ngroups + 255) / 256);
    cudaError_t e = cudaMemsetAsync(stat, 0, sizeof(unsigned int) * MMCB_HOT_STAT_WORDS, st);

    if (e != cudaSuccess) {
        return (int)e;
    }

    mmcb_hot_floor_kernel<<<1, 1, 0, st>>>(stat, raytet, alive, n0, maxgate);
    mmcb_hot_max_kernel<<<grid, 256, 0, st>>>((const acc_t*)field, ngroups, fieldlen, stat);
    mmcb_hot_hist_kernel<<<grid, 256, 0, st>>>((const acc_t*)field, ngroups, fieldlen, stat);
    mmcb_hot_select_kernel<<<grid, 256, 0, st>>>((const acc_t*)field, ngroups, fieldlen, stat, (uint2*)cand, cap);
    mmcb_hot_build_kernel<<<1, 256, 0, st>>>((const uint2*)cand, stat, cap, keys, minshare);
    return (int)cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------
// host-callable launchers (called from mmcb_host.cu)
// ----------------------------------------------------------------------------------------------------
extern "C" int mmcb_k_upload_param(const mmcb_kparam* hp, const float* det4, int detnum, cudaStream_t st) {
    cudaError_t e = cudaMemcpyToSymbolAsync(gp, hp, sizeof(mmcb_kparam), 0, cudaMemcpyHostToDevice, st);

    if (e == cudaSuccess && detnum > 0) {
        e = cudaMemcpyToSymbolAsync(gdet, det4, sizeof(float4) * detnum, 0, cudaMemcpyHostToDevice, st);
    }

    return (int)e;
}

// one entry per (tracer, detection, general-source) combination
typedef void (*photon_kernel_t)(const mmcb_kargs);
template <int METHOD>
static photon_kernel_t pick_kernel(int isdet, int isgeneral) {
    if (isdet) {
        return isgeneral ? mmcb_photon_kernel<METHOD, true, true> : mmcb_photon_kernel<METHOD, true, false>;
    }

    return isgeneral ? mmcb_photon_kernel<METHOD, false, true> : mmcb_photon_kernel<METHOD, false, false>;
}
template <int METHOD>
static photon_kernel_t pick_kernel_nodal(int isdet, int isgeneral) {
    if (isdet) {
        return isgeneral ? mmcb_photon_kernel<METHOD, true, true, false, false, true> : mmcb_photon_kernel<METHOD, true, false, false, false, true>;
    }

    return isgeneral ? mmcb_photon_kernel<METHOD, false, true, false, false, true> : mmcb_photon_kernel<METHOD, false, false, false, false, true>;
}
// variant: bit 0 = dual grid with capped segment loop (CAP), bit 1 = Havel / Plucker with nodal deposit (NODAL)
static photon_kernel_t pick_kernel(int method, int isdet, int isgeneral, int isrf, int variant = 0) {
    const int iscap = variant & 1;

    if (method <= 1 && (variant & 2)) {
        return method == 1 ? pick_kernel_nodal<1>(isdet, isgeneral) : pick_kernel_nodal<0>(isdet, isgeneral);
    }

    if (method == 4 && iscap && !isrf) {        // dual grid, long steps walked in pieces (gp.lcap)
        if (isdet) {
            return isgeneral ? mmcb_photon_kernel<4, true, true, false, true> : mmcb_photon_kernel<4, true, false, false, true>;
        }

        return isgeneral ? mmcb_photon_kernel<4, false, true, false, true> : mmcb_photon_kernel<4, false, false, false, true>;
    }

    if (isrf) {         // RF forward: general branch-less Badouel kernels only (mesh or dual-grid deposit)
        if (method == 4) {
            return isdet ? mmcb_photon_kernel<4, true, true, true> : mmcb_photon_kernel<4, false, true, true>;
        }

        return isdet ? mmcb_photon_kernel<3, true, true, true> : mmcb_photon_kernel<3, false, true, true>;
    }

    switch (method) {
        case 0:
            return pick_kernel<0>(isdet, isgeneral);

        case 1:
            return pick_kernel<1>(isdet, isgeneral);

        case 4:
            return pick_kernel<4>(isdet, isgeneral);

        default:
            return pick_kernel<3>(isdet, isgeneral);
    }
}

static photon_kernel_t pick_kernel_rp(int method, int isdet) {      // lane re-packing kernels (mmcb_kernel_rp.cuh)
    if (method == 4) {
        return isdet ? mmcb_photon_kernel_rp<4, true> : mmcb_photon_kernel_rp<4, false>;
    }

    return isdet ? mmcb_photon_kernel_rp<3, true> : mmcb_photon_kernel_rp<3, false>;
}

// shared memory the re-packing kernel needs on top of the media table (and hot-line cache): per warp 32 stashed walkers and the
// slot-rank scratch, plus two partial-path columns per thread
extern "C" size_t mmcb_k_rp_smem(int block, int isdet, int devreclen) {
    const size_t nwarp = (size_t)block / 32;
    return nwarp * ((isdet ? 7 : 5) * MMCB_RP_SLOTS * 16 + 32 * 4) + (isdet ? sizeof(float) * (size_t)devreclen * 2 * block : 0);
}

extern "C" int mmcb_k_launch_photons(const mmcb_kargs* a, int grid, int block, size_t smem, int method, int isdet, int isgeneral, int isrf, int carveout,
                                     int repack, int iscap, cudaStream_t st) {
    photon_kernel_t k = repack ? pick_kernel_rp(method, isdet) : pick_kernel(method, isdet, isgeneral, isrf, iscap);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

    if (e != cudaSuccess) {
        return (int)e;
    }

    // shared-memory carve-out in percent: exactly what the resident CTAs need, the rest of the 256 KB stays L1 for the record
    // gathers (the driver's own choice over-provisions kernels with detector columns: head-like 290 -> 273 ms)
    if (const char* co = getenv("MMCB_CARVEOUT")) {
        carveout = atoi(co);
    }

    if (carveout >= 0) {
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
    }

    k<<<grid, block, smem, st>>>(*a);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_max_block(int method, int repack) {      // largest (and default) block size the kernel of this tracer is compiled for
    return repack ? MMCB_RP_THREADS : ((method <= 1) ? MMCB_MAXTHREADS_HP : MMCB_MAXTHREADS);
}

extern "C" int mmcb_k_occupancy(int block, size_t smem, int method, int isdet, int isgeneral, int isrf, int repack, int iscap, int* blocks_per_sm) {
    photon_kernel_t k = repack ? pick_kernel_rp(method, isdet) : pick_kernel(method, isdet, isgeneral, isrf, iscap);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

    if (e == cudaSuccess) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, block, smem);
    }

    return (int)e;
}

extern "C" int mmcb_k_spread_nodes(const void* efield, double* nfield, const int* elem, int ne, int nn, int maxgate, int srcnum, cudaStream_t st) {
    mmcb_spread_nodes_kernel<<<148 * 8, 256, 0, st>>>((const acc_t*)efield, nfield, elem, ne, nn, maxgate, srcnum);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_acc_to_double(const void* in, double* out, size_t n, cudaStream_t st) {
    mmcb_acc_to_double_kernel<<<148 * 8, 256, 0, st>>>((const acc_t*)in, out, n);
    return (int)cudaGetLastError();
}

// RNG known-answer kernel: stream i draws `ndraw` floats with the same rand01() the photon kernel uses
__global__ void mmcb_rng_kernel(const uint32_t* __restrict__ seeds, int nstream, int ndraw, float* __restrict__ out,
                                unsigned long long* __restrict__ state_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i >= nstream) {
        return;
    }

    Rng r;
    r.t0 = ((unsigned long long)seeds[4 * i] << 32) | seeds[4 * i + 1];
    r.t1 = ((unsigned long long)seeds[4 * i + 2] << 32) | seeds[4 * i + 3];

    for (int k = 0; k < ndraw; k++) {
        out[(size_t)i * ndraw + k] = rand01(r);
    }

    state_out[2 * i] = r.t0;
    state_out[2 * i + 1] = r.t1;
}

extern "C" int mmcb_k_rng(const uint32_t* dseeds, int nstream, int ndraw, float* dout, unsigned long long* dstate, cudaStream_t st) {
    mmcb_rng_kernel<<<(nstream + 127) / 128, 128, 0, st>>>(dseeds, nstream, ndraw, dout, dstate);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_acc_is_double(void) {
    return sizeof(acc_t) == 8;
}
