/*
 * TEST INFRASTRUCTURE ONLY -- see mmc_oracle.h.  CPU oracle: plain-C restatement of the
 * reference's photon random walk.  Never linked into the product library.
 *
 * Parity status: PINNED against the unmodified reference binary oracle/_ref/mmc_ref
 * (tests/test_oracle_vs_ref.py, fixtures in tests/golden/).
 *
 * All path:line citations are relative to /root/reference/.
 * Compile with -ffp-contract=off: the reference CPU build has no FMA contraction and the
 * single-thread pin relies on identical rounding.
 */
#define _GNU_SOURCE
#include "mmc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* constants: src/mmc_mesh.h:65-74, src/mmc_const.h:45, src/mmc_raytrace.h:44-45 */
#define ORC_UNDEFINED   (3.40282347e+38F)
#define ID_UNDEFINED    0x7FFFFFFF
#define EPS             1e-6f
#define R_MIN_MUS       1e9f
#define DELTA_MUA       1e-4f
#define VERY_BIG        1e30f
#define R_C0            3.335640951981520e-12f
#define TWO_PI          (M_PI * 2.0)
#define MAX_TRIAL       3
#define FIX_PHOTON      1e-3f
#define JUST_BELOW_ONE  0.9998f

/* source types: src/mmc_const.h:49-66 */
enum { stPencil = 0, stIsotropic, stCone, stGaussian, stPlanar, stPattern, stFourier, stArcSin, stDisk,
       stFourierX, stFourier2D, stZGaussian, stLine, stSlit
     };

/* index tables: src/mmc_mesh.c:59-103, src/mmc_raytrace.c:69-93,152 */
static const int out_[4][3] = {{0, 3, 1}, {3, 2, 1}, {0, 2, 3}, {0, 1, 2}};
static const int facemap_[4] = {2, 0, 1, 3};
static const int ifacemap_[4] = {1, 2, 0, 3};
static const int faceorder_[5] = {1, 3, 2, 0, -1};
static const int ifaceorder_[4] = {3, 0, 2, 1};
static const int fc_[4][3] = {{0, 4, 2}, {3, 5, 4}, {2, 5, 1}, {1, 3, 0}};
static const int nc_[4][3] = {{3, 0, 1}, {3, 1, 2}, {2, 0, 3}, {1, 0, 2}};
static const int facelist_[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}}; /* src/mmc_highorder.cpp:50 */

static char g_err[512] = "";
const char* orc_last_error(void) {
    return g_err;
}
#define ORC_FAIL(...) do { snprintf(g_err, sizeof(g_err), __VA_ARGS__); return -1; } while (0)

#define MINI(a, b) ((a) < (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------
 * RNG -- src/mmc_rand_xorshift128p.c:55-76, src/mmc_rand_common.h:48-66
 * ---------------------------------------------------------------------------------------- */
void orc_rng_seed(const uint32_t seed4[4], uint64_t t[2]) {
    t[0] = ((uint64_t)seed4[0] << 32) | seed4[1];
    t[1] = ((uint64_t)seed4[2] << 32) | seed4[3];
}

float orc_rng_nextf(uint64_t t[2]) {
    union {
        uint64_t i;
        float f[2];
        uint32_t u[2];
    } s1;
    const uint64_t s0 = t[1];
    s1.i = t[0];
    t[0] = s0;
    s1.i ^= s1.i << 23;
    t[1] = s1.i ^ s0 ^ (s1.i >> 18) ^ (s0 >> 5);
    s1.i = t[1] + s0;
    s1.u[0] = 0x3F800000U | (s1.u[0] >> 9);
    return s1.f[0] - 1.0f;
}

static inline float rand01(uint64_t* t) {
    return orc_rng_nextf(t);
}
static inline float rand_scatlen(uint64_t* t) {
    return -logf(rand01(t) + EPS);
}

/* host seeding: src/mmc_host.c:240-245, src/mmc_cu_host.cu:438,532-534 */
void orc_host_seeds(int seed, int count, uint32_t* outv) {
    int i;
    srand((unsigned)seed);

    for (i = 0; i < count; i++) {
        outv[i] = (uint32_t)rand();
    }
}

/* ------------------------------------------------------------------------------------------
 * mesh preprocessing
 * ---------------------------------------------------------------------------------------- */
static inline const float* nd(const orc_mesh* m, int id1) {
    return m->node + 3 * (size_t)(id1 - 1);
}

/* src/mmc_mesh.c:910-948 */
static void mesh_getvolume(orc_mesh* m) {
    int i, j;
    m->evol = (float*)calloc(m->ne, sizeof(float));
    m->nvol = (float*)calloc(m->nn, sizeof(float));

    for (i = 0; i < m->ne; i++) {
        int* ee = m->elem + 4 * i;
        const float* n0 = nd(m, ee[0]), *n1 = nd(m, ee[1]), *n2 = nd(m, ee[2]), *n3 = nd(m, ee[3]);
        float dx = n2[0] - n3[0], dy = n2[1] - n3[1], dz = n2[2] - n3[2];
        float v;
        v = n1[0] * (n2[1] * n3[2] - n2[2] * n3[1]) - n1[1] * (n2[0] * n3[2] - n2[2] * n3[0]) + n1[2] * (n2[0] * n3[1] - n2[1] * n3[0]);
        v += -n0[0] * ((n2[1] * n3[2] - n2[2] * n3[1]) + n1[1] * dz - n1[2] * dy);
        v += +n0[1] * ((n2[0] * n3[2] - n2[2] * n3[0]) + n1[0] * dz - n1[2] * dx);
        v += -n0[2] * ((n2[0] * n3[1] - n2[1] * n3[0]) + n1[0] * dy - n1[1] * dx);
        v = -v;

        if (v < 0.f) {
            int e1 = ee[3];
            ee[3] = ee[2];
            ee[2] = e1;
            v = -v;
        }

        v *= (1.f / 6.f);
        m->evol[i] = v;

        if (m->type[i] == 0) {
            continue;
        }

        for (j = 0; j < 4; j++) {
            m->nvol[ee[j] - 1] += v * 0.25f;
        }
    }
}

/* src/mmc_highorder.cpp:124-159 -- same result via a sort-based matching (neighbours are unique) */
typedef struct {
    int a, b, c, slot;
} facekey;
static int facecmp(const void* x, const void* y) {
    const facekey* p = (const facekey*)x, *q = (const facekey*)y;

    if (p->a != q->a) {
        return p->a < q->a ? -1 : 1;
    }

    if (p->b != q->b) {
        return p->b < q->b ? -1 : 1;
    }

    if (p->c != q->c) {
        return p->c < q->c ? -1 : 1;
    }

    return p->slot < q->slot ? -1 : (p->slot > q->slot);
}
static void mesh_getfacenb(orc_mesh* m) {
    size_t nfk = (size_t)m->ne * 4, i;
    facekey* k = (facekey*)malloc(nfk * sizeof(facekey));
    m->facenb = (int*)calloc(nfk, sizeof(int));

    for (i = 0; i < (size_t)m->ne; i++) {
        int j, *ee = m->elem + 4 * i;

        for (j = 0; j < 4; j++) {
            int v[3] = {ee[facelist_[j][0]], ee[facelist_[j][1]], ee[facelist_[j][2]]}, t;

            if (v[0] > v[1]) {
                t = v[0];
                v[0] = v[1];
                v[1] = t;
            }

            if (v[1] > v[2]) {
                t = v[1];
                v[1] = v[2];
                v[2] = t;
            }

            if (v[0] > v[1]) {
                t = v[0];
                v[0] = v[1];
                v[1] = t;
            }

            k[i * 4 + j].a = v[0];
            k[i * 4 + j].b = v[1];
            k[i * 4 + j].c = v[2];
            k[i * 4 + j].slot = (int)(i * 4 + j);
        }
    }

    qsort(k, nfk, sizeof(facekey), facecmp);

    for (i = 0; i + 1 < nfk; i++) {
        if (k[i].a == k[i + 1].a && k[i].b == k[i + 1].b && k[i].c == k[i + 1].c) {
            m->facenb[k[i].slot] = (k[i + 1].slot >> 2) + 1;
            m->facenb[k[i + 1].slot] = (k[i].slot >> 2) + 1;
            i++;
        }
    }

    free(k);
}

/* src/mmc_mesh.c:390-427 + the -2 relabel of mesh_loadmedia :516-524 */
static void mesh_srcdetelem(orc_mesh* m) {
    int i, is = 0, id = 0;
    m->srcelemlen = m->detelemlen = 0;
    m->e0_from_src = 0;

    for (i = 0; i < m->ne; i++) {
        if (m->type[i] == -1) {
            m->srcelemlen++;

            if (!m->e0_from_src) {
                m->e0_from_src = i + 1;
            }
        }

        if (m->type[i] == -2) {
            m->detelemlen++;
            m->isextdet = 1;
        }
    }

    m->srcelem = (int*)calloc(m->srcelemlen + 1, sizeof(int));
    m->detelem = (int*)calloc(m->detelemlen + 1, sizeof(int));

    for (i = 0; i < m->ne; i++) {
        if (m->type[i] == -1) {
            m->srcelem[is++] = i + 1;
            m->type[i] = 0;
        } else if (m->type[i] == -2) {
            m->detelem[id++] = i + 1;
        }
    }
}

orc_mesh* orc_mesh_create(int nn, const float* node, int ne, const int* elem, const int* type,
                          int prop, const float* med, float nout, float unitinmm, const int* facenb_or_null,
                          const float* evol_or_null) {
    int i;
    orc_mesh* m = (orc_mesh*)calloc(1, sizeof(orc_mesh));
    m->nn = nn;
    m->ne = ne;
    m->prop = prop;
    m->node = (float*)malloc(sizeof(float) * 3 * nn);
    memcpy(m->node, node, sizeof(float) * 3 * nn);
    m->elem = (int*)malloc(sizeof(int) * 4 * ne);
    memcpy(m->elem, elem, sizeof(int) * 4 * ne);
    m->type = (int*)malloc(sizeof(int) * ne);
    memcpy(m->type, type, sizeof(int) * ne);
    mesh_srcdetelem(m);
    /* media: src/mmc_mesh.c:511-547 */
    m->med = (float*)calloc((size_t)(prop + 2) * 4, sizeof(float));
    memcpy(m->med, med, sizeof(float) * 4 * (prop + 1));
    m->med[0] = 0.f;
    m->med[1] = 0.f;
    m->med[2] = 1.f;
    m->med[3] = nout;

    if (m->isextdet) {
        memcpy(m->med + 4 * (prop + 1), m->med, 4 * sizeof(float));

        for (i = 0; i < ne; i++)
            if (m->type[i] == -2) {
                m->type[i] = prop + 1;
            }
    }

    if (unitinmm != 1.f)
        for (i = 1; i <= prop; i++) {
            m->med[4 * i + 1] *= unitinmm;
            m->med[4 * i] *= unitinmm;
        }

    if (evol_or_null) {   /* mesh_loadelemvol src/mmc_mesh.c:723-761: file volumes => no orientation swap */
        int j;
        m->evol = (float*)malloc(sizeof(float) * ne);
        memcpy(m->evol, evol_or_null, sizeof(float) * ne);
        m->nvol = (float*)calloc(nn, sizeof(float));

        for (i = 0; i < ne; i++) {
            if (m->type[i] == 0) {
                continue;
            }

            for (j = 0; j < 4; j++) {
                m->nvol[m->elem[4 * i + j] - 1] += m->evol[i] * 0.25f;
            }
        }
    } else {
        mesh_getvolume(m);
    }

    if (facenb_or_null) {
        m->facenb = (int*)malloc(sizeof(int) * 4 * ne);
        memcpy(m->facenb, facenb_or_null, sizeof(int) * 4 * ne);
    } else {
        mesh_getfacenb(m);
    }

    return m;
}

void orc_mesh_free(orc_mesh* m) {
    if (!m) {
        return;
    }

    free(m->node);
    free(m->elem);
    free(m->type);
    free(m->med);
    free(m->facenb);
    free(m->evol);
    free(m->nvol);
    free(m->srcelem);
    free(m->detelem);
    free(m->n);
    free(m->m);
    free(m->pd);
    free(m->pm);
    free(m);
}

static inline void diff3(const float* a, const float* b, float* r) {
    r[0] = b[0] - a[0];
    r[1] = b[1] - a[1];
    r[2] = b[2] - a[2];
}
static inline void cross3(const float* a, const float* b, float* r) {
    float x = a[1] * b[2] - a[2] * b[1];
    float y = a[2] * b[0] - a[0] * b[2];
    float z = a[0] * b[1] - a[1] * b[0];
    r[0] = x;
    r[1] = y;
    r[2] = z;
}
static inline float dot3(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* src/mmc_mesh.c:1496-1602 */
void orc_mesh_build_tracer(orc_mesh* m, int method) {
    static const int pairs[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    int i, j, ne = m->ne;
    float Rn2;
    free(m->n);
    free(m->m);
    free(m->pd);
    free(m->pm);
    m->n = m->m = m->pd = m->pm = NULL;

    if (method == ORC_PLUCKER) {
        m->pd = (float*)calloc((size_t)ne * 6 * 4, sizeof(float));
        m->pm = (float*)calloc((size_t)ne * 6 * 4, sizeof(float));

        for (i = 0; i < ne; i++)
            for (j = 0; j < 6; j++) {
                const float* p0 = nd(m, m->elem[4 * i + pairs[j][0]]), *p1 = nd(m, m->elem[4 * i + pairs[j][1]]);
                diff3(p0, p1, m->pd + ((size_t)i * 6 + j) * 4);
                cross3(p0, p1, m->pm + ((size_t)i * 6 + j) * 4);
            }
    } else if (method == ORC_HAVEL || method == ORC_BADOUEL) {
        m->m = (float*)calloc((size_t)ne * 12 * 4, sizeof(float));

        for (i = 0; i < ne; i++)
            for (j = 0; j < 4; j++) {
                float* vN = m->m + ((size_t)(4 * i + j) * 3) * 4, AB[3], AC[3];
                const float* a = nd(m, m->elem[4 * i + out_[j][0]]), *b = nd(m, m->elem[4 * i + out_[j][1]]), *c = nd(m, m->elem[4 * i + out_[j][2]]);
                int k;
                diff3(a, b, AB);
                diff3(a, c, AC);
                cross3(AB, AC, vN);
                cross3(AC, vN, vN + 4);
                cross3(vN, AB, vN + 8);
                Rn2 = 1.f / sqrt(dot3(vN, vN));

                for (k = 0; k < 3; k++) {
                    vN[k] = Rn2 * vN[k];
                }

                Rn2 *= Rn2;

                for (k = 0; k < 3; k++) {
                    vN[4 + k] = Rn2 * vN[4 + k];
                    vN[8 + k] = Rn2 * vN[8 + k];
                }

                vN[3] = dot3(vN, a);
                vN[7] = -dot3(vN + 4, a);
                vN[11] = -dot3(vN + 8, a);
            }
    }

    if (method == ORC_PLUCKER || method == ORC_BLBADOUEL || method == ORC_BLBADOUEL_GRID) {
        m->n = (float*)calloc((size_t)ne * 16, sizeof(float));

        for (i = 0; i < ne; i++) {
            float* vecN = m->n + (size_t)i * 16;

            for (j = 0; j < 4; j++) {
                float AB[3], AC[3], vN[3];
                const float* a = nd(m, m->elem[4 * i + out_[j][0]]), *b = nd(m, m->elem[4 * i + out_[j][1]]), *c = nd(m, m->elem[4 * i + out_[j][2]]);
                diff3(a, b, AB);
                diff3(a, c, AC);
                cross3(AB, AC, vN);
                Rn2 = 1.f / sqrt(dot3(vN, vN));
                vN[0] = Rn2 * vN[0];
                vN[1] = Rn2 * vN[1];
                vN[2] = Rn2 * vN[2];
                vecN[j] = vN[0];
                vecN[j + 4] = vN[1];
                vecN[j + 8] = vN[2];
                vecN[j + 12] = dot3(vN, a);
            }
        }
    }
}

/* src/mmc_mesh.c:1170-1206 */
static int mesh_barycentric(const orc_mesh* m, int e0, float* bary, const float* srcpos) {
    int i;
    const int* ee = m->elem + 4 * (size_t)(e0 - 1);
    float s = 0.f;

    if (e0 < 1 || e0 > m->ne) {
        return 1;
    }

    for (i = 0; i < 4; i++) {
        float AB[3], AC[3], S[3], N[3];
        const float* a = nd(m, ee[out_[i][0]]), *b = nd(m, ee[out_[i][1]]), *c = nd(m, ee[out_[i][2]]);
        diff3(a, b, AB);
        diff3(a, c, AC);
        diff3(a, srcpos, S);
        cross3(AB, AC, N);
        bary[facemap_[i]] = -dot3(S, N);
    }

    for (i = 0; i < 4; i++) {
        if (bary[i] < 0.f) {
            return 1;
        }

        s += bary[i];
    }

    for (i = 0; i < 4; i++) {
        bary[i] /= s;
    }

    return 0;
}

/* src/mmc_mesh.c:1060-1090 */
int orc_mesh_initelem(const orc_mesh* m, const float* srcpos, float* bary4) {
    int i, j;

    for (i = 0; i < m->ne; i++) {
        double pmin[3] = {VERY_BIG, VERY_BIG, VERY_BIG}, pmax[3] = { -VERY_BIG, -VERY_BIG, -VERY_BIG};
        const int* ee = m->elem + 4 * (size_t)i;

        for (j = 0; j < 4; j++) {
            const float* p = nd(m, ee[j]);
            int k;

            for (k = 0; k < 3; k++) {
                if (p[k] < pmin[k]) {
                    pmin[k] = p[k];
                }

                if (p[k] > pmax[k]) {
                    pmax[k] = p[k];
                }
            }
        }

        if (srcpos[0] <= pmax[0] && srcpos[0] >= pmin[0] && srcpos[1] <= pmax[1] && srcpos[1] >= pmin[1] &&
                srcpos[2] <= pmax[2] && srcpos[2] >= pmin[2]) {
            if (mesh_barycentric(m, i + 1, bary4, srcpos) == 0) {
                return i + 1;
            }
        }
    }

    return 0;
}

/* src/mmc_mesh.c:349-381 */
void orc_mesh_dualgrid(orc_mesh* m, float step, int* dim, unsigned int* crop) {
    int i, k;

    for (k = 0; k < 3; k++) {
        m->nmin[k] = VERY_BIG;
        m->nmax[k] = -VERY_BIG;
    }

    for (i = 0; i < m->nn; i++)
        for (k = 0; k < 3; k++) {
            float v = m->node[3 * i + k];

            if (v < m->nmin[k]) {
                m->nmin[k] = v;
            }

            if (v > m->nmax[k]) {
                m->nmax[k] = v;
            }
        }

    for (k = 0; k < 3; k++) {
        m->nmin[k] -= EPS;
        m->nmax[k] += EPS;
        dim[k] = (int)((m->nmax[k] - m->nmin[k]) / step) + 1;
    }

    crop[0] = dim[0];
    crop[1] = dim[1] * dim[0];
    crop[2] = dim[1] * dim[0] * dim[2];
}

/* src/mmc_mesh.c:2305-2334 */
static double mesh_getreff(double n_in, double n_out) {
    double oc = asin(1.0 / n_in);
    const double count = 1000.0;
    const double ostep = (M_PI / (2.0 * count));
    double r_phi = 0.0, r_j = 0.0, o, cosop, coso, r_fres, tmp;
    int i;

    for (i = 0; i < count; i++) {
        o = i * ostep;
        coso = cos(o);

        if (o < oc) {
            cosop = n_in * sin(o);
            cosop = sqrt(1. - cosop * cosop);
            tmp = (n_in * cosop - n_out * coso) / (n_in * cosop + n_out * coso);
            r_fres = 0.5 * tmp * tmp;
            tmp = (n_in * coso - n_out * cosop) / (n_in * coso + n_out * cosop);
            r_fres += 0.5 * tmp * tmp;
        } else {
            r_fres = 1.f;
        }

        r_phi += 2.0 * sin(o) * coso * r_fres;
        r_j += 3.0 * sin(o) * coso * coso * r_fres;
    }

    r_phi *= ostep;
    r_j *= ostep;
    return (r_phi + r_j) / (2.0 - r_phi + r_j);
}

/* surface-node nvol correction + exterior face numbering: src/mmc_mesh.c:1344-1386,1466-1474 */
static void tracer_prep_mesh(orc_mesh* m, const orc_config* cfg) {
    int i, j, k, ne = m->ne;

    if (m->nf > 0) {
        return;    /* already prepared (facenb already negative) */
    }

    if (cfg->isnormalized == 1 && cfg->method != ORC_BLBADOUEL_GRID && cfg->basisorder) {
        float* Reff = (float*)calloc(m->prop + 2, sizeof(float));

        if (cfg->isreflect) {
            for (i = 1; i <= m->prop; i++) {
                for (j = 1; j < i; j++) {
                    if (m->med[4 * j + 3] == m->med[4 * i + 3]) {
                        Reff[i] = Reff[j];
                        break;
                    }
                }

                if (Reff[i] == 0.f) {
                    Reff[i] = mesh_getreff(m->med[4 * i + 3], m->med[3]);
                }
            }
        }

        for (i = 0; i < ne; i++) {
            const int* ee = m->elem + 4 * (size_t)i, *enb = m->facenb + 4 * (size_t)i;

            for (j = 0; j < 4; j++) {
                if (enb[j] == 0) {
                    for (k = 0; k < 3; k++) {
                        int nid = ee[out_[ifaceorder_[j]][k]] - 1;

                        if (m->nvol[nid] > 0.f && m->type[i] >= 0) {
                            m->nvol[nid] *= -(2.f / (1.0 + Reff[m->type[i] > m->prop ? 0 : m->type[i]]));
                        }
                    }
                }
            }
        }

        free(Reff);

        for (i = 0; i < m->nn; i++) {
            if (m->nvol[i] < 0.f) {
                m->nvol[i] = -m->nvol[i];
            }
        }
    }

    m->nf = 0;

    for (i = 0; i < ne * 4; i++) {
        if (m->facenb[i] == 0) {
            m->facenb[i] = -(++m->nf);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * photon state -- src/mmc_raytrace.h:53-85 (ray), :93-106 (visitor)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float p0[4], vec[4], pout[3];
    float bary0[4];
    int eid, faceid, isend, nexteid;
    float weight, photontimer, slen0, slen, Lmove;
    double Eabsorb;
    unsigned int photonid;
    float* partialpath;
    float focus;
    unsigned int posidx, oldidx;
    double oldweight;
} oray;

typedef struct {
    const orc_mesh* mesh;
    orc_config* cfg;
    double* weight;            /* shared field */
    double* dref;
    float rtstep;
    int reclen, maxgate, framelen;
    unsigned int crop0[3];
    /* per-thread tallies */
    double launchweight[16], absorbweight[16], escweight[16], kahanc0[16], kahanc1[16];
    double raytet;
    float* detbuf;
    uint64_t* seedbuf;
    unsigned int* detcount;
    unsigned int detcap;
    float bary0[4];
    /* traj */
    float* traj;
    unsigned int* trajcount;
} octx;

static inline const float* medium_of(const orc_mesh* m, int type) {
    return m->med + 4 * (size_t)type;
}

static inline int is_fluence_like(int ot) {  /* src/mmc_raytrace.c:57-59 */
    return ot != ORC_ENERGY && ot != ORC_WP && ot != ORC_WL;
}
static inline float fluence_deposit(float w0, float w1, float mua, float len) { /* :61-63 */
    return (mua < EPS) ? (w0 * len) : ((w0 - w1) / mua);
}

static inline void atomic_add(double* p, double v) {
    #pragma omp atomic
    *p += v;
}

/* deposit helper for single/pattern sources (the repeated blocks at src/mmc_raytrace.c:1586-1611 etc.) */
static inline void deposit(octx* c, unsigned int idx, double w, const oray* r) {
    const orc_config* cfg = c->cfg;

    if (cfg->srctype != stPattern || cfg->srcnum == 1) {
        atomic_add(c->weight + idx, w);
    } else {
        int pidx;

        for (pidx = 0; pidx < cfg->srcnum; pidx++) {
            atomic_add(c->weight + (size_t)idx * cfg->srcnum + pidx, w * cfg->srcpattern[(size_t)r->posidx * cfg->srcnum + pidx]);
        }
    }
}

/* common head of every tracer step: time window, attenuation, replay weights
 * (src/mmc_raytrace.c:1510-1544 ; :357-388 ; :633-664).  ge=1 selects the ">=" window test of the
 * Plucker/Havel tracers, ge=0 the "> maxgate-1" test of BLB. */
static inline float step_common(octx* c, oray* r, const float* prop, float mus, int ge) {
    const orc_config* cfg = c->cfg;
    float rc = prop[3] * R_C0;
    float currweight = r->weight;
    int tg = (int)((r->photontimer + r->Lmove * rc - cfg->tstart) * c->rtstep);
    int hit = ge ? (tg >= (int)((cfg->tend - cfg->tstart) * c->rtstep)) : (tg > c->maxgate - 1);

    if (hit) {
        r->faceid = -2;
        r->pout[0] = ORC_UNDEFINED;
        r->Lmove = (cfg->tend - r->photontimer) / (prop[3] * R_C0) - 1e-4f;
    }

    r->weight *= expf(-prop[0] * r->Lmove);

    if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_JACOBIAN) {
        if (cfg->gpu_semantics) {           /* src/mmc_core.cl:815-818 */
            currweight = r->Lmove;
        } else {
            currweight = expf(-DELTA_MUA * r->Lmove);
        }

        currweight *= cfg->replayweight[r->photonid];
        currweight += r->weight;
    } else if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_WL) {
        currweight = r->Lmove;
        currweight *= cfg->replayweight[r->photonid];
        currweight += r->weight;
    }

    r->slen -= r->Lmove * mus;

    if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_WP) {
        if (r->slen0 < EPS) {
            currweight = 1;
        } else {
            currweight = r->Lmove * mus / r->slen0;
        }

        currweight *= cfg->replayweight[r->photonid];
        currweight += r->weight;
    }

    return currweight;
}

static inline int gate_shift(const octx* c, const oray* r, int framelen) {
    const orc_config* cfg = c->cfg;

    if (cfg->outputtype == ORC_WL || cfg->outputtype == ORC_WP) {
        return MINI(((int)(cfg->replaytime[r->photonid] * c->rtstep)), c->maxgate - 1) * framelen;
    }

    return MINI(((int)((r->photontimer - cfg->tstart) * c->rtstep)), c->maxgate - 1) * framelen;
}

/* ------------------------------------------------------------------------------------------
 * Branch-less Badouel -- src/mmc_raytrace.c:1414-1709 (mesh + dual-grid deposit)
 * ---------------------------------------------------------------------------------------- */
static float blb_raytet(oray* r, octx* c) {
    const orc_mesh* m = c->mesh;
    const orc_config* cfg = c->cfg;
    float T[4], S[4], Lmin, totalloss = 0.f, ww, currweight, dlen, rc, Lp0;
    int i, faceidx, eid, tshift, mask = 0;
    const float* N;

    if (m->n == NULL || r->eid <= 0 || r->eid > m->ne) {
        return -1;
    }

    eid = r->eid - 1;
    N = m->n + (size_t)eid * 16;
    r->pout[0] = ORC_UNDEFINED;
    r->faceid = -1;
    r->isend = 0;

    for (i = 0; i < 4; i++) {
        float t = N[i] * r->p0[0];
        float s = N[i] * r->vec[0];
        t = t + N[4 + i] * r->p0[1];
        t = t + N[8 + i] * r->p0[2];
        t = N[12 + i] - t;
        s = s + N[4 + i] * r->vec[1];
        s = s + N[8 + i] * r->vec[2];
        t = t / s;
        T[i] = (s > 0.f) ? (0.f + t) : (1e10f + 0.f);   /* andnot/and + add, :1459-1460 */
        S[i] = s;
    }

    {
        /* min via movehl/min/shuffle/min_ss, :1461-1464 */
        float a = (T[0] < T[2]) ? T[0] : T[2];   /* _mm_min_ps returns 2nd operand if equal/NaN: min(T,S) = T<S?T:S */
        float b = (T[1] < T[3]) ? T[1] : T[3];
        Lmin = (a < b) ? a : b;
    }

    for (i = 0; i < 4; i++)
        if (T[i] == Lmin) {
            mask |= (1 << i);
        }

    {
        static const char maskmap[16] = {4, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3};
        faceidx = maskmap[mask];

        if (cfg->gpu_semantics) { /* src/mmc_core.cl:770 picks the first tied face */
            faceidx = ((Lmin == 1e10f) ? 4 : Lmin == T[0] ? 0 : (Lmin == T[1] ? 1 : (Lmin == T[2] ? 2 : 3)));
        }
    }

    r->faceid = faceorder_[faceidx];

    if (r->faceid >= 0 && Lmin >= 0) {
        const float* prop = medium_of(m, m->type[eid]);
        const int* ee = m->elem + 4 * (size_t)eid;
        float mus = prop[1];
        rc = prop[3] * R_C0;
        r->nexteid = m->facenb[4 * (size_t)eid + r->faceid];
        dlen = (mus <= EPS) ? R_MIN_MUS : r->slen / mus;
        Lp0 = Lmin;
        r->isend = (Lp0 > dlen);
        r->Lmove = ((r->isend) ? dlen : Lp0);
        r->pout[0] = r->vec[0] * Lmin + r->p0[0];
        r->pout[1] = r->vec[1] * Lmin + r->p0[1];
        r->pout[2] = r->vec[2] * Lmin + r->p0[2];

        {
            /* step_common computes totalloss implicitly; BLB needs it for the grid split (:1517-1521) */
            float w_before = r->weight;
            (void)w_before;
        }

        /* :1510-1544, written out because totalloss is needed below */
        if ((int)((r->photontimer + r->Lmove * rc - cfg->tstart)*c->rtstep) > c->maxgate - 1) {
            r->faceid = -2;
            r->pout[0] = ORC_UNDEFINED;
            r->Lmove = (cfg->tend - r->photontimer) / (prop[3] * R_C0) - 1e-4f;
        }

        currweight = r->weight;
        totalloss = expf(-prop[0] * r->Lmove);
        r->weight *= totalloss;
        totalloss = 1.f - totalloss;

        if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_JACOBIAN) {
            currweight = cfg->gpu_semantics ? r->Lmove : expf(-DELTA_MUA * r->Lmove);
            currweight *= cfg->replayweight[r->photonid];
            currweight += r->weight;
        } else if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_WL) {
            currweight = r->Lmove;
            currweight *= cfg->replayweight[r->photonid];
            currweight += r->weight;
        }

        r->slen -= r->Lmove * mus;

        if (cfg->seed == ORC_SEED_FROM_FILE && cfg->outputtype == ORC_WP) {
            if (r->slen0 < EPS) {
                currweight = 1;
            } else {
                currweight = r->Lmove * mus / r->slen0;
            }

            currweight *= cfg->replayweight[r->photonid];
            currweight += r->weight;
        }

        {
            int framelen = (cfg->basisorder ? m->nn : m->ne);
            float S0[3], O[3];

            if (cfg->method == ORC_BLBADOUEL_GRID) {
                framelen = c->crop0[2];
            }

            ww = currweight - r->weight;
            r->photontimer += r->Lmove * rc;
            tshift = gate_shift(c, r, framelen);

            if (prop[0] > 0.f) {
                r->Eabsorb += ww;
            }

            if (is_fluence_like(cfg->outputtype)) {
                ww = fluence_deposit(currweight, r->weight, prop[0], r->Lmove);
            }

            for (i = 0; i < 3; i++) {
                O[i] = r->vec[i];
                S0[i] = r->p0[i];
                r->p0[i] = S0[i] + O[i] * r->Lmove;
            }

            if (!cfg->basisorder) {
                if (cfg->method == ORC_BLBADOUEL) {
                    unsigned int newidx = eid + tshift;
                    r->oldidx = (r->oldidx == 0xFFFFFFFF) ? newidx : r->oldidx;

                    if (newidx != r->oldidx) {
                        deposit(c, r->oldidx, r->oldweight, r);
                        r->oldidx = newidx;
                        r->oldweight = ww;
                    } else {
                        r->oldweight += ww;
                    }

                    if (r->faceid == -2 || !r->isend) {
                        deposit(c, newidx, r->oldweight, r);
                        r->oldweight = 0.f;
                    }
                } else {
                    /* dual-grid deposit :1616-1672 */
                    float dstep, segloss, w0, Tv[3], Sv[3];
                    int seg = (int)(r->Lmove / cfg->steps) + 1;
                    seg = (seg << 1);
                    dstep = r->Lmove / seg;
                    segloss = expf(-prop[0] * dstep);

                    for (i = 0; i < 3; i++) {
                        Tv[i] = O[i] * dstep;
                        Sv[i] = (S0[i] - m->nmin[i]) + Tv[i] * 0.5f;
                    }

                    dstep = 1.f / cfg->steps;
                    totalloss = (totalloss == 0.f) ? 0.f : (1.f - segloss) / totalloss;
                    w0 = ww;

                    for (i = 0; i < seg; i++) {
                        int ix = (int)(Sv[0] * dstep), iy = (int)(Sv[1] * dstep), iz = (int)(Sv[2] * dstep); /* cvttps */
                        unsigned int newidx = iz * c->crop0[1] + iy * c->crop0[0] + ix + tshift;
                        r->oldidx = (r->oldidx == 0xFFFFFFFF) ? newidx : r->oldidx;

                        if (newidx != r->oldidx) {
                            deposit(c, r->oldidx, r->oldweight, r);
                            r->oldidx = newidx;
                            r->oldweight = w0 * totalloss;
                        } else {
                            r->oldweight += w0 * totalloss;
                        }

                        if (r->faceid == -2 || !r->isend) {
                            deposit(c, newidx, r->oldweight, r);
                            r->oldweight = 0.f;
                        }

                        w0 *= segloss;
                        Sv[0] += Tv[0];
                        Sv[1] += Tv[1];
                        Sv[2] += Tv[2];
                    }
                }
            } else {
                /* nodal: 1/3 to the 3 nodes of the exit face :1675-1690 */
                ww *= 1.f / 3.f;

                if (cfg->srctype != stPattern || cfg->srcnum == 1) {
                    for (i = 0; i < 3; i++) {
                        atomic_add(c->weight + (ee[out_[faceidx][i]] - 1 + tshift), ww);
                    }
                } else {
                    int pidx;

                    for (pidx = 0; pidx < cfg->srcnum; pidx++)
                        for (i = 0; i < 3; i++) {
                            atomic_add(c->weight + (size_t)(ee[out_[faceidx][i]] - 1 + tshift) * cfg->srcnum + pidx,
                                       ww * cfg->srcpattern[(size_t)r->posidx * cfg->srcnum + pidx]);
                        }
                }
            }
        }
    }

    c->raytet++;

    if (r->faceid == -2) {
        return 0.f;
    }

    return r->slen;
}

/* ------------------------------------------------------------------------------------------
 * Havel -- src/mmc_raytrace.c:531-800 (true division instead of rcp+Newton)
 * ---------------------------------------------------------------------------------------- */
static int havel_face(const float* vecN, float* bary, const float* o, const float* d) {
    /* o=(p,1), d=(v,0); :531-561 */
    float det = vecN[0] * d[0] + vecN[1] * d[1] + vecN[2] * d[2];
    float dett, oldt, detp[4], detu, detv;
    union {
        float f;
        uint32_t u;
    } a, b;
    a.f = det;

    if (a.u & 0x80000000U) {
        return 0;
    }

    dett = (-vecN[0] * o[0] + -vecN[1] * o[1]) + (-vecN[2] * o[2] + vecN[3] * o[3]);
    oldt = bary[0];
    a.f = dett;
    b.f = oldt * det - dett;

    if (((a.u ^ b.u) & 0x80000000U) == 0) {
        int k;

        for (k = 0; k < 4; k++) {
            detp[k] = o[k] * det + dett * d[k];
        }

        detu = (detp[0] * vecN[4] + detp[1] * vecN[5]) + (detp[2] * vecN[6] + detp[3] * vecN[7]);
        a.f = detu;
        b.f = det - detu;

        if (((a.u ^ b.u) & 0x80000000U) == 0) {
            detv = (detp[0] * vecN[8] + detp[1] * vecN[9]) + (detp[2] * vecN[10] + detp[3] * vecN[11]);
            a.f = detv;
            b.f = det - (detu + detv);

            if (((a.u ^ b.u) & 0x80000000U) == 0) {
                float inv_det = 1.f / det;
                bary[0] = dett * inv_det;
                bary[1] = detu * inv_det;
                bary[2] = detv * inv_det;
                return (bary[0] == bary[0]);
            }
        }
    }

    return 0;
}

static float havel_raytet(oray* r, octx* c) {
    const orc_mesh* m = c->mesh;
    const orc_config* cfg = c->cfg;
    float bary[4] = {1e10f, 0.f, 0.f, 0.f}, barypout[4], O[4], T[4], S[4];
    float rc, currweight, dlen, ww, Lp0, mus;
    int i, j, k, tshift, eid;
    const int* enb = NULL, *nextenb = NULL;
    const float* prop;
    const int* ee;

    if (m->m == NULL || r->eid <= 0 || r->eid > m->ne) {
        return -1;
    }

    r->p0[3] = 1.f;
    r->vec[3] = 0.f;
    eid = r->eid - 1;
    memcpy(O, r->p0, sizeof(O));
    memcpy(T, r->vec, sizeof(T));
    ee = m->elem + 4 * (size_t)eid;
    prop = medium_of(m, m->type[eid]);
    rc = prop[3] * R_C0;
    mus = prop[1];
    r->pout[0] = ORC_UNDEFINED;
    r->faceid = -1;
    r->isend = 0;
    r->Lmove = 0.f;

    for (i = 0; i < 4; i++)
        if (havel_face(m->m + ((size_t)eid * 12 + i * 3) * 4, bary, O, T)) {
            r->faceid = faceorder_[i];
            dlen = (mus <= EPS) ? R_MIN_MUS : r->slen / mus;
            Lp0 = bary[0];
            r->isend = (Lp0 > dlen);
            r->Lmove = ((r->isend) ? dlen : Lp0);

            if (!r->isend) {
                enb = m->facenb + 4 * (size_t)eid;
                r->nexteid = enb[r->faceid];

                if (r->nexteid > 0) {
                    nextenb = m->elem + 4 * (size_t)(r->nexteid - 1);
                }
            }

            for (k = 0; k < 3; k++) {
                r->pout[k] = bary[0] * T[k] + O[k];
            }

            currweight = step_common(c, r, prop, mus, 1);

            if (bary[0] == 0.f) {
                break;
            }

            ww = currweight - r->weight;
            r->photontimer += r->Lmove * rc;

            if (prop[0] > 0.f) {
                r->Eabsorb += ww;
            }

            if (is_fluence_like(cfg->outputtype)) {
                ww = fluence_deposit(currweight, r->weight, prop[0], r->Lmove);
            }

            tshift = gate_shift(c, r, cfg->basisorder ? m->nn : m->ne);

            for (k = 0; k < 3; k++) {
                r->p0[k] = O[k] + T[k] * r->Lmove;
            }

            barypout[out_[i][0]] = 1.f - bary[1] - bary[2];
            barypout[out_[i][1]] = bary[1];
            barypout[out_[i][2]] = bary[2];
            barypout[facemap_[i]] = 0.f;
            dlen = r->Lmove / bary[0];

            for (k = 0; k < 4; k++) {
                float t = barypout[k], o = r->bary0[k];
                S[k] = r->isend ? (t * dlen + o * (1.f - dlen)) : t;
                O[k] = o;   /* O now holds bary at p0 */
            }

            memcpy(barypout, S, sizeof(S));

            if (nextenb && enb) {
                memset(r->bary0, 0, sizeof(r->bary0));

                for (j = 0; j < 3; j++)
                    for (k = 0; k < 4; k++) {
                        if (ee[out_[i][j]] == nextenb[k]) {
                            r->bary0[k] = barypout[out_[i][j]];
                            break;
                        }
                    }
            } else {
                memcpy(r->bary0, S, sizeof(S));
            }

            if (!cfg->basisorder) {
                deposit(c, eid + tshift, ww, r);
            } else {
                for (k = 0; k < 4; k++) {
                    barypout[k] = (O[k] + S[k]) * (ww * 0.5f);
                }

                if (cfg->srctype != stPattern || cfg->srcnum == 1) {
                    for (j = 0; j < 4; j++) {
                        atomic_add(c->weight + (ee[j] - 1 + tshift), barypout[j]);
                    }
                } else {
                    int pidx;

                    for (pidx = 0; pidx < cfg->srcnum; pidx++)
                        for (j = 0; j < 4; j++) {
                            atomic_add(c->weight + (size_t)(ee[j] - 1 + tshift) * cfg->srcnum + pidx,
                                       barypout[j] * cfg->srcpattern[(size_t)r->posidx * cfg->srcnum + pidx]);
                        }
                }
            }

            break;
        }

    c->raytet++;
    r->p0[3] = 0.f;

    if (r->faceid == -2) {
        return 0.f;
    }

    return r->slen;
}

/* ------------------------------------------------------------------------------------------
 * Plucker -- src/mmc_raytrace.c:227-508
 * ---------------------------------------------------------------------------------------- */
static float plucker_raytet(oray* r, octx* c) {
    const orc_mesh* m = c->mesh;
    const orc_config* cfg = c->cfg;
    float pcrx[3], p1[3], w[6], Rv, ww, currweight, dlen = 0.f, rc, mus, Lp0 = 0.f, ratio;
    float baryout[4] = {0.f, 0.f, 0.f, 0.f}, *baryp0 = r->bary0;
    int i, tshift, eid, faceidx = -1;
    const int* ee;
    const float* prop;
    union {
        float f;
        uint32_t u;
    } w0, w1, w2;

    if (m->pd == NULL || r->eid <= 0 || r->eid > m->ne) {
        return -1;
    }

    eid = r->eid - 1;
    r->faceid = -1;
    r->isend = 0;
    r->Lmove = 0.f;

    for (i = 0; i < 3; i++) {
        p1[i] = r->p0[i] + r->vec[i];
    }

    cross3(r->p0, p1, pcrx);
    ee = m->elem + 4 * (size_t)eid;
    prop = medium_of(m, m->type[eid]);
    rc = prop[3] * R_C0;
    currweight = r->weight;
    mus = prop[1];

    for (i = 0; i < 6; i++) {
        const float* D = m->pd + ((size_t)eid * 6 + i) * 4, *M = m->pm + ((size_t)eid * 6 + i) * 4;
        w[i] = dot3(r->vec, M) + dot3(pcrx, D);
    }

    r->pout[0] = ORC_UNDEFINED;

    for (i = 0; i < 4; i++) {
        if (i >= 2) {
            w[fc_[i][1]] = -w[fc_[i][1]];
        }

        w0.f = w[fc_[i][0]];
        w1.f = w[fc_[i][1]];
        w2.f = w[fc_[i][2]];

        if ((w0.u & 0x80000000U) & (w1.u & 0x80000000U) & ((w2.u ^ 0x80000000U))) {
            const float* q0 = nd(m, ee[nc_[i][0]]), *q1 = nd(m, ee[nc_[i][1]]), *q2 = nd(m, ee[nc_[i][2]]);
            int k;
            Rv = 1.f / (-w[fc_[i][0]] - w[fc_[i][1]] + w[fc_[i][2]]);
            baryout[nc_[i][0]] = -w[fc_[i][0]] * Rv;
            baryout[nc_[i][1]] = -w[fc_[i][1]] * Rv;
            baryout[nc_[i][2]] = w[fc_[i][2]] * Rv;

            for (k = 0; k < 3; k++) { /* getinterp src/mmc_raytrace.c:165-169 */
                r->pout[k] = baryout[nc_[i][0]] * q0[k] + baryout[nc_[i][1]] * q1[k] + baryout[nc_[i][2]] * q2[k];
            }

            Lp0 = sqrtf((r->pout[0] - r->p0[0]) * (r->pout[0] - r->p0[0]) + (r->pout[1] - r->p0[1]) * (r->pout[1] - r->p0[1]) +
                        (r->pout[2] - r->p0[2]) * (r->pout[2] - r->p0[2]));
            dlen = (mus <= EPS) ? R_MIN_MUS : r->slen / mus;
            faceidx = i;
            r->faceid = faceorder_[i];
            r->nexteid = m->facenb[4 * (size_t)eid + r->faceid];
            r->isend = (Lp0 > dlen);
            r->Lmove = ((r->isend) ? dlen : Lp0);
            break;
        }
    }

    c->raytet++;

    if (r->pout[0] != ORC_UNDEFINED) {
        currweight = step_common(c, r, prop, mus, 1);
        r->p0[0] += r->Lmove * r->vec[0];
        r->p0[1] += r->Lmove * r->vec[1];
        r->p0[2] += r->Lmove * r->vec[2];

        if (!cfg->basisorder) {
            ww = currweight - r->weight;
            r->Eabsorb += ww;

            if (is_fluence_like(cfg->outputtype)) {
                ww = fluence_deposit(currweight, r->weight, prop[0], r->Lmove);
            }

            r->photontimer += r->Lmove * rc;
            tshift = gate_shift(c, r, m->ne);

            if (cfg->srctype != stPattern || cfg->srcnum == 1) {
                atomic_add(c->weight + (eid + tshift), ww);
            } else {
                int pidx;

                for (pidx = 0; pidx < cfg->srcnum; pidx++) {
                    atomic_add(c->weight + (size_t)(eid + tshift) * cfg->srcnum + pidx, ww * cfg->srcpattern[(size_t)r->posidx * cfg->srcnum + pidx]);
                }
            }
        } else {
            if (Lp0 > EPS) {
                r->photontimer += r->Lmove * rc;
                ww = currweight - r->weight;

                if (prop[0] > 0.f || is_fluence_like(cfg->outputtype)) {
                    ratio = r->Lmove / Lp0;

                    if (prop[0] > 0.f) {
                        r->Eabsorb += ww;
                    }

                    if (is_fluence_like(cfg->outputtype)) {
                        ww = fluence_deposit(currweight, r->weight, prop[0], r->Lmove);
                    }

                    tshift = gate_shift(c, r, m->nn);
                    ww *= 0.5f;

                    if (r->isend) {
                        for (i = 0; i < 4; i++) {
                            baryout[i] = (1.f - ratio) * baryp0[i] + ratio * baryout[i];
                        }
                    }

                    if (cfg->srctype != stPattern || cfg->srcnum == 1) {
                        for (i = 0; i < 4; i++) {
                            atomic_add(c->weight + (ee[i] - 1 + tshift), ww * (baryp0[i] + baryout[i]));
                        }
                    } else {
                        int pidx;

                        for (pidx = 0; pidx < cfg->srcnum; pidx++)
                            for (i = 0; i < 4; i++) {
                                atomic_add(c->weight + (size_t)(ee[i] - 1 + tshift) * cfg->srcnum + pidx,
                                           ww * cfg->srcpattern[(size_t)r->posidx * cfg->srcnum + pidx] * (baryp0[i] + baryout[i]));
                            }
                    }
                }

                if (r->isend) {
                    memcpy(baryp0, baryout, 4 * sizeof(float));
                } else if (r->nexteid > 0 && faceidx >= 0) {
                    int j, k;
                    const int* nextenb = m->elem + 4 * (size_t)(r->nexteid - 1);
                    memset(baryp0, 0, 4 * sizeof(float));

                    for (j = 0; j < 3; j++)
                        for (k = 0; k < 4; k++) {
                            if (ee[nc_[faceidx][j]] == nextenb[k]) {
                                baryp0[k] = baryout[nc_[faceidx][j]];
                                break;
                            }
                        }
                }
            }
        }

        if (r->faceid == -2) {
            return 0.f;
        }
    }

    return r->slen;
}

/* ------------------------------------------------------------------------------------------
 * Fresnel -- src/mmc_raytrace.c:2257-2332
 * ---------------------------------------------------------------------------------------- */
static float reflectray(octx* c, float* c0, int* oldeid, int* eid, int faceid, uint64_t* ran) {
    const orc_mesh* m = c->mesh;
    const orc_config* cfg = c->cfg;
    float pn[3], Icos, Re, Im, Rtotal, tmp0, tmp1, tmp2, n1, n2;
    int offs = (*oldeid - 1) << 2, k;
    faceid = ifaceorder_[faceid];

    if (cfg->method == ORC_PLUCKER || cfg->method == ORC_BLBADOUEL || cfg->method == ORC_BLBADOUEL_GRID) {
        const float* nb = m->n + (size_t)offs * 4;
        pn[0] = nb[faceid];
        pn[1] = nb[faceid + 4];
        pn[2] = nb[faceid + 8];
    } else {
        const float* mb = m->m + ((size_t)(offs + faceid) * 3) * 4;
        pn[0] = mb[0];
        pn[1] = mb[1];
        pn[2] = mb[2];
    }

    Icos = fabs(dot3(c0, pn));
    n1 = (*oldeid != *eid) ? medium_of(m, m->type[*oldeid - 1])[3] : cfg->nout;
    n2 = (*eid > 0) ? medium_of(m, m->type[*eid - 1])[3] : cfg->nout;
    tmp0 = n1 * n1;
    tmp1 = n2 * n2;
    tmp2 = 1.f - tmp0 / tmp1 * (1.f - Icos * Icos);

    if (tmp2 > 0.f && !(*eid <= 0 && cfg->isreflect == ORC_BC_MIRROR)) {
        Re = tmp0 * Icos * Icos + tmp1 * tmp2;
        tmp2 = sqrtf(tmp2);
        Im = 2.f * n1 * n2 * Icos * tmp2;
        Rtotal = (Re - Im) / (Re + Im);
        Re = tmp1 * Icos * Icos + tmp0 * tmp2 * tmp2;
        Rtotal = (Rtotal + (Re - Im) / (Re + Im)) * 0.5f;

        if (*oldeid == *eid) {
            return Rtotal;
        }

        if (rand01(ran) <= Rtotal) {
            for (k = 0; k < 3; k++) {
                c0[k] = 1.f * c0[k] + (-2.f * Icos) * pn[k];
            }

            *eid = *oldeid;
        } else if (cfg->isspecular == 2 && *eid == 0) {
        } else {
            for (k = 0; k < 3; k++) {
                c0[k] = 1.f * c0[k] + (-Icos) * pn[k];
            }

            for (k = 0; k < 3; k++) {
                c0[k] = (n1 / n2) * c0[k] + tmp2 * pn[k];
            }
        }
    } else {
        for (k = 0; k < 3; k++) {
            c0[k] = 1.f * c0[k] + (-2.f * Icos) * pn[k];
        }

        *eid = *oldeid;
    }

    tmp0 = 1.f / sqrtf(dot3(c0, c0));

    for (k = 0; k < 3; k++) {
        c0[k] = tmp0 * c0[k];
    }

    return 1.f;
}

/* src/mmc_raytrace.c:171-199 rotatevector (launch) + src/mmc_mesh.c:1693-1705 (scatter) */
static void rotatevector(float* dir, float stheta, float ctheta, float sphi, float cphi, int renorm) {
    float p[3];

    if (dir[2] > -1.f + EPS && dir[2] < 1.f - EPS) {
        float tmp0 = 1.f - dir[2] * dir[2];
        float tmp1 = 1.f / sqrtf(tmp0);
        tmp1 = stheta * tmp1;
        p[0] = tmp1 * (dir[0] * dir[2] * cphi - dir[1] * sphi) + dir[0] * ctheta;
        p[1] = tmp1 * (dir[1] * dir[2] * cphi + dir[0] * sphi) + dir[1] * ctheta;
        p[2] = -tmp1 * tmp0 * cphi + dir[2] * ctheta;
    } else {
        p[0] = stheta * cphi;
        p[1] = stheta * sphi;
        p[2] = (dir[2] > 0.f) ? ctheta : -ctheta;
    }

    if (renorm) {  /* GPU only: src/mmc_core.cl:1325-1329 */
        float t = 1.f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        p[0] *= t;
        p[1] *= t;
        p[2] *= t;
    }

    dir[0] = p[0];
    dir[1] = p[1];
    dir[2] = p[2];
}

/* src/mmc_mesh.c:1644-1715 */
static float mc_next_scatter(octx* c, float g, float* dir, uint64_t* ran, float* pmom) {
    float nextslen, sphi, cphi, tmp0, theta, stheta, ctheta;
    nextslen = rand_scatlen(ran);
    tmp0 = TWO_PI * rand01(ran);
    sincosf(tmp0, &sphi, &cphi);

    if (g > EPS) {
        tmp0 = (1.f - g * g) / (1.f - g + 2.f * g * rand01(ran));
        tmp0 *= tmp0;
        tmp0 = (1.f + g * g - tmp0) / (2.f * g);

        if (tmp0 > 1.f) {
            tmp0 = 1.f;
        }

        if (tmp0 < -1.f) {
            tmp0 = -1.f;
        }

        stheta = sqrt(1.f - tmp0 * tmp0);
        ctheta = tmp0;
    } else {
        theta = acosf(2.f * rand01(ran) - 1.f);
        sincosf(theta, &stheta, &ctheta);
    }

    rotatevector(dir, stheta, ctheta, sphi, cphi, c->cfg->gpu_semantics);

    if (c->cfg->ismomentum) {
        pmom[0] += (1.f - ctheta);
    }

    return nextslen;
}

/* src/mmc_raytrace.c:201-213 */
static void fixphoton(float* p, const orc_mesh* m, const int* ee) {
    float c0[3] = {0.f, 0.f, 0.f};
    int i;

    for (i = 0; i < 4; i++) {
        const float* q = nd(m, ee[i]);
        c0[0] = c0[0] + q[0];
        c0[1] = c0[1] + q[1];
        c0[2] = c0[2] + q[2];
    }

    p[0] += (c0[0] * 0.25f - p[0]) * FIX_PHOTON;
    p[1] += (c0[1] * 0.25f - p[1]) * FIX_PHOTON;
    p[2] += (c0[2] * 0.25f - p[2]) * FIX_PHOTON;
}

/* ------------------------------------------------------------------------------------------
 * launch -- src/mmc_raytrace.c:2346-2671 (source types 0-13; CPU-only types 14-17 not restated)
 * ---------------------------------------------------------------------------------------- */
static int launchphoton(octx* c, oray* r, uint64_t* ran) {
    const orc_mesh* m = c->mesh;
    orc_config* cfg = c->cfg;
    int canfocus = 1, k;
    float origin[3] = {r->p0[0], r->p0[1], r->p0[2]};
    const float* sp1 = cfg->srcparam1, *sp2 = cfg->srcparam2, *sd = cfg->srcdir, *sp = cfg->srcpos;
    int st = cfg->srctype;

    r->slen = rand_scatlen(ran);

    if (st == stPencil) {
        if (r->eid > 0) {
            return 0;
        }
    } else if (st == stPlanar || st == stPattern || st == stFourier) {
        float rx = rand01(ran);
        float ry = rand01(ran);

        for (k = 0; k < 3; k++) {
            r->p0[k] = sp[k] + rx * sp1[k] + ry * sp2[k];
        }

        r->weight = 1.f;

        if (st == stPattern) {
            int xsize = (int)sp1[3], ysize = (int)sp2[3];
            r->posidx = MINI((int)(ry * ysize), ysize - 1) * xsize + MINI((int)(rx * xsize), xsize - 1);
            r->weight = (cfg->srcnum == 1) ? cfg->srcpattern[r->posidx] : 1.f;

            if (cfg->seed == ORC_SEED_FROM_FILE && (cfg->outputtype == ORC_WL || cfg->outputtype == ORC_WP)) {
                r->weight = cfg->srcpattern[MINI((int)(ry * sp2[3]), (int)sp2[3] - 1) * (int)(sp1[3]) + MINI((int)(rx * sp1[3]), (int)sp1[3] - 1)];
                cfg->replayweight[r->photonid] *= r->weight;
            }
        } else if (st == stFourier) {
            r->weight = (cosf((floorf(sp1[3]) * rx + floorf(sp2[3]) * ry + sp1[3] - floorf(sp1[3])) * TWO_PI) * (1.f - sp2[3] + floorf(sp2[3])) + 1.f) * 0.5f;
        }

        for (k = 0; k < 3; k++) {
            origin[k] += (sp1[k] + sp2[k]) * 0.5f;
        }
    } else if (st == stFourierX || st == stFourier2D) {
        float rx = rand01(ran);
        float ry = rand01(ran);
        float v2[4] = {sp1[0], sp1[1], sp1[2], sp1[3]};
        v2[3] *= 1.f / (sqrtf(sp1[0] * sp1[0] + sp1[1] * sp1[1] + sp1[2] * sp1[2]));
        v2[0] = v2[3] * (sd[1] * sp1[2] - sd[2] * sp1[1]);
        v2[1] = v2[3] * (sd[2] * sp1[0] - sd[0] * sp1[2]);
        v2[2] = v2[3] * (sd[0] * sp1[1] - sd[1] * sp1[0]);

        for (k = 0; k < 3; k++) {
            r->p0[k] = sp[k] + rx * sp1[k] + ry * v2[k];
        }

        if (st == stFourier2D) {
            r->weight = (sinf((sp2[0] * rx + sp2[2]) * TWO_PI) * sinf((sp2[1] * ry + sp2[3]) * TWO_PI) + 1.f) * 0.5f;
        } else {
            r->weight = (cosf((sp2[0] * rx + sp2[1] * ry + sp2[2]) * TWO_PI) * (1.f - sp2[3]) + 1.f) * 0.5f;
        }

        for (k = 0; k < 3; k++) {
            origin[k] += (sp1[k] + v2[k]) * 0.5f;
        }
    } else if (st == stDisk || st == stGaussian) {
        float sphi, cphi, phi, r0;
        phi = TWO_PI * rand01(ran);
        sphi = sinf(phi);
        cphi = cosf(phi);

        if (st == stDisk) {
            r0 = sqrtf(rand01(ran) * fabsf(sp1[0] * sp1[0] - sp1[1] * sp1[1]) + sp1[1] * sp1[1]);
        } else if (fabs(r->focus) < 1e-5f || fabs(sp1[1]) < 1e-5f) {
            r0 = sqrtf(-log(rand01(ran))) * sp1[0];
        } else {
            float z0 = sp1[0] * sp1[0] * M_PI / sp1[1];
            r0 = sqrtf(-log(rand01(ran)) * (1.f + (r->focus * r->focus / (z0 * z0)))) * sp1[0];
        }

        if (sd[2] > -1.f + EPS && sd[2] < 1.f - EPS) {
            float tmp0 = 1.f - sd[2] * sd[2];
            float tmp1 = r0 / sqrtf(tmp0);
            r->p0[0] = sp[0] + tmp1 * (sd[0] * sd[2] * cphi - sd[1] * sphi);
            r->p0[1] = sp[1] + tmp1 * (sd[1] * sd[2] * cphi + sd[0] * sphi);
            r->p0[2] = sp[2] - tmp1 * tmp0 * cphi;
        } else {
            r->p0[0] += r0 * cphi;
            r->p0[1] += r0 * sphi;
        }
    } else if (st == stCone || st == stIsotropic || st == stArcSin) {
        float ang, stheta, ctheta, sphi, cphi;
        ang = TWO_PI * rand01(ran);
        sphi = sinf(ang);
        cphi = cosf(ang);

        if (st == stCone) {
            do {
                ang = (sp1[1] > 0) ? TWO_PI * rand01(ran) : acosf(2.f * rand01(ran) - 1.f);
            } while (ang > sp1[0]);
        } else if (st == stIsotropic) {
            ang = acosf(2.f * rand01(ran) - 1.f);
        } else {
            ang = M_PI * rand01(ran);
        }

        stheta = sinf(ang);
        ctheta = cosf(ang);

        if (cfg->gpu_semantics) {   /* src/mmc_core.cl:1676-1678 */
            r->vec[0] = stheta * cphi;
            r->vec[1] = stheta * sphi;
            r->vec[2] = ctheta;
        } else {
            rotatevector(r->vec, stheta, ctheta, sphi, cphi, 1);   /* src/mmc_mesh.h:337-358 renormalises (with rsqrtss; exact here) */
        }

        canfocus = 0;

        if (r->eid > 0 && (!cfg->gpu_semantics || st == stIsotropic)) {
            return 0;
        }
    } else if (st == stZGaussian) {
        float ang, stheta, ctheta, sphi, cphi;
        ang = TWO_PI * rand01(ran);
        sphi = sinf(ang);
        cphi = cosf(ang);
        /* the reference draws the 2nd number from the unused ran0 stream (zero state => 0.f) :2492 */
        ang = sqrtf(-2.f * log(rand01(ran))) * (1.f - 2.f * (cfg->gpu_semantics ? rand01(ran) : 0.f)) * sp1[0];
        stheta = sinf(ang);
        ctheta = cosf(ang);

        if (cfg->gpu_semantics) {
            r->vec[0] = stheta * cphi;
            r->vec[1] = stheta * sphi;
            r->vec[2] = ctheta;
        } else {
            rotatevector(r->vec, stheta, ctheta, sphi, cphi, 1);   /* src/mmc_mesh.h:337-358 renormalises (with rsqrtss; exact here) */
        }

        canfocus = 0;
    } else if (st == stLine || st == stSlit) {
        float t = rand01(ran);

        for (k = 0; k < 3; k++) {
            r->p0[k] += t * sp1[k];
        }

        if (st == stLine) {
            float s, p, vv[3];
            t = 1.f - 2.f * rand01(ran);
            s = 1.f - 2.f * rand01(ran);
            p = sqrtf(1.f - r->vec[0] * r->vec[0] - r->vec[1] * r->vec[1]) * (rand01(ran) > 0.5f ? 1.f : -1.f);
            vv[0] = r->vec[1] * p - r->vec[2] * s;
            vv[1] = r->vec[2] * t - r->vec[0] * p;
            vv[2] = r->vec[0] * s - r->vec[1] * t;
            memcpy(r->vec, vv, sizeof(vv));
        }

        for (k = 0; k < 3; k++) {
            origin[k] += sp1[k] * 0.5f;
        }

        canfocus = (st == stSlit);
    } else {
        ORC_FAIL("source type %d is not restated by the oracle", st);
    }

    if (canfocus && cfg->gpu_semantics && (isnan(r->focus) || (r->focus < 0.f && isinf(r->focus)))) {
        /* GPU-only launch modes of the wide-field sources (src/mmc_core.cl:1743-1757; the CPU file has neither): focal length NaN = isotropic
         * directions, -inf = Lambertian (cosine-weighted) about srcdir */
        float ang = TWO_PI * rand01(ran), sphi = sinf(ang), cphi = cosf(ang), stheta, ctheta;

        if (isnan(r->focus)) {
            ang = acosf(2.f * rand01(ran) - 1.f);
            stheta = sinf(ang);
            ctheta = cosf(ang);
        } else {
            stheta = sqrtf(rand01(ran));
            ctheta = sqrtf(1.f - stheta * stheta);
        }

        rotatevector(r->vec, stheta, ctheta, sphi, cphi, 1);
    } else if (canfocus && r->focus != 0.f) {
        float Rn2;

        for (k = 0; k < 3; k++) {
            origin[k] += r->focus * r->vec[k];
        }

        for (k = 0; k < 3; k++) {
            r->vec[k] = (r->focus < 0.f) ? (r->p0[k] - origin[k]) : (origin[k] - r->p0[k]);
        }

        Rn2 = 1.f / sqrtf(dot3(r->vec, r->vec));

        for (k = 0; k < 3; k++) {
            r->vec[k] = Rn2 * r->vec[k];
        }
    }

    for (k = 0; k < 3; k++) {
        r->p0[k] = 1.f * r->p0[k] + EPS * r->vec[k];   /* vec_mult_add(p0,vec,1,EPS) :2591 */
    }

    /* enclosing element search :2593-2670 */
    {
        int is, i;
        float bary[4] = {0.f, 0.f, 0.f, 0.f};

        for (is = -1; is < m->srcelemlen; is++) {
            int include = 1;
            const int* ee;

            if (is < 0) {
                if (r->eid >= 0) {
                    if (r->eid == 0) {
                        continue;    /* the reference would read elem[-1]; treat e0==0 as "not set" */
                    }

                    ee = m->elem + 4 * (size_t)(r->eid - 1);
                } else {
                    continue;
                }
            } else {
                ee = m->elem + 4 * (size_t)(m->srcelem[is] - 1);
            }

            for (i = 0; i < 4; i++) {
                float AB[3], AC[3], S[3], N[3];
                const float* a = nd(m, ee[out_[i][0]]), *b = nd(m, ee[out_[i][1]]), *cc = nd(m, ee[out_[i][2]]);
                diff3(a, b, AB);
                diff3(a, cc, AC);
                diff3(a, r->p0, S);
                cross3(AB, AC, N);
                bary[facemap_[i]] = -dot3(S, N);
            }

            for (i = 0; i < 4; i++)
                if (bary[i] < -1e-4f) {
                    include = 0;
                }

            if (include) {
                float s = 0.f;
                r->eid = (is >= 0 ? m->srcelem[is] : r->eid);

                for (i = 0; i < 4; i++) {
                    s += bary[i];
                }

                for (i = 0; i < 4; i++) {
                    r->bary0[i] = bary[i] / s;
                }

                for (i = 0; i < 4; i++)
                    if ((bary[i] / s) < 1e-4f) {
                        r->faceid = ifacemap_[i] + 1;
                    }

                break;
            }
        }

        if (is == m->srcelemlen) {
            ORC_FAIL("initial element does not enclose the source!");
        }
    }

    return 0;
}

static void savedebug(octx* c, const oray* r, unsigned int id) {
    unsigned int pos;
    #pragma omp atomic capture
    pos = (*c->trajcount)++;

    if (pos < c->cfg->maxjumpdebug) {
        float* d = c->traj + (size_t)pos * 6;
        memcpy(d, &id, 4);
        d[1] = r->p0[0];
        d[2] = r->p0[1];
        d[3] = r->p0[2];
        d[4] = r->weight;
        memcpy(d + 5, &r->eid, 4);
    }
}

/* ------------------------------------------------------------------------------------------
 * one photon -- src/mmc_raytrace.c:1772-2126 (CPU flow); gpu_semantics follows src/mmc_core.cl:1851-2161
 * ---------------------------------------------------------------------------------------- */
static int onephoton(uint64_t id, octx* c, uint64_t* ran) {
    const orc_mesh* m = c->mesh;
    orc_config* cfg = c->cfg;
    int oldeid = 0, fixcount = 0, exitdet = 0, k;
    float mom;
    double kahany, kahant;
    uint64_t initseed[2] = {ran[0], ran[1]};
    oray r;
    float (*tracer)(oray*, octx*);
    float ppath[64];   /* reclen-1 floats */

    memset(&r, 0, sizeof(r));
    memcpy(r.p0, cfg->srcpos, sizeof(float) * 4);
    r.vec[0] = cfg->srcdir[0];
    r.vec[1] = cfg->srcdir[1];
    r.vec[2] = cfg->srcdir[2];
    r.pout[0] = ORC_UNDEFINED;
    memcpy(r.bary0, c->bary0, sizeof(float) * 4);   /* cfg->bary0 */
    r.eid = cfg->e0;
    r.faceid = -1;
    r.weight = 1.f;
    r.focus = cfg->srcdir[3];
    r.oldidx = 0xFFFFFFFF;
    r.photonid = (unsigned int)id;
    memset(ppath, 0, sizeof(ppath));
    r.partialpath = ppath;

    tracer = (cfg->method == ORC_PLUCKER) ? plucker_raytet : (cfg->method == ORC_HAVEL ? havel_raytet : blb_raytet);

    if (launchphoton(c, &r, ran)) {
        return -1;
    }

    if (cfg->savetraj) {
        savedebug(c, &r, (unsigned int)id);
    }

    if (cfg->srctype != stPattern || cfg->srcnum == 1) {
        r.partialpath[c->reclen - 2] = r.weight;

        /* CPU: the launched "energy" of a wl/wp replay is the detected weight (src/mmc_raytrace.c:1811-1816); the CUDA
         * kernel adds r.weight = 1 for every photon (src/mmc_core.cl:1907-1908) */
        if (cfg->seed == ORC_SEED_FROM_FILE && (cfg->outputtype == ORC_WL || cfg->outputtype == ORC_WP) && !cfg->gpu_semantics) {
            kahany = cfg->replayweight[r.photonid] - c->kahanc0[0];
        } else {
            kahany = r.weight - c->kahanc0[0];
        }

        kahant = c->launchweight[0] + kahany;
        c->kahanc0[0] = (kahant - c->launchweight[0]) - kahany;
        c->launchweight[0] = kahant;
    } else {
        int pidx;
        memcpy(r.partialpath + c->reclen - 2, &r.posidx, 4);

        for (pidx = 0; pidx < cfg->srcnum; pidx++) {
            kahany = r.weight * cfg->srcpattern[(size_t)r.posidx * cfg->srcnum + pidx] - c->kahanc0[pidx];
            kahant = c->launchweight[pidx] + kahany;
            c->kahanc0[pidx] = (kahant - c->launchweight[pidx]) - kahany;
            c->launchweight[pidx] = kahant;
        }
    }

    while (1) {
        r.slen = tracer(&r, c);

        if (r.pout[0] == ORC_UNDEFINED) {
            if (cfg->gpu_semantics && r.faceid == -2) {
                break;    /* src/mmc_core.cl:1928-1930 */
            }

            if (fixcount++ < MAX_TRIAL) {
                fixphoton(r.p0, m, m->elem + 4 * (size_t)(r.eid - 1));
                continue;
            }

            r.eid = ID_UNDEFINED;
            r.faceid = -1;
        }

        if (cfg->issavedet && r.Lmove > 0.f && r.eid != ID_UNDEFINED && m->type[r.eid - 1] > 0 && (r.faceid >= 0 || cfg->gpu_semantics)) {
            r.partialpath[m->prop - 1 + m->type[r.eid - 1]] += r.Lmove;
        }

        if (r.faceid == -2) {
            break;
        }

        while (r.faceid >= 0 && !r.isend) {
            memcpy(r.p0, r.pout, sizeof(float) * 3);
            oldeid = r.eid;
            r.eid = m->facenb[4 * (size_t)(r.eid - 1) + r.faceid];

            if (cfg->isreflect && (r.eid <= 0 || medium_of(m, m->type[r.eid - 1])[3] != medium_of(m, m->type[oldeid - 1])[3])) {
                if (!(r.eid <= 0 && ((medium_of(m, m->type[oldeid - 1])[3] == cfg->nout && cfg->isreflect != ORC_BC_MIRROR) || cfg->isreflect == ORC_BC_ABSORB_EXTERIOR))) {
                    reflectray(c, r.vec, &oldeid, &r.eid, r.faceid, ran);
                }
            }

            if (r.eid <= 0) {
                break;
            }

            if (m->type[oldeid - 1] == 0 && m->type[r.eid - 1]) {
                if (!cfg->voidtime) {
                    r.photontimer = 0.f;
                }
            }

            if (m->type[oldeid - 1] && m->type[r.eid - 1] == 0) {
                if (!m->isextdet) {
                    r.eid = 0;
                    break;
                }
            }

            r.slen = tracer(&r, c);

            if (cfg->issavedet && r.Lmove > 0.f && m->type[r.eid - 1] > 0) {
                r.partialpath[m->prop - 1 + m->type[r.eid - 1]] += r.Lmove;
            }

            if (cfg->gpu_semantics && r.faceid == -2) {
                break;    /* src/mmc_core.cl:2007-2009 */
            }

            fixcount = 0;

            while (r.pout[0] == ORC_UNDEFINED && fixcount++ < MAX_TRIAL) {
                fixphoton(r.p0, m, m->elem + 4 * (size_t)(r.eid - 1));
                r.slen = tracer(&r, c);

                if (cfg->issavedet && r.Lmove > 0.f && m->type[r.eid - 1] > 0) {
                    r.partialpath[m->prop - 1 + m->type[r.eid - 1]] += r.Lmove;
                }
            }

            if (r.pout[0] == ORC_UNDEFINED) {
                r.eid = ID_UNDEFINED;
                break;
            }
        }

        if (r.eid <= 0 || r.pout[0] == ORC_UNDEFINED) {
            if (r.eid != ID_UNDEFINED) {
                if (cfg->issavedet && cfg->issaveexit) {
                    memcpy(r.partialpath + (c->reclen - 2 - 6), r.p0, sizeof(float) * 3);
                    memcpy(r.partialpath + (c->reclen - 2 - 3), r.vec, sizeof(float) * 3);
                }

                if (cfg->issaveref && r.eid < 0 && c->dref) {
                    int tshift = MINI(((int)((r.photontimer - cfg->tstart) * c->rtstep)), c->maxgate - 1) * m->nf;
                    atomic_add(c->dref + (((-r.eid) - 1) + tshift), r.weight);
                }
            }

            if (cfg->issavedet && r.eid <= 0) {
                int i;

                if (cfg->detnum == 0 && m->isextdet && m->type[oldeid - 1] == m->prop + 1) {
                    exitdet = oldeid;
                } else
                    for (i = 0; i < cfg->detnum; i++) {
                        const float* d = cfg->detpos + 4 * i;

                        if ((d[0] - r.p0[0]) * (d[0] - r.p0[0]) + (d[1] - r.p0[1]) * (d[1] - r.p0[1]) + (d[2] - r.p0[2]) * (d[2] - r.p0[2]) < d[3] * d[3]) {
                            exitdet = i + 1;
                            break;
                        }
                    }
            }

            break;
        }

        if (cfg->minenergy > 0.f && r.weight < cfg->minenergy && (cfg->tend - cfg->tstart) * c->rtstep <= 1.f) {
            if (rand01(ran) * cfg->roulettesize <= 1.f) {
                r.weight *= cfg->roulettesize;
            } else {
                break;
            }
        }

        mom = 0.f;
        r.slen0 = mc_next_scatter(c, medium_of(m, m->type[r.eid - 1])[2], r.vec, ran, &mom);
        r.slen = r.slen0;

        if (cfg->savetraj) {
            savedebug(c, &r, (unsigned int)id);
        }

        if (cfg->ismomentum && m->type[r.eid - 1] > 0) {
            r.partialpath[(m->prop << 1) - 1 + m->type[r.eid - 1]] += mom;
        }

        if (m->type[r.eid - 1] > 0 || !cfg->gpu_semantics) {
            /* the CPU file increments partialpath[type-1] unconditionally (:2077); type==0 would hit index -1 */
            if (m->type[r.eid - 1] > 0) {
                r.partialpath[m->type[r.eid - 1] - 1]++;
            }
        }
    }

    if (cfg->issavedet && exitdet > 0) {
        unsigned int pos;
        #pragma omp atomic capture
        pos = (*c->detcount)++;

        if (pos < c->detcap) {
            float* rec = c->detbuf + (size_t)pos * c->reclen;
            rec[0] = exitdet;
            memcpy(rec + 1, r.partialpath, (c->reclen - 1) * sizeof(float));

            if (cfg->issaveseed && c->seedbuf) {
                c->seedbuf[2 * (size_t)pos] = initseed[0];
                c->seedbuf[2 * (size_t)pos + 1] = initseed[1];
            }
        }
    }

    if (cfg->savetraj) {
        savedebug(c, &r, (unsigned int)id);
    }

    /* tallies :2113-2125 (+ GPU-style escaped weight, src/mmc_core.cl:2150-2160) */
    if (cfg->srctype != stPattern || cfg->srcnum == 1) {
        kahany = r.Eabsorb - c->kahanc1[0];
        kahant = c->absorbweight[0] + kahany;
        c->kahanc1[0] = (kahant - c->absorbweight[0]) - kahany;
        c->absorbweight[0] = kahant;
        c->escweight[0] += r.weight;
    } else {
        int pidx;

        for (pidx = 0; pidx < cfg->srcnum; pidx++) {
            float pw = cfg->srcpattern[(size_t)r.posidx * cfg->srcnum + pidx];
            kahany = r.Eabsorb * pw - c->kahanc1[pidx];
            kahant = c->absorbweight[pidx] + kahany;
            c->kahanc1[pidx] = (kahant - c->absorbweight[pidx]) - kahany;
            c->absorbweight[pidx] = kahant;
            c->escweight[pidx] += r.weight * pw;
        }
    }

    (void)k;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * normalisation -- src/mmc_mesh.c:2154-2279 (single-source / pattern `pair`)
 * ---------------------------------------------------------------------------------------- */
static double mesh_normalize(const orc_mesh* m, const orc_config* cfg, orc_result* res, float Eabsorb, float Etotal, int pair) {
    int i, j, k, datalen = res->datalen, maxgate = res->maxgate, srcnum = cfg->srcnum;
    double energydeposit = 0.f, energyelem, normalizor;
    double* W = res->field;

    if (cfg->issaveref && res->dref) {
        float nz = 1.f / Etotal;

        for (i = 0; i < maxgate; i++)
            for (j = 0; j < m->nf; j++) {
                res->dref[i * m->nf + j] *= nz;
            }
    }

    if (cfg->seed == ORC_SEED_FROM_FILE && (cfg->outputtype == ORC_JACOBIAN || cfg->outputtype == ORC_WL || cfg->outputtype == ORC_WP)) {
        float nz = 1.f / (DELTA_MUA * cfg->nphoton);

        if (cfg->outputtype == ORC_WL || cfg->outputtype == ORC_WP) {
            nz = 1.f / Etotal;
        }

        for (i = 0; i < maxgate; i++)
            for (j = 0; j < datalen; j++) {
                W[((size_t)i * datalen + j)*srcnum + pair] *= nz;
            }

        return nz;
    }

    if (cfg->outputtype == ORC_ENERGY) {
        normalizor = 1.f / Etotal;

        for (i = 0; i < maxgate; i++)
            for (j = 0; j < datalen; j++) {
                W[((size_t)i * datalen + j)*srcnum + pair] *= normalizor;
            }

        return normalizor;
    }

    if (cfg->method == ORC_BLBADOUEL_GRID) {
        normalizor = 1.0 / (Etotal * cfg->unitinmm * cfg->unitinmm * cfg->unitinmm);
    } else if (cfg->basisorder) {
        for (i = 0; i < maxgate; i++)
            for (j = 0; j < datalen; j++)
                if (m->nvol[j] > 0.f) {
                    W[((size_t)i * datalen + j)*srcnum + pair] /= m->nvol[j];
                }

        for (i = 0; i < m->ne; i++) {
            const int* ee = m->elem + 4 * (size_t)i;
            energyelem = 0.f;

            for (j = 0; j < maxgate; j++)
                for (k = 0; k < 4; k++) {
                    float re_val = W[((size_t)j * m->nn + ee[k] - 1) * srcnum + pair];
                    energyelem += re_val;
                }

            energydeposit += energyelem * m->evol[i] * medium_of(m, m->type[i])[0];
        }

        normalizor = Eabsorb / (Etotal * energydeposit * 0.25f);
    } else {
        for (i = 0; i < datalen; i++)
            for (j = 0; j < maxgate; j++) {
                energydeposit += W[((size_t)j * datalen + i) * srcnum + pair];
            }

        for (i = 0; i < datalen; i++) {
            energyelem = m->evol[i] * medium_of(m, m->type[i])[0];

            for (j = 0; j < maxgate; j++) {
                W[((size_t)j * datalen + i) * srcnum + pair] /= energyelem;
            }
        }

        normalizor = Eabsorb / (Etotal * energydeposit);
    }

    if (cfg->outputtype == ORC_FLUX) {
        normalizor /= cfg->tstep;
    }

    for (i = 0; i < maxgate; i++)
        for (j = 0; j < datalen; j++) {
            W[((size_t)i * datalen + j)*srcnum + pair] *= normalizor;
        }

    return normalizor;
}

int orc_maxgate(const orc_config* cfg) {  /* src/mmc_utils.c:3527-3528 */
    return (int)((cfg->tend - cfg->tstart) / cfg->tstep + 0.5);
}
int orc_datalen(const orc_mesh* m, const orc_config* cfg) {
    if (cfg->method == ORC_BLBADOUEL_GRID) {
        int dim[3];
        unsigned int crop[3];
        orc_mesh tmp = *m;
        orc_mesh_dualgrid(&tmp, cfg->steps, dim, crop);
        return (int)crop[2];
    }

    return cfg->basisorder ? m->nn : m->ne;
}
int orc_reclen(const orc_mesh* m, const orc_config* cfg) { /* src/mmc_host.c:248 */
    return (2 + ((cfg->ismomentum) > 0)) * m->prop + (cfg->issaveexit > 0) * 6 + 2;
}

/* ------------------------------------------------------------------------------------------
 * driver -- src/mmc_host.c:178-379 (mmc_run_mp) with tracer_prep (src/mmc_mesh.c:1241-1483)
 * ---------------------------------------------------------------------------------------- */
int orc_run(orc_mesh* m, orc_config* cfg, orc_result* res) {
    int nthread = cfg->nthread > 0 ? cfg->nthread : 1, j, failed = 0;
    uint32_t* seeds;
    unsigned int trajcount = 0;
    int dim[3] = {0, 0, 0};
    unsigned int crop[3] = {0, 0, 0};
    float bary0[4] = {0.f, 0.f, 0.f, 0.f};
    g_err[0] = 0;

    if (cfg->srcnum < 1 || cfg->srcnum > 16) {
        ORC_FAIL("srcnum must be 1..16");
    }

    if (cfg->method == ORC_BLBADOUEL_GRID) {
        cfg->basisorder = 0;    /* src/mmc_utils.c:3554-3556 */
    }

    if (cfg->issavedet && cfg->detnum == 0 && m->isextdet == 0) {
        cfg->issavedet = 0;    /* src/mmc_utils.c:3729-3736 */
    }

    if (!cfg->issavedet) {
        cfg->ismomentum = 0;
        cfg->issaveexit = 0;
    }

    if (cfg->tstep > cfg->tend - cfg->tstart) {
        cfg->tstep = cfg->tend - cfg->tstart;
    }

    res->maxgate = orc_maxgate(cfg);
    cfg->tend = cfg->tstart + cfg->tstep * res->maxgate;

    if (cfg->e0 == 0 && m->e0_from_src) {
        cfg->e0 = m->e0_from_src;    /* src/mmc_mesh.c:399-401 */
    }

    if (!(cfg->method == ORC_PLUCKER ? (m->pd && m->n) : (cfg->method == ORC_HAVEL ? m->m != NULL : m->n != NULL))) {
        orc_mesh_build_tracer(m, cfg->method);
    }

    if (cfg->srctype == stPencil || cfg->srctype == stIsotropic || cfg->srctype == stCone || cfg->srctype == stArcSin) {
        if (cfg->e0 <= 0 || mesh_barycentric(m, cfg->e0, bary0, cfg->srcpos)) {
            cfg->e0 = orc_mesh_initelem(m, cfg->srcpos, bary0);

            if (cfg->e0 == 0) {
                ORC_FAIL("initial element does not enclose the source!");
            }
        }
    }

    tracer_prep_mesh(m, cfg);
    res->e0 = cfg->e0;

    if (cfg->method == ORC_BLBADOUEL_GRID) {
        orc_mesh_dualgrid(m, cfg->steps, dim, crop);
        res->datalen = (int)crop[2];
    } else {
        res->datalen = cfg->basisorder ? m->nn : m->ne;
    }

    res->reclen = orc_reclen(m, cfg);

    if (res->reclen - 1 > 64) {
        ORC_FAIL("too many media for the oracle's partial-path buffer");
    }

    seeds = (uint32_t*)malloc(sizeof(uint32_t) * 4 * nthread);
    orc_host_seeds(cfg->seed, 4 * nthread, seeds);
    memset(res->launchweight, 0, sizeof(res->launchweight));
    memset(res->absorbweight, 0, sizeof(res->absorbweight));
    memset(res->escweight, 0, sizeof(res->escweight));
    res->raytet = 0;
    res->detectedcount = 0;
    res->trajcount = 0;

    #pragma omp parallel num_threads(nthread)
    {
        octx c;
        uint64_t ran[2];
        int tid = 0;
        int64_t id;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        memset(&c, 0, sizeof(c));
        c.mesh = m;
        c.cfg = cfg;
        c.weight = res->field;
        c.dref = res->dref;
        c.rtstep = 1.f / cfg->tstep;
        c.reclen = res->reclen;
        c.maxgate = res->maxgate;
        c.crop0[0] = crop[0];
        c.crop0[1] = crop[1];
        c.crop0[2] = crop[2];
        c.detbuf = res->detected;
        c.seedbuf = res->detseed;
        c.detcap = cfg->maxdetphoton;
        c.detcount = &res->detectedcount;
        memcpy(c.bary0, bary0, sizeof(bary0));
        c.traj = res->traj;
        c.trajcount = &trajcount;
        orc_rng_seed(seeds + 4 * tid, ran);

        /* per-thread detected buffers are merged in thread order by the reference (src/mmc_host.c:308-335);
         * with nthread==1 the order is the photon order, which is what the pin test uses */
        #pragma omp for schedule(static)
        for (id = 0; id < (int64_t)cfg->nphoton; id++) {
            if (failed) {
                continue;
            }

            if (cfg->seed == ORC_SEED_FROM_FILE) {
                uint64_t rs[2] = {cfg->photonseed[2 * id], cfg->photonseed[2 * id + 1]};

                if (onephoton(id, &c, rs)) {
                    failed = 1;
                }
            } else if (onephoton(id, &c, ran)) {
                failed = 1;
            }
        }

        #pragma omp critical
        {
            for (j = 0; j < cfg->srcnum; j++) {
                res->launchweight[j] += c.launchweight[j];
                res->absorbweight[j] += c.absorbweight[j];
                res->escweight[j] += c.escweight[j];
            }

            res->raytet += c.raytet;
        }
    }

    free(seeds);

    if (failed) {
        return -1;
    }

    res->trajcount = trajcount < cfg->maxjumpdebug ? trajcount : cfg->maxjumpdebug;

    if (cfg->isnormalized) {
        double sum = 0;

        for (j = 0; j < cfg->srcnum; j++) {
            float Eabs = cfg->gpu_semantics ? (float)(res->launchweight[j] - res->escweight[j]) : (float)res->absorbweight[j];
            sum += mesh_normalize(m, cfg, res, Eabs, (float)res->launchweight[j], j);
        }

        res->normalizer = sum / cfg->srcnum;
    }

    return 0;
}
