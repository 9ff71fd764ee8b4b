// Mesh pre-processing on the GPU (sm_100a) -- SURVEY.md section 8(f) rank 4.  The reference does these on the host:
//   * mesh_getfacenb (src/mmc_highorder.cpp:124-159): face-neighbour table by matching the sorted node triples of the 4*ne faces.
//     Here: a device hash join.  Every face inserts its 63-bit key (three 21-bit node ids) into an open-addressing table with
//     atomicCAS; the first face to arrive owns the slot, the second one is its neighbour.  O(ne) instead of a 4*ne-key sort
//     (80-160 ms on the host for a 260k-element head mesh, the dominant cost of a one-call run with few photons).
//   * tracer_build for the branch-less Badouel tracer (src/mmc_mesh.c:1572-1600) + the per-face flags the photon kernel wants:
//     one thread per element writes the 96-byte record (mmcb_types.h) and the centroid straight into device memory, so the
//     records never exist on the host.  The arithmetic uses the round-to-nearest intrinsics, no FMA contraction, no fast-math
//     approximations: the records are bit-identical to the host builder (mmcb_host.cu: build_records), which stays as the
//     checker (tests/test_prep_gpu.py) and as the path for the Havel/Plucker tables.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmcb_types.h"
#include "../../include/mmc_b200.h"

#define EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long face_key(const int* __restrict__ elem, int f) {
    // faces j <-> nodes FACELIST[j] = {0,1,2},{0,1,3},{0,2,3},{1,2,3} (src/mmc_highorder.cpp:50)
    const int4 e = *(const int4*)(elem + 4 * (size_t)(f >> 2));
    const int j = f & 3;
    unsigned int a = (j == 3) ? e.y : e.x, b = (j < 2) ? e.y : e.z, c = (j == 0) ? e.z : e.w;
    unsigned int t;

    if (a > b) {
        t = a;
        a = b;
        b = t;
    }

    if (b > c) {
        t = b;
        b = c;
        c = t;
    }

    if (a > b) {
        t = a;
        a = b;
        b = t;
    }

    return ((unsigned long long)a << 42) | ((unsigned long long)b << 21) | (unsigned long long)c;
}

__device__ __forceinline__ unsigned int key_hash(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}

__global__ void mmcb_facenb_insert_kernel(const int* __restrict__ elem, int nface, unsigned long long* __restrict__ keys, int2* __restrict__ vals,
        unsigned int mask) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nface; f += gridDim.x * blockDim.x) {
        const unsigned long long key = face_key(elem, f);
        unsigned int h = key_hash(key) & mask;

        while (true) {
            const unsigned long long prev = atomicCAS(keys + h, EMPTY_KEY, key);

            if (prev == EMPTY_KEY || prev == key) {
                if (atomicCAS(&vals[h].x, -1, f) != -1) {
                    vals[h].y = f;          // second face with this node triple: the neighbour
                }

                break;
            }

            h = (h + 1) & mask;
        }
    }
}

__global__ void mmcb_facenb_lookup_kernel(const int* __restrict__ elem, int nface, const unsigned long long* __restrict__ keys,
        const int2* __restrict__ vals, unsigned int mask, int* __restrict__ facenb) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nface; f += gridDim.x * blockDim.x) {
        const unsigned long long key = face_key(elem, f);
        unsigned int h = key_hash(key) & mask;

        while (keys[h] != key) {
            h = (h + 1) & mask;
        }

        const int2 v = vals[h];
        const int other = (v.x == f) ? v.y : v.x;
        facenb[f] = (other >= 0) ? (other >> 2) + 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// BLB records: plane normals (unit, outward for positively oriented elements) and offsets of the faces in tracer order
// out[j] = {0,3,1},{3,2,1},{0,2,3},{0,1,2} (src/mmc_mesh.c:59), neighbours permuted by faceorder = {1,3,2,0} (:84), medium label,
// per-face flags (reflect / to void / from void; mmcb_types.h) and the centroid (fixphoton)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sub_rn(float a, float b) {
    return __fsub_rn(a, b);
}
__device__ __forceinline__ float cross_c(float a1, float b2, float a2, float b1) {      // a1*b2 - a2*b1 without contraction
    return __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
}

__global__ void mmcb_build_records_kernel(const float* __restrict__ node, const int* __restrict__ elem, const int* __restrict__ facenb,
        const int* __restrict__ type, const float* __restrict__ med_n, int ne, float nout, int isreflect,
        mmcb_tetrec* __restrict__ rec, float4* __restrict__ cent) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += gridDim.x * blockDim.x) {
        const int4 e4 = *(const int4*)(elem + 4 * (size_t)i);
        const int4 nb4 = *(const int4*)(facenb + 4 * (size_t)i);
        const int ee[4] = {e4.x, e4.y, e4.z, e4.w}, fnb[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
        float q[4][3];
        #pragma unroll

        for (int k = 0; k < 4; k++) {
            q[k][0] = node[3 * (size_t)(ee[k] - 1)];
            q[k][1] = node[3 * (size_t)(ee[k] - 1) + 1];
            q[k][2] = node[3 * (size_t)(ee[k] - 1) + 2];
        }

        const int ty = type[i];
        const float n_here = med_n[ty];
        mmcb_tetrec r;
        r.flags = 0;
        r.pad[0] = r.pad[1] = 0;
        r.type = ty;
        const int OUTJ[4][3] = {{0, 3, 1}, {3, 2, 1}, {0, 2, 3}, {0, 1, 2}}, FO[4] = {1, 3, 2, 0};
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            const float* a = q[OUTJ[j][0]], *b = q[OUTJ[j][1]], *c = q[OUTJ[j][2]];
            const float ABx = sub_rn(b[0], a[0]), ABy = sub_rn(b[1], a[1]), ABz = sub_rn(b[2], a[2]);
            const float ACx = sub_rn(c[0], a[0]), ACy = sub_rn(c[1], a[1]), ACz = sub_rn(c[2], a[2]);
            float Nx = cross_c(ABy, ACz, ABz, ACy), Ny = cross_c(ABz, ACx, ABx, ACz), Nz = cross_c(ABx, ACy, ABy, ACx);
            const float R = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(Nx, Nx), __fmul_rn(Ny, Ny)), __fmul_rn(Nz, Nz))));
            Nx = __fmul_rn(Nx, R);
            Ny = __fmul_rn(Ny, R);
            Nz = __fmul_rn(Nz, R);
            r.nx[j] = Nx;
            r.ny[j] = Ny;
            r.nz[j] = Nz;
            r.d[j] = __fadd_rn(__fadd_rn(__fmul_rn(Nx, a[0]), __fmul_rn(Ny, a[1])), __fmul_rn(Nz, a[2]));
            const int nb = fnb[FO[j]];
            r.nb[j] = nb;
            bool refl;

            if (nb <= 0) {      // src/mmc_core.cl:1957-1961
                refl = !((n_here == nout && isreflect != MMCB_BC_MIRROR) || isreflect == MMCB_BC_ABSORB_EXTERIOR);
            } else {
                const int tn = type[nb - 1];
                refl = (med_n[tn] != n_here);

                if (ty != 0 && tn == 0) {
                    r.flags |= MMCB_F_TO_VOID(j);
                }

                if (ty == 0 && tn != 0) {
                    r.flags |= MMCB_F_FROM_VOID(j);
                }
            }

            if (refl) {
                r.flags |= MMCB_F_REFLECT(j);
            }
        }

        rec[i] = r;
        const float cx = __fadd_rn(__fadd_rn(__fadd_rn(q[0][0], q[1][0]), q[2][0]), q[3][0]);
        const float cy = __fadd_rn(__fadd_rn(__fadd_rn(q[0][1], q[1][1]), q[2][1]), q[3][1]);
        const float cz = __fadd_rn(__fadd_rn(__fadd_rn(q[0][2], q[1][2]), q[2][2]), q[3][2]);
        cent[i] = make_float4(__fmul_rn(cx, 0.25f), __fmul_rn(cy, 0.25f), __fmul_rn(cz, 0.25f), 0.f);
    }
}

// Companion record of the Havel / Plucker kernels for nodal output (mmcb_types.h): for tracer face j the node opposite to it is
// local node {2,0,1,3}[j] (facemap, src/mmc_raytrace.c:60); its height above the face is d_j - n_j.x with the plane of the record.
__global__ void mmcb_build_hpaux_kernel(const float* __restrict__ node, const int* __restrict__ elem, const mmcb_tetrec* __restrict__ rec, int ne,
                                        float* __restrict__ aux) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += gridDim.x * blockDim.x) {
        const int4 e4 = *(const int4*)(elem + 4 * (size_t)i);
        const int opp[4] = {e4.z, e4.x, e4.y, e4.w};
        const mmcb_tetrec r = rec[i];
        float4 invh;
        float* o = &invh.x;
        #pragma unroll

        for (int j = 0; j < 4; j++) {
            const float* x = node + 3 * (size_t)(opp[j] - 1);
            const float h = r.d[j] - (r.nx[j] * x[0] + r.ny[j] * x[1] + r.nz[j] * x[2]);
            o[j] = (h != 0.f) ? __fdiv_rn(1.f, h) : 0.f;
        }

        float4* out = (float4*)(aux + MMCB_HPAUX_FLOATS * (size_t)i);
        out[0] = invh;
        out[1] = make_float4(__int_as_float(opp[0]), __int_as_float(opp[1]), __int_as_float(opp[2]), __int_as_float(opp[3]));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
static inline int grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b < 148 * 8 ? (b ? b : 1) : 148 * 8);
}

// facenb for `ne` elements already on the device (d_elem); d_facenb: ne*4 ints out.  Scratch is allocated from the stream's pool.
extern "C" int mmcb_k_facenb(const int* d_elem, int ne, int* d_facenb, cudaStream_t st) {
    const int nface = 4 * ne;
    unsigned int cap = 1024;

    while (cap < 2u * (unsigned int)nface) {
        cap <<= 1;
    }

    unsigned long long* keys = NULL;
    int2* vals = NULL;
    cudaError_t e = cudaMallocAsync(&keys, sizeof(unsigned long long) * cap, st);

    if (e == cudaSuccess) {
        e = cudaMallocAsync(&vals, sizeof(int2) * cap, st);
    }

    if (e == cudaSuccess) {
        e = cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * cap, st);
    }

    if (e == cudaSuccess) {
        e = cudaMemsetAsync(vals, 0xFF, sizeof(int2) * cap, st);        // -1, -1
    }

    if (e == cudaSuccess) {
        mmcb_facenb_insert_kernel<<<grid_for(nface), 256, 0, st>>>(d_elem, nface, keys, vals, cap - 1);
        mmcb_facenb_lookup_kernel<<<grid_for(nface), 256, 0, st>>>(d_elem, nface, keys, vals, cap - 1, d_facenb);
        e = cudaGetLastError();
    }

    if (keys) {
        cudaFreeAsync(keys, st);
    }

    if (vals) {
        cudaFreeAsync(vals, st);
    }

    return (int)e;
}

extern "C" int mmcb_k_build_records(const float* d_node, const int* d_elem, const int* d_facenb, const int* d_type, const float* d_med_n, int ne,
                                    float nout, int isreflect, mmcb_tetrec* d_rec, float4* d_cent, cudaStream_t st) {
    mmcb_build_records_kernel<<<grid_for((size_t)ne), 256, 0, st>>>(d_node, d_elem, d_facenb, d_type, d_med_n, ne, nout, isreflect, d_rec, d_cent);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_build_hpaux(const float* d_node, const int* d_elem, const mmcb_tetrec* d_rec, int ne, float* d_aux, cudaStream_t st) {
    mmcb_build_hpaux_kernel<<<grid_for((size_t)ne), 256, 0, st>>>(d_node, d_elem, d_rec, ne, d_aux);
    return (int)cudaGetLastError();
}
