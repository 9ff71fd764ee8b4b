// Post-processing kernels of mmc_b200 (sm_100a): mesh_normalize on the device (end of file) and the adjoint-Jacobian kernels, i.e. what the reference runs after a multi-slot (sources + detectors-as-sources)
// forward simulation to turn the slots' fluence volumes into Jacobians (src/mmc_core.cl:2218-2649, driven by
// src/mmc_cu_host.cu:1063-1395).  All of them are streaming, HBM/L2-bound kernels.
//
// Re-designed rather than ported: the reference re-sums the time gates of a detector slot for every (source, detector) pair and
// every finite-difference neighbour inside one thread per voxel (maxgate * Ns * Nd * 7 strided reads per voxel for J_D).  Here
//   1. one pass reduces the gates once per slot (coalesced, every field element is read exactly once)       -> cw[slot][i]
//   2. the pair kernels read cw only (nslots values per voxel / 4 nodes per element) and write Ns*Nd outputs
// and the mesh kernels compute the shape-function gradient products from the node coordinates in registers instead of reading the
// ne*10 deldotdel table of mesh_deldotdel (src/mmc_mesh.c:990-1049).
#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------------------------------------------------
// 1. CW sum over the time gates, mmc_cw_sum (src/mmc_core.cl:2227-2242): cw[slot*N + i] = sum_t field[i + (t + slot*maxgate)*N]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mmcb_adj_cw_kernel(const float* __restrict__ field, float* __restrict__ cw, size_t N, int maxgate, int nslots) {
    const size_t total = N * (size_t)nslots;

    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t slot = k / N, i = k - slot * N;
        const float* f = field + i + slot * (size_t)maxgate * N;
        float sum = 0.f;

        for (int t = 0; t < maxgate; t++) {
            sum += f[(size_t)t * N];
        }

        cw[k] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2a. grid mode, J_mua[v, s, d] = scale * phi_s[v] * phi_d[v]  (complex product for RF), mmc_adjoint_kernel (:2298-2333)
//     out[v + (s*Nd+d)*N] real, + Ns*Nd*N imaginary
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mmcb_adj_mua_kernel(const float* __restrict__ cw_re, const float* __restrict__ cw_im, float* __restrict__ out,
                                    size_t N, int Ns, int Nd, float scale) {
    const size_t adjlen = N * (size_t)Ns * Nd;

    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < N; v += (size_t)gridDim.x * blockDim.x) {
        for (int s = 0; s < Ns; s++) {
            const float sr = cw_re[s * N + v], si = cw_im ? cw_im[s * N + v] : 0.f;

            for (int d = 0; d < Nd; d++) {
                const float dr = cw_re[(size_t)(Ns + d) * N + v];
                const size_t o = v + (size_t)(s * Nd + d) * N;
                float re = sr * dr;

                if (cw_im) {
                    const float di = cw_im[(size_t)(Ns + d) * N + v];
                    re -= si * di;
                    out[o + adjlen] = scale * (sr * di + si * dr);
                }

                out[o] = scale * re;
            }
        }
    }
}

// 2nd-order finite difference along one axis, mmc_fd_grad (:2247-2285), on the CW volume of one slot
__device__ __forceinline__ float fd_grad(const float* __restrict__ f, size_t v, unsigned int i, unsigned int n, size_t stride) {
    if (n <= 1) {
        return 0.f;
    }

    const float f0 = f[v];

    if (i == 0) {
        const float p1 = f[v + stride];
        return (n == 2) ? (p1 - f0) : (-3.f * f0 + 4.f * p1 - f[v + 2 * stride]) * 0.5f;
    }

    if (i == n - 1) {
        const float m1 = f[v - stride];
        return (n == 2) ? (f0 - m1) : (f[v - 2 * stride] - 4.f * m1 + 3.f * f0) * 0.5f;
    }

    return (f[v + stride] - f[v - stride]) * 0.5f;
}

// 2b. grid mode, J_D[v, s, d] = scale * grad phi_s . grad phi_d, mmc_adjoint_dcoeff_kernel (:2343-2401)
__global__ void mmcb_adj_dcoeff_kernel(const float* __restrict__ cw_re, const float* __restrict__ cw_im, float* __restrict__ out,
                                       size_t N, int Ns, int Nd, unsigned int Nx, unsigned int Ny, float scale) {
    const size_t adjlen = N * (size_t)Ns * Nd;
    const size_t Nxy = (size_t)Nx * Ny;
    const unsigned int Nz = (unsigned int)(N / Nxy);

    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < N; v += (size_t)gridDim.x * blockDim.x) {
        const unsigned int ix = (unsigned int)(v % Nx), iy = (unsigned int)((v / Nx) % Ny), iz = (unsigned int)(v / Nxy);

        for (int s = 0; s < Ns; s++) {
            const float* fs = cw_re + s * N;
            const float sxr = fd_grad(fs, v, ix, Nx, 1), syr = fd_grad(fs, v, iy, Ny, Nx), szr = fd_grad(fs, v, iz, Nz, Nxy);
            float sxi = 0.f, syi = 0.f, szi = 0.f;

            if (cw_im) {
                const float* gs = cw_im + s * N;
                sxi = fd_grad(gs, v, ix, Nx, 1);
                syi = fd_grad(gs, v, iy, Ny, Nx);
                szi = fd_grad(gs, v, iz, Nz, Nxy);
            }

            for (int d = 0; d < Nd; d++) {
                const float* fd = cw_re + (size_t)(Ns + d) * N;
                const float dxr = fd_grad(fd, v, ix, Nx, 1), dyr = fd_grad(fd, v, iy, Ny, Nx), dzr = fd_grad(fd, v, iz, Nz, Nxy);
                const size_t o = v + (size_t)(s * Nd + d) * N;
                float re = sxr * dxr + syr * dyr + szr * dzr;

                if (cw_im) {
                    const float* gd = cw_im + (size_t)(Ns + d) * N;
                    const float dxi = fd_grad(gd, v, ix, Nx, 1), dyi = fd_grad(gd, v, iy, Ny, Nx), dzi = fd_grad(gd, v, iz, Nz, Nxy);
                    re -= sxi * dxi + syi * dyi + szi * dzi;
                    out[o + adjlen] = scale * (sxr * dxi + syr * dyi + szr * dzi + sxi * dxr + syi * dyr + szi * dzr);
                }

                out[o] = scale * re;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3a. mesh mode, full FEM form (rb_femjacobian), one thread per element, 0.25-weighted scatter to the 4 nodes
//     mmc_adjoint_mesh_full_kernel (:2438-2587):
//       J_mua(t) = -0.1 Ve [ sum_i ps_i pd_i + 0.5 sum_{i<j} (ps_i pd_j + ps_j pd_i) ]
//       J_D(t)   = -[ sum_i G_ii ps_i pd_i + sum_{i<j} G_ij (ps_i pd_j + ps_j pd_i) ],  G_ij = grad(N_i).grad(N_j) Ve
//     G is computed here from the node coordinates: grad N_i = n_i / (n_i . (p_i - a_i)), n_i the normal of the opposite face.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mmcb_adj_mesh_full_kernel(const float* __restrict__ cw_re, const float* __restrict__ cw_im, const int* __restrict__ elem,
        const float* __restrict__ node, const float* __restrict__ evol, float* __restrict__ jmua, float* __restrict__ jd,
        int ne, int nn, int Ns, int Nd) {
    const size_t adjlen = (size_t)nn * Ns * Nd;

    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ne; t += gridDim.x * blockDim.x) {
        int ee[4];
        float q[4][3];
        #pragma unroll

        for (int k = 0; k < 4; k++) {
            ee[k] = elem[4 * (size_t)t + k] - 1;
            q[k][0] = node[3 * (size_t)ee[k]];
            q[k][1] = node[3 * (size_t)ee[k] + 1];
            q[k][2] = node[3 * (size_t)ee[k] + 2];
        }

        const float Ve = evol[t];
        float g[4][3];

        if (jd) {
            #pragma unroll

            for (int i = 0; i < 4; i++) {
                const int a = (i + 1) & 3, b = (i + 2) & 3, c = (i + 3) & 3;
                const float ux = q[b][0] - q[a][0], uy = q[b][1] - q[a][1], uz = q[b][2] - q[a][2];
                const float vx = q[c][0] - q[a][0], vy = q[c][1] - q[a][1], vz = q[c][2] - q[a][2];
                const float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
                const float h = nx * (q[i][0] - q[a][0]) + ny * (q[i][1] - q[a][1]) + nz * (q[i][2] - q[a][2]);
                const float r = 1.f / h;
                g[i][0] = nx * r;
                g[i][1] = ny * r;
                g[i][2] = nz * r;
            }
        }

        for (int s = 0; s < Ns; s++) {
            float psr[4], psi[4];
            #pragma unroll

            for (int k = 0; k < 4; k++) {
                psr[k] = cw_re[(size_t)s * nn + ee[k]];
                psi[k] = cw_im ? cw_im[(size_t)s * nn + ee[k]] : 0.f;
            }

            for (int d = 0; d < Nd; d++) {
                float pdr[4], pdi[4];
                #pragma unroll

                for (int k = 0; k < 4; k++) {
                    pdr[k] = cw_re[(size_t)(Ns + d) * nn + ee[k]];
                    pdi[k] = cw_im ? cw_im[(size_t)(Ns + d) * nn + ee[k]] : 0.f;
                }

                float mr = 0.f, mi = 0.f, dr = 0.f, di = 0.f;
                #pragma unroll

                for (int i = 0; i < 4; i++) {
                    #pragma unroll

                    for (int j = i; j < 4; j++) {
                        float pre, pim;

                        if (i == j) {
                            pre = psr[i] * pdr[i] - psi[i] * pdi[i];
                            pim = psr[i] * pdi[i] + psi[i] * pdr[i];
                            mr += pre;
                            mi += pim;
                        } else {
                            pre = psr[i] * pdr[j] + psr[j] * pdr[i] - psi[i] * pdi[j] - psi[j] * pdi[i];
                            pim = psr[i] * pdi[j] + psr[j] * pdi[i] + psi[i] * pdr[j] + psi[j] * pdr[i];
                            mr += 0.5f * pre;
                            mi += 0.5f * pim;
                        }

                        if (jd) {
                            const float w = (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]) * Ve;
                            dr += w * pre;
                            di += w * pim;
                        }
                    }
                }

                mr *= -0.1f * Ve * 0.25f;
                mi *= -0.1f * Ve * 0.25f;
                dr *= -0.25f;
                di *= -0.25f;
                const size_t sd = (size_t)(s * Nd + d) * nn;
                #pragma unroll

                for (int k = 0; k < 4; k++) {
                    if (jmua) {
                        atomicAdd(jmua + sd + ee[k], mr);

                        if (cw_im) {
                            atomicAdd(jmua + sd + ee[k] + adjlen, mi);
                        }
                    }

                    if (jd) {
                        atomicAdd(jd + sd + ee[k], dr);

                        if (cw_im) {
                            atomicAdd(jd + sd + ee[k] + adjlen, di);
                        }
                    }
                }
            }
        }
    }
}

// 3b. mesh mode, nodal approximation J_mua[n] = -nvol[n] phi_s[n] phi_d[n], mmc_adjoint_mesh_nodal_kernel (:2609-2649)
__global__ void mmcb_adj_mesh_nodal_kernel(const float* __restrict__ cw_re, const float* __restrict__ cw_im, const float* __restrict__ nvol,
        float* __restrict__ jmua, int nn, int Ns, int Nd) {
    const size_t adjlen = (size_t)nn * Ns * Nd;

    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += gridDim.x * blockDim.x) {
        const float vol = nvol[n];

        for (int s = 0; s < Ns; s++) {
            const float sr = cw_re[(size_t)s * nn + n], si = cw_im ? cw_im[(size_t)s * nn + n] : 0.f;

            for (int d = 0; d < Nd; d++) {
                const float dr = cw_re[(size_t)(Ns + d) * nn + n];
                const size_t o = (size_t)n + (size_t)(s * Nd + d) * nn;

                if (cw_im) {
                    const float di = cw_im[(size_t)(Ns + d) * nn + n];
                    jmua[o] = -vol * (sr * dr - si * di);
                    jmua[o + adjlen] = -vol * (sr * di + si * dr);
                } else {
                    jmua[o] = -vol * sr * dr;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers (grids: multiples of the 148 SMs, grid-stride loops)
// ---------------------------------------------------------------------------------------------------------------------
static inline int grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b < 148 * 8 ? (b ? b : 1) : 148 * 8);
}

extern "C" int mmcb_k_adj_cw(const float* field, float* cw, size_t N, int maxgate, int nslots, cudaStream_t st) {
    mmcb_adj_cw_kernel<<<grid_for(N * nslots), 256, 0, st>>>(field, cw, N, maxgate, nslots);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_adj_mua(const float* cw_re, const float* cw_im, float* out, size_t N, int Ns, int Nd, float scale, cudaStream_t st) {
    mmcb_adj_mua_kernel<<<grid_for(N), 256, 0, st>>>(cw_re, cw_im, out, N, Ns, Nd, scale);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_adj_dcoeff(const float* cw_re, const float* cw_im, float* out, size_t N, int Ns, int Nd, unsigned int Nx, unsigned int Ny,
                                 float scale, cudaStream_t st) {
    mmcb_adj_dcoeff_kernel<<<grid_for(N), 256, 0, st>>>(cw_re, cw_im, out, N, Ns, Nd, Nx, Ny, scale);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_adj_mesh_full(const float* cw_re, const float* cw_im, const int* elem, const float* node, const float* evol, float* jmua,
                                    float* jd, int ne, int nn, int Ns, int Nd, cudaStream_t st) {
    mmcb_adj_mesh_full_kernel<<<grid_for((size_t)ne), 256, 0, st>>>(cw_re, cw_im, elem, node, evol, jmua, jd, ne, nn, Ns, Nd);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_adj_mesh_nodal(const float* cw_re, const float* cw_im, const float* nvol, float* jmua, int nn, int Ns, int Nd,
                                     cudaStream_t st) {
    mmcb_adj_mesh_nodal_kernel<<<grid_for((size_t)nn), 256, 0, st>>>(cw_re, cw_im, nvol, jmua, nn, Ns, Nd);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// mesh_normalize (src/mmc_mesh.c:2154-2279) on the device: the volume is reduced and scaled where it lives, and crosses PCIe once
// in its final form (the reference downloads raw floats, then makes several host passes over the double volume).
// Layout: W[(gate*datalen + i)*srcnum + pair], like mesh->weight.
// ---------------------------------------------------------------------------------------------------------------------
struct mmcb_normfac {
    double f[16];
};

__device__ __forceinline__ void block_add(double v, double* dst) {
    #pragma unroll

    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    }

    __shared__ double part[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;

    if (l == 0) {
        part[w] = v;
    }

    __syncthreads();

    if (w == 0) {
        v = (l < (blockDim.x >> 5)) ? part[l] : 0.0;
        #pragma unroll

        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        }

        if (l == 0) {
            atomicAdd(dst, v);
        }
    }
}

// basisorder 0: energydeposit[pair] = sum of all entries of the pair (:2247-2250)
__global__ void mmcb_norm_sum_kernel(const double* __restrict__ W, size_t nentry, int srcnum, double* __restrict__ dep) {
    const int p = blockIdx.y;
    double s = 0.0;

    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nentry; e += (size_t)gridDim.x * blockDim.x) {
        s += W[e * srcnum + p];
    }

    block_add(s, dep + p);
}

// basisorder 1, step 1: W /= nvol[node] where nvol > 0 (:2213-2222)
__global__ void mmcb_norm_nvol_kernel(double* __restrict__ W, size_t n, int nn, int srcnum, const float* __restrict__ nvol) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = nvol[(i / srcnum) % nn];

        if (v > 0.f) {
            W[i] /= v;
        }
    }
}

// basisorder 1, step 2: energydeposit[pair] = sum_e (sum_gates sum_4nodes (float)W) * evol_e * mua_e (:2224-2245); with an imaginary
// volume (RF, one pattern) the summand is |phi| = sqrtf(re^2 + im^2) (:2232-2237)
__global__ void mmcb_norm_elemdep_kernel(const double* __restrict__ W, const double* __restrict__ Wim, const int* __restrict__ elem,
        const float* __restrict__ evol, const float* __restrict__ emua, int ne, int nn, int maxgate, int srcnum, double* __restrict__ dep) {
    const int p = blockIdx.y;
    double s = 0.0;

    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
        const int4 ee = *(const int4*)(elem + 4 * (size_t)e);
        const int id[4] = {ee.x, ee.y, ee.z, ee.w};
        double energyelem = 0.0;

        for (int g = 0; g < maxgate; g++) {
            const size_t base = (size_t)g * nn;
            #pragma unroll

            for (int k = 0; k < 4; k++) {
                const size_t at = (base + id[k] - 1) * srcnum + p;
                const float re = (float)W[at];

                if (Wim) {
                    const float im = (float)Wim[at];
                    energyelem += sqrtf(re * re + im * im);
                } else {
                    energyelem += re;
                }
            }
        }

        s += energyelem * evol[e] * emua[e];
    }

    block_add(s, dep + p);
}

__global__ void mmcb_double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = (float)in[i];
    }
}

// final pass: out = (in / (evol*mua)) * fac[pair]   (division only for basisorder 0, :2252-2258), out may alias in
__global__ void mmcb_norm_scale_kernel(const double* __restrict__ in, double* __restrict__ out, size_t n, int datalen, int srcnum,
                                       const float* __restrict__ evol, const float* __restrict__ emua, const mmcb_normfac fac) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v = in[i];

        if (evol) {
            const size_t e = (i / srcnum) % datalen;
            v /= (double)__fmul_rn(evol[e], emua[e]);
        }

        out[i] = v * fac.f[i % srcnum];
    }
}

extern "C" int mmcb_k_norm_sum(const double* W, size_t nentry, int srcnum, double* dep, cudaStream_t st) {
    dim3 g(grid_for(nentry), srcnum);
    mmcb_norm_sum_kernel<<<g, 256, 0, st>>>(W, nentry, srcnum, dep);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_norm_nvol(double* W, size_t n, int nn, int srcnum, const float* nvol, cudaStream_t st) {
    mmcb_norm_nvol_kernel<<<grid_for(n), 256, 0, st>>>(W, n, nn, srcnum, nvol);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_norm_elemdep(const double* W, const double* Wim, const int* elem, const float* evol, const float* emua, int ne, int nn, int maxgate,
                                   int srcnum, double* dep, cudaStream_t st) {
    dim3 g(grid_for((size_t)ne), srcnum);
    mmcb_norm_elemdep_kernel<<<g, 256, 0, st>>>(W, Wim, elem, evol, emua, ne, nn, maxgate, srcnum, dep);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_double_to_float(const double* in, float* out, size_t n, cudaStream_t st) {
    mmcb_double_to_float_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n);
    return (int)cudaGetLastError();
}

extern "C" int mmcb_k_norm_scale(const double* in, double* out, size_t n, int datalen, int srcnum, const float* evol, const float* emua,
                                 const double* fac16, cudaStream_t st) {
    mmcb_normfac f;

    for (int i = 0; i < 16; i++) {
        f.f[i] = fac16[i];
    }

    mmcb_norm_scale_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n, datalen, srcnum, evol, emua, f);
    return (int)cudaGetLastError();
}
