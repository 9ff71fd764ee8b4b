// Micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not hold:
//   gather : dependent random gathers of one tetrahedron record per step (64/96/128 B, LDG.128 vs LDG.256)
//   red    : random fire-and-forget atomics (f32 / f64) into an ne*gates-sized volume, plus a hot-spot variant
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
// Prints one JSON object per line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\":\"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

// Each thread walks a chain: idx -> record -> next idx (stored in the record) ; BYTES per record read with W-bit loads
template <int BYTES, int WIDTH>
__global__ void gather_chain(const char* __restrict__ tab, int stride, int nrec, int steps, uint32_t* out) {
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u % nrec;
    float acc = 0.f;
    for (int s = 0; s < steps; s++) {
        const char* p = tab + (size_t)idx * stride;
        uint32_t nxt;
        if (WIDTH == 256) {
            float v[8];
            #pragma unroll
            for (int b = 0; b < BYTES / 32; b++) {
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p + 32 * b));
                acc += v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7];
                if (b == 0) nxt = __float_as_uint(v[0]);
            }
        } else {
            #pragma unroll
            for (int b = 0; b < BYTES / 16; b++) {
                float4 v = __ldg((const float4*)(p + 16 * b));
                acc += v.y + v.z + v.w;
                if (b == 0) nxt = __float_as_uint(v.x); else acc += v.x;
            }
        }
        idx = (nxt + (acc > 1e30f ? 1u : 0u)) % nrec;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx + (uint32_t)acc;
}

template <typename T>
__global__ void red_random(T* vol, uint32_t n, int steps, uint32_t hotmask) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 747796405u + 12345u;
    for (int i = 0; i < steps; i++) {
        uint32_t r = lcg(s) >> 4;
        uint32_t idx = hotmask ? (r & hotmask) : (r % n);
        atomicAdd(vol + idx, (T)1);
    }
}

template <int BYTES, int WIDTH>
void run_gather(int nrec, int stride, int threads_per_sm, int steps) {
    std::vector<uint32_t> h((size_t)nrec * stride / 4);
    uint32_t s = 777;
    for (size_t i = 0; i < h.size(); i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % nrec; }
    char* d; uint32_t* o;
    CK(cudaMalloc(&d, (size_t)nrec * stride)); CK(cudaMemcpy(d, h.data(), (size_t)nrec * stride, cudaMemcpyHostToDevice));
    int block = 128, grid = 148 * threads_per_sm / block;
    CK(cudaMalloc(&o, (size_t)grid * block * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_chain<BYTES, WIDTH><<<grid, block>>>(d, stride, nrec, steps / 4, o);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        gather_chain<BYTES, WIDTH><<<grid, block>>>(d, stride, nrec, steps, o);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double nsteps = (double)grid * block * steps;
    printf("{\"bench\":\"gather\",\"bytes\":%d,\"width\":%d,\"stride\":%d,\"nrec\":%d,\"table_MB\":%.1f,\"threads_per_sm\":%d,\"ms\":%.3f,\"Gsteps_s\":%.3f,\"GBs\":%.1f}\n",
           BYTES, WIDTH, stride, nrec, (double)nrec * stride / 1e6, threads_per_sm, best, nsteps / best / 1e6, nsteps * BYTES / best / 1e6);
    fflush(stdout);
    cudaFree(d); cudaFree(o);
}

template <typename T>
void run_red(uint32_t n, int threads_per_sm, int steps, uint32_t hotmask, const char* name) {
    T* d; CK(cudaMalloc(&d, (size_t)n * sizeof(T))); CK(cudaMemset(d, 0, (size_t)n * sizeof(T)));
    int block = 128, grid = 148 * threads_per_sm / block;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    red_random<T><<<grid, block>>>(d, n, steps / 4, hotmask); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        red_random<T><<<grid, block>>>(d, n, steps, hotmask);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double nops = (double)grid * block * steps;
    printf("{\"bench\":\"red\",\"type\":\"%s\",\"n\":%u,\"vol_MB\":%.1f,\"hotmask\":%u,\"threads_per_sm\":%d,\"ms\":%.3f,\"Gatomics_s\":%.3f}\n",
           name, n, (double)n * sizeof(T) / 1e6, hotmask, threads_per_sm, best, nops / best / 1e6);
    fflush(stdout);
    cudaFree(d);
}

// usage: microbench                      the whole sweep
//        microbench quick NREC NVOL [DEV] the two ceilings bench.py quotes: random 96-byte record gathers (256-bit loads) from a table of
//                                         NREC records, random f64 reductions into a volume of NVOL accumulators (one line each)
int main(int argc, char** argv) {
    if (argc >= 4 && argv[1][0] == 'q') {
        if (argc >= 5) {
            CK(cudaSetDevice(atoi(argv[4])));
        }

        run_gather<96, 256>(atoi(argv[2]), 96, 1024, 2000);
        run_red<double>((uint32_t)strtoul(argv[3], NULL, 10), 1024, 2000, 0, "f64");
        return 0;
    }

    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\":\"%s\",\"sm\":%d,\"l2_MB\":%.1f,\"clock_MHz\":%d}\n", p.name, p.multiProcessorCount, p.l2CacheSize / 1e6, p.clockRate / 1000);
    const int steps = 2000;
    for (int tps : {512, 1024, 2048}) {
        // cube60-sized (135k) and colin27-sized (420k) meshes
        for (int nrec : {135000, 420000}) {
            run_gather<96, 256>(nrec, 96, tps, steps);
            run_gather<96, 128>(nrec, 96, tps, steps);
            run_gather<64, 256>(nrec, 64, tps, steps);
            run_gather<64, 128>(nrec, 64, tps, steps);
            run_gather<128, 256>(nrec, 128, tps, steps);
            run_gather<32, 256>(nrec, 32, tps, steps);
        }
    }
    for (int tps : {1024, 2048}) {
        run_red<float>(6750000u, tps, steps, 0, "f32");
        run_red<double>(6750000u, tps, steps, 0, "f64");
        run_red<float>(80000000u, tps, steps, 0, "f32");
        run_red<double>(80000000u, tps, steps, 0, "f64");
        run_red<float>(6750000u, tps, steps, 0xFF, "f32");      // 256 hot addresses
        run_red<double>(6750000u, tps, steps, 0xFF, "f64");
        run_red<float>(6750000u, tps, steps, 0xFFFF, "f32");    // 64k addresses (L2-resident hot region)
        run_red<double>(6750000u, tps, steps, 0xFFFF, "f64");
    }
    return 0;
}
