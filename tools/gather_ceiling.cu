// Empirical random-sector ceiling for the occ-table access pattern (SURVEY.md §8d:
// "the harness must additionally measure an empirical ceiling with an independent
// random gather over a table of the index's size").
//
// Two access patterns over a table of T bytes:
//   indep : address_i = hash(thread, i)            -- no dependence between loads
//   chain : address_{i+1} = hash(value_i, thread)  -- one dependent chain per "query",
//           C chains interleaved per thread (what backward search looks like)
// Granule G = 16/32/64/128 bytes, aligned.  Prints G granules/s and GB/s.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_ceiling gather_ceiling.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template <int G>
__device__ __forceinline__ uint32_t load_granule(const char* p) {
    if constexpr (G == 16) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        return v.x ^ v.y ^ v.z ^ v.w;
    } else if constexpr (G == 32) {
        uint32_t a, b, c, d, e, f, g, h;
        asm volatile("ld.global.nc.L2::evict_normal.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
        return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    } else {
        uint32_t acc = 0;
#pragma unroll
        for (int o = 0; o < G; o += 32) {
            uint32_t a, b, c, d, e, f, g, h;
            asm volatile("ld.global.nc.L2::evict_normal.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + o));
            acc ^= a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
        }
        return acc;
    }
}

template <int G, int C>
__global__ void __launch_bounds__(256) chain_kernel(const char* tab, uint64_t ngran, int iters, uint32_t* out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st[C];
#pragma unroll
    for (int c = 0; c < C; ++c) st[c] = mix64(tid * C + c);
    for (int i = 0; i < iters; ++i) {
        uint32_t v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = load_granule<G>(tab + (st[c] % ngran) * G);
#pragma unroll
        for (int c = 0; c < C; ++c) st[c] = mix64(st[c] ^ v[c]);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) acc ^= (uint32_t)st[c];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int G>
__global__ void __launch_bounds__(256) indep_kernel(const char* tab, uint64_t ngran, int iters, uint32_t* out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        uint64_t a = mix64(tid * 1315423911ull + i) % ngran;
        acc ^= load_granule<G>(tab + a * G);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fill_kernel(uint4* p, uint64_t n16) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) {
        uint64_t a = mix64(i);
        p[i] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)i, 7u);
    }
}

template <typename F>
static float time_ms(F&& f, int reps = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 1.5;
    uint64_t tbytes = (uint64_t)(gb * 1e9) / 4096 * 4096;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs, table %.2f GB\n", prop.name, prop.multiProcessorCount, tbytes / 1e9);
    char* tab; CK(cudaMalloc(&tab, tbytes));
    uint32_t* out; CK(cudaMalloc(&out, 4));
    fill_kernel<<<prop.multiProcessorCount * 8, 256>>>((uint4*)tab, tbytes / 16);
    CK(cudaDeviceSynchronize());
    int sms = prop.multiProcessorCount;
    for (int gran_limit : {32, 64, 128}) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran_limit);
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("# L2 fetch granularity request %d -> %zu (%s)\n", gran_limit, got, cudaGetErrorString(e));
        const int iters = 256;
        for (int occ : {4, 8}) {
            int blocks = sms * occ;
            uint64_t threads = (uint64_t)blocks * 256;
#define RUN_INDEP(G) { float ms = time_ms([&] { indep_kernel<G><<<blocks, 256>>>(tab, tbytes / G, iters, out); }); \
            double n = (double)threads * iters; \
            printf("indep G=%3d blocks/SM=%d : %8.3f ms  %7.2f Ggran/s  %8.1f GB/s\n", G, occ, ms, n / ms / 1e6, n * G / ms / 1e6); }
            RUN_INDEP(16) RUN_INDEP(32) RUN_INDEP(64) RUN_INDEP(128)
#define RUN_CHAIN(G, C) { float ms = time_ms([&] { chain_kernel<G, C><<<blocks, 256>>>(tab, tbytes / G, iters, out); }); \
            double n = (double)threads * iters * C; \
            printf("chain G=%3d C=%d blocks/SM=%d : %8.3f ms  %7.2f Ggran/s  %8.1f GB/s\n", G, C, occ, ms, n / ms / 1e6, n * G / ms / 1e6); }
            RUN_CHAIN(32, 1) RUN_CHAIN(32, 2) RUN_CHAIN(32, 4) RUN_CHAIN(32, 8)
            RUN_CHAIN(64, 1) RUN_CHAIN(64, 2) RUN_CHAIN(64, 4)
            RUN_CHAIN(128, 1) RUN_CHAIN(128, 2) RUN_CHAIN(128, 4)
            RUN_CHAIN(16, 2) RUN_CHAIN(16, 4)
        }
    }
    return 0;
}
