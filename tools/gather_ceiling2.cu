// Random-granule HBM ceiling, cooperative variant: a group of LANES lanes fetches one aligned G-byte granule
// with ONE load instruction (each lane 16 or 32 bytes), so that a 64/128/256-byte granule reaches the memory
// system as one coalesced request instead of several 32-byte requests (tools/gather_ceiling.cu did the latter).
// Also measures a streaming read of the same table for reference.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_ceiling2 gather_ceiling2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// PER = bytes per lane (16 or 32), LANES = lanes per granule; G = PER*LANES
template <int PER, int LANES>
__global__ void __launch_bounds__(256) coop_kernel(const char* tab, uint64_t ngran, int iters, uint32_t* out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t group = tid / LANES;
    uint32_t lane = tid % LANES;
    uint32_t acc = 0;
    uint64_t st = mix64(group);
    for (int i = 0; i < iters; ++i) {
        const char* p = tab + (st % ngran) * (uint64_t)(PER * LANES) + lane * PER;
        uint32_t v;
        if (PER == 16) {
            uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
            v = x.x ^ x.y ^ x.z ^ x.w;
        } else {
            uint32_t a, b, c, d, e, f, g, h;
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
            v = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
        }
        acc ^= v;
        // chain: next address depends on the loaded value of lane 0 (like backward search)
        uint32_t v0 = __shfl_sync(0xFFFFFFFFu, v, (threadIdx.x & 31) / LANES * LANES);
        st = mix64(st ^ v0);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void stream_kernel(const uint4* tab, uint64_t n16, uint32_t* out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i < n16; i += stride) { uint4 x = __ldg(tab + i); acc ^= x.x ^ x.y ^ x.z ^ x.w; }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void fill_kernel(uint4* p, uint64_t n16) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) { uint64_t a = mix64(i); p[i] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)i, 7u); }
}

template <typename F>
static float time_ms(F&& f, int reps = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 3.0;
    uint64_t tbytes = (uint64_t)(gb * 1e9) / 4096 * 4096;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs, table %.2f GB\n", prop.name, prop.multiProcessorCount, tbytes / 1e9);
    char* tab; CK(cudaMalloc(&tab, tbytes));
    uint32_t* out; CK(cudaMalloc(&out, 4));
    fill_kernel<<<prop.multiProcessorCount * 8, 256>>>((uint4*)tab, tbytes / 16);
    CK(cudaDeviceSynchronize());
    int sms = prop.multiProcessorCount;
    {
        float ms = time_ms([&] { stream_kernel<<<sms * 8, 256>>>((const uint4*)tab, tbytes / 16, out); });
        printf("stream read                      : %8.3f ms  %8.1f GB/s\n", ms, tbytes / ms / 1e6);
    }
    const int iters = 128;
    for (int occ : {4, 8}) {
        int blocks = sms * occ;
        uint64_t threads = (uint64_t)blocks * 256;
#define RUN(PER, LANES) { float ms = time_ms([&] { coop_kernel<PER, LANES><<<blocks, 256>>>(tab, tbytes / (PER * LANES), iters, out); }); \
        double n = (double)threads / LANES * iters; \
        printf("coop G=%4d (%2d lanes x %2d B) blocks/SM=%d : %8.3f ms  %7.2f Ggran/s  %8.1f GB/s\n", PER * LANES, LANES, PER, occ, ms, n / ms / 1e6, n * PER * LANES / ms / 1e6); }
        RUN(16, 1) RUN(32, 1) RUN(16, 2) RUN(32, 2) RUN(16, 4) RUN(32, 4) RUN(16, 8) RUN(32, 8) RUN(16, 16) RUN(32, 16) RUN(16, 32) RUN(32, 32)
    }
    return 0;
}
