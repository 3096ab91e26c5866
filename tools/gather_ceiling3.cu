// How many DRAM bytes does one random 32-byte lookup cost on B200, and can the fetch granularity be lowered?
// ncu on exact_search_kernel showed ~3.7 sectors fetched per 32-byte LDG (L1->L2 and L2->DRAM), i.e. the full
// 128-byte line.  This tool times a dependent random 32-byte gather over a 3 GB table with different load
// flavours and device limits; since the gather is DRAM-bandwidth bound, Ggran/s tells the bytes per lookup.
//   argv[1] = table GB, argv[2] = L2 fetch granularity limit to set BEFORE any allocation (0 = leave default)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_ceiling3 gather_ceiling3.cu
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

enum Flavor { NC_NOALLOC = 0, PLAIN, CG, CS, CV, VOLATILE_, NC_2x16, PLAIN_2x16, NC_1x16, NC_EVICT_FIRST, RELAXED_GPU, NFLAVORS };
static const char* kNames[] = {"ld.global.nc.L1::no_allocate.v8", "ld.global.v8 (plain)", "ld.global.cg.v8", "ld.global.cs.v8", "ld.global.cv.v8",
                               "ld.volatile.global.v8", "2 x ld.global.nc.v4 (16B)", "2 x ld.global.v4 (16B)", "1 x ld.global.nc.v4 (16B only)",
                               "ld.global.nc.L2::evict_first.v8", "ld.relaxed.gpu.global.v8"};

template <int F>
__device__ __forceinline__ uint32_t load32(const char* p) {
    uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, h = 0;
    if (F == NC_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == PLAIN)
        asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == CG)
        asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == CS)
        asm volatile("ld.global.cs.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == CV)
        asm volatile("ld.global.cv.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == VOLATILE_)
        asm volatile("ld.volatile.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == NC_2x16) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (F == PLAIN_2x16) {
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (F == NC_1x16) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    } else if (F == NC_EVICT_FIRST)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (F == RELAXED_GPU)
        asm volatile("ld.relaxed.gpu.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int F>
__global__ void __launch_bounds__(256) chain_kernel(const char* tab, uint64_t ngran, int iters, uint32_t* out) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st = mix64(tid);
    for (int i = 0; i < iters; ++i) {
        uint32_t v = load32<F>(tab + (st % ngran) * 32);
        st = mix64(st ^ v);
    }
    if ((uint32_t)st == 0x12345678u) out[0] = (uint32_t)st;
}

__global__ void fill_kernel(uint4* p, uint64_t n16) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) { uint64_t a = mix64(i); p[i] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)i, 7u); }
}

template <typename Fn>
static float time_ms(Fn&& f, int reps = 3) {
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

template <int F>
static void run(const char* tab, uint64_t tbytes, int sms, uint32_t* out) {
    const int iters = 128, occ = 8;
    int blocks = sms * occ;
    float ms = time_ms([&] { chain_kernel<F><<<blocks, 256>>>(tab, tbytes / 32, iters, out); });
    double n = (double)blocks * 256 * iters;
    printf("%-40s : %8.3f ms  %7.2f Ggran/s  (%6.1f GB/s useful; at 4.9 TB/s DRAM that is %5.1f B fetched per lookup)\n", kNames[F], ms,
           n / ms / 1e6, n * 32 / ms / 1e6, 4.9e12 / (n / ms * 1e3));
}

int main(int argc, char** argv) {
    double gb = argc > 1 ? atof(argv[1]) : 3.0;
    int gran = argc > 2 ? atoi(argv[2]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);     // before the first allocation
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("# cudaLimitMaxL2FetchGranularity request %d -> %zu (%s)\n", gran, got, cudaGetErrorString(e));
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("# cudaLimitMaxL2FetchGranularity default = %zu\n", got);
    }
    uint64_t tbytes = (uint64_t)(gb * 1e9) / 4096 * 4096;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs, table %.2f GB\n", prop.name, prop.multiProcessorCount, tbytes / 1e9);
    char* tab; CK(cudaMalloc(&tab, tbytes));
    uint32_t* out; CK(cudaMalloc(&out, 4));
    fill_kernel<<<prop.multiProcessorCount * 8, 256>>>((uint4*)tab, tbytes / 16);
    CK(cudaDeviceSynchronize());
    int sms = prop.multiProcessorCount;
    run<NC_NOALLOC>(tab, tbytes, sms, out);
    run<PLAIN>(tab, tbytes, sms, out);
    run<CG>(tab, tbytes, sms, out);
    run<CS>(tab, tbytes, sms, out);
    run<CV>(tab, tbytes, sms, out);
    run<VOLATILE_>(tab, tbytes, sms, out);
    run<NC_2x16>(tab, tbytes, sms, out);
    run<PLAIN_2x16>(tab, tbytes, sms, out);
    run<NC_1x16>(tab, tbytes, sms, out);
    run<NC_EVICT_FIRST>(tab, tbytes, sms, out);
    run<RELAXED_GPU>(tab, tbytes, sms, out);
    return 0;
}
