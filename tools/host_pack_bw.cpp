// host-side 2-bit packing bandwidth (decides whether packing reads on the host can beat sending raw bytes over PCIe)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static inline uint32_t pack8(uint64_t w, uint64_t& bad) {
    uint64_t v = ((w & 0x7F7F7F7F7F7F7F7Full) + 0x0303030303030303ull) & 0x0303030303030303ull;
    // invalid: byte == 0 or byte > 4  <=>  ((byte - 1) & 0xFC) != 0 (per byte, no borrows needed for the test below)
    uint64_t t = (w | 0x8080808080808080ull) - 0x0101010101010101ull;     // per byte: byte-1 (bit 7 protects from borrows), bit7 set unless byte==0 & ...
    bad |= ((t ^ 0x8080808080808080ull) & 0xFCFCFCFCFCFCFCFCull) | (w & 0x8080808080808080ull);
    // gather 8 x 2 bits: two 32-bit halves with the multiply trick
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    uint32_t a = (lo * 0x00041041u >> 18) & 0xFF, b = (hi * 0x00041041u >> 18) & 0xFF;
    return a | (b << 8);
}
int main(int argc, char** argv) {
    size_t bytes = argc > 1 ? (size_t)atof(argv[1]) : (size_t)1500000000;
    std::vector<uint8_t> src(bytes);
    for (size_t i = 0; i < bytes; ++i) src[i] = 1 + (uint8_t)((i * 2654435761u) >> 30);
    std::vector<uint16_t> dst(bytes / 8 + 8);
    for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
        if (T > (int)std::thread::hardware_concurrency() * 2) break;
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            std::vector<uint64_t> bads(T, 0);
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    size_t words = bytes / 8, b = words * t / T, e = words * (t + 1) / T;
                    const uint64_t* s = reinterpret_cast<const uint64_t*>(src.data());
                    uint64_t bad = 0;
                    for (size_t i = b; i < e; ++i) dst[i] = (uint16_t)pack8(s[i], bad);
                    bads[t] = bad;
                });
            for (auto& x : th) x.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (s < best) best = s;
        }
        printf("threads %2d: %.1f ms  %.1f GB/s\n", T, best * 1e3, bytes / best / 1e9);
    }
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
}
