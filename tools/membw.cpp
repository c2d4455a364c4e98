// Host DRAM read bandwidth by thread count (reference for the packer's numbers): g++ -O2 -mavx2 -pthread tools/membw.cpp
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
int main() {
    const size_t N = 2ull << 30;
    uint8_t *a = (uint8_t *)aligned_alloc(4096, N);
    for (size_t i = 0; i < N; i += 4096) a[i] = (uint8_t)i;
    for (int T : {1, 2, 4, 8, 12, 16}) {
        std::vector<uint64_t> sink(T * 16);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                const size_t b0 = N / T * t, b1 = N / T * (t + 1);
                __m256i s = _mm256_setzero_si256();
                for (size_t i = b0; i + 64 <= b1; i += 64) {
                    s = _mm256_add_epi64(s, _mm256_load_si256((const __m256i *)(a + i)));
                    s = _mm256_add_epi64(s, _mm256_load_si256((const __m256i *)(a + i + 32)));
                }
                sink[t * 16] = (uint64_t)_mm256_extract_epi64(s, 0);
            });
        for (auto &x : th) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%2d threads: %.1f GB/s read\n", T, N / dt / 1e9);
    }
    return 0;
}
