// pack_stress.cpp -- sanitizer driver for csrc/pack.cpp (the host packer and its worker pool): several caller threads
// pack batches of many sizes with varying thread counts at the same time, every result is checked against the table.
// Built twice by tools/asan/run_pack.sh: -fsanitize=address,undefined and -fsanitize=thread.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"

static uint8_t nt4(uint8_t b) {
    switch (b) {
        case 'A': case 'a': case 0: return 0;
        case 'C': case 'c': case 1: return 1;
        case 'G': case 'g': case 2: return 2;
        case 'T': case 't': case 'U': case 'u': case 3: return 3;
        default: return 4;
    }
}

int main() {
    std::atomic<int> bad{0};
    std::atomic<long> calls{0};
    auto caller = [&](int id) {
        std::mt19937_64 rng(1234 + id);
        const char *alpha = "ACGTacgtUuN\x00\x01\x02\x03";
        for (int it = 0; it < 60 && !bad; it++) {
            const uint64_t n = (it % 7 == 0) ? rng() % 3000000 : rng() % 200000;
            std::vector<uint8_t> b(n);
            const int mode = (int)(rng() % 3);
            for (auto &x : b) x = mode == 0 ? (uint8_t)rng() : mode == 1 ? (uint8_t)"ACGT"[rng() & 3] : (uint8_t)alpha[rng() % 15];
            std::vector<uint8_t> packed((n + 3) / 4 + 1, 0xEE);
            const uint64_t cap = 1 + rng() % (n + 1);
            std::vector<uint32_t> exc(cap);
            uint64_t n_exc = 0;
            const int threads = 1 + (int)(rng() % 6);
            if (hulk_b200_pack_bases(b.data(), n, packed.data(), exc.data(), cap, &n_exc, threads) != HULK_B200_OK) { bad = 1; break; }
            calls++;
            uint64_t want = 0;
            for (uint64_t i = 0; i < n; i++) {
                const uint8_t c = nt4(b[i]);
                if (c == 4) {
                    if (want < cap && exc[want] != i) { bad = 2; break; }
                    want++;
                } else if (((packed[i >> 2] >> (2 * (i & 3))) & 3) != c) { bad = 3; break; }
            }
            if (want != n_exc) bad = 4;
            if (packed[(n + 3) / 4] != 0xEE) bad = 5;                       // nothing behind the last byte was touched
        }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < 4; i++) th.emplace_back(caller, i);
    for (auto &t : th) t.join();
    printf("pack stress: %ld calls from 4 concurrent callers, verdict %d\n", calls.load(), bad.load());
    return bad ? 1 : 0;
}
