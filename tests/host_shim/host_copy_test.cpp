// CPU test of csrc/host_copy.hpp (the copy threads and the page-residency probe of the pageable download path).
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../wgpu-voxel-engine_b200/csrc/host_copy.hpp"

static int fails = 0;
#define EXPECT(c)                                                        \
    do {                                                                 \
        if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); ++fails; } \
    } while (0)

int main()
{
    std::mt19937 rng(7);
    const size_t sizes[] = {0, 1, 4095, 4096, 4097, (256u << 10) - 1, 256u << 10, (256u << 10) + 1, (5u << 20) + 13, 16u << 20};
    for (unsigned helpers : {0u, 1u, 3u, 7u}) {
        ParallelCopier pc(helpers);
        for (size_t n : sizes) {
            std::vector<unsigned char> src(n + 64), dst(n + 64, 0xEE);
            for (auto &b : src) b = (unsigned char)rng();
            for (int rep = 0; rep < 3; ++rep) {                       // the pool is reused across copies
                pc.copy(dst.data() + 7, src.data() + 3, n);           // unaligned on both sides
                EXPECT(memcmp(dst.data() + 7, src.data() + 3, n) == 0);
                bool guard = true;
                for (size_t i = 0; i < 7; ++i) guard = guard && dst[i] == 0xEE;
                for (size_t i = n + 7; i < n + 64; ++i) guard = guard && dst[i] == 0xEE;
                EXPECT(guard);                                        // nothing written outside [dst, dst + n)
            }
        }
    }
    // page residency: an anonymous mapping is fresh until it is written
    const size_t len = 8u << 20;
    void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    EXPECT(m != MAP_FAILED);
    EXPECT(host_is_fresh(m, len));
    EXPECT(host_is_fresh((char *)m + 100, len - 1000));
    memset(m, 1, len / 4);                                            // a quarter touched: still mostly fresh
    EXPECT(host_is_fresh(m, len));
    memset(m, 1, len);
    EXPECT(!host_is_fresh(m, len));
    munmap(m, len);
    std::vector<unsigned char> touched(4u << 20, 3);
    EXPECT(!host_is_fresh(touched.data(), touched.size()));
    EXPECT(!host_is_fresh(touched.data(), 100));                      // too small to bother
    printf(fails ? "host_copy: %d failures\n" : "host_copy ok\n", fails);
    return fails ? 1 : 0;
}
