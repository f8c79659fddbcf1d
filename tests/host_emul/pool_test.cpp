// Host test of csrc/copy_pool.h: many parallel range copies of odd sizes, each compared with the source.
#include "../../2022-entries_b200/csrc/copy_pool.h"
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 4;
    const size_t cap = (size_t)9 << 20;
    std::vector<uint8_t> src(cap), dst(cap + 64);
    for (size_t i = 0; i < cap; i++) src[i] = (uint8_t)((i * 2654435761u) >> 13);
    copy_pool_t pool;
    pool.start(threads);
    const size_t sizes[] = {0, 1, 63, 64, 65, 4097, (size_t)1 << 20, ((size_t)3 << 20) + 17, cap};
    for (int rep = 0; rep < 200; rep++)
        for (size_t n : sizes) {
            memset(dst.data(), 0xA5, n + 64);
            pool.run(dst.data(), src.data() + (rep % 7), n - (n > 7 ? rep % 7 : 0));
            const size_t len = n - (n > 7 ? rep % 7 : 0);
            if (memcmp(dst.data(), src.data() + (rep % 7), len) != 0) { printf("mismatch at size %zu\n", len); return 1; }
            for (int k = 0; k < 64; k++) if (dst[len + k] != 0xA5) { printf("overrun at size %zu\n", len); return 1; }
        }
    printf("ok\n");
    return 0;
}
