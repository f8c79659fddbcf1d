// Host build of csrc/digits.cuh (TEST ONLY): the signed-digit window decomposition of k_digits / k_digits_bin, run on the
// CPU so tests/test_host_emul.py can check keys and point references against Python integers without a GPU.
#include <cstddef>
#include "../../2022-entries_b200/csrc/digits.cuh"
using namespace b200msm;

template <class C> static int run(const uint32_t *scalars, uint32_t i, uint32_t gi, uint32_t table_stride, int c, int W, int F, int maxw16,
                                  uint32_t *keys, uint32_t *refs, int *windows) {
    int cnt = 0;
    if (maxw16) for_each_digit<C, 16>(scalars, i, gi, table_stride, c, W, F, [&](int w, uint32_t k, uint32_t r) { keys[cnt] = k; refs[cnt] = r; windows[cnt] = w; cnt++; });
    else for_each_digit<C>(scalars, i, gi, table_stride, c, W, F, [&](uint32_t k, uint32_t r) { keys[cnt] = k; refs[cnt] = r; windows[cnt] = -1; cnt++; });
    return cnt;
}

// scalars: n x 8 words; pairs of scalar i (base number first + i) are written at out[i * 32 ...]; counts[i] = how many
extern "C" void emul_digits(int curve, const uint32_t *scalars, size_t n, uint32_t first, uint32_t table_stride, int c, int W, int F, int maxw16,
                            uint32_t *keys, uint32_t *refs, int *windows, int *counts) {
    for (size_t i = 0; i < n; i++)
        counts[i] = curve == 0 ? run<Bls12_377>(scalars, (uint32_t)i, first + (uint32_t)i, table_stride, c, W, F, maxw16, keys + i * 32, refs + i * 32, windows + i * 32)
                               : run<Bls12_381>(scalars, (uint32_t)i, first + (uint32_t)i, table_stride, c, W, F, maxw16, keys + i * 32, refs + i * 32, windows + i * 32);
}
