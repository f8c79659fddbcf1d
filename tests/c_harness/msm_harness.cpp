// C++ twin of the reference's integration test (P1A/combined-top-solutions/tests/msm.rs:14-40) over the C ABI of
// libb200msm.so: generate bases and 4 batches of scalars, run multi_scalar_mult_init / multi_scalar_mult through BOTH
// symbol families exactly as the Rust shims do (P1A/storswiftlabs/src/lib.rs:43-98, P1A/yrrid/src/lib.rs:38-90), and
// compare every batch with arkworks' algorithm (the CPU oracle, linked here as the checker only).
// TEST CODE: built and run by tests/test_gpu_parity.py::test_c_harness (needs a GPU).
//   g++ -std=c++17 -O2 msm_harness.cpp -I../../include -L../../2022-entries_b200 -lb200msm -L../../oracle -loracle
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "b200msm.h"

extern "C" {
int oracle_ark_msm(int curve, const void *bases, const void *scalars, size_t n, void *out, int nthreads, int c_override);
int oracle_jac_to_affine(int curve, const void *jac, size_t n, void *aff);
int oracle_gen_points(int curve, uint64_t seed, size_t n, void *aff_out);
int oracle_gen_scalars(int curve, uint64_t seed, size_t n, void *out);
}

static const size_t AFF = 104, SC = 32, PROJ = 144;

// util::generate_points_scalars (P1A/yrrid/src/util.rs:11-35): `distinct` random points, vector-doubled up to len
static void generate(size_t len, size_t batches, size_t distinct, std::vector<uint8_t> &points, std::vector<uint8_t> &scalars) {
    if (distinct > len) distinct = len;
    points.resize(len * AFF);
    oracle_gen_points(0, 2022, distinct, points.data());
    for (size_t have = distinct; have < len;) {
        size_t take = have < len - have ? have : len - have;
        memcpy(points.data() + have * AFF, points.data(), take * AFF);
        have += take;
    }
    scalars.resize(len * batches * SC);
    oracle_gen_scalars(0, 2023, len * batches, scalars.data());
}

static bool same_point(const uint8_t *proj_a, const uint8_t *proj_b) {
    uint8_t a[AFF], b[AFF];
    oracle_jac_to_affine(0, proj_a, 1, a);
    oracle_jac_to_affine(0, proj_b, 1, b);
    return memcmp(a, b, 97) == 0;
}

int main(int argc, char **argv) {
    const int npow = argc > 1 ? atoi(argv[1]) : 12;          // TEST_NPOW
    const size_t batches = 4, n = (size_t)1 << npow;
    std::vector<uint8_t> points, scalars;
    generate(n, batches, 1 << 8, points, scalars);

    // family A: the organisers' shim
    MultiScalarMultContext ctx = {nullptr};
    RustError err = mult_pippenger_init(&ctx, points.data(), n, AFF);
    if (err.code != 0) { fprintf(stderr, "mult_pippenger_init: %d %s\n", err.code, err.message ? err.message : ""); free(err.message); return 2; }
    std::vector<uint8_t> res_a(batches * PROJ, 0);
    err = mult_pippenger_inf(&ctx, (uint64_t *)res_a.data(), points.data(), n, batches, scalars.data(), AFF);
    if (err.code != 0) { fprintf(stderr, "mult_pippenger_inf: %d %s\n", err.code, err.message ? err.message : ""); free(err.message); return 2; }
    mult_pippenger_free(&ctx);

    // family B: yrrid's shim
    void *y = MSMAllocContext(1 << 26, 16);
    if (!y) { fprintf(stderr, "MSMAllocContext failed\n"); return 2; }
    int32_t rc = MSMPreprocessPoints(y, points.data(), (uint32_t)n);
    if (rc != 0) { fprintf(stderr, "Error %d occurred in C code\n", rc); return 2; }
    std::vector<uint8_t> res_b(batches * PROJ, 0);
    rc = MSMRun(y, (uint64_t *)res_b.data(), scalars.data(), (uint32_t)(n * batches));
    if (rc != 0) { fprintf(stderr, "Error %d occurred in C code\n", rc); return 2; }
    MSMFreeContext(y);

    int bad = 0;
    for (size_t b = 0; b < batches; b++) {
        uint8_t want[PROJ];
        oracle_ark_msm(0, points.data(), scalars.data() + b * n * SC, n, want, 8, 0);
        const bool ok_a = same_point(res_a.data() + b * PROJ, want), ok_b = same_point(res_b.data() + b * PROJ, want);
        const bool same_bytes = memcmp(res_a.data() + b * PROJ, res_b.data() + b * PROJ, PROJ) == 0;
        printf("batch %zu: family A %s, family B %s, A == B bytewise %s\n", b, ok_a ? "ok" : "MISMATCH", ok_b ? "ok" : "MISMATCH", same_bytes ? "yes" : "no");
        bad += !ok_a + !ok_b + !same_bytes;
    }
    printf(bad ? "FAILED\n" : "msm_harness ok: %zu batches of 2^%d points, both ABI families, bit-exact vs arkworks' algorithm\n", batches, npow);
    return bad ? 1 : 0;
}
