/*
 * b200msm.h -- C ABI of libb200msm.so: BLS12-377 / BLS12-381 G1 multi-scalar multiplication on B200 (sm_100a).
 *
 * This is the drop-in seam of the ZPrize-2022 prize-1a harness.  Paths below are relative
 * to the reference tree, P1A = open-division/prize1-msm/prize1a-msm-gpu.  Three groups of symbols:
 *
 *   1. Family A -- the organisers' (sppark) FFI, bound by P1A/storswiftlabs/src/lib.rs:18-41 and implemented in the
 *      reference by P1A/storswiftlabs/cuda/pippenger_inf.cu:51-173.
 *   2. Family B -- the yrrid / combined-top-solutions FFI, bound by P1A/yrrid/src/lib.rs:22-36, declared in
 *      P1A/yrrid/yrrid-msm/MSM.h:65-76.
 *   3. b200msm_* -- the native interface both families are thin adapters over (adds what the reference lacks:
 *      a destroy call, device-resident scalars, per-GPU partial results for point sharding, stage timings).
 *
 * Data layouts are arkworks' in-memory ones (P1B/hardcaml/.../rust_types.h:73-83, yrrid MSM.cu:93-95,378):
 *   G1Affine      104 bytes: x, y = 6 x u64 little-endian Montgomery (R = 2^384), u8 infinity flag at offset 96
 *   scalar         32 bytes: 4 x u64 little-endian integer in [0, r), used as-is (not Montgomery-converted)
 *   G1Projective  144 bytes: Jacobian x, y, z (x/z^2, y/z^3); this library returns (x_affine, y_affine, 1) or the
 *                            arkworks zero (1, 1, 0) -- any representative is valid, the harness compares projectively
 *
 * All functions return 0 on success; -1 = bad argument, -2 = CUDA extension unusable (no device), -3 = bases the
 * selected bucket arithmetic cannot represent (b200msm_set_bucket_arithmetic), any positive value is a cudaError_t.  After a failure b200msm_last_error() describes it.  A context is used from one thread at a
 * time (the harness calls multi_scalar_mult repeatedly on the same context from one thread).
 */
#ifndef B200MSM_H
#define B200MSM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 3. native interface ------------------------------------------------------------------------------------ */
typedef struct b200msm_ctx b200msm_ctx;

enum { B200MSM_BLS12_377 = 0, B200MSM_BLS12_381 = 1 };
enum { B200MSM_BASES_HARNESS = 0, B200MSM_BASES_DISTINCT = 1 };

/* window_bits = 0 and tables = 0 select the size-tuned plan (see DESIGN.md "window plan"). */
int b200msm_create(b200msm_ctx **out, int curve, int device, int window_bits, int tables);
int b200msm_destroy(b200msm_ctx *ctx);
const char *b200msm_last_error(const b200msm_ctx *ctx);   /* ctx may be NULL: error of the last failed create */

/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*); NULL = own stream. */
int b200msm_set_stream(b200msm_ctx *ctx, void *cuda_stream);

/* Bucket arithmetic of the context; call before the bases are set (it drops tables already built).
 *   0  the curve's default: twisted-Edwards extended coordinates on BLS12-377 (7 M unified mixed add), XYZZ on BLS12-381.
 *      The twisted-Edwards model represents every point of G1 and every generic on-curve point; the few on-curve points
 *      WITHOUT an image (y = 0, or s (x + 1) = -1: all outside the prime-order subgroup) make b200msm_set_bases fail with
 *      -3 instead of producing a wrong sum.  Not detected: two bases outside G1 whose difference is one of the model's
 *      exceptional low-order points meeting in one bucket -- impossible for bases in G1, the harness contract.
 *   1  XYZZ (8 M + 2 S mixed add): complete on every on-curve point, the contract of arkworks' add_assign_mixed
 *      (P1B/nickray/driver/algebra/ec/src/models/short_weierstrass_jacobian.rs:662-729).  About 1.35x slower.
 * The environment variable B200MSM_GROUP=xyzz selects 1 for every BLS12-377 context the process creates (Families A, B
 * and the one-call form, which have no such argument). */
int b200msm_set_bucket_arithmetic(b200msm_ctx *ctx, int arithmetic);

/* Pipelining of consecutive MSMs on one context.  The last ~3 ms of an MSM (upper levels of the bucket reduction, finalize,
 * the multi-GPU combine) are chains of dependent curve additions in tiny kernels; they run on an internal high-priority stream
 * so that the next MSM's digit / scatter kernels can fill the GPU meanwhile (the reference overlaps the same part on the host:
 * yrrid MSM.cu:484-494).  b200msm_msm_host and the one-call form always do this internally.  For the device-resident calls
 * (b200msm_msm_device, *_partial_*, b200msm_combine_device) it is opt-in, because it changes WHERE the result is ordered:
 * with pipelining on, the result buffer is written on b200msm_result_stream(ctx) -- consume it there (or after b200msm_sync);
 * off (default), it is ordered on the context's own stream as documented below. */
int b200msm_set_pipelined(b200msm_ctx *ctx, int on);
void *b200msm_result_stream(b200msm_ctx *ctx);      /* cudaStream_t: the context's stream, or the tail stream when pipelined */

/* How a call that takes scalars from host memory splits its FIRST batch so that the upload overlaps the accumulation: `chunks`
 * pieces of sizes 1 : ratio : ratio^2 ... (0 = chosen by size: 3 pieces from 2^24 points, 2 from 2^22; ratio 4).  Only the
 * first piece's upload is exposed; a ratio up to (accumulation time per point) / (upload time per point) keeps every later
 * piece in place before it is needed -- ~5 on one GPU with a PCIe link of its own, ~2 when 8 GPUs share the host. */
int b200msm_set_host_chunks(b200msm_ctx *ctx, int chunks, int ratio);

/* multi_scalar_mult_init: upload `n` arkworks G1Affine records (stride bytes apart, 96 <= stride; a base is at infinity
 * when its flag byte at offset 96 is set (stride > 96) or when x = y = 0, the encoding of the flag-less 96-byte
 * blst_p1_affine records, sppark ec/xyzz_t.hpp:44-45) and build the base tables 2^(c*F*j) * P_i.  Replaces
 * mult_pippenger_init / MSMPreprocessPoints + precomputePointsKernel (P1A/yrrid/yrrid-msm/PrecomputePoints.cu:10-39). */
int b200msm_set_bases(b200msm_ctx *ctx, const void *affine_host, uint64_t n, uint64_t stride);

/* Synthetic bases with known discrete logs, generated on the device (tests and bench.py; 2^26 arkworks records
 * would be 6.5 GiB of host traffic).  HARNESS: `distinct` random multiples of the generator replicated to n, the
 * shape of the reference generator (P1A/yrrid/src/util.rs:15-28); DISTINCT: (seed + i) * G.                  */
int b200msm_set_bases_generated(b200msm_ctx *ctx, uint64_t n, uint64_t seed, int mode, uint64_t distinct);
/* Copy base table 0 back as n x 96-byte (x, y) records (tests). */
int b200msm_get_bases(b200msm_ctx *ctx, void *affine96_host, uint64_t first, uint64_t n);

/* Uniform scalars in [0, r) written to device memory (n x 32 bytes). */
int b200msm_generate_scalars(b200msm_ctx *ctx, void *scalars_dev, uint64_t n, uint64_t seed);

/* multi_scalar_mult, one batch, scalars already in device memory.  Asynchronous on the context's stream: the
 * 144-byte Jacobian result is written to out_dev (device memory) when the stream reaches it. */
int b200msm_msm_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_jacobian_dev);
/* Same, scalars and result in host memory (pinned or pageable); `batches` results of 144 bytes; returns after the
 * results are in place.  Scalar upload of batch b+1 overlaps the compute of batch b. */
int b200msm_msm_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, uint64_t batches, void *out_jacobian_host);

/* Point sharding across GPUs: the local shard's sum as one B200MSM_PARTIAL_BYTES record in device memory
 * (asynchronous; XYZZ coordinates in the device representation, opaque to the caller), and the combine step run after
 * the all-gather: sum `count` such records -> 144-byte Jacobian (device memory). */
#define B200MSM_PARTIAL_BYTES 224
int b200msm_msm_partial_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_xyzz_dev);
int b200msm_combine_device(b200msm_ctx *ctx, const void *xyzz_dev, int count, void *out_jacobian_dev);
/* The partial sum with this GPU's scalars in host memory (pinned for overlap): asynchronous like the device form; the
 * upload is chunked against the accumulation exactly as in b200msm_msm_host.  scalars_host must stay valid until the
 * stream reaches the end of the call (b200msm_sync). */
int b200msm_msm_partial_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, void *out_xyzz_dev);

/* Multi-GPU BUCKET REDUCE-SCATTER over NVLink (one process per GPU; point sharding as above).  With point sharding every GPU
 * ends the accumulation with partial sums of ALL buckets, and reducing them (2 curve additions per bucket, latency-bound)
 * does not shrink with the shard.  After attaching, rank g reduces only the buckets [g B / G, (g + 1) B / G), reading the
 * partial sums of that slice from every peer's HBM through mapped pointers INSIDE the level-0 reduction kernel (no staging
 * copy, no separate collective; cross-GPU ordering by two release/acquire counters per pair).  The result is again one
 * B200MSM_PARTIAL_BYTES record per rank, combined as before (all-gather + b200msm_combine_device).
 *   b200msm_peer_export   after b200msm_set_bases*: writes this rank's handle (CUDA IPC handles of the bucket arrays; `chunks` =
 *                         the most chunks a call will split its upload into: 1 for device-resident scalars).  Needs a plan
 *                         with every window folded into one (all base tables resident), the same n on every rank.
 *   b200msm_peer_attach   handles = world x B200MSM_PEER_HANDLE_BYTES, indexed by rank (exchanged by the caller, e.g.
 *                         torch.distributed.all_gather); world a power of two <= 32.  Collective in effect: every rank must
 *                         attach before any rank runs a peer call, and all ranks run the same sequence of peer calls.
 *   b200msm_msm_peer_partial_device / _host   replace b200msm_msm_partial_device / _host.
 *   b200msm_peer_accumulate_device + b200msm_peer_reduce   the same call in its two halves (tests run several ranks on one GPU,
 *                         where one rank's waiting reduction kernel would starve the other rank's accumulation).
 *   info[0..5] = world, rank, chunks, bytes one call reads from peers over NVLink, and of the last peer call (after a stream
 *                synchronisation): microseconds this rank waited for the slowest peer's accumulation, microseconds of the exchange
 *                kernel (NVLink reads + curve additions). */
#define B200MSM_PEER_HANDLE_BYTES 320
int b200msm_peer_export(b200msm_ctx *ctx, int chunks, void *handle_out);
int b200msm_peer_attach(b200msm_ctx *ctx, int rank, int world, const void *handles);
int b200msm_peer_detach(b200msm_ctx *ctx);
int b200msm_peer_info(b200msm_ctx *ctx, uint64_t info[6]);
int b200msm_msm_peer_partial_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_xyzz_dev);
int b200msm_msm_peer_partial_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, void *out_xyzz_dev);
int b200msm_peer_accumulate_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n);
int b200msm_peer_reduce(b200msm_ctx *ctx, void *out_xyzz_dev);

/* Opt-in subgroup check of the resident bases (table 0): *outside = number of bases P with r P != infinity.  The default
 * bucket arithmetic reduces scalars mod r and folds their top bit into a point negation, which is exact only for bases of
 * order r (G1, the harness contract; arkworks checks it when a G1Affine is deserialised); with B200MSM_CHECK_G1=1 in the
 * environment b200msm_set_bases runs this check itself and fails with -3.  Cost: ~380 curve operations per base. */
int b200msm_check_bases_in_g1(b200msm_ctx *ctx, uint64_t *outside);

int b200msm_sync(b200msm_ctx *ctx);

/* arkworks' VariableBaseMSM::multi_scalar_mul(bases, scalars) in ONE call, no context: the first
 * min(nbases, nscalars) pairs are used, an empty input gives the arkworks zero (ark-ec 0.3.0, restated in the reference
 * at P1A/speakspeak/src/util.rs:51-58), everything in host memory, result = one 144-byte Jacobian point.  This is also the
 * shape of snarkVM's VariableBase::msm -> cuda::msm_cuda dispatch (team-division/prize2-posw-gpu/storswiftlabs/snarkVM/
 * algorithms/src/msm/variable_base/mod.rs:120-157).  Bases are uploaded on every call and used without table
 * precomputation (one table); the library keeps one context per (curve, device) between calls so repeated calls of one
 * size re-use their device buffers -- b200msm_variable_base_release() frees them.  Calls are serialised by a mutex.
 * On failure b200msm_last_error(NULL) describes it.  No CPU path: -2 without a CUDA device, the empty input included. */
int b200msm_variable_base_msm(int curve, int device, const void *affine_host, uint64_t nbases, uint64_t stride,
                              const void *scalars_host, uint64_t nscalars, void *out_jacobian_host);
int b200msm_variable_base_release(void);

/* Plan and measurements of the last run.  info[0..7] = n, window bits c, windows W, tables T, folded windows F,
 * buckets, Lmax, kernel launches of the last MSM.  stage_ms[0..5] = digits+histogram, scan+scatter, pieces,
 * accumulate, bucket reduce, finalize (CUDA events on the context's stream; valid after b200msm_sync). */
int b200msm_get_info(const b200msm_ctx *ctx, uint64_t info[8]);
/* The window plan by itself, without a device: what b200msm_create(window_bits, tables) + set_bases(n) would choose with
 * free_bytes of device memory, the bucket reduction summing over `chunks` chunks (1 for the resident-bases calls).
 * info[0..7] = n, window bits c, windows W, tables T, folded windows F, buckets, bytes of base tables, point references. */
int b200msm_plan_info(int curve, uint64_t n, uint64_t free_bytes, int window_bits, int tables, int chunks, uint64_t info[8]);
int b200msm_get_stage_ms(b200msm_ctx *ctx, float stage_ms[6]);
/* The same six stage times summed over the MSMs run since b200msm_reset_stage_stats (at most the last 32 are kept);
 * *steps = how many MSMs the sums cover.  bench.py derives the per-launch duration of the accumulate kernel from it. */
int b200msm_reset_stage_stats(b200msm_ctx *ctx);
int b200msm_get_stage_ms_sum(b200msm_ctx *ctx, float stage_ms_sum[6], uint64_t *steps);

/* Element-wise device arithmetic on host arrays, for the parity tests of the field and curve layers.
 * Operands and results are arkworks field elements (48 bytes, Montgomery R = 2^384, canonical); the library converts
 * to and from its device representation.  fp op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv, 5 representation round trip.
 * ec op: 0 acc += affine(96 B), 1 acc += xyzz(192 B), 2 acc = 2 acc, 3 acc = 2 affine; acc is XYZZ (4 x 48 B). */
int b200msm_test_fp_op(int curve, int op, const void *a, const void *b, void *r, uint64_t n);
int b200msm_test_ec_op(int curve, int op, void *acc, const void *other, uint64_t n, int reps);
/* Bucket coordinate system of the curve (twisted-Edwards extended on BLS12-377, XYZZ on BLS12-381), through the same
 * table-entry conversion and mixed add the accumulate kernel uses.  a, b: n arkworks (x, y) records of 96 bytes, points
 * of G1; out: n G1Affine records of 104 bytes.  op 0: a + reps*b (mixed add on a table entry), 1: the same through the
 * full add, 2: 2^reps a, 3: a - reps*b. */
int b200msm_test_grp_op(int curve, int op, const void *a96, const void *b96, void *out104, uint64_t n, int reps);

/* ---- 1. Family A: sppark / organisers' harness ---------------------------------------------------------------- */
typedef struct { int code; char *message; } RustError;   /* message is strdup'd; the Rust side frees it
                                                            (P1A/6block/sppark/rust/src/lib.rs:17-24) */
typedef struct { void *context; } MultiScalarMultContext; /* P1A/storswiftlabs/src/lib.rs:18-21 */

RustError mult_pippenger_init(MultiScalarMultContext *context, const void *points_with_infinity, size_t npoints,
                              size_t ffi_affine_sz);
RustError mult_pippenger_inf(MultiScalarMultContext *context, uint64_t *out, const void *points_with_infinity,
                             size_t npoints, size_t batch_size, const void *scalars, size_t ffi_affine_sz);
/* not in the reference (its contexts leak): release a Family-A context */
RustError mult_pippenger_free(MultiScalarMultContext *context);
/* upstream sppark's one-shot call (team-division/prize2-posw-gpu/storswiftlabs/snarkVM/algorithms/cuda/pippenger.cu:19-25,
 * bound by .../Jackytan2018/extern/sppark/sppark/poc/blst-cuda/src/lib.rs:21-53): npoints blst_p1_affine records of
 * 96 bytes (x, y Montgomery, no infinity flag), npoints scalars, out = blst_p1 (144-byte Jacobian); BLS12-377, device 0.
 * An adapter over b200msm_variable_base_msm. */
RustError mult_pippenger(void *out, const void *points, size_t npoints, const void *scalars);

/* ---- 2. Family B: yrrid / combined-top-solutions -------------------------------------------------------------- */
void *MSMAllocContext(int32_t maxPoints, int32_t maxBatches);
int32_t MSMFreeContext(void *context);
int32_t MSMPreprocessPoints(void *context, void *affinePointsPtr, uint32_t points);
int32_t MSMRun(void *context, uint64_t *projectiveResultsPtr, void *scalarsPtr, uint32_t scalars);

#ifdef __cplusplus
}
#endif
#endif
