// msm.cu -- host orchestrator and C ABI of libb200msm.so (include/b200msm.h).
//
// Stands in for the reference's L3 layer: yrrid MSMContext (MSM.cu:139-503), sppark
// pippenger_t + pippenger_inf.cu, matter-labs msm.cu schedule_execution.  One context = one device, one set of
// base tables, one stream; every MSM is the fixed kernel sequence in run_msm() below.
#include <cuda_runtime.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/b200msm.h"
#include "kernels_common.cuh"
#include "launch.h"
#include "copy_pool.h"

using namespace b200msm;

static thread_local std::string g_create_error;

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return ctx->fail((int)e__, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)


// bucket arithmetic of a context: 0 = BLS12-377 twisted Edwards (default), 1 = BLS12-381 XYZZ, 2 = BLS12-377 XYZZ
#define BY_GROUP(kind, ...)                                              \
    switch (kind) {                                                      \
    case 0: { typedef Bls12_377 C; __VA_ARGS__; } break;                 \
    case 1: { typedef Bls12_381 C; __VA_ARGS__; } break;                 \
    default: { typedef Bls12_377_sw C; __VA_ARGS__; } break;             \
    }

struct red_step_t {
    int kind;                 // 0: reduce_slices through the piece table, 1: reduce_slices, 2: block_sum
    const xyzz_t *in;
    uint32_t n_out, block, m; // block_sum: n_out groups of m inputs
    xyzz_t *outS, *outT;
};

struct b200msm_ctx {
    int curve = 0, device = 0;
    int kind = 0;             // BY_GROUP selector: curve, except 2 = BLS12-377 with XYZZ buckets (b200msm_set_bucket_arithmetic)
    int req_c = 0, req_tables = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, conv_stream = nullptr;   // conv: entry conversion of the one-call form (high priority)
    // The TAIL of an MSM -- the upper levels of the bucket reduction, finalize, and (multi-GPU) the combine step: a few dozen
    // dependent curve additions in kernels of 1 .. 64 blocks, ~3 ms that no amount of hardware shortens -- runs on a stream of
    // its own (high priority), so that the NEXT MSM's digit and scatter kernels fill the GPU meanwhile.  The reference overlaps
    // the same part of its pipeline, there on the host (yrrid MSM.cu:484-494: hostReduce of batch k under the kernels of batch
    // k + 1).  msm_host / the one-call form always work this way; the device-resident calls do when the caller has opted in
    // (b200msm_set_pipelined) and consumes results on b200msm_result_stream or after b200msm_sync.
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t ev_l0 = nullptr, ev_tail = nullptr;
    bool tail_pending = false;          // ev_tail has been recorded: the next level 0 must not overwrite the level buffers before it
    bool pipeline = false;              // b200msm_set_pipelined
    cudaEvent_t ev_conv = nullptr;
    bool own_stream = true;
    std::string err;
    int sticky = 0;

    // plan
    uint64_t n = 0;
    bool plan_ok = false;     // alloc_plan completed for n (a later set_bases of the same n reuses every buffer)
    int c = 0, W = 0, T = 0, F = 0;
    uint32_t nbuckets = 0, lmax = 0, max_pieces = 0;
    uint64_t max_refs = 0;
    int launches = 0;
    int chunks_alloc = 1;     // copies of piece_off / res (one per chunk of a chunked MSM)
    int host_chunks = 0;      // msm_host: split the first batch into this many chunks (0 = by size)
    int chunk_ratio = 0;      // ... of sizes 1 : q : q^2 ... (B200MSM_CHUNK_RATIO / b200msm_set_host_chunks; 0 = 4)
    cudaEvent_t ev_chunk[8] = {};
    int acc_mix = 8;          // accumulate kernel flavour: 8 = integer pipe (default), 0 = FP64 pipe (B200MSM_ACC=fp64)

    // device memory
    uint8_t *tables = nullptr;          // T tables of n entries in the curve's bucket coordinates (grp.cuh)
    size_t entry_bytes = 0;
    uint32_t *infmask = nullptr;
    uint32_t *bad = nullptr;            // bases without an image in the bucket coordinate system (counted by make_entries)
    uint32_t *counts = nullptr, *bucket_off = nullptr, *cursor = nullptr, *piece_off = nullptr;
    uint32_t *piece_start = nullptr, *piece_len = nullptr, *order = nullptr, *len_hist = nullptr, *len_cursor = nullptr;
    uint32_t *scan_tmp = nullptr, *refs = nullptr;
    uint32_t *bin_cursor = nullptr;     // two-level scatter (large problems only)
    uint2 *staging = nullptr;
    int bin_shift = 0;
    uint32_t nbins = 0;
    size_t bs2_smem = 0;                // dynamic shared memory k_bin_scatter_tiled has been configured for
    bool scatter_atomic = false;        // B200MSM_SCATTER=atomic: the round-1 second level (one global atomic per reference)
    xyzz_t *res = nullptr, *red_buf = nullptr, *part_xyzz = nullptr;
    xyzz_t *fin_scratch = nullptr;      // k_finalize: one product per folded window
    uint32_t *fin_ticket = nullptr;
    uint8_t *scal_buf[2] = {nullptr, nullptr};
    uint8_t *raw_buf[2] = {nullptr, nullptr};   // one-call form: arkworks records of one upload step, double-buffered
    affine_t *sw_buf = nullptr;                 //                the same step in the device representation
    size_t raw_cap = 0;
    cudaEvent_t ev_raw_copied[2] = {}, ev_raw_free[2] = {};
    static const size_t SLOT_BYTES = 32u << 20;
    uint8_t *slot[2] = {nullptr, nullptr};      // pinned staging for pageable sources (h2d)
    cudaEvent_t slot_ev[2] = {};
    bool slot_busy[2] = {false, false};
    uint64_t slot_k = 0;
    copy_pool_t *pool = nullptr;
    double h2d_ms = 0, h2d_wait_ms = 0;   // host time spent in h2d / of it waiting for a free slot (B200MSM_TRACE)
    int plan_chunks = 1;      // chunks the cost model charges the bucket reduction for (one-call form)
    uint8_t *out_dev = nullptr;      // batches x 144
    uint64_t out_dev_cap = 0;
    std::vector<red_step_t> red_steps;
    fin_args_t fin;
    red_src_t *src_tab = nullptr;    // device: sources of the level-0 reduction (this GPU's chunks, then the peers' chunks)
    int src_tab_chunks = -1;         // chunk count the local table was last written for
    // multi-GPU bucket reduce-scatter (b200msm_peer_*): rank g reduces buckets [g nb / G, (g + 1) nb / G) of every GPU
    struct peer_t {
        int rank = 0, world = 0, chunks = 1;
        uint32_t epoch = 0;
        uint32_t *flags = nullptr;           // this GPU's counters: ready[PEER_MAX], done[PEER_MAX] (peers write them over NVLink)
        uint32_t **flag_ptrs = nullptr;      // device array: every peer's flag buffer, mapped
        void *mapped[PEER_MAX][3] = {};      // cudaIpcOpenMemHandle results to close again
        red_src_t *src_tab = nullptr;        // device: world x chunks sources
        xyzz_t *red_buf = nullptr;
        xyzz_t *merged = nullptr;            // this rank's slice of the buckets, summed over every GPU (k_peer_merge)
        std::vector<red_step_t> steps;
        fin_args_t fin;
        uint64_t nvlink_bytes = 0;           // bytes one level-0 launch reads from peers (algorithmic)
        cudaEvent_t ev[3] = {};              // last MSM: before the wait for the peers' `ready`, after it, after k_peer_merge
    } peer;
    // Stage boundaries: one event set per CHUNK (stages 0-3: digits, scatter, pieces, accumulate); the set of an MSM's
    // last chunk also takes the reduction and finalize boundaries (ev_last = 6) and closes the MSM.  A plain MSM is one set.
    static const int EV_RING = 64;
    cudaEvent_t evr[EV_RING][7] = {};
    uint8_t ev_last[EV_RING] = {};      // index of the last event recorded in the set (4 or 6)
    uint64_t ev_sets = 0;               // sets recorded since the last stats reset
    uint64_t ev_steps = 0;              // MSMs completed since the last stats reset
    cudaEvent_t *ev = evr[0];           // the set the current chunk records into
    cudaEvent_t ev_copied[2] = {}, ev_free[2] = {};

    int fail(int code, const std::string &msg) {
        err = msg;
        if (code > 0) sticky = code;
        fprintf(stderr, "b200msm: %s\n", msg.c_str());
        return code;
    }
    void peer_release() {
        for (auto &m : peer.mapped) for (void *&q : m) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
        void *ptrs[] = {peer.flags, peer.flag_ptrs, peer.src_tab, peer.red_buf, peer.merged};
        for (void *q : ptrs) if (q) cudaFree(q);
        peer.flags = nullptr; peer.flag_ptrs = nullptr; peer.src_tab = nullptr; peer.red_buf = nullptr; peer.merged = nullptr;
        peer.world = 0; peer.epoch = 0; peer.steps.clear();
        for (auto &e : peer.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    }
    void free_plan() {
        void *ptrs[] = {tables, infmask, bad, counts, bucket_off, cursor, piece_off, piece_start, piece_len, order, len_hist, len_cursor,
                        scan_tmp, refs, res, red_buf, part_xyzz, fin_scratch, fin_ticket, scal_buf[0], scal_buf[1], bin_cursor, staging, raw_buf[0], raw_buf[1], sw_buf};
        for (void *p : ptrs) if (p) cudaFree(p);
        tables = nullptr; infmask = nullptr; bad = nullptr; counts = bucket_off = cursor = piece_off = piece_start = piece_len = order = len_hist = len_cursor = nullptr;
        scan_tmp = refs = nullptr; res = red_buf = part_xyzz = nullptr; fin_scratch = nullptr; fin_ticket = nullptr; scal_buf[0] = scal_buf[1] = nullptr; bin_cursor = nullptr; staging = nullptr; nbins = 0;
        raw_buf[0] = raw_buf[1] = nullptr; sw_buf = nullptr; raw_cap = 0;
        if (src_tab) cudaFree(src_tab);
        src_tab = nullptr; src_tab_chunks = -1;
        peer_release();
        red_steps.clear();
        n = 0;
        plan_ok = false;
        chunks_alloc = 1;
    }
};

static int scalar_bits(int kind) { return kind == 0 ? Bls12_377::WINDOW_BITS : kind == 1 ? Bls12_381::WINDOW_BITS : Bls12_377_sw::WINDOW_BITS; }

// ---------------------------------------------------------------------------------------------
// window plan: choose c (window bits), T (tables) and F (folded windows) for n points.
// cost model: W*n mixed adds (10 modmul) + 2 full adds (14 modmul) per bucket, buckets = F * 2^(c-1);
// T limited by device memory (96 n bytes per table) and by the 31-bit point reference (T*n < 2^31).
static bool choose_plan(b200msm_ctx *ctx, uint64_t n, size_t free_bytes) {
    const int sb = scalar_bits(ctx->kind);
    uint64_t tmax_mem = (uint64_t)((double)free_bytes * 0.72 / ((double)ctx->entry_bytes * (double)n));
    uint64_t tmax_ref = ((1ull << 31) - 1) / n;
    uint64_t tmax = tmax_mem < tmax_ref ? tmax_mem : tmax_ref;
    if (tmax < 1) tmax = 1;
    if (ctx->req_tables > 0 && (uint64_t)ctx->req_tables < tmax) tmax = ctx->req_tables;
    int best_c = 0;
    double best = 1e300;
    int lo = ctx->req_c > 0 ? ctx->req_c : 4, hi = ctx->req_c > 0 ? ctx->req_c : 24;
    for (int c = lo; c <= hi; c++) {
        int W = (sb + c - 1) / c;
        int T = (int)(tmax < (uint64_t)W ? tmax : W);
        int F = (W + T - 1) / T;
        double buckets = (double)F * (double)(1ull << (c - 1));
        if (buckets > (double)(1u << 27)) continue;
        double red = (28.0 + 14.0 * (ctx->plan_chunks - 1)) * buckets;      // + one full add per bucket per extra chunk
        double cost = 10.0 * (double)W * (double)n + red;
        if (cost < best) { best = cost; best_c = c; }
    }
    if (best_c == 0) return false;      // every candidate needs more than 2^27 buckets (a forced window of 25 / 26 bits with few tables)
    ctx->c = best_c;
    ctx->W = (sb + best_c - 1) / best_c;
    int T = (int)(tmax < (uint64_t)ctx->W ? tmax : ctx->W);
    ctx->F = (ctx->W + T - 1) / T;
    ctx->T = (ctx->W + ctx->F - 1) / ctx->F;
    ctx->nbuckets = (uint32_t)ctx->F << (ctx->c - 1);
    return true;
}

// the context's stream AND the tail stream its MSMs hand their last kernels to
static cudaError_t sync_streams(b200msm_ctx *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && ctx->tail_stream) e = cudaStreamSynchronize(ctx->tail_stream);
    if (e == cudaSuccess) ctx->tail_pending = false;
    return e;
}

template <class T> static cudaError_t dmalloc(T **p, size_t count) { return cudaMalloc((void **)p, count * sizeof(T)); }

// Bucket-reduction plan (see k_reduce_slices): level sizes, buffers carved from one allocation, the terms k_finalize sums.
// div = 1: all buckets of this GPU.  div = G > 1 (bucket reduce-scatter, one folded window): the same tree restricted to the
// rank-th 1/G of every level -- the slice of a level's outputs is the reduction of the slice of its inputs, so the only
// place the position shows is the last level, whose 32 / G entries weigh (rank * 32 / G + k).
static int make_red_plan(b200msm_ctx *ctx, uint32_t div, uint32_t rank, xyzz_t **buf_out, std::vector<red_step_t> &steps, fin_args_t &fin) {
    struct lvl { uint32_t n_in, m, n_out, block, nT; };
    std::vector<lvl> lv;
    {
        uint32_t cur = 1u << (ctx->c - 1);
        bool first = true;
        while (first || cur > 32) {
            uint32_t m = cur >= 256 ? 8 : cur / 32;
            if (m < 1) m = 1;
            uint32_t n_out = cur / m / div, block = n_out < 128 ? n_out : 128;
            if (n_out == 0 || (cur / m) % div) return ctx->fail(-1, "bucket reduce-scatter: the window is too small to split over this many GPUs");
            lv.push_back({cur / div, m, n_out, block, n_out / block});
            cur = cur / m;
            first = false;
        }
    }
    const uint32_t F = ctx->F;
    size_t need = 0;
    for (auto &l : lv) need += (size_t)F * (l.n_out + l.nT + (l.nT > 32 ? 32 : 0));
    CK(dmalloc(buf_out, need + 1));
    xyzz_t *bump = *buf_out;
    auto take = [&](size_t cnt) { xyzz_t *p = bump; bump += cnt; return p; };
    steps.clear();
    memset(&fin, 0, sizeof(fin));
    fin.F = ctx->F; fin.c = ctx->c;
    const xyzz_t *cur_arr = nullptr;          // level 0 reads through the source table
    uint32_t shift = 0;
    for (size_t li = 0; li < lv.size(); li++) {
        lvl &l = lv[li];
        xyzz_t *S = take((size_t)F * l.n_out), *Tt = take((size_t)F * l.nT);
        steps.push_back({li == 0 ? 0 : 1, cur_arr, F * l.n_out, l.block, l.m, S, Tt});
        uint32_t nT = l.nT;
        if (nT > 32) {
            xyzz_t *T2 = take((size_t)F * 32);
            steps.push_back({2, Tt, F * 32, 0, nT / 32, T2, nullptr});
            Tt = T2; nT = 32;
        }
        if (fin.nterms >= FIN_MAX_TERMS - 2) return ctx->fail(-1, "reduction plan has too many terms");
        fin.term[fin.nterms++] = {Tt, nT, nT, shift, 0, 0};              // 2^shift * sum_t T_t of this level
        uint32_t lg = 0; while ((1u << lg) < l.m) lg++;
        shift += lg;
        cur_arr = S;
    }
    uint32_t last_n = lv.back().n_out;
    fin.term[fin.nterms++] = {cur_arr, last_n, last_n, 0, 0, 0};                     // sum of all buckets
    fin.term[fin.nterms++] = {cur_arr, last_n, last_n, shift, 1, rank * last_n};     // 2^shift * sum_k (base + k) * S_last[k]
    return 0;
}

static int alloc_plan(b200msm_ctx *ctx, uint64_t n) {
    if (ctx->plan_ok && ctx->n == n) return 0;        // same size again (one-shot calls): plan and buffers stand
    int rc = 0;
    ctx->free_plan();
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (!choose_plan(ctx, n, free_b)) return ctx->fail(-1, "window_bits / tables combination needs more than 2^27 buckets");
    ctx->n = n;
    ctx->max_refs = (uint64_t)ctx->W * n;
    if (ctx->max_refs >= (1ull << 32)) return ctx->fail(-1, "too many point references for 32-bit list offsets");
    // Lmax: uniform scalars give ~W*n/buckets points per bucket; pieces only split buckets far above that
    uint64_t mean = ctx->max_refs / ctx->nbuckets + 1;
    uint32_t lmax = 256;
    while (lmax < 4 * mean && lmax < 4096) lmax <<= 1;
    ctx->lmax = lmax;
    ctx->max_pieces = ctx->nbuckets + (uint32_t)(ctx->max_refs / lmax) + 1;
    const uint32_t nb = ctx->nbuckets;
    CK(dmalloc(&ctx->tables, (size_t)ctx->T * n * ctx->entry_bytes));
    CK(dmalloc(&ctx->infmask, (size_t)cdiv(n, 32) + 1));
    CK(dmalloc(&ctx->bad, 1));
    CK(cudaMemsetAsync(ctx->bad, 0, 4, ctx->stream));
    CK(dmalloc(&ctx->counts, (size_t)nb + 1));
    CK(dmalloc(&ctx->bucket_off, (size_t)nb + 1));
    CK(dmalloc(&ctx->cursor, (size_t)nb + 1));
    CK(dmalloc(&ctx->piece_off, (size_t)nb + 1));
    CK(dmalloc(&ctx->piece_start, (size_t)ctx->max_pieces));
    CK(dmalloc(&ctx->piece_len, (size_t)ctx->max_pieces));
    CK(dmalloc(&ctx->order, (size_t)ctx->max_pieces));
    CK(dmalloc(&ctx->len_hist, (size_t)lmax + 1));
    CK(dmalloc(&ctx->len_cursor, (size_t)lmax + 1));
    CK(dmalloc(&ctx->scan_tmp, (size_t)cdiv(nb, SCAN_TPB * SCAN_EPT) + 2));
    CK(dmalloc(&ctx->refs, (size_t)ctx->max_refs + 1));
    CK(dmalloc(&ctx->res, (size_t)ctx->max_pieces));
    CK(dmalloc(&ctx->part_xyzz, 1));
    CK(dmalloc(&ctx->fin_scratch, (size_t)ctx->F));
    CK(dmalloc(&ctx->fin_ticket, 1));
    CK(cudaMemsetAsync(ctx->fin_ticket, 0, 4, ctx->stream));
    // two-level scatter when the list array is far beyond L2: bins of <= 24 MB of list each, at most BIN_MAX bins
    ctx->nbins = 0;
    // from 256 MB of lists on.  Measured (profiles/r02_scaling.txt): 2^23 points, c = 23, 369 MB of lists (the 8-GPU shard of
    // 2^26): direct scatter 3.42 ms, two-level 1.66 ms; 2^22 points, c = 20, 218 MB: direct 0.79 ms, two-level 1.25 ms
    uint64_t bin_min_bytes = 256ull << 20;
    uint64_t bin_bytes = 24ull << 20;
    if (const char *e = getenv("B200MSM_BIN_MIN_MB")) bin_min_bytes = (uint64_t)atoll(e) << 20;      // tests: force the two-level path
    if (const char *e = getenv("B200MSM_BIN_KB")) { bin_bytes = (uint64_t)atoll(e) << 10; if (bin_bytes == 0) bin_bytes = 1024; }   // ... with many bins at small n
    if (ctx->max_refs * 4 > bin_min_bytes && ctx->W <= 16 && !getenv("B200MSM_DIRECT_SCATTER")) {
        uint64_t want = (ctx->max_refs * 4 + bin_bytes - 1) / bin_bytes;               // bins needed
        int lg_nb = 0; while ((1ull << lg_nb) < nb) lg_nb++;
        int lg_bins = 0; while ((1ull << lg_bins) < want && lg_bins < 8) lg_bins++;
        while (lg_nb - lg_bins > BS2_MAX_SHIFT && lg_bins < 8) lg_bins++;             // the tiled second level keeps one bin's counters in shared memory
        if (lg_bins > lg_nb) lg_bins = lg_nb;
        ctx->bin_shift = lg_nb - lg_bins;
        ctx->nbins = (uint32_t)((nb + (1u << ctx->bin_shift) - 1) >> ctx->bin_shift);
        CK(dmalloc(&ctx->bin_cursor, (size_t)ctx->nbins + 1));
        CK(dmalloc(&ctx->staging, (size_t)ctx->max_refs + 1));
    }

    CK(dmalloc(&ctx->src_tab, (size_t)RED_MAX_SRC));
    rc = make_red_plan(ctx, 1, 0, &ctx->red_buf, ctx->red_steps, ctx->fin); if (rc) return rc;
    ctx->plan_ok = true;
    return 0;
}

// dst (device) <- src (caller's host memory) on stream st.  Pinned sources and small copies go straight to cudaMemcpyAsync;
// pageable sources are staged through the context's pinned slots by the copy pool.  Returns once src has been read.
static int h2d(b200msm_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    size_t min_bytes = 4u << 20;
    if (const char *e = getenv("B200MSM_STAGE_MIN_KB")) min_bytes = (size_t)atoll(e) << 10;        // tests: stage small copies too
    bool direct = bytes < min_bytes || getenv("B200MSM_NO_STAGING");
    if (!direct) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, src) != cudaSuccess) cudaGetLastError();      // unknown to the runtime: pageable
        else direct = a.type != cudaMemoryTypeUnregistered;
    }
    if (direct) { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)); return 0; }
    if (!ctx->pool) {
        for (int i = 0; i < 2; i++) {
            if (!ctx->slot[i]) CK(cudaHostAlloc((void **)&ctx->slot[i], b200msm_ctx::SLOT_BYTES, cudaHostAllocDefault));      // (a retry after a failed first attempt)
            if (!ctx->slot_ev[i]) CK(cudaEventCreateWithFlags(&ctx->slot_ev[i], cudaEventDisableTiming));
        }
        int p = (int)std::thread::hardware_concurrency() / 2;
        if (const char *e = getenv("B200MSM_COPY_THREADS")) p = atoi(e);
        p = p < 1 ? 1 : (p > 16 ? 16 : p);
        ctx->pool = new copy_pool_t();
        ctx->pool->start(p);
    }
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    for (size_t off = 0; off < bytes; off += b200msm_ctx::SLOT_BYTES, ctx->slot_k++) {
        const size_t len = bytes - off < b200msm_ctx::SLOT_BYTES ? bytes - off : b200msm_ctx::SLOT_BYTES;
        const int b = (int)(ctx->slot_k & 1);
        if (ctx->slot_busy[b]) { const auto tw = std::chrono::steady_clock::now(); CK(cudaEventSynchronize(ctx->slot_ev[b])); ctx->h2d_wait_ms += ms_since(tw); }
        ctx->pool->run(ctx->slot[b], (const uint8_t *)src + off, len);
        CK(cudaMemcpyAsync((uint8_t *)dst + off, ctx->slot[b], len, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ctx->slot_ev[b], st));
        ctx->slot_busy[b] = true;
    }
    ctx->h2d_ms += ms_since(t0);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// a11: base tables.  Per chunk of <= 2^22 points: table 0 entry = the point itself, table j = 2^(c*F) * table j-1,
// doubled in XYZZ, normalised to short-Weierstrass affine (one inversion per 16 points), then converted to the curve's
// table-entry form.  Only two affine chunk buffers and one XYZZ chunk buffer of scratch are alive.
struct table_scratch_t {           // freed on every exit path, error returns included
    affine_t *sw[2] = {nullptr, nullptr}; xyzz_t *tmp = nullptr; uint8_t *raw = nullptr; uint32_t cap = 0;
    ~table_scratch_t() { cudaFree(sw[0]); cudaFree(sw[1]); cudaFree(tmp); cudaFree(raw); }
};
static int scratch_alloc(b200msm_ctx *ctx, table_scratch_t &sc, uint32_t cap) {
    sc.cap = cap;
    CK(dmalloc(&sc.sw[0], cap)); CK(dmalloc(&sc.sw[1], cap)); CK(dmalloc(&sc.tmp, cap));
    return 0;
}
// sc.sw[0] holds cnt affine points starting at index `first`: emit their entries into every table
// infmask: the infinity bits of these cnt points (bit 0 = point `first`), or nullptr when none is at infinity
template <class C> static void emit_tables(b200msm_ctx *ctx, table_scratch_t &sc, uint32_t first, uint32_t cnt, const uint32_t *infmask) {
    const size_t eb = ctx->entry_bytes;
    int cur = 0;
    for (int j = 0; j < ctx->T; j++) {
        if (j > 0) {
            LaunchCold<C>::double_n(ctx->stream, sc.sw[cur], cnt, ctx->c * ctx->F, sc.tmp);
            LaunchCold<C>::batch_normalize(ctx->stream, sc.tmp, cnt, sc.sw[cur ^ 1], infmask, ctx->bad);
            cur ^= 1;
        }
        LaunchCold<C>::make_entries(ctx->stream, sc.sw[cur], cnt, ctx->tables + ((size_t)j * ctx->n + first) * eb, ctx->acc_mix == 0, infmask, ctx->bad);
    }
    if (cur != 0) { affine_t *t = sc.sw[0]; sc.sw[0] = sc.sw[1]; sc.sw[1] = t; }
}

template <class C> static int gen_bases_t(b200msm_ctx *ctx, uint64_t seed, int mode, uint64_t distinct, const fp &gx, const fp &gy) {
    const uint32_t n = (uint32_t)ctx->n;
    uint32_t gen_n = mode == 0 ? (uint32_t)(distinct < n ? distinct : n) : n;
    const uint32_t chunk = 1u << 22;
    table_scratch_t sc;
    int rc = scratch_alloc(ctx, sc, gen_n < chunk ? gen_n : chunk); if (rc) return rc;
    for (uint32_t first = 0; first < gen_n; first += chunk) {
        uint32_t cnt = gen_n - first < chunk ? gen_n - first : chunk;
        LaunchCold<C>::gen_points(ctx->stream, first, cnt, seed, mode, (uint32_t)(distinct ? distinct : 1), gx, gy, sc.tmp);
        LaunchCold<C>::batch_normalize(ctx->stream, sc.tmp, cnt, sc.sw[0], nullptr, nullptr);
        emit_tables<C>(ctx, sc, first, cnt, nullptr);
    }
    if (gen_n < n) {        // the replicas of a point have the same entry in every table
        const uint32_t vec = (uint32_t)(ctx->entry_bytes / 16);
        for (int j = 0; j < ctx->T; j++) {
            uint64_t work = (uint64_t)(n - gen_n) * vec;
            k_replicate<<<cdiv(work, 256), 256, 0, ctx->stream>>>((uint4 *)(ctx->tables + (size_t)j * n * ctx->entry_bytes), gen_n, n, vec);
        }
    }
    CK(cudaMemsetAsync(ctx->infmask, 0, ((size_t)cdiv(n, 32) + 1) * 4, ctx->stream));
    CK(cudaGetLastError());
    CK(sync_streams(ctx));
    return 0;
}

// Bases the bucket coordinate system cannot represent (twisted Edwards on BLS12-377: points with y = 0 or s (x + 1) = -1,
// which lie outside the prime-order subgroup) were counted while the entries were built: refuse them loudly.  The XYZZ
// arithmetic (b200msm_set_bucket_arithmetic) takes every on-curve point, as arkworks does
// (P1B/nickray/driver/algebra/ec/src/models/short_weierstrass_jacobian.rs:662-729).
static int check_bad_entries(b200msm_ctx *ctx) {
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(sync_streams(ctx));
    if (bad == 0) return 0;
    CK(cudaMemsetAsync(ctx->bad, 0, 4, ctx->stream));
    ctx->n = 0;                          // the tables are unusable: a later MSM call fails with "bases not set"
    ctx->plan_ok = false;
    return ctx->fail(-3, std::to_string(bad) + " table entries belong to bases outside the prime-order subgroup with no twisted-Edwards image; "
                         "select the XYZZ bucket arithmetic (b200msm_set_bucket_arithmetic(ctx, 1) or B200MSM_GROUP=xyzz)");
}

template <class C> static int set_bases_t(b200msm_ctx *ctx, const void *affine_host, uint64_t n, uint64_t stride) {
    const uint64_t chunk = 1ull << 22;
    table_scratch_t sc;
    int rc = scratch_alloc(ctx, sc, (uint32_t)(n < chunk ? n : chunk)); if (rc) return rc;
    CK(cudaMalloc((void **)&sc.raw, (size_t)((n < chunk ? n : chunk) * stride)));
    uint8_t *raw = sc.raw;
    CK(cudaMemsetAsync(ctx->infmask, 0, ((size_t)cdiv(n, 32) + 1) * 4, ctx->stream));
    for (uint64_t first = 0; first < n; first += chunk) {
        uint64_t cnt = n - first < chunk ? n - first : chunk;
        rc = h2d(ctx, raw, (const uint8_t *)affine_host + first * stride, (size_t)(cnt * stride), ctx->stream); if (rc) return rc;
        LaunchCold<C>::convert_points(ctx->stream, raw, (uint32_t)cnt, (uint32_t)stride, sc.sw[0], ctx->infmask + first / 32);
        emit_tables<C>(ctx, sc, (uint32_t)first, (uint32_t)cnt, ctx->infmask + first / 32);
        CK(cudaGetLastError());
        CK(sync_streams(ctx));
    }
    return check_bad_entries(ctx);
}

// generator in the device representation (P1B/nickray/driver/bls12-377/src/curves/g1.rs:159-165)
static void generator(int curve, fp &gx, fp &gy) {
    static const uint32_t g377x[12] = {0xb21be9efu, 0xeab9b16eu, 0xffcd394eu, 0xd5481512u, 0xbd37cb5cu, 0x188282c8u, 0xaa9d41bbu, 0x85951e2cu, 0xbf87ff54u, 0xc8fc6225u, 0xfe740a67u, 0x008848deu};
    static const uint32_t g377y[12] = {0x559c8ea6u, 0xfd82de55u, 0x34a9591au, 0xc2fe3d36u, 0x4fb82305u, 0x6d182ad4u, 0xca3e52d9u, 0xbd7fb348u, 0x30afeec4u, 0x1f674f5du, 0xc5102effu, 0x01914a69u};
    static const uint32_t g381x[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu, 0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
    static const uint32_t g381y[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu, 0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
    // canonical -> x * 2^392 mod p on the host (schoolbook, only here: 392 modular doublings), then 28-bit limbs
    auto to_mont = [&](const uint32_t *v, fp &out) {
        uint32_t a[13];
        for (int i = 0; i < 12; i++) a[i] = v[i];
        a[12] = 0;
        static const uint32_t p377[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu, 0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
        static const uint32_t p381[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        auto p = [&](int i) { return curve == 0 ? p377[i] : p381[i]; };
        for (int s = 0; s < FP_LIMBS * FP_BITS; s++) {
            uint32_t carry = 0;
            for (int i = 0; i < 13; i++) { uint32_t nc = a[i] >> 31; a[i] = (a[i] << 1) | carry; carry = nc; }
            // a >= p ? a -= p
            bool ge = a[12] != 0;
            if (!ge) { ge = true; for (int i = 11; i >= 0; i--) if (a[i] != p(i)) { ge = a[i] > p(i); break; } }
            if (ge) { uint64_t br = 0; for (int i = 0; i < 12; i++) { uint64_t t = (uint64_t)a[i] - p(i) - br; a[i] = (uint32_t)t; br = (t >> 32) & 1; } a[12] -= (uint32_t)br; }
        }
        fp_abi packed;
        for (int i = 0; i < 12; i++) packed.w[i] = a[i];
        Fp<Bls12_377>::repack_from_abi(out, packed);      // pure bit slicing, curve-independent
    };
    to_mont(curve == 0 ? g377x : g381x, gx);
    to_mont(curve == 0 ? g377y : g381y, gy);
}

// ---------------------------------------------------------------------------------------------
// one MSM on the context's stream.  scalars: device, n x 32 bytes.  Exactly one of out_jac / out_xyzz is set.
// Stages 1-5 for the bases [first, first + count) with their scalars at scalars_dev: digits, grouping, accumulation into
// this chunk's own bucket-sum array.  A plain MSM is one chunk covering everything; msm_host splits the first batch of
// a call so that the upload of chunk k+1 overlaps the accumulation of chunk k (the reduction then sums over chunks).
template <class C> static int run_chunk_t(b200msm_ctx *ctx, const void *scalars_dev, uint32_t first, uint32_t count, int chunk) {
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)ctx->n, nb = ctx->nbuckets, lmax = ctx->lmax;
    const uint32_t *sc = (const uint32_t *)scalars_dev;
    uint32_t *piece_off = ctx->piece_off + (size_t)chunk * (nb + 1);
    xyzz_t *res = ctx->res + (size_t)chunk * ctx->max_pieces;
    int launches = 0;
    {
        const int slot = (int)(ctx->ev_sets % b200msm_ctx::EV_RING);
        ctx->ev = ctx->evr[slot];
        ctx->ev_last[slot] = 4;
        ctx->ev_sets++;
        cudaEventRecord(ctx->ev[0], st);
    }
    CK(cudaMemsetAsync(ctx->counts, 0, ((size_t)nb + 1) * 4, st));
    LaunchHot<C>::digits(0, st, sc, ctx->infmask, count, first, n, ctx->c, ctx->W, ctx->F, ctx->counts, nullptr);
    launches++;
    cudaEventRecord(ctx->ev[1], st);
    // bucket offsets (and scatter cursors), then scatter the point references
    const uint32_t sblocks = cdiv(nb, SCAN_TPB * SCAN_EPT);
    k_scan_block_sums<0><<<sblocks, SCAN_TPB, 0, st>>>(ctx->counts, nb, lmax, ctx->scan_tmp);
    k_scan_top<<<1, 1024, 0, st>>>(ctx->scan_tmp, sblocks);
    k_scan_apply<0><<<sblocks, SCAN_TPB, 0, st>>>(ctx->counts, nb, lmax, ctx->scan_tmp, ctx->bucket_off, ctx->cursor);
    if (ctx->nbins > 1) {
        k_bin_init<<<cdiv(ctx->nbins, 256), 256, 0, st>>>(ctx->bucket_off, nb, ctx->bin_shift, ctx->nbins, ctx->bin_cursor);
        LaunchHot<C>::digits_bin(st, sc, ctx->infmask, count, first, n, ctx->c, ctx->W, ctx->F, ctx->bin_shift, ctx->nbins, ctx->bin_cursor, ctx->staging);
        if (ctx->bin_shift <= BS2_MAX_SHIFT && !ctx->scatter_atomic) {
            const size_t hist_bytes = (size_t)4 << ctx->bin_shift;
            if (ctx->bs2_smem < hist_bytes) { CK(cudaFuncSetAttribute(k_bin_scatter_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes)); ctx->bs2_smem = hist_bytes; }
            uint32_t tile = BS2_TILE;
            if (const char *e = getenv("B200MSM_SCATTER_TILE_K")) tile = (uint32_t)atoi(e) << 10;      // experiments
            k_bin_scatter_tiled<<<cdiv((uint64_t)ctx->W * count, tile), BS2_TPB, hist_bytes, st>>>(ctx->staging, ctx->bucket_off, nb, ctx->bin_shift, tile, ctx->cursor, ctx->refs);
        } else {
            k_bin_scatter<<<cdiv((uint64_t)ctx->W * count, BINSC_TPB * BINSC_EPT), BINSC_TPB, 0, st>>>(ctx->staging, ctx->bucket_off + nb, ctx->cursor, ctx->refs);
        }
        launches += 2;
    } else {
        LaunchHot<C>::digits(1, st, sc, ctx->infmask, count, first, n, ctx->c, ctx->W, ctx->F, ctx->cursor, ctx->refs);
    }
    launches += 4;
    cudaEventRecord(ctx->ev[2], st);
    // pieces.  Bucket reduce-scatter: peers read piece_off / res of the previous MSM over NVLink -- nothing of them is overwritten
    // before every peer has published `done` for that MSM.
    if (ctx->peer.world && ctx->peer.epoch > 1 && chunk == 0) {
        k_peer_wait<<<1, PEER_MAX, 0, st>>>(ctx->peer.flags + PEER_MAX, (uint32_t)ctx->peer.world, ctx->peer.epoch - 1);
        launches++;
    }
    k_scan_block_sums<1><<<sblocks, SCAN_TPB, 0, st>>>(ctx->counts, nb, lmax, ctx->scan_tmp);
    k_scan_top<<<1, 1024, 0, st>>>(ctx->scan_tmp, sblocks);
    k_scan_apply<1><<<sblocks, SCAN_TPB, 0, st>>>(ctx->counts, nb, lmax, ctx->scan_tmp, piece_off, nullptr);
    CK(cudaMemsetAsync(ctx->len_hist, 0, ((size_t)lmax + 1) * 4, st));
    const size_t hist_smem = ((size_t)lmax + 1) * 4;
    k_piece_fill<<<cdiv(nb, PIECE_TPB), PIECE_TPB, hist_smem, st>>>(ctx->bucket_off, piece_off, nb, lmax, ctx->piece_start, ctx->piece_len, ctx->len_hist);
    k_piece_len_scan<<<1, 1024, 0, st>>>(ctx->len_hist, lmax, ctx->len_cursor);
    const uint32_t *npieces_ptr = piece_off + nb;
    k_piece_order<<<cdiv(ctx->max_pieces, PIECE_TPB), PIECE_TPB, hist_smem, st>>>(ctx->piece_len, npieces_ptr, lmax, ctx->len_cursor, ctx->order);
    launches += 6;
    cudaEventRecord(ctx->ev[3], st);
    // the hot loop
    LaunchHot<C>::accumulate(st, ctx->max_pieces, ctx->tables, ctx->refs, ctx->order, ctx->piece_start, ctx->piece_len, npieces_ptr, res, ctx->acc_mix);
    launches++;
    cudaEventRecord(ctx->ev[4], st);
    ctx->launches = (chunk == 0 ? 0 : ctx->launches) + launches;
    CK(cudaGetLastError());
    return 0;
}
// Stages 6-7: bucket reduction over the chunks' bucket sums, window combine, export.
// the level-0 source table of this GPU alone: entry ch = chunk ch's arrays, whatever number of chunks a call uses, so it is
// written once per allocation of the arrays (a synchronising copy here would break the upload / compute overlap of msm_host)
static int local_sources(b200msm_ctx *ctx, int nchunks) {
    if (ctx->src_tab_chunks >= nchunks) return 0;
    red_src_t h[RED_MAX_SRC];
    const int cnt = ctx->chunks_alloc;
    for (int ch = 0; ch < cnt; ch++) h[ch] = {ctx->res + (size_t)ch * ctx->max_pieces, ctx->piece_off + (size_t)ch * ((size_t)ctx->nbuckets + 1)};
    CK(cudaMemcpy(ctx->src_tab, h, sizeof(red_src_t) * cnt, cudaMemcpyHostToDevice));
    ctx->src_tab_chunks = cnt;
    return 0;
}
// peer = false: every bucket of this GPU (nchunks sources).  peer = true: this rank's slice of the buckets of every GPU
// (world x chunks sources over NVLink, waiting for the peers' `ready` counters inside the level-0 kernel).
// join: the context's stream waits for the tail before the call returns (the result is ordered on that stream, as the
// device-resident calls promise); without it the caller orders its consumer after ev_tail itself.
template <class C> static int run_reduce_t(b200msm_ctx *ctx, int nchunks, void *out_jac, void *out_xyzz, bool peer, bool join) {
    cudaStream_t st = ctx->stream;
    int launches = 0;
    if (!peer) { int rc = local_sources(ctx, nchunks); if (rc) return rc; }
    auto &P = ctx->peer;
    if (ctx->tail_pending) { CK(cudaStreamWaitEvent(st, ctx->ev_tail, 0)); ctx->tail_pending = false; }      // level 0 overwrites what the previous tail reads
    bool first = true;
    for (const red_step_t &s : peer ? P.steps : ctx->red_steps) {
        if (!first && st == ctx->stream) {          // level 0 is done on the context's stream: everything after it is the tail
            CK(cudaEventRecord(ctx->ev_l0, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_l0, 0));
            st = ctx->tail_stream;
        }
        first = false;
        if (s.kind == 2) LaunchCold<C>::block_sum(st, s.in, s.n_out, s.m, s.outS);
        else if (s.kind == 1) LaunchCold<C>::reduce_slices(st, s.in, nullptr, 0, 0, s.n_out, s.block, s.m, s.outS, s.outT, nullptr, 0, 0);
        else if (!peer) LaunchCold<C>::reduce_slices(st, nullptr, ctx->src_tab, (uint32_t)nchunks, 0, s.n_out, s.block, s.m, s.outS, s.outT, nullptr, 0, 0);
        else {              // the exchange (k_peer_merge reads the peers over NVLink), `done` to the peers, then the ordinary level 0 on the merged slice
            const uint32_t nslice = ctx->nbuckets / (uint32_t)P.world;
            // the wait for the slowest peer's accumulation gets a launch (and a pair of events) of its own, so that the time of
            // the exchange kernel is transfer + additions only (b200msm_peer_info)
            cudaEventRecord(P.ev[0], st);
            k_peer_wait<<<1, PEER_MAX, 0, st>>>(P.flags, (uint32_t)P.world, P.epoch);
            cudaEventRecord(P.ev[1], st);
            LaunchCold<C>::peer_merge(st, P.src_tab, (uint32_t)(P.world * P.chunks), (uint32_t)P.rank * nslice, nslice, P.merged, P.flags, (uint32_t)P.world, P.epoch);
            cudaEventRecord(P.ev[2], st);
            k_peer_signal<<<1, PEER_MAX, 0, st>>>(P.flag_ptrs, (uint32_t)P.world, PEER_MAX + (uint32_t)P.rank, P.epoch);      // done reading
            LaunchCold<C>::reduce_slices(st, P.merged, nullptr, 0, 0, s.n_out, s.block, s.m, s.outS, s.outT, nullptr, 0, 0);
            launches += 3;
        }
        launches++;
    }
    if (st == ctx->stream) {                        // a one-level plan: the hand-over still happens before finalize
        CK(cudaEventRecord(ctx->ev_l0, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_l0, 0));
        st = ctx->tail_stream;
    }
    cudaEventRecord(ctx->ev[5], st);
    LaunchCold<C>::finalize(st, peer ? P.fin : ctx->fin, (xyzz_t *)out_xyzz, (uint32_t *)out_jac, ctx->fin_scratch, ctx->fin_ticket);
    launches++;
    cudaEventRecord(ctx->ev[6], st);
    CK(cudaEventRecord(ctx->ev_tail, st));
    ctx->tail_pending = true;
    if (join) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tail, 0)); ctx->tail_pending = false; }
    ctx->ev_last[(ctx->ev_sets - 1) % b200msm_ctx::EV_RING] = 6;
    ctx->ev_steps++;
    ctx->launches += launches;
    CK(cudaGetLastError());
    return 0;
}
// the context's stream waits for the last tail (before something on it consumes a result)
static int join_tail(b200msm_ctx *ctx) {
    if (ctx->tail_pending) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tail, 0)); ctx->tail_pending = false; }
    return 0;
}
static int run_chunk(b200msm_ctx *ctx, const void *scalars_dev, uint32_t first, uint32_t count, int chunk) {
    int rc = 0;
    BY_GROUP(ctx->kind, rc = run_chunk_t<C>(ctx, scalars_dev, first, count, chunk));
    return rc;
}
static int run_reduce(b200msm_ctx *ctx, int nchunks, void *out_jac, void *out_xyzz, bool peer = false, bool join = true) {
    int rc = 0;
    BY_GROUP(ctx->kind, rc = run_reduce_t<C>(ctx, nchunks, out_jac, out_xyzz, peer, join));
    return rc;
}
// one MSM on the context's stream.  scalars: device, n x 32 bytes.  Exactly one of out_jac / out_xyzz is set.
static int run_msm(b200msm_ctx *ctx, const void *scalars_dev, void *out_jac, void *out_xyzz, bool join) {
    int rc = run_chunk(ctx, scalars_dev, 0, (uint32_t)ctx->n, 0);
    if (rc) return rc;
    return run_reduce(ctx, 1, out_jac, out_xyzz, false, join);
}
// piece_off / res hold one copy per chunk
static int ensure_chunks(b200msm_ctx *ctx, int k) {
    if (k <= ctx->chunks_alloc) return 0;
    if (ctx->peer.world) return ctx->fail(-1, "the bucket arrays are mapped by peer GPUs and cannot grow: attach with the chunk count the calls will use");
    CK(sync_streams(ctx));
    cudaFree(ctx->piece_off); cudaFree(ctx->res);
    ctx->piece_off = nullptr; ctx->res = nullptr;
    CK(dmalloc(&ctx->piece_off, (size_t)k * ((size_t)ctx->nbuckets + 1)));
    CK(dmalloc(&ctx->res, (size_t)k * ctx->max_pieces));
    ctx->chunks_alloc = k;
    ctx->src_tab_chunks = -1;          // the level-0 source table points into the old arrays
    return 0;
}

static int check_ctx(b200msm_ctx *ctx) {
    if (!ctx) return -1;
    if (ctx->sticky) return ctx->sticky;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ctx->fail((int)e, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int b200msm_create(b200msm_ctx **out, int curve, int device, int window_bits, int tables) {
    if (!out || (curve != 0 && curve != 1) || window_bits < 0 || window_bits > 26 || (window_bits > 0 && window_bits < 2) || tables < 0) {
        g_create_error = "b200msm_create: bad argument";
        return -1;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("b200msm_create: no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
        fprintf(stderr, "%s\n", g_create_error.c_str());
        return -2;
    }
    if (device < 0 || device >= ndev) { g_create_error = "b200msm_create: bad device index"; return -1; }
    b200msm_ctx *ctx = new b200msm_ctx();
    ctx->curve = curve; ctx->device = device; ctx->req_c = window_bits; ctx->req_tables = tables;
    if (const char *e = getenv("B200MSM_ACC")) ctx->acc_mix = strcmp(e, "fp64") == 0 ? 0 : 8;
    if (const char *e = getenv("B200MSM_HOST_CHUNKS")) ctx->host_chunks = atoi(e);
    if (const char *e = getenv("B200MSM_SCATTER")) ctx->scatter_atomic = strcmp(e, "atomic") == 0;
    if (const char *e = getenv("B200MSM_CHUNK_RATIO")) { int v = atoi(e); if (v >= 2 && v <= 16) ctx->chunk_ratio = v; }
    if (const char *g = getenv("B200MSM_GROUP")) if (curve == 0 && strcmp(g, "xyzz") == 0) ctx->kind = 2;
    if (curve == 1) ctx->kind = 1;
    BY_GROUP(ctx->kind, ctx->entry_bytes = LaunchCold<C>::entry_bytes());
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); e = cudaStreamCreateWithPriority(&ctx->conv_stream, cudaStreamNonBlocking, hi); }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_conv, cudaEventDisableTiming);
    if (e == cudaSuccess) { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); e = cudaStreamCreateWithPriority(&ctx->tail_stream, cudaStreamNonBlocking, hi); }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_l0, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_tail, cudaEventDisableTiming);
    for (int r = 0; r < b200msm_ctx::EV_RING; r++) for (int i = 0; i < 7 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->evr[r][i]);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) { e = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming); if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming); }
    for (int i = 0; i < 2 && e == cudaSuccess; i++) { e = cudaEventCreateWithFlags(&ctx->ev_raw_copied[i], cudaEventDisableTiming); if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_raw_free[i], cudaEventDisableTiming); }
    if (e != cudaSuccess) {
        g_create_error = std::string("b200msm_create: ") + cudaGetErrorString(e);
        delete ctx;
        return (int)e;
    }
    *out = ctx;
    return 0;
}

int b200msm_destroy(b200msm_ctx *ctx) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->free_plan();
    if (ctx->out_dev) cudaFree(ctx->out_dev);
    for (auto &row : ctx->evr) for (auto &e : row) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) { if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]); if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]); }
    for (int i = 0; i < 2; i++) { if (ctx->ev_raw_copied[i]) cudaEventDestroy(ctx->ev_raw_copied[i]); if (ctx->ev_raw_free[i]) cudaEventDestroy(ctx->ev_raw_free[i]); }
    delete ctx->pool;
    for (int i = 0; i < 2; i++) { if (ctx->slot[i]) cudaFreeHost(ctx->slot[i]); if (ctx->slot_ev[i]) cudaEventDestroy(ctx->slot_ev[i]); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->conv_stream) cudaStreamDestroy(ctx->conv_stream);
    if (ctx->ev_conv) cudaEventDestroy(ctx->ev_conv);
    if (ctx->tail_stream) cudaStreamDestroy(ctx->tail_stream);
    if (ctx->ev_l0) cudaEventDestroy(ctx->ev_l0);
    if (ctx->ev_tail) cudaEventDestroy(ctx->ev_tail);
    delete ctx;
    return 0;
}

const char *b200msm_last_error(const b200msm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b200msm_set_stream(b200msm_ctx *ctx, void *cuda_stream) {
    int rc = check_ctx(ctx); if (rc) return rc;
    CK(sync_streams(ctx));
    if (cuda_stream) {
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else if (!ctx->own_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return 0;
}

int b200msm_set_bucket_arithmetic(b200msm_ctx *ctx, int arithmetic) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (arithmetic != 0 && arithmetic != 1) return ctx->fail(-1, "b200msm_set_bucket_arithmetic: 0 = the curve's default, 1 = XYZZ");
    const int kind = ctx->curve == 1 ? 1 : (arithmetic == 1 ? 2 : 0);
    if (kind == ctx->kind) return 0;
    CK(sync_streams(ctx));
    ctx->free_plan();                     // entry size and tables change: bases must be set again
    ctx->kind = kind;
    BY_GROUP(ctx->kind, ctx->entry_bytes = LaunchCold<C>::entry_bytes());
    return 0;
}

int b200msm_set_pipelined(b200msm_ctx *ctx, int on) {
    int rc = check_ctx(ctx); if (rc) return rc;
    CK(sync_streams(ctx));
    ctx->pipeline = on != 0;
    return 0;
}
void *b200msm_result_stream(b200msm_ctx *ctx) { return ctx ? (void *)(ctx->pipeline ? ctx->tail_stream : ctx->stream) : nullptr; }

int b200msm_set_host_chunks(b200msm_ctx *ctx, int chunks, int ratio) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (chunks < 0 || chunks > 8 || (ratio != 0 && (ratio < 2 || ratio > 16))) return ctx->fail(-1, "b200msm_set_host_chunks: chunks 0..8 (0 = by size), ratio 0 (default) or 2..16");
    ctx->host_chunks = chunks;
    ctx->chunk_ratio = ratio;
    return 0;
}

int b200msm_set_bases(b200msm_ctx *ctx, const void *affine_host, uint64_t n, uint64_t stride) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!affine_host || n == 0 || n >= (1ull << 31) || stride < 96 || (stride & 7)) return ctx->fail(-1, "b200msm_set_bases: bad argument");
    rc = alloc_plan(ctx, n); if (rc) return rc;
    BY_GROUP(ctx->kind, rc = set_bases_t<C>(ctx, affine_host, n, stride));
    if (rc == 0 && ctx->kind != 2) {
        const char *e = getenv("B200MSM_CHECK_G1");
        if (e && atoi(e)) {              // opt-in: the default arithmetic folds scalars mod r, which is only right for bases of order r
            uint64_t outside = 0;
            rc = b200msm_check_bases_in_g1(ctx, &outside);
            if (rc == 0 && outside) {
                ctx->n = 0; ctx->plan_ok = false;
                return ctx->fail(-3, std::to_string(outside) + " bases lie outside the prime-order subgroup; select the XYZZ bucket arithmetic (b200msm_set_bucket_arithmetic(ctx, 1) or B200MSM_GROUP=xyzz)");
            }
        }
    }
    return rc;
}

int b200msm_check_bases_in_g1(b200msm_ctx *ctx, uint64_t *outside) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!outside || ctx->n == 0) return ctx->fail(-1, "b200msm_check_bases_in_g1: bad argument (bases not set)");
    CK(cudaMemsetAsync(ctx->bad, 0, 4, ctx->stream));
    BY_GROUP(ctx->kind, LaunchCold<C>::check_order_r(ctx->stream, ctx->tables, ctx->infmask, (uint32_t)ctx->n, ctx->acc_mix == 0, ctx->bad));
    uint32_t cnt = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(sync_streams(ctx));
    CK(cudaMemsetAsync(ctx->bad, 0, 4, ctx->stream));
    *outside = cnt;
    return 0;
}

int b200msm_set_bases_generated(b200msm_ctx *ctx, uint64_t n, uint64_t seed, int mode, uint64_t distinct) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (n == 0 || n >= (1ull << 31) || (mode != 0 && mode != 1) || (mode == 0 && distinct == 0) || (mode == 1 && seed == 0))
        return ctx->fail(-1, "b200msm_set_bases_generated: bad argument");
    rc = alloc_plan(ctx, n); if (rc) return rc;
    fp gx, gy;
    generator(ctx->curve, gx, gy);
    BY_GROUP(ctx->kind, rc = gen_bases_t<C>(ctx, seed, mode, distinct, gx, gy));
    return rc;
}

int b200msm_get_bases(b200msm_ctx *ctx, void *affine96_host, uint64_t first, uint64_t n) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!affine96_host || first + n > ctx->n) return ctx->fail(-1, "b200msm_get_bases: bad argument");
    CK(sync_streams(ctx));
    uint32_t *tmp = nullptr;
    CK(cudaMalloc((void **)&tmp, (size_t)n * 96));
    BY_GROUP(ctx->kind, LaunchCold<C>::export_points(ctx->stream, ctx->tables + first * ctx->entry_bytes, (uint32_t)n, tmp, ctx->acc_mix == 0));
    CK(sync_streams(ctx));
    CK(cudaMemcpy(affine96_host, tmp, (size_t)n * 96, cudaMemcpyDeviceToHost));
    cudaFree(tmp);
    return 0;
}

int b200msm_generate_scalars(b200msm_ctx *ctx, void *scalars_dev, uint64_t n, uint64_t seed) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_dev || n == 0 || n >= (1ull << 32)) return ctx->fail(-1, "b200msm_generate_scalars: bad argument");
    if (ctx->curve == 0) LaunchCold<Bls12_377>::gen_scalars(ctx->stream, (uint32_t)n, seed, (uint32_t *)scalars_dev);
    else LaunchCold<Bls12_381>::gen_scalars(ctx->stream, (uint32_t)n, seed, (uint32_t *)scalars_dev);
    CK(cudaGetLastError());
    return 0;
}

int b200msm_msm_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_jacobian_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_dev || !out_jacobian_dev || ctx->n == 0 || n != ctx->n) return ctx->fail(-1, "b200msm_msm_device: bad argument (bases not set or length mismatch)");
    return run_msm(ctx, scalars_dev, out_jacobian_dev, nullptr, !ctx->pipeline);
}

int b200msm_msm_partial_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_xyzz_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_dev || !out_xyzz_dev || ctx->n == 0 || n != ctx->n) return ctx->fail(-1, "b200msm_msm_partial_device: bad argument");
    return run_msm(ctx, scalars_dev, nullptr, out_xyzz_dev, !ctx->pipeline);
}

int b200msm_combine_device(b200msm_ctx *ctx, const void *xyzz_dev, int count, void *out_jacobian_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!xyzz_dev || count < 1 || !out_jacobian_dev) return ctx->fail(-1, "b200msm_combine_device: bad argument");
    // pipelined contexts: the partials were produced (and gathered) on the tail stream; the combine step continues there
    BY_GROUP(ctx->kind, LaunchCold<C>::combine(ctx->pipeline ? ctx->tail_stream : ctx->stream, (const xyzz_t *)xyzz_dev, count, (uint32_t *)out_jacobian_dev));
    CK(cudaGetLastError());
    return 0;
}

// Upload one batch of scalars from host memory and run it, the batch split into chunks of bases when it is large so that
// chunk k+1 uploads while chunk k accumulates (each chunk into its own bucket sums, one shared reduction at the end).
// Nothing else can hide the upload of the first batch of a call.  Stands in for the 1/4 + 3/4 split of the reference's
// fastest entry (P1A/combined-top-solutions/combined-msm/MSM.cu:419-435, 477-492).  buf: device staging of n x 32 bytes.
static int default_host_chunks(const b200msm_ctx *ctx, uint64_t n) {
    // measured end to end (profiles/r02_chunks.txt): 2^26: 3 chunks of sizes 1 : 4 : 16 236.6 ms (1 : 5 : 25 236.1), 4 of 1 : 2 : 4 : 8 (round 1) 240.7,
    // 2 of 1 : 5 238.2; 2^23 (the 8-GPU shard): 2 chunks 1 : 8 39.7 ms, one chunk 40.8
    int nch = ctx->host_chunks > 0 ? ctx->host_chunks : (n >= (1ull << 24) ? 3 : (n >= (1ull << 22) ? 2 : 1));
    return nch > 8 ? 8 : nch;
}
static int peer_clear_unused_chunks(b200msm_ctx *ctx, int used);
static int upload_and_run(b200msm_ctx *ctx, const uint8_t *scalars_host, uint64_t n, uint8_t *buf, cudaEvent_t buf_free, void *out_jac, void *out_xyzz, bool peer, bool join) {
    int rc = 0;
    int nch = peer ? ctx->peer.chunks : default_host_chunks(ctx, n);
    if ((uint64_t)nch > n) nch = 1;
    if (nch > 1) { rc = ensure_chunks(ctx, nch); if (rc) return rc; }
    CK(cudaStreamWaitEvent(ctx->copy_stream, buf_free, 0));
    // chunk sizes double (1 : 2 : 4 : ...): only the first chunk's upload is exposed, and accumulation (3.8 ns per point)
    // is so much slower than the upload (0.6 ns per point) that every later chunk is in place long before it is needed
    // chunk sizes grow geometrically (1 : q : q^2 ...): only the first chunk's upload is exposed.  q up to ~5 keeps every later
    // chunk in place before it is needed (accumulation 3.1 ns per point against 0.6 ns of upload)
    // default ratio 4: within 0.5 ms of the best measured on one GPU (1 : 5 : 25 at 2^26, 1 : 8 at 2^23) and still safe when the
    // ranks of an 8-GPU box share the host's links (27 GB/s per GPU: 8 x B200, 2^23 points each: 1 : 4 41.6 ms, 1 : 8 44.1 ms)
    uint64_t q = ctx->chunk_ratio > 0 ? (uint64_t)ctx->chunk_ratio : 4, pw[9] = {1};
    for (int i = 1; i <= nch; i++) pw[i] = pw[i - 1] * q;
    const uint64_t denom = (pw[nch] - 1) / (q - 1);          // 1 + q + ... + q^(nch-1)
    auto upto = [&](int ch) { return n * ((pw[ch] - 1) / (q - 1)) / denom; };
    int used = 0;
    for (int ch = 0; ch < nch; ch++) {
        const uint64_t first = upto(ch), end = ch + 1 == nch ? n : upto(ch + 1);
        const uint64_t cnt = end - first;
        if (cnt == 0) continue;                  // tiny n: fewer, non-empty chunks
        rc = h2d(ctx, buf + first * 32, scalars_host + first * 32, (size_t)cnt * 32, ctx->copy_stream); if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_chunk[used], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[used], 0));
        rc = run_chunk(ctx, buf + first * 32, (uint32_t)first, (uint32_t)cnt, used);
        if (rc) return rc;
        used++;
    }
    if (peer) {
        rc = peer_clear_unused_chunks(ctx, used); if (rc) return rc;
        k_peer_signal<<<1, PEER_MAX, 0, ctx->stream>>>(ctx->peer.flag_ptrs, (uint32_t)ctx->peer.world, (uint32_t)ctx->peer.rank, ctx->peer.epoch);
        return run_reduce(ctx, ctx->peer.chunks, nullptr, out_xyzz, true, join);
    }
    return run_reduce(ctx, used, out_jac, out_xyzz, false, join);
}

int b200msm_msm_partial_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, void *out_xyzz_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_host || !out_xyzz_dev || ctx->n == 0 || n != ctx->n) return ctx->fail(-1, "b200msm_msm_partial_host: bad argument");
    if (!ctx->scal_buf[0]) CK(cudaMalloc((void **)&ctx->scal_buf[0], (size_t)n * 32));
    CK(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    return upload_and_run(ctx, (const uint8_t *)scalars_host, n, ctx->scal_buf[0], ctx->ev_free[0], nullptr, out_xyzz_dev, false, !ctx->pipeline);
}

// ---- multi-GPU bucket reduce-scatter -------------------------------------------------------------------------------------
// One process per GPU (or several contexts in one process: tests).  After b200msm_set_bases every rank exports a handle,
// the handles are exchanged by the caller (torch.distributed all_gather in sharded.py), every rank attaches, and from then on
// b200msm_msm_peer_partial_* replaces b200msm_msm_partial_*: the result is the sum over ALL GPUs' points of the buckets
// of this rank's slice, still one 224-byte record, combined exactly as before (all-gather + b200msm_combine_device).
struct peer_handle_t {
    uint32_t magic, kind;
    int32_t pid, device;
    uint32_t c, F, nbuckets, max_pieces, chunks, pad;
    uint64_t n;
    uint64_t raw_res, raw_poff, raw_flags;          // the same process maps nothing: plain pointers
    cudaIpcMemHandle_t res, poff, flags;
};
static_assert(sizeof(peer_handle_t) <= B200MSM_PEER_HANDLE_BYTES, "peer handle size");

int b200msm_peer_export(b200msm_ctx *ctx, int chunks, void *handle_out) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!handle_out || chunks < 1 || chunks > 8 || !ctx->plan_ok) return ctx->fail(-1, "b200msm_peer_export: bad argument (bases not set?)");
    if (ctx->F != 1) return ctx->fail(-1, "b200msm_peer_export: the bucket reduce-scatter needs a one-window plan (all windows folded)");
    if (ctx->peer.world) return ctx->fail(-1, "b200msm_peer_export: already attached; detach first");
    rc = ensure_chunks(ctx, chunks); if (rc) return rc;
    if (!ctx->peer.flags) {
        CK(dmalloc(&ctx->peer.flags, 2 * PEER_MAX));
        CK(cudaMemset(ctx->peer.flags, 0, 2 * PEER_MAX * 4));
    }
    ctx->peer.chunks = chunks;
    peer_handle_t h;
    memset(&h, 0, sizeof(h));
    h.magic = 0xb200feedu; h.kind = (uint32_t)ctx->kind; h.pid = (int32_t)getpid(); h.device = ctx->device;
    h.c = (uint32_t)ctx->c; h.F = (uint32_t)ctx->F; h.nbuckets = ctx->nbuckets; h.max_pieces = ctx->max_pieces; h.chunks = (uint32_t)chunks; h.n = ctx->n;
    h.raw_res = (uint64_t)ctx->res; h.raw_poff = (uint64_t)ctx->piece_off; h.raw_flags = (uint64_t)ctx->peer.flags;
    CK(cudaIpcGetMemHandle(&h.res, ctx->res));
    CK(cudaIpcGetMemHandle(&h.poff, ctx->piece_off));
    CK(cudaIpcGetMemHandle(&h.flags, ctx->peer.flags));
    memset(handle_out, 0, B200MSM_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int b200msm_peer_attach(b200msm_ctx *ctx, int rank, int world, const void *handles) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!handles || world < 2 || world > PEER_MAX || rank < 0 || rank >= world || !ctx->peer.flags) return ctx->fail(-1, "b200msm_peer_attach: bad argument (export first)");
    if ((world & (world - 1)) || ctx->nbuckets % (uint32_t)world || ctx->peer.chunks * world > RED_MAX_SRC) return ctx->fail(-1, "b200msm_peer_attach: world must be a power of two dividing the bucket count");
    auto &P = ctx->peer;
    const uint8_t *hb = (const uint8_t *)handles;
    red_src_t src[RED_MAX_SRC];
    uint32_t *fl[PEER_MAX];
    const size_t poff_stride = (size_t)ctx->nbuckets + 1;
    for (int k = 0; k < world; k++) {
        const int p = (rank + k) % world;              // own data first, then the peers in ring order: spreads the NVLink load
        peer_handle_t h;
        memcpy(&h, hb + (size_t)p * B200MSM_PEER_HANDLE_BYTES, sizeof(h));
        if (h.magic != 0xb200feedu || h.kind != (uint32_t)ctx->kind || h.c != (uint32_t)ctx->c || h.F != 1 || h.nbuckets != ctx->nbuckets || h.max_pieces != ctx->max_pieces ||
            h.chunks != (uint32_t)P.chunks || h.n != ctx->n)
            return ctx->fail(-1, "b200msm_peer_attach: rank " + std::to_string(p) + " exported a different plan (every rank must hold the same number of points)");
        const xyzz_t *res; const uint32_t *poff; uint32_t *flags;
        if (p == rank) { res = ctx->res; poff = ctx->piece_off; flags = P.flags; }
        else if (h.pid == (int32_t)getpid()) {         // another context of this process
            res = (const xyzz_t *)h.raw_res; poff = (const uint32_t *)h.raw_poff; flags = (uint32_t *)h.raw_flags;
            if (h.device != ctx->device) { cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError(); }
        } else {
            CK(cudaIpcOpenMemHandle(&P.mapped[p][0], h.res, cudaIpcMemLazyEnablePeerAccess));
            CK(cudaIpcOpenMemHandle(&P.mapped[p][1], h.poff, cudaIpcMemLazyEnablePeerAccess));
            CK(cudaIpcOpenMemHandle(&P.mapped[p][2], h.flags, cudaIpcMemLazyEnablePeerAccess));
            res = (const xyzz_t *)P.mapped[p][0]; poff = (const uint32_t *)P.mapped[p][1]; flags = (uint32_t *)P.mapped[p][2];
        }
        for (int ch = 0; ch < P.chunks; ch++) src[k * P.chunks + ch] = {res + (size_t)ch * ctx->max_pieces, poff + (size_t)ch * poff_stride};
        fl[p] = flags;
    }
    CK(dmalloc(&P.src_tab, (size_t)world * P.chunks));
    CK(cudaMalloc((void **)&P.flag_ptrs, sizeof(uint32_t *) * PEER_MAX));
    CK(cudaMemcpy(P.src_tab, src, sizeof(red_src_t) * world * P.chunks, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P.flag_ptrs, fl, sizeof(uint32_t *) * world, cudaMemcpyHostToDevice));
    P.rank = rank; P.world = world; P.epoch = 0;
    for (auto &e : P.ev) if (!e) CK(cudaEventCreate(&e));
    rc = make_red_plan(ctx, (uint32_t)world, (uint32_t)rank, &P.red_buf, P.steps, P.fin); if (rc) { P.world = 0; return rc; }
    CK(dmalloc(&P.merged, (size_t)ctx->nbuckets / world));
    P.nvlink_bytes = (uint64_t)(world - 1) * (ctx->nbuckets / (uint32_t)world) * (uint64_t)P.chunks * (sizeof(xyzz_t) + 8);
    return 0;
}

int b200msm_peer_detach(b200msm_ctx *ctx) {
    int rc = check_ctx(ctx); if (rc) return rc;
    CK(sync_streams(ctx));
    ctx->peer_release();
    return 0;
}

int b200msm_peer_info(b200msm_ctx *ctx, uint64_t info[6]) {
    if (!ctx || !info) return -1;
    info[0] = (uint64_t)ctx->peer.world; info[1] = (uint64_t)ctx->peer.rank; info[2] = (uint64_t)ctx->peer.chunks; info[3] = ctx->peer.nvlink_bytes;
    info[4] = info[5] = 0;
    if (ctx->peer.world && ctx->peer.epoch && sync_streams(ctx) == cudaSuccess) {
        float w = 0.f, m = 0.f;
        if (cudaEventElapsedTime(&w, ctx->peer.ev[0], ctx->peer.ev[1]) == cudaSuccess && cudaEventElapsedTime(&m, ctx->peer.ev[1], ctx->peer.ev[2]) == cudaSuccess) {
            info[4] = (uint64_t)(w * 1000.f); info[5] = (uint64_t)(m * 1000.f);
        }
        cudaGetLastError();
    }
    return 0;
}

// chunks of a peer call the caller did not use hold no pieces, so that the peers' level-0 walk finds nothing there
static int peer_clear_unused_chunks(b200msm_ctx *ctx, int used) {
    for (int ch = used; ch < ctx->peer.chunks; ch++)
        CK(cudaMemsetAsync(ctx->piece_off + (size_t)ch * ((size_t)ctx->nbuckets + 1), 0, ((size_t)ctx->nbuckets + 1) * 4, ctx->stream));
    return 0;
}
// stages 1-5 on this GPU's points, then `ready` to every peer
int b200msm_peer_accumulate_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_dev || ctx->n == 0 || n != ctx->n || !ctx->peer.world) return ctx->fail(-1, "b200msm_peer_accumulate_device: bad argument (not attached?)");
    ctx->peer.epoch++;
    rc = run_chunk(ctx, scalars_dev, 0, (uint32_t)ctx->n, 0); if (rc) return rc;
    rc = peer_clear_unused_chunks(ctx, 1); if (rc) return rc;
    k_peer_signal<<<1, PEER_MAX, 0, ctx->stream>>>(ctx->peer.flag_ptrs, (uint32_t)ctx->peer.world, (uint32_t)ctx->peer.rank, ctx->peer.epoch);
    CK(cudaGetLastError());
    return 0;
}
// this rank's slice of the buckets summed over every GPU (level 0 reads the peers over NVLink), reduced to one partial record
int b200msm_peer_reduce(b200msm_ctx *ctx, void *out_xyzz_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!out_xyzz_dev || !ctx->peer.world || ctx->peer.epoch == 0) return ctx->fail(-1, "b200msm_peer_reduce: bad argument");
    return run_reduce(ctx, ctx->peer.chunks, nullptr, out_xyzz_dev, true, !ctx->pipeline);
}
int b200msm_msm_peer_partial_device(b200msm_ctx *ctx, const void *scalars_dev, uint64_t n, void *out_xyzz_dev) {
    int rc = b200msm_peer_accumulate_device(ctx, scalars_dev, n); if (rc) return rc;
    return b200msm_peer_reduce(ctx, out_xyzz_dev);
}
int b200msm_msm_peer_partial_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, void *out_xyzz_dev) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_host || !out_xyzz_dev || ctx->n == 0 || n != ctx->n || !ctx->peer.world) return ctx->fail(-1, "b200msm_msm_peer_partial_host: bad argument (not attached?)");
    if (!ctx->scal_buf[0]) CK(cudaMalloc((void **)&ctx->scal_buf[0], (size_t)n * 32));
    CK(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    ctx->peer.epoch++;
    return upload_and_run(ctx, (const uint8_t *)scalars_host, n, ctx->scal_buf[0], ctx->ev_free[0], nullptr, out_xyzz_dev, true, !ctx->pipeline);
}

int b200msm_msm_host(b200msm_ctx *ctx, const void *scalars_host, uint64_t n, uint64_t batches, void *out_jacobian_host) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!scalars_host || !out_jacobian_host || ctx->n == 0 || n != ctx->n || batches == 0) return ctx->fail(-1, "b200msm_msm_host: bad argument (bases not set or length mismatch)");
    const size_t bytes = (size_t)n * 32;
    for (int i = 0; i < 2; i++) if (!ctx->scal_buf[i] && (i == 0 || batches > 1)) CK(cudaMalloc((void **)&ctx->scal_buf[i], bytes));
    if (ctx->out_dev_cap < batches) {
        if (ctx->out_dev) cudaFree(ctx->out_dev);
        CK(cudaMalloc((void **)&ctx->out_dev, (size_t)batches * 144));
        ctx->out_dev_cap = batches;
    }
    // the caller's stream may hold earlier work on these buffers
    CK(cudaEventRecord(ctx->ev_free[0], ctx->stream));
    CK(cudaEventRecord(ctx->ev_free[1], ctx->stream));
    // The upload of batch b+1 overlaps the compute of batch b; the first batch is chunked (upload_and_run).
    for (uint64_t b = 0; b < batches; b++) {
        int k = (int)(b & 1);
        if (b == 0) {
            rc = upload_and_run(ctx, (const uint8_t *)scalars_host, n, ctx->scal_buf[k], ctx->ev_free[k], ctx->out_dev, nullptr, false, false);
            if (rc) return rc;
            CK(cudaEventRecord(ctx->ev_free[k], ctx->stream));
            continue;
        }
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[k], 0));
        rc = h2d(ctx, ctx->scal_buf[k], (const uint8_t *)scalars_host + b * bytes, bytes, ctx->copy_stream); if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_copied[k], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[k], 0));
        rc = run_msm(ctx, ctx->scal_buf[k], ctx->out_dev + b * 144, nullptr, false);      // its tail runs beside the next batch's planning
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_free[k], ctx->stream));
    }
    rc = join_tail(ctx); if (rc) return rc;
    CK(cudaMemcpyAsync(out_jacobian_host, ctx->out_dev, (size_t)batches * 144, cudaMemcpyDeviceToHost, ctx->stream));
    CK(sync_streams(ctx));
    return 0;
}

int b200msm_sync(b200msm_ctx *ctx) {
    int rc = check_ctx(ctx); if (rc) return rc;
    CK(sync_streams(ctx));
    return 0;
}

// The window plan alone (pure host arithmetic, no device needed): what a context created with (window_bits, tables) would
// choose for n points with free_bytes of device memory, the reduction summing over `chunks` chunks.
int b200msm_plan_info(int curve, uint64_t n, uint64_t free_bytes, int window_bits, int tables, int chunks, uint64_t info[8]) {
    if ((curve != 0 && curve != 1) || n == 0 || n >= (1ull << 31) || !info || window_bits < 0 || window_bits > 26 || (window_bits > 0 && window_bits < 2) ||
        tables < 0 || chunks < 1 || chunks > 8) return -1;
    b200msm_ctx tmp;
    tmp.curve = curve; tmp.req_c = window_bits; tmp.req_tables = tables; tmp.plan_chunks = chunks;
    tmp.kind = curve;
    if (const char *g = getenv("B200MSM_GROUP")) if (curve == 0 && strcmp(g, "xyzz") == 0) tmp.kind = 2;
    BY_GROUP(tmp.kind, tmp.entry_bytes = LaunchCold<C>::entry_bytes());
    if (!choose_plan(&tmp, n, (size_t)free_bytes)) return -1;
    info[0] = n; info[1] = tmp.c; info[2] = tmp.W; info[3] = tmp.T; info[4] = tmp.F; info[5] = tmp.nbuckets;
    info[6] = (uint64_t)tmp.T * n * tmp.entry_bytes;      // bytes of base tables
    info[7] = (uint64_t)tmp.W * n;                        // point references per MSM
    return 0;
}

int b200msm_get_info(const b200msm_ctx *ctx, uint64_t info[8]) {
    if (!ctx || !info) return -1;
    info[0] = ctx->n; info[1] = ctx->c; info[2] = ctx->W; info[3] = ctx->T; info[4] = ctx->F; info[5] = ctx->nbuckets; info[6] = ctx->lmax; info[7] = ctx->launches;
    return 0;
}

// Sums the stage times of the newest `want` complete MSMs still in the ring (every chunk's set counted): walks the sets from
// the newest back; a set with ev_last == 6 closes an MSM, the sets before it down to the previous closing set are its chunks.
static int stage_sums(b200msm_ctx *ctx, uint64_t want, float sum[6], uint64_t *covered) {
    CK(sync_streams(ctx));
    for (int i = 0; i < 6; i++) sum[i] = 0.f;
    *covered = 0;
    const uint64_t have = ctx->ev_sets < (uint64_t)b200msm_ctx::EV_RING ? ctx->ev_sets : (uint64_t)b200msm_ctx::EV_RING;
    float pend[6] = {};
    bool open = false;                    // pend holds the sets of an MSM whose first chunk has not been reached yet
    for (uint64_t k = 0; k < have && *covered < want; k++) {
        const int slot = (int)((ctx->ev_sets - 1 - k) % b200msm_ctx::EV_RING);
        const int last = ctx->ev_last[slot];
        if (last == 6) {                  // a closing set: whatever was pending is complete
            if (open) { for (int i = 0; i < 6; i++) sum[i] += pend[i]; (*covered)++; if (*covered == want) break; }
            for (int i = 0; i < 6; i++) pend[i] = 0.f;
            open = true;
        } else if (!open) continue;       // chunks of an MSM that has not been reduced yet
        for (int i = 0; i < last; i++) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->evr[slot][i], ctx->evr[slot][i + 1])); pend[i] += ms; }
    }
    if (open && *covered < want && ctx->ev_sets <= (uint64_t)b200msm_ctx::EV_RING) { for (int i = 0; i < 6; i++) sum[i] += pend[i]; (*covered)++; }
    return 0;
}

int b200msm_get_stage_ms(b200msm_ctx *ctx, float stage_ms[6]) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!stage_ms) return -1;
    uint64_t covered = 0;
    return stage_sums(ctx, 1, stage_ms, &covered);
}

int b200msm_reset_stage_stats(b200msm_ctx *ctx) {
    int rc = check_ctx(ctx); if (rc) return rc;
    CK(sync_streams(ctx));
    ctx->ev_steps = 0;
    ctx->ev_sets = 0;
    return 0;
}

int b200msm_get_stage_ms_sum(b200msm_ctx *ctx, float stage_ms_sum[6], uint64_t *steps) {
    int rc = check_ctx(ctx); if (rc) return rc;
    if (!stage_ms_sum || !steps) return -1;
    return stage_sums(ctx, ctx->ev_steps, stage_ms_sum, steps);
}

// ---- one-call form: arkworks VariableBaseMSM::multi_scalar_mul(bases, scalars) ---------------------------------
// The literal shape of ark-ec 0.3 multi_scalar_mul (P1A/speakspeak/src/util.rs:51-58: size = min(bases.len(), scalars.len()),
// empty -> zero), of upstream sppark's one-shot mult_pippenger (team-division/prize2-posw-gpu/storswiftlabs/snarkVM/
// algorithms/cuda/pippenger.cu:19-25, bound by .../Jackytan2018/extern/sppark/sppark/poc/blst-cuda/src/lib.rs:21-53) and of
// snarkVM's VariableBase::msm -> cuda::msm_cuda dispatch (.../snarkVM/algorithms/src/msm/variable_base/mod.rs:120-157).
// Bases change from call to call, so the plan is ONE table (no 2^(cj) P precompute: building T tables costs ~23 (T - 1)
// doublings per point, far more than the MSM it would save) and the context of a (curve, device) pair is kept between
// calls (as sppark keeps a static pippenger_t per device), so repeated calls of one size re-use every device buffer.
static const uint64_t R_MOD_P[2][6] = {     // 2^384 mod p: the Montgomery one of the arkworks zero (1, 1, 0)
    {0x02cdffffffffff68ull, 0x51409f837fffffb1ull, 0x9f7db3a98a7d3ff2ull, 0x7b4e97b76e7c6305ull, 0x4cf495bf803c84e8ull, 0x008d6661e2fdf49aull},
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull}};
static const int ONESHOT_MAX_DEV = 16;
static std::mutex g_oneshot_mu;
static b200msm_ctx *g_oneshot[2][ONESHOT_MAX_DEV] = {};

}  // extern "C"

// Bases and scalars of one call stream through the device in steps: while step k's records are converted (arkworks ->
// 28-bit limbs -> table entry: one shared inversion per GRP_BATCH points on BLS12-377) and while the chunk before
// accumulates, the copy stream already uploads step k + 1.  The accumulate chunks are the ones of upload_and_run (own
// bucket sums per chunk, one reduction over all of them), equal-sized here because the call is upload-bound, not
// compute-bound: the exposed tail is one chunk's accumulation, not one chunk's upload.
static int one_call_chunks(const b200msm_ctx *ctx, uint64_t n) {
    int nch = ctx->host_chunks > 0 ? ctx->host_chunks : (n >= (1ull << 23) ? 4 : (n >= (1ull << 21) ? 2 : 1));
    if (nch > 8) nch = 8;
    if ((uint64_t)nch * 32 > n) nch = 1;
    return nch;
}
template <class C> static int one_call_t(b200msm_ctx *ctx, const uint8_t *bases, uint64_t n, uint64_t stride, const uint8_t *scalars, void *out_host) {
    if (ctx->T != 1) return ctx->fail(-1, "one-call form needs a one-table plan");
    const uint64_t step = 1ull << 22;                          // 4 M records (416 MB) per upload: 65536 converting threads
    const auto t_call = std::chrono::steady_clock::now();
    ctx->h2d_ms = ctx->h2d_wait_ms = 0;
    const uint64_t cap = n < step ? n : step;
    if (ctx->raw_cap < cap * stride) {
        CK(sync_streams(ctx));
        for (int i = 0; i < 2; i++) { cudaFree(ctx->raw_buf[i]); ctx->raw_buf[i] = nullptr; }
        ctx->raw_cap = 0;
        for (int i = 0; i < 2; i++) CK(cudaMalloc((void **)&ctx->raw_buf[i], (size_t)(cap * stride)));
        ctx->raw_cap = (size_t)(cap * stride);
    }
    if (!ctx->sw_buf) CK(dmalloc(&ctx->sw_buf, (size_t)cap));
    if (!ctx->scal_buf[0]) CK(cudaMalloc((void **)&ctx->scal_buf[0], (size_t)n * 32));
    if (ctx->out_dev_cap < 1) { CK(cudaMalloc((void **)&ctx->out_dev, 144)); ctx->out_dev_cap = 1; }
    const int nch = one_call_chunks(ctx, n);
    int rc = 0;
    if (nch > 1) { rc = ensure_chunks(ctx, nch); if (rc) return rc; }
    CK(cudaMemsetAsync(ctx->infmask, 0, ((size_t)cdiv(n, 32) + 1) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev_free[0], ctx->stream));         // earlier work on this stream may still read the buffers
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
    // The conversion of chunk k + 1 runs on its own high-priority stream beside the accumulation of chunk k: its threads are
    // latency-bound (one serial inversion chain per 64 points), the accumulate kernel is pipe-bound, together they fill the SMs.
    const char *conv_env = getenv("B200MSM_CONV_STREAM");
    const bool conv_async = !(conv_env && atoi(conv_env) == 0);
    cudaStream_t cs = conv_async ? ctx->conv_stream : ctx->stream;
    if (conv_async) CK(cudaStreamWaitEvent(cs, ctx->ev_free[0], 0));
    uint64_t k = 0;
    int used = 0;
    for (int ch = 0; ch < nch; ch++) {
        // chunk boundaries on multiples of 32: the conversion writes whole words of the infinity mask
        const uint64_t first = (n * ch / nch) & ~31ull, end = ch + 1 == nch ? n : (n * (ch + 1) / nch) & ~31ull;
        if (end <= first) continue;
        for (uint64_t s = first; s < end; s += step, k++) {
            const uint64_t cnt = end - s < step ? end - s : step;
            const int b = (int)(k & 1);
            if (k >= 2) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_raw_free[b], 0));
            rc = h2d(ctx, ctx->raw_buf[b], bases + s * stride, (size_t)(cnt * stride), ctx->copy_stream); if (rc) return rc;
            CK(cudaEventRecord(ctx->ev_raw_copied[b], ctx->copy_stream));
            CK(cudaStreamWaitEvent(cs, ctx->ev_raw_copied[b], 0));
            LaunchCold<C>::convert_points(cs, ctx->raw_buf[b], (uint32_t)cnt, (uint32_t)stride, ctx->sw_buf, ctx->infmask + s / 32);
            LaunchCold<C>::make_entries(cs, ctx->sw_buf, (uint32_t)cnt, ctx->tables + s * ctx->entry_bytes, ctx->acc_mix == 0, ctx->infmask + s / 32, ctx->bad);
            CK(cudaEventRecord(ctx->ev_raw_free[b], cs));
        }
        if (conv_async) { CK(cudaEventRecord(ctx->ev_conv, cs)); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_conv, 0)); }
        rc = h2d(ctx, ctx->scal_buf[0] + first * 32, scalars + first * 32, (size_t)(end - first) * 32, ctx->copy_stream); if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_chunk[used], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[used], 0));
        rc = run_chunk(ctx, ctx->scal_buf[0] + first * 32, (uint32_t)first, (uint32_t)(end - first), used);
        if (rc) return rc;
        used++;
    }
    rc = run_reduce(ctx, used, ctx->out_dev, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_host, ctx->out_dev, 144, cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    const double issue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
    CK(sync_streams(ctx));
    if (getenv("B200MSM_TRACE")) {
        float st[6] = {};
        uint64_t cov = 0;
        stage_sums(ctx, 1, st, &cov);                                                               // summed over the call's chunks
        fprintf(stderr, "b200msm one-call n=%llu c=%d W=%d F=%d chunks=%d: total %.1f ms, issued after %.1f ms, in h2d %.1f ms (waiting for a slot %.1f ms); "
                "all chunks: digits %.1f scatter %.1f pieces %.1f accumulate %.1f, reduce %.1f finalize %.1f ms\n", (unsigned long long)n, ctx->c, ctx->W, ctx->F, nch,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(), issue_ms, ctx->h2d_ms, ctx->h2d_wait_ms,
                st[0], st[1], st[2], st[3], st[4], st[5]);
    }
    if (bad) {
        CK(cudaMemsetAsync(ctx->bad, 0, 4, ctx->stream));
        return ctx->fail(-3, std::to_string(bad) + " bases lie outside the prime-order subgroup and have no twisted-Edwards image; set B200MSM_GROUP=xyzz");
    }
    return 0;
}

extern "C" {

int b200msm_variable_base_msm(int curve, int device, const void *affine_host, uint64_t nbases, uint64_t stride,
                              const void *scalars_host, uint64_t nscalars, void *out_jacobian_host) {
    if ((curve != 0 && curve != 1) || device < 0 || device >= ONESHOT_MAX_DEV || !out_jacobian_host || stride < 96 || (stride & 7)) {
        g_create_error = "b200msm_variable_base_msm: bad argument";
        return -1;
    }
    const uint64_t n = nbases < nscalars ? nbases : nscalars;
    if (n >= (1ull << 31) || (n > 0 && (!affine_host || !scalars_host))) { g_create_error = "b200msm_variable_base_msm: bad input"; return -1; }
    std::lock_guard<std::mutex> lock(g_oneshot_mu);
    b200msm_ctx *&ctx = g_oneshot[curve][device];
    if (!ctx) {                                       // also the no-device failure: the empty MSM has no CPU path either
        int rc = b200msm_create(&ctx, curve, device, 0, 1);
        if (rc) { ctx = nullptr; return rc; }
    }
    if (n == 0) {
        uint64_t *o = (uint64_t *)out_jacobian_host;
        for (int i = 0; i < 6; i++) { o[i] = R_MOD_P[curve][i]; o[6 + i] = R_MOD_P[curve][i]; o[12 + i] = 0; }
        return 0;
    }
    int rc = check_ctx(ctx);
    ctx->plan_chunks = one_call_chunks(ctx, n);       // the window plan knows the reduction sums over this many chunks
    if (rc == 0) rc = alloc_plan(ctx, n);
    if (rc == 0) BY_GROUP(ctx->kind, rc = one_call_t<C>(ctx, (const uint8_t *)affine_host, n, stride, (const uint8_t *)scalars_host, out_jacobian_host));
    if (rc) {                                         // keep the message, drop the context: the next call starts clean
        g_create_error = ctx->err;
        b200msm_destroy(ctx);
        ctx = nullptr;
    }
    return rc;
}

int b200msm_variable_base_release(void) {
    std::lock_guard<std::mutex> lock(g_oneshot_mu);
    for (auto &row : g_oneshot) for (b200msm_ctx *&c : row) if (c) { b200msm_destroy(c); c = nullptr; }
    return 0;
}

// ---- element-wise test entry points ----------------------------------------------------------
static int test_fail(const char *what, cudaError_t e) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
    fprintf(stderr, "b200msm: %s\n", g_create_error.c_str());
    return (int)e;
}
#define TCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return test_fail(#call, e__); } while (0)

int b200msm_test_fp_op(int curve, int op, const void *a, const void *b, void *r, uint64_t n) {
    if (!a || !b || !r || n == 0 || (curve != 0 && curve != 1)) return -1;
    fp_abi *da, *db, *dr;
    size_t bytes = (size_t)n * 48;
    TCK(cudaMalloc((void **)&da, bytes)); TCK(cudaMalloc((void **)&db, bytes)); TCK(cudaMalloc((void **)&dr, bytes));
    TCK(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice)); TCK(cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice));
    if (curve == 0) LaunchCold<Bls12_377>::test_fp(op, da, db, dr, (uint32_t)n);
    else LaunchCold<Bls12_381>::test_fp(op, da, db, dr, (uint32_t)n);
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(r, dr, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dr);
    return 0;
}

int b200msm_test_ec_op(int curve, int op, void *acc, const void *other, uint64_t n, int reps) {
    if (!acc || n == 0 || (curve != 0 && curve != 1) || op < 0 || op > 3 || (op != 2 && !other)) return -1;
    xyzz_abi_t *dacc; void *doth = nullptr;
    size_t ob = (op == 1 ? 192 : 96) * (size_t)n;
    TCK(cudaMalloc((void **)&dacc, (size_t)n * 192));
    TCK(cudaMemcpy(dacc, acc, (size_t)n * 192, cudaMemcpyHostToDevice));
    if (op != 2) { TCK(cudaMalloc(&doth, ob)); TCK(cudaMemcpy(doth, other, ob, cudaMemcpyHostToDevice)); }
    if (curve == 0) LaunchCold<Bls12_377>::test_ec(op, dacc, doth, (uint32_t)n, reps);
    else LaunchCold<Bls12_381>::test_ec(op, dacc, doth, (uint32_t)n, reps);
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(acc, dacc, (size_t)n * 192, cudaMemcpyDeviceToHost));
    cudaFree(dacc); if (doth) cudaFree(doth);
    return 0;
}

int b200msm_test_grp_op(int curve, int op, const void *a96, const void *b96, void *out104, uint64_t n, int reps) {
    if (!a96 || !b96 || !out104 || n == 0 || curve < 0 || curve > 2 || op < 0 || op > 3 || reps < 0) return -1;      // 2: BLS12-377 in XYZZ
    affine_abi_t *da, *db; uint32_t *dout;
    TCK(cudaMalloc((void **)&da, (size_t)n * 96)); TCK(cudaMalloc((void **)&db, (size_t)n * 96)); TCK(cudaMalloc((void **)&dout, (size_t)n * 104));
    TCK(cudaMemcpy(da, a96, (size_t)n * 96, cudaMemcpyHostToDevice)); TCK(cudaMemcpy(db, b96, (size_t)n * 96, cudaMemcpyHostToDevice));
    BY_GROUP(curve, LaunchCold<C>::test_grp(op, da, db, dout, (uint32_t)n, reps));
    TCK(cudaGetLastError());
    TCK(cudaMemcpy(out104, dout, (size_t)n * 104, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}

// ---- Family A (sppark) -----------------------------------------------------------------------
static RustError rust_error(int code, const char *msg) {
    RustError e;
    e.code = code;
    e.message = (code != 0 && msg) ? strdup(msg) : nullptr;
    return e;
}

RustError mult_pippenger_init(MultiScalarMultContext *context, const void *points_with_infinity, size_t npoints, size_t ffi_affine_sz) {
    if (!context) return rust_error(-1, "mult_pippenger_init: null context");
    context->context = nullptr;
    b200msm_ctx *ctx = nullptr;
    int rc = b200msm_create(&ctx, B200MSM_BLS12_377, 0, 0, 0);
    if (rc) return rust_error(rc, b200msm_last_error(nullptr));
    rc = b200msm_set_bases(ctx, points_with_infinity, npoints, ffi_affine_sz);
    if (rc) { RustError e = rust_error(rc, b200msm_last_error(ctx)); b200msm_destroy(ctx); return e; }
    context->context = ctx;
    return rust_error(0, nullptr);
}

RustError mult_pippenger_inf(MultiScalarMultContext *context, uint64_t *out, const void *points_with_infinity, size_t npoints, size_t batch_size,
                             const void *scalars, size_t ffi_affine_sz) {
    (void)points_with_infinity; (void)ffi_affine_sz;    // bases were taken at init, exactly as in the reference
    if (!context || !context->context) return rust_error(-1, "mult_pippenger_inf: context not initialised");
    b200msm_ctx *ctx = (b200msm_ctx *)context->context;
    int rc = b200msm_msm_host(ctx, scalars, npoints, batch_size, out);
    if (rc) return rust_error(rc, b200msm_last_error(ctx));
    return rust_error(0, nullptr);
}

RustError mult_pippenger_free(MultiScalarMultContext *context) {
    if (!context || !context->context) return rust_error(-1, "mult_pippenger_free: context not initialised");
    b200msm_destroy((b200msm_ctx *)context->context);
    context->context = nullptr;
    return rust_error(0, nullptr);
}

// upstream sppark's one-shot form: blst_p1_affine points (96 bytes: x, y, no infinity flag), BLS12-377 in the snarkVM build
RustError mult_pippenger(void *out, const void *points, size_t npoints, const void *scalars) {
    int rc = b200msm_variable_base_msm(B200MSM_BLS12_377, 0, points, npoints, 96, scalars, npoints, out);
    if (rc) return rust_error(rc, b200msm_last_error(nullptr));
    return rust_error(0, nullptr);
}

// ---- Family B (yrrid) ------------------------------------------------------------------------
struct yrrid_ctx_t { b200msm_ctx *ctx; int32_t max_points, max_batches; uint32_t points; };

void *MSMAllocContext(int32_t maxPoints, int32_t maxBatches) {
    if (maxPoints <= 0 || maxBatches <= 0) return nullptr;
    b200msm_ctx *ctx = nullptr;
    if (b200msm_create(&ctx, B200MSM_BLS12_377, 0, 0, 0) != 0) return nullptr;
    yrrid_ctx_t *y = new yrrid_ctx_t{ctx, maxPoints, maxBatches, 0};
    return y;
}
int32_t MSMFreeContext(void *context) {
    if (!context) return -1;
    yrrid_ctx_t *y = (yrrid_ctx_t *)context;
    b200msm_destroy(y->ctx);
    delete y;
    return 0;
}
int32_t MSMPreprocessPoints(void *context, void *affinePointsPtr, uint32_t points) {
    if (!context) return -1;
    yrrid_ctx_t *y = (yrrid_ctx_t *)context;
    if (points == 0 || (int64_t)points > (int64_t)y->max_points) return -1;
    int rc = b200msm_set_bases(y->ctx, affinePointsPtr, points, 104);
    if (rc == 0) y->points = points;
    return rc;
}
int32_t MSMRun(void *context, uint64_t *projectiveResultsPtr, void *scalarsPtr, uint32_t scalars) {
    if (!context) return -1;
    yrrid_ctx_t *y = (yrrid_ctx_t *)context;
    if (y->points == 0 || scalars == 0 || scalars % y->points != 0) return -1;
    uint32_t batches = scalars / y->points;
    if ((int64_t)batches > (int64_t)y->max_batches) return -1;
    return b200msm_msm_host(y->ctx, scalarsPtr, y->points, batches, projectiveResultsPtr);
}

}  // extern "C"
