// grp.cuh -- the bucket coordinate system of a curve, as the kernels see it.  Grp<C> is
//   GrpTe<C>    twisted-Edwards extended coordinates, 7 M unified mixed add, 144-byte packed table entries (BLS12-377)
//   GrpXyzz<C>  short-Weierstrass XYZZ, 8 M + 2 S mixed add with infinity / doubling cases, 112-byte entries (BLS12-381)
// A bucket is four field elements in both (struct xyzz_t).  Base tables are built in short-Weierstrass affine form
// (doubling chains + batched normalisation, kernels_cold.cuh) and converted to entries chunk by chunk.
#pragma once
#include "te.cuh"

namespace b200msm {

#define GRP_BATCH 64      // points per thread in the entry conversion (one inversion per batch)

template <class C> struct GrpXyzz {
    typedef Fp<C> F;
    typedef affine_t entry_t;
    static constexpr int ENTRY_VEC = sizeof(affine_t) / 16;       // 7
    typedef xyzz_t acc_t;                                         // the accumulate kernel's register accumulator
    static constexpr bool INF_ENTRIES = true;                     // a table entry (0, 0) stands for the point at infinity
    FP_HD static void set_inf(xyzz_t &a) { Ec<C>::set_inf(a); }
#if defined(__CUDACC__)
    __device__ __forceinline__ static void acc_init(acc_t &a) { Ec<C>::set_inf(a); }
    __device__ __forceinline__ static void acc_madd_staged(acc_t &acc, const uint4 *st, int stride, bool negate) { madd_staged(acc, st, stride, negate); }
    __device__ __forceinline__ static void acc_export(xyzz_t &out, const acc_t &a) { out = a; }
    // acc += entry, the entry staged as ENTRY_VEC 16-byte chunks `stride` uint4 apart
    __device__ __forceinline__ static void madd_staged(xyzz_t &acc, const uint4 *st, int stride, bool negate) {
        affine_t pt;
        uint4 *t = reinterpret_cast<uint4 *>(&pt);
#pragma unroll
        for (int k = 0; k < ENTRY_VEC; k++) t[k] = st[k * stride];
        // (0, 0): 2^(c F j) P at infinity for a base P of small even order (outside G1); lies on neither curve
        if (F::is_literal_zero(pt.x) && F::is_literal_zero(pt.y)) return;
        if (negate) F::template neg<2>(pt.y, pt.y);
        Ec<C>::madd(acc, pt.x, pt.y);
    }
    __device__ __noinline__ static void add_ni(xyzz_t &a, const xyzz_t &b) { Ec<C>::add(a, b); }
    __device__ __noinline__ static void dbl_ni(xyzz_t &a) { Ec<C>::dbl(a); }
    // acc += the table entry at e (global memory); cold kernels only
    __device__ __noinline__ static void madd_entry_ni(xyzz_t &acc, const entry_t *e, bool /*ark*/) { madd_staged(acc, reinterpret_cast<const uint4 *>(e), 1, false); }
    __device__ static bool is_neutral(const xyzz_t &a) { return Ec<C>::is_inf(a); }
    __device__ static bool to_sw_affine(fp &x, fp &y, const xyzz_t &a) {
        if (Ec<C>::is_inf(a)) return false;
        Ec<C>::to_affine_ni(x, y, a);
        return true;
    }
    // thread t converts src[t*GRP_BATCH ...] (short-Weierstrass affine, canonical) to entries: a copy here
    __device__ static void make_entries(const affine_t *__restrict__ src, uint32_t n, entry_t *__restrict__ dst, uint32_t t, bool /*ark*/,
                                        const uint32_t * /*infmask*/, uint32_t * /*bad*/) {
        for (uint32_t k = 0; k < GRP_BATCH; k++) {
            const uint32_t i = t * GRP_BATCH + k;
            if (i >= n) return;
            const uint4 *s = reinterpret_cast<const uint4 *>(src + i);
            uint4 *d = reinterpret_cast<uint4 *>(dst + i);
#pragma unroll
            for (int v = 0; v < ENTRY_VEC; v++) d[v] = s[v];
        }
    }
    __device__ static void entry_to_sw(fp &x, fp &y, const entry_t &e, bool /*ark*/) { x = e.x; y = e.y; }
#endif
};

template <class C> struct GrpTe {
    typedef Fp<C> F;
    typedef te_entry_t entry_t;
    static constexpr int ENTRY_VEC = sizeof(te_entry_t) / 16;     // 9
    typedef ted_t acc_t;                                          // FP64-pipe accumulator (te.cuh TeD)
    static constexpr bool INF_ENTRIES = false;                    // every entry is a finite point (unmappable bases are refused)
    FP_HD static void set_inf(xyzz_t &a) { Te<C>::set_inf(a); }
#if defined(__CUDACC__)
    __device__ __forceinline__ static void acc_init(acc_t &a) { TeD<C>::set_inf(a); }
    __device__ __forceinline__ static void acc_madd_staged(acc_t &acc, const uint4 *st, int stride, bool negate) {
        te_entry_t e;
        uint4 *t = reinterpret_cast<uint4 *>(&e);
#pragma unroll
        for (int k = 0; k < ENTRY_VEC; k++) t[k] = st[k * stride];
        ted_pre_t q;
        TeD<C>::unpack(q, e);
        if (negate) TeD<C>::negate(q);
        TeD<C>::madd(acc, q);
    }
    __device__ __noinline__ static void acc_export(xyzz_t &out, const acc_t &a) { TeD<C>::to_int(out, a); }
    // acc (cold representation) += staged entry THROUGH THE HOT ARITHMETIC (fp29.cuh, Te29::madd): the group-law test kernel
    // exercises exactly what k_accumulate runs; conversions in and out cost 8 multiplies, irrelevant there
    __device__ __noinline__ static void madd_staged(xyzz_t &acc, const uint4 *st, int stride, bool negate) {
        te29_t a;
        Te29::import_from(a, acc);
        fp29 v[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            fp_abi t;
            const uint4 v0 = st[(3 * k) * stride], v1 = st[(3 * k + 1) * stride], v2 = st[(3 * k + 2) * stride];
            t.w[0] = v0.x; t.w[1] = v0.y; t.w[2] = v0.z; t.w[3] = v0.w; t.w[4] = v1.x; t.w[5] = v1.y; t.w[6] = v1.z; t.w[7] = v1.w;
            t.w[8] = v2.x; t.w[9] = v2.y; t.w[10] = v2.z; t.w[11] = v2.w;
            Fp29::repack_from_abi(v[k], t);
        }
        Te29::madd(a, v[0], v[1], v[2], negate);
        Te29::export_to(acc, a);
    }
    __device__ __noinline__ static void add_ni(xyzz_t &a, const xyzz_t &b) { Te<C>::add(a, b); }
    __device__ __noinline__ static void dbl_ni(xyzz_t &a) { Te<C>::dbl(a); }
    __device__ __noinline__ static void madd_entry_ni(xyzz_t &acc, const entry_t *e, bool ark) {
        te_pre_t q;
        Te<C>::unpack(q, *e, ark);
        Te<C>::madd(acc, q);
    }
    // (0 : 1 : 1) exactly: X = 0 also holds for the point of order 2, (0 : -1 : 1)
    __device__ static bool is_neutral(const xyzz_t &a) {
        fp t;
        F::template sub<3>(t, a.y, a.zz);
        return F::is_zero(a.x) && F::is_zero(t);
    }
    __device__ __noinline__ static bool to_sw_affine(fp &x, fp &y, const xyzz_t &a) {
        if (Te<C>::is_inf(a)) return false;
        fp zmy, zpy, den, inv, t, one, c3, cy;
        F::template sub<3>(zmy, a.zz, a.y);
        F::add(zpy, a.zz, a.y);
        Ec<C>::mul_ni(den, a.x, zmy);
        Ec<C>::inv_ni(inv, den);
        Ec<C>::mul_ni(t, zpy, inv);
        Te<C>::load_const(c3, 3);
        Te<C>::load_const(cy, 4);
        F::set_one(one);
        Ec<C>::mul_ni(x, t, a.x);
        Ec<C>::mul_ni(x, x, c3);
        F::template sub<2>(x, x, one);
        Ec<C>::mul_ni(y, t, a.zz);
        Ec<C>::mul_ni(y, y, cy);
        Ec<C>::mul_ni(x, x, one);                 // canonical representatives: x * R / R, then < p
        Ec<C>::mul_ni(y, y, one);
        F::csub_p(x);
        F::csub_p(y);
        return true;
    }
    // Montgomery's trick over GRP_BATCH points: one inversion per thread.  A zero denominator (y = 0 or
    // s (x + 1) = -1: only for points outside G1, or the (0, 0) placeholder of a base flagged infinity, which the digit
    // kernel never references) is replaced by 1 so that it cannot poison its batch; unless the base is flagged infinity
    // (infmask: bit i = point i of src) it is COUNTED in *bad -- such a point has no image in this model and the host
    // refuses the base set (msm.cu check_bad_entries) instead of returning a wrong sum.
    __device__ static void make_entries(const affine_t *__restrict__ src, uint32_t n, entry_t *__restrict__ dst, uint32_t t, bool ark,
                                        const uint32_t *__restrict__ infmask, uint32_t *__restrict__ bad) {
        const uint32_t base = t * GRP_BATCH;
        if (base >= n) return;
        const uint32_t cnt = n - base < GRP_BATCH ? n - base : GRP_BATCH;
        fp prefix[GRP_BATCH];
        fp acc;
        F::set_one(acc);
        for (uint32_t k = 0; k < cnt; k++) {
            affine_t a;
            load_affine_plain(a, src + base + k);
            fp nu, nv, dv, den;
            Te<C>::map_prepare(nu, nv, dv, den, a.x, a.y);
            if (F::is_zero(den)) {
                F::set_one(den);
                const uint32_t i = base + k;
                if (bad && !(infmask && (infmask[i >> 5] >> (i & 31)) & 1)) atomicAdd(bad, 1u);
            }
            prefix[k] = acc;
            Ec<C>::mul_ni(acc, acc, den);
        }
        fp inv;
        Ec<C>::inv_ni(inv, acc);
        for (uint32_t k = cnt; k-- > 0;) {
            affine_t a;
            load_affine_plain(a, src + base + k);
            fp nu, nv, dv, den, i1;
            Te<C>::map_prepare(nu, nv, dv, den, a.x, a.y);       // recomputed: cheaper than keeping 4 x GRP_BATCH values
            if (F::is_zero(den)) F::set_one(den);
            Ec<C>::mul_ni(i1, inv, prefix[k]);
            Ec<C>::mul_ni(inv, inv, den);
            te_entry_t e;
            Te<C>::map_finish(e, nu, nv, dv, a.y, i1, ark);
            uint4 *d = reinterpret_cast<uint4 *>(dst + base + k);
            const uint4 *s = reinterpret_cast<const uint4 *>(&e);
#pragma unroll
            for (int v = 0; v < ENTRY_VEC; v++) d[v] = s[v];
        }
    }
    __device__ static void entry_to_sw(fp &x, fp &y, const entry_t &e, bool ark) {
        xyzz_t a;
        Te<C>::entry_to_point(a, e, ark);
        if (!to_sw_affine(x, y, a)) { F::set_zero(x); F::set_zero(y); }
    }
    __device__ static void load_affine_plain(affine_t &d, const affine_t *src) {
        const uint4 *s = reinterpret_cast<const uint4 *>(src);
        uint4 *t = reinterpret_cast<uint4 *>(&d);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(affine_t) / 16); i++) t[i] = s[i];
    }
#endif
};

#if defined(__CUDACC__)
// accumulator flavours of the accumulate kernel: which registers hold the running sum and which pipe multiplies
template <class G> struct AccNative {          // the group's own integer arithmetic
    typedef xyzz_t acc_t;
    __device__ __forceinline__ static void init(acc_t &a) { G::set_inf(a); }
    __device__ __forceinline__ static void madd_staged(acc_t &acc, const uint4 *st, int stride, bool negate) { G::madd_staged(acc, st, stride, negate); }
    __device__ __forceinline__ static void export_to(xyzz_t &out, const acc_t &a) { out = a; }
};
template <class C> struct AccTe29 {            // twisted Edwards in the 13 x 29-bit field (fp29.cuh): the default on BLS12-377
    typedef te29_t acc_t;
    __device__ __forceinline__ static void init(acc_t &a) { Te29::set_inf(a); }
    __device__ __forceinline__ static void unpack3(fp29 &r, const uint4 *st, int stride) {
        fp_abi t;
        const uint4 v0 = st[0], v1 = st[stride], v2 = st[2 * stride];
        t.w[0] = v0.x; t.w[1] = v0.y; t.w[2] = v0.z; t.w[3] = v0.w; t.w[4] = v1.x; t.w[5] = v1.y; t.w[6] = v1.z; t.w[7] = v1.w;
        t.w[8] = v2.x; t.w[9] = v2.y; t.w[10] = v2.z; t.w[11] = v2.w;
        Fp29::repack_from_abi(r, t);
    }
    __device__ __forceinline__ static void madd_staged(acc_t &acc, const uint4 *st, int stride, bool negate) {
        fp29 ymx, ypx, kt;
        unpack3(ymx, st, stride); unpack3(ypx, st + 3 * stride, stride); unpack3(kt, st + 6 * stride, stride);
        Te29::madd(acc, ymx, ypx, kt, negate);
    }
    __device__ __forceinline__ static void export_to(xyzz_t &out, const acc_t &a) { Te29::export_to(out, a); }
};
template <class C> struct AccFp64Te {          // twisted Edwards on the FP64 pipe
    typedef GrpTe<C> G;
    typedef typename G::acc_t acc_t;
    __device__ __forceinline__ static void init(acc_t &a) { G::acc_init(a); }
    __device__ __forceinline__ static void madd_staged(acc_t &acc, const uint4 *st, int stride, bool negate) { G::acc_madd_staged(acc, st, stride, negate); }
    __device__ __forceinline__ static void export_to(xyzz_t &out, const acc_t &a) { G::acc_export(out, a); }
};
#endif

template <class C, bool TE = C::HAS_TE> struct GrpSel { typedef GrpXyzz<C> type; };
template <class C> struct GrpSel<C, true> { typedef GrpTe<C> type; };
template <class C> using Grp = typename GrpSel<C>::type;

}  // namespace b200msm
