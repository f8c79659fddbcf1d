// Host build of the device field/curve headers with the PTX carry flag emulated (TEST ONLY).
// Lets the CPU suite check the limb-level algorithms of csrc/fp.cuh and csrc/ec.cuh against the oracle
// without a GPU.  Built by tests/test_host_emul.py with g++; never part of the product library.
#include <cstring>
#include <cstddef>
#include "../../2022-entries_b200/csrc/ec.cuh"
using namespace b200msm;

template <class C> static void fp_ops(int op, const fp *a, const fp *b, fp *r, size_t n) {
    typedef Fp<C> F;
    for (size_t i = 0; i < n; i++) {
        fp t;
        switch (op) {
        case 0: F::mul(t, a[i], b[i]); break;
        case 1: F::add(t, a[i], b[i]); break;
        case 2: F::template sub<2>(t, a[i], b[i]); break;
        case 3: F::sqr(t, a[i]); break;
        case 4: F::inv(t, a[i]); break;
        case 5: F::to_mont(t, a[i]); break;
        case 6: F::from_mont(t, a[i]); break;
        default: t = a[i];
        }
        F::reduce(t);
        r[i] = t;
    }
}
template <class C> static void canon(xyzz_t &q) {
    typedef Fp<C> F;
    F::reduce(q.x); F::reduce(q.y); F::reduce(q.zz); F::reduce(q.zzz);
}
// op 0: acc += aff (madd); 1: acc += other (add); 2: acc = 2 acc; 3: acc = 2*aff
template <class C> static void ec_ops(int op, xyzz_t *acc, const void *other, size_t n, int reps) {
    typedef Ec<C> E;
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < reps; k++) {
            switch (op) {
            case 0: E::madd(acc[i], ((const affine_t *)other)[i].x, ((const affine_t *)other)[i].y); break;
            case 1: E::add(acc[i], ((const xyzz_t *)other)[i]); break;
            case 2: E::dbl(acc[i]); break;
            case 3: E::dbl_affine(acc[i], ((const affine_t *)other)[i].x, ((const affine_t *)other)[i].y); break;
            }
        }
        canon<C>(acc[i]);
    }
}
// lazy-bound stress: multiply k1*p + a by k2*p + b (same residues, larger representatives)
template <class C> static void fp_mul_lazy(const fp *a, const fp *b, fp *r, size_t n, int k1, int k2) {
    typedef Fp<C> F;
    for (size_t i = 0; i < n; i++) {
        fp x = a[i], y = b[i], pp;
        for (int j = 0; j < 12; j++) pp.l[j] = C::p(j);
        for (int k = 0; k < k1; k++) F::add(x, x, pp);
        for (int k = 0; k < k2; k++) F::add(y, y, pp);
        fp t; F::mul(t, x, y);
        // record whether the documented bound (< 2p) held: subtract 2p must borrow
        fp chk = t; F::template csub<2>(chk);
        r[i] = t;
        F::reduce(r[i]);
        if (memcmp(&chk, &t, sizeof(fp)) != 0) r[i].l[11] |= 0x80000000u;   // flag bound violation
    }
}
extern "C" {
int emul_fp_op(int curve, int op, const void *a, const void *b, void *r, size_t n) {
    if (curve == 0) fp_ops<Bls12_377>(op, (const fp *)a, (const fp *)b, (fp *)r, n); else fp_ops<Bls12_381>(op, (const fp *)a, (const fp *)b, (fp *)r, n);
    return 0;
}
int emul_ec_op(int curve, int op, void *acc, const void *other, size_t n, int reps) {
    if (curve == 0) ec_ops<Bls12_377>(op, (xyzz_t *)acc, other, n, reps); else ec_ops<Bls12_381>(op, (xyzz_t *)acc, other, n, reps);
    return 0;
}
int emul_fp_mul_lazy(int curve, const void *a, const void *b, void *r, size_t n, int k1, int k2) {
    if (curve == 0) fp_mul_lazy<Bls12_377>((const fp *)a, (const fp *)b, (fp *)r, n, k1, k2); else fp_mul_lazy<Bls12_381>((const fp *)a, (const fp *)b, (fp *)r, n, k1, k2);
    return 0;
}
}
