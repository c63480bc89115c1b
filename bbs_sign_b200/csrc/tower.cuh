// Fp2 / Fp6 / Fp12 tower on top of field.cuh, for both curves:
//   Fp2 = Fp[u]/(u^2+1),  Fp6 = Fp2[v]/(v^3 - xi),  Fp12 = Fp6[w]/(w^2 - v)
//   xi = 1+u (BLS12-381), 9+u (BN254).
// Replaces ark-ff's Fp2/Fp6/Fp12 under `E::pairing` (verify.rs:88-92, proof_verify.rs:112-115).
// Elements are flat uint32_t arrays (Montgomery form): Fp2 = [c0|c1], Fp6 = [c0|c1|c2] of Fp2,
// Fp12 = [c0|c1] of Fp6.  GT values are never observable in the reference (only `== ONE`), so the
// representation is free; SURVEY 8a note (ii).
#pragma once
#include "field.cuh"

namespace bbs {

// ---- curve parameter packs -----------------------------------------------------------------------
struct Bls {
    using Fp = BlsFp;
    using Fr = BlsFr;
    static constexpr int ID = 1;
    static constexpr bool M_TWIST = true;     // line is sparse at (0,1,4); BN (D-type) at (0,3,4)
    static constexpr int G1_BYTES = 48;       // compressed
    static constexpr int G2_BYTES = 96;
    static BBS_HD const uint32_t* B() { return BLS_B(); }
    static BBS_HD const uint32_t* P1() { return BLS_P1(); }
    static BBS_HD const uint32_t* G2() { return BLS_G2(); }
    static BBS_HD const uint32_t* GLV_BETA() { return BLS_GLV_BETA(); }
    static BBS_HD const uint32_t* FROB(int j) { return j == 1 ? BLS_FROB1() : (j == 2 ? BLS_FROB2() : BLS_FROB3()); }
    // r = a * (1+u)
    static BBS_HD void mul_xi(uint32_t* r, const uint32_t* a) {
        BBS_A16 uint32_t t0[12], t1[12];
        fe_sub<Fp>(t0, a, a + 12);
        fe_add<Fp>(t1, a, a + 12);
        bn_copy<12>(r, t0);
        bn_copy<12>(r + 12, t1);
    }
};

struct Bn {
    using Fp = BnFp;
    using Fr = BnFr;
    static constexpr int ID = 2;
    static constexpr bool M_TWIST = false;
    static constexpr int G1_BYTES = 32;
    static constexpr int G2_BYTES = 64;
    static BBS_HD const uint32_t* B() { return BN_B(); }
    static BBS_HD const uint32_t* P1() { return BN_P1(); }
    static BBS_HD const uint32_t* G2() { return BN_G2(); }
    static BBS_HD const uint32_t* GLV_BETA() { return BN_GLV_BETA(); }
    static BBS_HD const uint32_t* FROB(int j) { return j == 1 ? BN_FROB1() : (j == 2 ? BN_FROB2() : BN_FROB3()); }
    // r = a * (9+u) = (9a0 - a1) + (9a1 + a0) u
    static BBS_HD void mul_xi(uint32_t* r, const uint32_t* a) {
        BBS_A16 uint32_t t0[8], t1[8], x[8];
        fe_dbl<Fp>(x, a); fe_dbl<Fp>(x, x); fe_dbl<Fp>(x, x); fe_add<Fp>(x, x, a);          // 9 a0
        fe_sub<Fp>(t0, x, a + 8);
        fe_dbl<Fp>(x, a + 8); fe_dbl<Fp>(x, x); fe_dbl<Fp>(x, x); fe_add<Fp>(x, x, a + 8);  // 9 a1
        fe_add<Fp>(t1, x, a);
        bn_copy<8>(r, t0);
        bn_copy<8>(r + 8, t1);
    }
};

// ---- Fp2 -----------------------------------------------------------------------------------------
#define FPN (C::Fp::N)
#define F2N (2 * C::Fp::N)
#define F6N (6 * C::Fp::N)
#define F12N (12 * C::Fp::N)

template <class C> BBS_HD void f2_copy(uint32_t* r, const uint32_t* a) { bn_copy<2 * C::Fp::N>(r, a); }
template <class C> BBS_HD void f2_zero(uint32_t* r) { bn_zero<2 * C::Fp::N>(r); }
template <class C> BBS_HD void f2_one(uint32_t* r) { fe_set_one<typename C::Fp>(r); bn_zero<C::Fp::N>(r + FPN); }
template <class C> BBS_HD bool f2_is_zero(const uint32_t* a) { return bn_is_zero<2 * C::Fp::N>(a); }
template <class C> BBS_HD bool f2_eq(const uint32_t* a, const uint32_t* b) { return bn_eq<2 * C::Fp::N>(a, b); }
template <class C> BBS_HD void f2_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    fe_add<typename C::Fp>(r, a, b); fe_add<typename C::Fp>(r + FPN, a + FPN, b + FPN);
}
template <class C> BBS_HD void f2_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    fe_sub<typename C::Fp>(r, a, b); fe_sub<typename C::Fp>(r + FPN, a + FPN, b + FPN);
}
template <class C> BBS_HD void f2_dbl(uint32_t* r, const uint32_t* a) { f2_add<C>(r, a, a); }
template <class C> BBS_HD void f2_neg(uint32_t* r, const uint32_t* a) {
    fe_neg<typename C::Fp>(r, a); fe_neg<typename C::Fp>(r + FPN, a + FPN);
}
template <class C> BBS_HD void f2_conj(uint32_t* r, const uint32_t* a) {
    bn_copy<C::Fp::N>(r, a); fe_neg<typename C::Fp>(r + FPN, a + FPN);
}
// Karatsuba, 3 base-field products
template <class C> BBS_HDN void f2_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    using F = typename C::Fp;
    BBS_A16 uint32_t t0[FPN], t1[FPN], s0[FPN], s1[FPN], t2[FPN];
    fe_mul<F>(t0, a, b);
    fe_mul<F>(t1, a + FPN, b + FPN);
    fe_add<F>(s0, a, a + FPN);
    fe_add<F>(s1, b, b + FPN);
    fe_mul<F>(t2, s0, s1);
    fe_sub<F>(r, t0, t1);
    fe_sub<F>(t2, t2, t0);
    fe_sub<F>(r + FPN, t2, t1);
}
// (a0+a1)(a0-a1), 2 a0 a1
template <class C> BBS_HDN void f2_sqr(uint32_t* r, const uint32_t* a) {
    using F = typename C::Fp;
    BBS_A16 uint32_t s[FPN], d[FPN], m[FPN];
    fe_add<F>(s, a, a + FPN);
    fe_sub<F>(d, a, a + FPN);
    fe_mul<F>(m, a, a + FPN);
    fe_mul<F>(r, s, d);
    fe_dbl<F>(r + FPN, m);
}
template <class C> BBS_HD void f2_mul_fp(uint32_t* r, const uint32_t* a, const uint32_t* k) {
    fe_mul<typename C::Fp>(r, a, k); fe_mul<typename C::Fp>(r + FPN, a + FPN, k);
}
template <class C> BBS_HDN void f2_inv(uint32_t* r, const uint32_t* a) {
    using F = typename C::Fp;
    BBS_A16 uint32_t n[FPN], t[FPN];
    fe_sqr<F>(n, a);
    fe_sqr<F>(t, a + FPN);
    fe_add<F>(n, n, t);
    fe_inv<F>(n, n);
    fe_mul<F>(r, a, n);
    fe_mul<F>(t, a + FPN, n);
    fe_neg<F>(r + FPN, t);
}

// variable-time variant for context creation (one thread, public data): field.cuh fe_inv_vt
template <class C> BBS_HDN void f2_inv_vt(uint32_t* r, const uint32_t* a) {
    using F = typename C::Fp;
    BBS_A16 uint32_t n[FPN], t[FPN];
    fe_sqr<F>(n, a);
    fe_sqr<F>(t, a + FPN);
    fe_add<F>(n, n, t);
    fe_inv_vt<F>(n, n);
    fe_mul<F>(r, a, n);
    fe_mul<F>(t, a + FPN, n);
    fe_neg<F>(r + FPN, t);
}

// ---- Fp6 -----------------------------------------------------------------------------------------
template <class C> BBS_HD void f6_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    for (int i = 0; i < 3; i++) f2_add<C>(r + i * F2N, a + i * F2N, b + i * F2N);
}
template <class C> BBS_HD void f6_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    for (int i = 0; i < 3; i++) f2_sub<C>(r + i * F2N, a + i * F2N, b + i * F2N);
}
template <class C> BBS_HD void f6_neg(uint32_t* r, const uint32_t* a) {
    for (int i = 0; i < 3; i++) f2_neg<C>(r + i * F2N, a + i * F2N);
}
template <class C> BBS_HD void f6_copy(uint32_t* r, const uint32_t* a) { bn_copy<6 * C::Fp::N>(r, a); }
// r = a * v  (in place safe)
template <class C> BBS_HD void f6_mul_by_v(uint32_t* r, const uint32_t* a) {
    BBS_A16 uint32_t t[F2N], a0[F2N], a1[F2N];
    C::mul_xi(t, a + 2 * F2N);
    f2_copy<C>(a0, a);
    f2_copy<C>(a1, a + F2N);
    f2_copy<C>(r, t);
    f2_copy<C>(r + F2N, a0);
    f2_copy<C>(r + 2 * F2N, a1);
}
// Karatsuba: 6 Fp2 products
template <class C> BBS_HDN void f6_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    const uint32_t *a0 = a, *a1 = a + F2N, *a2 = a + 2 * F2N, *b0 = b, *b1 = b + F2N, *b2 = b + 2 * F2N;
    BBS_A16 uint32_t v0[F2N], v1[F2N], v2[F2N], s[F2N], t[F2N], c0[F2N], c1[F2N], c2[F2N];
    f2_mul<C>(v0, a0, b0);
    f2_mul<C>(v1, a1, b1);
    f2_mul<C>(v2, a2, b2);
    // c0 = v0 + xi((a1+a2)(b1+b2) - v1 - v2)
    f2_add<C>(s, a1, a2); f2_add<C>(t, b1, b2); f2_mul<C>(c0, s, t);
    f2_sub<C>(c0, c0, v1); f2_sub<C>(c0, c0, v2); C::mul_xi(c0, c0); f2_add<C>(c0, c0, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1 + xi v2
    f2_add<C>(s, a0, a1); f2_add<C>(t, b0, b1); f2_mul<C>(c1, s, t);
    f2_sub<C>(c1, c1, v0); f2_sub<C>(c1, c1, v1); C::mul_xi(t, v2); f2_add<C>(c1, c1, t);
    // c2 = (a0+a2)(b0+b2) - v0 - v2 + v1
    f2_add<C>(s, a0, a2); f2_add<C>(t, b0, b2); f2_mul<C>(c2, s, t);
    f2_sub<C>(c2, c2, v0); f2_sub<C>(c2, c2, v2); f2_add<C>(c2, c2, v1);
    f2_copy<C>(r, c0); f2_copy<C>(r + F2N, c1); f2_copy<C>(r + 2 * F2N, c2);
}
// a * (b0 + b1 v): 5 Fp2 products
template <class C> BBS_HDN void f6_mul_by_01(uint32_t* r, const uint32_t* a, const uint32_t* b0, const uint32_t* b1) {
    const uint32_t *a0 = a, *a1 = a + F2N, *a2 = a + 2 * F2N;
    BBS_A16 uint32_t aa[F2N], bb[F2N], s[F2N], t[F2N], c0[F2N], c1[F2N], c2[F2N];
    f2_mul<C>(aa, a0, b0);
    f2_mul<C>(bb, a1, b1);
    f2_add<C>(s, a1, a2); f2_mul<C>(c0, s, b1); f2_sub<C>(c0, c0, bb); C::mul_xi(c0, c0); f2_add<C>(c0, c0, aa);
    f2_add<C>(s, a0, a2); f2_mul<C>(c2, s, b0); f2_sub<C>(c2, c2, aa); f2_add<C>(c2, c2, bb);
    f2_add<C>(s, a0, a1); f2_add<C>(t, b0, b1); f2_mul<C>(c1, s, t); f2_sub<C>(c1, c1, aa); f2_sub<C>(c1, c1, bb);
    f2_copy<C>(r, c0); f2_copy<C>(r + F2N, c1); f2_copy<C>(r + 2 * F2N, c2);
}
// a * (b1 v): 3 Fp2 products
template <class C> BBS_HD void f6_mul_by_1(uint32_t* r, const uint32_t* a, const uint32_t* b1) {
    BBS_A16 uint32_t c0[F2N], c1[F2N], c2[F2N];
    f2_mul<C>(c0, a + 2 * F2N, b1); C::mul_xi(c0, c0);
    f2_mul<C>(c1, a, b1);
    f2_mul<C>(c2, a + F2N, b1);
    f2_copy<C>(r, c0); f2_copy<C>(r + F2N, c1); f2_copy<C>(r + 2 * F2N, c2);
}
template <class C> BBS_HDN void f6_inv(uint32_t* r, const uint32_t* a) {
    const uint32_t *a0 = a, *a1 = a + F2N, *a2 = a + 2 * F2N;
    BBS_A16 uint32_t c0[F2N], c1[F2N], c2[F2N], t[F2N], d[F2N];
    // c0 = a0^2 - xi a1 a2 ; c1 = xi a2^2 - a0 a1 ; c2 = a1^2 - a0 a2
    f2_sqr<C>(c0, a0); f2_mul<C>(t, a1, a2); C::mul_xi(t, t); f2_sub<C>(c0, c0, t);
    f2_sqr<C>(c1, a2); C::mul_xi(c1, c1); f2_mul<C>(t, a0, a1); f2_sub<C>(c1, c1, t);
    f2_sqr<C>(c2, a1); f2_mul<C>(t, a0, a2); f2_sub<C>(c2, c2, t);
    // d = a0 c0 + xi (a2 c1 + a1 c2)
    f2_mul<C>(d, a2, c1); f2_mul<C>(t, a1, c2); f2_add<C>(d, d, t); C::mul_xi(d, d);
    f2_mul<C>(t, a0, c0); f2_add<C>(d, d, t);
    f2_inv<C>(d, d);
    f2_mul<C>(r, c0, d); f2_mul<C>(r + F2N, c1, d); f2_mul<C>(r + 2 * F2N, c2, d);
}

// ---- Fp12 ----------------------------------------------------------------------------------------
template <class C> BBS_HD void f12_copy(uint32_t* r, const uint32_t* a) { bn_copy<12 * C::Fp::N>(r, a); }
template <class C> BBS_HD void f12_one(uint32_t* r) { bn_zero<12 * C::Fp::N>(r); fe_set_one<typename C::Fp>(r); }
template <class C> BBS_HD bool f12_is_one(const uint32_t* a) {
    BBS_A16 uint32_t one[FPN];
    fe_set_one<typename C::Fp>(one);
    uint32_t o = 0;
    for (int i = 0; i < FPN; i++) o |= a[i] ^ one[i];
    for (int i = FPN; i < F12N; i++) o |= a[i];
    return o == 0;
}
template <class C> BBS_HD void f12_conj(uint32_t* r, const uint32_t* a) {
    f6_copy<C>(r, a); f6_neg<C>(r + F6N, a + F6N);
}
template <class C> BBS_HDN void f12_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    BBS_A16 uint32_t aa[F6N], bb[F6N], s[F6N], t[F6N];
    f6_mul<C>(aa, a, b);
    f6_mul<C>(bb, a + F6N, b + F6N);
    f6_add<C>(s, a, a + F6N);
    f6_add<C>(t, b, b + F6N);
    f6_mul<C>(s, s, t);
    f6_sub<C>(s, s, aa);
    f6_sub<C>(r + F6N, s, bb);
    f6_mul_by_v<C>(bb, bb);
    f6_add<C>(r, aa, bb);
}
// complex squaring: 2 Fp6 products
template <class C> BBS_HDN void f12_sqr(uint32_t* r, const uint32_t* a) {
    BBS_A16 uint32_t ab[F6N], s[F6N], t[F6N];
    f6_mul<C>(ab, a, a + F6N);
    f6_add<C>(s, a, a + F6N);
    f6_mul_by_v<C>(t, a + F6N);
    f6_add<C>(t, t, a);
    f6_mul<C>(s, s, t);          // (a0+a1)(a0+v a1) = a0^2 + v a1^2 + (1+v) a0 a1
    f6_sub<C>(s, s, ab);
    f6_mul_by_v<C>(t, ab);
    f6_sub<C>(r, s, t);
    f6_add<C>(r + F6N, ab, ab);
}
template <class C> BBS_HDN void f12_inv(uint32_t* r, const uint32_t* a) {
    BBS_A16 uint32_t t0[F6N], t1[F6N];
    f6_mul<C>(t0, a, a);
    f6_mul<C>(t1, a + F6N, a + F6N);
    f6_mul_by_v<C>(t1, t1);
    f6_sub<C>(t0, t0, t1);       // a0^2 - v a1^2
    f6_inv<C>(t0, t0);
    f6_mul<C>(r, a, t0);
    f6_mul<C>(t1, a + F6N, t0);
    f6_neg<C>(r + F6N, t1);
}
// slot s of the tower layout [c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2] carries w-power {0,2,4,1,3,5}[s]
template <class C> BBS_HDN void f12_frob(uint32_t* r, const uint32_t* a, int j) {
    const uint32_t* g = C::FROB(j);
    const int wpow[6] = {0, 2, 4, 1, 3, 5};
    for (int s = 0; s < 6; s++) {
        BBS_A16 uint32_t t[F2N];
        if (j & 1) f2_conj<C>(t, a + s * F2N); else f2_copy<C>(t, a + s * F2N);
        f2_mul<C>(r + s * F2N, t, g + wpow[s] * F2N);
    }
}
// f *= (c0 + c1 v) + (c4 v) w      [M-type twist line; BLS12-381]
template <class C> BBS_HDN void f12_mul_by_014(uint32_t* f, const uint32_t* c0, const uint32_t* c1, const uint32_t* c4) {
    BBS_A16 uint32_t aa[F6N], bb[F6N], s[F6N], o[F2N];
    f6_mul_by_01<C>(aa, f, c0, c1);
    f6_mul_by_1<C>(bb, f + F6N, c4);
    f2_add<C>(o, c1, c4);
    f6_add<C>(s, f, f + F6N);
    f6_mul_by_01<C>(s, s, c0, o);
    f6_sub<C>(s, s, aa);
    f6_sub<C>(f + F6N, s, bb);
    f6_mul_by_v<C>(bb, bb);
    f6_add<C>(f, aa, bb);
}
// f *= c0 + (c3 + c4 v) w          [D-type twist line; BN254]
template <class C> BBS_HDN void f12_mul_by_034(uint32_t* f, const uint32_t* c0, const uint32_t* c3, const uint32_t* c4) {
    BBS_A16 uint32_t a[F6N], b[F6N], e[F6N], o[F2N];
    for (int i = 0; i < 3; i++) f2_mul<C>(a + i * F2N, f + i * F2N, c0);
    f6_mul_by_01<C>(b, f + F6N, c3, c4);
    f2_add<C>(o, c0, c3);
    f6_add<C>(e, f, f + F6N);
    f6_mul_by_01<C>(e, e, o, c4);
    f6_sub<C>(e, e, a);
    f6_sub<C>(f + F6N, e, b);
    f6_mul_by_v<C>(b, b);
    f6_add<C>(f, a, b);
}
// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part)
template <class C> BBS_HDN void f12_cyc_sqr(uint32_t* r, const uint32_t* a) {
    // z0=c0.c0 z4=c0.c1 z3=c0.c2 z2=c1.c0 z1=c1.c1 z5=c1.c2
    const uint32_t *z0 = a, *z4 = a + F2N, *z3 = a + 2 * F2N, *z2 = a + 3 * F2N, *z1 = a + 4 * F2N, *z5 = a + 5 * F2N;
    BBS_A16 uint32_t t0[F2N], t1[F2N], t2[F2N], t3[F2N], t4[F2N], t5[F2N], tmp[F2N], s[F2N], u[F2N];
    // (x + y Y)^2 in Fp4 = Fp2[Y]/(Y^2 - xi):  (x^2 + xi y^2) + 2xy Y
#define BBS_FP4_SQR(x, y, lo, hi)                                                     \
    f2_mul<C>(tmp, x, y);                                                             \
    f2_add<C>(s, x, y); C::mul_xi(u, y); f2_add<C>(u, u, x); f2_mul<C>(lo, s, u);     \
    f2_sub<C>(lo, lo, tmp); C::mul_xi(u, tmp); f2_sub<C>(lo, lo, u);                  \
    f2_dbl<C>(hi, tmp);
    BBS_FP4_SQR(z0, z1, t0, t1)
    BBS_FP4_SQR(z2, z3, t2, t3)
    BBS_FP4_SQR(z4, z5, t4, t5)
#undef BBS_FP4_SQR
    BBS_A16 uint32_t o[F12N];
    // z0' = 3 t0 - 2 z0 ; z1' = 3 t1 + 2 z1
    f2_sub<C>(s, t0, z0); f2_dbl<C>(s, s); f2_add<C>(o, s, t0);
    f2_add<C>(s, t1, z1); f2_dbl<C>(s, s); f2_add<C>(o + 4 * F2N, s, t1);
    // z2' = 3 xi t5 + 2 z2 ; z3' = 3 t4 - 2 z3
    C::mul_xi(tmp, t5);
    f2_add<C>(s, tmp, z2); f2_dbl<C>(s, s); f2_add<C>(o + 3 * F2N, s, tmp);
    f2_sub<C>(s, t4, z3); f2_dbl<C>(s, s); f2_add<C>(o + 2 * F2N, s, t4);
    // z4' = 3 t2 - 2 z4 ; z5' = 3 t3 + 2 z5
    f2_sub<C>(s, t2, z4); f2_dbl<C>(s, s); f2_add<C>(o + F2N, s, t2);
    f2_add<C>(s, t3, z5); f2_dbl<C>(s, s); f2_add<C>(o + 5 * F2N, s, t3);
    f12_copy<C>(r, o);
}

}  // namespace bbs
