// Self-test kernels behind bbs_selftest_* (include/bbs_b200.h): they expose the field, G1 and pairing
// layers one level below the BBS operations so the parity tests can pin each layer against the oracle.
#pragma once
#include "kernels.cuh"

namespace bbs {

struct FieldTestArgs { int op; const uint8_t* a; const uint8_t* b; uint8_t* out; };

template <class C> BBS_HD void field_test_item(const FieldTestArgs& t, uint32_t i) {
    using Fp = typename C::Fp;
    using Fr = typename C::Fr;
    if (t.op <= 4 || t.op == 7) {
        constexpr int N = Fp::N;
        BBS_A16 uint32_t a[N], b[N], r[N];
        limbs_from_le<N>(a, t.a + (size_t)i * 4 * N);
        limbs_from_le<N>(b, t.b + (size_t)i * 4 * N);
        fe_to_mont<Fp>(a, a); fe_to_mont<Fp>(b, b);
        if (t.op == 0) fe_mul<Fp>(r, a, b);
        else if (t.op == 1) fe_add<Fp>(r, a, b);
        else if (t.op == 2) fe_sub<Fp>(r, a, b);
        else if (t.op == 3) fe_inv<Fp>(r, a);
        else if (t.op == 7) fe_inv_vt<Fp>(r, a);
        else { if (!fe_sqrt<Fp>(r, a)) bn_zero<N>(r); }
        fe_from_mont<Fp>(r, r);
        limbs_to_le<N>(t.out + (size_t)i * 4 * N, r);
    } else {
        BBS_A16 uint32_t a[8], b[8], r[8];
        limbs_from_le<8>(a, t.a + (size_t)i * 32);
        limbs_from_le<8>(b, t.b + (size_t)i * 32);
        fe_to_mont<Fr>(a, a); fe_to_mont<Fr>(b, b);
        if (t.op == 5) fe_mul<Fr>(r, a, b); else if (t.op == 6) fe_inv<Fr>(r, a); else fe_inv_vt<Fr>(r, a);
        fe_from_mont<Fr>(r, r);
        limbs_to_le<8>(t.out + (size_t)i * 32, r);
    }
}


struct G1MulTestArgs { const uint8_t* pts; const uint8_t* sc; uint8_t* out; };
template <class C> BBS_HD void g1_mul_test_item(const G1MulTestArgs& t, uint32_t i) {
    BBS_A16 uint32_t A[G1A], k[8], R[G1J];
    uint8_t* o = t.out + (size_t)i * C::G1_BYTES;
    int st = g1_decompress<C>(A, t.pts + (size_t)i * C::G1_BYTES);
    limbs_from_le<8>(k, t.sc + (size_t)i * 32);
    if (st == PT_BAD) { for (int j = 0; j < C::G1_BYTES; j++) o[j] = 0xff; return; }
    // canonical scalars take the production path (windowed GLV, g1_mul_scalar); others the plain double-and-add
    if (st == PT_INF) g1_set_inf<C>(R);
    else if (fe_is_canonical<typename C::Fr>(k)) g1_mul_scalar<C>(R, A, k);
    else g1_mul_affine<C>(R, A, k, 256);
    g1_compress<C>(o, R);
}

struct PairTestPrepArgs { const uint8_t* q; uint32_t* W; uint32_t* lines; uint32_t* st; };
template <class C> BBS_HD void pair_test_prep_item(const PairTestPrepArgs& t, uint32_t i) {
    if (i == 0) {
        int s = g2_decompress<C>(t.W, t.q);
        t.st[0] = (uint32_t)s;
        if (s == PT_OK) g2_precompute_lines<C>(t.lines, t.W, 0);
    } else {
        g2_precompute_lines<C>(t.lines, C::G2(), 1);
    }
}
struct PairTestArgs { const uint8_t* p; const uint8_t* r; const uint32_t* lines; const uint32_t* st; uint8_t* status; };
template <class C> BBS_HD void pair_test_item(const PairTestArgs& t, uint32_t i) {
    BBS_A16 uint32_t P[G1A], R[G1A], a0[3 * FPN], a1[3 * FPN], f[F12N];
    int sp = g1_decompress<C>(P, t.p + (size_t)i * C::G1_BYTES);
    int sr = g1_decompress<C>(R, t.r + (size_t)i * C::G1_BYTES);
    if (sp == PT_BAD || sr == PT_BAD || t.st[0] == PT_BAD) { t.status[i] = ST_ERR_MALFORMED; return; }
    bn_copy<2 * C::Fp::N>(a0, P); fe_set_one<typename C::Fp>(a0 + 2 * FPN);
    bn_copy<2 * C::Fp::N>(a1, R); fe_set_one<typename C::Fp>(a1 + 2 * FPN);
    miller2<C>(f, t.lines, a0, sp == PT_INF || t.st[0] == PT_INF, a1, sr == PT_INF);
    t.status[i] = final_exp_is_one<C>(f) ? ST_ACCEPT : ST_REJECT;
}

}  // namespace bbs
