// G1 arithmetic (short Weierstrass, a = 0) in Jacobian coordinates over Fp, plus the two wire encodings.
// Replaces ark-ec `Projective: Add / Mul<Fr>` and ark-serialize `serialize_compressed` on G1 under
// verify.rs:81-86, sign.rs:120-130, proof_verify.rs:163-182, proof_gen.rs:304-311.
// Point layouts (Montgomery limbs): affine = [x|y] (2N words), Jacobian = [X|Y|Z] (3N), Z = 0 <=> identity.
#pragma once
#include "tower.cuh"

namespace bbs {

#define G1A (2 * C::Fp::N)
#define G1J (3 * C::Fp::N)

template <class C> BBS_HD bool g1_is_inf(const uint32_t* p) { return bn_is_zero<C::Fp::N>(p + 2 * FPN); }
template <class C> BBS_HD void g1_set_inf(uint32_t* p) {
    fe_set_one<typename C::Fp>(p); fe_set_one<typename C::Fp>(p + FPN); bn_zero<C::Fp::N>(p + 2 * FPN);
}
template <class C> BBS_HD void g1_copy(uint32_t* r, const uint32_t* p) { bn_copy<3 * C::Fp::N>(r, p); }
template <class C> BBS_HD void g1_from_affine(uint32_t* r, const uint32_t* a) {
    bn_copy<2 * C::Fp::N>(r, a); fe_set_one<typename C::Fp>(r + 2 * FPN);
}
template <class C> BBS_HD void g1_neg(uint32_t* r, const uint32_t* p) {
    bn_copy<C::Fp::N>(r, p); fe_neg<typename C::Fp>(r + FPN, p + FPN); bn_copy<C::Fp::N>(r + 2 * FPN, p + 2 * FPN);
}

// dbl-2009-l: 2M + 5S
template <class C> BBS_HDN void g1_dbl(uint32_t* r, const uint32_t* p) {
    using F = typename C::Fp;
    const uint32_t *X = p, *Y = p + FPN, *Z = p + 2 * FPN;
    uint32_t A[FPN], B[FPN], Cc[FPN], D[FPN], E[FPN], Fq[FPN], t[FPN], Z3[FPN];
    fe_sqr<F>(A, X);
    fe_sqr<F>(B, Y);
    fe_sqr<F>(Cc, B);
    fe_add<F>(t, X, B); fe_sqr<F>(t, t); fe_sub<F>(t, t, A); fe_sub<F>(t, t, Cc); fe_dbl<F>(D, t);
    fe_dbl<F>(E, A); fe_add<F>(E, E, A);
    fe_sqr<F>(Fq, E);
    fe_mul<F>(Z3, Y, Z); fe_dbl<F>(Z3, Z3);
    fe_dbl<F>(t, D); fe_sub<F>(r, Fq, t);                       // X3 = F - 2D
    fe_sub<F>(t, D, r); fe_mul<F>(t, E, t);
    fe_dbl<F>(Cc, Cc); fe_dbl<F>(Cc, Cc); fe_dbl<F>(Cc, Cc);    // 8C
    fe_sub<F>(r + FPN, t, Cc);
    bn_copy<C::Fp::N>(r + 2 * FPN, Z3);                         // identity stays identity (Z3 = 0)
}

// madd-2007-bl (Jacobian += affine): 7M + 4S, complete via the rare-case branches
template <class C> BBS_HDN void g1_add_mixed(uint32_t* r, const uint32_t* p, const uint32_t* q /*affine*/) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { g1_from_affine<C>(r, q); return; }
    const uint32_t *X1 = p, *Y1 = p + FPN, *Z1 = p + 2 * FPN, *X2 = q, *Y2 = q + FPN;
    uint32_t Z1Z1[FPN], U2[FPN], S2[FPN], H[FPN], HH[FPN], I[FPN], J[FPN], rr[FPN], V[FPN], t[FPN], X3[FPN], Y3[FPN], Z3[FPN];
    fe_sqr<F>(Z1Z1, Z1);
    fe_mul<F>(U2, X2, Z1Z1);
    fe_mul<F>(S2, Y2, Z1); fe_mul<F>(S2, S2, Z1Z1);
    fe_sub<F>(H, U2, X1);
    fe_sub<F>(rr, S2, Y1);
    if (bn_is_zero<C::Fp::N>(H)) {
        if (bn_is_zero<C::Fp::N>(rr)) { g1_dbl<C>(r, p); } else { g1_set_inf<C>(r); }
        return;
    }
    fe_dbl<F>(rr, rr);
    fe_sqr<F>(HH, H);
    fe_dbl<F>(I, HH); fe_dbl<F>(I, I);
    fe_mul<F>(J, H, I);
    fe_mul<F>(V, X1, I);
    fe_sqr<F>(X3, rr); fe_sub<F>(X3, X3, J); fe_sub<F>(X3, X3, V); fe_sub<F>(X3, X3, V);
    fe_sub<F>(t, V, X3); fe_mul<F>(Y3, rr, t);
    fe_mul<F>(t, Y1, J); fe_dbl<F>(t, t); fe_sub<F>(Y3, Y3, t);
    fe_add<F>(Z3, Z1, H); fe_sqr<F>(Z3, Z3); fe_sub<F>(Z3, Z3, Z1Z1); fe_sub<F>(Z3, Z3, HH);
    bn_copy<C::Fp::N>(r, X3); bn_copy<C::Fp::N>(r + FPN, Y3); bn_copy<C::Fp::N>(r + 2 * FPN, Z3);
}

// add-2007-bl (Jacobian += Jacobian): 11M + 5S, complete via the rare-case branches
template <class C> BBS_HDN void g1_add(uint32_t* r, const uint32_t* p, const uint32_t* q) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { g1_copy<C>(r, q); return; }
    if (g1_is_inf<C>(q)) { g1_copy<C>(r, p); return; }
    const uint32_t *X1 = p, *Y1 = p + FPN, *Z1 = p + 2 * FPN, *X2 = q, *Y2 = q + FPN, *Z2 = q + 2 * FPN;
    uint32_t Z1Z1[FPN], Z2Z2[FPN], U1[FPN], U2[FPN], S1[FPN], S2[FPN], H[FPN], I[FPN], J[FPN], rr[FPN], V[FPN], t[FPN],
        X3[FPN], Y3[FPN], Z3[FPN];
    fe_sqr<F>(Z1Z1, Z1);
    fe_sqr<F>(Z2Z2, Z2);
    fe_mul<F>(U1, X1, Z2Z2);
    fe_mul<F>(U2, X2, Z1Z1);
    fe_mul<F>(S1, Y1, Z2); fe_mul<F>(S1, S1, Z2Z2);
    fe_mul<F>(S2, Y2, Z1); fe_mul<F>(S2, S2, Z1Z1);
    fe_sub<F>(H, U2, U1);
    fe_sub<F>(rr, S2, S1);
    if (bn_is_zero<C::Fp::N>(H)) {
        if (bn_is_zero<C::Fp::N>(rr)) { g1_dbl<C>(r, p); } else { g1_set_inf<C>(r); }
        return;
    }
    fe_dbl<F>(rr, rr);
    fe_dbl<F>(I, H); fe_sqr<F>(I, I);
    fe_mul<F>(J, H, I);
    fe_mul<F>(V, U1, I);
    fe_sqr<F>(X3, rr); fe_sub<F>(X3, X3, J); fe_sub<F>(X3, X3, V); fe_sub<F>(X3, X3, V);
    fe_sub<F>(t, V, X3); fe_mul<F>(Y3, rr, t);
    fe_mul<F>(t, S1, J); fe_dbl<F>(t, t); fe_sub<F>(Y3, Y3, t);
    fe_add<F>(Z3, Z1, Z2); fe_sqr<F>(Z3, Z3); fe_sub<F>(Z3, Z3, Z1Z1); fe_sub<F>(Z3, Z3, Z2Z2); fe_mul<F>(Z3, Z3, H);
    bn_copy<C::Fp::N>(r, X3); bn_copy<C::Fp::N>(r + FPN, Y3); bn_copy<C::Fp::N>(r + 2 * FPN, Z3);
}

// Jacobian -> affine (x, y); returns false for the identity (then r is zeroed)
template <class C> BBS_HDN bool g1_to_affine(uint32_t* r, const uint32_t* p) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { bn_zero<2 * C::Fp::N>(r); return false; }
    uint32_t zi[FPN], zi2[FPN];
    fe_inv<F>(zi, p + 2 * FPN);
    fe_sqr<F>(zi2, zi);
    fe_mul<F>(r, p, zi2);
    fe_mul<F>(zi2, zi2, zi);
    fe_mul<F>(r + FPN, p + FPN, zi2);
    return true;
}

// y^2 == x^3 + b  (affine, Montgomery)
template <class C> BBS_HDN bool g1_on_curve(const uint32_t* a) {
    using F = typename C::Fp;
    uint32_t l[FPN], rr[FPN];
    fe_sqr<F>(l, a + FPN);
    fe_sqr<F>(rr, a); fe_mul<F>(rr, rr, a); fe_add<F>(rr, rr, C::B());
    return bn_eq<C::Fp::N>(l, rr);
}

// r = k * P  (P Jacobian), k = canonical little-endian limbs, `bits` significant bits; MSB-first
// double-and-add (what ark-ec's `Projective * Fr` does; variable time like the reference).
template <class C> BBS_HDN void g1_mul(uint32_t* r, const uint32_t* p, const uint32_t* k, int bits) {
    uint32_t acc[G1J], base[G1J];
    g1_copy<C>(base, p);
    g1_set_inf<C>(acc);
    for (int i = bits - 1; i >= 0; i--) {
        g1_dbl<C>(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) g1_add<C>(acc, acc, base);
    }
    g1_copy<C>(r, acc);
}
// same with an affine base (mixed additions)
template <class C> BBS_HDN void g1_mul_affine(uint32_t* r, const uint32_t* a, const uint32_t* k, int bits) {
    uint32_t acc[G1J];
    g1_set_inf<C>(acc);
    for (int i = bits - 1; i >= 0; i--) {
        g1_dbl<C>(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) g1_add_mixed<C>(acc, acc, a);
    }
    g1_copy<C>(r, acc);
}

// ---- encodings -------------------------------------------------------
enum : int { PT_OK = 0, PT_INF = 1, PT_BAD = 2 };

// compressed bytes -> affine Montgomery point.  PT_INF for the identity encoding, PT_BAD for anything
// ark's deserializer would refuse (bad flags, x >= p, x not on the curve).  No subgroup check: the
// reference's verify functions take already-typed points and perform none.
template <class C> BBS_HDN int g1_decompress(uint32_t* r /*affine*/, const uint8_t* in);
template <class C> BBS_HDN void g1_compress_affine(uint8_t* out, const uint32_t* a /*affine*/, bool inf);

template <class C> BBS_HDN int g1_finish_decompress(uint32_t* r, uint32_t* x_canon, bool want_high) {
    using F = typename C::Fp;
    if (!fe_is_canonical<F>(x_canon)) return PT_BAD;
    uint32_t x[FPN], rhs[FPN], y[FPN];
    fe_to_mont<F>(x, x_canon);
    fe_sqr<F>(rhs, x); fe_mul<F>(rhs, rhs, x); fe_add<F>(rhs, rhs, C::B());
    if (!fe_sqrt<F>(y, rhs)) return PT_BAD;
    if (fe_is_high<F>(y) != want_high) fe_neg<F>(y, y);
    bn_copy<C::Fp::N>(r, x); bn_copy<C::Fp::N>(r + FPN, y);
    return PT_OK;
}

// BLS12-381: zcash / IETF format, 48 bytes big-endian x, flags in the top 3 bits of byte 0
template <> BBS_HDN int g1_decompress<Bls>(uint32_t* r, const uint8_t* in) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return PT_BAD;
    if (b0 & 0x40) {
        uint32_t o = b0 & 0x3f;
        for (int i = 1; i < 48; i++) o |= in[i];
        bn_zero<24>(r);
        return o ? PT_BAD : PT_INF;
    }
    uint8_t tmp[48];
    for (int i = 0; i < 48; i++) tmp[i] = in[i];
    tmp[0] = b0 & 0x1f;
    uint32_t x[12];
    limbs_from_be<12>(x, tmp);
    return g1_finish_decompress<Bls>(r, x, (b0 & 0x20) != 0);
}
template <> BBS_HDN void g1_compress_affine<Bls>(uint8_t* out, const uint32_t* a, bool inf) {
    if (inf) { out[0] = 0xc0; for (int i = 1; i < 48; i++) out[i] = 0; return; }
    uint32_t x[12];
    fe_from_mont<BlsFp>(x, a);
    limbs_to_be<12>(out, x);
    out[0] |= 0x80;
    if (fe_is_high<BlsFp>(a + 12)) out[0] |= 0x20;
}
// BN254: arkworks default SW format, 32 bytes little-endian x, flags in the top 2 bits of byte 31
template <> BBS_HDN int g1_decompress<Bn>(uint32_t* r, const uint8_t* in) {
    uint8_t fl = in[31] & 0xc0;
    if (fl == 0xc0) return PT_BAD;
    if (fl & 0x40) { bn_zero<16>(r); return PT_INF; }
    uint8_t tmp[32];
    for (int i = 0; i < 32; i++) tmp[i] = in[i];
    tmp[31] &= 0x3f;
    uint32_t x[8];
    limbs_from_le<8>(x, tmp);
    return g1_finish_decompress<Bn>(r, x, (fl & 0x80) != 0);
}
template <> BBS_HDN void g1_compress_affine<Bn>(uint8_t* out, const uint32_t* a, bool inf) {
    if (inf) { for (int i = 0; i < 32; i++) out[i] = 0; out[31] = 0x40; return; }
    uint32_t x[8];
    fe_from_mont<BnFp>(x, a);
    limbs_to_le<8>(out, x);
    if (fe_is_high<BnFp>(a + 8)) out[31] |= 0x80;
}

// Jacobian -> compressed bytes (one inversion)
template <class C> BBS_HDN void g1_compress(uint8_t* out, const uint32_t* p) {
    uint32_t a[G1A];
    bool ok = g1_to_affine<C>(a, p);
    g1_compress_affine<C>(out, a, !ok);
}

// ---- scalars --------------------------------------------------------------------------------------
// 32-byte little-endian (ark `Fr::serialize_compressed`) -> canonical limbs; false if >= r
template <class C> BBS_HD bool fr_from_le32(uint32_t* r, const uint8_t* in) {
    limbs_from_le<8>(r, in);
    return fe_is_canonical<typename C::Fr>(r);
}

}  // namespace bbs
