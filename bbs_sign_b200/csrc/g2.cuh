// G2 helpers needed once per issuer key: decoding the public key (ark `CanonicalDeserialize` of
// `PublicKey{pk: E::G2}`, key_gen.rs:12-15): twist-curve membership and the prime-order subgroup test.
#pragma once
#include "pairing.cuh"

namespace bbs {

template <class C> BBS_HDN void f2_pow(uint32_t* r, const uint32_t* a, const uint32_t* e, int ebits) {
    BBS_A16 uint32_t acc[F2N];
    f2_one<C>(acc);
    for (int i = ebits - 1; i >= 0; i--) {
        f2_sqr<C>(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) f2_mul<C>(acc, acc, a);
    }
    f2_copy<C>(r, acc);
}

// sqrt in Fp2 for p = 3 mod 4 (Adj-Rodriguez-Henriquez, the algorithm ark-ff uses for this case)
template <class C> BBS_HDN bool f2_sqrt(uint32_t* r, const uint32_t* a) {
    using F = typename C::Fp;
    if (f2_is_zero<C>(a)) { f2_zero<C>(r); return true; }
    BBS_A16 uint32_t a1[F2N], alpha[F2N], x0[F2N], m1[F2N], cand[F2N], chk[F2N];
    f2_pow<C>(a1, a, F::EXP_PM3D4(), F::BITS);
    f2_sqr<C>(alpha, a1); f2_mul<C>(alpha, alpha, a);
    f2_mul<C>(x0, a1, a);
    f2_one<C>(m1); f2_neg<C>(m1, m1);
    if (f2_eq<C>(alpha, m1)) {
        // cand = u * x0
        fe_neg<F>(cand, x0 + FPN); bn_copy<C::Fp::N>(cand + FPN, x0);
    } else {
        BBS_A16 uint32_t b[F2N];
        f2_one<C>(b); f2_add<C>(b, b, alpha);
        f2_pow<C>(b, b, F::HALF(), F::BITS);
        f2_mul<C>(cand, b, x0);
    }
    f2_sqr<C>(chk, cand);
    f2_copy<C>(r, cand);
    return f2_eq<C>(chk, a);
}

// arkworks QuadExt ordering: a > b compares c1 first, then c0 (canonical integers)
template <class C> BBS_HD bool f2_is_high(const uint32_t* y) {
    using F = typename C::Fp;
    // y > -y  <=>  (c1 > (p-1)/2) or (c1 == 0 and c0 > (p-1)/2)
    if (!bn_is_zero<C::Fp::N>(y + FPN)) return fe_is_high<F>(y + FPN);
    return fe_is_high<F>(y);
}

template <class C> BBS_HDN void g2_twist_b(uint32_t* b2);
template <> BBS_HDN void g2_twist_b<Bls>(uint32_t* b2) {   // 4 (1+u)
    BBS_A16 uint32_t four[12];
    fe_set_one<BlsFp>(four); fe_dbl<BlsFp>(four, four); fe_dbl<BlsFp>(four, four);
    bn_copy<12>(b2, four); bn_copy<12>(b2 + 12, four);
}
template <> BBS_HDN void g2_twist_b<Bn>(uint32_t* b2) {    // 3 / (9+u)
    BBS_A16 uint32_t xi[16], three[16];
    bn_copy<16>(xi, BN_XI());
    f2_inv<Bn>(xi, xi);
    fe_set_one<BnFp>(three); fe_dbl<BnFp>(three + 8, three); fe_add<BnFp>(three, three, three + 8); bn_zero<8>(three + 8);
    f2_mul<Bn>(b2, three, xi);
}

// ---- subgroup membership of a public key ---------------------------------------------------------------
// ark `CanonicalDeserialize` of PublicKey{pk: E::G2} (key_gen.rs:12-15) validates the r-torsion; both twists have large
// cofactors.  Test with the untwist-Frobenius-twist endomorphism psi [Scott, "A note on group membership tests"]:
//   BLS12-381:  psi(Q) == [x] Q, x < 0   i.e.  [|x|] Q == -psi(Q);       BN254:  psi(Q) == [6 t^2] Q.
// Sufficient: psi^2 - tr psi + p = 0 on E'(Fp2), and gcd(k^2 - tr k + p, #E'(Fp2)) = r for k = x resp. 6 t^2
// (checked in tests/test_oracle_kat.py).  Jacobian coordinates over Fp2, public data, once per context.
#define G2J (6 * C::Fp::N)
template <class C> BBS_HDN void g2j_dbl(uint32_t* r, const uint32_t* p) {          // dbl-2009-l
    const uint32_t *X = p, *Y = p + F2N, *Z = p + 2 * F2N;
    BBS_A16 uint32_t A[F2N], B[F2N], Cc[F2N], D[F2N], E[F2N], Fq[F2N], t[F2N], Z3[F2N];
    f2_sqr<C>(A, X);
    f2_sqr<C>(B, Y);
    f2_sqr<C>(Cc, B);
    f2_add<C>(t, X, B); f2_sqr<C>(t, t); f2_sub<C>(t, t, A); f2_sub<C>(t, t, Cc); f2_dbl<C>(D, t);
    f2_dbl<C>(E, A); f2_add<C>(E, E, A);
    f2_sqr<C>(Fq, E);
    f2_mul<C>(Z3, Y, Z); f2_dbl<C>(Z3, Z3);
    f2_dbl<C>(t, D); f2_sub<C>(r, Fq, t);
    f2_sub<C>(t, D, r); f2_mul<C>(t, E, t);
    f2_dbl<C>(Cc, Cc); f2_dbl<C>(Cc, Cc); f2_dbl<C>(Cc, Cc);
    f2_sub<C>(r + F2N, t, Cc);
    f2_copy<C>(r + 2 * F2N, Z3);
}
template <class C> BBS_HDN void g2j_add_mixed(uint32_t* r, const uint32_t* p, const uint32_t* q /*affine*/) {   // madd-2007-bl
    const uint32_t *X1 = p, *Y1 = p + F2N, *Z1 = p + 2 * F2N, *X2 = q, *Y2 = q + F2N;
    if (f2_is_zero<C>(Z1)) { f2_copy<C>(r, X2); f2_copy<C>(r + F2N, Y2); f2_one<C>(r + 2 * F2N); return; }
    BBS_A16 uint32_t Z1Z1[F2N], U2[F2N], S2[F2N], H[F2N], HH[F2N], I[F2N], J[F2N], rr[F2N], V[F2N], t[F2N], X3[F2N], Y3[F2N], Z3[F2N];
    f2_sqr<C>(Z1Z1, Z1);
    f2_mul<C>(U2, X2, Z1Z1);
    f2_mul<C>(S2, Y2, Z1); f2_mul<C>(S2, S2, Z1Z1);
    f2_sub<C>(H, U2, X1);
    f2_sub<C>(rr, S2, Y1);
    if (f2_is_zero<C>(H)) {
        if (f2_is_zero<C>(rr)) { g2j_dbl<C>(r, p); } else { f2_one<C>(r); f2_one<C>(r + F2N); f2_zero<C>(r + 2 * F2N); }
        return;
    }
    f2_dbl<C>(rr, rr);
    f2_sqr<C>(HH, H);
    f2_dbl<C>(I, HH); f2_dbl<C>(I, I);
    f2_mul<C>(J, H, I);
    f2_mul<C>(V, X1, I);
    f2_sqr<C>(X3, rr); f2_sub<C>(X3, X3, J); f2_sub<C>(X3, X3, V); f2_sub<C>(X3, X3, V);
    f2_sub<C>(t, V, X3); f2_mul<C>(Y3, rr, t);
    f2_mul<C>(t, Y1, J); f2_dbl<C>(t, t); f2_sub<C>(Y3, Y3, t);
    f2_add<C>(Z3, Z1, H); f2_sqr<C>(Z3, Z3); f2_sub<C>(Z3, Z3, Z1Z1); f2_sub<C>(Z3, Z3, HH);
    f2_copy<C>(r, X3); f2_copy<C>(r + F2N, Y3); f2_copy<C>(r + 2 * F2N, Z3);
}
// the multiplier k of the test, little-endian limbs: |x| (BLS12-381), 6 t^2 (BN254)
template <class C> struct G2Check;
template <> struct G2Check<Bls> {
    static constexpr int BITS = 64;
    static BBS_HD uint32_t limb(int i) { return (uint32_t)(BLS_X_ABS >> (32 * i)); }
};
template <> struct G2Check<Bn> {
    static constexpr int BITS = 128;
    static BBS_HD uint32_t limb(int i) {
        const unsigned __int128 k = (unsigned __int128)6 * BN_T * BN_T;
        return (uint32_t)(k >> (32 * i));
    }
};
template <class C> BBS_HDN bool g2_in_subgroup(const uint32_t* q /*affine [x|y], on the twist*/) {
    BBS_A16 uint32_t acc[G2J];
    f2_one<C>(acc); f2_one<C>(acc + F2N); f2_zero<C>(acc + 2 * F2N);
    for (int i = G2Check<C>::BITS - 1; i >= 0; i--) {
        g2j_dbl<C>(acc, acc);
        if ((G2Check<C>::limb(i >> 5) >> (i & 31)) & 1) g2j_add_mixed<C>(acc, acc, q);
    }
    if (f2_is_zero<C>(acc + 2 * F2N)) return false;
    // psi(x, y) = (conj(x) g_x, conj(y) g_y) with g_x = xi^((p-1)/3), g_y = xi^((p-1)/2) on the D-type twist (BN254) and
    // their inverses on the M-type twist (BLS12-381); compared cross-multiplied, so no inversion:
    //   D: X == conj(x) g_x Z^2,  Y == conj(y) g_y Z^3          M (and the sign of x < 0): X g_x == conj(x) Z^2,  Y g_y == -conj(y) Z^3
    const uint32_t *gx = C::FROB(1) + 2 * F2N, *gy = C::FROB(1) + 3 * F2N;
    BBS_A16 uint32_t zz[F2N], zzz[F2N], lx[F2N], ly[F2N], rx[F2N], ry[F2N];
    f2_sqr<C>(zz, acc + 2 * F2N); f2_mul<C>(zzz, zz, acc + 2 * F2N);
    f2_conj<C>(rx, q); f2_mul<C>(rx, rx, zz);
    f2_conj<C>(ry, q + F2N); f2_mul<C>(ry, ry, zzz);
    if (C::M_TWIST) {
        f2_mul<C>(lx, acc, gx); f2_mul<C>(ly, acc + F2N, gy); f2_neg<C>(ry, ry);
    } else {
        f2_copy<C>(lx, acc); f2_copy<C>(ly, acc + F2N); f2_mul<C>(rx, rx, gx); f2_mul<C>(ry, ry, gy);
    }
    BBS_OPAQUE_CALL_BARRIER();
    const bool ok = f2_eq<C>(lx, rx) && f2_eq<C>(ly, ry);
    BBS_OPAQUE_CALL_BARRIER();
    return ok;
}

template <class C> BBS_HDN int g2_finish_decompress(uint32_t* r, uint32_t* xc0, uint32_t* xc1, bool want_high) {
    using F = typename C::Fp;
    if (!fe_is_canonical<F>(xc0) || !fe_is_canonical<F>(xc1)) return PT_BAD;
    BBS_A16 uint32_t x[F2N], rhs[F2N], y[F2N], b2[F2N];
    fe_to_mont<F>(x, xc0); fe_to_mont<F>(x + FPN, xc1);
    g2_twist_b<C>(b2);
    f2_sqr<C>(rhs, x); f2_mul<C>(rhs, rhs, x); f2_add<C>(rhs, rhs, b2);
    if (!f2_sqrt<C>(y, rhs)) return PT_BAD;
    if (f2_is_high<C>(y) != want_high) f2_neg<C>(y, y);
    f2_copy<C>(r, x); f2_copy<C>(r + F2N, y);
    return g2_in_subgroup<C>(r) ? PT_OK : PT_BAD;
}

template <class C> BBS_HDN int g2_decompress(uint32_t* r /*affine [x|y] Fp2*/, const uint8_t* in);
// BLS12-381 (zcash): 96 bytes = BE(x.c1) || BE(x.c0), flags in byte 0
template <> BBS_HDN int g2_decompress<Bls>(uint32_t* r, const uint8_t* in) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return PT_BAD;
    if (b0 & 0x40) {
        uint32_t o = b0 & 0x3f;
        for (int i = 1; i < 96; i++) o |= in[i];
        bn_zero<48>(r);
        return o ? PT_BAD : PT_INF;
    }
    uint8_t tmp[48];
    for (int i = 0; i < 48; i++) tmp[i] = in[i];
    tmp[0] = b0 & 0x1f;
    BBS_A16 uint32_t c1[12], c0[12];
    limbs_from_be<12>(c1, tmp);
    limbs_from_be<12>(c0, in + 48);
    return g2_finish_decompress<Bls>(r, c0, c1, (b0 & 0x20) != 0);
}
// BN254 (ark default): 64 bytes = LE(x.c0) || LE(x.c1), flags in byte 63
template <> BBS_HDN int g2_decompress<Bn>(uint32_t* r, const uint8_t* in) {
    uint8_t fl = in[63] & 0xc0;
    if (fl == 0xc0) return PT_BAD;
    if (fl & 0x40) { bn_zero<32>(r); return PT_INF; }
    uint8_t tmp[32];
    for (int i = 0; i < 32; i++) tmp[i] = in[32 + i];
    tmp[31] &= 0x3f;
    BBS_A16 uint32_t c0[8], c1[8];
    limbs_from_le<8>(c0, in);
    limbs_from_le<8>(c1, tmp);
    return g2_finish_decompress<Bn>(r, c0, c1, (fl & 0x80) != 0);
}

}  // namespace bbs
