// G2 helpers needed once per issuer key: decoding the public key (ark `CanonicalDeserialize` of
// `PublicKey{pk: E::G2}`, key_gen.rs:12-15) and the twist-curve membership test.  Affine, Fp2.
#pragma once
#include "pairing.cuh"

namespace bbs {

template <class C> BBS_HDN void f2_pow(uint32_t* r, const uint32_t* a, const uint32_t* e, int ebits) {
    uint32_t acc[F2N];
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
    uint32_t a1[F2N], alpha[F2N], x0[F2N], m1[F2N], cand[F2N], chk[F2N];
    f2_pow<C>(a1, a, F::EXP_PM3D4(), F::BITS);
    f2_sqr<C>(alpha, a1); f2_mul<C>(alpha, alpha, a);
    f2_mul<C>(x0, a1, a);
    f2_one<C>(m1); f2_neg<C>(m1, m1);
    if (f2_eq<C>(alpha, m1)) {
        // cand = u * x0
        fe_neg<F>(cand, x0 + FPN); bn_copy<C::Fp::N>(cand + FPN, x0);
    } else {
        uint32_t b[F2N];
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
    uint32_t four[12];
    fe_set_one<BlsFp>(four); fe_dbl<BlsFp>(four, four); fe_dbl<BlsFp>(four, four);
    bn_copy<12>(b2, four); bn_copy<12>(b2 + 12, four);
}
template <> BBS_HDN void g2_twist_b<Bn>(uint32_t* b2) {    // 3 / (9+u)
    uint32_t xi[16], three[16];
    bn_copy<16>(xi, BN_XI());
    f2_inv<Bn>(xi, xi);
    fe_set_one<BnFp>(three); fe_dbl<BnFp>(three + 8, three); fe_add<BnFp>(three, three, three + 8); bn_zero<8>(three + 8);
    f2_mul<Bn>(b2, three, xi);
}

template <class C> BBS_HDN int g2_finish_decompress(uint32_t* r, uint32_t* xc0, uint32_t* xc1, bool want_high) {
    using F = typename C::Fp;
    if (!fe_is_canonical<F>(xc0) || !fe_is_canonical<F>(xc1)) return PT_BAD;
    uint32_t x[F2N], rhs[F2N], y[F2N], b2[F2N];
    fe_to_mont<F>(x, xc0); fe_to_mont<F>(x + FPN, xc1);
    g2_twist_b<C>(b2);
    f2_sqr<C>(rhs, x); f2_mul<C>(rhs, rhs, x); f2_add<C>(rhs, rhs, b2);
    if (!f2_sqrt<C>(y, rhs)) return PT_BAD;
    if (f2_is_high<C>(y) != want_high) f2_neg<C>(y, y);
    f2_copy<C>(r, x); f2_copy<C>(r + F2N, y);
    return PT_OK;
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
    uint32_t c1[12], c0[12];
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
    uint32_t c0[8], c1[8];
    limbs_from_le<8>(c0, in);
    limbs_from_le<8>(c1, tmp);
    return g2_finish_decompress<Bn>(r, c0, c1, (fl & 0x80) != 0);
}

}  // namespace bbs
