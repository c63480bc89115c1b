// create_generators + hash-to-G1 on the device: src/utils/interface_utilities.rs:47-73 with the two
// `HashToG1` implementations :24-44 -- zkcrypto's BLS12381G1_XMD:SHA-256_SSWU_RO_ and bn254_hash2curve's
// BN254G1_XMD:SHA-256_SVDW_RO_ (RFC 9380).  Setup-path code: one thread per generator, plain field calls.
//   generators: v = xmd(seed, seed_dst, 48); for i = 1..count: v = xmd(v || I2OSP(i, 8), seed_dst, 48); G_i = H2G(v, gen_dst)
#pragma once
#include "g1.cuh"
#include "sha256.cuh"

namespace bbs {

// expand_message_xmd(msg, dst, len) for len <= 128 bytes (utilities_helper.rs:42-97); out: len/4 big-endian words
BBS_HDN void xmd_expand(uint32_t* out, int len_bytes, const uint8_t* m1, uint32_t m1_len, const uint8_t* m2,
                        uint32_t m2_len, const uint8_t* dst, uint32_t dst_len) {
    Sha256 s;
    s.init_after_zpad();
    s.update(m1, m1_len);
    s.update(m2, m2_len);
    s.put((uint8_t)(len_bytes >> 8)); s.put((uint8_t)len_bytes); s.put(0);
    s.update(dst, dst_len); s.put((uint8_t)dst_len);
    BBS_A16 uint32_t b0[8], bi[8];
    s.finish(b0);
    const int ell = (len_bytes + 31) / 32;
    for (int i = 1; i <= ell; i++) {
        BBS_A16 uint32_t x[8];
        for (int k = 0; k < 8; k++) x[k] = i == 1 ? b0[k] : (b0[k] ^ bi[k]);
        s.init(); s.update_words(x, 8); s.put((uint8_t)i); s.update(dst, dst_len); s.put((uint8_t)dst_len);
        s.finish(bi);
        for (int k = 0; k < 8; k++) {
            int w = (i - 1) * 8 + k;
            if (w * 4 < len_bytes) out[w] = bi[k];
        }
    }
}

// big-endian words (nwords = N + 4) -> Montgomery element = value mod p.  The value is cut into c0 (low N-1 limbs) and c1
// (the 5 limbs above): both are < p, which the device Montgomery product requires of its operands (field.cuh fe_mul).
template <class F> BBS_HDN void fe_from_wide_be(uint32_t* r, const uint32_t* be, int nwords) {
    constexpr int N = F::N;
    BBS_A16 uint32_t c0[N], c1[N], sh[N];
    for (int i = 0; i < N; i++) {
        c0[i] = i < N - 1 ? be[nwords - 1 - i] : 0u;
        c1[i] = (N - 1 + i) < nwords ? be[nwords - 1 - (N - 1) - i] : 0u;
        sh[i] = i == N - 1 ? 1u : 0u;          // 2^(32 (N-1))
    }
    fe_mul<F>(c0, c0, F::R2());
    fe_mul<F>(c1, c1, F::R2());
    fe_mul<F>(sh, sh, F::R2());
    fe_mul<F>(c1, c1, sh);
    fe_add<F>(r, c0, c1);
}

// out of line on purpose: cicc folds comparisons of values an out-of-line callee has just written (DESIGN.md compiler note)
template <int N> BBS_HDN bool h2c_is_zero(const uint32_t* a) {
    uint32_t o = 0;
    for (int i = 0; i < N; i++) o |= a[i];
    return o == 0;
}
// sgn0(u) != sgn0(y) for Montgomery inputs (sgn0 = parity of the canonical integer)
template <class F> BBS_HDN bool h2c_sgn_differs(const uint32_t* u, const uint32_t* y) {
    BBS_A16 uint32_t uc[F::N], yc[F::N];
    fe_from_mont<F>(uc, u);
    fe_from_mont<F>(yc, y);
    return ((uc[0] ^ yc[0]) & 1u) != 0;
}

// ---- BLS12-381: simplified SWU onto E', 11-isogeny by Velu's formulas, cofactor clearing by h_eff ---------------------
BBS_HDN void bls_sswu(uint32_t* X, uint32_t* Y, const uint32_t* u) {
    using F = BlsFp;
    constexpr int N = 12;
    BBS_A16 uint32_t u2[N], tv[N], tv1[N], x1[N], g[N], t[N], one[N];
    fe_set_one<F>(one);
    fe_sqr<F>(u2, u);
    fe_mul<F>(t, u2, BLS_H2C_Z());             // Z u^2
    fe_sqr<F>(tv, t);                          // Z^2 u^4
    fe_add<F>(tv, tv, t);
    fe_inv_vt<F>(tv1, tv);                     // inv0
    if (h2c_is_zero<N>(tv1)) bn_copy<N>(x1, BLS_H2C_BZA());
    else { fe_add<F>(x1, one, tv1); fe_mul<F>(x1, x1, BLS_H2C_NBA()); }
    // g(x1) = x1^3 + A x1 + B
    fe_sqr<F>(g, x1); fe_add<F>(g, g, BLS_H2C_A()); fe_mul<F>(g, g, x1); fe_add<F>(g, g, BLS_H2C_B());
    BBS_A16 uint32_t y[N];
    if (fe_sqrt<F>(y, g)) { bn_copy<N>(X, x1); }
    else {
        fe_mul<F>(X, t, x1);                   // Z u^2 x1
        fe_sqr<F>(g, X); fe_add<F>(g, g, BLS_H2C_A()); fe_mul<F>(g, g, X); fe_add<F>(g, g, BLS_H2C_B());
        fe_sqrt<F>(y, g);
    }
    if (h2c_sgn_differs<F>(u, y)) fe_neg<F>(y, y);
    bn_copy<N>(Y, y);
}
BBS_HDN void bls_iso11(uint32_t* xo, uint32_t* yo, const uint32_t* X, const uint32_t* Y) {
    using F = BlsFp;
    constexpr int N = 12;
    BBS_A16 uint32_t ax[N], ay[N];
    bn_copy<N>(ax, X);
    bn_copy<N>(ay, Y);
    for (int q = 0; q < 5; q++) {
        const uint32_t* tb = BLS_H2C_VELU() + q * 5 * N;      // xq, yq, gx*gy, vq, uq
        BBS_A16 uint32_t d[N], d2[N], d3[N], t[N], s[N];
        fe_sub<F>(d, X, tb);
        fe_inv_vt<F>(d, d);
        fe_sqr<F>(d2, d);
        fe_mul<F>(d3, d2, d);
        fe_mul<F>(t, tb + 3 * N, d); fe_add<F>(ax, ax, t);            // + vq d
        fe_mul<F>(t, tb + 4 * N, d2); fe_add<F>(ax, ax, t);           // + uq d^2
        fe_mul<F>(t, tb + 4 * N, Y); fe_dbl<F>(t, t); fe_mul<F>(t, t, d3); fe_sub<F>(ay, ay, t);   // - 2 uq Y d^3
        fe_sub<F>(s, Y, tb + N); fe_mul<F>(s, s, tb + 3 * N); fe_mul<F>(s, s, d2); fe_sub<F>(ay, ay, s);   // - vq (Y - yq) d^2
        fe_mul<F>(t, tb + 2 * N, d2); fe_add<F>(ay, ay, t);           // + gx gy d^2
    }
    fe_mul<F>(xo, ax, BLS_H2C_I11_2());
    fe_mul<F>(yo, ay, BLS_H2C_I11_3());
}

// out: affine Montgomery point (never the identity in practice)
template <class C> BBS_HDN void hash_to_g1(uint32_t* out_aff, const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len);
template <> BBS_HDN void hash_to_g1<Bls>(uint32_t* out_aff, const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len) {
    using F = BlsFp;
    BBS_A16 uint32_t ub[32], u[12], P0[24], P1[24], x[12], y[12];
    xmd_expand(ub, 128, msg, len, nullptr, 0, dst, dst_len);
    fe_from_wide_be<F>(u, ub, 16);
    bls_sswu(x, y, u); bls_iso11(P0, P0 + 12, x, y);
    fe_from_wide_be<F>(u, ub + 16, 16);
    bls_sswu(x, y, u); bls_iso11(P1, P1 + 12, x, y);
    BBS_A16 uint32_t R[36], Ra[24], acc[36];
    g1_from_affine<Bls>(R, P0);
    g1_add_mixed<Bls>(R, R, P1);
    g1_to_affine_vt<Bls>(Ra, R);
    BBS_A16 uint32_t k[2] = {(uint32_t)BLS_H2C_HEFF, (uint32_t)(BLS_H2C_HEFF >> 32)};
    g1_mul_affine<Bls>(acc, Ra, k, 64);
    g1_to_affine_vt<Bls>(out_aff, acc);
}

// ---- BN254: Shallue-van de Woestijne, Z = 1 (cofactor 1) -------------------------------------------------------------
BBS_HDN void bn_svdw(uint32_t* X, uint32_t* Y, const uint32_t* u) {
    using F = BnFp;
    constexpr int N = 8;
    BBS_A16 uint32_t one[N], tv1[N], tv2[N], tv3[N], tv4[N], x1[N], x2[N], x3[N], g[N], y[N], t[N];
    fe_set_one<F>(one);
    fe_sqr<F>(tv1, u); fe_mul<F>(tv1, tv1, BN_H2C_C1());
    fe_add<F>(tv2, one, tv1);
    fe_sub<F>(tv1, one, tv1);
    fe_mul<F>(tv3, tv1, tv2); fe_inv_vt<F>(tv3, tv3);
    fe_mul<F>(tv4, u, tv1); fe_mul<F>(tv4, tv4, tv3); fe_mul<F>(tv4, tv4, BN_H2C_C3());
    fe_sub<F>(x1, BN_H2C_C2(), tv4);
    fe_add<F>(x2, BN_H2C_C2(), tv4);
    fe_sqr<F>(x3, tv2); fe_mul<F>(x3, x3, tv3); fe_sqr<F>(x3, x3); fe_mul<F>(x3, x3, BN_H2C_C4()); fe_add<F>(x3, x3, one);
    // g(x) = x^3 + 3
    fe_sqr<F>(g, x1); fe_mul<F>(g, g, x1); fe_add<F>(g, g, BN_B());
    bool e1 = fe_sqrt<F>(y, g);
    if (e1) bn_copy<N>(X, x1);
    else {
        fe_sqr<F>(g, x2); fe_mul<F>(g, g, x2); fe_add<F>(g, g, BN_B());
        if (fe_sqrt<F>(y, g)) bn_copy<N>(X, x2);
        else {
            fe_sqr<F>(g, x3); fe_mul<F>(g, g, x3); fe_add<F>(g, g, BN_B());
            fe_sqrt<F>(y, g);
            bn_copy<N>(X, x3);
        }
    }
    if (h2c_sgn_differs<F>(u, y)) fe_neg<F>(y, y);
    bn_copy<N>(Y, y);
    (void)t;
}
template <> BBS_HDN void hash_to_g1<Bn>(uint32_t* out_aff, const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len) {
    using F = BnFp;
    BBS_A16 uint32_t ub[24], u[8], P0[16], P1[16];
    xmd_expand(ub, 96, msg, len, nullptr, 0, dst, dst_len);
    fe_from_wide_be<F>(u, ub, 12);
    bn_svdw(P0, P0 + 8, u);
    fe_from_wide_be<F>(u, ub + 12, 12);
    bn_svdw(P1, P1 + 8, u);
    BBS_A16 uint32_t R[24];
    g1_from_affine<Bn>(R, P0);
    g1_add_mixed<Bn>(R, R, P1);
    g1_to_affine_vt<Bn>(out_aff, R);
}

// ---- create_generators -----------------------------------------------------------------------------------------------
struct GenSeedArgs {
    const uint8_t* api_id; uint32_t api_id_len; uint32_t count;
    uint8_t* v_out;            // count x 48 bytes: the hash-to-curve inputs v_1 .. v_count
};
// one thread: the sequential seed chain (interface_utilities.rs:52-67)
BBS_HD void gen_seed_item(const GenSeedArgs& a, uint32_t) {
    uint8_t seed[160], seed_dst[160];
    uint32_t n = a.api_id_len;
    for (uint32_t i = 0; i < n; i++) { seed[i] = a.api_id[i]; seed_dst[i] = a.api_id[i]; }
    const char* s1 = "MESSAGE_GENERATOR_SEED";
    const char* s2 = "SIG_GENERATOR_SEED_";
    uint32_t l1 = 0, l2 = 0;
    while (s1[l1]) { seed[n + l1] = (uint8_t)s1[l1]; l1++; }
    while (s2[l2]) { seed_dst[n + l2] = (uint8_t)s2[l2]; l2++; }
    BBS_A16 uint32_t v[12];
    xmd_expand(v, 48, seed, n + l1, nullptr, 0, seed_dst, n + l2);
    for (uint32_t i = 1; i <= a.count; i++) {
        uint8_t vb[48], ctr[8];
        for (int k = 0; k < 12; k++) { vb[4 * k] = (uint8_t)(v[k] >> 24); vb[4 * k + 1] = (uint8_t)(v[k] >> 16); vb[4 * k + 2] = (uint8_t)(v[k] >> 8); vb[4 * k + 3] = (uint8_t)v[k]; }
        for (int k = 0; k < 8; k++) ctr[k] = (uint8_t)((uint64_t)i >> (8 * (7 - k)));
        xmd_expand(v, 48, vb, 48, ctr, 8, seed_dst, n + l2);
        uint8_t* o = a.v_out + (size_t)(i - 1) * 48;
        for (int k = 0; k < 12; k++) { o[4 * k] = (uint8_t)(v[k] >> 24); o[4 * k + 1] = (uint8_t)(v[k] >> 16); o[4 * k + 2] = (uint8_t)(v[k] >> 8); o[4 * k + 3] = (uint8_t)v[k]; }
    }
}
struct GenPointArgs {
    const uint8_t* api_id; uint32_t api_id_len;
    const uint8_t* v;          // count x 48
    uint8_t* out;              // count x G1 compressed
};
template <class C> BBS_HD void gen_point_item(const GenPointArgs& a, uint32_t i) {
    uint8_t dst[160];
    uint32_t n = a.api_id_len;
    for (uint32_t k = 0; k < n; k++) dst[k] = a.api_id[k];
    const char* s = "SIG_GENERATOR_DST_";
    uint32_t l = 0;
    while (s[l]) { dst[n + l] = (uint8_t)s[l]; l++; }
    BBS_A16 uint32_t P[2 * C::Fp::N];
    hash_to_g1<C>(P, a.v + (size_t)i * 48, 48, dst, n + l);
    g1_compress_affine<C>(a.out + (size_t)i * C::G1_BYTES, P, false);
}

}  // namespace bbs
