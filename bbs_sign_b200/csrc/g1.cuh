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
// out-of-line identity test for points that were just written through pointers by other out-of-line functions: cicc 12.9
// has folded such comparisons when they were inlined into the caller (field.cuh, BBS_OPAQUE_CALL_BARRIER)
template <class C> BBS_HDN bool g1_is_inf_ool(const uint32_t* p) {
    BBS_OPAQUE_CALL_BARRIER();
    bool r = bn_is_zero<C::Fp::N>(p + 2 * FPN);
    BBS_OPAQUE_CALL_BARRIER();
    return r;
}
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

// The three group operations write their result in place and keep as few temporaries as the formulas allow: every
// temporary is a 48- / 32-byte array in the thread's local memory, and the hot set of those arrays (x 512 resident
// threads per SM) is what decides the L1 hit rate of the per-thread kernels (profiles/summary_r02.md).  r may alias p
// (the accumulator pattern acc = acc + q); r never aliases q.

// dbl-2009-l: 2M + 5S
template <class C> BBS_HDN void g1_dbl(uint32_t* r, const uint32_t* p) {
    using F = typename C::Fp;
    const uint32_t *X = p, *Y = p + FPN, *Z = p + 2 * FPN;
    BBS_A16 uint32_t A[FPN], B[FPN], Cc[FPN], D[FPN], Fq[FPN];
    fe_sqr<F>(A, X);
    fe_sqr<F>(B, Y);
    fe_sqr<F>(Cc, B);
    fe_add<F>(D, X, B); fe_sqr<F>(D, D); fe_sub<F>(D, D, A); fe_sub<F>(D, D, Cc); fe_dbl<F>(D, D);   // D = 2((X+B)^2 - A - C)
    fe_dbl<F>(B, A); fe_add<F>(A, B, A);                        // E = 3A (in A)
    fe_sqr<F>(Fq, A);
    fe_mul<F>(B, Y, Z); fe_dbl<F>(r + 2 * FPN, B);              // Z3 = 2 Y Z; identity stays identity (Z3 = 0); X, Y are dead now
    fe_dbl<F>(B, D); fe_sub<F>(r, Fq, B);                       // X3 = F - 2D
    fe_sub<F>(D, D, r); fe_mul<F>(D, A, D);
    fe_dbl<F>(Cc, Cc); fe_dbl<F>(Cc, Cc); fe_dbl<F>(Cc, Cc);    // 8C
    fe_sub<F>(r + FPN, D, Cc);
}

// madd-2007-bl (Jacobian += affine): 7M + 4S, complete via the rare-case branches
template <class C> BBS_HDN void g1_add_mixed(uint32_t* r, const uint32_t* p, const uint32_t* q /*affine*/) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { g1_from_affine<C>(r, q); return; }
    const uint32_t *X1 = p, *Y1 = p + FPN, *Z1 = p + 2 * FPN, *X2 = q, *Y2 = q + FPN;
    BBS_A16 uint32_t Z1Z1[FPN], H[FPN], rr[FPN], HH[FPN], V[FPN], J[FPN];
    fe_sqr<F>(Z1Z1, Z1);
    fe_mul<F>(H, X2, Z1Z1);                                     // U2
    fe_mul<F>(rr, Y2, Z1); fe_mul<F>(rr, rr, Z1Z1);             // S2
    fe_sub<F>(H, H, X1);
    fe_sub<F>(rr, rr, Y1);
    if (bn_is_zero<C::Fp::N>(H)) {
        if (bn_is_zero<C::Fp::N>(rr)) { g1_dbl<C>(r, p); } else { g1_set_inf<C>(r); }
        return;
    }
    fe_dbl<F>(rr, rr);
    fe_sqr<F>(HH, H);
    fe_dbl<F>(V, HH); fe_dbl<F>(V, V);                          // I = 4 HH
    fe_mul<F>(J, H, V);
    fe_mul<F>(V, X1, V);                                        // V = X1 I
    fe_add<F>(H, Z1, H); fe_sqr<F>(H, H); fe_sub<F>(H, H, Z1Z1); fe_sub<F>(r + 2 * FPN, H, HH);   // Z3; Z1, X1 are dead now
    fe_sqr<F>(HH, rr); fe_sub<F>(HH, HH, J); fe_sub<F>(HH, HH, V); fe_sub<F>(r, HH, V);           // X3 = r^2 - J - 2V
    fe_sub<F>(V, V, r); fe_mul<F>(V, rr, V);
    fe_mul<F>(J, Y1, J); fe_dbl<F>(J, J);
    fe_sub<F>(r + FPN, V, J);                                   // Y3 = r (V - X3) - 2 Y1 J
}

// add-2007-bl (Jacobian += Jacobian): 11M + 5S, complete via the rare-case branches
template <class C> BBS_HDN void g1_add(uint32_t* r, const uint32_t* p, const uint32_t* q) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { g1_copy<C>(r, q); return; }
    if (g1_is_inf<C>(q)) { g1_copy<C>(r, p); return; }
    const uint32_t *X1 = p, *Y1 = p + FPN, *Z1 = p + 2 * FPN, *X2 = q, *Y2 = q + FPN, *Z2 = q + 2 * FPN;
    BBS_A16 uint32_t Z1Z1[FPN], Z2Z2[FPN], U1[FPN], S1[FPN], H[FPN], rr[FPN], V[FPN], J[FPN];
    fe_sqr<F>(Z1Z1, Z1);
    fe_sqr<F>(Z2Z2, Z2);
    fe_mul<F>(U1, X1, Z2Z2);
    fe_mul<F>(H, X2, Z1Z1);                                     // U2
    fe_mul<F>(S1, Y1, Z2); fe_mul<F>(S1, S1, Z2Z2);
    fe_mul<F>(rr, Y2, Z1); fe_mul<F>(rr, rr, Z1Z1);             // S2
    fe_sub<F>(H, H, U1);
    fe_sub<F>(rr, rr, S1);
    if (bn_is_zero<C::Fp::N>(H)) {
        if (bn_is_zero<C::Fp::N>(rr)) { g1_dbl<C>(r, p); } else { g1_set_inf<C>(r); }
        return;
    }
    fe_dbl<F>(rr, rr);
    fe_add<F>(V, Z1, Z2); fe_sqr<F>(V, V); fe_sub<F>(V, V, Z1Z1); fe_sub<F>(V, V, Z2Z2);
    fe_mul<F>(r + 2 * FPN, V, H);                               // Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) H; p and q are dead now
    fe_dbl<F>(V, H); fe_sqr<F>(V, V);                           // I = (2H)^2
    fe_mul<F>(J, H, V);
    fe_mul<F>(V, U1, V);                                        // V = U1 I
    fe_sqr<F>(H, rr); fe_sub<F>(H, H, J); fe_sub<F>(H, H, V); fe_sub<F>(r, H, V);                 // X3
    fe_sub<F>(V, V, r); fe_mul<F>(V, rr, V);
    fe_mul<F>(J, S1, J); fe_dbl<F>(J, J);
    fe_sub<F>(r + FPN, V, J);                                   // Y3
}

// Jacobian -> affine (x, y); returns false for the identity (then r is zeroed)
template <class C> BBS_HDN bool g1_to_affine(uint32_t* r, const uint32_t* p) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { bn_zero<2 * C::Fp::N>(r); return false; }
    BBS_A16 uint32_t zi[FPN], zi2[FPN];
    fe_inv<F>(zi, p + 2 * FPN);
    fe_sqr<F>(zi2, zi);
    fe_mul<F>(r, p, zi2);
    fe_mul<F>(zi2, zi2, zi);
    fe_mul<F>(r + FPN, p + FPN, zi2);
    return true;
}

// the same with the variable-time inversion: for one-thread tails and context creation (public points)
template <class C> BBS_HDN bool g1_to_affine_vt(uint32_t* r, const uint32_t* p) {
    using F = typename C::Fp;
    if (g1_is_inf<C>(p)) { bn_zero<2 * C::Fp::N>(r); return false; }
    BBS_A16 uint32_t zi[FPN], zi2[FPN];
    fe_inv_vt<F>(zi, p + 2 * FPN);
    fe_sqr<F>(zi2, zi);
    fe_mul<F>(r, p, zi2);
    fe_mul<F>(zi2, zi2, zi);
    fe_mul<F>(r + FPN, p + FPN, zi2);
    return true;
}

// y^2 == x^3 + b  (affine, Montgomery)
template <class C> BBS_HDN bool g1_on_curve(const uint32_t* a) {
    using F = typename C::Fp;
    BBS_A16 uint32_t l[FPN], rr[FPN];
    fe_sqr<F>(l, a + FPN);
    fe_sqr<F>(rr, a); fe_mul<F>(rr, rr, a); fe_add<F>(rr, rr, C::B());
    return bn_eq<C::Fp::N>(l, rr);
}

// r = k * P, k = little-endian limbs, `bits` significant bits; MSB-first double-and-add (what ark-ec's
// `Projective * Fr` does).  Used where the scalar is a public constant (context creation, self tests); the per-item
// multiplications use g1_msm_win4 below.
template <class C> BBS_HDN void g1_mul_affine(uint32_t* r, const uint32_t* a, const uint32_t* k, int bits) {
    BBS_A16 uint32_t acc[G1J];
    g1_set_inf<C>(acc);
    for (int i = bits - 1; i >= 0; i--) {
        g1_dbl<C>(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) g1_add_mixed<C>(acc, acc, a);
    }
    g1_copy<C>(r, acc);
}

// ---- windowed scalar multiplication with the GLV endomorphism ----------------------------------------------
// Variable-base k*P for the per-item points (e*A in core_verify, D*r3^ in proof_verify_init).  ark-ec's
// `Projective * Fr` is a bit-serial double-and-add; on a GPU its data-dependent additions diverge inside a warp
// (every lane pays for every addition), so this uses fixed signed 5-bit windows (one table look-up and one addition per
// window for every lane) and the curve endomorphism phi(x, y) = (beta x, y) = lambda P (both curves have j = 0):
// k = k1 + k2 lambda with both halves below 2^128, so k P = k1 P + k2 phi(P) needs 128 doublings instead of 255.  On
// BLS12-381 lambda = x^2 - 1 is itself 128 bits and k2 = floor(k / lambda), k1 = k mod lambda; on BN254 lambda is 254 bits and
// the halves come from a lattice reduction (bn_glv_split).  Same group element as the reference computes; only the
// addition chain differs.
// Signed 5-bit windows without carries: with C = sum_w 16 * 32^w, the base-32 digits u_w of k + C give
// k = sum_w (u_w - 16) 32^w, digits in [-16, 15]: 26 windows for k < 2^128 (k + C < 2^130), a table of 1..16 times the point,
// one addition per non-zero digit (the negative ones negate y).
constexpr int WIN5_WINDOWS = 26;
BBS_HD void win5_recode(uint32_t* kk /*5 limbs*/, const uint32_t* k /*>= 4 limbs, value < 2^128*/) {
    const uint32_t C[5] = {0x21084210u, 0x08421084u, 0x42108421u, 0x10842108u, 0x2u};
    uint64_t c = 0;
    for (int i = 0; i < 5; i++) { c += (uint64_t)(i < 4 ? k[i] : 0u) + C[i]; kk[i] = (uint32_t)c; c >>= 32; }
}
BBS_HD int win5_digit(const uint32_t* kk, int w) {
    const int bit = 5 * w, word = bit >> 5, off = bit & 31;
    uint64_t v = kk[word];
    if (word < 4) v |= (uint64_t)kk[word + 1] << 32;
    return (int)((uint32_t)(v >> off) & 31u) - 16;
}

// r = sum_j k1[j] * P_j (+ k2[j] * phi(P_j) when beta != nullptr) over NP affine points (pts[j] == nullptr: the
// identity, skipped), Straus with shared doublings; all k1[j], k2[j] below 2^128 (`bits` <= 128)
template <class C, int NP> BBS_HDN void g1_msm_win4(uint32_t* r, const uint32_t* const* pts, const uint32_t (*k1)[9],
                                                    const uint32_t (*k2)[9], int bits, const uint32_t* beta) {
    using F = typename C::Fp;
    (void)bits;
    BBS_A16 uint32_t tab[NP][16][3 * C::Fp::N];            // d * P_j, d = 1..16 (Jacobian)
    BBS_A16 uint32_t kk1[NP][5], kk2[NP][5];
    for (int j = 0; j < NP; j++) {
        if (!pts[j]) continue;
        g1_from_affine<C>(tab[j][0], pts[j]);
        g1_dbl<C>(tab[j][1], tab[j][0]);
        for (int d = 3; d <= 16; d++) g1_add_mixed<C>(tab[j][d - 1], tab[j][d - 2], pts[j]);
        win5_recode(kk1[j], k1[j]);
        if (beta) win5_recode(kk2[j], k2[j]);
    }
    BBS_A16 uint32_t acc[G1J];
    g1_set_inf<C>(acc);
    for (int w = WIN5_WINDOWS - 1; w >= 0; w--) {
        if (!g1_is_inf<C>(acc)) { g1_dbl<C>(acc, acc); g1_dbl<C>(acc, acc); g1_dbl<C>(acc, acc); g1_dbl<C>(acc, acc); g1_dbl<C>(acc, acc); }
        for (int j = 0; j < NP; j++) {
            if (!pts[j]) continue;
            const int d1 = win5_digit(kk1[j], w);
            if (d1) {
                BBS_A16 uint32_t t[G1J];
                const uint32_t* e = tab[j][(d1 > 0 ? d1 : -d1) - 1];
                bn_copy<C::Fp::N>(t, e);
                if (d1 > 0) bn_copy<C::Fp::N>(t + FPN, e + FPN); else fe_neg<F>(t + FPN, e + FPN);
                bn_copy<C::Fp::N>(t + 2 * FPN, e + 2 * FPN);
                g1_add<C>(acc, acc, t);
            }
            if (beta) {
                const int d2 = win5_digit(kk2[j], w);
                if (d2) {
                    BBS_A16 uint32_t t[G1J];
                    const uint32_t* e = tab[j][(d2 > 0 ? d2 : -d2) - 1];
                    fe_mul<F>(t, e, beta);
                    if (d2 > 0) bn_copy<C::Fp::N>(t + FPN, e + FPN); else fe_neg<F>(t + FPN, e + FPN);
                    bn_copy<C::Fp::N>(t + 2 * FPN, e + 2 * FPN);
                    g1_add<C>(acc, acc, t);
                }
            }
        }
    }
    g1_copy<C>(r, acc);
}

// k (8 canonical limbs, < r) = k1 + k2 * lambda, 0 <= k1 < lambda, k2 <= lambda + 1   [BLS12-381]
BBS_HD void bls_glv_split(uint32_t* k1, uint32_t* k2, const uint32_t* k) {
    const uint32_t* lam = BLS_GLV_LAMBDA();
    const uint32_t* mu = BLS_GLV_MU();       // floor(2^256 / lambda), 5 limbs
    BBS_A16 uint32_t prod[13];
    for (int i = 0; i < 13; i++) prod[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 5; j++) {
            c += (uint64_t)k[i] * mu[j] + prod[i + j];
            prod[i + j] = (uint32_t)c;
            c >>= 32;
        }
        prod[i + 5] = (uint32_t)c;
    }
    BBS_A16 uint32_t q[5];
    for (int i = 0; i < 5; i++) q[i] = prod[8 + i];          // q^ = floor(k mu / 2^256) in {q - 1, q}
    BBS_A16 uint32_t ql[9];
    for (int i = 0; i < 9; i++) ql[i] = 0;
    for (int i = 0; i < 5; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) {
            c += (uint64_t)q[i] * lam[j] + ql[i + j];
            ql[i + j] = (uint32_t)c;
            c >>= 32;
        }
        if (i + 4 < 9) ql[i + 4] = (uint32_t)c;
    }
    BBS_A16 uint32_t rem[5];
    {
        int64_t c = 0;
        for (int i = 0; i < 5; i++) {
            c += (int64_t)(i < 8 ? k[i] : 0) - (int64_t)ql[i];
            rem[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    for (int it = 0; it < 2; it++) {          // at most one correction is needed; two for safety
        BBS_A16 uint32_t d[5];
        int64_t c = 0;
        for (int i = 0; i < 5; i++) {
            c += (int64_t)rem[i] - (int64_t)(i < 4 ? lam[i] : 0);
            d[i] = (uint32_t)c;
            c >>= 32;
        }
        if (c == 0) {                          // rem >= lambda
            for (int i = 0; i < 5; i++) rem[i] = d[i];
            uint64_t cc = 1;
            for (int i = 0; i < 5; i++) { cc += q[i]; q[i] = (uint32_t)cc; cc >>= 32; }
        }
    }
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
    k1[4] = rem[4]; k2[4] = q[4];              // k1[4] == 0; k2[4] == 0 (k2 <= lambda + 1 < 2^128)
}

// out[0 .. NX + NY) = x * y (schoolbook on 32-bit limbs; setup-sized operands)
template <int NX, int NY> BBS_HD void mp_mul(uint32_t* out, const uint32_t* x, const uint32_t* y) {
    for (int i = 0; i < NX + NY; i++) out[i] = 0;
    for (int i = 0; i < NX; i++) {
        uint64_t c = 0;
        for (int j = 0; j < NY; j++) {
            c += (uint64_t)x[i] * y[j] + out[i + j];
            out[i + j] = (uint32_t)c;
            c >>= 32;
        }
        out[i + NY] = (uint32_t)c;
    }
}

// k (8 canonical limbs, < r) = k1 + k2 * lambda mod r with 0 <= k1, k2 < 2^128   [BN254]
// lambda is a 254-bit root of x^2 + x + 1, so the halves come from the lattice {(x, y): x + y lambda = 0 mod r} with basis
// v1 = (a, -b), v2 = (b, a + b), a = 6t^2 + 2t, b = 2t + 1 (det = r):  (k1, k2) = (k + r, 0) - c1 v1 - c2 v2  with
//   c1 = floor(((a + b)(k + r) + 2 b^2) / r),   c2 = floor((b (k + r) - 2 a b) / r)
// i.e. the coordinates of (k + r, -2b) rounded DOWN, which puts the remainder into the fundamental parallelogram shifted by
// (0, 2b): both halves non-negative (no signed digits, no point negations).  The two quotients are taken with reciprocals
// floor(2^384 / r), floor(2^320 / r): each may come out one too small, which adds one basis vector to the remainder; the
// halves stay below 2 (a + 2b) < 2^128 (tools/gen_constants.py checks the constants; the split is checked against
// k1 + k2 lambda = k on the device by the G1 self test and against the oracle by every BN254 parity test).
BBS_HD void bn_glv_split(uint32_t* k1, uint32_t* k2, const uint32_t* k) {
    BBS_A16 uint32_t kp[8];
    {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (uint64_t)k[i] + BN_FR_P()[i]; kp[i] = (uint32_t)c; c >>= 32; }   // < 2^255
    }
    BBS_A16 uint32_t n1[12], n2[10];
    mp_mul<4, 8>(n1, BN_GLV_AB(), kp);
    {
        uint64_t c = 0;
        for (int i = 0; i < 12; i++) { c += (uint64_t)n1[i] + (i < 5 ? BN_GLV_2BB()[i] : 0u); n1[i] = (uint32_t)c; c >>= 32; }
    }
    mp_mul<2, 8>(n2, BN_GLV_B(), kp);
    {
        int64_t c = 0;
        for (int i = 0; i < 10; i++) { c += (int64_t)n2[i] - (int64_t)(i < 6 ? BN_GLV_2AB()[i] : 0u); n2[i] = (uint32_t)c; c >>= 32; }
    }
    BBS_A16 uint32_t q1[17], q2[13];
    mp_mul<12, 5>(q1, n1, BN_GLV_G384());
    mp_mul<10, 3>(q2, n2, BN_GLV_G320());
    const uint32_t* c1 = q1 + 12;      // 5 limbs
    const uint32_t* c2 = q2 + 10;      // 3 limbs
    BBS_A16 uint32_t c1a[9], c2b[5], c1b[7], c2ab[7];
    mp_mul<5, 4>(c1a, c1, BN_GLV_A());
    mp_mul<3, 2>(c2b, c2, BN_GLV_B());
    mp_mul<5, 2>(c1b, c1, BN_GLV_B());
    mp_mul<3, 4>(c2ab, c2, BN_GLV_AB());
    int64_t ca = 0, cb = 0;
    for (int i = 0; i < 5; i++) {      // mod 2^160: both results are below 2^128
        ca += (int64_t)kp[i] - (int64_t)c1a[i] - (int64_t)c2b[i];
        k1[i] = (uint32_t)ca;
        ca >>= 32;
        cb += (int64_t)c1b[i] - (int64_t)c2ab[i];
        k2[i] = (uint32_t)cb;
        cb >>= 32;
    }
}

template <class C> BBS_HD void glv_split(uint32_t* k1, uint32_t* k2, const uint32_t* k);
template <> BBS_HD void glv_split<Bls>(uint32_t* k1, uint32_t* k2, const uint32_t* k) { bls_glv_split(k1, k2, k); }
template <> BBS_HD void glv_split<Bn>(uint32_t* k1, uint32_t* k2, const uint32_t* k) { bn_glv_split(k1, k2, k); }

// r = sum_j k_j * P_j for NP affine points (nullptr = identity) and canonical scalars k_j (8 limbs each)
template <class C, int NP> BBS_HDN void g1_msm_scalar(uint32_t* r, const uint32_t* const* pts, const uint32_t* const* ks);
template <class C, int NP> struct G1Msm {
    static BBS_HD void run(uint32_t* r, const uint32_t* const* pts, const uint32_t* const* ks) {
        BBS_A16 uint32_t k1[NP][9], k2[NP][9];
        for (int j = 0; j < NP; j++) {
            for (int i = 0; i < 9; i++) { k1[j][i] = 0; k2[j][i] = 0; }
            glv_split<C>(k1[j], k2[j], ks[j]);
        }
        g1_msm_win4<C, NP>(r, pts, k1, k2, 128, C::GLV_BETA());   // both halves < 2^128
    }
};
template <class C, int NP> BBS_HDN void g1_msm_scalar(uint32_t* r, const uint32_t* const* pts, const uint32_t* const* ks) {
    G1Msm<C, NP>::run(r, pts, ks);
}
// r = k * P for an affine P and a canonical scalar k (8 limbs)
template <class C> BBS_HDN void g1_mul_scalar(uint32_t* r, const uint32_t* a, const uint32_t* k) {
    const uint32_t* pts[1] = {a};
    const uint32_t* ks[1] = {k};
    G1Msm<C, 1>::run(r, pts, ks);
}

// ---- encodings -------------------------------------------------------
enum : int { PT_OK = 0, PT_INF = 1, PT_BAD = 2 };

// ---- subgroup membership -----------------------------------------------------------------------------
// ark-serialize `deserialize_compressed` (Validate::Yes; derived for Signature sign.rs:18, Proof proof_gen.rs:29,
// PublicKey key_gen.rs:12) ends with `is_in_correct_subgroup_assuming_on_curve`.  BN254's G1 has cofactor 1.  On
// BLS12-381 (cofactor (x-1)^2/3) the test is the endomorphism identity of G1 [Scott, "A note on group membership tests"]:
// with phi(x, y) = (beta x, y) acting as lambda = x^2 - 1 on G1 (the GLV map above), phi^2 acts as -x^2, so
//     P in G1  <=>  [x^2] P == -phi^2(P) = (beta^2 x, -y),   beta^2 = -beta - 1.
// Sufficient because phi^2 + phi + 1 = 0 on all of E(Fp): the identity forces (x^4 - x^2 + 1) P = r P = O.
// Cost: two passes over |x| (Hamming weight 6): 126 doublings + 10 additions.  Every variable-base product of the path
// relies on it: the GLV split is only a scalar multiplication on G1.
template <class C> BBS_HDN bool g1_in_subgroup(const uint32_t* a /*affine, on the curve*/);
template <> BBS_HDN bool g1_in_subgroup<Bn>(const uint32_t*) { return true; }
template <> BBS_HDN bool g1_in_subgroup<Bls>(const uint32_t* a) {
    using C = Bls;
    using F = BlsFp;
    BBS_A16 uint32_t acc[G1J], base[G1J];
    g1_from_affine<C>(acc, a);
    for (int i = 62; i >= 0; i--) {                      // |x| P
        g1_dbl<C>(acc, acc);
        if ((BLS_X_ABS >> i) & 1) g1_add_mixed<C>(acc, acc, a);
    }
    g1_copy<C>(base, acc);
    for (int i = 62; i >= 0; i--) {                      // x^2 P
        g1_dbl<C>(acc, acc);
        if ((BLS_X_ABS >> i) & 1) g1_add<C>(acc, acc, base);
    }
    if (g1_is_inf_ool<C>(acc)) return false;
    // (X, Y, Z) == (beta^2 x, -y)  <=>  X + (beta x + x) Z^2 == 0  and  Y + y Z^3 == 0
    BBS_A16 uint32_t zz[FPN], t[FPN], s[FPN];
    fe_sqr<F>(zz, acc + 2 * FPN);
    fe_mul<F>(t, a, C::GLV_BETA()); fe_add<F>(t, t, a); fe_mul<F>(t, t, zz); fe_add<F>(t, t, acc);
    fe_mul<F>(zz, zz, acc + 2 * FPN); fe_mul<F>(s, a + FPN, zz); fe_add<F>(s, s, acc + FPN);
    BBS_OPAQUE_CALL_BARRIER();
    const bool ok = bn_is_zero<12>(t) && bn_is_zero<12>(s);
    BBS_OPAQUE_CALL_BARRIER();
    return ok;
}

// compressed bytes -> affine Montgomery point.  PT_INF for the identity encoding, PT_BAD for anything
// ark's deserializer would refuse (bad flags, x >= p, x not on the curve, point outside the prime-order subgroup).
template <class C> BBS_HDN int g1_decompress(uint32_t* r /*affine*/, const uint8_t* in);
template <class C> BBS_HDN void g1_compress_affine(uint8_t* out, const uint32_t* a /*affine*/, bool inf);

template <class C> BBS_HDN int g1_finish_decompress(uint32_t* r, uint32_t* x_canon, bool want_high) {
    using F = typename C::Fp;
    if (!fe_is_canonical<F>(x_canon)) return PT_BAD;
    BBS_A16 uint32_t x[FPN], rhs[FPN], y[FPN];
    fe_to_mont<F>(x, x_canon);
    fe_sqr<F>(rhs, x); fe_mul<F>(rhs, rhs, x); fe_add<F>(rhs, rhs, C::B());
    if (!fe_sqrt<F>(y, rhs)) return PT_BAD;
    if (fe_is_high<F>(y) != want_high) fe_neg<F>(y, y);
    bn_copy<C::Fp::N>(r, x); bn_copy<C::Fp::N>(r + FPN, y);
    return g1_in_subgroup<C>(r) ? PT_OK : PT_BAD;
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
    BBS_A16 uint32_t x[12];
    limbs_from_be<12>(x, tmp);
    return g1_finish_decompress<Bls>(r, x, (b0 & 0x20) != 0);
}
template <> BBS_HDN void g1_compress_affine<Bls>(uint8_t* out, const uint32_t* a, bool inf) {
    if (inf) { out[0] = 0xc0; for (int i = 1; i < 48; i++) out[i] = 0; return; }
    BBS_A16 uint32_t x[12];
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
    BBS_A16 uint32_t x[8];
    limbs_from_le<8>(x, tmp);
    return g1_finish_decompress<Bn>(r, x, (fl & 0x80) != 0);
}
template <> BBS_HDN void g1_compress_affine<Bn>(uint8_t* out, const uint32_t* a, bool inf) {
    if (inf) { for (int i = 0; i < 32; i++) out[i] = 0; out[31] = 0x40; return; }
    BBS_A16 uint32_t x[8];
    fe_from_mont<BnFp>(x, a);
    limbs_to_le<8>(out, x);
    if (fe_is_high<BnFp>(a + 8)) out[31] |= 0x80;
}

// Jacobian -> compressed bytes (one inversion)
template <class C> BBS_HDN void g1_compress(uint8_t* out, const uint32_t* p) {
    BBS_A16 uint32_t a[G1A];
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
