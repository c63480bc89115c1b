// Optimal-ate pairing pieces for the BBS verdict  e(P0, Q0) * e(P1, Q1) == 1  with BOTH G2 arguments
// fixed per issuer (Q0 = W, Q1 = BP2; SURVEY 8a note (i)):
//   * g2_precompute_lines : walks the ate loop once per G2 point on the twist (affine, Fp2) and stores
//                           the line coefficients (A = lambda*xT - yT, Bc = -lambda) of every step;
//   * miller2             : two-pair shared-squaring Miller loop that only evaluates stored lines;
//   * final_exp_is_one    : one final exponentiation (easy part + cyclotomic hard part) and == 1.
// Replaces the two `E::pairing` calls of verify.rs:88-92 and proof_verify.rs:112-115 (ark-ec bls12 / bn
// models).  GT values are unobservable in the reference, so factors that die in the final exponentiation
// (the w^3 line scaling, Z^3 of a projective P, the x<0 conjugation) are dropped.
#pragma once
#include "g1.cuh"

namespace bbs {

#define LINE_WORDS (4 * C::Fp::N)   // A (Fp2) | Bc (Fp2)

// ---- ate loop schedule ----------------------------------------------------------------------------
template <class C> struct Ate;
template <> struct Ate<Bls> {
    static constexpr int STEPS = 63;          // bits 62..0 of |x|, below the leading one
    static constexpr int TAIL = 0;
    static BBS_HD int digit(int i) { return (int)((BLS_X_ABS >> i) & 1); }
    static constexpr int LINES = 63 + 5;      // |x| has Hamming weight 6
};
template <> struct Ate<Bn> {
    static constexpr int STEPS = BN_ATE_NAF_LEN - 1;
    static constexpr int TAIL = 2;            // pi(Q), -pi^2(Q)
    static BBS_HD int digit(int i) { return (int)BN_ATE_NAF()[i] - 1; }
    static constexpr int LINES_MAX = 2 * BN_ATE_NAF_LEN + 2;
};

template <class C> BBS_HD int ate_line_count() {
    int n = Ate<C>::STEPS + Ate<C>::TAIL;
    for (int i = 0; i < Ate<C>::STEPS; i++) n += Ate<C>::digit(i) != 0;
    return n;
}

// ---- G2 affine steps on the twist, emitting line coefficients ------------------------------------
// T <- 2T ; line: lambda = 3 xT^2 / (2 yT)
template <class C> BBS_HDN void g2_dbl_step(uint32_t* line, uint32_t* T) {
    BBS_A16 uint32_t lam[F2N], t[F2N], x3[F2N], y3[F2N];
    f2_sqr<C>(lam, T); f2_dbl<C>(t, lam); f2_add<C>(lam, lam, t);
    f2_dbl<C>(t, T + F2N); f2_inv_vt<C>(t, t);
    f2_mul<C>(lam, lam, t);
    f2_mul<C>(line, lam, T); f2_sub<C>(line, line, T + F2N);       // A = lambda xT - yT
    f2_neg<C>(line + F2N, lam);                                    // Bc = -lambda
    f2_sqr<C>(x3, lam); f2_sub<C>(x3, x3, T); f2_sub<C>(x3, x3, T);
    f2_sub<C>(t, T, x3); f2_mul<C>(y3, lam, t); f2_sub<C>(y3, y3, T + F2N);
    f2_copy<C>(T, x3); f2_copy<C>(T + F2N, y3);
}
// T <- T + Q ; line through T and Q
template <class C> BBS_HDN void g2_add_step(uint32_t* line, uint32_t* T, const uint32_t* Q) {
    BBS_A16 uint32_t lam[F2N], t[F2N], x3[F2N], y3[F2N];
    f2_sub<C>(lam, Q + F2N, T + F2N);
    f2_sub<C>(t, Q, T); f2_inv_vt<C>(t, t);
    f2_mul<C>(lam, lam, t);
    f2_mul<C>(line, lam, T); f2_sub<C>(line, line, T + F2N);
    f2_neg<C>(line + F2N, lam);
    f2_sqr<C>(x3, lam); f2_sub<C>(x3, x3, T); f2_sub<C>(x3, x3, Q);
    f2_sub<C>(t, T, x3); f2_mul<C>(y3, lam, t); f2_sub<C>(y3, y3, T + F2N);
    f2_copy<C>(T, x3); f2_copy<C>(T + F2N, y3);
}

// Q: affine twist point [x(Fp2)|y(Fp2)], must not be the identity and must have order r.
// out: line k of pair `pair` lives at out + (2k + pair) * LINE_WORDS.
template <class C> BBS_HDN void g2_precompute_lines(uint32_t* out, const uint32_t* Q, int pair) {
    BBS_A16 uint32_t T[2 * F2N], nQ[2 * F2N];
    bn_copy<4 * C::Fp::N>(T, Q);
    f2_copy<C>(nQ, Q); f2_neg<C>(nQ + F2N, Q + F2N);
    int k = 0;
    for (int i = Ate<C>::STEPS - 1; i >= 0; i--) {
        g2_dbl_step<C>(out + (2 * k + pair) * LINE_WORDS, T); k++;
        int d = Ate<C>::digit(i);
        if (d) { g2_add_step<C>(out + (2 * k + pair) * LINE_WORDS, T, d > 0 ? Q : nQ); k++; }
    }
    if (Ate<C>::TAIL) {
        // Q1 = pi(Q) = (conj(x) g1[2], conj(y) g1[3]);  Q2 = -pi^2(Q) = (x g2[2], -y g2[3])  (D-type twist)
        BBS_A16 uint32_t Q1[2 * F2N], Q2[2 * F2N], t[F2N];
        const uint32_t *g1 = C::FROB(1), *g2 = C::FROB(2);
        f2_conj<C>(t, Q); f2_mul<C>(Q1, t, g1 + 2 * F2N);
        f2_conj<C>(t, Q + F2N); f2_mul<C>(Q1 + F2N, t, g1 + 3 * F2N);
        f2_mul<C>(Q2, Q, g2 + 2 * F2N);
        f2_mul<C>(t, Q + F2N, g2 + 3 * F2N); f2_neg<C>(Q2 + F2N, t);
        g2_add_step<C>(out + (2 * k + pair) * LINE_WORDS, T, Q1); k++;
        g2_add_step<C>(out + (2 * k + pair) * LINE_WORDS, T, Q2); k++;
    }
}

// ---- line evaluation --------------------------------------------------------------------------------
// P is given as (px, py, pz) = (X*Z, Y, Z^3) of a Jacobian point (affine: (x, y, 1)); the evaluated line is
// scaled by Z^3 in Fp, which the final exponentiation kills.
template <class C> BBS_HDN void f12_mul_line(uint32_t* f, const uint32_t* line, const uint32_t* P) {
    BBS_A16 uint32_t a[F2N], b[F2N], y[F2N];
    f2_mul_fp<C>(a, line, P + 2 * FPN);          // A * pz
    f2_mul_fp<C>(b, line + F2N, P);              // Bc * px
    bn_copy<C::Fp::N>(y, P + FPN); bn_zero<C::Fp::N>(y + FPN);
    if (C::M_TWIST) f12_mul_by_014<C>(f, a, b, y);   // A + (Bc px) v + py (v w)
    else            f12_mul_by_034<C>(f, y, b, a);   // py + (Bc px) w + A (v w)
}

// f = f_{Q0}(P0) * f_{Q1}(P1) over the shared loop; a skipped pair contributes 1 (ark-ec filters pairs
// with an identity argument: SURVEY 4 / Appendix C).
template <class C> BBS_HDN void miller2(uint32_t* f, const uint32_t* lines, const uint32_t* P0, bool skip0,
                                       const uint32_t* P1, bool skip1) {
    f12_one<C>(f);
    int k = 0;
    bool first = true;
    for (int i = Ate<C>::STEPS - 1; i >= 0; i--) {
        if (!first) f12_sqr<C>(f, f);
        first = false;
        if (!skip0) f12_mul_line<C>(f, lines + (2 * k) * LINE_WORDS, P0);
        if (!skip1) f12_mul_line<C>(f, lines + (2 * k + 1) * LINE_WORDS, P1);
        k++;
        if (Ate<C>::digit(i)) {
            if (!skip0) f12_mul_line<C>(f, lines + (2 * k) * LINE_WORDS, P0);
            if (!skip1) f12_mul_line<C>(f, lines + (2 * k + 1) * LINE_WORDS, P1);
            k++;
        }
    }
    for (int j = 0; j < Ate<C>::TAIL; j++) {
        if (!skip0) f12_mul_line<C>(f, lines + (2 * k) * LINE_WORDS, P0);
        if (!skip1) f12_mul_line<C>(f, lines + (2 * k + 1) * LINE_WORDS, P1);
        k++;
    }
}

// ---- final exponentiation ---------------------------------------------------------------------------
// r = a^e for a in the cyclotomic subgroup, e > 0 (a public constant)
template <class C> BBS_HDN void f12_cyc_pow(uint32_t* r, const uint32_t* a, uint64_t e) {
    BBS_A16 uint32_t acc[F12N];
    f12_copy<C>(acc, a);
    int top = 63;
    while (!((e >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        f12_cyc_sqr<C>(acc, acc);
        if ((e >> i) & 1) f12_mul<C>(acc, acc, a);
    }
    f12_copy<C>(r, acc);
}

template <class C> BBS_HDN void final_exp_hard(uint32_t* r, const uint32_t* f);

// BLS12 (x < 0): 3 (p^4-p^2+1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3   [Hayashida-Hayasaka-Teruya, eprint 2020/875]
// f^x = conj(f^|x|) in the cyclotomic subgroup.
template <> BBS_HDN void final_exp_hard<Bls>(uint32_t* r, const uint32_t* f) {
    using C = Bls;
    BBS_A16 uint32_t a[F12N], b[F12N], c[F12N];
    f12_cyc_pow<C>(a, f, BLS_X_ABS); f12_conj<C>(a, a); f12_conj<C>(b, f); f12_mul<C>(a, a, b);      // f^(x-1)
    f12_cyc_pow<C>(b, a, BLS_X_ABS); f12_conj<C>(b, b); f12_conj<C>(c, a); f12_mul<C>(a, b, c);      // ^(x-1)
    f12_cyc_pow<C>(b, a, BLS_X_ABS); f12_conj<C>(b, b); f12_frob<C>(c, a, 1); f12_mul<C>(a, b, c);   // ^(x+p)
    f12_cyc_pow<C>(b, a, BLS_X_ABS); f12_conj<C>(b, b);
    f12_cyc_pow<C>(b, b, BLS_X_ABS); f12_conj<C>(b, b);                                              // a^(x^2)
    f12_frob<C>(c, a, 2); f12_mul<C>(b, b, c); f12_conj<C>(c, a); f12_mul<C>(a, b, c);               // ^(x^2+p^2-1)
    f12_cyc_sqr<C>(b, f); f12_mul<C>(b, b, f);                                                       // f^3
    f12_mul<C>(r, a, b);
}

// BN (x = t > 0): Fuentes-Castaneda et al. addition chain (the one ark-ec's bn model uses)
template <> BBS_HDN void final_exp_hard<Bn>(uint32_t* r, const uint32_t* f) {
    using C = Bn;
    BBS_A16 uint32_t y0[F12N], y1[F12N], y3[F12N], y4[F12N], y6[F12N], y8[F12N], y9[F12N], t[F12N], u[F12N];
    f12_cyc_pow<C>(y0, f, BN_T); f12_conj<C>(y0, y0);        // f^-x
    f12_cyc_sqr<C>(y1, y0);                                  // y1
    f12_cyc_sqr<C>(t, y1);                                   // y2
    f12_mul<C>(y3, t, y1);                                   // y3
    f12_cyc_pow<C>(y4, y3, BN_T); f12_conj<C>(y4, y4);       // y4
    f12_cyc_sqr<C>(t, y4);                                   // y5
    f12_cyc_pow<C>(y6, t, BN_T); f12_conj<C>(y6, y6);        // y6
    f12_conj<C>(y3, y3); f12_conj<C>(y6, y6);
    f12_mul<C>(t, y6, y4);                                   // y7
    f12_mul<C>(y8, t, y3);                                   // y8
    f12_mul<C>(y9, y8, y1);                                  // y9
    f12_mul<C>(t, y8, y4);                                   // y10
    f12_mul<C>(t, t, f);                                     // y11
    f12_frob<C>(u, y9, 1);                                   // y12
    f12_mul<C>(t, u, t);                                     // y13
    f12_frob<C>(y8, y8, 2);
    f12_mul<C>(t, y8, t);                                    // y14
    f12_conj<C>(u, f); f12_mul<C>(u, u, y9); f12_frob<C>(u, u, 3);   // y15
    f12_mul<C>(r, u, t);                                     // y16
}

// f^((p^12-1)/r * c) == 1 ?   (c = 3 for BLS12, 1 for BN; gcd(c, r) = 1 so the verdict is unchanged)
template <class C> BBS_HDN bool final_exp_is_one(const uint32_t* f) {
    BBS_A16 uint32_t a[F12N], b[F12N];
    f12_inv<C>(a, f);
    f12_conj<C>(b, f);
    f12_mul<C>(a, a, b);            // f^(p^6-1)
    f12_frob<C>(b, a, 2);
    f12_mul<C>(a, a, b);            // ^(p^2+1)
    final_exp_hard<C>(b, a);
    return f12_is_one<C>(b);
}

}  // namespace bbs
