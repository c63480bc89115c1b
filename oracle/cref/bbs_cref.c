/* bbs_cref.c -- CPU restatement in plain C of the reference's per-item hot path for BLS12-381
 * (TEST / BENCH INFRASTRUCTURE: the oracle's fast leg; never linked into or called by the product).
 *
 * Follows hashcloak/bbs_sign as it executes on arkworks 0.4.x, per item and with no hoisting:
 *   msg_to_scalars            src/utils/interface_utilities.rs:76-88  (expand_message_xmd utilities_helper.rs:42-97)
 *   calculate_domain          src/utils/core_utilities.rs:24-63       (L+2 point compressions + hash, every call)
 *   B = P1 + Q1 d + sum H m   src/verify.rs:81-86                     (MSB-first double-and-add, Jacobian)
 *   pk + BP2 * e              src/verify.rs:89                        (G2 double-and-add)
 *   E::pairing x 2, product   src/verify.rs:88-92                     (two Miller loops, two final exponentiations)
 * Third-party arithmetic (ark-ff / ark-ec / ark-bls12-381 / sha2, not under /root/reference) is restated from
 * the published algorithms: 6x64-bit Montgomery, Fp2/Fp6/Fp12 Karatsuba tower, Jacobian G1/G2, the optimal-ate
 * Miller loop with homogeneous projective line steps (Costello-Lange-Naehrig as in ark-ec bls12), final
 * exponentiation with cyclotomic squarings, zcash point encoding.  Parity: pinned by the IRTF signature
 * fixture (test_vector.rs:164-192) and by agreement with oracle/bbs_oracle.py (tests/test_cref.py).
 *
 * Build: make -C oracle/cref   ->  oracle/_ref/libbbs_cref.so
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 l[6]; } fp;
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;
typedef struct { u64 l[4]; } fr;

/* ---- Fp: 6x64 Montgomery ---- */
static const u64 P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL,
                         0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 PINV = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */
static fp FP_ONE, FP_R2, FP_ZERO;

static int ge6(const u64* a, const u64* b) { for (int i = 5; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; } return 1; }
static void sub6(u64* r, const u64* a, const u64* b) { u128 br = 0; for (int i = 0; i < 6; i++) { u128 t = (u128)a[i] - b[i] - (u64)br; r[i] = (u64)t; br = (t >> 64) & 1; } }
static u64 add6(u64* r, const u64* a, const u64* b) { u128 c = 0; for (int i = 0; i < 6; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; } return (u64)c; }
static void fp_add(fp* r, const fp* a, const fp* b) { add6(r->l, a->l, b->l); if (ge6(r->l, P)) sub6(r->l, r->l, P); }
static void fp_sub(fp* r, const fp* a, const fp* b) { if (ge6(a->l, b->l)) sub6(r->l, a->l, b->l); else { u64 t[6]; sub6(t, b->l, a->l); sub6(r->l, P, t); } }
static int fp_is_zero(const fp* a) { u64 o = 0; for (int i = 0; i < 6; i++) o |= a->l[i]; return o == 0; }
static int fp_eq(const fp* a, const fp* b) { return memcmp(a, b, sizeof(fp)) == 0; }
static void fp_neg(fp* r, const fp* a) { if (fp_is_zero(a)) *r = *a; else sub6(r->l, P, a->l); }
static void fp_dbl(fp* r, const fp* a) { fp_add(r, a, a); }
static void fp_mul(fp* r, const fp* a, const fp* b) {
    u64 t[8] = {0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[6]; t[6] = (u64)c; t[7] = (u64)(c >> 64);
        u64 m = t[0] * PINV;
        c = ((u128)m * P[0] + t[0]) >> 64;
        for (int j = 1; j < 6; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[6]; t[5] = (u64)c; t[6] = t[7] + (u64)(c >> 64);
    }
    if (t[6] || ge6(t, P)) sub6(r->l, t, P); else memcpy(r->l, t, 48);
}
static void fp_sqr(fp* r, const fp* a) { fp_mul(r, a, a); }
static void fp_pow(fp* r, const fp* a, const u64* e, int nlimbs) {
    fp acc = FP_ONE; int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) { if (started) fp_mul(&acc, &acc, a); else { acc = *a; started = 1; } }
    }
    *r = acc;
}
static void fp_inv(fp* r, const fp* a) { u64 e[6]; memcpy(e, P, 48); e[0] -= 2; fp_pow(r, a, e, 6); }
static void fp_from_u64(fp* r, u64 v) { fp t = FP_ZERO; t.l[0] = v; fp_mul(r, &t, &FP_R2); }
static void fp_to_canon(u64* out, const fp* a) { fp one = FP_ZERO; one.l[0] = 1; fp t; fp_mul(&t, a, &one); memcpy(out, t.l, 48); }
static void fp_from_be48(fp* r, const uint8_t* b) { fp t; for (int i = 0; i < 6; i++) { u64 v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | b[8 * (5 - i) + k]; t.l[i] = v; } fp_mul(r, &t, &FP_R2); }
static void fp_to_be48(uint8_t* b, const fp* a) { u64 c[6]; fp_to_canon(c, a); for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[8 * (5 - i) + k] = (uint8_t)(c[i] >> (56 - 8 * k)); }
static int fp_is_high(const fp* a) { /* canonical(a) > (p-1)/2 */
    u64 c[6], h[6]; fp_to_canon(c, a);
    u64 carry = 0; for (int i = 5; i >= 0; i--) { u64 v = P[i]; h[i] = (v >> 1) | (carry << 63); carry = v & 1; } /* (p-1)/2 = p>>1 */
    for (int i = 5; i >= 0; i--) { if (c[i] != h[i]) return c[i] > h[i]; } return 0;
}

/* ---- Fr: 4x64, only what the path needs (mod-r reduction of 48 bytes, add) ---- */
static const u64 RMOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static int ge4(const u64* a, const u64* b) { for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; } return 1; }
static void sub4(u64* r, const u64* a, const u64* b) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - (u64)br; r[i] = (u64)t; br = (t >> 64) & 1; } }
/* x (48 bytes big-endian) mod r by shift-and-subtract on 384 bits (from_okm, utilities_helper.rs:30-40) */
static void fr_from_okm(fr* out, const uint8_t* be48) {
    u64 acc[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 48; i++) {
        for (int bit = 7; bit >= 0; bit--) {
            u64 top = acc[3] >> 63;
            for (int k = 3; k > 0; k--) acc[k] = (acc[k] << 1) | (acc[k - 1] >> 63);
            acc[0] = (acc[0] << 1) | ((be48[i] >> bit) & 1);
            if (top || ge4(acc, RMOD)) sub4(acc, acc, RMOD);
        }
    }
    memcpy(out->l, acc, 32);
}
static void fr_to_be32(uint8_t* b, const fr* a) { for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[8 * (3 - i) + k] = (uint8_t)(a->l[i] >> (56 - 8 * k)); }
static void fr_from_le32(fr* r, const uint8_t* b) { for (int i = 0; i < 4; i++) { u64 v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | b[8 * i + k]; r->l[i] = v; } }

/* ---- SHA-256 + expand_message_xmd ---- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
typedef struct { uint32_t h[8]; uint8_t buf[64]; size_t fill; u64 total; } sha256;
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha_block(sha256* s, const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) { uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10); w[i] = w[i - 16] + s0 + w[i - 7] + s1; }
    uint32_t a = s->h[0], b = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4], f = s->h[5], g = s->h[6], h = s->h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha_init(sha256* s) { static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19}; memcpy(s->h, iv, 32); s->fill = 0; s->total = 0; }
static void sha_update(sha256* s, const uint8_t* p, size_t n) {
    s->total += n;
    while (n) { size_t k = 64 - s->fill; if (k > n) k = n; memcpy(s->buf + s->fill, p, k); s->fill += k; p += k; n -= k; if (s->fill == 64) { sha_block(s, s->buf); s->fill = 0; } }
}
static void sha_final(sha256* s, uint8_t* out) {
    u64 bits = s->total * 8; uint8_t pad[72] = {0x80}; size_t k = (s->fill < 56) ? 56 - s->fill : 120 - s->fill;
    sha_update(s, pad, k); uint8_t len[8]; for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i)); sha_update(s, len, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = s->h[i] >> 24; out[4 * i + 1] = s->h[i] >> 16; out[4 * i + 2] = s->h[i] >> 8; out[4 * i + 3] = s->h[i]; }
}
/* expand_message_xmd(msg, dst, 48) (utilities_helper.rs:42-97); msg given as two pieces to avoid copies */
static void xmd48(uint8_t out[48], const uint8_t* m1, size_t n1, const uint8_t* m2, size_t n2, const uint8_t* dst, size_t dlen) {
    uint8_t z[64] = {0}, b0[32], b1[32], b2[32], t[33]; uint8_t dl = (uint8_t)dlen; uint8_t lib[3] = {0, 48, 0};
    sha256 s; sha_init(&s); sha_update(&s, z, 64); sha_update(&s, m1, n1); sha_update(&s, m2, n2); sha_update(&s, lib, 3); sha_update(&s, dst, dlen); sha_update(&s, &dl, 1); sha_final(&s, b0);
    memcpy(t, b0, 32); t[32] = 1; sha_init(&s); sha_update(&s, t, 33); sha_update(&s, dst, dlen); sha_update(&s, &dl, 1); sha_final(&s, b1);
    for (int i = 0; i < 32; i++) t[i] = b0[i] ^ b1[i]; t[32] = 2; sha_init(&s); sha_update(&s, t, 33); sha_update(&s, dst, dlen); sha_update(&s, &dl, 1); sha_final(&s, b2);
    memcpy(out, b1, 32); memcpy(out + 32, b2, 16);
}
static void hash_to_scalar(fr* r, const uint8_t* msg, size_t n, const uint8_t* dst, size_t dlen) { uint8_t okm[48]; xmd48(okm, msg, n, NULL, 0, dst, dlen); fr_from_okm(r, okm); }

/* expand_message_xmd for any output length <= 8160 (utilities_helper.rs:42-97) */
static void xmd(uint8_t* out, size_t len, const uint8_t* m1, size_t n1, const uint8_t* m2, size_t n2, const uint8_t* dst, size_t dlen) {
    uint8_t z[64] = {0}, b0[32], bi[32], t[33]; uint8_t dl = (uint8_t)dlen; uint8_t lib[3] = {(uint8_t)(len >> 8), (uint8_t)len, 0};
    size_t ell = (len + 31) / 32;
    sha256 s; sha_init(&s); sha_update(&s, z, 64); sha_update(&s, m1, n1); sha_update(&s, m2, n2); sha_update(&s, lib, 3); sha_update(&s, dst, dlen); sha_update(&s, &dl, 1); sha_final(&s, b0);
    memset(bi, 0, 32);
    for (size_t i = 1; i <= ell; i++) {
        for (int k = 0; k < 32; k++) t[k] = (i == 1) ? b0[k] : (uint8_t)(b0[k] ^ bi[k]);
        t[32] = (uint8_t)i; sha_init(&s); sha_update(&s, t, 33); sha_update(&s, dst, dlen); sha_update(&s, &dl, 1); sha_final(&s, bi);
        size_t off = 32 * (i - 1), take = len - off < 32 ? len - off : 32; memcpy(out + off, bi, take);
    }
}

/* ---- Fp2 ---- */
static void f2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static void f2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static void f2_neg(fp2* r, const fp2* a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static void f2_dbl(fp2* r, const fp2* a) { f2_add(r, a, a); }
static void f2_conj(fp2* r, const fp2* a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
static void f2_mul(fp2* r, const fp2* a, const fp2* b) {
    fp t0, t1, s0, s1, t2; fp_mul(&t0, &a->c0, &b->c0); fp_mul(&t1, &a->c1, &b->c1); fp_add(&s0, &a->c0, &a->c1); fp_add(&s1, &b->c0, &b->c1); fp_mul(&t2, &s0, &s1);
    fp_sub(&r->c0, &t0, &t1); fp_sub(&t2, &t2, &t0); fp_sub(&r->c1, &t2, &t1);
}
static void f2_sqr(fp2* r, const fp2* a) { fp s, d, m; fp_add(&s, &a->c0, &a->c1); fp_sub(&d, &a->c0, &a->c1); fp_mul(&m, &a->c0, &a->c1); fp_mul(&r->c0, &s, &d); fp_dbl(&r->c1, &m); }
static void f2_mul_fp(fp2* r, const fp2* a, const fp* k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); }
static void f2_mul_xi(fp2* r, const fp2* a) { fp t0, t1; fp_sub(&t0, &a->c0, &a->c1); fp_add(&t1, &a->c0, &a->c1); r->c0 = t0; r->c1 = t1; }
static void f2_inv(fp2* r, const fp2* a) { fp n, t; fp_sqr(&n, &a->c0); fp_sqr(&t, &a->c1); fp_add(&n, &n, &t); fp_inv(&n, &n); fp_mul(&r->c0, &a->c0, &n); fp_mul(&t, &a->c1, &n); fp_neg(&r->c1, &t); }
static int f2_is_zero(const fp2* a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static fp2 F2_ZERO, F2_ONE;

/* ---- Fp6 / Fp12 ---- */
static void f6_add(fp6* r, const fp6* a, const fp6* b) { f2_add(&r->c0, &a->c0, &b->c0); f2_add(&r->c1, &a->c1, &b->c1); f2_add(&r->c2, &a->c2, &b->c2); }
static void f6_sub(fp6* r, const fp6* a, const fp6* b) { f2_sub(&r->c0, &a->c0, &b->c0); f2_sub(&r->c1, &a->c1, &b->c1); f2_sub(&r->c2, &a->c2, &b->c2); }
static void f6_neg(fp6* r, const fp6* a) { f2_neg(&r->c0, &a->c0); f2_neg(&r->c1, &a->c1); f2_neg(&r->c2, &a->c2); }
static void f6_mul_v(fp6* r, const fp6* a) { fp2 t; f2_mul_xi(&t, &a->c2); fp2 a0 = a->c0, a1 = a->c1; r->c0 = t; r->c1 = a0; r->c2 = a1; }
static void f6_mul(fp6* r, const fp6* a, const fp6* b) {
    fp2 v0, v1, v2, s, t, c0, c1, c2;
    f2_mul(&v0, &a->c0, &b->c0); f2_mul(&v1, &a->c1, &b->c1); f2_mul(&v2, &a->c2, &b->c2);
    f2_add(&s, &a->c1, &a->c2); f2_add(&t, &b->c1, &b->c2); f2_mul(&c0, &s, &t); f2_sub(&c0, &c0, &v1); f2_sub(&c0, &c0, &v2); f2_mul_xi(&c0, &c0); f2_add(&c0, &c0, &v0);
    f2_add(&s, &a->c0, &a->c1); f2_add(&t, &b->c0, &b->c1); f2_mul(&c1, &s, &t); f2_sub(&c1, &c1, &v0); f2_sub(&c1, &c1, &v1); f2_mul_xi(&t, &v2); f2_add(&c1, &c1, &t);
    f2_add(&s, &a->c0, &a->c2); f2_add(&t, &b->c0, &b->c2); f2_mul(&c2, &s, &t); f2_sub(&c2, &c2, &v0); f2_sub(&c2, &c2, &v2); f2_add(&c2, &c2, &v1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void f6_mul_by_01(fp6* r, const fp6* a, const fp2* b0, const fp2* b1) {
    fp2 aa, bb, s, t, c0, c1, c2;
    f2_mul(&aa, &a->c0, b0); f2_mul(&bb, &a->c1, b1);
    f2_add(&s, &a->c1, &a->c2); f2_mul(&c0, &s, b1); f2_sub(&c0, &c0, &bb); f2_mul_xi(&c0, &c0); f2_add(&c0, &c0, &aa);
    f2_add(&s, &a->c0, &a->c2); f2_mul(&c2, &s, b0); f2_sub(&c2, &c2, &aa); f2_add(&c2, &c2, &bb);
    f2_add(&s, &a->c0, &a->c1); f2_add(&t, b0, b1); f2_mul(&c1, &s, &t); f2_sub(&c1, &c1, &aa); f2_sub(&c1, &c1, &bb);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void f6_mul_by_1(fp6* r, const fp6* a, const fp2* b1) { fp2 c0, c1, c2; f2_mul(&c0, &a->c2, b1); f2_mul_xi(&c0, &c0); f2_mul(&c1, &a->c0, b1); f2_mul(&c2, &a->c1, b1); r->c0 = c0; r->c1 = c1; r->c2 = c2; }
static void f6_inv(fp6* r, const fp6* a) {
    fp2 c0, c1, c2, t, d;
    f2_sqr(&c0, &a->c0); f2_mul(&t, &a->c1, &a->c2); f2_mul_xi(&t, &t); f2_sub(&c0, &c0, &t);
    f2_sqr(&c1, &a->c2); f2_mul_xi(&c1, &c1); f2_mul(&t, &a->c0, &a->c1); f2_sub(&c1, &c1, &t);
    f2_sqr(&c2, &a->c1); f2_mul(&t, &a->c0, &a->c2); f2_sub(&c2, &c2, &t);
    f2_mul(&d, &a->c2, &c1); f2_mul(&t, &a->c1, &c2); f2_add(&d, &d, &t); f2_mul_xi(&d, &d); f2_mul(&t, &a->c0, &c0); f2_add(&d, &d, &t);
    f2_inv(&d, &d); f2_mul(&r->c0, &c0, &d); f2_mul(&r->c1, &c1, &d); f2_mul(&r->c2, &c2, &d);
}
static fp12 F12_ONE;
static void f12_mul(fp12* r, const fp12* a, const fp12* b) {
    fp6 aa, bb, s, t; f6_mul(&aa, &a->c0, &b->c0); f6_mul(&bb, &a->c1, &b->c1); f6_add(&s, &a->c0, &a->c1); f6_add(&t, &b->c0, &b->c1); f6_mul(&s, &s, &t);
    f6_sub(&s, &s, &aa); f6_sub(&r->c1, &s, &bb); f6_mul_v(&bb, &bb); f6_add(&r->c0, &aa, &bb);
}
static void f12_sqr(fp12* r, const fp12* a) {
    fp6 ab, s, t; f6_mul(&ab, &a->c0, &a->c1); f6_add(&s, &a->c0, &a->c1); f6_mul_v(&t, &a->c1); f6_add(&t, &t, &a->c0); f6_mul(&s, &s, &t);
    f6_sub(&s, &s, &ab); f6_mul_v(&t, &ab); f6_sub(&r->c0, &s, &t); f6_add(&r->c1, &ab, &ab);
}
static void f12_conj(fp12* r, const fp12* a) { r->c0 = a->c0; f6_neg(&r->c1, &a->c1); }
static void f12_inv(fp12* r, const fp12* a) { fp6 t0, t1; f6_mul(&t0, &a->c0, &a->c0); f6_mul(&t1, &a->c1, &a->c1); f6_mul_v(&t1, &t1); f6_sub(&t0, &t0, &t1); f6_inv(&t0, &t0); f6_mul(&r->c0, &a->c0, &t0); f6_mul(&t1, &a->c1, &t0); f6_neg(&r->c1, &t1); }
static void f12_mul_by_014(fp12* f, const fp2* c0, const fp2* c1, const fp2* c4) {
    fp6 aa, bb, s; fp2 o; f6_mul_by_01(&aa, &f->c0, c0, c1); f6_mul_by_1(&bb, &f->c1, c4); f2_add(&o, c1, c4); f6_add(&s, &f->c0, &f->c1); f6_mul_by_01(&s, &s, c0, &o);
    f6_sub(&s, &s, &aa); f6_sub(&f->c1, &s, &bb); f6_mul_v(&bb, &bb); f6_add(&f->c0, &aa, &bb);
}
/* Frobenius coefficients xi^(k (p^j - 1)/6), computed at init */
static fp2 FROB[3][6];
static void f12_frob(fp12* r, const fp12* a, int j) {
    const fp2* in[6] = {&a->c0.c0, &a->c0.c1, &a->c0.c2, &a->c1.c0, &a->c1.c1, &a->c1.c2};
    fp2* out[6] = {&r->c0.c0, &r->c0.c1, &r->c0.c2, &r->c1.c0, &r->c1.c1, &r->c1.c2};
    static const int wp[6] = {0, 2, 4, 1, 3, 5};
    for (int s = 0; s < 6; s++) { fp2 t; if (j & 1) f2_conj(&t, in[s]); else t = *in[s]; f2_mul(out[s], &t, &FROB[j - 1][wp[s]]); }
}
static void f12_cyc_sqr(fp12* r, const fp12* a) {
    const fp2 *z0 = &a->c0.c0, *z4 = &a->c0.c1, *z3 = &a->c0.c2, *z2 = &a->c1.c0, *z1 = &a->c1.c1, *z5 = &a->c1.c2;
    fp2 t0, t1, t2, t3, t4, t5, tmp, s, u; fp12 o;
#define FP4SQ(x, y, lo, hi) f2_mul(&tmp, x, y); f2_add(&s, x, y); f2_mul_xi(&u, y); f2_add(&u, &u, x); f2_mul(&lo, &s, &u); f2_sub(&lo, &lo, &tmp); f2_mul_xi(&u, &tmp); f2_sub(&lo, &lo, &u); f2_dbl(&hi, &tmp);
    FP4SQ(z0, z1, t0, t1) FP4SQ(z2, z3, t2, t3) FP4SQ(z4, z5, t4, t5)
    f2_sub(&s, &t0, z0); f2_dbl(&s, &s); f2_add(&o.c0.c0, &s, &t0);
    f2_add(&s, &t1, z1); f2_dbl(&s, &s); f2_add(&o.c1.c1, &s, &t1);
    f2_mul_xi(&tmp, &t5); f2_add(&s, &tmp, z2); f2_dbl(&s, &s); f2_add(&o.c1.c0, &s, &tmp);
    f2_sub(&s, &t4, z3); f2_dbl(&s, &s); f2_add(&o.c0.c2, &s, &t4);
    f2_sub(&s, &t2, z4); f2_dbl(&s, &s); f2_add(&o.c0.c1, &s, &t2);
    f2_add(&s, &t3, z5); f2_dbl(&s, &s); f2_add(&o.c1.c2, &s, &t3);
    *r = o;
}
#define X_ABS 0xd201000000010000ULL
static void f12_exp_x(fp12* r, const fp12* a) { /* a^x, x < 0: conj(a^|x|), a in the cyclotomic subgroup */
    fp12 acc = *a; for (int i = 62; i >= 0; i--) { f12_cyc_sqr(&acc, &acc); if ((X_ABS >> i) & 1) f12_mul(&acc, &acc, a); } f12_conj(r, &acc);
}
/* ark-ec bls12 final_exponentiation: easy part then the hard part of eprint 2020/875 */
static void final_exp(fp12* r, const fp12* f) {
    fp12 a, b, c, t; f12_inv(&a, f); f12_conj(&b, f); f12_mul(&a, &a, &b); f12_frob(&b, &a, 2); f12_mul(&t, &a, &b);
    f12_exp_x(&a, &t); f12_conj(&b, &t); f12_mul(&a, &a, &b);
    f12_exp_x(&b, &a); f12_conj(&c, &a); f12_mul(&a, &b, &c);
    f12_exp_x(&b, &a); f12_frob(&c, &a, 1); f12_mul(&a, &b, &c);
    f12_exp_x(&b, &a); f12_exp_x(&b, &b); f12_frob(&c, &a, 2); f12_mul(&b, &b, &c); f12_conj(&c, &a); f12_mul(&a, &b, &c);
    f12_cyc_sqr(&b, &t); f12_mul(&b, &b, &t); f12_mul(r, &a, &b);
}

/* ---- G1 (Jacobian over Fp) and G2 (Jacobian over Fp2) ---- */
typedef struct { fp x, y, z; } g1;
typedef struct { fp2 x, y, z; } g2;
static fp B1; static fp2 B2; static g1 G1_P1; static g2 G2_GEN;
static int g1_inf(const g1* p) { return fp_is_zero(&p->z); }
static void g1_dbl(g1* r, const g1* p) {
    fp A, B, C, D, E, F, t, z3; fp_sqr(&A, &p->x); fp_sqr(&B, &p->y); fp_sqr(&C, &B);
    fp_add(&t, &p->x, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_dbl(&D, &t); fp_dbl(&E, &A); fp_add(&E, &E, &A); fp_sqr(&F, &E);
    fp_mul(&z3, &p->y, &p->z); fp_dbl(&z3, &z3); fp_dbl(&t, &D); fp_sub(&r->x, &F, &t); fp_sub(&t, &D, &r->x); fp_mul(&t, &E, &t);
    fp_dbl(&C, &C); fp_dbl(&C, &C); fp_dbl(&C, &C); fp_sub(&r->y, &t, &C); r->z = z3;
}
static void g1_add(g1* r, const g1* p, const g1* q) {
    if (g1_inf(p)) { *r = *q; return; } if (g1_inf(q)) { *r = *p; return; }
    fp z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    fp_sqr(&z1z1, &p->z); fp_sqr(&z2z2, &q->z); fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s1, &p->y, &q->z); fp_mul(&s1, &s1, &z2z2); fp_mul(&s2, &q->y, &p->z); fp_mul(&s2, &s2, &z1z1);
    fp_sub(&h, &u2, &u1); fp_sub(&rr, &s2, &s1);
    if (fp_is_zero(&h)) { if (fp_is_zero(&rr)) { g1_dbl(r, p); } else { r->x = FP_ONE; r->y = FP_ONE; r->z = FP_ZERO; } return; }
    fp_dbl(&rr, &rr); fp_dbl(&i, &h); fp_sqr(&i, &i); fp_mul(&j, &h, &i); fp_mul(&v, &u1, &i);
    fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &j); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp_sub(&t, &v, &x3); fp_mul(&y3, &rr, &t); fp_mul(&t, &s1, &j); fp_dbl(&t, &t); fp_sub(&y3, &y3, &t);
    fp_add(&z3, &p->z, &q->z); fp_sqr(&z3, &z3); fp_sub(&z3, &z3, &z1z1); fp_sub(&z3, &z3, &z2z2); fp_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
/* ark-ec `Projective * Fr`: MSB-first double-and-add over the scalar's bits */
static void g1_mul(g1* r, const g1* p, const fr* k) {
    g1 acc; acc.x = FP_ONE; acc.y = FP_ONE; acc.z = FP_ZERO; int started = 0;
    for (int i = 255; i >= 0; i--) { if (started) g1_dbl(&acc, &acc); if ((k->l[i >> 6] >> (i & 63)) & 1) { g1_add(&acc, &acc, p); started = 1; } }
    *r = acc;
}
static void g1_compress(uint8_t out[48], const g1* p) {
    if (g1_inf(p)) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fp zi, zi2, x, y; fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi); fp_mul(&x, &p->x, &zi2); fp_mul(&zi2, &zi2, &zi); fp_mul(&y, &p->y, &zi2);
    fp_to_be48(out, &x); out[0] |= 0x80; if (fp_is_high(&y)) out[0] |= 0x20;
}
static int g1_decompress(g1* r, const uint8_t in[48]) { /* 0 ok, 1 identity, -1 bad */
    if (!(in[0] & 0x80)) return -1;
    if (in[0] & 0x40) { r->x = FP_ONE; r->y = FP_ONE; r->z = FP_ZERO; return 1; }
    uint8_t t[48]; memcpy(t, in, 48); t[0] &= 0x1f; fp x, rhs, y; fp_from_be48(&x, t);
    fp_sqr(&rhs, &x); fp_mul(&rhs, &rhs, &x); fp_add(&rhs, &rhs, &B1);
    u64 e[6], c = 0; /* (p+1)/4 */ { u64 tmp[6]; memcpy(tmp, P, 48); tmp[0] += 1; for (int i = 5; i >= 0; i--) { u64 v = tmp[i]; e[i] = (v >> 2) | (c << 62); c = v & 3; } }
    fp_pow(&y, &rhs, e, 6); fp chk; fp_sqr(&chk, &y); if (!fp_eq(&chk, &rhs)) return -1;
    if (fp_is_high(&y) != ((in[0] & 0x20) != 0)) fp_neg(&y, &y);
    r->x = x; r->y = y; r->z = FP_ONE;
    /* Validate::Yes: ark-bls12-381 0.4.0 `is_in_correct_subgroup_assuming_on_curve` for G1 = the endomorphism test
       -[x^2]P == endomorphism(P) (two multiplications by the 64-bit |x|); here against beta^2: [x^2]P == (beta^2 x, -y) */
    {
        static const uint8_t BETA2_BE[48] = {0x00,0x00,0x00,0x00,0x00,0x00,0x00,0x00,0x5f,0x19,0x67,0x2f,0xdf,0x76,0xce,0x51,0xba,0x69,0xc6,0x07,
            0x6a,0x0f,0x77,0xea,0xdd,0xb3,0xa9,0x3b,0xe6,0xf8,0x96,0x88,0xde,0x17,0xd8,0x13,0x62,0x0a,0x00,0x02,0x2e,0x01,0xff,0xff,
            0xff,0xfe,0xff,0xfe};
        fr k; memset(&k, 0, sizeof k); k.l[0] = 0xd201000000010000ull;
        g1 q; g1_mul(&q, r, &k); g1_mul(&q, &q, &k);
        if (g1_inf(&q)) return -1;
        fp b2, zz, t; fp_from_be48(&b2, BETA2_BE); fp_sqr(&zz, &q.z); fp_mul(&t, &b2, &x); fp_mul(&t, &t, &zz);
        if (!fp_eq(&t, &q.x)) return -1;
        fp_mul(&zz, &zz, &q.z); fp_mul(&t, &y, &zz); fp_neg(&t, &t);
        if (!fp_eq(&t, &q.y)) return -1;
    }
    return 0;
}
static int g2_inf(const g2* p) { return f2_is_zero(&p->z); }
static void g2_dbl(g2* r, const g2* p) {
    fp2 A, B, C, D, E, F, t, z3; f2_sqr(&A, &p->x); f2_sqr(&B, &p->y); f2_sqr(&C, &B);
    f2_add(&t, &p->x, &B); f2_sqr(&t, &t); f2_sub(&t, &t, &A); f2_sub(&t, &t, &C); f2_dbl(&D, &t); f2_dbl(&E, &A); f2_add(&E, &E, &A); f2_sqr(&F, &E);
    f2_mul(&z3, &p->y, &p->z); f2_dbl(&z3, &z3); f2_dbl(&t, &D); f2_sub(&r->x, &F, &t); f2_sub(&t, &D, &r->x); f2_mul(&t, &E, &t);
    f2_dbl(&C, &C); f2_dbl(&C, &C); f2_dbl(&C, &C); f2_sub(&r->y, &t, &C); r->z = z3;
}
static void g2_add(g2* r, const g2* p, const g2* q) {
    if (g2_inf(p)) { *r = *q; return; } if (g2_inf(q)) { *r = *p; return; }
    fp2 z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    f2_sqr(&z1z1, &p->z); f2_sqr(&z2z2, &q->z); f2_mul(&u1, &p->x, &z2z2); f2_mul(&u2, &q->x, &z1z1);
    f2_mul(&s1, &p->y, &q->z); f2_mul(&s1, &s1, &z2z2); f2_mul(&s2, &q->y, &p->z); f2_mul(&s2, &s2, &z1z1);
    f2_sub(&h, &u2, &u1); f2_sub(&rr, &s2, &s1);
    if (f2_is_zero(&h)) { if (f2_is_zero(&rr)) { g2_dbl(r, p); } else { r->x = F2_ONE; r->y = F2_ONE; r->z = F2_ZERO; } return; }
    f2_dbl(&rr, &rr); f2_dbl(&i, &h); f2_sqr(&i, &i); f2_mul(&j, &h, &i); f2_mul(&v, &u1, &i);
    f2_sqr(&x3, &rr); f2_sub(&x3, &x3, &j); f2_sub(&x3, &x3, &v); f2_sub(&x3, &x3, &v);
    f2_sub(&t, &v, &x3); f2_mul(&y3, &rr, &t); f2_mul(&t, &s1, &j); f2_dbl(&t, &t); f2_sub(&y3, &y3, &t);
    f2_add(&z3, &p->z, &q->z); f2_sqr(&z3, &z3); f2_sub(&z3, &z3, &z1z1); f2_sub(&z3, &z3, &z2z2); f2_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g2_mul(g2* r, const g2* p, const fr* k) {
    g2 acc; acc.x = F2_ONE; acc.y = F2_ONE; acc.z = F2_ZERO; int started = 0;
    for (int i = 255; i >= 0; i--) { if (started) g2_dbl(&acc, &acc); if ((k->l[i >> 6] >> (i & 63)) & 1) { g2_add(&acc, &acc, p); started = 1; } }
    *r = acc;
}
static void g2_to_affine(fp2* x, fp2* y, const g2* p) { fp2 zi, zi2; f2_inv(&zi, &p->z); f2_sqr(&zi2, &zi); f2_mul(x, &p->x, &zi2); f2_mul(&zi2, &zi2, &zi); f2_mul(y, &p->y, &zi2); }
static int f2_is_high(const fp2* y) { return fp_is_zero(&y->c1) ? fp_is_high(&y->c0) : fp_is_high(&y->c1); }
static void g2_compress(uint8_t out[96], const g2* p) {
    if (g2_inf(p)) { memset(out, 0, 96); out[0] = 0xc0; return; }
    fp2 x, y; g2_to_affine(&x, &y, p); fp_to_be48(out, &x.c1); fp_to_be48(out + 48, &x.c0); out[0] |= 0x80; if (f2_is_high(&y)) out[0] |= 0x20;
}
static void f2_pow(fp2* r, const fp2* a, const u64* e, int nl) { fp2 acc = F2_ONE; for (int i = nl * 64 - 1; i >= 0; i--) { f2_sqr(&acc, &acc); if ((e[i >> 6] >> (i & 63)) & 1) f2_mul(&acc, &acc, a); } *r = acc; }
static int g2_decompress(g2* r, const uint8_t in[96]) {
    if (!(in[0] & 0x80)) return -1;
    if (in[0] & 0x40) { r->x = F2_ONE; r->y = F2_ONE; r->z = F2_ZERO; return 1; }
    uint8_t t[48]; memcpy(t, in, 48); t[0] &= 0x1f; fp2 x, rhs, y; fp_from_be48(&x.c1, t); fp_from_be48(&x.c0, in + 48);
    f2_sqr(&rhs, &x); f2_mul(&rhs, &rhs, &x); f2_add(&rhs, &rhs, &B2);
    /* sqrt in Fp2, p = 3 mod 4 */
    u64 e1[6], e2[6], c = 0; { u64 tmp[6]; memcpy(tmp, P, 48); tmp[0] -= 3; for (int i = 5; i >= 0; i--) { u64 v = tmp[i]; e1[i] = (v >> 2) | (c << 62); c = v & 3; } c = 0; for (int i = 5; i >= 0; i--) { u64 v = P[i]; e2[i] = (v >> 1) | (c << 63); c = v & 1; } }
    fp2 a1, alpha, x0, m1, cand; f2_pow(&a1, &rhs, e1, 6); f2_sqr(&alpha, &a1); f2_mul(&alpha, &alpha, &rhs); f2_mul(&x0, &a1, &rhs); f2_neg(&m1, &F2_ONE);
    if (memcmp(&alpha, &m1, sizeof(fp2)) == 0) { fp_neg(&cand.c0, &x0.c1); cand.c1 = x0.c0; } else { fp2 b; f2_add(&b, &F2_ONE, &alpha); f2_pow(&b, &b, e2, 6); f2_mul(&cand, &b, &x0); }
    fp2 chk; f2_sqr(&chk, &cand); if (memcmp(&chk, &rhs, sizeof(fp2)) != 0) return -1;
    y = cand; if (f2_is_high(&y) != ((in[0] & 0x20) != 0)) f2_neg(&y, &y);
    r->x = x; r->y = y; r->z = F2_ONE; return 0;
}

/* ---- Miller loop, ark-ec bls12 style: G2 prepared with homogeneous projective steps ---- */
typedef struct { fp2 c0, c1, c2; } ell_coeff;
static fp TWO_INV;
static void dbl_step(ell_coeff* o, g2* r) {
    fp2 a, b, c, e, f, g, h, i, j, e2, t;
    f2_mul(&a, &r->x, &r->y); f2_mul_fp(&a, &a, &TWO_INV); f2_sqr(&b, &r->y); f2_sqr(&c, &r->z);
    f2_dbl(&t, &c); f2_add(&t, &t, &c); f2_mul(&e, &B2, &t); f2_dbl(&f, &e); f2_add(&f, &f, &e);
    f2_add(&g, &b, &f); f2_mul_fp(&g, &g, &TWO_INV); f2_add(&h, &r->y, &r->z); f2_sqr(&h, &h); f2_add(&t, &b, &c); f2_sub(&h, &h, &t);
    f2_sub(&i, &e, &b); f2_sqr(&j, &r->x); f2_sqr(&e2, &e);
    f2_sub(&t, &b, &f); f2_mul(&r->x, &a, &t); f2_sqr(&g, &g); f2_dbl(&t, &e2); f2_add(&t, &t, &e2); f2_sub(&r->y, &g, &t); f2_mul(&r->z, &b, &h);
    o->c0 = i; f2_dbl(&t, &j); f2_add(&o->c1, &t, &j); f2_neg(&o->c2, &h);   /* M-type: (i, 3j, -h) */
}
static void add_step(ell_coeff* o, g2* r, const fp2* qx, const fp2* qy) {
    fp2 theta, lambda, c, d, e, f, g, h, t, j;
    f2_mul(&t, qy, &r->z); f2_sub(&theta, &r->y, &t); f2_mul(&t, qx, &r->z); f2_sub(&lambda, &r->x, &t);
    f2_sqr(&c, &theta); f2_sqr(&d, &lambda); f2_mul(&e, &lambda, &d); f2_mul(&f, &r->z, &c); f2_mul(&g, &r->x, &d);
    f2_add(&h, &e, &f); f2_dbl(&t, &g); f2_sub(&h, &h, &t);
    f2_mul(&r->x, &lambda, &h); f2_sub(&t, &g, &h); f2_mul(&t, &theta, &t); f2_mul(&c, &e, &r->y); f2_sub(&r->y, &t, &c); f2_mul(&r->z, &r->z, &e);
    f2_mul(&j, &theta, qx); f2_mul(&t, &lambda, qy); f2_sub(&j, &j, &t);
    o->c0 = j; f2_neg(&o->c1, &theta); o->c2 = lambda;                         /* M-type: (j, -theta, lambda) */
}
/* e(P, Q) as ark-ec computes it: multi_miller_loop over one pair + final_exponentiation; identity arguments give 1 */
static void pairing(fp12* out, const g1* p, const g2* q) {
    if (g1_inf(p) || g2_inf(q)) { *out = F12_ONE; return; }
    fp px, py; { fp zi, zi2; fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi); fp_mul(&px, &p->x, &zi2); fp_mul(&zi2, &zi2, &zi); fp_mul(&py, &p->y, &zi2); }
    fp2 qx, qy; g2_to_affine(&qx, &qy, q);
    ell_coeff co[68]; int n = 0; g2 r; r.x = qx; r.y = qy; r.z = F2_ONE;
    for (int i = 62; i >= 0; i--) { dbl_step(&co[n++], &r); if ((X_ABS >> i) & 1) add_step(&co[n++], &r, &qx, &qy); }
    fp12 f = F12_ONE; n = 0;
    for (int i = 62; i >= 0; i--) {
        if (i != 62) f12_sqr(&f, &f);
        fp2 c1, c2; f2_mul_fp(&c2, &co[n].c2, &py); f2_mul_fp(&c1, &co[n].c1, &px); f12_mul_by_014(&f, &co[n].c0, &c1, &c2); n++;
        if ((X_ABS >> i) & 1) { f2_mul_fp(&c2, &co[n].c2, &py); f2_mul_fp(&c1, &co[n].c1, &px); f12_mul_by_014(&f, &co[n].c0, &c1, &c2); n++; }
    }
    fp12 fc; f12_conj(&fc, &f);   /* x < 0 */
    final_exp(out, &fc);
}

/* ---- context: the `generators` argument and the issuer key, decoded once (they are typed values in the reference) ---- */
/* ---- create_generators (interface_utilities.rs:47-73) with HashToG1Bls12381 (:30-44) = zkcrypto bls12_381 hash_to_curve,
 * RFC 9380 BLS12381G1_XMD:SHA-256_SSWU_RO_: hash_to_field (two elements from 128 bytes), simplified SWU onto the 11-isogenous
 * curve E': y^2 = x^3 + A'x + B' (Z = 11), the isogeny E' -> E evaluated by Velu's formulas from the x-coordinates of its
 * rational order-11 kernel (the five denominators share one inversion), cofactor clearing by
 * h_eff = 0xd201000000010001.  The reference runs this for L + 1 generators inside EVERY verify (verify.rs:35). ---- */
static fp H2C_A, H2C_B, H2C_Z, H2C_NBA, H2C_BZA, H2C_I11_2, H2C_I11_3;
static struct { fp xq, yq, gxgy, vq, uq; } VELU[5];
static u64 EXP_SQRT[6];
static void fp_from_hex(fp* r, const char* hex) {       /* 96 hex digits, big-endian */
    uint8_t b[48]; for (int i = 0; i < 48; i++) { unsigned v; sscanf(hex + 2 * i, "%2x", &v); b[i] = (uint8_t)v; } fp_from_be48(r, b);
}
static int fp_sqrt(fp* y, const fp* a) { fp c; fp_pow(y, a, EXP_SQRT, 6); fp_sqr(&c, y); return fp_eq(&c, a); }
static int fp_parity(const fp* a) { u64 c[6]; fp_to_canon(c, a); return (int)(c[0] & 1); }
static void h2c_init(void) {
    fp_from_hex(&H2C_A, "00144698a3b8e9433d693a02c96d4982b0ea985383ee66a8d8e8981aefd881ac98936f8da0e0f97f5cf428082d584c1d");
    fp_from_hex(&H2C_B, "12e2908d11688030018b12e8753eee3b2016c1f0f24f4070a0b9c14fcef35ef55a23215a316ceaa5d1cc48e98e172be0");
    fp_from_u64(&H2C_Z, 11);
    fp ai, t; fp_inv(&ai, &H2C_A); fp_mul(&t, &H2C_B, &ai); fp_neg(&H2C_NBA, &t);                 /* -B/A */
    fp za; fp_mul(&za, &H2C_Z, &H2C_A); fp_inv(&za, &za); fp_mul(&H2C_BZA, &H2C_B, &za);       /* B/(Z A) */
    fp e11; fp_from_u64(&e11, 11); fp_inv(&e11, &e11); fp_sqr(&H2C_I11_2, &e11); fp_mul(&H2C_I11_3, &H2C_I11_2, &e11);
    { u64 tmp[6]; memcpy(tmp, P, 48); tmp[0] += 1; u64 c = 0; for (int i = 5; i >= 0; i--) { u64 v = tmp[i]; EXP_SQRT[i] = (v >> 2) | (c << 62); c = v & 3; } }
    static const char* KER[5] = {
        "010ef325dd1e98bdf0d97a4c6b7f968ed7f31f2fbff088acb39d5319cfc261ea18773405f325612742f0c5d90634bcf4",
        "0d7f2d0d03ae035321eed4c1479d13251abf0e9a96479623eb5380b575e319851fb5e5a8b43b9c1a46880f54bf2b2f7c",
        "105249b4cac630ce5aa18e6c1189a18c82019b4e12e491fbac012c259ca3a67f638560b8bb416af02a4724385ed0fc8e",
        "140d41735b10ce710727cd9356905701a2b866b803baa468948b7f423ddcc560c9a8f1cd5f8ed4297c37464fb8bfe4a7",
        "1665a9c648e78314490a94f654d9b1039ab85847223bfaed9aa54f0f07736d122d1ceca1ac0e9123e753fde16e97c3d7"};
    for (int k = 0; k < 5; k++) {
        fp xq, rhs, yq, gx, gy, t2; fp_from_hex(&xq, KER[k]);
        fp_sqr(&rhs, &xq); fp_mul(&rhs, &rhs, &xq); fp_mul(&t2, &H2C_A, &xq); fp_add(&rhs, &rhs, &t2); fp_add(&rhs, &rhs, &H2C_B);
        fp_sqrt(&yq, &rhs);
        fp_sqr(&gx, &xq); fp_dbl(&t2, &gx); fp_add(&gx, &gx, &t2); fp_add(&gx, &gx, &H2C_A);            /* 3 xq^2 + A */
        fp_dbl(&gy, &yq); fp_neg(&gy, &gy);                                                              /* -2 yq */
        VELU[k].xq = xq; VELU[k].yq = yq; fp_mul(&VELU[k].gxgy, &gx, &gy); fp_dbl(&VELU[k].vq, &gx); fp_sqr(&VELU[k].uq, &gy);
    }
}
static void fp_from_be64_mod(fp* r, const uint8_t b[64]) {      /* OS2IP(64 bytes) mod p, Montgomery form */
    uint8_t hi[48] = {0}; memcpy(hi + 32, b, 16); fp h, l; fp_from_be48(&h, hi); fp_from_be48(&l, b + 16); fp_mul(&h, &h, &FP_R2); fp_add(r, &h, &l);
}
static void sswu(fp* X, fp* Y, const fp* u) {
    fp u2, tv, tv1, x1, g, t, y;
    fp_sqr(&u2, u); fp_mul(&tv, &H2C_Z, &u2);                       /* Z u^2 */
    fp_sqr(&t, &tv); fp_add(&tv1, &t, &tv);                          /* Z^2 u^4 + Z u^2 */
    if (fp_is_zero(&tv1)) x1 = H2C_BZA;
    else { fp_inv(&tv1, &tv1); fp_add(&t, &FP_ONE, &tv1); fp_mul(&x1, &H2C_NBA, &t); }
    fp_sqr(&g, &x1); fp_mul(&g, &g, &x1); fp_mul(&t, &H2C_A, &x1); fp_add(&g, &g, &t); fp_add(&g, &g, &H2C_B);
    if (fp_sqrt(&y, &g)) { *X = x1; }
    else {
        fp_mul(X, &tv, &x1);                                         /* Z u^2 x1 */
        fp_sqr(&g, X); fp_mul(&g, &g, X); fp_mul(&t, &H2C_A, X); fp_add(&g, &g, &t); fp_add(&g, &g, &H2C_B);
        fp_sqrt(&y, &g);
    }
    if (fp_parity(u) != fp_parity(&y)) fp_neg(&y, &y);
    *Y = y;
}
static void iso11(g1* out, const fp* X, const fp* Y) {
    fp den[5], pre[5], inv, xo = *X, yo = *Y, t, d, d2, d3, a, b;
    for (int k = 0; k < 5; k++) { fp_sub(&den[k], X, &VELU[k].xq); if (k == 0) pre[0] = den[0]; else fp_mul(&pre[k], &pre[k - 1], &den[k]); }
    fp_inv(&inv, &pre[4]);
    for (int k = 4; k >= 0; k--) {
        if (k > 0) { fp_mul(&d, &inv, &pre[k - 1]); fp_mul(&inv, &inv, &den[k]); } else d = inv;
        fp_sqr(&d2, &d); fp_mul(&d3, &d2, &d);
        fp_mul(&a, &VELU[k].vq, &d); fp_mul(&b, &VELU[k].uq, &d2); fp_add(&xo, &xo, &a); fp_add(&xo, &xo, &b);
        fp_mul(&a, &VELU[k].uq, Y); fp_dbl(&a, &a); fp_mul(&a, &a, &d3);                       /* 2 uq Y d^3 */
        fp_sub(&t, Y, &VELU[k].yq); fp_mul(&b, &VELU[k].vq, &t); fp_mul(&b, &b, &d2);            /* vq (Y - yq) d^2 */
        fp_add(&a, &a, &b); fp_mul(&b, &VELU[k].gxgy, &d2); fp_sub(&a, &a, &b);                  /* - gx gy d^2 */
        fp_sub(&yo, &yo, &a);
    }
    fp_mul(&out->x, &xo, &H2C_I11_2); fp_mul(&out->y, &yo, &H2C_I11_3); out->z = FP_ONE;
}
static void hash_to_g1(g1* out, const uint8_t* msg, size_t n, const uint8_t* dst, size_t dlen) {
    uint8_t ub[128]; xmd(ub, 128, msg, n, NULL, 0, dst, dlen);
    fp u0, u1, X, Y; g1 q0, q1, r; fp_from_be64_mod(&u0, ub); fp_from_be64_mod(&u1, ub + 64);
    sswu(&X, &Y, &u0); iso11(&q0, &X, &Y); sswu(&X, &Y, &u1); iso11(&q1, &X, &Y);
    g1_add(&r, &q0, &q1);
    fr h; memset(&h, 0, sizeof h); h.l[0] = 0xd201000000010001ULL; g1_mul(out, &r, &h);
}
static void create_generators(g1* out, int count, const uint8_t* api_id, size_t api_len) {
    uint8_t seed_dst[300], gen_dst[300], gen_seed[300], v[56];
    memcpy(seed_dst, api_id, api_len); memcpy(seed_dst + api_len, "SIG_GENERATOR_SEED_", 19);
    memcpy(gen_dst, api_id, api_len); memcpy(gen_dst + api_len, "SIG_GENERATOR_DST_", 18);
    memcpy(gen_seed, api_id, api_len); memcpy(gen_seed + api_len, "MESSAGE_GENERATOR_SEED", 22);
    xmd(v, 48, gen_seed, api_len + 22, NULL, 0, seed_dst, api_len + 19);
    for (int i = 1; i <= count; i++) {
        for (int k = 0; k < 8; k++) v[48 + k] = (uint8_t)((u64)i >> (56 - 8 * k));
        uint8_t nv[48]; xmd(nv, 48, v, 56, NULL, 0, seed_dst, api_len + 19); memcpy(v, nv, 48);
        hash_to_g1(&out[i - 1], v, 48, gen_dst, api_len + 18);
    }
}

typedef struct { int L; g1* gens; g2 pk; uint8_t api_id[256]; size_t api_len; uint8_t dst_h2s[256], dst_map[256]; size_t dst_h2s_len, dst_map_len; } cref_ctx;

static int inited = 0;
static void init_consts(void) {
    if (inited) return;
    memset(&FP_ZERO, 0, sizeof FP_ZERO);
    /* R mod p and R^2 mod p by repeated doubling of 1 (Montgomery form of 1 is R mod p) */
    fp one; memset(&one, 0, sizeof one); one.l[0] = 1; fp t = one;
    for (int i = 0; i < 384; i++) fp_add(&t, &t, &t);
    FP_ONE = t; for (int i = 0; i < 384; i++) fp_add(&t, &t, &t); FP_R2 = t;
    F2_ZERO.c0 = FP_ZERO; F2_ZERO.c1 = FP_ZERO; F2_ONE.c0 = FP_ONE; F2_ONE.c1 = FP_ZERO;
    memset(&F12_ONE, 0, sizeof F12_ONE); F12_ONE.c0.c0.c0 = FP_ONE;
    fp_from_u64(&B1, 4); B2.c0 = B1; B2.c1 = B1;
    fp two; fp_from_u64(&two, 2); fp_inv(&TWO_INV, &two);
    /* Frobenius coefficients: xi^(k (p^j-1)/6) = (xi^((p^j-1)/6))^k; (p^j-1)/6 as a big exponent via repeated Fp2 pow by p */
    fp2 xi; xi.c0 = FP_ONE; xi.c1 = FP_ONE;
    /* g1 = xi^((p-1)/6): exponent (p-1)/6 fits 6 limbs */
    u64 e[6]; { u64 tmp[6]; memcpy(tmp, P, 48); tmp[0] -= 1; u128 rem = 0; for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | tmp[i]; e[i] = (u64)(cur / 6); rem = cur % 6; } }
    fp2 g[3]; f2_pow(&g[0], &xi, e, 6);
    /* xi^((p^2-1)/6) = g1^(p+1) = conj(g1) * g1 ; xi^((p^3-1)/6) = g1^(p^2+p+1) = g1 * conj(g1)^... use Frobenius: a^p = conj(a) in Fp2 */
    fp2 c; f2_conj(&c, &g[0]); f2_mul(&g[1], &c, &g[0]);             /* g1^p * g1 */
    fp2 c2; f2_conj(&c2, &g[1]); f2_mul(&g[2], &c2, &g[0]);           /* (g1^(p+1))^p * g1 = g1^(p^2+p+1) */
    for (int j = 0; j < 3; j++) { FROB[j][0] = F2_ONE; for (int k = 1; k < 6; k++) f2_mul(&FROB[j][k], &FROB[j][k - 1], &g[j]); }
    /* P1 (constants.rs:75-78) and BP2 from their compressed IRTF encodings (test_vector.rs:60-68) */
    static const uint8_t p1c[48] = {0xa8,0xce,0x25,0x61,0x02,0x84,0x08,0x21,0xa3,0xe9,0x4e,0xa9,0x02,0x5e,0x46,0x62,0xb2,0x05,0x76,0x2f,0x97,0x76,0xb3,0xa7,0x66,0xc8,0x72,0xb9,0x48,0xf1,0xfd,0x22,0x5e,0x7c,0x59,0x69,0x85,0x88,0xe7,0x0d,0x11,0x40,0x6d,0x16,0x1b,0x4e,0x28,0xc9};
    static const uint8_t bp2c[96] = {0x93,0xe0,0x2b,0x60,0x52,0x71,0x9f,0x60,0x7d,0xac,0xd3,0xa0,0x88,0x27,0x4f,0x65,0x59,0x6b,0xd0,0xd0,0x99,0x20,0xb6,0x1a,0xb5,0xda,0x61,0xbb,0xdc,0x7f,0x50,0x49,0x33,0x4c,0xf1,0x12,0x13,0x94,0x5d,0x57,0xe5,0xac,0x7d,0x05,0x5d,0x04,0x2b,0x7e,
                                     0x02,0x4a,0xa2,0xb2,0xf0,0x8f,0x0a,0x91,0x26,0x08,0x05,0x27,0x2d,0xc5,0x10,0x51,0xc6,0xe4,0x7a,0xd4,0xfa,0x40,0x3b,0x02,0xb4,0x51,0x0b,0x64,0x7a,0xe3,0xd1,0x77,0x0b,0xac,0x03,0x26,0xa8,0x05,0xbb,0xef,0xd4,0x80,0x56,0xc8,0xc1,0x21,0xbd,0xb8};
    g1_decompress(&G1_P1, p1c); g2_decompress(&G2_GEN, bp2c);
    h2c_init();
    inited = 1;
}

void* cref_ctx_create(const uint8_t* pk96, const uint8_t* gens48, int n_gens, const uint8_t* api_id, size_t api_len) {
    init_consts();
    cref_ctx* c = (cref_ctx*)calloc(1, sizeof(cref_ctx)); c->L = n_gens - 1; c->gens = (g1*)calloc(n_gens, sizeof(g1));
    if (g2_decompress(&c->pk, pk96) < 0) { free(c->gens); free(c); return NULL; }
    for (int i = 0; i < n_gens; i++) if (g1_decompress(&c->gens[i], gens48 + 48 * i) != 0) { free(c->gens); free(c); return NULL; }
    memcpy(c->api_id, api_id, api_len); c->api_len = api_len;
    memcpy(c->dst_h2s, api_id, api_len); memcpy(c->dst_h2s + api_len, "H2S_", 4); c->dst_h2s_len = api_len + 4;
    memcpy(c->dst_map, api_id, api_len); memcpy(c->dst_map + api_len, "MAP_MSG_TO_SCALAR_AS_HASH_", 26); c->dst_map_len = api_len + 26;
    return c;
}
void cref_ctx_destroy(void* p) { cref_ctx* c = (cref_ctx*)p; if (c) { free(c->gens); free(c); } }

/* calculate_domain (core_utilities.rs:24-63): compresses pk, Q1 and every H_i on each call, as the reference does */
static void calc_domain(fr* dom, const cref_ctx* c, const uint8_t* header, size_t hlen) {
    size_t L = (size_t)c->L, n = 96 + 8 + 48 * (L + 1) + c->api_len + 8 + hlen; uint8_t* buf = (uint8_t*)malloc(n), *q = buf;
    g2_compress(q, &c->pk); q += 96;
    for (int i = 0; i < 8; i++) *q++ = (uint8_t)((u64)L >> (56 - 8 * i));
    for (size_t i = 0; i <= L; i++) { g1_compress(q, &c->gens[i]); q += 48; }
    memcpy(q, c->api_id, c->api_len); q += c->api_len;
    for (int i = 0; i < 8; i++) *q++ = (uint8_t)((u64)hlen >> (56 - 8 * i));
    memcpy(q, header, hlen);
    hash_to_scalar(dom, buf, n, c->dst_h2s, c->dst_h2s_len); free(buf);
}
static void compute_B(g1* B, const cref_ctx* c, const fr* dom, const fr* m) {
    g1 t; g1_mul(&t, &c->gens[0], dom); g1_add(B, &G1_P1, &t);
    for (int i = 1; i <= c->L; i++) { g1_mul(&t, &c->gens[i], &m[i - 1]); g1_add(B, B, &t); }
}

/* PublicKey::verify for one item (verify.rs:18-93, generators given): returns 1 / 0, -1 for an undecodable signature */
static int verify_one(const cref_ctx* c, const uint8_t* sig80, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen);
int cref_verify_one(void* p, const uint8_t* sig80, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen) {
    return verify_one((const cref_ctx*)p, sig80, msgs, offs, header, hlen);
}
/* mode (A) of BASELINE.md: exactly PublicKey::verify (verify.rs:18-50) -- create_generators(L + 1) recomputed for the item */
int cref_verify_one_as_reference(void* p, const uint8_t* sig80, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen) {
    cref_ctx local = *(const cref_ctx*)p;
    g1* gens = (g1*)malloc((size_t)(local.L + 1) * sizeof(g1));
    create_generators(gens, local.L + 1, local.api_id, local.api_len);
    local.gens = gens;
    int r = verify_one(&local, sig80, msgs, offs, header, hlen);
    free(gens);
    return r;
}
void cref_create_generators(const uint8_t* api_id, size_t api_len, int count, uint8_t* out48) {
    init_consts();
    g1* g = (g1*)malloc((size_t)count * sizeof(g1));
    create_generators(g, count, api_id, api_len);
    for (int i = 0; i < count; i++) g1_compress(out48 + 48 * i, &g[i]);
    free(g);
}
static int verify_one(const cref_ctx* c, const uint8_t* sig80, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen) {
    g1 A; fr e; fr m[256];
    if (g1_decompress(&A, sig80) < 0) return -1;
    fr_from_le32(&e, sig80 + 48);
    for (int j = 0; j < c->L; j++) hash_to_scalar(&m[j], msgs + offs[j], (size_t)(offs[j + 1] - offs[j]), c->dst_map, c->dst_map_len);
    fr dom; calc_domain(&dom, c, header, hlen);
    g1 B; compute_B(&B, c, &dom, m);
    g2 w, t; g2_mul(&t, &G2_GEN, &e); g2_add(&w, &c->pk, &t);
    g2 nbp2 = G2_GEN; f2_neg(&nbp2.y, &nbp2.y);
    fp12 e1, e2, prod; pairing(&e1, &A, &w); pairing(&e2, &B, &nbp2); f12_mul(&prod, &e1, &e2);
    return memcmp(&prod, &F12_ONE, sizeof(fp12)) == 0;
}
/* batch over items with OpenMP: item i owns messages i*L .. i*L+L-1 of the flat message array */
void cref_verify_batch(void* p, size_t n, const uint8_t* sigs, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen, uint8_t* status, int threads) {
    cref_ctx* c = (cref_ctx*)p;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long i = 0; i < (long)n; i++) { int r = cref_verify_one(p, sigs + 80 * i, msgs, offs + (size_t)i * c->L, header, hlen); status[i] = r < 0 ? 5 : (uint8_t)r; }
}
void cref_verify_batch_as_reference(void* p, size_t n, const uint8_t* sigs, const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen, uint8_t* status, int threads) {
    cref_ctx* c = (cref_ctx*)p;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long i = 0; i < (long)n; i++) { int r = cref_verify_one_as_reference(p, sigs + 80 * i, msgs, offs + (size_t)i * c->L, header, hlen); status[i] = r < 0 ? 5 : (uint8_t)r; }
}
/* SecretKey::sign for one item (sign.rs:32-133, generators and pk given): B and A = B * (sk+e)^-1 */
int cref_sign_one(void* p, const uint8_t sk_le32[32], const uint8_t* msgs, const u64* offs, const uint8_t* header, size_t hlen, uint8_t sig80[80], uint8_t b48[48]) {
    cref_ctx* c = (cref_ctx*)p; fr sk, m[256], dom, e; fr_from_le32(&sk, sk_le32);
    for (int j = 0; j < c->L; j++) hash_to_scalar(&m[j], msgs + offs[j], (size_t)(offs[j + 1] - offs[j]), c->dst_map, c->dst_map_len);
    calc_domain(&dom, c, header, hlen);
    size_t n = 32 * ((size_t)c->L + 2); uint8_t* buf = (uint8_t*)malloc(n); fr_to_be32(buf, &sk);
    for (int j = 0; j < c->L; j++) fr_to_be32(buf + 32 * (j + 1), &m[j]); fr_to_be32(buf + 32 * (c->L + 1), &dom);
    hash_to_scalar(&e, buf, n, c->dst_h2s, c->dst_h2s_len); free(buf);
    g1 B; compute_B(&B, c, &dom, m); if (b48) g1_compress(b48, &B);
    /* (sk+e)^-1 mod r by Fermat with plain 256-bit shift-add arithmetic (slow but only used to make test data) */
    u64 s[5]; { u128 cc = 0; for (int i = 0; i < 4; i++) { cc += (u128)sk.l[i] + e.l[i]; s[i] = (u64)cc; cc >>= 64; } s[4] = (u64)cc; if (s[4] || ge4(s, RMOD)) sub4(s, s, RMOD); }
    /* modular multiplication by double-and-add */
    fr base, acc; memcpy(base.l, s, 32); memset(&acc, 0, sizeof acc); acc.l[0] = 1; u64 ex[4]; memcpy(ex, RMOD, 32); ex[0] -= 2;
    for (int i = 255; i >= 0; i--) {
        for (int rep = 0; rep < 2; rep++) {
            if (rep == 1 && !((ex[i >> 6] >> (i & 63)) & 1)) break;
            const fr* y = rep == 0 ? &acc : &base; fr x = acc, res; memset(&res, 0, sizeof res);
            for (int b = 255; b >= 0; b--) {
                u64 top = res.l[3] >> 63; for (int k = 3; k > 0; k--) res.l[k] = (res.l[k] << 1) | (res.l[k - 1] >> 63); res.l[0] <<= 1;
                if (top || ge4(res.l, RMOD)) sub4(res.l, res.l, RMOD);
                if ((y->l[b >> 6] >> (b & 63)) & 1) { u128 cc = 0; u64 tt[5]; for (int k = 0; k < 4; k++) { cc += (u128)res.l[k] + x.l[k]; tt[k] = (u64)cc; cc >>= 64; } tt[4] = (u64)cc; if (tt[4] || ge4(tt, RMOD)) sub4(tt, tt, RMOD); memcpy(res.l, tt, 32); }
            }
            acc = res;
        }
    }
    g1 A; g1_mul(&A, &B, &acc); g1_compress(sig80, &A);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) sig80[48 + 8 * i + k] = (uint8_t)(e.l[i] >> (8 * k));
    return 0;
}
