// SHA-256, RFC 9380 expand_message_xmd and hash_to_scalar on device.
// Replaces utilities_helper.rs:42-97 (expand_message), :15-40 (from_okm) and
// core_utilities.rs:11-21 (hash_to_scalar) -- rows a4/a5/a6 of SURVEY 8a.
#pragma once
#include "field.cuh"

namespace bbs {

BBS_CONST_ARRAY(SHA256_K, 64,
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u)

BBS_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

struct Sha256 {
    uint32_t h[8];
    uint32_t w[16];   // current block, big-endian words
    uint32_t fill;    // bytes in w
    uint32_t total;   // bytes absorbed (messages on this path are < 2^29 bytes)

    BBS_HD void init() {
        h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
        h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
        for (int i = 0; i < 16; i++) w[i] = 0;
        fill = 0; total = 0;
    }
    BBS_HDN void compress() {
        const uint32_t* K = SHA256_K();
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        uint32_t m[16];
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = w[i];
#pragma unroll
        for (int i = 0; i < 64; i++) {
            uint32_t wi;
            if (i < 16) {
                wi = m[i];
            } else {
                uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
                uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
                m[i & 15] = wi;
            }
            uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + K[i] + wi;
            uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
        fill = 0;
    }
    BBS_HD void put(uint8_t b) {
        w[fill >> 2] |= (uint32_t)b << (24 - 8 * (fill & 3));
        fill++; total++;
        if (fill == 64) compress();
    }
    // one big-endian word; the block position must be word-aligned
    BBS_HD void put_word(uint32_t x) {
        w[fill >> 2] = x;
        fill += 4; total += 4;
        if (fill == 64) compress();
    }
    BBS_HDN void update(const uint8_t* p, uint32_t n) {
        uint32_t i = 0;
        while (i < n && (fill & 3)) put(p[i++]);
        for (; i + 4 <= n; i += 4)
            put_word(((uint32_t)p[i] << 24) | ((uint32_t)p[i + 1] << 16) | ((uint32_t)p[i + 2] << 8) | p[i + 3]);
        for (; i < n; i++) put(p[i]);
    }
    // 32 bytes given as 8 big-endian words
    BBS_HD void update_words(const uint32_t* x, int nwords) {
        for (int i = 0; i < nwords; i++) {
            if ((fill & 3) == 0) put_word(x[i]);
            else { put((uint8_t)(x[i] >> 24)); put((uint8_t)(x[i] >> 16)); put((uint8_t)(x[i] >> 8)); put((uint8_t)x[i]); }
        }
    }
    // state after absorbing the 64 zero bytes Z_pad of expand_message_xmd (utilities_helper.rs:55): a constant
    BBS_HD void init_after_zpad() {
        h[0] = 0xda5698beu; h[1] = 0x17b9b469u; h[2] = 0x62335799u; h[3] = 0x779fbecau;
        h[4] = 0x8ce5d491u; h[5] = 0xc0d26243u; h[6] = 0xbafef9eau; h[7] = 0x1837a9d8u;
        for (int i = 0; i < 16; i++) w[i] = 0;
        fill = 0; total = 64;
    }
    BBS_HD void put_be64(uint64_t v) {
        for (int i = 7; i >= 0; i--) put((uint8_t)(v >> (8 * i)));
    }
    BBS_HDN void finish(uint32_t* out8) {
        uint32_t bits_lo = total << 3, bits_hi = total >> 29;
        uint32_t t = total;
        put(0x80);
        while (fill & 3) put(0);
        while (fill != 56) put_word(0);
        w[14] = bits_hi; w[15] = bits_lo;
        compress();
        total = t;
        for (int i = 0; i < 8; i++) out8[i] = h[i];
    }
};

// Streaming expand_message_xmd(msg, dst, 48): the caller absorbs the message between begin() and
// finish().  Output: 48 uniform bytes as 12 big-endian words.
struct Xmd48 {
    Sha256 s;
    BBS_HD void begin() { s.init_after_zpad(); }     // Z_pad (utilities_helper.rs:55) absorbed as a precomputed midstate
    BBS_HDN void finish(const uint8_t* dst, uint32_t dst_len, uint32_t* out12) {
        s.put(0); s.put(48); s.put(0);           // I2OSP(48, 2) || 0x00 (utilities_helper.rs:57)
        s.update(dst, dst_len); s.put((uint8_t)dst_len);
        uint32_t b0[8], b1[8], b2[8];
        s.finish(b0);
        s.init(); s.update_words(b0, 8); s.put(1); s.update(dst, dst_len); s.put((uint8_t)dst_len);
        s.finish(b1);
        uint32_t x[8];
        for (int i = 0; i < 8; i++) x[i] = b0[i] ^ b1[i];
        s.init(); s.update_words(x, 8); s.put(2); s.update(dst, dst_len); s.put((uint8_t)dst_len);
        s.finish(b2);
        for (int i = 0; i < 8; i++) out12[i] = b1[i];
        for (int i = 0; i < 4; i++) out12[8 + i] = b2[i];
    }
};

// from_okm: OS2IP(48 bytes) mod r -> canonical limbs (utilities_helper.rs:15-40).
// x = hi * 2^256 + lo with hi < 2^128: lo mod r by conditional subtractions (2^256 < 3r for BLS12-381,
// < 6r for BN254), hi * 2^256 mod r = mont_mul(hi, R^2) since R = 2^256.
template <class Fr>
BBS_HDN void okm48_to_scalar(uint32_t* r, const uint32_t* be12) {
    BBS_A16 uint32_t lo[8], hi[8], t[8];
    for (int i = 0; i < 8; i++) lo[i] = be12[11 - i];
    for (int i = 0; i < 4; i++) hi[i] = be12[3 - i];
    for (int i = 4; i < 8; i++) hi[i] = 0;
    for (int k = 0; k < 6; k++) {
        uint32_t borrow = bn_sub<8>(t, lo, Fr::P());
        if (!borrow) bn_copy<8>(lo, t);
    }
    fe_mul<Fr>(hi, hi, Fr::R2());   // hi * R mod r, canonical
    fe_add<Fr>(r, lo, hi);
}

// hash_to_scalar(msg, dst) for a byte-string message
template <class Fr>
BBS_HDN void hash_to_scalar(uint32_t* r, const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len) {
    Xmd48 x;
    x.begin();
    x.s.update(msg, len);
    uint32_t okm[12];
    x.finish(dst, dst_len, okm);
    okm48_to_scalar<Fr>(r, okm);
}

}  // namespace bbs
