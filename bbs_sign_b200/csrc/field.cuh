// Fixed-width Montgomery arithmetic on 32-bit limbs for the four prime fields of the path:
// BLS12-381 Fp (12 limbs) / Fr (8), BN254 Fp (8) / Fr (8).  Replaces ark-ff `Fp<MontBackend>` under
// every hot-path row of SURVEY 8a (a10).  R = 2^(32N), the same Montgomery radix arkworks uses for
// its 64-bit limbs, so values are interchangeable limb-for-limb.
//
// Device code: each row of the operand-scanning product is a pure mad.lo.cc / madc.hi.cc chain
// (gen_mont_asm.cuh), i.e. carries ride on the IMAD pipe; no separate carry instructions.
// Host code (`BBS_HOSTSIM`, tests only): portable 64-bit fallback of the same functions so the whole
// pipeline can be debugged in a container without a GPU.  The shipped library is CUDA-only.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define BBS_HD __host__ __device__ __forceinline__
#define BBS_HDN inline __host__ __device__ __noinline__
#else
#define BBS_HD inline
#define BBS_HDN inline __attribute__((noinline))
#ifndef __constant__
#define __constant__
#endif
#endif

// Field elements are moved with 128-bit loads / stores on the device (fe_ld / fe_st below), so every array that can hold
// one -- local arrays, shared-memory trees, constant tables, records in global memory -- is 16-byte aligned: local arrays
// are declared with BBS_A16, element sizes (32 / 48 bytes) keep sub-arrays aligned.
#if defined(__CUDACC__)
#define BBS_A16 __align__(16)
#else
#define BBS_A16 __attribute__((aligned(16)))
#endif

#if defined(__CUDACC__)
#define BBS_CONST_ARRAY(name, n, ...)                                   \
    static __constant__ BBS_A16 uint32_t name##_dev[n] = {__VA_ARGS__};  \
    static const BBS_A16 uint32_t name##_host[n] = {__VA_ARGS__};        \
    BBS_HD const uint32_t* name() {                                      \
        return BBS_SELECT_DEV(name##_dev, name##_host);                  \
    }
#ifdef __CUDA_ARCH__
#define BBS_SELECT_DEV(d, h) (d)
#else
#define BBS_SELECT_DEV(d, h) (h)
#endif
#else
#define BBS_CONST_ARRAY(name, n, ...)                                   \
    static const BBS_A16 uint32_t name##_host[n] = {__VA_ARGS__};        \
    inline const uint32_t* name() { return name##_host; }
#endif

// Every out-of-line device function starts and ends with a compiler-level memory barrier.  Without it
// cicc 12.9's inter-procedural attribute inference was observed to constant-fold a comparison of two
// buffers written by such calls (fe_sqrt's y^2 == x check became `false`); the barrier makes the callee
// opaque ("may read / write anything") at zero run-time cost.  Repro notes: tools/dbg/.
#ifdef __CUDA_ARCH__
#define BBS_OPAQUE_CALL_BARRIER() asm volatile("" ::: "memory")
#else
#define BBS_OPAQUE_CALL_BARRIER() ((void)0)
#endif

#include "gen_constants.cuh"
#include "gen_mont_asm.cuh"
#include "gen_mont_mul.cuh"
#include "gen_mont_sqr.cuh"

namespace bbs {

// ---- field parameter packs ---------------------------------------------------------------------
struct BlsFp {
    static constexpr int N = 12;
    static constexpr int BITS = BLS_FP_BITS;
    static constexpr uint32_t INV = BLS_FP_INV;
    static BBS_HD const uint32_t* P() { return BLS_FP_P(); }
    static BBS_HD const uint32_t* ONE() { return BLS_FP_ONE(); }
    static BBS_HD const uint32_t* R2() { return BLS_FP_R2(); }
    static BBS_HD const uint32_t* EXP_INV() { return BLS_FP_EXP_INV(); }
    static BBS_HD const uint32_t* EXP_SQRT() { return BLS_FP_EXP_SQRT(); }
    static BBS_HD const uint32_t* HALF() { return BLS_FP_HALF(); }
    static BBS_HD const uint32_t* EXP_PM3D4() { return BLS_FP_EXP_PM3D4(); }
};
struct BlsFr {
    static constexpr int N = 8;
    static constexpr int BITS = BLS_FR_BITS;
    static constexpr uint32_t INV = BLS_FR_INV;
    static BBS_HD const uint32_t* P() { return BLS_FR_P(); }
    static BBS_HD const uint32_t* ONE() { return BLS_FR_ONE(); }
    static BBS_HD const uint32_t* R2() { return BLS_FR_R2(); }
    static BBS_HD const uint32_t* EXP_INV() { return BLS_FR_EXP_INV(); }
};
struct BnFp {
    static constexpr int N = 8;
    static constexpr int BITS = BN_FP_BITS;
    static constexpr uint32_t INV = BN_FP_INV;
    static BBS_HD const uint32_t* P() { return BN_FP_P(); }
    static BBS_HD const uint32_t* ONE() { return BN_FP_ONE(); }
    static BBS_HD const uint32_t* R2() { return BN_FP_R2(); }
    static BBS_HD const uint32_t* EXP_INV() { return BN_FP_EXP_INV(); }
    static BBS_HD const uint32_t* EXP_SQRT() { return BN_FP_EXP_SQRT(); }
    static BBS_HD const uint32_t* HALF() { return BN_FP_HALF(); }
    static BBS_HD const uint32_t* EXP_PM3D4() { return BN_FP_EXP_PM3D4(); }
};
struct BnFr {
    static constexpr int N = 8;
    static constexpr int BITS = BN_FR_BITS;
    static constexpr uint32_t INV = BN_FR_INV;
    static BBS_HD const uint32_t* P() { return BN_FR_P(); }
    static BBS_HD const uint32_t* ONE() { return BN_FR_ONE(); }
    static BBS_HD const uint32_t* R2() { return BN_FR_R2(); }
    static BBS_HD const uint32_t* EXP_INV() { return BN_FR_EXP_INV(); }
};

// ---- 128-bit moves between memory and register arrays (memory operands of the out-of-line functions) ----------
#ifdef __CUDA_ARCH__
template <int N> __device__ __forceinline__ void fe_ld(uint32_t* x, const uint32_t* p) {
    static_assert(N % 4 == 0, "limb count");
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < N / 4; i++) { uint4 v = q[i]; x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w; }
}
template <int N> __device__ __forceinline__ void fe_st(uint32_t* p, const uint32_t* x) {
    static_assert(N % 4 == 0, "limb count");
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < N / 4; i++) q[i] = make_uint4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}
#endif

#ifdef __CUDA_ARCH__
template <int N> __device__ __forceinline__ void bn_ld(uint32_t* x, const uint32_t* p) {
    if constexpr (N % 4 == 0) fe_ld<N>(x, p);
    else {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = p[i];
    }
}
template <int N> __device__ __forceinline__ void bn_st(uint32_t* p, const uint32_t* x) {
    if constexpr (N % 4 == 0) fe_st<N>(p, x);
    else {
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = x[i];
    }
}
#endif

// ---- raw multi-limb helpers ----------------------------------------------------------------------
template <int N>
BBS_HD uint32_t bn_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // returns carry
#ifdef __CUDA_ARCH__
    BBS_A16 uint32_t x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = a[i]; y[i] = b[i]; }
    uint32_t c = bbs_addn<N>(z, x, y);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = z[i];
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < N; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}

template <int N>
BBS_HD uint32_t bn_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // returns borrow mask
#ifdef __CUDA_ARCH__
    BBS_A16 uint32_t x[N], y[N], z[N];
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = a[i]; y[i] = b[i]; }
    uint32_t c = bbs_subn<N>(z, x, y);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = z[i];
    return c;
#else
    int64_t c = 0;
    for (int i = 0; i < N; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return c ? 0xffffffffu : 0u;
#endif
}

template <int N>
BBS_HD bool bn_is_zero(const uint32_t* a) {
    uint32_t o = 0;
#ifdef __CUDA_ARCH__
    uint32_t x[N];
    bn_ld<N>(x, a);
#pragma unroll
    for (int i = 0; i < N; i++) o |= x[i];
#else
    for (int i = 0; i < N; i++) o |= a[i];
#endif
    return o == 0;
}

template <int N>
BBS_HD bool bn_eq(const uint32_t* a, const uint32_t* b) {
    uint32_t o = 0;
#ifdef __CUDA_ARCH__
    uint32_t x[N], y[N];
    bn_ld<N>(x, a);
    bn_ld<N>(y, b);
#pragma unroll
    for (int i = 0; i < N; i++) o |= x[i] ^ y[i];
#else
    for (int i = 0; i < N; i++) o |= a[i] ^ b[i];
#endif
    return o == 0;
}

// Element copies go through an opaque register move on the device so NVVM cannot turn them into a
// memcpy: cicc 12.9 mis-optimises memcpy chains between the stack arrays of inlined callees (observed:
// fe_sqrt's final comparison constant-folded to false; tools/dbg/).
template <int N>
BBS_HD void bn_copy(uint32_t* r, const uint32_t* a) {
#ifdef __CUDA_ARCH__
    uint32_t x[N];
    bn_ld<N>(x, a);
#pragma unroll
    for (int i = 0; i < N; i++) asm volatile("" : "+r"(x[i]));
    bn_st<N>(r, x);
#else
    for (int i = 0; i < N; i++) r[i] = a[i];
#endif
}

template <int N>
BBS_HD void bn_zero(uint32_t* r) {
#ifdef __CUDA_ARCH__
    uint32_t x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 0;
    bn_st<N>(r, x);
#else
    for (int i = 0; i < N; i++) r[i] = 0;
#endif
}

// a > b as little-endian integers
template <int N>
BBS_HD bool bn_gt(const uint32_t* a, const uint32_t* b) {
    BBS_A16 uint32_t t[N];
    return bn_sub<N>(t, b, a) != 0;  // b - a borrows  <=>  a > b
}

// ---- modular add / sub / neg (canonical representatives in [0, p)) ------------------------------
template <class F>
BBS_HD void fe_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    constexpr int N = F::N;
    BBS_A16 uint32_t s[N], d[N];
    bn_add<N>(s, a, b);  // 2p < 2^(32N) for every modulus here: no carry out
    uint32_t borrow = bn_sub<N>(d, s, F::P());
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = borrow ? s[i] : d[i];
    bn_copy<N>(r, d);
}

template <class F>
BBS_HD void fe_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    constexpr int N = F::N;
    BBS_A16 uint32_t d[N], s[N];
    uint32_t borrow = bn_sub<N>(d, a, b);
    bn_add<N>(s, d, F::P());
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = borrow ? s[i] : d[i];
    bn_copy<N>(r, d);
}

template <class F>
BBS_HD void fe_neg(uint32_t* r, const uint32_t* a) {
    constexpr int N = F::N;
    BBS_A16 uint32_t d[N];
    bool z = bn_is_zero<N>(a);
    bn_sub<N>(d, F::P(), a);
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = z ? 0u : d[i];
    bn_copy<N>(r, d);
}

template <class F>
BBS_HD void fe_dbl(uint32_t* r, const uint32_t* a) { fe_add<F>(r, a, a); }

// ---- Montgomery product ---------------------------------------------------------------------------
// CIOS, one limb of b per row.  Bound: with a, b < p and 2p < R the running value stays < 2p, so the
// accumulator needs N+1 words and the result needs one conditional subtraction.
template <class F>
BBS_HD void fe_mul_inl(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    constexpr int N = F::N;
#ifdef __CUDA_ARCH__
    BBS_A16 uint32_t A[N], B[N], M[N], t[N];
    const uint32_t* pm = F::P();
    fe_ld<N>(A, a);                           // memory operands: 128-bit loads (16-byte aligned arrays, see BBS_A16)
    fe_ld<N>(B, b);
    fe_ld<N>(M, pm);
    bbs_mont_mul<N>(t, A, B, M, F::INV);      // two-accumulator CIOS, result < 2p (gen_mont_mul.cuh)
    BBS_A16 uint32_t d[N];
    uint32_t borrow = bbs_subn<N>(d, t, M);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : d[i];
    fe_st<N>(r, t);
#else
    BBS_A16 uint32_t t[N + 2];
    const uint32_t* pm = F::P();
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
        uint64_t c = 0;
        for (int j = 0; j < N; j++) {
            c += (uint64_t)a[j] * b[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[N];
        t[N] = (uint32_t)c;
        t[N + 1] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * F::INV;
        c = ((uint64_t)m * pm[0] + t[0]) >> 32;
        for (int j = 1; j < N; j++) {
            c += (uint64_t)m * pm[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[N];
        t[N - 1] = (uint32_t)c;
        t[N] = t[N + 1] + (uint32_t)(c >> 32);
    }
    BBS_A16 uint32_t d[N];
    uint32_t borrow = bn_sub<N>(d, t, pm);
    bool ge = (t[N] != 0) || !borrow;
    for (int i = 0; i < N; i++) r[i] = ge ? d[i] : t[i];
#endif
}

// Out-of-line instance: the 1-thread-per-item kernels call this so code size stays bounded
// (one ~0.7k-instruction body per field instead of one per call site).
template <class F>
BBS_HDN void fe_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    BBS_OPAQUE_CALL_BARRIER();
    fe_mul_inl<F>(r, a, b);
    BBS_OPAQUE_CALL_BARRIER();
}

// Dedicated squaring for the two base fields (gen_mont_sqr.cuh: n(n+1)/2 + n^2 + n products instead of 2n^2 + n); the
// scalar fields and the host build square by multiplying.  A translation unit may opt out (BBS_NO_DEDICATED_SQR): a second
// ~7 KB body competes with the multiplication body for the instruction cache, which costs sign_kernel (squarings and
// multiplications finely interleaved in its 160 mixed additions per item; ncu: 2.3 no-instruction stalls per issue) more
// than the 22 % fewer products save -- measured at fixed clocks: sign +4..9 % slower, verify_g1 -5 %, rlc_prep -9 %, the
// Fermat inversion inside pairing_coop_kernel -3.7 % of that kernel.
template <class F> struct FeSqr {
    static BBS_HD void run(uint32_t* r, const uint32_t* a) { fe_mul<F>(r, a, a); }
};
#if defined(__CUDA_ARCH__) && !defined(BBS_NO_DEDICATED_SQR)
template <class F, void (*SQR)(uint32_t*, const uint32_t*)>
__device__ __noinline__ void fe_sqr_ool(uint32_t* r, const uint32_t* a) {
    constexpr int N = F::N;
    BBS_OPAQUE_CALL_BARRIER();
    uint32_t A[N], M[N], t[N], d[N];
    fe_ld<N>(A, a);
    fe_ld<N>(M, F::P());
    SQR(t, A);                                // < 2p
    uint32_t borrow = bbs_subn<N>(d, t, M);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = borrow ? t[i] : d[i];
    fe_st<N>(r, t);
    BBS_OPAQUE_CALL_BARRIER();
}
template <> struct FeSqr<BlsFp> {
    static __device__ __forceinline__ void run(uint32_t* r, const uint32_t* a) { fe_sqr_ool<BlsFp, bbs_mont_sqr_bls_fp>(r, a); }
};
template <> struct FeSqr<BnFp> {
    static __device__ __forceinline__ void run(uint32_t* r, const uint32_t* a) { fe_sqr_ool<BnFp, bbs_mont_sqr_bn_fp>(r, a); }
};
#endif
template <class F>
BBS_HD void fe_sqr(uint32_t* r, const uint32_t* a) { FeSqr<F>::run(r, a); }
template <class F>
BBS_HD void fe_to_mont(uint32_t* r, const uint32_t* a) { fe_mul<F>(r, a, F::R2()); }

template <class F>
BBS_HD void fe_from_mont(uint32_t* r, const uint32_t* a) {
    BBS_A16 uint32_t one[F::N];
    bn_zero<F::N>(one);
    one[0] = 1;
    fe_mul<F>(r, a, one);
}

template <class F>
BBS_HD void fe_set_one(uint32_t* r) { bn_copy<F::N>(r, F::ONE()); }

// r = a^e, e given as little-endian 32-bit limbs (a public constant: uniform control flow across the
// warp); fixed 4-bit windows, MSB first.
template <class F>
BBS_HDN void fe_pow(uint32_t* r, const uint32_t* a, const uint32_t* e, int ebits) {
    constexpr int N = F::N;
    BBS_A16 uint32_t tab[16][N];
    fe_set_one<F>(tab[0]);
    bn_copy<N>(tab[1], a);
    for (int i = 2; i < 16; i++) fe_mul<F>(tab[i], tab[i - 1], a);
    BBS_A16 uint32_t acc[N];
    int top = (ebits + 3) / 4 - 1;
    {
        uint32_t d = (e[(top * 4) >> 5] >> ((top * 4) & 31)) & 15;
        bn_copy<N>(acc, tab[d]);
    }
    for (int w = top - 1; w >= 0; w--) {
        fe_sqr<F>(acc, acc);
        fe_sqr<F>(acc, acc);
        fe_sqr<F>(acc, acc);
        fe_sqr<F>(acc, acc);
        uint32_t d = (e[(w * 4) >> 5] >> ((w * 4) & 31)) & 15;
        if (d) fe_mul<F>(acc, acc, tab[d]);
    }
    bn_copy<N>(r, acc);
}

// Fermat inversion; inv(0) = 0 (matches the `inv0` convention; callers test for zero where it matters)
template <class F>
BBS_HDN void fe_inv(uint32_t* r, const uint32_t* a) { fe_pow<F>(r, a, F::EXP_INV(), F::BITS); }

// Variable-time inversion (binary extended Euclid), for the spots where ONE thread inverts while its block or the whole
// GPU waits (root of the block-wide inversion tree, the one-block tail kernels of the batch mode, context creation) and the
// input is public: ~4x fewer instructions than the Fermat ladder.  Never used on secret data (core_sign's 1 / (sk + e)
// keeps fe_inv).  Montgomery in, Montgomery out; inv(0) = 0.
template <int N> BBS_HD void bn_shr1(uint32_t* a, uint32_t top) {
    for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[N - 1] = (a[N - 1] >> 1) | (top << 31);
}
template <int N> BBS_HD bool bn_is_one(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
    for (int i = 1; i < N; i++) o |= a[i];
    return o == 0;
}
template <class F>
BBS_HDN void fe_inv_vt(uint32_t* r, const uint32_t* a) {
    constexpr int N = F::N;
    if (bn_is_zero<N>(a)) { bn_zero<N>(r); return; }
    BBS_A16 uint32_t u[N], v[N], x1[N], x2[N], t[N];
    bn_copy<N>(u, a); bn_copy<N>(v, F::P());
    bn_zero<N>(x1); x1[0] = 1; bn_zero<N>(x2);
    // invariants: x1 * a = u, x2 * a = v (mod p); u, v odd-or-being-halved, gcd(u, v) = 1
    while (!bn_is_one<N>(u) && !bn_is_one<N>(v)) {
        while (!(u[0] & 1u)) {
            bn_shr1<N>(u, 0);
            uint32_t c = 0;
            if (x1[0] & 1u) c = bn_add<N>(x1, x1, F::P());
            bn_shr1<N>(x1, c);
        }
        while (!(v[0] & 1u)) {
            bn_shr1<N>(v, 0);
            uint32_t c = 0;
            if (x2[0] & 1u) c = bn_add<N>(x2, x2, F::P());
            bn_shr1<N>(x2, c);
        }
        if (bn_sub<N>(t, u, v) == 0) {          // u >= v
            bn_copy<N>(u, t);
            if (bn_sub<N>(x1, x1, x2)) bn_add<N>(x1, x1, F::P());
        } else {
            bn_sub<N>(v, v, u);
            if (bn_sub<N>(x2, x2, x1)) bn_add<N>(x2, x2, F::P());
        }
    }
    // x = (a R)^-1 = a^-1 R^-1 for the Montgomery input a R; two products by R^2 give a^-1 R
    fe_mul<F>(t, bn_is_one<N>(u) ? x1 : x2, F::R2());
    fe_mul<F>(r, t, F::R2());
}

// square root for p = 3 mod 4; returns false when a is a non-residue
template <class F>
BBS_HDN bool fe_sqrt(uint32_t* r, const uint32_t* a) {
    BBS_A16 uint32_t s[F::N], q[F::N];
    fe_pow<F>(s, a, F::EXP_SQRT(), F::BITS - 1);
    fe_sqr<F>(q, s);
    bool ok = bn_eq<F::N>(q, a);
    bn_copy<F::N>(r, s);
    return ok;
}

// canonical(a) > (p-1)/2   (the "y is lexicographically largest / negative" flag of both encodings)
template <class F>
BBS_HDN bool fe_is_high(const uint32_t* a_mont) {
    BBS_A16 uint32_t c[F::N];
    fe_from_mont<F>(c, a_mont);
    return bn_gt<F::N>(c, F::HALF());
}

// ---- byte <-> limb conversion (canonical, non-Montgomery) -----------------------------------------
template <int N>
BBS_HD void limbs_from_be(uint32_t* r, const uint8_t* b) {  // 4N big-endian bytes
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = b + 4 * (N - 1 - i);
        r[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
}
template <int N>
BBS_HD void limbs_to_be(uint8_t* b, const uint32_t* a) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint8_t* q = b + 4 * (N - 1 - i);
        q[0] = (uint8_t)(a[i] >> 24); q[1] = (uint8_t)(a[i] >> 16); q[2] = (uint8_t)(a[i] >> 8); q[3] = (uint8_t)a[i];
    }
}
template <int N>
BBS_HD void limbs_from_le(uint32_t* r, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = b + 4 * i;
        r[i] = ((uint32_t)q[3] << 24) | ((uint32_t)q[2] << 16) | ((uint32_t)q[1] << 8) | q[0];
    }
}
template <int N>
BBS_HD void limbs_to_le(uint8_t* b, const uint32_t* a) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint8_t* q = b + 4 * i;
        q[0] = (uint8_t)a[i]; q[1] = (uint8_t)(a[i] >> 8); q[2] = (uint8_t)(a[i] >> 16); q[3] = (uint8_t)(a[i] >> 24);
    }
}

// canonical limbs < p ?
template <class F>
BBS_HD bool fe_is_canonical(const uint32_t* a) {
    return bn_gt<F::N>(F::P(), a);
}

}  // namespace bbs
