// Cooperative pairing kernel: "e(P0, Q0) e(P1, Q1) == 1" for 32 items per thread block, SIX warps per block.
// Warp k (a ROLE) owns the Fp2 coefficient c_k of every Fp12 value f = sum c_k w^k in Fp2[w]/(w^6 - xi); lane =
// item.  All values live in shared-memory cells; each role interprets its own instruction stream
// (gen_pairing_prog.cuh, built and verified against the oracle by tools/coop_prog.py):
//     EP  : acc_R +-= X*Y, acc_I +-= X*Y     one 12x12-limb product (two fresh half products on IMAD.WIDE chains)
//                                            fed to the double-width real / imaginary accumulators
//     FIN : cell = canon(REDC(3^t acc +- 2^d Z R + KP))   ONE Montgomery reduction per output coefficient
// so an Fp12 product costs 18 EPs + 2 reductions per role (Karatsuba at the Fp2 level, schoolbook above, lazy
// reduction) with no operand pre-additions beyond c0 +- c1, and the Fp12 state never touches local memory.
// Replaces the per-thread pairing_item for BLS12-381 (verify.rs:88-92, proof_verify.rs:112-115).
#pragma once
#include "kernels.cuh"

#ifdef __CUDACC__
#include "gen_coop.cuh"
#define COOP_BN_QCANON_M 1354u                 // floor(2^32 / ((p_bn254 >> 232) + 1)), checked by tools/gen_coop.py
#define COOP_BLS_QCANON_M 165164498u           // floor(2^56 / ((p_bls12_381 >> 352) + 1)), checked by tools/gen_coop.py
#ifdef BBS_COOP_PROG_HEADER
#include BBS_COOP_PROG_HEADER              // timing experiments: alternative (not result-correct) programs
#else
#include "gen_pairing_prog.cuh"
#endif

namespace bbs {

constexpr int COOP_ROLES = 6;
constexpr int COOP_ITEMS = 32;                 // items per block = lanes
constexpr int COOP_CELLS = 24;
// occupancy knobs per curve (-D overridable): independent 32-item groups per block (each with its own named barrier) and
// the register cap.  BLS12-381: 2 groups (73 KB of cells each), 128 registers.  BN254: 4 groups (49 KB each), 80 registers:
// its EPs are short (64 products), so twice the warps hide more of the shared-memory and barrier latency (measured).
#ifndef BBS_COOP_GROUPS_BLS
#define BBS_COOP_GROUPS_BLS 2
#endif
#ifndef BBS_COOP_GROUPS_BN
#define BBS_COOP_GROUPS_BN 4
#endif
#ifndef BBS_COOP_MAXREG_BLS
#define BBS_COOP_MAXREG_BLS 128
#endif
#ifndef BBS_COOP_MAXREG_BN
#define BBS_COOP_MAXREG_BN 80
#endif

struct CoopArgs {
    const uint32_t* lines;     // per ate line: B'0, A'0, B'1, A'1 (Fp2 each, Montgomery): 8N words
    const uint32_t* pair;      // per item: P0 = (x, y, 1), P1 = (x, y, 1) affine Montgomery (6N words)
    const uint32_t* flags;
    uint8_t* status;
    uint32_t* gscratch;        // per block: COOP_ROLES cells
    uint32_t n;
    const uint32_t* item_issuer;   // MULTI kernel: item i reads the line table at lines + item_issuer[i] * line_stride
    uint32_t line_stride;
    uint32_t split2_max;           // batches of up to this many items (<= 32) run the SPLIT = 2 kernel (two warps per role)
};

// ---- per-curve bindings of the generated primitives ---------------------------------------------------------
template <class C> struct Coop;
template <> struct Coop<Bls> {
    static constexpr int N = 12;
    static constexpr int GROUPS = BBS_COOP_GROUPS_BLS;
    static constexpr int MAXREG = BBS_COOP_MAXREG_BLS;
    static constexpr int RW = 12;              // words of a reduction result
    static constexpr int MAXK = 8;             // largest canonicalisation step (multiples of p)
    static constexpr int PARK = 1;             // Fp12 values a group parks in HBM (GSAVE / GLOAD slots)
    static constexpr bool ACC_XI = false;      // xi = 1 + u: multiplication by xi is routed through the EP signs
    static constexpr bool POINT_RATIO = false; // the item's points enter as (x, y)
    // w (12 words, < R < 10 p) -> w mod p: q^ = (w[11] * floor(2^56 / (top 29 bits of p + 1))) >> 56 is q or q - 1
    // (tools/gen_coop.py check_canon_q_bls), q^ p comes from a 10-row table in shared memory (the row index differs per lane),
    // then ONE conditional subtraction: 42 instructions per component where the step ladder (8p, 4p, 2p, p) took 50-100
    static constexpr bool QCANON = true;
    static constexpr int QCANON_MIN = 1;       // canon levels below this keep the ladder (level 0: one step)
    static constexpr int QTAB_UINT4 = 10 * 3;  // shared-memory table: q p for q < 10
    static __device__ __forceinline__ const uint32_t* qtab() { return COOP_QP_BLS; }
    static __device__ __forceinline__ void canon_q(uint32_t* w, const uint4* tab) {
        const uint32_t q = __umulhi(w[11], COOP_BLS_QCANON_M) >> 24;
        const uint4 a = tab[q * 3], b = tab[q * 3 + 1], c = tab[q * 3 + 2];
        asm("sub.cc.u32 %0, %0, %12; subc.cc.u32 %1, %1, %13; subc.cc.u32 %2, %2, %14; subc.cc.u32 %3, %3, %15; subc.cc.u32 %4, %4, %16; "
            "subc.cc.u32 %5, %5, %17; subc.cc.u32 %6, %6, %18; subc.cc.u32 %7, %7, %19; subc.cc.u32 %8, %8, %20; subc.cc.u32 %9, %9, %21; "
            "subc.cc.u32 %10, %10, %22; subc.u32 %11, %11, %23;"
            : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]),
              "+r"(w[10]), "+r"(w[11])
            : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w));
        uint32_t d[12];
        const uint32_t bo = coop_sub_1p_bls(d, w);
#pragma unroll
        for (int i = 0; i < 12; i++) w[i] = bo ? w[i] : d[i];
    }
    static constexpr bool WARP_INV = true;     // INV shares one inversion among the 32 lanes (coop_warp_inverse): -4.5 %
    static constexpr bool WARP_INV_PRO = false;
    static constexpr bool REP_SMEM = false;
    static constexpr int KPW = 12;             // KP[k] = k p R: the table holds k p, added to the high half of the accumulator
    static __device__ __forceinline__ void add_kp(uint32_t* acc, const uint32_t* k) { coop_acc_add_hi12(acc, k); }
    static __device__ __forceinline__ void xi(uint32_t*, uint32_t*) {}
    static __device__ __forceinline__ const uint32_t* prog() { return COOP_PROG_BLS; }
    static __device__ __forceinline__ const uint32_t* prog_off() { return COOP_PROG_OFF_BLS; }
    static __device__ __forceinline__ const uint32_t* consts() { return COOP_CONSTS_BLS; }
    static __device__ __forceinline__ const uint32_t* kp() { return COOP_KP_BLS; }
    static __device__ __forceinline__ void wmul_e(uint32_t* w, const uint32_t* a, const uint32_t* b) { coop_wmul_e12(w, a, b); }
    static __device__ __forceinline__ void wmul_o(uint32_t* w, const uint32_t* a, const uint32_t* b) { coop_wmul_o12(w, a, b); }
    static __device__ __forceinline__ void merge(uint32_t* w, const uint32_t* v) { coop_merge12(w, v); }
    static __device__ __forceinline__ void add_e(uint32_t* acc, const uint32_t* w) { coop_acc_add_e12(acc, w); }
    static __device__ __forceinline__ void sub_e(uint32_t* acc, const uint32_t* w) { coop_acc_sub_e12(acc, w); }
    static __device__ __forceinline__ void add_hi(uint32_t* acc, const uint32_t* z) { coop_acc_add_hi12(acc, z); }
    static __device__ __forceinline__ void sub_hi(uint32_t* acc, const uint32_t* z) { coop_acc_sub_hi12(acc, z); }
    static __device__ __forceinline__ void addn(uint32_t* d, const uint32_t* s) { coop_addn12(d, s); }
    static __device__ __forceinline__ void subn(uint32_t* d, const uint32_t* s) { coop_subn12(d, s); }
    static __device__ __forceinline__ void add_p(uint32_t* d) { coop_add_p_bls(d); }
    static __device__ __forceinline__ void redc(uint32_t* r, uint32_t* t) { coop_redc_bls(r, t); }
    template <int K> static __device__ __forceinline__ uint32_t sub_kp(uint32_t* d, const uint32_t* r) {
        if (K == 8) return coop_sub_8p_bls(d, r);
        if (K == 4) return coop_sub_4p_bls(d, r);
        if (K == 2) return coop_sub_2p_bls(d, r);
        return coop_sub_1p_bls(d, r);
    }
};

// BN254: 8 limbs, Montgomery radix 2^256 as everywhere else.  R - p leaves no head-room for lazy sums and xi = 9 + u is not
// a sign choice, so (i) the accumulators use all 17 words, xi multiplies the ACCUMULATORS (instruction XI, after the
// wrap-around products of a coefficient), (ii) the reduction returns 9 words (< 128 p) and canonicalises in up to 7 steps.
template <> struct Coop<Bn> {
    static constexpr int N = 8;
    static constexpr int GROUPS = BBS_COOP_GROUPS_BN;
    static constexpr int MAXREG = BBS_COOP_MAXREG_BN;
    static constexpr int RW = 9;
    static constexpr int MAXK = 64;            // (bound of the reduction output: 128 p)
    static constexpr int PARK = 3;
    static constexpr bool ACC_XI = true;
    static constexpr bool POINT_RATIO = true;  // D-type twist: the line is normalised by 1/y, the points enter as (x/y, 1/y)
    // the lanes of a warp share ONE inversion (coop_warp_inverse) in the prologue (x / y, 1 / y: -2.2 %) but not for the INV
    // instruction: inside the interpreter loop the call makes ptxas re-allocate this kernel's 80 registers (+4.8 %)
    static constexpr bool WARP_INV = false;
    static constexpr bool WARP_INV_PRO = true;
    static constexpr bool REP_SMEM = true;
    static __device__ __forceinline__ const uint32_t* prog() { return COOP_PROG_BN; }
    static __device__ __forceinline__ const uint32_t* prog_off() { return COOP_PROG_OFF_BN; }
    static __device__ __forceinline__ const uint32_t* consts() { return COOP_CONSTS_BN; }
    static __device__ __forceinline__ const uint32_t* kp() { return COOP_KP_BN; }
    static __device__ __forceinline__ void wmul_e(uint32_t* w, const uint32_t* a, const uint32_t* b) { coop_wmul_e8(w, a, b); }
    static __device__ __forceinline__ void wmul_o(uint32_t* w, const uint32_t* a, const uint32_t* b) { coop_wmul_o8(w, a, b); }
    static __device__ __forceinline__ void merge(uint32_t* w, const uint32_t* v) { coop_merge8(w, v); }
    static __device__ __forceinline__ void add_e(uint32_t* acc, const uint32_t* w) { coop_acc_add_e8(acc, w); }
    static __device__ __forceinline__ void sub_e(uint32_t* acc, const uint32_t* w) { coop_acc_sub_e8(acc, w); }
    static __device__ __forceinline__ void add_hi(uint32_t* acc, const uint32_t* z) { coop_acc_add_hi8(acc, z); }
    static __device__ __forceinline__ void sub_hi(uint32_t* acc, const uint32_t* z) { coop_acc_sub_hi8(acc, z); }
    static __device__ __forceinline__ void addn(uint32_t* d, const uint32_t* s) { coop_addn8(d, s); }
    static __device__ __forceinline__ void subn(uint32_t* d, const uint32_t* s) { coop_subn8(d, s); }
    static __device__ __forceinline__ void add_p(uint32_t* d) { coop_add_p_bn(d); }
    static __device__ __forceinline__ void redc(uint32_t* r, uint32_t* t) { coop_redc_wide_bn(r, t); }
    static constexpr int KPW = 9;              // KP[k] = k p R (k <= 97: nine words of k p), added to acc[8..16]
    static __device__ __forceinline__ void add_kp(uint32_t* acc, const uint32_t* k) {
        asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, %10; addc.cc.u32 %2, %2, %11; addc.cc.u32 %3, %3, %12; addc.cc.u32 %4, %4, %13; "
            "addc.cc.u32 %5, %5, %14; addc.cc.u32 %6, %6, %15; addc.cc.u32 %7, %7, %16; addc.u32 %8, %8, %17;"
            : "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11]), "+r"(acc[12]), "+r"(acc[13]), "+r"(acc[14]), "+r"(acc[15]), "+r"(acc[16])
            : "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]), "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]), "r"(k[8]));
    }
    template <int K> static __device__ __forceinline__ uint32_t sub_kp(uint32_t* d, const uint32_t* r) {
        if (K == 64) return coop_sub_64p_bn9(d, r);
        if (K == 32) return coop_sub_32p_bn9(d, r);
        if (K == 16) return coop_sub_16p_bn9(d, r);
        if (K == 8) return coop_sub_8p_bn9(d, r);
        if (K == 4) return coop_sub_4p_bn9(d, r);
        if (K == 2) return coop_sub_2p_bn9(d, r);
        return coop_sub_1p_bn9(d, r);
    }
    // w (9 words, < 128 p) -> w mod p by a quotient estimate: q^ = floor(top 29 bits of w * floor(2^32 / (top 22 bits of
    // p + 1)) / 2^32) is q or q - 1 (q = floor(w / p) < 128; bound in tools/gen_coop.py check_canon_q), so one multiply-
    // subtract by q^ p and ONE conditional subtraction replace the 5..7 conditional subtractions of the step ladder
    static constexpr bool QCANON = true;
    static constexpr int QCANON_MIN = 2;
    static constexpr int QTAB_UINT4 = 0;       // no table: q^ p is computed (q < 128)
    static __device__ __forceinline__ const uint32_t* qtab() { return nullptr; }
    static __device__ __forceinline__ void canon_q(uint32_t* w, const uint4*) {
        const uint32_t* p = BN_FP_P();
        const uint32_t v = (w[8] << 24) | (w[7] >> 8);
        const uint32_t q = __umulhi(v, COOP_BN_QCANON_M);
        uint32_t t[9];
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { c += (uint64_t)p[i] * q; t[i] = (uint32_t)c; c >>= 32; }
        t[8] = (uint32_t)c;
        asm("sub.cc.u32 %0, %0, %9; subc.cc.u32 %1, %1, %10; subc.cc.u32 %2, %2, %11; subc.cc.u32 %3, %3, %12; subc.cc.u32 %4, %4, %13; "
            "subc.cc.u32 %5, %5, %14; subc.cc.u32 %6, %6, %15; subc.cc.u32 %7, %7, %16; subc.u32 %8, %8, %17;"
            : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8])
            : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]));
        uint32_t d[9];
        const uint32_t b = coop_sub_1p_bn9(d, w);
#pragma unroll
        for (int i = 0; i < 9; i++) w[i] = b ? w[i] : d[i];
    }
    // (R, I) <- xi (R, I) = (9R - I, 9I + R) on the 17-word two's-complement accumulators
    static __device__ __forceinline__ void xi(uint32_t* R, uint32_t* I) {
        uint32_t t[17], u[17];
#pragma unroll
        for (int i = 16; i > 0; i--) { t[i] = __funnelshift_l(R[i - 1], R[i], 3); u[i] = __funnelshift_l(I[i - 1], I[i], 3); }
        t[0] = R[0] << 3; u[0] = I[0] << 3;
        coop_acc_add_f8(t, R);
        coop_acc_add_f8(u, I);
        coop_acc_sub_f8(t, I);
        coop_acc_add_f8(u, R);
#pragma unroll
        for (int i = 0; i < 17; i++) { R[i] = t[i]; I[i] = u[i]; }
    }
};

// ---- shared-memory cells: cell c = [comp 2][quad N/4][lane 32] uint4 ---------------------------------------------
template <int N> __device__ __forceinline__ void coop_load(uint32_t* x, const uint4* p) {
#pragma unroll
    for (int q = 0; q < N / 4; q++) {
        uint4 v = p[q * 32];
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
}
template <int N> __device__ __forceinline__ void coop_store(uint4* p, const uint32_t* x) {
#pragma unroll
    for (int q = 0; q < N / 4; q++) p[q * 32] = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}
template <int N> __device__ __forceinline__ void coop_load_g(uint32_t* x, const uint32_t* p) {
#pragma unroll
    for (int q = 0; q < N / 4; q++) {
        uint4 v = __ldg((const uint4*)p + q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
}
template <int N> __device__ __forceinline__ void coop_shl(uint32_t* x, int s) {
#pragma unroll
    for (int i = N - 1; i > 0; i--) x[i] = __funnelshift_l(x[i - 1], x[i], s);
    x[0] <<= s;
}

// operand in one of the four forms of the instruction set (tools/coop_prog.py): c0, c1, c0 + c1, c0 - c1 + p
template <class C, bool GLOBAL> __device__ __forceinline__ void coop_operand(uint32_t* x, const void* cellp, int form, int shift) {
    constexpr int N = Coop<C>::N;
    uint32_t t[N];
    if (GLOBAL) {
        const uint32_t* g = (const uint32_t*)cellp;
        coop_load_g<N>(x, g + (form == 1 ? N : 0));
        if (form >= 2) {
            coop_load_g<N>(t, g + N);
            if (form == 2) Coop<C>::addn(x, t); else { Coop<C>::subn(x, t); Coop<C>::add_p(x); }
        }
    } else {
        const uint4* s = (const uint4*)cellp;
        coop_load<N>(x, s + (form == 1 ? (N / 4) * 32 : 0));
        if (form >= 2) {
            coop_load<N>(t, s + (N / 4) * 32);
            if (form == 2) Coop<C>::addn(x, t); else { Coop<C>::subn(x, t); Coop<C>::add_p(x); }
        }
    }
    if (shift) coop_shl<N>(x, shift);
}

// o = v^-1 for the 32 values of a role-warp (one per lane = item; 0 -> 0 as the Fermat inversion gives), all 32 lanes
// calling.  A SIMT warp pays for a Fermat ladder (381 squarings + 95 multiplications) whether one lane needs it or all, so
// the lanes share ONE inversion instead: prefix and suffix products by shuffles (Montgomery's trick, 12 multiplications per
// lane), the total inverted by the variable-time binary Euclid -- on the SAME value in every lane, hence without
// divergence, and the data are public (pairing values of public inputs).  ~3.8x cheaper than 32 ladders in lock-step.
template <class C> __device__ __noinline__ void coop_warp_inverse(uint32_t* o, const uint32_t* v) {
    using F = typename C::Fp;
    constexpr int N = F::N;
    const int lane = threadIdx.x & 31;
    BBS_A16 uint32_t x[N], pre[N], suf[N], t[N], u[N];
    const bool z = bn_is_zero<N>(v);
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = z ? F::ONE()[i] : v[i]; pre[i] = x[i]; suf[i] = x[i]; }
    for (int s = 1; s < 32; s <<= 1) {
#pragma unroll
        for (int i = 0; i < N; i++) { t[i] = __shfl_up_sync(0xffffffffu, pre[i], s); u[i] = __shfl_down_sync(0xffffffffu, suf[i], s); }
        if (lane < s) bn_copy<N>(t, F::ONE());
        if (lane + s > 31) bn_copy<N>(u, F::ONE());
        fe_mul<F>(pre, pre, t);                      // pre = x_0 .. x_lane
        fe_mul<F>(suf, suf, u);                      // suf = x_lane .. x_31
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = __shfl_sync(0xffffffffu, pre[i], 31);                  // the product of all 32
        t[i] = __shfl_up_sync(0xffffffffu, pre[i], 1);                // x_0 .. x_{lane-1}
        x[i] = __shfl_down_sync(0xffffffffu, suf[i], 1);              // x_{lane+1} .. x_31
    }
    if (lane == 0) bn_copy<N>(t, F::ONE());
    if (lane == 31) bn_copy<N>(x, F::ONE());
    fe_inv_vt<F>(pre, u);
    fe_mul<F>(t, t, x);
    fe_mul<F>(t, t, pre);
#pragma unroll
    for (int i = 0; i < N; i++) o[i] = z ? 0u : t[i];
}

// both output components: (accR, accI) -> canonical Fp2.  The two reductions are independent; they are written
// step by step side by side (one basic block per step) so that their dependency chains overlap.
template <class C> __device__ __forceinline__ void coop_finish2(uint32_t* r0, uint32_t* r1, uint32_t* R, uint32_t* I,
                                                                uint32_t ins, const uint4* zp, const uint4* qtab) {
    constexpr int N = Coop<C>::N;
    constexpr int Q = N / 4;
    if ((ins >> 10) & 1) {                       // triple
        uint32_t t[2 * N + 1], u[2 * N + 1];
#pragma unroll
        for (int i = 2 * N; i > 0; i--) { t[i] = __funnelshift_l(R[i - 1], R[i], 1); u[i] = __funnelshift_l(I[i - 1], I[i], 1); }
        t[0] = R[0] << 1; u[0] = I[0] << 1;
        Coop<C>::add_e(R, t);                    // low 2N words + carry into the top word
        Coop<C>::add_e(I, u);
        R[2 * N] += t[2 * N];
        I[2 * N] += u[2 * N];
    }
    const uint32_t zs = (ins >> 11) & 3;
    if (zs) {
        uint32_t z0[N], z1[N];
        coop_load<N>(z0, zp);
        coop_load<N>(z1, zp + Q * 32);
        if ((ins >> 13) & 1) { coop_shl<N>(z0, 1); coop_shl<N>(z1, 1); }
        if (zs == 1) { Coop<C>::add_hi(R, z0); Coop<C>::add_hi(I, z1); } else { Coop<C>::sub_hi(R, z0); Coop<C>::sub_hi(I, z1); }
    }
    const uint32_t kp = ((ins >> 22) & 3) | (((ins >> 31) & 1) << 2);
    if (kp) {
        constexpr int KPW = Coop<C>::KPW;
        uint32_t k[KPW];
        const uint32_t* kt = Coop<C>::kp() + kp * KPW;
#pragma unroll
        for (int i = 0; i < KPW; i++) k[i] = kt[i];
        Coop<C>::add_kp(R, k);
        Coop<C>::add_kp(I, k);
    }
    constexpr int RW = Coop<C>::RW;
    uint32_t w0[RW], w1[RW];
    Coop<C>::redc(w0, R);
    Coop<C>::redc(w1, I);
    const uint32_t canon = ((ins >> 24) & 3) | (((ins >> 30) & 1) << 2);
    uint32_t d0[RW], d1[RW], b0, b1;
#define COOP_CANON_STEP(K)                                                                     \
    b0 = Coop<C>::template sub_kp<K>(d0, w0); b1 = Coop<C>::template sub_kp<K>(d1, w1);         \
    _Pragma("unroll") for (int i = 0; i < RW; i++) { w0[i] = b0 ? w0[i] : d0[i]; w1[i] = b1 ? w1[i] : d1[i]; }
    if (Coop<C>::QCANON && canon >= Coop<C>::QCANON_MIN) {
        Coop<C>::canon_q(w0, qtab);
        Coop<C>::canon_q(w1, qtab);
    } else {
        if constexpr (!Coop<C>::QCANON) {
            if (canon >= 3) { COOP_CANON_STEP(8) }
            if (canon >= 2) { COOP_CANON_STEP(4) }
        }
        if (!(Coop<C>::QCANON && Coop<C>::QCANON_MIN <= 1) && canon >= 1) { COOP_CANON_STEP(2) }
        COOP_CANON_STEP(1)
    }
#undef COOP_CANON_STEP
#pragma unroll
    for (int i = 0; i < N; i++) { r0[i] = w0[i]; r1[i] = w1[i]; }
}

// ONE output component (the SPLIT = 2 kernel: each warp of a role pair reduces one accumulator): M -> canonical Fp
template <class C> __device__ __forceinline__ void coop_finish1(uint32_t* r, uint32_t* M, uint32_t ins, const uint4* zp, const uint4* qtab) {
    constexpr int N = Coop<C>::N;
    if ((ins >> 10) & 1) {                       // triple
        uint32_t t[2 * N + 1];
#pragma unroll
        for (int i = 2 * N; i > 0; i--) t[i] = __funnelshift_l(M[i - 1], M[i], 1);
        t[0] = M[0] << 1;
        Coop<C>::add_e(M, t);
        M[2 * N] += t[2 * N];
    }
    const uint32_t zs = (ins >> 11) & 3;
    if (zs) {
        uint32_t z[N];
        coop_load<N>(z, zp);
        if ((ins >> 13) & 1) coop_shl<N>(z, 1);
        if (zs == 1) Coop<C>::add_hi(M, z); else Coop<C>::sub_hi(M, z);
    }
    const uint32_t kp = ((ins >> 22) & 3) | (((ins >> 31) & 1) << 2);
    if (kp) {
        constexpr int KPW = Coop<C>::KPW;
        uint32_t k[KPW];
        const uint32_t* kt = Coop<C>::kp() + kp * KPW;
#pragma unroll
        for (int i = 0; i < KPW; i++) k[i] = kt[i];
        Coop<C>::add_kp(M, k);
    }
    constexpr int RW = Coop<C>::RW;
    uint32_t w[RW];
    Coop<C>::redc(w, M);
    const uint32_t canon = ((ins >> 24) & 3) | (((ins >> 30) & 1) << 2);
    uint32_t d[RW], b;
#define COOP_CANON_STEP(K)                                                                     \
    b = Coop<C>::template sub_kp<K>(d, w);                                                     \
    _Pragma("unroll") for (int i = 0; i < RW; i++) w[i] = b ? w[i] : d[i];
    if (Coop<C>::QCANON && canon >= Coop<C>::QCANON_MIN) {
        Coop<C>::canon_q(w, qtab);
    } else {
        if constexpr (!Coop<C>::QCANON) {
            if (canon >= 3) { COOP_CANON_STEP(8) }
            if (canon >= 2) { COOP_CANON_STEP(4) }
        }
        if (!(Coop<C>::QCANON && Coop<C>::QCANON_MIN <= 1) && canon >= 1) { COOP_CANON_STEP(2) }
        COOP_CANON_STEP(1)
    }
#undef COOP_CANON_STEP
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = w[i];
}

// MULTI: every lane (item) reads the line constants of its own issuer (issuer sets, kernels.cuh); the loads stay 128-bit
// but are no longer warp-uniform, everything else is identical.
// SPLIT = 2 (small batches: one block of up to 32 items, TWELVE warps): a role is a PAIR of warps.  With a handful of
// items a role-warp has one or a few active lanes and the kernel is pure dependent-issue latency, so the work of a role is
// cut in two: every EP of the stream carries a half tag (bit 31, tools/coop_prog.py) and is executed by that half only;
// at a FIN the halves exchange one accumulator each through shared memory (half 0 ends up with the whole real sum, half
// 1 with the whole imaginary sum; half 1 routes the EP signs crosswise, so "its" sum is always the first array), reduce ONE
// component each and meet at a pair barrier before either reads the cell.  Same program, same results, about half the
// latency.
template <class C, bool MULTI, int SPLIT>
__global__ void __maxnreg__(Coop<C>::MAXREG) pairing_coop_kernel(const CoopArgs a) {
    constexpr int COOP_GROUPS = SPLIT == 2 ? 1 : Coop<C>::GROUPS;
    constexpr int GW = COOP_ROLES * SPLIT;     // warps per group
    constexpr int N = Coop<C>::N;
    constexpr int Q = N / 4;
    constexpr int CELL = 2 * Q * 32;           // uint4 per cell
    extern __shared__ uint4 smem_all[];
    const int lane = threadIdx.x & 31, wig = (threadIdx.x >> 5) % GW, role = wig % COOP_ROLES;
    const int half = SPLIT == 2 ? wig / COOP_ROLES : 0;
    const int group = threadIdx.x / (GW * 32);
    uint4* smem = smem_all + (size_t)group * (COOP_CELLS * CELL + COOP_ROLES * 32 / 4 + GW);
    uint32_t* votes = (uint32_t*)(smem + COOP_CELLS * CELL);      // [role][lane]
    const uint32_t gblock = blockIdx.x * COOP_GROUPS + group;      // 32-item group index
    const uint32_t item = gblock * COOP_ITEMS + lane;
#define COOP_BAR() asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GW * 32) : "memory")
#define COOP_PAIR_BAR() do { if constexpr (SPLIT == 2) asm volatile("bar.sync %0, 64;" ::"r"(8 + role) : "memory"); } while (0)
    const uint4* qtab = smem_all + (size_t)COOP_GROUPS * (COOP_CELLS * CELL + COOP_ROLES * 32 / 4 + GW);
    if constexpr (Coop<C>::QTAB_UINT4 > 0) {
        // q p table of canon_q: written by the first threads of every group with the same values (a group whose items are
        // all beyond n leaves below, so no group may depend on another one's writes); ordered by the group barrier below
        if ((int)(threadIdx.x % (GW * 32)) < Coop<C>::QTAB_UINT4)
            ((uint4*)qtab)[threadIdx.x % (GW * 32)] = ((const uint4*)Coop<C>::qtab())[threadIdx.x % (GW * 32)];
    }
    // SPLIT = 2: exchange buffer of the role pairs, [role][half][word][lane]
    uint32_t* xbuf = (uint32_t*)(qtab + Coop<C>::QTAB_UINT4);
    (void)xbuf;
    if (gblock * COOP_ITEMS >= a.n) return;          // a whole group without items (its named barrier is its own)
    const bool valid = item < a.n;
    const uint32_t fl = valid ? a.flags[item] : (uint32_t)(FL_DONE | FL_SKIP0 | FL_SKIP1);
    uint4* cells = smem + lane;
    const uint32_t* lines = a.lines;
    if constexpr (MULTI) lines += (size_t)((valid && !(fl & FL_DONE)) ? a.item_issuer[item] : 0u) * a.line_stride;

    // prologue: f = 1 in slot 0 (cells 0..5), P0 -> cell 16, P1 -> cell 17 (x in c0, y in c1); by the first half of a pair
    {
        uint32_t v[N];
        if (half == 0) {
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = (role == 0) ? C::Fp::ONE()[i] : 0u;
            coop_store<N>(cells + role * CELL, v);
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = 0u;
            coop_store<N>(cells + role * CELL + Q * 32, v);
            if (role < 4) {
                // role r loads coordinate (r & 1) of point (r >> 1)
                const uint32_t* src = a.pair + (size_t)(valid ? item : 0) * (6 * N) + (role >> 1) * 3 * N + (role & 1) * N;
#pragma unroll
                for (int i = 0; i < N; i++) v[i] = valid ? src[i] : 0u;
                coop_store<N>(cells + (16 + (role >> 1)) * CELL + (role & 1) * Q * 32, v);
            }
        }
        if constexpr (Coop<C>::POINT_RATIO) {
            // (x, y) -> (x / y, 1 / y), one point per role-warp (y != 0 on a curve of odd order; skipped pairs hold zeros)
            COOP_BAR();
            if (role < 2 && half == 0) {
                BBS_A16 uint32_t x[N], y[N], yi[N];
                coop_load<N>(x, cells + (16 + role) * CELL);
                coop_load<N>(y, cells + (16 + role) * CELL + Q * 32);
                if constexpr (Coop<C>::WARP_INV_PRO) coop_warp_inverse<C>(yi, y); else fe_inv<typename C::Fp>(yi, y);
                fe_mul<typename C::Fp>(x, x, yi);
                coop_store<N>(cells + (16 + role) * CELL, x);
                coop_store<N>(cells + (16 + role) * CELL + Q * 32, yi);
            }
        }
    }
    COOP_BAR();

    const uint32_t* prog = Coop<C>::prog() + Coop<C>::prog_off()[role];
    uint32_t pc = 0, line = 0;
    // REP / ENDREP state (two nesting levels): in registers, or in shared memory (Coop<C>::REP_SMEM) -- it is touched ~1,000
    // times per item against ~23,000 EP / FIN.  Measured: BN254 (80-register cap) -1.0 % in shared memory; BLS12-381 +3.7 %
    // although its 8-byte spill in the EP path disappears (ptxas schedules the loop differently), so it keeps registers.
    volatile uint32_t* rep = votes + COOP_ROLES * 32 + wig * 4;        // [pc0, cnt0, pc1, cnt1] of this warp
    uint32_t rep_pc[2] = {0, 0}, rep_cnt[2] = {0, 0};
    int rep_sp = 0;
    // R, I: real / imaginary accumulator (SPLIT = 2, half 1: the other way round -- R is always the one this warp reduces)
    uint32_t R[2 * N + 1], I[2 * N + 1];
#pragma unroll
    for (int i = 0; i <= 2 * N; i++) { R[i] = 0; I[i] = 0; }
    uint32_t ins = __ldg(prog);
    for (;;) {
        // every lane of a role-warp holds the same word; the redux makes that visible to the compiler (uniform register),
        // so the decode runs on the uniform datapath and the branches below need no reconvergence bookkeeping
        const uint32_t cur = __reduce_or_sync(0xffffffffu, ins);
        pc++;
        ins = __ldg(prog + pc);                  // prefetch (streams end with END, one word of slack is harmless)
        const uint32_t kind = cur & 3;
        if (kind == 0) {
            // ---- EP ------------------------------------------------------------------------------------------
            if constexpr (SPLIT == 2) { if (((cur >> 31) & 1) != (uint32_t)half) continue; }       // the other half's product
            uint32_t x[N], y[N], w[2 * N], v[2 * N];
            coop_operand<C, false>(x, cells + ((cur >> 6) & 255) * CELL, (cur >> 14) & 3, (cur >> 16) & 3);
            const uint32_t yc = (cur >> 18) & 255;
            if ((cur >> 30) & 1) {
                const uint32_t* g = yc >= 128 ? lines + ((size_t)line * 4 + (yc - 128)) * (2 * N)
                                              : Coop<C>::consts() + yc * (2 * N);
                coop_operand<C, true>(y, g, (cur >> 26) & 3, (cur >> 28) & 3);
            } else {
                coop_operand<C, false>(y, cells + yc * CELL, (cur >> 26) & 3, (cur >> 28) & 3);
            }
            uint32_t sR = (cur >> 2) & 3, sI = (cur >> 4) & 3;
            if constexpr (SPLIT == 2) { if (half) { const uint32_t t = sR; sR = sI; sI = t; } }
            // the two half products are independent carry chains: issued back to back they interleave on the IMAD pipe
            Coop<C>::wmul_e(w, x, y);
            Coop<C>::wmul_o(v, x, y);
            Coop<C>::merge(w, v);            // full product: one 2N-word addition per accumulator instead of two
            if (sR == 1) Coop<C>::add_e(R, w); else if (sR == 2) Coop<C>::sub_e(R, w);
            if (sI == 1) Coop<C>::add_e(I, w); else if (sI == 2) Coop<C>::sub_e(I, w);
        } else if (kind == 3) {
            // ---- XI: accumulators *= xi (curves whose xi is not a sign choice) ----------------------------------
            // (linear, so each half of a SPLIT = 2 pair applies it to its partial sums; half 1 holds them crosswise)
            if constexpr (Coop<C>::ACC_XI) { if (SPLIT == 2 && half) Coop<C>::xi(I, R); else Coop<C>::xi(R, I); }
        } else if (kind == 1) {
            // ---- FIN -----------------------------------------------------------------------------------------
            const uint4* zp = cells + ((cur >> 14) & 255) * CELL;
            uint4* dp = cells + ((cur >> 2) & 255) * CELL;
            const uint32_t sk = (cur >> 27) & 3;
            const bool zero = ((sk & 1) && (fl & ((sk & 2) ? FL_SKIP1 : FL_SKIP0))), fp_only = (cur >> 29) & 1;
            if constexpr (SPLIT == 2) {
                // hand the accumulator this warp does not reduce to its partner, take the partner's share of the own one
                uint32_t* xo = xbuf + ((role * 2 + half) * (2 * N + 1)) * 32 + lane;
#pragma unroll
                for (int i = 0; i <= 2 * N; i++) xo[i * 32] = I[i];
                COOP_PAIR_BAR();
                const uint32_t* xi_ = xbuf + ((role * 2 + (half ^ 1)) * (2 * N + 1)) * 32 + lane;
                uint32_t t[2 * N + 1];
#pragma unroll
                for (int i = 0; i <= 2 * N; i++) t[i] = xi_[i * 32];
                Coop<C>::add_e(R, t);
                R[2 * N] += t[2 * N];
                uint32_t r[N];
                coop_finish1<C>(r, R, cur, zp + half * Q * 32, qtab);
                if ((sk & 1) || fp_only) {
#pragma unroll
                    for (int i = 0; i < N; i++) r[i] = (zero || (fp_only && half)) ? 0u : r[i];
                }
                coop_store<N>(dp + half * Q * 32, r);
            } else {
                uint32_t r0[N], r1[N];
                coop_finish2<C>(r0, r1, R, I, cur, zp, qtab);
                if ((sk & 1) || fp_only) {
#pragma unroll
                    for (int i = 0; i < N; i++) { r0[i] = zero ? 0u : r0[i]; r1[i] = (zero || fp_only) ? 0u : r1[i]; }
                }
                coop_store<N>(dp, r0);
                coop_store<N>(dp + Q * 32, r1);
            }
#pragma unroll
            for (int i = 0; i <= 2 * N; i++) { R[i] = 0; I[i] = 0; }
            COOP_PAIR_BAR();                 // both components of the cell are visible to both halves; the exchange buffer is free
            if ((cur >> 26) & 1) COOP_BAR();
        } else {
            // ---- CTL -----------------------------------------------------------------------------------------
            const uint32_t sub = (cur >> 2) & 15, arg = cur >> 6;
            if (sub == 0) break;                                           // END
            if (sub == 1) {                                                // REP
                if constexpr (Coop<C>::REP_SMEM) { if (lane == 0) { rep[2 * rep_sp] = pc; rep[2 * rep_sp + 1] = arg; } __syncwarp(); }
                else { rep_pc[rep_sp] = pc; rep_cnt[rep_sp] = arg; }
                rep_sp++;
            }
            else if (sub == 2) {                                           // ENDREP
                if constexpr (Coop<C>::REP_SMEM) {
                    const uint32_t left = rep[2 * rep_sp - 1] - 1;
                    __syncwarp();
                    if (left > 0) { if (lane == 0) rep[2 * rep_sp - 1] = left; pc = rep[2 * rep_sp - 2]; ins = __ldg(prog + pc); } else rep_sp--;
                    __syncwarp();
                } else {
                    if (--rep_cnt[rep_sp - 1] > 0) { pc = rep_pc[rep_sp - 1]; ins = __ldg(prog + pc); } else rep_sp--;
                }
            }
            else if (sub == 3) line++;                                     // NEXTLINE
            else if (sub == 4) COOP_BAR();                            // BAR
            else if (sub == 5 || sub == 6) {                               // GSAVE / GLOAD (own cell <-> global)
                if (half == 0) {
                    uint4* g = (uint4*)a.gscratch + (((size_t)gblock * Coop<C>::PARK + (arg >> 8)) * COOP_ROLES + role) * CELL + lane;
                    uint4* c = cells + (arg & 255) * CELL;
#pragma unroll
                    for (int q = 0; q < 2 * Q; q++) { if (sub == 5) g[q * 32] = c[q * 32]; else c[q * 32] = g[q * 32]; }
                }
                COOP_PAIR_BAR();
            }
            else if (sub == 8) {                                           // INV: cell.c0 = cell.c0^-1 (one inversion per warp)
                if (half == 0) {
                    BBS_A16 uint32_t v[N], o[N];
                    coop_load<N>(v, cells + arg * CELL);
                    if constexpr (Coop<C>::WARP_INV) coop_warp_inverse<C>(o, v); else fe_inv<typename C::Fp>(o, v);
                    coop_store<N>(cells + arg * CELL, o);
                }
                // INV only ever follows a FIN (tools/coop_prog.py fp_inverse): the accumulators are zero here.  Saying so
                // makes them dead across the call, so that nothing of the hot loop's state has to be spilled around it.
                if constexpr (Coop<C>::WARP_INV) {
#pragma unroll
                    for (int i = 0; i <= 2 * N; i++) { R[i] = 0; I[i] = 0; }
                }
                COOP_PAIR_BAR();
            }
            else if (sub == 7) {                                           // CHECK: result == 1 ?
                if (half == 0) {
                    uint32_t v[N], o = 0;
                    coop_load<N>(v, cells + arg * CELL);
#pragma unroll
                    for (int i = 0; i < N; i++) o |= v[i] ^ ((role == 0) ? C::Fp::ONE()[i] : 0u);
                    coop_load<N>(v, cells + arg * CELL + Q * 32);
#pragma unroll
                    for (int i = 0; i < N; i++) o |= v[i];
                    votes[role * 32 + lane] = o;
                }
                COOP_BAR();
                if (role == 0 && half == 0) {
                    uint32_t all = 0;
#pragma unroll
                    for (int k = 0; k < COOP_ROLES; k++) all |= votes[k * 32 + lane];
                    if (valid && !(fl & FL_DONE)) a.status[item] = all == 0 ? ST_ACCEPT : ST_REJECT;
                }
            }
        }
    }
}

template <class C> constexpr size_t coop_gscratch_bytes(size_t n) {
    constexpr int COOP_GROUPS = Coop<C>::GROUPS;
    const size_t per_block = (size_t)COOP_ITEMS * COOP_GROUPS;
    return ((n + per_block - 1) / per_block) * COOP_GROUPS * Coop<C>::PARK * COOP_ROLES * (2 * (Coop<C>::N / 4) * 32) * sizeof(uint4);
}
template <class C, int SPLIT = 1> constexpr size_t coop_smem_bytes() {
    constexpr size_t groups = SPLIT == 2 ? 1 : Coop<C>::GROUPS, gw = COOP_ROLES * SPLIT;
    return groups * ((size_t)COOP_CELLS * (2 * (Coop<C>::N / 4) * 32) * sizeof(uint4) + COOP_ROLES * 32 * sizeof(uint32_t) +
                     gw * 4 * sizeof(uint32_t)) +
           Coop<C>::QTAB_UINT4 * sizeof(uint4) +
           (SPLIT == 2 ? (size_t)COOP_ROLES * 2 * (2 * Coop<C>::N + 1) * 32 * sizeof(uint32_t) : 0);
}

}  // namespace bbs
#endif  // __CUDACC__
