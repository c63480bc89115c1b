// Random-linear-combination batch verification (the north star's optional mode): n signatures
// under one issuer key are accepted together iff
//     e( sum r_i A_i , W ) * e( sum r_i (e_i A_i - B_i) , BP2 ) == 1,      B_i = K + sum_j m_ij H_j  (verify.rs:81-92)
// with public-coin 128-bit coefficients r_i = first 16 bytes (big-endian integer) of SHA-256(seed || BE64(i)), r_i = 1
// if that is 0.  A batch-level, probabilistic verdict (a false accept needs a 2^-128 event over the seed); the per-item
// entry points remain the reference-exact path.  The B_i terms collapse to L+1 fixed-base products:
//     sum r_i B_i = (sum r_i) K + sum_j (sum_i r_i m_ij) H_j
// so the per-item work is one decompression and the item's share of two bucket MSMs (rlc_msm.cuh).  Every GPU reduces
// its shard to two G1 points; the partial points of all shards are added and checked with ONE pairing product by
// rlc_combine_kernel (no collective: 2 compressed points per GPU travel through the host).
#pragma once
#include "kernels.cuh"

#ifdef __CUDACC__
namespace bbs {

constexpr int RLC_TPB = 128;

struct RlcArgs {
    CtxView ctx;
    const uint8_t* sigs;        // n x (G1 compressed || LE32 e)
    const uint8_t* scalars;     // n x n_msgs x LE32
    uint32_t n_msgs, n;
    uint64_t index_base;        // global index of item 0 of this shard (coefficients depend on the global index)
    BBS_A16 uint32_t seed[8];           // the 32-byte seed as big-endian words
    uint32_t* sc_part;          // [blocks][n_msgs + 1][8]  canonical Fr limbs
    uint32_t* bad;              // != 0: some item was malformed
};

// r_i (8 limbs, upper 4 zero)
__device__ __forceinline__ void rlc_coeff(uint32_t* r, const uint32_t* seed_be, uint64_t index) {
    Sha256 s;
    s.init();
    s.update_words(seed_be, 8);
    s.put_be64(index);
    BBS_A16 uint32_t h[8];
    s.finish(h);
    r[0] = h[3]; r[1] = h[2]; r[2] = h[1]; r[3] = h[0];
    for (int i = 4; i < 8; i++) r[i] = 0;
    if ((r[0] | r[1] | r[2] | r[3]) == 0) r[0] = 1;
}

// block-wide sum of one Jacobian point per thread (result in sp[0])
template <class C> __device__ __forceinline__ void rlc_block_sum_points(uint32_t (*sp)[3 * C::Fp::N], uint32_t* mine) {
    const int t = threadIdx.x;
    g1_copy<C>(sp[t], mine);
    __syncthreads();
    for (int s = RLC_TPB / 2; s >= 1; s >>= 1) {
        if (t < s) g1_add<C>(sp[t], sp[t], sp[t + s]);
        __syncthreads();
    }
}
// block-wide sum mod r of one Fr value per thread (result in ss[0])
template <class C> __device__ __forceinline__ void rlc_block_sum_fr(uint32_t (*ss)[8], const uint32_t* mine) {
    const int t = threadIdx.x;
    bn_copy<8>(ss[t], mine);
    __syncthreads();
    for (int s = RLC_TPB / 2; s >= 1; s >>= 1) {
        if (t < s) fe_add<typename C::Fr>(ss[t], ss[t], ss[t + s]);
        __syncthreads();
    }
}

struct RlcCombineArgs {
    CtxView ctx;
    const uint8_t* parts; uint32_t n_parts;      // n_parts x (comp(S1) || comp(S2))
    uint32_t* pair; uint32_t* flags; uint8_t* status;
};
// one block: the shards' compressed partial points are decompressed one per thread (a square root each, all at once),
// added per sum by a block tree, normalised in two warps and handed to the single pairing check
template <class C> __global__ void __launch_bounds__(RLC_TPB, 1) rlc_combine_kernel(const RlcCombineArgs a) {
    using F = typename C::Fp;
    __shared__ BBS_A16 uint32_t sp[RLC_TPB][3 * C::Fp::N];
    __shared__ BBS_A16 uint32_t S[2][3 * C::Fp::N];
    __shared__ uint32_t bad, skip[2];
    const CtxView& cx = a.ctx;
    const uint32_t t = threadIdx.x;
    if (t == 0) bad = 0;
    __syncthreads();
    BBS_A16 uint32_t acc[G1J];
    g1_set_inf<C>(acc);
    for (uint32_t idx = t; idx < 2 * a.n_parts; idx += RLC_TPB) {     // RLC_TPB is even: a thread only meets its own sum
        BBS_A16 uint32_t p[G1A];
        const int st = g1_decompress<C>(p, a.parts + (size_t)idx * C::G1_BYTES);
        if (st == PT_BAD) atomicOr(&bad, 1u);
        if (st == PT_OK) g1_add_mixed<C>(acc, acc, p);
    }
    for (uint32_t w = 0; w < 2; w++) {
        BBS_A16 uint32_t mine[G1J];
        if ((t & 1) == w) g1_copy<C>(mine, acc); else g1_set_inf<C>(mine);
        rlc_block_sum_points<C>(sp, mine);
        if (t == 0) g1_copy<C>(S[w], sp[0]);
        __syncthreads();
    }
    if (bad) {
        if (t == 0) { a.status[0] = ST_ERR_MALFORMED; a.flags[0] = FL_DONE; }
        return;
    }
    if ((t & 31) == 0 && (t >> 5) < 2) {
        const uint32_t w = t >> 5;
        BBS_A16 uint32_t aff[G1A];
        const bool fin = g1_to_affine_vt<C>(aff, S[w]);
        uint32_t* pr = a.pair + w * 3 * FPN;
        bn_copy<2 * C::Fp::N>(pr, aff);
        fe_set_one<F>(pr + 2 * FPN);
        skip[w] = fin ? 0u : 1u;
    }
    __syncthreads();
    if (t == 0) {
        a.flags[0] = ((skip[0] || cx.w_inf) ? FL_SKIP0 : 0u) | (skip[1] ? FL_SKIP1 : 0u);
        a.status[0] = ST_REJECT;
    }
}

}  // namespace bbs
#endif  // __CUDACC__
