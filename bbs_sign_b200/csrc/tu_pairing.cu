// Translation unit for the `pairing` kernels (see launchers.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN;
// with neither macro both curves are instantiated (host-simulation build).
#include "launchers.cuh"

// occupancy knobs (threads per block, min resident blocks per SM -> register cap); -D overridable
#ifndef BBS_PAIRING_TPB
#define BBS_PAIRING_TPB 128
#endif
#ifndef BBS_PAIRING_MINB
#define BBS_PAIRING_MINB 2
#endif

namespace bbs {

template <class C> int launch_pairing(const PairingArgs& a, uint32_t n, rt_stream_t s) {
    return rt_launch<PairingArgs, &pairing_item<C>, BBS_PAIRING_TPB, BBS_PAIRING_MINB>(a, n, s);
}

#ifndef BBS_HOSTSIM
template <class C> size_t coop_gscratch_size(size_t n) { return coop_gscratch_bytes<C>(n); }
template <class C, bool MULTI, int SPLIT> static int launch_pairing_coop_t(const CoopArgs& a, rt_stream_t s) {
    constexpr size_t smem = coop_smem_bytes<C, SPLIT>();
    RT_CHECK(cudaFuncSetAttribute(pairing_coop_kernel<C, MULTI, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RT_CHECK(cudaFuncSetAttribute(pairing_coop_kernel<C, MULTI, SPLIT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const uint32_t groups = SPLIT == 2 ? 1 : Coop<C>::GROUPS, per_block = COOP_ITEMS * groups;
    pairing_coop_kernel<C, MULTI, SPLIT><<<(a.n + per_block - 1) / per_block, COOP_ROLES * 32 * SPLIT * groups, smem, s>>>(a);
    RT_CHECK(cudaGetLastError());
    return 0;
}
template <class C> int launch_pairing_coop(const CoopArgs& a, rt_stream_t s) {
    if (a.n == 0) return 0;
    // a batch that fits one group is latency-bound: two warps per role (pairing_coop.cuh, SPLIT = 2)
    if (a.n <= COOP_ITEMS && a.n <= a.split2_max)
        return a.item_issuer ? launch_pairing_coop_t<C, true, 2>(a, s) : launch_pairing_coop_t<C, false, 2>(a, s);
    return a.item_issuer ? launch_pairing_coop_t<C, true, 1>(a, s) : launch_pairing_coop_t<C, false, 1>(a, s);
}
#endif

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_pairing<Bls>(const PairingArgs&, uint32_t, rt_stream_t);
#ifndef BBS_HOSTSIM
template size_t coop_gscratch_size<Bls>(size_t);
template int launch_pairing_coop<Bls>(const CoopArgs&, rt_stream_t);
#endif
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_pairing<Bn>(const PairingArgs&, uint32_t, rt_stream_t);
#ifndef BBS_HOSTSIM
template size_t coop_gscratch_size<Bn>(size_t);
template int launch_pairing_coop<Bn>(const CoopArgs&, rt_stream_t);
#endif
#endif

}  // namespace bbs
