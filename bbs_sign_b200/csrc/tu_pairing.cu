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

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_pairing<Bls>(const PairingArgs&, uint32_t, rt_stream_t);
#ifndef BBS_HOSTSIM
size_t coop_gscratch_bytes_bls(size_t n) {
    const size_t per_block = (size_t)COOP_ITEMS * COOP_GROUPS;
    return ((n + per_block - 1) / per_block) * COOP_GROUPS * COOP_ROLES * (2 * (Coop<Bls>::N / 4) * 32) * sizeof(uint4);
}
int launch_pairing_coop_bls(const CoopArgs& a, rt_stream_t s) {
    if (a.n == 0) return 0;
    constexpr size_t smem = coop_smem_bytes<Bls>();
    RT_CHECK(cudaFuncSetAttribute(pairing_coop_kernel<Bls>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RT_CHECK(cudaFuncSetAttribute(pairing_coop_kernel<Bls>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const uint32_t per_block = COOP_ITEMS * COOP_GROUPS;
    pairing_coop_kernel<Bls><<<(a.n + per_block - 1) / per_block, COOP_TPB, smem, s>>>(a);
    RT_CHECK(cudaGetLastError());
    return 0;
}
#endif
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_pairing<Bn>(const PairingArgs&, uint32_t, rt_stream_t);
#endif

}  // namespace bbs
