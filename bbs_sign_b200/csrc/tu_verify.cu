// Translation unit for the `verify` kernels (see launchers.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN;
// with neither macro both curves are instantiated (host-simulation build).
#include "launchers.cuh"

// occupancy knobs (threads per block, min resident blocks per SM -> register cap); -D overridable
#ifndef BBS_VERIFY_G1_TPB
#define BBS_VERIFY_G1_TPB 256
#endif
#ifndef BBS_VERIFY_G1_MINB
#define BBS_VERIFY_G1_MINB 2
#endif

namespace bbs {

template <class C> int launch_verify_g1(const VerifyG1Args& a, uint32_t n, rt_stream_t s) {
#ifdef BBS_HOSTSIM
    return rt_launch<VerifyG1Args, &verify_g1_item<C>, BBS_VERIFY_G1_TPB, BBS_VERIFY_G1_MINB>(a, n, s);
#else
    if (n == 0) return 0;
    verify_g1_kernel<C, BBS_VERIFY_G1_TPB, BBS_VERIFY_G1_MINB><<<(n + BBS_VERIFY_G1_TPB - 1) / BBS_VERIFY_G1_TPB, BBS_VERIFY_G1_TPB, 0, s>>>(a, n);
    RT_CHECK(cudaGetLastError());
    return 0;
#endif
}

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_verify_g1<Bls>(const VerifyG1Args&, uint32_t, rt_stream_t);
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_verify_g1<Bn>(const VerifyG1Args&, uint32_t, rt_stream_t);
#endif

}  // namespace bbs
