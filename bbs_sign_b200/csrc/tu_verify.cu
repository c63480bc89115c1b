// Translation unit for the `verify` kernels (see launchers.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN;
// with neither macro both curves are instantiated (host-simulation build).
#include "launchers.cuh"

namespace bbs {

template <class C> int launch_verify_g1(const VerifyG1Args& a, uint32_t n, rt_stream_t s) {
    return rt_launch<VerifyG1Args, &verify_g1_item<C>, 128>(a, n, s);
}

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_verify_g1<Bls>(const VerifyG1Args&, uint32_t, rt_stream_t);
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_verify_g1<Bn>(const VerifyG1Args&, uint32_t, rt_stream_t);
#endif

}  // namespace bbs
