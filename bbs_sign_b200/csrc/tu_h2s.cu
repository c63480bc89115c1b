// Translation unit for the `h2s` kernels (see launchers.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN;
// with neither macro both curves are instantiated (host-simulation build).
#include "launchers.cuh"

namespace bbs {

template <class C> int launch_h2s(const H2sArgs& a, uint32_t n, rt_stream_t s) {
    return rt_launch<H2sArgs, &h2s_item<C>, 128>(a, n, s);
}

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_h2s<Bls>(const H2sArgs&, uint32_t, rt_stream_t);
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_h2s<Bn>(const H2sArgs&, uint32_t, rt_stream_t);
#endif

}  // namespace bbs
