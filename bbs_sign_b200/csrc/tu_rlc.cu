// Translation unit for the random-linear-combination batch mode (rlc.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN.
#include "launchers.cuh"

#ifndef BBS_HOSTSIM
namespace bbs {

template <class C> int launch_rlc_partial(const RlcArgs& a, uint32_t n_blocks, rt_stream_t s) {
    if (n_blocks == 0) return 0;
    rlc_partial_kernel<C><<<n_blocks, RLC_TPB, 0, s>>>(a);
    RT_CHECK(cudaGetLastError());
    return 0;
}
template <class C> int launch_rlc_finish(const RlcFinishArgs& a, rt_stream_t s) {
    rlc_finish_kernel<C><<<1, RLC_TPB, 0, s>>>(a);
    RT_CHECK(cudaGetLastError());
    return 0;
}
template <class C> int launch_rlc_combine(const RlcCombineArgs& a, rt_stream_t s) {
    return rt_launch<RlcCombineArgs, &rlc_combine_item<C>, 32>(a, 1, s);
}

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_rlc_partial<Bls>(const RlcArgs&, uint32_t, rt_stream_t);
template int launch_rlc_finish<Bls>(const RlcFinishArgs&, rt_stream_t);
template int launch_rlc_combine<Bls>(const RlcCombineArgs&, rt_stream_t);
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_rlc_partial<Bn>(const RlcArgs&, uint32_t, rt_stream_t);
template int launch_rlc_finish<Bn>(const RlcFinishArgs&, rt_stream_t);
template int launch_rlc_combine<Bn>(const RlcCombineArgs&, rt_stream_t);
#endif

}  // namespace bbs
#endif
