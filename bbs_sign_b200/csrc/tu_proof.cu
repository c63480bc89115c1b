// Translation unit for the `proof` kernels (see launchers.cuh).  Built once per curve: -DBBS_TU_BLS / -DBBS_TU_BN;
// with neither macro both curves are instantiated (host-simulation build).
#include "launchers.cuh"

// occupancy knobs (threads per block, min resident blocks per SM -> register cap); -D overridable
#ifndef BBS_PROOF_G1_TPB
#define BBS_PROOF_G1_TPB 128
#endif
#ifndef BBS_PROOF_G1_MINB
#define BBS_PROOF_G1_MINB 4
#endif

namespace bbs {

template <class C> int launch_proof_g1(const ProofG1Args& a, uint32_t n, rt_stream_t s) {
    return rt_launch<ProofG1Args, &proof_g1_item<C>, BBS_PROOF_G1_TPB, BBS_PROOF_G1_MINB>(a, n, s);
}

template <class C> int launch_proof_gen(const ProofGenArgs& a, uint32_t n, rt_stream_t s) {
    return rt_launch<ProofGenArgs, &proof_gen_item<C>, BBS_PROOF_G1_TPB, BBS_PROOF_G1_MINB>(a, n, s);
}

#if defined(BBS_TU_BLS) || !defined(BBS_TU_BN)
template int launch_proof_gen<Bls>(const ProofGenArgs&, uint32_t, rt_stream_t);
template int launch_proof_g1<Bls>(const ProofG1Args&, uint32_t, rt_stream_t);
#endif
#if defined(BBS_TU_BN) || !defined(BBS_TU_BLS)
template int launch_proof_gen<Bn>(const ProofGenArgs&, uint32_t, rt_stream_t);
template int launch_proof_g1<Bn>(const ProofG1Args&, uint32_t, rt_stream_t);
#endif

}  // namespace bbs
