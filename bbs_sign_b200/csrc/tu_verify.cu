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
// BN254: 64-thread blocks (8 per SM) measured 9 % faster than 256-thread blocks; BLS12-381 the other way round (6 %)
#ifndef BBS_VERIFY_G1_TPB_BN
#define BBS_VERIFY_G1_TPB_BN 64
#endif
#ifndef BBS_VERIFY_G1_MINB_BN
#define BBS_VERIFY_G1_MINB_BN 8
#endif

namespace bbs {

#ifndef BBS_VERIFY_SPLIT_TPB
#define BBS_VERIFY_SPLIT_TPB 128
#endif
#ifndef BBS_VERIFY_SPLIT_MINB
#define BBS_VERIFY_SPLIT_MINB 4
#endif
template <class C> struct VerifyG1Geom {
    static constexpr int TPB = BBS_VERIFY_G1_TPB, MINB = BBS_VERIFY_G1_MINB;
    static constexpr int SPLIT_TPB = BBS_VERIFY_SPLIT_TPB, SPLIT_MINB = BBS_VERIFY_SPLIT_MINB;
};
template <> struct VerifyG1Geom<Bn> {
    static constexpr int TPB = BBS_VERIFY_G1_TPB_BN, MINB = BBS_VERIFY_G1_MINB_BN;
    static constexpr int SPLIT_TPB = BBS_VERIFY_G1_TPB_BN, SPLIT_MINB = BBS_VERIFY_G1_MINB_BN;
};

template <class C> int launch_verify_g1(const VerifyG1Args& a, uint32_t n, rt_stream_t s) {
#ifdef BBS_HOSTSIM
    if (a.part_v) {                                     // the task functions of the split path, run in sequence
        uint8_t* fbad = a.part_st + n;
        for (uint32_t i = 0; i < n; i++) { verify_task_v<C>(a, i); verify_task_f<C>(a, i, fbad + i); }
        for (uint32_t i = 0; i < n; i++) verify_join_item<C>(a, i, fbad);
        return 0;
    }
    return rt_launch<VerifyG1Args, &verify_g1_item<C>, BBS_VERIFY_G1_TPB, BBS_VERIFY_G1_MINB>(a, n, s);
#else
    if (n == 0) return 0;
    constexpr int TPB = VerifyG1Geom<C>::TPB, MINB = VerifyG1Geom<C>::MINB;
    if (a.part_v) {
        // two tasks per item (kernels.cuh verify_task_v / verify_task_f) and the join; the state bytes of F live behind those of V
        constexpr int ST = VerifyG1Geom<C>::SPLIT_TPB, SM = VerifyG1Geom<C>::SPLIT_MINB;
        const uint32_t nb = (n + ST - 1) / ST;
        uint8_t* fbad = a.part_st + n;
        verify_g1_split_kernel<C, ST, SM><<<2 * nb, ST, 0, s>>>(a, n, nb, fbad);
        RT_CHECK(cudaGetLastError());
        verify_g1_combine_kernel<C, 128><<<(n + 127) / 128, 128, 0, s>>>(a, n, fbad);
        RT_CHECK(cudaGetLastError());
        return 0;
    }
    verify_g1_kernel<C, TPB, MINB><<<(n + TPB - 1) / TPB, TPB, 0, s>>>(a, n);
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
