// Thin runtime layer under the C ABI: device memory, copies and the generic one-thread-per-item launcher.
// CUDA build: cudaMalloc / cudaMemcpyAsync / <<<>>> on sm_100a.
// BBS_HOSTSIM build (tests only, never shipped or loaded by the package): the same entry points backed by
// malloc/memcpy and a host loop, so the pipeline logic can be debugged where no GPU exists.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

namespace bbs {

inline char* rt_errbuf() { static thread_local char buf[512] = ""; return buf; }
inline void rt_set_error(const char* what, const char* detail) { snprintf(rt_errbuf(), 512, "%s: %s", what, detail); }

#ifdef BBS_HOSTSIM

typedef void* rt_stream_t;
inline int rt_set_device(int) { return 0; }
inline int rt_malloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : -2; }
inline void rt_free(void* p) { free(p); }
inline int rt_h2d(void* d, const void* h, size_t n, rt_stream_t) { if (n) memcpy(d, h, n); return 0; }
inline int rt_d2h(void* h, const void* d, size_t n, rt_stream_t) { if (n) memcpy(h, d, n); return 0; }
inline int rt_memset(void* d, int v, size_t n, rt_stream_t) { if (n) memset(d, v, n); return 0; }
inline int rt_sync(rt_stream_t) { return 0; }
inline int rt_stream_create(rt_stream_t* s) { *s = nullptr; return 0; }
inline void rt_stream_destroy(rt_stream_t) {}
inline int rt_set_stack(size_t) { return 0; }

template <class Args, void (*Body)(const Args&, uint32_t), int TPB, int MINB = 1>
inline int rt_launch(const Args& a, uint32_t n, rt_stream_t) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n; i++) Body(a, (uint32_t)i);
    return 0;
}

#else  // CUDA

typedef cudaStream_t rt_stream_t;
#define RT_CHECK(expr)                                                        \
    do {                                                                      \
        cudaError_t e_ = (expr);                                              \
        if (e_ != cudaSuccess) { rt_set_error(#expr, cudaGetErrorString(e_)); return -2; } \
    } while (0)

inline int rt_set_device(int d) { RT_CHECK(cudaSetDevice(d)); return 0; }
inline int rt_malloc(void** p, size_t n) { RT_CHECK(cudaMalloc(p, n ? n : 1)); return 0; }
inline void rt_free(void* p) { if (p) cudaFree(p); }
inline int rt_h2d(void* d, const void* h, size_t n, rt_stream_t s) {
    if (n) RT_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s));
    return 0;
}
inline int rt_d2h(void* h, const void* d, size_t n, rt_stream_t s) {
    if (n) RT_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s));
    return 0;
}
inline int rt_memset(void* d, int v, size_t n, rt_stream_t s) { if (n) RT_CHECK(cudaMemsetAsync(d, v, n, s)); return 0; }
inline int rt_sync(rt_stream_t s) { RT_CHECK(cudaStreamSynchronize(s)); return 0; }
inline int rt_stream_create(rt_stream_t* s) { RT_CHECK(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)); return 0; }
inline void rt_stream_destroy(rt_stream_t s) { if (s) cudaStreamDestroy(s); }
inline int rt_set_stack(size_t bytes) {
    size_t cur = 0;
    RT_CHECK(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
    if (cur < bytes) RT_CHECK(cudaDeviceSetLimit(cudaLimitStackSize, bytes));
    return 0;
}

template <class Args, void (*Body)(const Args&, uint32_t), int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_items(const Args a, uint32_t n) {
    uint32_t i = blockIdx.x * TPB + threadIdx.x;
    if (i < n) Body(a, i);
}

template <class Args, void (*Body)(const Args&, uint32_t), int TPB, int MINB = 1>
inline int rt_launch(const Args& a, uint32_t n, rt_stream_t s) {
    if (n == 0) return 0;
    k_items<Args, Body, TPB, MINB><<<(n + TPB - 1) / TPB, TPB, 0, s>>>(a, n);
    RT_CHECK(cudaGetLastError());
    return 0;
}

#endif

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        rt_free(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        int rc = rt_malloc(&p, want);
        if (rc) return rc;
        cap = want;
        return 0;
    }
    void release() { rt_free(p); p = nullptr; cap = 0; }
};

}  // namespace bbs
