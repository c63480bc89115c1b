#include <cstdio>
#include <cstring>
#include "../../bbs_sign_b200/csrc/field.cuh"
using namespace bbs;
using F = BlsFp;
#define DUMP(x) do { for (int i_ = 0; i_ < 12; i_++) out[n * 12 + i_] = (x)[i_]; n++; } while (0)
__host__ __device__ void seq(uint32_t* out, const uint32_t* in) {
    int n = 0;
    uint32_t a[12], r[12], c[12];
    for (int i = 0; i < 12; i++) a[i] = in[i];
    fe_to_mont<F>(a, a);
    DUMP(a);
    uint32_t s[F::N], q[F::N];
    fe_pow<F>(s, a, F::EXP_SQRT(), F::BITS - 1);
    DUMP(s);
    fe_sqr<F>(q, s);
    DUMP(q);
    bool ok = bn_eq<F::N>(q, a);
    out[15 * 12] = ok;
    uint32_t o = 0;
    for (int i = 0; i < 12; i++) o |= q[i] ^ a[i];
    out[15 * 12 + 1] = o;
    bool ok2 = fe_sqrt<F>(r, a);
    out[15 * 12 + 2] = ok2;
    DUMP(r);
}
__global__ void k(uint32_t* out, const uint32_t* in) { seq(out, in); }
int main() {
    uint32_t *d, *din; cudaMalloc(&d, 4 * 12 * 16); cudaMemset(d, 0, 4 * 12 * 16); cudaMalloc(&din, 48);
    uint32_t in[12] = {4}; cudaMemcpy(din, in, 48, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(d, din);
    uint32_t h[12 * 16], g[12 * 16]; memset(g, 0, sizeof g);
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    seq(g, in);
    for (int b = 0; b < 4; b++) {
        printf("%2d %s dev:", b, memcmp(h + b * 12, g + b * 12, 48) ? "DIFF" : "same");
        for (int i = 11; i >= 0; i--) printf("%08x", h[b * 12 + i]);
        printf("\n        host:"); for (int i = 11; i >= 0; i--) printf("%08x", g[b * 12 + i]); printf("\n");
    }
    printf("flags dev %u %08x %u host %u %08x %u\n", h[180], h[181], h[182], g[180], g[181], g[182]);
}
