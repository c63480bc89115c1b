#include <cstdio>
#include <cstring>
#include "../../bbs_sign_b200/csrc/field.cuh"
using namespace bbs;
using F = BlsFp;
// trace: run the same sequence on host and device, dump every intermediate
#define DUMP(x) do { for (int i_ = 0; i_ < 12; i_++) out[n * 12 + i_] = (x)[i_]; n++; } while (0)
__host__ __device__ void seq(uint32_t* out) {
    int n = 0;
    uint32_t a[12], s[12], q[12], c[12];
    bn_zero<12>(a); a[0] = 4;
    DUMP(a);
    fe_to_mont<F>(a, a); DUMP(a);
    fe_mul<F>(s, a, a); DUMP(s);
    fe_sqr<F>(q, s); DUMP(q);
    fe_from_mont<F>(c, q); DUMP(c);
    uint32_t tab[4][12];
    fe_set_one<F>(tab[0]); bn_copy<12>(tab[1], a);
    for (int i = 2; i < 4; i++) fe_mul<F>(tab[i], tab[i - 1], a);
    DUMP(tab[2]); DUMP(tab[3]);
    fe_pow<F>(s, a, F::EXP_SQRT(), F::BITS - 1); DUMP(s);
    fe_pow<F>(s, a, F::EXP_INV(), F::BITS); DUMP(s);
    uint32_t e3[12]; bn_zero<12>(e3); e3[0] = 3;
    fe_pow<F>(s, a, e3, 2); DUMP(s);
    fe_pow<F>(s, a, e3, 4); DUMP(s);
    e3[0] = 0x35;
    fe_pow<F>(s, a, e3, 8); DUMP(s);
}
__global__ void k(uint32_t* out) { seq(out); }
int main() {
    uint32_t* d; cudaMalloc(&d, 4 * 12 * 16); cudaMemset(d, 0, 4 * 12 * 16);
    k<<<1, 1>>>(d);
    uint32_t h[12 * 16], g[12 * 16]; memset(g, 0, sizeof g);
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    seq(g);
    for (int b = 0; b < 12; b++) {
        printf("%2d %s dev:", b, memcmp(h + b * 12, g + b * 12, 48) ? "DIFF" : "same");
        for (int i = 11; i >= 0; i--) printf("%08x", h[b * 12 + i]);
        printf("\n        host:"); for (int i = 11; i >= 0; i--) printf("%08x", g[b * 12 + i]); printf("\n");
    }
}
