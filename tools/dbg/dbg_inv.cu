#include <cstdio>
#include "../../bbs_sign_b200/csrc/launchers.cuh"
using namespace bbs;
__global__ void k(uint32_t* out) {
    using F = BlsFp;
    __shared__ uint32_t tree[2 * 256][12];
    uint32_t z[12], own[12], chk[12];
    // z = (t+2) in Montgomery form
    uint32_t c[12]; for (int i = 0; i < 12; i++) c[i] = 0; c[0] = threadIdx.x + 2;
    fe_to_mont<F>(z, c);
    if (threadIdx.x & 1) fe_set_one<F>(z);
    fe_inv<F>(own, z);
    block_batch_inverse<F, 256>(z, tree);
    int bad = 0;
    for (int i = 0; i < 12; i++) bad |= (own[i] != z[i]);
    out[threadIdx.x] = bad;
}
int main() {
    cudaDeviceSetLimit(cudaLimitStackSize, 48 * 1024);
    uint32_t* d; cudaMalloc(&d, 1024);
    k<<<1, 256>>>(d);
    uint32_t h[256]; cudaMemcpy(h, d, 1024, cudaMemcpyDeviceToHost);
    int nb = 0; for (int i = 0; i < 256; i++) nb += h[i];
    printf("err %s, mismatching threads: %d (first: ", cudaGetErrorString(cudaGetLastError()), nb);
    for (int i = 0; i < 256 && nb; i++) if (h[i]) { printf("%d ", i); if (--nb < -8) break; }
    printf(")\n");
}
