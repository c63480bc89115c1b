// Micro-benchmark of the cooperative kernel's inner primitives (tuning aid, not part of the library):
// cycles per EP (12x12-limb product into two signed accumulators) and per REDC pair as a function of the
// number of resident warps per SM.    nvcc -arch=sm_100a -O3 -o ep_ubench ep_ubench.cu && ./ep_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../bbs_sign_b200/csrc/gen_coop.cuh"

template <int MODE>
__global__ void k(uint32_t* out, const uint32_t* in, int iters, long long* cyc) {
    uint32_t x[12], y[12], w[24], v[24], R[25], I[25];
    for (int i = 0; i < 12; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 12 + i]; }
    for (int i = 0; i < 25; i++) { R[i] = in[i + 30]; I[i] = in[i + 60]; }
    for (int i = 0; i < 24; i++) { w[i] = in[threadIdx.x + i + 100]; v[i] = in[threadIdx.x + i + 200]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // products only
            coop_wmul_e12(w, x, y); coop_wmul_o12(v, x, y);
            for (int i = 0; i < 12; i++) { x[i] ^= w[i + 3]; y[i] ^= v[i + 5]; }
        } else if (MODE == 1) {     // EP: products + both accumulators (static signs)
            coop_wmul_e12(w, x, y); coop_wmul_o12(v, x, y);
            coop_acc_add_e12(R, w); coop_acc_sub_e12(I, w);
            coop_acc_add_o12(R, v); coop_acc_sub_o12(I, v);
            x[0] ^= R[3]; y[0] ^= I[5];
        } else if (MODE == 2) {     // REDC pair
            uint32_t r0[12], r1[12];
            coop_redc_bls(r0, R); coop_redc_bls(r1, I);
            for (int i = 0; i < 12; i++) { R[i] = r0[i]; R[12 + i] = r1[i] ^ x[i]; I[i] = r1[i]; I[12 + i] = r0[i] ^ y[i]; }
            R[24] = 0; I[24] = 0; R[23] &= 0xffffff; I[23] &= 0xffffff;
        } else if (MODE == 3) {     // accumulate chains only
            coop_acc_add_e12(R, w); coop_acc_sub_e12(I, w);
            coop_acc_add_o12(R, v); coop_acc_sub_o12(I, v);
            w[0] ^= R[7]; v[0] ^= I[9];
        } else if (MODE == 4) {     // products + 96 independent LOP3
            coop_wmul_e12(w, x, y); coop_wmul_o12(v, x, y);
            for (int i = 0; i < 24; i++) { R[i] = (R[i] & w[i]) ^ v[i]; I[i] = (I[i] | v[i]) ^ w[i]; }
            for (int i = 0; i < 24; i++) { R[i] = (R[i] & I[i]) ^ v[i]; I[i] = (I[i] | R[i]) ^ w[i]; }
            x[0] ^= R[3]; y[0] ^= I[5];
        } else if (MODE == 5) {     // products + 96 independent IADD3 (no carries)
            coop_wmul_e12(w, x, y); coop_wmul_o12(v, x, y);
            for (int i = 0; i < 24; i++) { R[i] = R[i] + w[i] + v[i]; I[i] = I[i] + v[i] + w[i]; }
            for (int i = 0; i < 24; i++) { R[i] = R[i] + I[i] + v[i]; I[i] = I[i] + R[i] + w[i]; }
            x[0] ^= R[3]; y[0] ^= I[5];
        }
    }
    long long t1 = clock64();
    uint32_t o = 0;
    for (int i = 0; i < 25; i++) o ^= R[i] ^ I[i];
    for (int i = 0; i < 12; i++) o ^= x[i] ^ y[i];
    if (MODE >= 3) for (int i = 0; i < 24; i++) o ^= w[i] ^ v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, uint32_t* d_out, uint32_t* d_in, long long* d_cyc) {
    int sms = 148;
    for (int wps : {4, 8, 12}) {      // warps per SM (one block per SM)
        int iters = 2000;
        k<MODE><<<sms, wps * 32>>>(d_out, d_in, iters, d_cyc);
        cudaDeviceSynchronize();
        k<MODE><<<sms, wps * 32>>>(d_out, d_in, iters, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[148];
        cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; i++) avg += c[i];
        avg /= sms;
        // per SMSP: warps wps/4, each doing `iters` iterations in avg cycles
        printf("%-28s warps/SM %2d: %7.1f cycles/iter/warp, %7.1f cycles/iter per SMSP  (%s)\n", name, wps, avg / iters,
               avg / iters / (wps / 4.0), cudaGetErrorString(e));
    }
}

int main() {
    uint32_t *d_out, *d_in; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_in, 4096 * 4); cudaMalloc(&d_cyc, 148 * 8);
    uint32_t h[4096];
    for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1);
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("products (144 IMAD.WIDE)", d_out, d_in, d_cyc);
    run<1>("EP (products + 4 acc chains)", d_out, d_in, d_cyc);
    run<2>("REDC pair (288 IMAD.WIDE)", d_out, d_in, d_cyc);
    run<3>("4 acc chains only (98 IADD3)", d_out, d_in, d_cyc);
    run<4>("products + 96 LOP3", d_out, d_in, d_cyc);
    run<5>("products + 96 IADD3", d_out, d_in, d_cyc);
    return 0;
}
