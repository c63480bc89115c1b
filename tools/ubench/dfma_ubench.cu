// Which multiplier is there besides IMAD.WIDE?  Measures, per SM and clock, the rate of (a) IMAD.WIDE.U32 alone, (b) DFMA
// alone, (c) both in the same warps, interleaved 1:1 (do the two pipes overlap?), (d) DFMA with two IADD3 per DFMA (the
// integer column sums a floating-point limb product needs).  Not part of the library: evidence for DESIGN.md section 6.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_ubench dfma_ubench.cu && ./dfma_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, double fa, int iters) {
    unsigned long long y0 = a, y1 = a + 1, y2 = a + 2, y3 = a + 3;
    double d0 = fa, d1 = fa + 1, d2 = fa + 2, d3 = fa + 3, m = fa * 1.0000001 + threadIdx.x;
    uint32_t x = threadIdx.x * 0x9e3779b9u + a, b0 = a ^ 5, b1 = a + 9, b2 = a * 3, b3 = a + 77;
    uint32_t s0 = a, s1 = a + 3, s2 = a + 5, s3 = a + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0 || MODE == 2)
                asm volatile("mad.wide.u32 %0, %4, %5, %0;\n\tmad.wide.u32 %1, %4, %6, %1;\n\tmad.wide.u32 %2, %4, %7, %2;\n\t"
                             "mad.wide.u32 %3, %4, %8, %3;\n\tadd.u32 %4, %4, 0x632be5ab;"
                             : "+l"(y0), "+l"(y1), "+l"(y2), "+l"(y3), "+r"(x) : "r"(b0), "r"(b1), "r"(b2), "r"(b3));
            if (MODE == 1 || MODE == 2 || MODE == 3)
                asm volatile("fma.rn.f64 %0, %0, %4, %5;\n\tfma.rn.f64 %1, %1, %4, %5;\n\tfma.rn.f64 %2, %2, %4, %5;\n\t"
                             "fma.rn.f64 %3, %3, %4, %5;"
                             : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(m), "d"(fa));
            if (MODE == 3)
                asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;\n\t"
                             "add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.u32 %3, %3, %4;"
                             : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3) : "r"(b0), "r"(b1), "r"(b2), "r"(b3));
        }
    }
    unsigned long long r = y0 ^ y1 ^ y2 ^ y3 ^ __double_as_longlong(d0 + d1 + d2 + d3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32) ^ x ^ s0 ^ s1 ^ s2 ^ s3;
}

template <int MODE> double run(int sms, int iters, uint32_t* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<sms * 8, 256>>>(d, 0x9e3779b9u + rep, 1.000000001 + rep, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
    const int iters = 2000;
    const double groups = (double)sms * 8 * 256 * iters * 16.0;       // groups of 4 operations per thread
    const double clk = khz * 1e3;
    const char* name[4] = {"IMAD.WIDE alone", "DFMA alone", "IMAD.WIDE + DFMA 1:1", "DFMA + 2 IADD3 per DFMA"};
    double ms[4] = {run<0>(sms, iters, d), run<1>(sms, iters, d), run<2>(sms, iters, d), run<3>(sms, iters, d)};
    for (int mo = 0; mo < 4; mo++) {
        const double per_clk_sm = groups * 4 / (ms[mo] * 1e-3) / clk / sms;
        printf("%-28s %8.3f ms  %6.2f ops of each kind / clk / SM (at the nominal %d MHz)\n", name[mo], ms[mo], per_clk_sm, khz / 1000);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
