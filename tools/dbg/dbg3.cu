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
    bool ok = fe_sqrt<F>(r, a);
    if (!ok) bn_zero<12>(r);
    fe_from_mont<F>(c, r);
    DUMP(c);
    out[12] = ok;
}
__global__ void k(uint32_t* out, const uint32_t* in) { seq(out, in); }
int main() { return 0; }
