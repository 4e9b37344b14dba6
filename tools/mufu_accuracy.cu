// Measures the relative accuracy of the MUFU seeds (rcp/rsqrt.approx.ftz.f64) and of the Newton
// sequences built on them in ort_internal.cuh, against correctly rounded IEEE results.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../include -I../opticalraytracing.jl_b200/csrc mufu_accuracy.cu
#include <cstdio>
#include <cmath>
#include "ort_internal.cuh"

__device__ double u64_to_unit(unsigned long long h) { return (h >> 11) * (1.0 / 9007199254740992.0); }
__device__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
// candidate shorter sequences
__device__ __forceinline__ double div3(double n, double d) { double r = mufu_rcp(d); double e = fma(-d, r, 1.0); r = fma(r, fma(e, e, e), r); return n * r; }
__device__ __forceinline__ double div4(double n, double d) { double r = mufu_rcp(d); double e = fma(-d, r, 1.0); r = fma(r, fma(e, e, e), r); double q = n * r; return fma(fma(-d, q, n), r, q); }
__device__ __forceinline__ double sqrt3(double a) { double r = mufu_rsqrt(a); double g = a * r; double h = half_of(r); double e = fma(-h, g, 0.5); return fma(g, e, g); }
__device__ __forceinline__ double sqrt4(double a) { double r = mufu_rsqrt(a); double g = a * r; double h = half_of(r); double e = fma(-h, g, 0.5); return fma(g, fma(1.5 * e, e, e), g); }

__global__ void k(double* out, int n_per_thread)
{
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long s = mix(blockIdx.x * 1024ULL + threadIdx.x + 1);
    for (int i = 0; i < n_per_thread; i++) {
        s = mix(s + i);
        double a = exp2(40.0 * (u64_to_unit(s) - 0.5)) * (1.0 + u64_to_unit(mix(s ^ 0x9e3779b97f4a7c15ULL)));
        s = mix(s);
        double b = exp2(20.0 * (u64_to_unit(s) - 0.5)) * (1.0 + u64_to_unit(mix(s ^ 0x1234567ULL)));
        double rs = 1.0 / sqrt(a), rc = 1.0 / a, sq = sqrt(a), dv = b / a;
        m[0] = fmax(m[0], fabs(mufu_rsqrt(a) - rs) / rs);
        m[1] = fmax(m[1], fabs(mufu_rcp(a) - rc) / rc);
        m[2] = fmax(m[2], fabs(fast_sqrt(a) - sq) / sq);
        m[3] = fmax(m[3], fabs(fast_div(b, a) - dv) / fabs(dv));
        m[4] = fmax(m[4], fabs(fast_rsqrt(a) - rs) / rs);
        m[5] = fmax(m[5], fabs(div3(b, a) - dv) / fabs(dv));
        m[6] = fmax(m[6], fabs(sqrt3(a) - sq) / sq);
        m[7] = fmax(m[7], fmax(fabs(div4(b, a) - dv) / fabs(dv), 0.0) + 0.0 * fabs(sqrt4(a) - sq));
    }
    for (int j = 0; j < 8; j++) {
        double v = m[j];
        for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(~0u, v, d));
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[j], (unsigned long long)__double_as_longlong(v));
    }
    // sqrt4 separately
    double m8 = 0; s = mix(blockIdx.x * 7777ULL + threadIdx.x + 5);
    for (int i = 0; i < n_per_thread; i++) { s = mix(s + i); double a = exp2(40.0 * (u64_to_unit(s) - 0.5)) * (1.0 + u64_to_unit(mix(s ^ 99ULL))); double sq = sqrt(a); m8 = fmax(m8, fabs(sqrt4(a) - sq) / sq); }
    for (int d = 16; d; d >>= 1) m8 = fmax(m8, __shfl_xor_sync(~0u, m8, d));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[8], (unsigned long long)__double_as_longlong(m8));
}

int main()
{
    double* d; cudaMalloc(&d, 9 * 8); cudaMemset(d, 0, 9 * 8);
    k<<<592, 256>>>(d, 4000);
    double h[9]; cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost);
    const char* names[9] = {"MUFU.RSQ64H seed", "MUFU.RCP64H seed", "fast_sqrt (5 op)", "fast_div (5 op)", "fast_rsqrt", "div3 (cubic, 4 op)", "sqrt3 (1 Newton, 3 op)", "div4 (cubic+resid, 6 op)", "sqrt4 (cubic, 5 op)"};
    for (int j = 0; j < 9; j++) printf("%-28s max rel err = %.3e = 2^%.2f  (%.2f ulp)\n", names[j], h[j], log2(h[j]), h[j] / 1.1102230246251565e-16);
    return 0;
}
