// Accuracy of the fp64 MUFU seeds (rcp / rsqrt .approx.ftz.f64) on B200 and of the refinement schemes built on
// them: two quadratic Newton steps (round 1) against ONE cubic step. Prints max relative errors over 2^24
// arguments spread over [2^-40, 2^40).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o seed_accuracy seed_accuracy.cu && ./seed_accuracy
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double amax(double a, double b) { return a > b ? a : b; }
__global__ void k(double* out) {
    double e_seed_r = 0, e_seed_q = 0, e_r2 = 0, e_r3 = 0, e_q2 = 0, e_q3 = 0;
    const unsigned n = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned it = 0; it < 64; it++) {
        const unsigned u = (t + n * it) * 2654435761u;
        const double m = 1.0 + (double)(u & 0xffffff) / 16777216.0 + (double)(u >> 24) / 4294967296.0 / 16.0;
        const double x = ldexp(m, (int)((u >> 8) % 81) - 40);
        double r, q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
        const double rt = 1.0 / x, qt = 1.0 / sqrt(x);
        e_seed_r = amax(e_seed_r, fabs(r - rt) / rt);
        e_seed_q = amax(e_seed_q, fabs(q - qt) / qt);
        {   // two Newton steps
            double a = r, e = fma(-x, a, 1.0); a = fma(a, e, a); e = fma(-x, a, 1.0); a = fma(a, e, a);
            e_r2 = amax(e_r2, fabs(a - rt) / rt);
            double y = q; const double hx = 0.5 * x;
            double f = fma(-hx * y, y, 0.5); y = fma(y, f, y); f = fma(-hx * y, y, 0.5); y = fma(y, f, y);
            e_q2 = amax(e_q2, fabs(y - qt) / qt);
        }
        {   // one cubic step
            const double e = fma(-x, r, 1.0);
            const double a = fma(r, fma(e, e, e), r);
            e_r3 = amax(e_r3, fabs(a - rt) / rt);
            const double f = fma(-x * q, q, 1.0);
            const double y = fma(q * f, fma(0.375, f, 0.5), q);
            e_q3 = amax(e_q3, fabs(y - qt) / qt);
        }
    }
    double v[6] = {e_seed_r, e_seed_q, e_r2, e_r3, e_q2, e_q3};
    for (int k = 0; k < 6; k++) {
        double s = v[k];
        for (int o = 16; o > 0; o >>= 1) s = amax(s, __shfl_xor_sync(0xffffffffu, s, o));
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[k], (unsigned long long)__double_as_longlong(s));
    }
}
int main() {
    double* d; cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
    k<<<1024, 256>>>(d);
    double h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("max relative error: rcp seed %.3e (2^%.1f)  rsqrt seed %.3e (2^%.1f)\n", h[0], log2(h[0]), h[1], log2(h[1]));
    printf("rcp:   two Newton steps %.3e   one cubic step %.3e\n", h[2], h[3]);
    printf("rsqrt: two Newton steps %.3e   one cubic step %.3e\n", h[4], h[5]);
    return 0;
}
