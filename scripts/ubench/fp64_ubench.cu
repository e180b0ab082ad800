// B200 fp64 pipe microbenchmark: dependent-issue latency and throughput of DFMA / MUFU.RSQ64H,
// as a function of warps per SM sub-partition and independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int u = 0; u < ILP; u++) x[u] = threadIdx.x * 1e-3 + u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < ILP; u++) x[u] = fma(x[u], a, b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; u++) s += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(double* out, int iters) {
    double x = threadIdx.x + 1.5, s = 0;
    for (int it = 0; it < iters; it++) {
        double y;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        s += y; x += 1.0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024 * 8);
    int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    printf("clock attr %d kHz\n", dev_clk);
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
        int blocks = 148 * (warps_per_sm * 32 / threads);
        float t1 = timeit([&] { k_dfma<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t2 = timeit([&] { k_dfma<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t4 = timeit([&] { k_dfma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t8 = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float tm = timeit([&] { k_mufu<<<blocks, threads>>>(out, iters); });
        double n = (double)blocks * threads * iters;
        printf("warps/SM %2d: DFMA Tflop/s ilp1 %.2f ilp2 %.2f ilp4 %.2f ilp8 %.2f | ns per dependent DFMA (ilp1) %.2f | MUFU.RSQ64H Gop/s %.1f\n",
               warps_per_sm, 2 * n / t1 * 1e-9, 2 * 2 * n / t2 * 1e-9, 4 * 2 * n / t4 * 1e-9, 8 * 2 * n / t8 * 1e-9,
               t1 * 1e6 / iters, n / tm * 1e-6);
    }
    return 0;
}
