// FP64 issue throughput per SM: W warps x 8 independent DFMA chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int nt : {32, 64, 128, 256, 416, 512, 1024}) {
        k<<<1, nt>>>(out, cyc, iters); cudaDeviceSynchronize();
        k<<<1, nt>>>(out, cyc, iters); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double winstr = (double)(nt / 32) * 8 * iters;
        printf("threads %4d: %lld cycles, %.2f cycles per warp-DFMA per SM, %.1f DFMA lanes/clk/SM\n", nt, c, c / winstr, 32.0 * winstr / c);
    }
    return 0;
}
