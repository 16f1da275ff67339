// Latency microbenchmarks for the wavefront step budget: dependent fp64 ops, shared-memory round trip, block barrier.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, double a, double b, int n, long long* cyc) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double* out, double a, int n, long long* cyc) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = x + a; x = x * a; x = x + a; x = x * a; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ddiv(double* out, double a, double b, int n, long long* cyc) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = x / b + a; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(int n, long long* cyc) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_smem(double* out, int n, long long* cyc) {
    __shared__ volatile double s[1024];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double x = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x += s[(threadIdx.x + (int)x) & 1023]; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// one wavefront-like step: barrier, 4 volatile smem reads, 13 dependent fp64 ops, smem write
__global__ void k_step(double* out, double a, int n, long long* cyc) {
    extern __shared__ volatile double s[];
    const int t = threadIdx.x, N = blockDim.x;
    s[t] = t; s[N + t] = 1; s[2 * N + t] = 2; s[3 * N + t] = 3;
    __syncthreads();
    double prev = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        __syncthreads();
        const int sl = (i & 3) * N;
        double c = s[sl + t], jp = s[((i + 1) & 3) * N + t], ip = s[sl + ((t + 1) % N)], im = s[sl + ((t + N - 1) % N)];
        double x = (ip - 2.0 * c + im) * a; x = fma(x, a, x); x = fma(x, a, x);
        double y = (jp - 2.0 * c + prev) * a; y = fma(y, a, y); y = fma(y, a, y);
        double r = (x + y) * a; r = a - r; r = r * a; r = fma(r, a, r); r = fma(r, a, r);
        prev = c + r;
        s[((i + 2) & 3) * N + t] = prev;
    }
    long long t1 = clock64();
    out[t] = prev; if (t == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8192 * 8); cudaMallocManaged(&cyc, 8);
    int n = 10000;
    k_dfma<<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc); cudaDeviceSynchronize(); printf("dependent DFMA: %.1f cycles/op\n", cyc[0] / (4.0 * n));
    k_dadd<<<1, 32>>>(out, 1.0000001, n, cyc); cudaDeviceSynchronize(); printf("dependent DADD/DMUL: %.1f cycles/op\n", cyc[0] / (4.0 * n));
    k_ddiv<<<1, 32>>>(out, 1.0000001, 3.0, n, cyc); cudaDeviceSynchronize(); printf("dependent DDIV+DADD: %.1f cycles\n", cyc[0] / (1.0 * n));
    k_smem<<<1, 32>>>(out, n, cyc); cudaDeviceSynchronize(); printf("dependent volatile LDS.64 + DADD + index: %.1f cycles\n", cyc[0] / (1.0 * n));
    for (int nt : {32, 128, 256, 512, 1024}) { k_bar<<<1, nt>>>(n, cyc); cudaDeviceSynchronize(); printf("__syncthreads %4d threads: %.1f cycles\n", nt, cyc[0] / (1.0 * n)); }
    for (int nt : {32, 128, 256, 512, 896, 1024}) { k_step<<<1, nt, 4 * nt * 8>>>(out, 0.3, n, cyc); cudaDeviceSynchronize(); printf("model step %4d threads: %.1f cycles/step\n", nt, cyc[0] / (1.0 * n)); }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
    return 0;
}
