// Flag hand-off latency between two CTAs (different SMs): release store -> relaxed poll (+fence) -> reply.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int ld_relaxed(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
template <int MODE>   // 0: release + relaxed poll; 1: + __threadfence both sides; 2: mode 1 + nanosleep(20) in the poll loop
__global__ void pingpong(int* flags, int n, long long* ns, double* junk) {
    if (threadIdx.x != 0) {   // background stores by the other threads, like the compute warps
        if (MODE >= 1) for (int i = 0; i < n * 4; ++i) junk[(blockIdx.x * blockDim.x + threadIdx.x) * 16 + (i & 15)] = i;
        return;
    }
    int* mine = flags + blockIdx.x * 32; int* other = flags + (1 - blockIdx.x) * 32;
    long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 1; i <= n; ++i) {
        if (blockIdx.x == 0) {
            if (MODE >= 1) __threadfence();
            st_release(mine, i);
            while (ld_relaxed(other) < i) { if (MODE == 2) __nanosleep(20); }
            if (MODE >= 1) __threadfence();
        } else {
            while (ld_relaxed(other) < i) { if (MODE == 2) __nanosleep(20); }
            if (MODE >= 1) __threadfence();
            st_release(mine, i);
        }
    }
    long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (blockIdx.x == 0) ns[0] = t1 - t0;
}
int main() {
    int* flags; long long* ns; double* junk;
    cudaMalloc(&flags, 1024); cudaMallocManaged(&ns, 8); cudaMalloc(&junk, 2 * 1024 * 16 * 8);
    int n = 2000;
    for (int threads : {32, 704}) {
        cudaMemset(flags, 0, 1024); pingpong<0><<<2, threads>>>(flags, n, ns, junk); cudaDeviceSynchronize();
        printf("threads %d mode0 (release/relaxed):           round trip %.0f ns\n", threads, (double)ns[0] / n);
        cudaMemset(flags, 0, 1024); pingpong<1><<<2, threads>>>(flags, n, ns, junk); cudaDeviceSynchronize();
        printf("threads %d mode1 (+threadfence, bg stores):   round trip %.0f ns\n", threads, (double)ns[0] / n);
        cudaMemset(flags, 0, 1024); pingpong<2><<<2, threads>>>(flags, n, ns, junk); cudaDeviceSynchronize();
        printf("threads %d mode2 (+nanosleep 20):             round trip %.0f ns\n", threads, (double)ns[0] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
