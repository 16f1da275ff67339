// common.cuh -- shared device/host definitions for libsrcfd (sm_100a).
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, every expression is
// written in the evaluation order of the reference's Python source and true IEEE fp64 divisions
// are kept, so a cell update here produces the same bits as the reference's numba kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srcfd {

enum { OP_PRESSURE = 0, OP_UPWIND = 1, OP_QUICK = 2 };

// Grid geometry + physical constants handed to every kernel by value.
struct Consts {
    int nx, ny;
    int pitch;            // ny + 2 (reference layout, unpadded)
    long long plane;      // (nx+2)*(ny+2)
    double dx, dy, volp, dt, nu, rho;
    // derived with the reference's own expressions (see make_consts in api.cu)
    double dx2, dy2;      // dx*dx, dy*dy
    double ap_d;          // -volp*(2/(dx*dx) + 2/(dy*dy))            LDC.py:236
    double volp_dt;       // volp/dt                                     LDC.py:260-261
    double rho_dt;        // rho/dt                                      LDC.py:305
    double neg_nu;        // -nu
    double neg_nu_ap_d;   // (-nu)*ap_d
    double dt_rho;        // dt/rho                                      LDC.py:321
    double mdt_rho;       // -dt/rho                                     LDC.py:243
    double two_dx, two_dy;// 2*dx, 2*dy
};

struct BcSpec {
    int types[3][4];
    double values[3][4];
    int bfs;              // left-boundary override of BFS.py:524-562
    double step_h, h, Ub;
    int skip_lo, skip_hi; // slab decomposition: row 0 / row nx+1 of the local plane belongs to a neighbour, not to the boundary
};

// Device-resident control block.  Written only by single-thread epilogue kernels or by thread 0 of
// block 0 of a solve kernel; read by every kernel at entry (stop => early exit).
struct Ctrl {
    int stop;                 // converged, NaN/Inf or deadlock guard: all later kernels are no-ops
    int converged;
    int nan_flag;
    int deadlock;
    long long iterations;
    double residual[3];       // CFDSolver.residual
    double rms[3];            // sqrt(residual/(nx*ny))/dt
    int last_sweeps[3];
    int guess[3];             // sweep-count guess per inner solve (u, v, p) for the pipelined GS order
    long long total_sweeps[3];
    double last_inner_rms[3];
    long long n_hist;
    long long hist_cap;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Fixed-order block sum: lanes by xor-shuffle, then warps in index order by thread 0.
// Deterministic for a given block size.  `scratch` needs >= 32 doubles.  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0 && threadIdx.y == 0)
        for (int w = 0; w < nwarp; ++w) s += scratch[w];
    return s;
}

}  // namespace srcfd
