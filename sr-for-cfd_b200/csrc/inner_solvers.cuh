// inner_solvers.cuh -- the three inner relaxation loops (solve_pressure, solve_momentum_upwind,
// solve_momentum_quick; LDC.py:248-314) as persistent cooperative kernels.  One launch runs the whole
// "repeat <= max_iter sweeps, stop after the first sweep with rms < tol" loop on the device.
//
// Two families:
//   k_solve_sync<OP,ORDER>  Jacobi / red-black: one grid-wide barrier per (half) sweep; the break test is
//                           evaluated by every block from the same fixed-order residual sum.
//   k_solve_gs<OP>          reference order (in-place lexicographic Gauss-Seidel), executed as a
//                           pipelined wavefront; see the comment block above it.
#pragma once
#include <cooperative_groups.h>
#include "cell_ops.cuh"

namespace srcfd {
namespace cg = cooperative_groups;

struct SolveArgs {
    double* Var;              // (3, nx+2, ny+2) + tail padding
    const double* VarOld;
    const double* Ff;         // (4, nx+2, ny+2)
    const double* rhs;        // (nx+2, ny+2) pressure right-hand side
    double* scratch;          // one plane: Jacobi second buffer / Gauss-Seidel rollback snapshot
    double* partials;         // residual partial sums
    int* prog;                // wavefront progress flags [sweep][band]
    Ctrl* ctrl;
    Consts K;
    int k;                    // plane relaxed (0 u, 1 v, 2 p)
    int slot;                 // which Ctrl counters to update (0 u, 1 v, 2 p)
    double tol;
    int max_iter;
    int nbands, band_rows;    // wavefront row bands
    int spin_limit;
    int guess_bias;           // first group runs guess + bias sweeps (bias <= 0 under-guesses)
    double omega;             // RED_BLACK pressure: successive over-relaxation factor (1 = plain sweep)
};

// ---------------------------------------------------------------------------------------------------
// Stencil evaluation from memory (Jacobi / red-black).  S = base of the plane being READ for plane k.
// Second neighbours that fall outside the (nx+2, ny+2) plane follow the reference's flat-buffer
// behaviour (SURVEY.md hazard H4); those locations are ghost cells, so they are always read from Var.
// ---------------------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ double eval_cell(const SolveArgs& a, const double* __restrict__ S, int i, int j, double& R) {
    const Consts& K = a.K;
    const long long c = (long long)i * K.pitch + j;
    const double vc = __ldcg(S + c), vip = __ldcg(S + c + K.pitch), vim = __ldcg(S + c - K.pitch);
    const double vjp = __ldcg(S + c + 1), vjm = __ldcg(S + c - 1);
    if (OP == OP_PRESSURE) return pressure_cell(vc, vip, vim, vjp, vjm, __ldg(a.rhs + c), K, R);
    const long long kb = (long long)a.k * K.plane;
    const double vold = __ldg(a.VarOld + kb + c);
    const double fE = __ldg(a.Ff + c), fN = __ldg(a.Ff + K.plane + c);
    const double fW = __ldg(a.Ff + 2 * K.plane + c), fS = __ldg(a.Ff + 3 * K.plane + c);
    if (OP == OP_UPWIND) return upwind_cell(vc, vip, vim, vjp, vjm, vold, fE, fN, fW, fS, K, R);
    const double* G = a.Var + kb;   // ghost source
    const double vip2 = (i + 2 <= K.nx + 1) ? __ldcg(S + c + 2 * K.pitch) : __ldcg(G + (long long)(K.nx + 2) * K.pitch + j);
    const double vim2 = (i - 2 >= 0) ? __ldcg(S + c - 2 * K.pitch) : __ldcg(G + (long long)(K.nx + 1) * K.pitch + j);
    const double vjp2 = (j + 2 <= K.ny + 1) ? __ldcg(S + c + 2) : __ldcg(G + (long long)(i + 1) * K.pitch);
    const double vjm2 = (j - 2 >= 0) ? __ldcg(S + c - 2) : __ldcg(G + (long long)i * K.pitch + K.ny + 1);
    return quick_cell(vc, vip, vim, vjp, vjm, vip2, vim2, vjp2, vjm2, vold, fE, fN, fW, fS, K, R);
}

constexpr int SYNC_THREADS = 256;

template <int OP, int ORDER>
__global__ void __launch_bounds__(SYNC_THREADS) k_solve_sync(SolveArgs a) {
    cg::grid_group grid = cg::this_grid();
    if (a.ctrl->stop) return;
    const Consts& K = a.K;
    __shared__ double red[32];
    __shared__ double s_tot;
    double* A = a.Var + (long long)a.k * K.plane;
    double* Bp = a.scratch;
    const long long ncell = (long long)K.nx * K.ny;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    if (ORDER == 1) {  // Jacobi: second buffer needs the ghost cells too
        for (long long t = gtid; t < K.plane; t += gsize) Bp[t] = A[t];
        grid.sync();
    }
    const double* src = A;
    double* dst = (ORDER == 1) ? Bp : A;
    int n = 0;
    double rms = 0.0;
    for (int it = 0; it < a.max_iter; ++it) {
        double acc = 0.0;
        if (ORDER == 1) {
            for (long long idx = gtid; idx < ncell; idx += gsize) {
                const int i = (int)(idx / K.ny) + 1, j = (int)(idx % K.ny) + 1;
                double R;
                const double nv = eval_cell<OP>(a, src, i, j, R);
                dst[(long long)i * K.pitch + j] = nv;
                acc += R * R;
            }
        } else {
            for (int colour = 0; colour < 2; ++colour) {
                for (long long idx = gtid; idx < ncell; idx += gsize) {
                    const int i = (int)(idx / K.ny) + 1, j = (int)(idx % K.ny) + 1;
                    if (((i + j) & 1) != colour) continue;
                    double R;
                    double nv = eval_cell<OP>(a, A, i, j, R);
                    if (OP == OP_PRESSURE && a.omega != 1.0)         // red-black SOR: p += omega * R/ap (R as in the plain sweep)
                        nv = __ldcg(A + (long long)i * K.pitch + j) + a.omega * (R / K.ap_d);
                    A[(long long)i * K.pitch + j] = nv;
                    acc += R * R;
                }
                if (colour == 0) grid.sync();
            }
        }
        double* part = a.partials + (size_t)(it & 1) * gridDim.x;
        const double tot = block_sum(acc, red);
        if (threadIdx.x == 0) part[blockIdx.x] = tot;
        grid.sync();
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(part + b);
        const double all = block_sum(s, red);
        if (threadIdx.x == 0) s_tot = all;
        __syncthreads();
        rms = sqrt(s_tot / (double)ncell);
        n = it + 1;
        if (ORDER == 1) { const double* t = src; src = dst; dst = const_cast<double*>(t); }
        if (rms < a.tol) break;
    }
    if (ORDER == 1 && src != A) {  // latest iterate lives in the scratch plane
        for (long long idx = gtid; idx < ncell; idx += gsize) {
            const long long c = (idx / K.ny + 1) * K.pitch + (idx % K.ny) + 1;
            A[c] = __ldcg(src + c);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctrl->last_sweeps[a.slot] = n;
        a.ctrl->total_sweeps[a.slot] += n;
        a.ctrl->last_inner_rms[a.slot] = rms;
    }
}

// ---------------------------------------------------------------------------------------------------
// Reference-order Gauss-Seidel as a pipelined wavefront.
//
// The reference updates cells in place in lexicographic (i outer, j inner) order, so cell (i,j) of
// sweep s reads (i-1,j),(i,j-1) already updated in sweep s and (i+1,j),(i,j+1) still holding sweep
// s-1.  Any schedule that respects those data dependences produces the same bits.  Here:
//   * a TASK is (sweep s, row band b); one CTA runs one task at a time; thread r owns grid row
//     i0+r and walks it along j (the contiguous axis), one step behind thread r-1, so at step tau it
//     updates column jj = tau - r + 1.  "new" neighbours arrive from thread r-1 / its own previous
//     step, "old" neighbours from thread r+1's look-ahead window / its own window, both through
//     double-buffered shared memory with ONE block barrier per step;
//   * two extra passive threads above and below the band stream the neighbouring rows (already-new
//     rows of band b-1 in this sweep, still-old rows of band b+1 from sweep s-1);
//   * tasks are dealt round-robin to the persistent CTAs in (s,b) order, and every dependence points
//     to an earlier task, so the grid is a feed-forward pipeline: sweep s+1 trails sweep s by a few
//     columns, synchronised by per-task progress counters in global memory (release/acquire);
//   * each thread reads its own row a few columns ahead into a register ring (L1-bypassing loads: the data is
//     produced by other SMs during the kernel).
// With W = R + 1 (R = ring size, see WfShape), before step tau of task (s,b) starts the following must have been published
// (P = completed steps of a task, capped at that task's step count):
//     P(s-1, b)   >= tau + W                     own rows hold sweep s-1
//     P(s-1, b+1) >= tau + W - nrows(b)          rows below hold sweep s-1
//     P(s,   b-1) >= tau + W + nrows(b-1)        rows above hold sweep s
// The same three conditions also order every in-place overwrite after its last reader.
//
// The break test needs the rms of a COMPLETED sweep, but later sweeps are already in flight by then.
// So sweeps run in speculative groups: snapshot the plane, run n sweeps, then look for the first sweep
// with rms < tol; if that is not the last one of the group, restore the snapshot and re-run exactly
// that many sweeps.  The group size is the previous outer iteration's count, so the common case is
// one pass.
// ---------------------------------------------------------------------------------------------------
constexpr int WF_SVC = 64;           // two service warps: publisher (first lane of warp -2) and poller (warp -1)
constexpr int WF_MAX_THREADS = 512;  // band rows + 4 passive threads (warp-rounded) + service warps
constexpr int WF_MAX_BAND = WF_MAX_THREADS - WF_SVC - 4;

// Per-operator pipeline shape.  WIN = columns jj..jj+WIN-1 a thread must hold (itself + what it
// publishes to the rows above); D = extra columns in flight from L2; R = WIN + D is both the register
// ring size and the unroll factor of the step loop (ring slots are addressed by step parity, so no
// register is ever moved while its load is outstanding).  AR = ring for the read-only inputs.
template <int OP> struct WfShape;
template <> struct WfShape<OP_PRESSURE> { static constexpr int WIN = 3, D = 3, R = 6, AR = 3; };
template <> struct WfShape<OP_UPWIND>   { static constexpr int WIN = 3, D = 3, R = 6, AR = 3; };
template <> struct WfShape<OP_QUICK>    { static constexpr int WIN = 4, D = 4, R = 8, AR = 4; };

template <int OP> struct WfAux;
template <> struct WfAux<OP_PRESSURE> { double rhs; };
template <> struct WfAux<OP_UPWIND> { double vold, fE, fN, fW, fS; };
template <> struct WfAux<OP_QUICK> { double vold, fE, fN, fW, fS; };

__device__ __forceinline__ int lds_volatile(const int* p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(int* p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bar_compute(int nthreads) {   // named barrier 1: compute warps only
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

constexpr int WF_INF = 0x3fffffff;

// One task = (sweep s, band b).  Threads [0, NT-WF_SVC) compute; the two service warps run beside them,
// off the per-step critical path:
//   publisher  watches s_sync[0] (= number of completed steps, written by a compute thread right after
//              the step barrier), then fence + release-store of that number to the task's global flag;
//   poller     reads the three dependence flags and keeps s_sync[1] = highest step index that may start.
template <int OP>
__device__ void wf_task(const SolveArgs& a, const int s, const int b, double* smem, int* s_sync, double& acc_out) {
    const Consts& K = a.K;
    constexpr bool Q = (OP == OP_QUICK);
    constexpr int R = WfShape<OP>::R, AR = WfShape<OP>::AR;
    constexpr int W = R + 1;               // dependence look-ahead in steps (see the block comment above)
    const int NT = blockDim.x, tid = threadIdx.x, SP = NT + 4;
    const int NCOMP = NT - WF_SVC;
    // volatile: the per-step barrier is inline PTX (named barrier 1); keep the compiler from caching these across it
    volatile double* s_new = smem + 2;     // [2][SP], index = buf*SP + tid, valid tid range -2 .. NT+1
    volatile double* s_c   = smem + 2 + 2 * SP;
    volatile double* s_im  = smem + 2 + 4 * SP;
    volatile double* s_c2  = smem + 2 + 6 * SP;

    const int B = a.nbands;
    const int i0 = 1 + b * a.band_rows;
    const int nrows = min(a.band_rows, K.nx - i0 + 1);
    const int nsteps = K.ny + nrows - 1;

    const int* f_prev  = (s > 0) ? a.prog + (size_t)(s - 1) * B + b : nullptr;
    const int* f_below = (s > 0 && b + 1 < B) ? a.prog + (size_t)(s - 1) * B + b + 1 : nullptr;
    const int* f_above = (b > 0) ? a.prog + (size_t)s * B + b - 1 : nullptr;
    int* my_flag = a.prog + (size_t)s * B + b;

    if (tid == 0) {
        sts_volatile(&s_sync[0], 0);
        sts_volatile(&s_sync[1], (f_prev || f_below || f_above) ? -WF_INF : WF_INF);
    }
    __syncthreads();
    acc_out = 0.0;

    if (tid >= NCOMP) {
        if (tid == NT - WF_SVC) {                       // ---- publisher
            int last = 0;
            while (last < nsteps) {
                const int d = lds_volatile(&s_sync[0]);
                if (d > last) { __threadfence(); st_release(my_flag, d); last = d; }
                else __nanosleep(20);
            }
        } else if (tid == NT - 32 && (f_prev || f_below || f_above)) {   // ---- poller
            const int nrows_below = (b + 1 < B) ? min(a.band_rows, K.nx - (i0 + a.band_rows) + 1) : 0;
            const int nsteps_below = K.ny + nrows_below - 1;
            const int nsteps_above = K.ny + a.band_rows - 1;
            int c1 = 0, c2 = 0, c3 = 0, cur = -WF_INF, spins = 0;
            while (true) {
                if (f_prev && c1 < nsteps) c1 = ld_relaxed(f_prev);
                if (f_below && c2 < nsteps_below) c2 = ld_relaxed(f_below);
                if (f_above && c3 < nsteps_above) c3 = ld_relaxed(f_above);
                const int a1 = (!f_prev || c1 >= nsteps) ? WF_INF : c1 - W;
                const int a2 = (!f_below || c2 >= nsteps_below) ? WF_INF : c2 - W + nrows;
                const int a3 = (!f_above || c3 >= nsteps_above) ? WF_INF : c3 - W - a.band_rows;
                const int al = min(a1, min(a2, a3));
                if (al > cur) {
                    __threadfence();                      // acquire side of the producers' release stores
                    sts_volatile(&s_sync[1], al);
                    cur = al; spins = 0;
                }
                if (al >= nsteps) break;
                if (++spins > a.spin_limit || ld_volatile(&a.ctrl->deadlock)) {
                    a.ctrl->deadlock = 1;                 // guard: never hang the GPU (host reports SRCFD_ERR_DEADLOCK)
                    sts_volatile(&s_sync[1], WF_INF);
                    break;
                }
                __nanosleep(20);
            }
        }
        __syncthreads();
        return;
    }

    // ------------------------------------------ compute threads --------------------------------
    const int r = tid - 2;
    const int irow = i0 + r;
    const bool regular = (r >= 0 && r < nrows);
    const bool passive = Q ? (r == -1 || r == -2 || r == nrows || r == nrows + 1) : (r == -1 || r == nrows);
    const bool live = regular || passive;

    // row bases; irow = -1 wraps to nx+1, irow = nx+2 runs on into the next plane (hazard H4)
    const long long rowoff = (irow < 0) ? (long long)(K.nx + 2 + irow) * K.pitch : (long long)irow * K.pitch;
    const double* rowp = a.Var + (long long)a.k * K.plane + (live ? rowoff : 0);
    double* wrow = a.Var + (long long)a.k * K.plane + rowoff;
    const double* aux0 = (OP == OP_PRESSURE) ? a.rhs + rowoff : a.VarOld + (long long)a.k * K.plane + rowoff;
    const double* auxF = a.Ff + rowoff;
    const int colmax = K.ny + 2;

    auto ldc = [&](int col) -> double {
        if (col == -1) col = K.ny + 1;
        return (live && col >= 0 && col <= colmax) ? __ldcg(rowp + col) : 0.0;
    };
    auto lda = [&](int col) -> WfAux<OP> {
        WfAux<OP> x;
        const bool ok = regular && col >= 1 && col <= K.ny;
        if constexpr (OP == OP_PRESSURE) {
            x.rhs = ok ? __ldg(aux0 + col) : 0.0;
        } else {
            if (ok) {
                x.vold = __ldg(aux0 + col);
                x.fE = __ldg(auxF + col); x.fN = __ldg(auxF + K.plane + col);
                x.fW = __ldg(auxF + 2 * K.plane + col); x.fS = __ldg(auxF + 3 * K.plane + col);
            } else { x.vold = x.fE = x.fN = x.fW = x.fS = 0.0; }
        }
        return x;
    };
    int allowed = -WF_INF;
    auto wait_allowed = [&](int tau) {
        while (allowed < tau) allowed = lds_volatile(&s_sync[1]);
    };

    // ---- window set-up at tau = -2: ring[m] = C[jj+m] ---------------------------------------------
    wait_allowed(-2);
    int jj = -2 - r + 1;
    double ring[R];
#pragma unroll
    for (int m = 0; m < R; ++m) ring[m] = ldc(jj + m);
    WfAux<OP> ax[AR];
#pragma unroll
    for (int m = 0; m < AR; ++m) ax[m] = lda(jj + m);
    double prev1 = regular ? ldc(0) : 0.0;          // (i, 0)   ghost column
    double prev2 = (regular && Q) ? ldc(-1) : 0.0;  // (i, -1) -> (i, ny+1) by index wrap
    double acc = 0.0;

    for (int tau0 = -2; tau0 < nsteps; tau0 += R) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int tau = tau0 + u;
            if (tau >= nsteps) break;
            bar_compute(NCOMP);
            const int pb = tau & 1, cb = pb ^ 1;     // read what step tau-1 published
            if (tid == 0 && tau > 0) sts_volatile(&s_sync[0], tau);   // steps < tau are complete
            wait_allowed(tau);

            const double w0 = ring[u % R], w1 = ring[(u + 1) % R], w2 = ring[(u + 2) % R];
            const double w3 = Q ? ring[(u + 3) % R] : 0.0;
            const WfAux<OP> x = ax[u % AR];
            const double im = s_new[pb * SP + tid - 1];
            const double ip = s_c[pb * SP + tid + 1];
            double im2 = 0.0, ip2 = 0.0;
            if (Q) { im2 = s_im[pb * SP + tid - 1]; ip2 = s_c2[pb * SP + tid + 2]; }

            double outv = w0;
            if (regular && jj >= 1 && jj <= K.ny) {
                double Rr, nv;
                if constexpr (OP == OP_PRESSURE) nv = pressure_cell(w0, ip, im, w1, prev1, x.rhs, K, Rr);
                else if constexpr (OP == OP_UPWIND)
                    nv = upwind_cell(w0, ip, im, w1, prev1, x.vold, x.fE, x.fN, x.fW, x.fS, K, Rr);
                else
                    nv = quick_cell(w0, ip, im, w1, prev1, ip2, im2, w2, prev2, x.vold, x.fE, x.fN, x.fW, x.fS, K, Rr);
                wrow[jj] = nv;
                acc += Rr * Rr;
                outv = nv;
                prev2 = prev1; prev1 = nv;
            }
            s_new[cb * SP + tid] = outv;
            s_c[cb * SP + tid] = w2;
            if (Q) { s_im[cb * SP + tid] = im; s_c2[cb * SP + tid] = w3; }

            // refill the slots just consumed: column jj+R of the row, column jj+AR of the inputs
            ring[u % R] = ldc(jj + R);
            ax[u % AR] = lda(jj + AR);
            ++jj;
        }
    }
    bar_compute(NCOMP);
    if (tid == 0) sts_volatile(&s_sync[0], nsteps);
    acc_out = acc;
    __syncthreads();
}

template <int OP>
__device__ void wf_run(const SolveArgs& a, int n_sweeps, double* smem, int* s_sync) {
    const int B = a.nbands;
    const int ntasks = n_sweeps * B;
    double* red = smem + 8 * (blockDim.x + 4) + 4;
    for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
        double acc;
        wf_task<OP>(a, t / B, t % B, smem, s_sync, acc);
        const double tot = block_sum(acc, red);
        if (threadIdx.x == 0) a.partials[t] = tot;
        __syncthreads();
    }
}

// rms of sweep s of the last group, from the per-band partials in band order
__device__ __forceinline__ double wf_sweep_rms(const SolveArgs& a, int s) {
    double ssq = 0.0;
    for (int b = 0; b < a.nbands; ++b) ssq += __ldcg(a.partials + (size_t)s * a.nbands + b);
    return sqrt(ssq / (double)((long long)a.K.nx * (long long)a.K.ny));
}

template <int OP>
__global__ void __launch_bounds__(WF_MAX_THREADS, 1) k_solve_gs(SolveArgs a) {
    cg::grid_group grid = cg::this_grid();
    if (a.ctrl->stop) return;
    extern __shared__ double smem[];
    __shared__ int s_first;
    __shared__ int s_sync[2];
    const Consts& K = a.K;
    double* A = a.Var + (long long)a.k * K.plane;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;

    int guess = a.ctrl->guess[a.slot] + a.guess_bias;
    guess = max(1, min(guess, a.max_iter));
    int n_done = 0, grow = 1;
    double last_rms = 0.0;
    bool first_group = true;
    while (true) {
        const int n_run = min(first_group ? guess : grow, a.max_iter - n_done);
        // snapshot + clear progress flags
        for (long long t = gtid; t < K.plane; t += gsize) a.scratch[t] = __ldcg(A + t);
        for (long long t = gtid; t < (long long)n_run * a.nbands; t += gsize) a.prog[t] = 0;
        if (threadIdx.x == 0) s_first = 0x7fffffff;
        grid.sync();
        wf_run<OP>(a, n_run, smem, s_sync);
        grid.sync();
        for (int s = threadIdx.x; s < n_run; s += blockDim.x)
            if (wf_sweep_rms(a, s) < a.tol) atomicMin(&s_first, s);
        __syncthreads();
        const int first = s_first;
        __syncthreads();
        if (first == 0x7fffffff) {            // no sweep of this group met the tolerance
            n_done += n_run;
            last_rms = wf_sweep_rms(a, n_run - 1);
            if (n_done >= a.max_iter) break;
            if (!first_group) grow = min(grow * 2, 64);
            first_group = false;
            continue;
        }
        if (first == n_run - 1) { n_done += n_run; last_rms = wf_sweep_rms(a, first); break; }
        // overshoot: roll back and run exactly first+1 sweeps
        last_rms = wf_sweep_rms(a, first);
        grid.sync();                          // everyone has read the partials of the speculative group
        for (long long t = gtid; t < K.plane; t += gsize) A[t] = __ldcg(a.scratch + t);
        for (long long t = gtid; t < (long long)(first + 1) * a.nbands; t += gsize) a.prog[t] = 0;
        grid.sync();
        wf_run<OP>(a, first + 1, smem, s_sync);
        n_done += first + 1;
        break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctrl->last_sweeps[a.slot] = n_done;
        a.ctrl->total_sweeps[a.slot] += n_done;
        a.ctrl->last_inner_rms[a.slot] = last_rms;
        a.ctrl->guess[a.slot] = n_done;
    }
}

}  // namespace srcfd
