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
                    const double nv = eval_cell<OP>(a, A, i, j, R);
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
//   * each thread reads its own row WF_D columns ahead into registers (L1-bypassing loads: the data is
//     produced by other SMs during the kernel).
// With W = WF_D + 5, before step tau of task (s,b) starts the following must have been published
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
constexpr int WF_D = 4;             // row look-ahead (steps) kept in registers
constexpr int WF_W = WF_D + 5;      // dependence look-ahead, see above
constexpr int WF_C = 4;             // publish progress every WF_C steps
constexpr int WF_MAX_THREADS = 512; // band rows + 4 passive threads, rounded to a warp multiple
constexpr int WF_MAX_BAND = WF_MAX_THREADS - 4;

template <int OP> struct WfAux;
template <> struct WfAux<OP_PRESSURE> { double rhs; };
template <> struct WfAux<OP_UPWIND> { double vold, fE, fN, fW, fS; };
template <> struct WfAux<OP_QUICK> { double vold, fE, fN, fW, fS; };

__device__ __forceinline__ void wf_wait(const int* flag, int need, int& cache, const SolveArgs& a) {
    if (cache >= need) return;
    int spins = 0;
    while ((cache = ld_acquire(flag)) < need) {
        if (++spins > a.spin_limit || ld_volatile(&a.ctrl->deadlock)) {
            a.ctrl->deadlock = 1;     // guard: never hang the GPU; the host reports SRCFD_ERR_DEADLOCK
            cache = 0x7fffffff;
            break;
        }
        __nanosleep(64);
    }
}

template <int OP>
__device__ void wf_task(const SolveArgs& a, const int s, const int b, double* smem, double& acc_out) {
    const Consts& K = a.K;
    constexpr bool Q = (OP == OP_QUICK);
    const int NT = blockDim.x, tid = threadIdx.x, SP = NT + 4;
    double* s_new = smem + 2;              // [2][SP], index = buf*SP + tid, valid tid range -2 .. NT+1
    double* s_c   = smem + 2 + 2 * SP;
    double* s_im  = smem + 2 + 4 * SP;
    double* s_c2  = smem + 2 + 6 * SP;

    const int B = a.nbands;
    const int i0 = 1 + b * a.band_rows;
    const int nrows = min(a.band_rows, K.nx - i0 + 1);
    const int nsteps = K.ny + nrows - 1;
    const int r = tid - 2;
    const int irow = i0 + r;
    const bool regular = (r >= 0 && r < nrows);
    const bool passive = Q ? (r == -1 || r == -2 || r == nrows || r == nrows + 1) : (r == -1 || r == nrows);
    const bool live = regular || passive;

    // row bases; irow = -1 wraps to nx+1, irow = nx+2 runs on into the next plane (hazard H4)
    const long long rowoff = (irow < 0) ? (long long)(K.nx + 2 + irow) * K.pitch : (long long)irow * K.pitch;
    const double* rowp = a.Var + (long long)a.k * K.plane + rowoff;
    double* wrow = a.Var + (long long)a.k * K.plane + rowoff;
    const double* aux0 = (OP == OP_PRESSURE) ? a.rhs + rowoff : a.VarOld + (long long)a.k * K.plane + rowoff;
    const double* auxF = a.Ff + rowoff;

    auto ldc = [&](int col) -> double {
        if (!live) return 0.0;
        if (col == -1) col = K.ny + 1;
        if (col < 0 || col > K.ny + 2) return 0.0;
        return __ldcg(rowp + col);
    };
    auto lda = [&](int col) -> WfAux<OP> {
        WfAux<OP> x;
        if constexpr (OP == OP_PRESSURE) {
            x.rhs = (regular && col >= 1 && col <= K.ny) ? __ldg(aux0 + col) : 0.0;
        } else {
            if (regular && col >= 1 && col <= K.ny) {
                x.vold = __ldg(aux0 + col);
                x.fE = __ldg(auxF + col); x.fN = __ldg(auxF + K.plane + col);
                x.fW = __ldg(auxF + 2 * K.plane + col); x.fS = __ldg(auxF + 3 * K.plane + col);
            } else { x.vold = x.fE = x.fN = x.fW = x.fS = 0.0; }
        }
        return x;
    };

    // ---- dependence bookkeeping (one polling thread) -------------------------------------------
    const int poller = NT - 1;
    const int* f_prev  = (s > 0) ? a.prog + (size_t)(s - 1) * B + b : nullptr;
    const int* f_below = (s > 0 && b + 1 < B) ? a.prog + (size_t)(s - 1) * B + b + 1 : nullptr;
    const int* f_above = (b > 0) ? a.prog + (size_t)s * B + b - 1 : nullptr;
    const int nrows_below = (b + 1 < B) ? min(a.band_rows, K.nx - (i0 + a.band_rows) + 1) : 0;
    const int nsteps_below = K.ny + nrows_below - 1;
    const int nsteps_above = K.ny + a.band_rows - 1;
    int c_prev = 0, c_below = 0, c_above = 0;
    auto ensure = [&](int tau) {
        if (f_prev) wf_wait(f_prev, min(nsteps, tau + WF_W), c_prev, a);
        if (f_below) { const int need = min(nsteps_below, tau + WF_W - nrows); if (need > 0) wf_wait(f_below, need, c_below, a); }
        if (f_above) wf_wait(f_above, min(nsteps_above, tau + WF_W + a.band_rows), c_above, a);
    };
    int* my_flag = a.prog + (size_t)s * B + b;

    if (tid == poller) ensure(-2);
    __syncthreads();

    // ---- window set-up at tau = -2 ---------------------------------------------------------------
    int jj = -2 - r + 1;
    double w0 = ldc(jj), w1 = ldc(jj + 1), w2 = ldc(jj + 2), w3 = ldc(jj + 3);
    double q[WF_D];
#pragma unroll
    for (int d = 0; d < WF_D; ++d) q[d] = ldc(jj + 4 + d);
    WfAux<OP> ax[WF_D];
#pragma unroll
    for (int d = 0; d < WF_D; ++d) ax[d] = lda(jj + d);
    double prev1 = regular ? ldc(0) : 0.0;      // (i, 0)   ghost column
    double prev2 = (regular && Q) ? ldc(-1) : 0.0;  // (i, -1) -> (i, ny+1) by index wrap
    double acc = 0.0;

    for (int tau = -2; tau < nsteps; ++tau) {
        __syncthreads();
        const int pb = tau & 1, cb = pb ^ 1;     // buffers: read what step tau-1 published
        if (tid == 0 && tau > 0 && (tau % WF_C) == 0) { __threadfence(); st_release(my_flag, tau); }
        const double qn = ldc(jj + 4 + WF_D);
        const WfAux<OP> an = lda(jj + WF_D);

        const double im = s_new[pb * SP + tid - 1];
        const double ip = s_c[pb * SP + tid + 1];
        double im2 = 0.0, ip2 = 0.0;
        if (Q) { im2 = s_im[pb * SP + tid - 1]; ip2 = s_c2[pb * SP + tid + 2]; }

        double outv = w0;
        if (regular && jj >= 1 && jj <= K.ny) {
            double R, nv;
            if constexpr (OP == OP_PRESSURE) nv = pressure_cell(w0, ip, im, w1, prev1, ax[0].rhs, K, R);
            else if constexpr (OP == OP_UPWIND)
                nv = upwind_cell(w0, ip, im, w1, prev1, ax[0].vold, ax[0].fE, ax[0].fN, ax[0].fW, ax[0].fS, K, R);
            else
                nv = quick_cell(w0, ip, im, w1, prev1, ip2, im2, w2, prev2, ax[0].vold, ax[0].fE, ax[0].fN, ax[0].fW, ax[0].fS, K, R);
            wrow[jj] = nv;
            acc += R * R;
            outv = nv;
            prev2 = prev1; prev1 = nv;
        }
        s_new[cb * SP + tid] = outv;
        s_c[cb * SP + tid] = w2;
        if (Q) { s_im[cb * SP + tid] = im; s_c2[cb * SP + tid] = w3; }

        w0 = w1; w1 = w2; w2 = w3; w3 = q[0];
#pragma unroll
        for (int d = 0; d + 1 < WF_D; ++d) { q[d] = q[d + 1]; ax[d] = ax[d + 1]; }
        q[WF_D - 1] = qn; ax[WF_D - 1] = an;
        ++jj;
        if (tid == poller) ensure(tau + 1);
    }
    __syncthreads();
    acc_out = acc;
}

template <int OP>
__device__ void wf_run(const SolveArgs& a, int n_sweeps, double* smem) {
    const int B = a.nbands;
    const int ntasks = n_sweeps * B;
    double* red = smem + 8 * (blockDim.x + 4) + 4;
    for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const int s = t / B, b = t % B;
        double acc;
        wf_task<OP>(a, s, b, smem, acc);
        const double tot = block_sum(acc, red);
        if (threadIdx.x == 0) {
            a.partials[t] = tot;
            const int nrows = min(a.band_rows, a.K.nx - (1 + b * a.band_rows) + 1);
            __threadfence();
            st_release(a.prog + t, a.K.ny + nrows - 1);
        }
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
        wf_run<OP>(a, n_run, smem);
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
        wf_run<OP>(a, first + 1, smem);
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
