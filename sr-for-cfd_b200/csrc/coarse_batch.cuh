// coarse_batch.cuh -- whole CFDSolver.solve() runs of MANY small cases in one launch: one CTA per case, the
// complete solver state (Var, VarOld, Ff) in shared memory, no host round trip and no global memory traffic
// between the first and the last outer iteration.
//
// This is the coarse stage of the ML-accelerated workflow (run_coarse_simulation, PyCFD_ML_accelerated.py:696-761 /
// bfs_ml_accelerated.py:893-977): up to 100 000 outer iterations of a 10x10 problem per case.  On a grid that small
// the whole-GPU kernels are pure launch latency; here an outer iteration costs a few hundred block barriers, and an
// ensemble's coarse solves run side by side, one SM each.
//
// Sweep order: the reference's in-place lexicographic Gauss-Seidel (numba on one thread), bit for bit.  Cell (i,j)
// of sweep s runs at step  t = L*s + i + j  (L = 2 for the 5-point stencils, 3 for QUICK, whose second neighbours
// would otherwise be overwritten one step too early), so at any step 1/L of the cells -- belonging to up to
// (nx+ny)/L consecutive sweeps -- update concurrently and every stencil value a cell reads is exactly the one the
// sequential loop would have read.  One thread owns L consecutive cells of a row and updates one of them per step.
// A dedicated warp adds up each finished sweep's R^2 (fixed order) one step behind the workers and raises `hit`
// when the rms meets the tolerance (LDC.py:266-268; evaluated as an exactly equivalent threshold on the sum, see
// coarse_threshold); its verdict rides on the step barrier.  Sweeps past the hit have then been started already, so
// an inner solve runs in segments: snapshot, `limit` sweeps (the previous outer iteration's count, or 2 below it for
// long solves), and only when the tolerance was met before the limit: restore and replay exactly that many sweeps.
#pragma once
#include "common.cuh"
#include "cell_ops.cuh"
#include "inner_gs2.cuh"

namespace srcfd {

struct CoarseCase {
    Consts K;
    BcSpec bc;
    int scheme;              // SRCFD_SCHEME_*
    int relax_enabled;
    double relax[3];
    double inner_tol;
    int inner_max;
    double crit[3];
    long long max_iterations;
};

struct CoarseOut {
    long long iterations;
    int converged, nan_flag;
    double rms[3];
    long long total_sweeps[3];
    long long n_hist;
    double last_inner_rms[3];
    double residual[3];      // the sums correct_velocity accumulated in the last iteration (CFDSolver.residual)
    int last_sweeps[3];
};

struct CoarseShared {
    int hit;
    int flag;
    double last_v;           // sum of R^2 of the last sweep the reducer looked at
    double v_hit;            // smallest sum whose rms is NOT below the tolerance (see coarse_threshold)
    double red[3][32];
};

// The break test of LDC.py:266-268 is  sqrt(sum / (Nx*Ny)) < tolerance.  Division by a positive constant and sqrt are
// correctly rounded, hence monotone, so the test is EXACTLY  sum < T  with T the smallest non-negative double whose rms
// is >= tolerance -- found once per case by bisection on the bit pattern.  That takes a division and a square root (two
// long dependent chains) off every sweep's critical path.
__device__ inline double coarse_threshold(double tol, double ncells) {
    unsigned long long lo = 0ull, hi = 0x7ff0000000000000ull;          // +0 .. +inf
    if (!(sqrt(__longlong_as_double((long long)hi) / ncells) >= tol)) return 0.0;   // NaN tolerance: never below
    while (lo < hi) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (sqrt(__longlong_as_double((long long)mid) / ncells) >= tol) hi = mid; else lo = mid + 1;
    }
    return __longlong_as_double((long long)lo);
}

// `limit` pipelined sweeps of plane k.  Returns the 1-based index of the first sweep whose rms met the tolerance
// (check only), 0 when none did.  Block-uniform; ends with every thread past a barrier.
// A worker's active steps are consecutive (from t_first on it updates one cell per step), so its position inside the
// sweep (r), the sweep (s) and the ring slot advance by counters: no division in the step loop.  The loop-invariant
// divisors go through the exact reciprocal sequence of inner_gs2.cuh (same bits as '/', a third of the latency).
template <int OP>
__device__ __forceinline__ int coarse_sweeps(double* __restrict__ Var, const double* __restrict__ VarOld,
                                             const double* __restrict__ Ff, const double* __restrict__ rhs,
                                             double* __restrict__ part, CoarseShared* sh, int k, int limit, bool check,
                                             const Consts& K, const Gs2Div& D, double tol, int ring, int stride) {
    constexpr int L = (OP == OP_QUICK) ? 3 : 2;
    const int nx = K.nx, ny = K.ny, pitch = K.pitch;
    const int P = (int)K.plane;
    const int M = (ny + L - 1) / L;
    const int nwork = nx * M;
    const int tid = threadIdx.x;
    const bool reducer = tid >= (int)blockDim.x - 32;
    const bool worker = tid < nwork;
    const int i0 = tid / M, m = tid - i0 * M;
    const int tail = nx - 1 + L * M - 1;      // steps from a sweep's first cell until every worker has posted its partial sum
    const int t_end = L * (limit - 1) + tail + (check ? 1 : 0);
    double* Vk = Var + k * P;
    const double* Ok = VarOld + k * P;
    const int t_first = worker ? i0 + L * m : 0x7fffffff;
    const int c0 = (i0 + 1) * pitch + L * m + 1;
    const int jleft = ny - L * m;             // cells of this thread inside the row: r < jleft
    // QUICK second neighbours on the flat buffer (SURVEY.md hazard H4): a negative index wraps to the other end of its
    // axis, one past the end runs on into the next row / plane
    const int off_im2 = ((i0 - 1 < 0) ? i0 - 1 + nx + 2 : i0 - 1) * pitch - (i0 + 1) * pitch;
    int r = 0, s = 0, slot = 0;               // worker: position in the sweep, sweep, ring slot
    int next_t = tail + 1, s_r = 0, slot_r = 0;   // reducer: step at which sweep s_r is complete
    double acc = 0.0;
    if (tid == 0) sh->hit = 0;
    __syncthreads();
    const double v_hit = sh->v_hit;
    int found = 0;
    for (int t = 0; t <= t_end; ++t) {
        if (t >= t_first && s < limit) {
            if (r == 0) acc = 0.0;
            if (r < jleft) {
                const int c = c0 + r;
                const double cc = Vk[c], ip = Vk[c + pitch], im = Vk[c - pitch], jp = Vk[c + 1], jm = Vk[c - 1];
                double R, nv;
                if (OP == OP_PRESSURE) {
                    nv = pressure_cell2(cc, ip, im, jp, jm, rhs[c], K, D, R);
                } else if (OP == OP_UPWIND) {
                    nv = upwind_cell2(cc, ip, im, jp, jm, Ok[c], Ff[c], Ff[P + c], Ff[2 * P + c], Ff[3 * P + c], K, D, R);
                } else {
                    const int j0 = L * m + r;
                    const int cm2 = (j0 - 1 < 0) ? j0 - 1 + ny + 2 : j0 - 1;
                    const double ip2 = Vk[c + 2 * pitch], im2 = Vk[c + off_im2];
                    const double jp2 = Vk[c + 2], jm2 = Vk[(i0 + 1) * pitch + cm2];
                    nv = quick_cell2(cc, ip, im, jp, jm, ip2, im2, jp2, jm2, Ok[c], Ff[c], Ff[P + c], Ff[2 * P + c],
                                     Ff[3 * P + c], K, D, R);
                }
                Vk[c] = nv;
                acc += R * R;
            }
            if (++r == L) {
                if (check) part[slot * stride + tid] = acc;
                r = 0; ++s;
                if (++slot == ring) slot = 0;
            }
        } else if (reducer && check && t == next_t && s_r < limit) {
            const int lane = tid & 31;
            const double* ps = part + slot_r * stride;
            double v = 0.0;
            for (int w = lane; w < nwork; w += 32) v += ps[w];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            found = v < v_hit;
            if (lane == 0) {
                sh->last_v = v;
                if (found) sh->hit = s_r + 1;
            }
            next_t += L; ++s_r;
            if (++slot_r == ring) slot_r = 0;
        }
        if (__syncthreads_or(found)) break;      // the step barrier also carries the reducer's verdict
    }
    __syncthreads();
    const int hit = sh->hit;
    __syncthreads();
    return hit;
}

// One inner solve (LDC.py:248-314): <= inner_max sweeps, stop after the first whose rms < tolerance.
// Guess policy: a run that meets the tolerance BEFORE its limit costs a restore and a full replay, one that falls
// short only a short follow-up segment.  Long solves (>= 32 sweeps, whose counts drift down by a sweep every few outer
// iterations) therefore aim 2 sweeps below the previous count and finish in segments of 4, 8, 16, ... sweeps.
template <int OP>
__device__ int coarse_inner(double* Var, const double* VarOld, const double* Ff, const double* rhs, double* snap,
                            double* part, CoarseShared* sh, int k, int& guess, const Consts& K, const Gs2Div& D,
                            double tol, int maxs, int ring, int stride) {
    const int ncell = K.nx * K.ny, P = (int)K.plane;
    double* Vk = Var + k * P;
    const bool under = guess >= 32;
    int done = 0, g = guess < 1 ? 1 : (under ? guess - 2 : guess), follow = 4;
    while (true) {
        for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
            const int c = (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
            snap[c] = Vk[c];
        }
        const int limit = g < maxs - done ? g : maxs - done;
        const int hit = coarse_sweeps<OP>(Var, VarOld, Ff, rhs, part, sh, k, limit, true, K, D, tol, ring, stride);
        if (hit == 0) {
            done += limit;
            if (done >= maxs) break;
            if (under) { g = follow; follow *= 2; }
            else g = limit / 4 < 4 ? 4 : limit / 4;
            continue;
        }
        if (hit < limit) {
            for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
                const int c = (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
                Vk[c] = snap[c];
            }
            coarse_sweeps<OP>(Var, VarOld, Ff, rhs, part, sh, k, hit, false, K, D, tol, ring, stride);
        }
        done += hit;
        break;
    }
    guess = done;
    return done;
}

// _apply_bc_wrapper (LDC.py:391-394 / BFS.py:564-569) on the shared-memory planes; same arithmetic as k_apply_bc.
__device__ __forceinline__ void coarse_bc(double* Var, int k, const Consts& K, const BcSpec& bc) {
    const int P = (int)K.plane;
    double* V = Var + k * P;
    const int n = K.nx > K.ny ? K.nx : K.ny;
    for (int t = threadIdx.x + 1; t <= n; t += blockDim.x) {
        if (t <= K.ny) {
            const int j = t;
            if (bc.types[k][0] == 0) V[j] = 2 * bc.values[k][0] - V[K.pitch + j];
            else                     V[j] = V[K.pitch + j];
            const int r = (K.nx + 1) * K.pitch + j, q = K.nx * K.pitch + j;
            if (bc.types[k][1] == 0) V[r] = 2 * bc.values[k][1] - V[q];
            else                     V[r] = V[q];
            if (bc.bfs && (k == 0 || k == 1)) {
                const double y = (j - 0.5) * K.dy;
                if (y < bc.step_h) {
                    V[j] = -V[K.pitch + j];
                } else if (k == 1) {
                    V[j] = -V[K.pitch + j];
                } else {
                    double yprime = y - bc.step_h;
                    if (yprime < 0.0) yprime = 0.0;
                    if (yprime > bc.h) yprime = bc.h;
                    const double u_in = 6.0 * bc.Ub * (yprime / bc.h) * (1.0 - (yprime / bc.h));
                    V[j] = 2.0 * u_in - V[K.pitch + j];
                    double* Vv = Var + P;
                    Vv[j] = -Vv[K.pitch + j];
                }
            }
        }
        if (t <= K.nx) {
            const int row = t * K.pitch;
            if (bc.types[k][2] == 0) V[row + K.ny + 1] = 2 * bc.values[k][2] - V[row + K.ny];
            else                     V[row + K.ny + 1] = V[row + K.ny];
            if (bc.types[k][3] == 0) V[row] = 2 * bc.values[k][3] - V[row + 1];
            else                     V[row] = V[row + 1];
        }
    }
    __syncthreads();
}

// LDC.py:147-154 linear_interpolation + the pressure right-hand side rho/dt*(fE+fN+fW+fS) (LDC.py:305)
__device__ __forceinline__ void coarse_interpolate(const double* Var, double* Ff, double* rhs, const Consts& K) {
    const int P = (int)K.plane, ncell = K.nx * K.ny;
    const double* U = Var; const double* V = Var + P;
    for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
        const int c = (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
        const double fE = (U[c] + U[c + K.pitch]) * K.dy * 0.5;
        const double fN = (V[c] + V[c + 1]) * K.dx * 0.5;
        const double fW = -(U[c] + U[c - K.pitch]) * K.dy * 0.5;
        const double fS = -(V[c] + V[c - 1]) * K.dx * 0.5;
        Ff[c] = fE; Ff[P + c] = fN; Ff[2 * P + c] = fW; Ff[3 * P + c] = fS;
        rhs[c] = K.rho_dt * (fE + fN + fW + fS);
    }
    __syncthreads();
}

// CFDSolver.__init__ + solve() (LDC.py:333-430 / BFS.py:473-706) of case blockIdx.x.  resume != 0 starts from the
// state in gVar/gVarOld/gFf instead of _initialize_fields.
__global__ void __launch_bounds__(640, 1) k_coarse_solve(const CoarseCase* __restrict__ cases, double* __restrict__ gVar, double* __restrict__ gVarOld,
                               double* __restrict__ gFf, CoarseOut* __restrict__ out, double* __restrict__ hist,
                               long long hist_cap, int ring, int stride, int resume) {
    extern __shared__ double co_sm[];
    __shared__ CoarseShared sh;
    __shared__ CoarseCase cs_sh;
    if (threadIdx.x == 0) cs_sh = cases[blockIdx.x];
    __syncthreads();
    const Consts K = cs_sh.K;
    const int P = (int)K.plane, ncell = K.nx * K.ny, tid = threadIdx.x, nth = blockDim.x;
    double* Var = co_sm;
    double* VarOld = Var + 3 * P;
    double* Ff = VarOld + 3 * P;
    double* rhs = Ff + 4 * P;
    double* snap = rhs + P;
    double* part = snap + P;
    const long long cb = blockIdx.x;
    double* myhist = hist ? hist + cb * hist_cap * 3 : nullptr;

    if (resume) {
        for (int t = tid; t < 3 * P; t += nth) { Var[t] = gVar[cb * 3 * P + t]; VarOld[t] = gVarOld[cb * 3 * P + t]; }
        for (int t = tid; t < 4 * P; t += nth) Ff[t] = gFf[cb * 4 * P + t];
        for (int t = tid; t < 2 * P; t += nth) rhs[t] = 0.0;
        __syncthreads();
    } else {
        // _initialize_fields (LDC.py:377-389)
        for (int t = tid; t < 12 * P; t += nth) co_sm[t] = 0.0;
        __syncthreads();
        for (int k = 0; k < 3; ++k) coarse_bc(Var, k, K, cs_sh.bc);
        for (int t = tid; t < 3 * P; t += nth) VarOld[t] = Var[t];
        coarse_interpolate(Var, Ff, rhs, K);
    }

    long long count = 0, total[3] = {0, 0, 0}, n_hist = 0;
    int converged = 0, nan_flag = 0;
    int guess[3] = {4, 4, 32};
    double rms[3] = {0.0, 0.0, 0.0}, inner_rms[3] = {0.0, 0.0, 0.0}, resid[3] = {0.0, 0.0, 0.0};
    int last_n[3] = {0, 0, 0};
    const double tol = cs_sh.inner_tol;
    const int maxs = cs_sh.inner_max;
    if (tid == 0) { sh.v_hit = coarse_threshold(tol, (double)((long long)K.nx * (long long)K.ny)); sh.last_v = 0.0; }
    __syncthreads();
    Gs2Div D;
    D.dx2 = make_invdiv(K.dx2); D.dy2 = make_invdiv(K.dy2); D.apd = make_invdiv(K.ap_d);
    while (!converged && count < cs_sh.max_iterations) {
        count += 1;
        // _implicit_solve (LDC.py:432-467 / BFS.py:622-673)
        for (int k = 0; k < 2; ++k) {
            int n;
            if (cs_sh.scheme == 1) n = coarse_inner<OP_QUICK>(Var, VarOld, Ff, rhs, snap, part, &sh, k, guess[k], K, D, tol, maxs, ring, stride);
            else                   n = coarse_inner<OP_UPWIND>(Var, VarOld, Ff, rhs, snap, part, &sh, k, guess[k], K, D, tol, maxs, ring, stride);
            total[k] += n; last_n[k] = n;
            inner_rms[k] = sqrt(sh.last_v / (double)((long long)K.nx * (long long)K.ny));
            if (cs_sh.relax_enabled) {
                const double a = cs_sh.relax[k];
                for (int idx = tid; idx < ncell; idx += nth) {
                    const int o = k * P + (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
                    Var[o] = VarOld[o] + a * (Var[o] - VarOld[o]);
                }
            }
            __syncthreads();
            coarse_bc(Var, k, K, cs_sh.bc);
        }
        coarse_interpolate(Var, Ff, rhs, K);
        {
            const int n = coarse_inner<OP_PRESSURE>(Var, VarOld, Ff, rhs, snap, part, &sh, 2, guess[2], K, D, tol, maxs, ring, stride);
            total[2] += n; last_n[2] = n;
            inner_rms[2] = sqrt(sh.last_v / (double)((long long)K.nx * (long long)K.ny));
            if (cs_sh.relax_enabled) {
                const double a = cs_sh.relax[2];
                for (int idx = tid; idx < ncell; idx += nth) {
                    const int o = 2 * P + (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
                    Var[o] = VarOld[o] + a * (Var[o] - VarOld[o]);
                }
            }
            __syncthreads();
            coarse_bc(Var, 2, K, cs_sh.bc);
        }
        // correct_velocity (LDC.py:316-328) with the three residual sums
        double du2 = 0.0, dv2 = 0.0, dp2 = 0.0;
        {
            double* U = Var; double* V = Var + P; const double* Pp = Var + 2 * P;
            const double* UO = VarOld; const double* VO = VarOld + P; const double* PO = VarOld + 2 * P;
            for (int idx = tid; idx < ncell; idx += nth) {
                const int c = (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
                const double u = U[c] - K.dt_rho * (Pp[c + K.pitch] - Pp[c - K.pitch]) / K.two_dx;
                const double v = V[c] - K.dt_rho * (Pp[c + 1] - Pp[c - 1]) / K.two_dy;
                U[c] = u; V[c] = v;
                const double du = u - UO[c], dv = v - VO[c], dp = Pp[c] - PO[c];
                du2 += du * du; dv2 += dv * dv; dp2 += dp * dp;
            }
        }
        const double su = block_sum(du2, sh.red[0]);
        const double sv = block_sum(dv2, sh.red[1]);
        const double sp = block_sum(dp2, sh.red[2]);
        resid[0] = su; resid[1] = sv; resid[2] = sp;
        __syncthreads();
        coarse_bc(Var, 0, K, cs_sh.bc);
        coarse_bc(Var, 1, K, cs_sh.bc);
        // update_flux (LDC.py:239-246)
        {
            const double* Pp = Var + 2 * P;
            for (int idx = tid; idx < ncell; idx += nth) {
                const int c = (idx / K.ny + 1) * K.pitch + idx % K.ny + 1;
                const double p = Pp[c];
                Ff[c]         += K.mdt_rho * (Pp[c + K.pitch] - p) * K.dy / K.dx;
                Ff[P + c]     += K.mdt_rho * (Pp[c + 1] - p) * K.dx / K.dy;
                Ff[2 * P + c] += K.mdt_rho * (Pp[c - K.pitch] - p) * K.dy / K.dx;
                Ff[3 * P + c] += K.mdt_rho * (Pp[c - 1] - p) * K.dx / K.dy;
            }
        }
        // _convergence_check (LDC.py:469-501)
        if (tid == 0) {
            const double res[3] = {su, sv, sp};
            bool bad = false, conv = true;
            for (int k = 0; k < 3; ++k) {
                double r = sqrt(res[k] / (double)((long long)K.nx * (long long)K.ny));
                r = r / K.dt;
                sh.red[0][k] = r;
                if (isnan(r) || isinf(r)) bad = true;
                if (r > cs_sh.crit[k]) conv = false;
            }
            sh.flag = bad ? 2 : (conv ? 1 : 0);
        }
        __syncthreads();
        const int flag = sh.flag;
        rms[0] = sh.red[0][0]; rms[1] = sh.red[0][1]; rms[2] = sh.red[0][2];
        if (flag == 2) { nan_flag = 1; __syncthreads(); break; }
        converged = (flag == 1);
        if (count % 100 == 0 && myhist && n_hist < hist_cap) {
            if (tid == 0) { myhist[3 * n_hist] = rms[0]; myhist[3 * n_hist + 1] = rms[1]; myhist[3 * n_hist + 2] = rms[2]; }
            n_hist += 1;
        }
        if (!converged)
            for (int t = tid; t < 3 * P; t += nth) VarOld[t] = Var[t];
        __syncthreads();
    }

    for (int t = tid; t < 3 * P; t += nth) { gVar[cb * 3 * P + t] = Var[t]; gVarOld[cb * 3 * P + t] = VarOld[t]; }
    for (int t = tid; t < 4 * P; t += nth) gFf[cb * 4 * P + t] = Ff[t];
    if (tid == 0) {
        CoarseOut o;
        o.iterations = count; o.converged = converged; o.nan_flag = nan_flag; o.n_hist = n_hist;
        for (int k = 0; k < 3; ++k) { o.rms[k] = rms[k]; o.total_sweeps[k] = total[k]; o.last_inner_rms[k] = inner_rms[k];
                                      o.residual[k] = resid[k]; o.last_sweeps[k] = last_n[k]; }
        out[blockIdx.x] = o;
    }
}

}  // namespace srcfd
