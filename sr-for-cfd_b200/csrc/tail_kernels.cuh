// tail_kernels.cuh -- the once-per-outer-iteration kernels of _implicit_solve / _convergence_check.
// All of them are streaming, HBM/L2-bound passes: one thread per interior cell, j (the contiguous
// axis) fastest across a warp so every global access is coalesced.
#pragma once
#include "common.cuh"

namespace srcfd {

constexpr int TAIL_THREADS = 256;

__device__ __forceinline__ bool cell_of_thread(const Consts& K, int& i, int& j, long long& c) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)K.nx * K.ny) return false;
    i = (int)(idx / K.ny) + 1;
    j = (int)(idx % K.ny) + 1;
    c = (long long)i * K.pitch + j;
    return true;
}

// LDC.py:110-115 copy_new_to_old (all three planes, ghosts included)
__global__ void k_copy_new_to_old(const double* __restrict__ Var, double* __restrict__ VarOld, long long n,
                                  const Ctrl* ctrl) {
    if (ctrl->stop) return;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        VarOld[t] = Var[t];
}

// _apply_bc_wrapper: apply_bc_configured (LDC.py:117-145) then _apply_bfs_inlet (BFS.py:524-562).
// One thread per boundary index t in [1, max(nx,ny)]; corners are never written.
// mode: 0 = both (the wrapper), 1 = apply_bc_configured only, 2 = _apply_bfs_inlet only.
// nk > 1: planes k .. k+nk-1 in ONE launch, each thread taking its boundary index through the planes in order -- every cell
// a pass touches belongs to its own index (corners are never written), so this is the sequence of separate launches.
__global__ void k_apply_bc(double* __restrict__ Var, int k0, Consts K, BcSpec bc, int mode, const Ctrl* ctrl, int nk) {
    if (ctrl->stop) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x + 1;
  for (int k = k0; k < k0 + nk; ++k) {
    const bool generic = (mode != 2), inlet = (mode != 1) && bc.bfs && (k == 0 || k == 1) && !bc.skip_lo;
    double* V = Var + (long long)k * K.plane;
    if (mode == 4) {
        // only the side effect of the k = 0 inlet pass on the v ghost column (BFS.py:562), so that a paired u/v solve
        // sees the column the reference's v solve would see
        if (bc.bfs && !bc.skip_lo && t <= K.ny && (t - 0.5) * K.dy >= bc.step_h) {
            double* Vv = Var + K.plane;
            Vv[t] = -Vv[K.pitch + t];
        }
        return;
    }
    if (t <= K.ny) {
        const int j = t;
        if (generic && !bc.skip_lo) {
            if (bc.types[k][0] == 0) V[j] = 2 * bc.values[k][0] - V[K.pitch + j];
            else                     V[j] = V[K.pitch + j];
        }
        if (generic && !bc.skip_hi) {
            const long long r = (long long)(K.nx + 1) * K.pitch + j, q = (long long)K.nx * K.pitch + j;
            if (bc.types[k][1] == 0) V[r] = 2 * bc.values[k][1] - V[q];
            else                     V[r] = V[q];
        }
        if (inlet) {
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
                double* Vv = Var + K.plane;          // the k=0 pass also resets the v ghost (BFS.py:562)
                Vv[j] = -Vv[K.pitch + j];
            }
        }
    }
    if (generic && t <= K.nx) {
        const long long row = (long long)t * K.pitch;
        if (bc.types[k][2] == 0) V[row + K.ny + 1] = 2 * bc.values[k][2] - V[row + K.ny];
        else                     V[row + K.ny + 1] = V[row + K.ny];
        if (bc.types[k][3] == 0) V[row] = 2 * bc.values[k][3] - V[row + 1];
        else                     V[row] = V[row + 1];
    }
  }
}

// LDC.py:147-154 linear_interpolation; also emits rhs = rho/dt*(fE+fN+fW+fS) (LDC.py:305) when rhs != null,
// which is exactly what solve_pressure recomputes per sweep from the same four fluxes.
__global__ void k_linear_interpolation(const double* __restrict__ Var, double* __restrict__ Ff,
                                       double* __restrict__ rhs, Consts K, const Ctrl* ctrl) {
    if (ctrl->stop) return;
    int i, j; long long c;
    if (!cell_of_thread(K, i, j, c)) return;
    const double* U = Var; const double* V = Var + K.plane;
    const double fE = (U[c] + U[c + K.pitch]) * K.dy * 0.5;
    const double fN = (V[c] + V[c + 1]) * K.dx * 0.5;
    const double fW = -(U[c] + U[c - K.pitch]) * K.dy * 0.5;
    const double fS = -(V[c] + V[c - 1]) * K.dx * 0.5;
    Ff[c] = fE; Ff[K.plane + c] = fN; Ff[2 * K.plane + c] = fW; Ff[3 * K.plane + c] = fS;
    if (rhs) rhs[c] = K.rho_dt * (fE + fN + fW + fS);
}

// rhs from an existing Ff (kernel-level solve_pressure entry point)
__global__ void k_pressure_rhs(const double* __restrict__ Ff, double* __restrict__ rhs, Consts K, const Ctrl* ctrl) {
    if (ctrl->stop) return;
    int i, j; long long c;
    if (!cell_of_thread(K, i, j, c)) return;
    rhs[c] = K.rho_dt * (Ff[c] + Ff[K.plane + c] + Ff[2 * K.plane + c] + Ff[3 * K.plane + c]);
}

// LDC.py:239-246 update_flux
__global__ void k_update_flux(const double* __restrict__ Var, double* __restrict__ Ff, Consts K, const Ctrl* ctrl) {
    if (ctrl->stop) return;
    int i, j; long long c;
    if (!cell_of_thread(K, i, j, c)) return;
    const double* P = Var + 2 * K.plane;
    const double p = P[c];
    Ff[c]               += K.mdt_rho * (P[c + K.pitch] - p) * K.dy / K.dx;
    Ff[K.plane + c]     += K.mdt_rho * (P[c + 1] - p) * K.dx / K.dy;
    Ff[2 * K.plane + c] += K.mdt_rho * (P[c - K.pitch] - p) * K.dy / K.dx;
    Ff[3 * K.plane + c] += K.mdt_rho * (P[c - 1] - p) * K.dx / K.dy;
}

// BFS.py:371-375 under_relax_field
// k2 >= 0: plane k2 with factor alpha2 in the same launch (u and v after a paired momentum solve)
__global__ void k_under_relax(double* __restrict__ Var, const double* __restrict__ VarOld, int k, double alpha,
                              Consts K, const Ctrl* ctrl, int k2, double alpha2) {
    if (ctrl->stop) return;
    int i, j; long long c;
    if (!cell_of_thread(K, i, j, c)) return;
    const long long o = (long long)k * K.plane + c;
    Var[o] = VarOld[o] + alpha * (Var[o] - VarOld[o]);
    if (k2 >= 0) {
        const long long o2 = (long long)k2 * K.plane + c;
        Var[o2] = VarOld[o2] + alpha2 * (Var[o2] - VarOld[o2]);
    }
}

// LDC.py:316-328 correct_velocity.  Residual sums: per-block partials in a fixed order, finished by
// k_residual_finish -- deterministic, but not the reference's sequential order (differs in the last bits).
__global__ void k_correct_velocity(double* __restrict__ Var, const double* __restrict__ VarOld,
                                   double* __restrict__ partials, Consts K, const Ctrl* ctrl, int res_r0, int res_r1) {
    if (ctrl->stop) return;
    __shared__ double scratch[3][32];
    int i, j; long long c;
    double du2 = 0.0, dv2 = 0.0, dp2 = 0.0;
    if (cell_of_thread(K, i, j, c)) {
        double* U = Var; double* V = Var + K.plane; const double* P = Var + 2 * K.plane;
        const double* UO = VarOld; const double* VO = VarOld + K.plane; const double* PO = VarOld + 2 * K.plane;
        const double u = U[c] - K.dt_rho * (P[c + K.pitch] - P[c - K.pitch]) / K.two_dx;
        const double v = V[c] - K.dt_rho * (P[c + 1] - P[c - 1]) / K.two_dy;
        U[c] = u; V[c] = v;
        const double du = u - UO[c], dv = v - VO[c], dp = P[c] - PO[c];
        if (i >= res_r0 && i <= res_r1) { du2 = du * du; dv2 = dv * dv; dp2 = dp * dp; }   // a slab counts its own rows only
    }
    const double su = block_sum(du2, scratch[0]);
    const double sv = block_sum(dv2, scratch[1]);
    const double sp = block_sum(dp2, scratch[2]);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x + 0] = su; partials[3 * blockIdx.x + 1] = sv; partials[3 * blockIdx.x + 2] = sp;
    }
}

// residual[k] += sum of partials (fixed order: strided lanes, then block_sum).  One block.
// out != null: the three sums go there instead (a slab hands them to the exchange that adds the ranks up).
// assign: residual[k] = tot (the composed iteration: the reference zeroes the array at the top of _implicit_solve and nothing
// else adds to it, and 0.0 + tot == tot) -- saves the zeroing launch.
__global__ void k_residual_finish(const double* __restrict__ partials, int nblocks, Ctrl* ctrl, double* __restrict__ out, int assign) {
    if (ctrl->stop) return;
    __shared__ double scratch[32];
    for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[3 * b + k];
        const double tot = block_sum(s, scratch);
        if (threadIdx.x == 0) { if (out) out[k] = tot; else if (assign) ctrl->residual[k] = 0.0 + tot; else ctrl->residual[k] += tot; }
        __syncthreads();
    }
}

// A solve()/step() that ended converged or with NaN leaves stop = 1 so that the launches queued behind it are no-ops;
// the next kernel-level or composed call starts from a clean verdict (the reference's methods always run).
__global__ void k_clear_stop(Ctrl* ctrl) { ctrl->stop = 0; ctrl->converged = 0; ctrl->nan_flag = 0; }


// _convergence_check (LDC.py:469-501) + the bookkeeping of solve() (LDC.py:408-419).  One thread.
// The copy_new_to_old that follows is a separate launch that sees stop == 1 when converged.
__global__ void k_convergence_check(Ctrl* ctrl, double* hist, Consts K, double cu, double cv, double cp) {
    if (ctrl->stop) return;
    double rms[3];
    bool bad = false;
    for (int k = 0; k < 3; ++k) {
        rms[k] = sqrt(ctrl->residual[k] / (double)((long long)K.nx * (long long)K.ny));
        rms[k] = rms[k] / K.dt;
        ctrl->rms[k] = rms[k];
        if (isnan(rms[k]) || isinf(rms[k])) bad = true;
    }
    ctrl->iterations += 1;
    if (bad) { ctrl->nan_flag = 1; ctrl->stop = 1; return; }
    bool converged = true;
    if (rms[0] > cu) converged = false;
    if (rms[1] > cv) converged = false;
    if (rms[2] > cp) converged = false;
    if (ctrl->iterations % 100 == 0 && hist && ctrl->n_hist < ctrl->hist_cap) {
        hist[3 * ctrl->n_hist + 0] = rms[0]; hist[3 * ctrl->n_hist + 1] = rms[1]; hist[3 * ctrl->n_hist + 2] = rms[2];
        ctrl->n_hist += 1;
    }
    if (converged) { ctrl->converged = 1; ctrl->stop = 1; }
}

// The verdict above and the copy_new_to_old that follows it in ONE launch (srcfd_step): every thread evaluates the verdict
// from the three residual sums (a pure function of ctrl->residual and the criteria, the same operations as
// k_convergence_check), copies only if the iteration neither converged nor failed, and thread 0 of block 0 does the
// bookkeeping.  A block that starts after stop was raised returns at once -- the case in which nothing is copied anyway.
__global__ void k_check_and_copy(const double* __restrict__ Var, double* __restrict__ VarOld, long long n, Ctrl* ctrl,
                                 double* hist, Consts K, double cu, double cv, double cp) {
    if (ctrl->stop) return;
    double rms[3];
    bool bad = false;
    for (int k = 0; k < 3; ++k) {
        rms[k] = sqrt(ctrl->residual[k] / (double)((long long)K.nx * (long long)K.ny));
        rms[k] = rms[k] / K.dt;
        if (isnan(rms[k]) || isinf(rms[k])) bad = true;
    }
    bool converged = true;
    if (rms[0] > cu) converged = false;
    if (rms[1] > cv) converged = false;
    if (rms[2] > cp) converged = false;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) ctrl->rms[k] = rms[k];
        ctrl->iterations += 1;
        if (bad) { ctrl->nan_flag = 1; ctrl->stop = 1; }
        else {
            if (ctrl->iterations % 100 == 0 && hist && ctrl->n_hist < ctrl->hist_cap) {
                hist[3 * ctrl->n_hist + 0] = rms[0]; hist[3 * ctrl->n_hist + 1] = rms[1]; hist[3 * ctrl->n_hist + 2] = rms[2];
                ctrl->n_hist += 1;
            }
            if (converged) { ctrl->converged = 1; ctrl->stop = 1; }
        }
    }
    if (bad || converged) return;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        VarOld[t] = Var[t];
}

// Warm-start injection (LDC.py:936-938): Var[k, 1:-1, 1:-1] = field_k.T with field_k of shape (ny, nx).
template <typename T>
__global__ void k_inject_fields(double* __restrict__ Var, const T* __restrict__ fields, Consts K) {
    // tile transpose through shared memory: reads coalesced along x (nx), writes coalesced along j (ny)
    __shared__ double tile[32][33];
    const int k = blockIdx.z;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;   // x: reference i-1, y: reference j-1
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int x = x0 + threadIdx.x, y = y0 + dy;
        if (x < K.nx && y < K.ny) tile[dy][threadIdx.x] = (double)fields[((long long)k * K.ny + y) * K.nx + x];
    }
    __syncthreads();
    for (int dx = threadIdx.y; dx < 32; dx += blockDim.y) {
        const int x = x0 + dx, y = y0 + threadIdx.x;
        if (x < K.nx && y < K.ny) Var[(long long)k * K.plane + (long long)(x + 1) * K.pitch + (y + 1)] = tile[threadIdx.x][dx];
    }
}

}  // namespace srcfd
