/*
 * oracle/pycfd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the PyCFD fine-grid Navier-Stokes hot path of
 * bitseal02/SR-for-CFD (the numba @njit kernels and the CFDSolver outer loop).
 * It exists only so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check the CUDA path
 * against and to time beside it.  Nothing under sr-for-cfd_b200/ may import,
 * link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py (runs where
 * /root/reference is mounted) checks every function below bit-for-bit against
 * the reference's own numba kernels run with NUMBA_NUM_THREADS=1, and
 * tests/test_golden.py checks the composed solver against the committed
 * golden vectors (outputs/bfs_Re400_centerline.dat and the coarse .h5 fields).
 *
 * Arithmetic rules that make "bit-for-bit" possible: every expression below is
 * written in the evaluation order of the Python source it restates, the file
 * is compiled with -ffp-contract=off (numba/LLVM does not fuse mul+add without
 * fastmath), true IEEE divisions are kept, and the residual sums run in
 * lexicographic (i outer, j inner) order like a 1-thread prange.
 *
 * Reference citations use the abbreviations of SURVEY.md:
 *   LDC.py = PyCFD_ML_accelerated.py,  BFS.py = bfs_ml_accelerated.py.
 *
 * Array layout (the reference's, unchanged): Var[k][i][j], k in {u,v,p},
 * shape (3, Nx+2, Ny+2), C order (j fastest); Ff[f][i][j], f in {E,N,W,S}.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_ORDER_GS_LEX 0   /* in-place lexicographic sweep = the reference run on 1 thread  */
#define ORC_ORDER_JACOBI 1   /* every cell from the previous iterate                          */
#define ORC_ORDER_RB     2   /* in-place red-black: (i+j) even first, then (i+j) odd          */
#define ORC_ORDER_GS_OMP 3   /* in-place, rows split over OpenMP threads (racy seams, like    */
                             /* numba prange with >1 thread) -- CPU-baseline timing only      */

#define ORC_SCHEME_UPWIND 0
#define ORC_SCHEME_QUICK  1

#define ORC_OP_PRESSURE 0
#define ORC_OP_UPWIND   1
#define ORC_OP_QUICK    2

typedef struct {
    int Nx, Ny;
    size_t sI;   /* stride of i  = Ny+2          */
    size_t P;    /* plane stride = (Nx+2)(Ny+2)  */
} grid_t;

static grid_t mk_grid(int Nx, int Ny) {
    grid_t g; g.Nx = Nx; g.Ny = Ny; g.sI = (size_t)Ny + 2; g.P = ((size_t)Nx + 2) * ((size_t)Ny + 2);
    return g;
}

/* ---- LDC.py:110-115 copy_new_to_old ------------------------------------ */
void orc_copy_new_to_old(const double *Var, double *VarOld, int nVar, int Nx, int Ny) {
    grid_t g = mk_grid(Nx, Ny);
    memcpy(VarOld, Var, sizeof(double) * g.P * (size_t)nVar);
}

/* ---- LDC.py:117-145 apply_bc_configured --------------------------------
 * sides ordered [left(i=0), right(i=Nx+1), top(j=Ny+1), bottom(j=0)];
 * type 0 = Dirichlet (ghost = 2*value - interior), else Neumann (ghost = interior).
 * Corners are never written. */
void orc_apply_bc_configured(double *Var, int k, int Nx, int Ny,
                             const int32_t *bc_types, const double *bc_values) {
    grid_t g = mk_grid(Nx, Ny);
    double *V = Var + (size_t)k * g.P;
    for (int j = 1; j <= Ny; ++j) {
        if (bc_types[0] == 0) V[0 * g.sI + j] = 2 * bc_values[0] - V[1 * g.sI + j];
        else                  V[0 * g.sI + j] = V[1 * g.sI + j];
        if (bc_types[1] == 0) V[(size_t)(Nx + 1) * g.sI + j] = 2 * bc_values[1] - V[(size_t)Nx * g.sI + j];
        else                  V[(size_t)(Nx + 1) * g.sI + j] = V[(size_t)Nx * g.sI + j];
    }
    for (int i = 1; i <= Nx; ++i) {
        if (bc_types[2] == 0) V[(size_t)i * g.sI + Ny + 1] = 2 * bc_values[2] - V[(size_t)i * g.sI + Ny];
        else                  V[(size_t)i * g.sI + Ny + 1] = V[(size_t)i * g.sI + Ny];
        if (bc_types[3] == 0) V[(size_t)i * g.sI + 0] = 2 * bc_values[3] - V[(size_t)i * g.sI + 1];
        else                  V[(size_t)i * g.sI + 0] = V[(size_t)i * g.sI + 1];
    }
}

/* ---- BFS.py:524-562 CFDSolver._apply_bfs_inlet -------------------------
 * Left ghost column override for k in {0,1}: below the step a no-slip wall,
 * above it a parabolic u profile and v = 0. */
void orc_apply_bfs_inlet(double *Var, int k, int Nx, int Ny, double dy,
                         double step_h, double h, double Ub) {
    if (k != 0 && k != 1) return;
    grid_t g = mk_grid(Nx, Ny);
    double *U = Var, *Vv = Var + g.P, *Vk = Var + (size_t)k * g.P;
    for (int j = 1; j <= Ny; ++j) {
        double y = (j - 0.5) * dy;
        if (y < step_h) {
            Vk[j] = -Vk[g.sI + j];
        } else if (k == 1) {
            Vv[j] = -Vv[g.sI + j];
        } else {
            double yprime = y - step_h;
            if (yprime < 0.0) yprime = 0.0;
            if (yprime > h) yprime = h;
            double u_in = 6.0 * Ub * (yprime / h) * (1.0 - (yprime / h));
            U[j] = 2.0 * u_in - U[g.sI + j];
            Vv[j] = -Vv[g.sI + j];
        }
    }
}

/* ---- LDC.py:147-154 linear_interpolation -------------------------------- */
void orc_linear_interpolation(const double *Var, double *Ff, int Nx, int Ny, double dx, double dy) {
    grid_t g = mk_grid(Nx, Ny);
    const double *U = Var, *V = Var + g.P;
    for (int i = 1; i <= Nx; ++i)
        for (int j = 1; j <= Ny; ++j) {
            size_t c = (size_t)i * g.sI + j;
            Ff[0 * g.P + c] = (U[c] + U[c + g.sI]) * dy * 0.5;
            Ff[1 * g.P + c] = (V[c] + V[c + 1]) * dx * 0.5;
            Ff[2 * g.P + c] = -(U[c] + U[c - g.sI]) * dy * 0.5;
            Ff[3 * g.P + c] = -(V[c] + V[c - 1]) * dx * 0.5;
        }
}

/* ---- LDC.py:239-246 update_flux ----------------------------------------- */
void orc_update_flux(const double *Var, double *Ff, double dt, double rho, int Nx, int Ny,
                     double dx, double dy) {
    grid_t g = mk_grid(Nx, Ny);
    const double *Pp = Var + 2 * g.P;
    for (int i = 1; i <= Nx; ++i)
        for (int j = 1; j <= Ny; ++j) {
            size_t c = (size_t)i * g.sI + j;
            Ff[0 * g.P + c] += -dt / rho * (Pp[c + g.sI] - Pp[c]) * dy / dx;
            Ff[1 * g.P + c] += -dt / rho * (Pp[c + 1] - Pp[c]) * dx / dy;
            Ff[2 * g.P + c] += -dt / rho * (Pp[c - g.sI] - Pp[c]) * dy / dx;
            Ff[3 * g.P + c] += -dt / rho * (Pp[c - 1] - Pp[c]) * dx / dy;
        }
}

/* ---- BFS.py:371-375 under_relax_field ----------------------------------- */
void orc_under_relax_field(double *Var, const double *VarOld, int k, int Nx, int Ny, double alpha) {
    grid_t g = mk_grid(Nx, Ny);
    double *V = Var + (size_t)k * g.P;
    const double *O = VarOld + (size_t)k * g.P;
    for (int i = 1; i <= Nx; ++i)
        for (int j = 1; j <= Ny; ++j) {
            size_t c = (size_t)i * g.sI + j;
            V[c] = O[c] + alpha * (V[c] - O[c]);
        }
}

/* ---- LDC.py:232-237 diffusive_flux -------------------------------------- */
static inline double diff_flux(const double *V, size_t c, size_t sI, double dx, double dy, double volp) {
    return volp * ((V[c + sI] - 2.0 * V[c] + V[c - sI]) / (dx * dx) +
                   (V[c + 1] - 2.0 * V[c] + V[c - 1]) / (dy * dy));
}
static inline double diff_ap(double dx, double dy, double volp) {
    return -volp * (2.0 / (dx * dx) + 2.0 / (dy * dy));
}

/* ---- LDC.py:156-188 simple_upwind --------------------------------------- */
static inline void upwind_cell(const double *V, const double *Ff, size_t P, size_t c, size_t sI,
                               double volp, double *Fc, double *ap_c) {
    double fE = Ff[0 * P + c], fN = Ff[1 * P + c], fW = Ff[2 * P + c], fS = Ff[3 * P + c];
    double ue, uw, un, us, sum_flux = 0.0;
    if (fE >= 0) { ue = V[c]; sum_flux += fE; } else ue = V[c + sI];
    if (fW >= 0) { uw = V[c]; sum_flux += fW; } else uw = V[c - sI];
    if (fN >= 0) { un = V[c]; sum_flux += fN; } else un = V[c + 1];
    if (fS >= 0) { us = V[c]; sum_flux += fS; } else us = V[c - 1];
    *Fc = ue * fE + uw * fW + un * fN + us * fS;
    *ap_c = sum_flux * volp;
}

/* ---- LDC.py:190-230 quick_scheme ----------------------------------------
 * Second-neighbour reads that leave the (Nx+2, Ny+2) plane follow what the
 * numba code does on the flat buffer (SURVEY.md hazard H4): a negative index
 * wraps to the other end of that axis, an index one past the end runs on into
 * the next row / next plane.  qrd() takes the ABSOLUTE plane base (k*P). */
static inline double qrd(const double *Var, size_t kbase, int Nx, int Ny, size_t sI, int i, int j) {
    if (i < 0) i += Nx + 2;
    if (j < 0) j += Ny + 2;
    return Var[kbase + (size_t)i * sI + (size_t)j];
}
static inline void quick_cell(const double *Var, size_t kbase, const double *Ff, size_t P, int Nx, int Ny,
                              size_t sI, int i, int j, double volp, double *Fc, double *ap_c) {
    size_t c = (size_t)i * sI + j;
    const double *V = Var + kbase;
    double fE = Ff[0 * P + c], fN = Ff[1 * P + c], fW = Ff[2 * P + c], fS = Ff[3 * P + c];
    double ue, uw, un, us, sum_flux = 0.0;
    if (fE >= 0) { ue = 0.75 * V[c] + 0.375 * V[c + sI] - 0.125 * V[c - sI]; sum_flux += 0.75 * fE; }
    else { ue = 0.75 * V[c + sI] + 0.375 * V[c] - 0.125 * qrd(Var, kbase, Nx, Ny, sI, i + 2, j); sum_flux += 0.375 * fE; }
    if (fW >= 0) { uw = 0.75 * V[c] + 0.375 * V[c - sI] - 0.125 * V[c + sI]; sum_flux += 0.75 * fW; }
    else { uw = 0.75 * V[c - sI] + 0.375 * V[c] - 0.125 * qrd(Var, kbase, Nx, Ny, sI, i - 2, j); sum_flux += 0.375 * fW; }
    if (fN >= 0) { un = 0.75 * V[c] + 0.375 * V[c + 1] - 0.125 * V[c - 1]; sum_flux += 0.75 * fN; }
    else { un = 0.75 * V[c + 1] + 0.375 * V[c] - 0.125 * qrd(Var, kbase, Nx, Ny, sI, i, j + 2); sum_flux += 0.375 * fN; }
    if (fS >= 0) { us = 0.75 * V[c] + 0.375 * V[c - 1] - 0.125 * V[c + 1]; sum_flux += 0.75 * fS; }
    else { us = 0.75 * V[c - 1] + 0.375 * V[c] - 0.125 * qrd(Var, kbase, Nx, Ny, sI, i, j - 2); sum_flux += 0.375 * fS; }
    *Fc = ue * fE + uw * fW + un * fN + us * fS;
    *ap_c = sum_flux * volp;
}

/* Successive over-relaxation factor of the RED-BLACK pressure sweep (north_star's "red-black-SOR"; not in the reference,
 * whose update is omega = 1): p += omega * R/ap with R and ap exactly as in solve_pressure (LDC.py:300-310).
 * Process-wide because the oracle is test infrastructure; 1.0 = the plain sweep. */
static double g_sor_omega = 1.0;
void orc_set_sor_omega(double w) { g_sor_omega = (w > 0.0) ? w : 1.0; }

/* One cell of one relaxation sweep.  `src` is the array the stencil reads
 * (== Var for the in-place orders, the previous iterate for Jacobi); the new
 * value is returned, R is handed back for the residual norm.
 *   pressure : LDC.py:300-310     momentum : LDC.py:255-264 / 277-286 */
static inline double relax_cell(int op, const double *src, const double *VarOld, const double *Ff,
                                int k, int Nx, int Ny, size_t sI, size_t P, int i, int j,
                                double dx, double dy, double dt, double nu_or_rho, double volp,
                                double *Rout) {
    size_t kbase = (size_t)k * P, c = (size_t)i * sI + j;
    const double *V = src + kbase;
    double Fd = diff_flux(V, c, sI, dx, dy, volp);
    double ap_d = diff_ap(dx, dy, volp);
    double R, ap;
    if (op == ORC_OP_PRESSURE) {
        double rho = nu_or_rho;
        double RHS = rho / dt * (Ff[0 * P + c] + Ff[1 * P + c] + Ff[2 * P + c] + Ff[3 * P + c]);
        R = RHS - Fd;
        ap = ap_d;
    } else {
        double nu = nu_or_rho, Fc, ap_c;
        if (op == ORC_OP_UPWIND) upwind_cell(V, Ff, P, c, sI, volp, &Fc, &ap_c);
        else quick_cell(src, kbase, Ff, P, Nx, Ny, sI, i, j, volp, &Fc, &ap_c);
        R = -(volp / dt * (V[c] - VarOld[kbase + c]) + Fc + (-nu) * Fd);
        ap = volp / dt + ap_c + (-nu) * ap_d;
    }
    *Rout = R;
    return V[c] + R / ap;
}

/* Generic inner solve: <= max_iter sweeps, stop after the first sweep whose
 * rms = sqrt(sum R^2 / (Nx*Ny)) is < tolerance (LDC.py:253-268, 275-290,
 * 298-314).  Returns the number of sweeps executed; *last_rms gets the rms of
 * the last one.  rms_hist (may be NULL) receives one value per sweep. */
static int inner_solve(int op, double *Var, const double *VarOld, const double *Ff, int k,
                       int Nx, int Ny, double dx, double dy, double dt, double nu_or_rho, double volp,
                       int order, double tolerance, int max_iter, double *last_rms, double *rms_hist) {
    grid_t g = mk_grid(Nx, Ny);
    size_t nplanes = 3;
    double *prev = NULL;
    if (order == ORC_ORDER_JACOBI) prev = (double *)malloc(sizeof(double) * g.P * nplanes);
    int sweeps = 0;
    double rms = 0.0;
    for (int it = 0; it < max_iter; ++it) {
        rms = 0.0;
        if (order == ORC_ORDER_GS_LEX) {
            for (int i = 1; i <= Nx; ++i)
                for (int j = 1; j <= Ny; ++j) {
                    double R;
                    double nv = relax_cell(op, Var, VarOld, Ff, k, Nx, Ny, g.sI, g.P, i, j, dx, dy, dt, nu_or_rho, volp, &R);
                    Var[(size_t)k * g.P + (size_t)i * g.sI + j] = nv;
                    rms += R * R;
                }
        } else if (order == ORC_ORDER_GS_OMP) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : rms)
#endif
            for (int i = 1; i <= Nx; ++i)
                for (int j = 1; j <= Ny; ++j) {
                    double R;
                    double nv = relax_cell(op, Var, VarOld, Ff, k, Nx, Ny, g.sI, g.P, i, j, dx, dy, dt, nu_or_rho, volp, &R);
                    Var[(size_t)k * g.P + (size_t)i * g.sI + j] = nv;
                    rms += R * R;
                }
        } else if (order == ORC_ORDER_JACOBI) {
            memcpy(prev, Var, sizeof(double) * g.P * nplanes);
            for (int i = 1; i <= Nx; ++i)
                for (int j = 1; j <= Ny; ++j) {
                    double R;
                    double nv = relax_cell(op, prev, VarOld, Ff, k, Nx, Ny, g.sI, g.P, i, j, dx, dy, dt, nu_or_rho, volp, &R);
                    Var[(size_t)k * g.P + (size_t)i * g.sI + j] = nv;
                    rms += R * R;
                }
        } else { /* red-black, in place */
            for (int colour = 0; colour < 2; ++colour)
                for (int i = 1; i <= Nx; ++i)
                    for (int j = 1; j <= Ny; ++j) {
                        if (((i + j) & 1) != colour) continue;
                        double R;
                        double nv = relax_cell(op, Var, VarOld, Ff, k, Nx, Ny, g.sI, g.P, i, j, dx, dy, dt, nu_or_rho, volp, &R);
                        if (op == ORC_OP_PRESSURE && g_sor_omega != 1.0)
                            nv = Var[(size_t)k * g.P + (size_t)i * g.sI + j] + g_sor_omega * (R / diff_ap(dx, dy, volp));
                        Var[(size_t)k * g.P + (size_t)i * g.sI + j] = nv;
                        rms += R * R;
                    }
        }
        rms = sqrt(rms / (double)((int64_t)Nx * (int64_t)Ny));
        if (rms_hist) rms_hist[it] = rms;
        sweeps = it + 1;
        if (rms < tolerance) break;
    }
    if (prev) free(prev);
    if (last_rms) *last_rms = rms;
    return sweeps;
}

/* ---- LDC.py:292-314 solve_pressure -------------------------------------- */
int orc_solve_pressure(double *Var, const double *Ff, int Nx, int Ny, double dx, double dy, double dt,
                       double rho, double volp, int order, double tolerance, int max_iter,
                       double *last_rms, double *rms_hist) {
    return inner_solve(ORC_OP_PRESSURE, Var, NULL, Ff, 2, Nx, Ny, dx, dy, dt, rho, volp, order,
                       tolerance, max_iter, last_rms, rms_hist);
}
/* ---- LDC.py:270-290 solve_momentum_upwind ------------------------------- */
int orc_solve_momentum_upwind(double *Var, const double *VarOld, const double *Ff, int k, int Nx, int Ny,
                              double dx, double dy, double dt, double nu, double volp, int order,
                              double tolerance, int max_iter, double *last_rms, double *rms_hist) {
    return inner_solve(ORC_OP_UPWIND, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order,
                       tolerance, max_iter, last_rms, rms_hist);
}
/* ---- LDC.py:248-268 solve_momentum_quick -------------------------------- */
int orc_solve_momentum_quick(double *Var, const double *VarOld, const double *Ff, int k, int Nx, int Ny,
                             double dx, double dy, double dt, double nu, double volp, int order,
                             double tolerance, int max_iter, double *last_rms, double *rms_hist) {
    return inner_solve(ORC_OP_QUICK, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order,
                       tolerance, max_iter, last_rms, rms_hist);
}

/* ---- LDC.py:316-328 / BFS.py:445-464 correct_velocity --------------------
 * residual[0..2] += sum (Var_k - VarOld_k)^2 in lexicographic order. */
void orc_correct_velocity(double *Var, const double *VarOld, double dt, double rho, int Nx, int Ny,
                          double dx, double dy, double *residual) {
    grid_t g = mk_grid(Nx, Ny);
    double *U = Var, *V = Var + g.P;
    const double *Pp = Var + 2 * g.P;
    const double *UO = VarOld, *VO = VarOld + g.P, *PO = VarOld + 2 * g.P;
    for (int i = 1; i <= Nx; ++i)
        for (int j = 1; j <= Ny; ++j) {
            size_t c = (size_t)i * g.sI + j;
            U[c] = U[c] - dt / rho * (Pp[c + g.sI] - Pp[c - g.sI]) / (2 * dx);
            V[c] = V[c] - dt / rho * (Pp[c + 1] - Pp[c - 1]) / (2 * dy);
            double du = U[c] - UO[c], dv = V[c] - VO[c], dp = Pp[c] - PO[c];
            residual[0] += du * du;
            residual[1] += dv * dv;
            residual[2] += dp * dp;
        }
}

/* ======================================================================= *
 *  Composed solver: CFDSolver (LDC.py:331-501, BFS.py:471-706)            *
 * ======================================================================= */
typedef struct {
    int32_t Nx, Ny;
    double dx, dy, volp, dt, nu, rho;
    int32_t scheme;              /* ORC_SCHEME_*                                          */
    int32_t bc_types[3][4];      /* [k][left,right,top,bottom]                            */
    double  bc_values[3][4];
    int32_t bfs;                 /* apply _apply_bfs_inlet after every BC (BFS.py:564-569)*/
    double  step_h, h, Ub;
    int32_t use_relax;           /* BFS.py:643-659 under-relaxation calls                 */
    double  alpha[3];
    int32_t order;               /* ORC_ORDER_*                                           */
    double  inner_tol;           /* 1e-6  (LDC.py:250)                                    */
    int32_t inner_max;           /* 1000  (LDC.py:251)                                    */
} orc_params;

void orc_apply_bc_wrapper(const orc_params *p, double *Var, int k) {
    orc_apply_bc_configured(Var, k, p->Nx, p->Ny, p->bc_types[k], p->bc_values[k]);
    if (p->bfs) orc_apply_bfs_inlet(Var, k, p->Nx, p->Ny, p->dy, p->step_h, p->h, p->Ub);
}

/* LDC.py:377-389 _initialize_fields (also the tail of the warm-start
 * injection, LDC.py:942-948, when zero_first == 0). */
void orc_initialize_fields(const orc_params *p, double *Var, double *VarOld, double *Ff, int zero_first) {
    grid_t g = mk_grid(p->Nx, p->Ny);
    if (zero_first) {
        memset(Var, 0, sizeof(double) * 3 * g.P);
        memset(VarOld, 0, sizeof(double) * 3 * g.P);
        memset(Ff, 0, sizeof(double) * 4 * g.P);
    }
    for (int k = 0; k < 3; ++k) orc_apply_bc_wrapper(p, Var, k);
    orc_copy_new_to_old(Var, VarOld, 3, p->Nx, p->Ny);
    orc_linear_interpolation(Var, Ff, p->Nx, p->Ny, p->dx, p->dy);
}

/* LDC.py:432-467 / BFS.py:622-673 _implicit_solve.  sweeps[3] receives the
 * inner sweep counts (u, v, p). */
void orc_implicit_solve(const orc_params *p, double *Var, double *VarOld, double *Ff,
                        double *residual, int32_t *sweeps) {
    /* order 4 (RB_JACOBI): momentum in Jacobi order, pressure red-black (with the SOR factor when one is set) */
    const int mo = p->order == 4 ? ORC_ORDER_JACOBI : p->order, po = p->order == 4 ? 2 : p->order;
    residual[0] = residual[1] = residual[2] = 0.0;
    for (int k = 0; k < 2; ++k) {
        int n;
        if (p->scheme == ORC_SCHEME_QUICK)
            n = orc_solve_momentum_quick(Var, VarOld, Ff, k, p->Nx, p->Ny, p->dx, p->dy, p->dt, p->nu,
                                         p->volp, mo, p->inner_tol, p->inner_max, NULL, NULL);
        else
            n = orc_solve_momentum_upwind(Var, VarOld, Ff, k, p->Nx, p->Ny, p->dx, p->dy, p->dt, p->nu,
                                          p->volp, mo, p->inner_tol, p->inner_max, NULL, NULL);
        if (sweeps) sweeps[k] = n;
        if (p->use_relax) orc_under_relax_field(Var, VarOld, k, p->Nx, p->Ny, p->alpha[k]);
        orc_apply_bc_wrapper(p, Var, k);
    }
    orc_linear_interpolation(Var, Ff, p->Nx, p->Ny, p->dx, p->dy);
    int n = orc_solve_pressure(Var, Ff, p->Nx, p->Ny, p->dx, p->dy, p->dt, p->rho, p->volp, po,
                               p->inner_tol, p->inner_max, NULL, NULL);
    if (sweeps) sweeps[2] = n;
    if (p->use_relax) orc_under_relax_field(Var, VarOld, 2, p->Nx, p->Ny, p->alpha[2]);
    orc_apply_bc_wrapper(p, Var, 2);
    orc_correct_velocity(Var, VarOld, p->dt, p->rho, p->Nx, p->Ny, p->dx, p->dy, residual);
    orc_apply_bc_wrapper(p, Var, 0);
    orc_apply_bc_wrapper(p, Var, 1);
    orc_update_flux(Var, Ff, p->dt, p->rho, p->Nx, p->Ny, p->dx, p->dy);
}

/* LDC.py:469-501 _convergence_check.  Returns 1 converged, 0 not, -1 NaN/Inf
 * (the reference raises ValueError there). */
int orc_convergence_check(const orc_params *p, double *Var, double *VarOld, const double *residual,
                          const double *crit, double *rms) {
    for (int k = 0; k < 3; ++k) {
        rms[k] = sqrt(residual[k] / (double)((int64_t)p->Nx * (int64_t)p->Ny));
        rms[k] = rms[k] / p->dt;
    }
    for (int k = 0; k < 3; ++k)
        if (isnan(rms[k]) || isinf(rms[k])) return -1;
    int converged = 1;
    if (rms[0] > crit[0]) converged = 0;
    if (rms[1] > crit[1]) converged = 0;
    if (rms[2] > crit[2]) converged = 0;
    if (!converged) orc_copy_new_to_old(Var, VarOld, 3, p->Nx, p->Ny);
    return converged;
}

/* LDC.py:396-430 solve (loop only; no printing, no saving).  hist (may be
 * NULL) receives rms triplets sampled when count % 100 == 0, at most hist_cap
 * of them; *n_hist gets how many.  total_sweeps[3] accumulates inner sweeps.
 * Returns the iteration count, or -count if NaN/Inf appeared at `count`. */
int64_t orc_solve(const orc_params *p, double *Var, double *VarOld, double *Ff, int64_t max_iterations,
                  const double *crit, double *last_rms, double *hist, int64_t hist_cap, int64_t *n_hist,
                  int64_t *total_sweeps) {
    int64_t count = 0, nh = 0;
    int converged = 0;
    double residual[3], rms[3] = {0, 0, 0};
    int32_t sw[3];
    if (total_sweeps) total_sweeps[0] = total_sweeps[1] = total_sweeps[2] = 0;
    while (!converged && count < max_iterations) {
        count += 1;
        orc_implicit_solve(p, Var, VarOld, Ff, residual, sw);
        if (total_sweeps) for (int k = 0; k < 3; ++k) total_sweeps[k] += sw[k];
        int c = orc_convergence_check(p, Var, VarOld, residual, crit, rms);
        if (c < 0) { if (last_rms) memcpy(last_rms, rms, sizeof rms); if (n_hist) *n_hist = nh; return -count; }
        converged = c;
        if (count % 100 == 0 && hist && nh < hist_cap) { memcpy(hist + 3 * nh, rms, sizeof rms); nh++; }
    }
    if (last_rms) memcpy(last_rms, rms, sizeof rms);
    if (n_hist) *n_hist = nh;
    return count;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
