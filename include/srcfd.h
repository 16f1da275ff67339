/*
 * srcfd.h -- C ABI of libsrcfd.so: the B200 (sm_100a) implementation of the fine-grid
 * Navier-Stokes hot path of bitseal02/SR-for-CFD.
 *
 * The reference has no FFI of its own: its "operator interface" is the set of module-level
 * numba kernels and the CFDSolver methods of PyCFD_ML_accelerated.py (LDC.py below) and
 * bfs_ml_accelerated.py (BFS.py).  Every entry point here names the reference function it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns 0 on success, non-zero on failure, and
 *     srcfd_last_error() returns a message for the calling thread's last failure;
 *   - the caller owns all host buffers; the library owns device state behind the handle;
 *   - one handle = one flow case on one CUDA device with its own stream; handles are not
 *     thread-safe, distinct handles may be driven from distinct threads;
 *   - host arrays use the reference layout: Var/VarOld (3, nx+2, ny+2), Ff (4, nx+2, ny+2),
 *     float64, C order (j fastest), k = 0:u 1:v 2:p, f = 0:E 1:N 2:W 3:S (LDC.py:342-345);
 *   - there is no CPU fallback: without a CUDA device srcfd_create fails.
 */
#ifndef SRCFD_H
#define SRCFD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCFD_ABI_VERSION 2   /* 2: srcfd_slab_* entry points, srcfd_params.sor_omega (carved out of reserved[]; layout and size unchanged) */

enum { SRCFD_SCHEME_UPWIND = 0, SRCFD_SCHEME_QUICK = 1 };           /* SolverSettings.scheme, LDC.py:95 */
enum {
    SRCFD_ORDER_GS_LEX = 0,   /* reference order: in-place lexicographic Gauss-Seidel (numba, 1 thread),  */
                              /* run as a pipelined wavefront -- results bit-identical to the reference  */
    SRCFD_ORDER_JACOBI = 1,   /* every cell from the previous iterate                                     */
    SRCFD_ORDER_RED_BLACK = 2,/* in-place two-colour sweep                                                */
    SRCFD_ORDER_RB_JACOBI = 3 /* pressure: red-black (SOR with srcfd_params.sor_omega); momentum: Jacobi.  The   */
                              /* combination north_star names ("Jacobi/red-black-SOR sweeps"); works with QUICK   */
};
enum { SRCFD_BC_DIRICHLET = 0, SRCFD_BC_NEUMANN = 1 };              /* _get_bc_arrays, LDC.py:351-375   */

typedef struct srcfd_params {
    int32_t nx, ny;              /* MeshParameters, LDC.py:69-78                                   */
    double  dx, dy, volp;        /*   dx = lx/nx, dy = ly/ny, volp = dx*dy (computed by the caller) */
    double  dt;                  /* SolverSettings.dt                                               */
    double  nu, rho;             /* FluidProperties: nu = 1/Re, LDC.py:80-86                        */
    int32_t scheme;              /* SRCFD_SCHEME_*                                                  */
    int32_t bc_types[3][4];      /* [k][left,right,top,bottom], SRCFD_BC_*                          */
    double  bc_values[3][4];
    int32_t bfs_enabled;         /* CFDSolver.case_type == 'BFS': _apply_bfs_inlet, BFS.py:524-569   */
    double  bfs_step_h, bfs_h, bfs_Ub;
    int32_t relax_enabled;       /* under_relax_field calls of BFS.py:643-659                        */
    double  relax[3];            /*   alpha_u, alpha_v, alpha_p                                      */
    double  inner_tol;           /* 1e-6, hard-coded at LDC.py:250,272,294                           */
    int32_t inner_max;           /* 1000, hard-coded at LDC.py:251,273,295                           */
    int32_t sweep_order;         /* SRCFD_ORDER_*                                                    */
    int32_t device;              /* CUDA device ordinal                                              */
    int32_t max_ctas;            /* 0 = whole GPU; >0 caps the persistent grids (ensemble members)   */
    double  sor_omega;           /* over-relaxation factor of the RED_BLACK pressure sweep (red-black SOR, north_star):  */
                                 /*   p += omega * R/ap instead of p += R/ap; 0 or 1 = plain sweep.  Other orders ignore it */
    int32_t reserved[6];
} srcfd_params;

typedef struct srcfd_handle srcfd_handle;

/* ---- life cycle --------------------------------------------------------------------------- */
int srcfd_abi_version(void);
const char *srcfd_last_error(void);
int srcfd_device_count(int *count);
/* CFDSolver.__init__ (LDC.py:333-349 / BFS.py:473-496) without _initialize_fields: device arrays are zero. */
int srcfd_create(const srcfd_params *params, srcfd_handle **out);
int srcfd_destroy(srcfd_handle *h);
/* Change everything except nx, ny, device (the reference mutates bc / case_type after construction). */
int srcfd_set_params(srcfd_handle *h, const srcfd_params *params);
int srcfd_synchronize(srcfd_handle *h);
/* The CUDA stream (cudaStream_t as an integer) the handle launches on. */
int srcfd_stream(srcfd_handle *h, uint64_t *stream);

/* ---- state transfer (any pointer may be NULL = skip) --------------------------------------- */
int srcfd_upload(srcfd_handle *h, const double *Var, const double *VarOld, const double *Ff, const double *residual);
int srcfd_download(srcfd_handle *h, double *Var, double *VarOld, double *Ff, double *residual);
/* Device addresses of the state arrays, for callers that keep inputs resident in HBM. */
int srcfd_device_ptrs(srcfd_handle *h, uint64_t *Var, uint64_t *VarOld, uint64_t *Ff);
/* Page-locked host memory for the arrays handed to srcfd_upload / srcfd_download (the reference keeps Var, VarOld, Ff
 * as numpy arrays, PyCFD_ML_accelerated.py:340-345; allocating them here lets the copies run at full PCIe rate).
 * Any host pointer is accepted by upload/download; pageable memory is simply slower. */
int srcfd_host_alloc(uint64_t bytes, void **out);
int srcfd_host_free(void *p);

/* ---- composed path ------------------------------------------------------------------------- */
/* _initialize_fields (LDC.py:377-389): optional zero fill, BC on u,v,p, copy_new_to_old, linear_interpolation. */
int srcfd_initialize_fields(srcfd_handle *h, int zero_first);
/* Warm-start injection (LDC.py:936-948): fields = (3, ny, nx) float64 or float32 on the host; writes
 * Var[k,1:-1,1:-1] = fields[k].T, then BC x3, copy_new_to_old, linear_interpolation. */
int srcfd_set_fields(srcfd_handle *h, const void *fields, int is_float32);
/* Enqueue n_outer x (_implicit_solve + _convergence_check) (LDC.py:408-419, 432-501) with no host
 * round trip; iterations after convergence / NaN are skipped on the device.  crit = {u, v, p}. */
int srcfd_step(srcfd_handle *h, int64_t n_outer, const double crit[3]);
/* Wait for enqueued work and report: iterations done since the last srcfd_reset_counters, converged
 * flag, last rms triplet (sqrt(res/(nx*ny))/dt), last inner sweep counts, cumulative inner sweeps.
 * Returns SRCFD_ERR_NAN (3) when the reference would raise ValueError("Solver failed: NaN/Inf in residuals"). */
int srcfd_status(srcfd_handle *h, int64_t *iterations, int32_t *converged, double rms[3],
                 int32_t last_sweeps[3], int64_t total_sweeps[3]);
int srcfd_reset_counters(srcfd_handle *h);
/* CFDSolver.solve loop (LDC.py:396-430) without printing/saving: runs until converged or max_iterations.
 * hist receives the rms triplets sampled when count % 100 == 0 (at most hist_cap triplets). */
int srcfd_solve(srcfd_handle *h, int64_t max_iterations, const double crit[3], int64_t *iterations,
                double *seconds, double *hist, int64_t hist_cap, int64_t *n_hist);

/* ---- kernel-level entry points, one per reference kernel, on the handle's device arrays ---- */
int srcfd_k_copy_new_to_old(srcfd_handle *h);                         /* LDC.py:110-115 */
int srcfd_k_apply_bc(srcfd_handle *h, int k);                         /* _apply_bc_wrapper: LDC.py:391-394, BFS.py:564-569 */
int srcfd_k_apply_bc_configured(srcfd_handle *h, int k);              /* apply_bc_configured alone: LDC.py:117-145 */
int srcfd_k_apply_bfs_inlet(srcfd_handle *h, int k);                  /* _apply_bfs_inlet alone: BFS.py:524-562 */
int srcfd_k_linear_interpolation(srcfd_handle *h);                    /* LDC.py:147-154 */
int srcfd_k_update_flux(srcfd_handle *h);                             /* LDC.py:239-246 */
int srcfd_k_under_relax(srcfd_handle *h, int k, double alpha);        /* BFS.py:371-375 */
int srcfd_k_correct_velocity(srcfd_handle *h, double residual_out[3]);/* LDC.py:316-328 (residual += sums) */
int srcfd_k_solve_pressure(srcfd_handle *h, int32_t *sweeps, double *last_rms);          /* LDC.py:292-314 */
/* One temporally blocked JACOBI pass of the pressure relaxation on its own (slab decomposition, srcfd/slab.py):
 * nsweeps (1..srcfd_jacobi_pass_max) sweeps of plane 2 in place, every cell of a sweep from the previous iterate
 * (LDC.py:300-310 with ORDER_JACOBI of the oracle); sums[t] = sum of R^2 of sweep t over interior rows own_row0..own_row1
 * (1-based, inclusive: a slab's halo rows are relaxed but not counted).  recompute_rhs != 0 rebuilds the right-hand
 * side from Ff first, as solve_pressure does (LDC.py:305).  commit == 0 leaves the plane untouched (the result waits in
 * the scratch plane) so that the caller can apply the break rule to the globally reduced sums first: accept with
 * srcfd_k_jacobi_commit, or repeat the pass with fewer sweeps. */
int srcfd_jacobi_pass_max(srcfd_handle *h, int *H);
int srcfd_k_jacobi_pass(srcfd_handle *h, int nsweeps, int own_row0, int own_row1, int recompute_rhs, int commit, int slot,
                        double *sums);
int srcfd_k_jacobi_commit(srcfd_handle *h);
/* Device address of the per-sweep sums, [16 slots][8]: a pass writes slot `slot` (sums == NULL above skips the host copy
 * and the stream synchronisation, so a multi-GPU caller can run a block of passes and all-reduce all their sums at once).
 * srcfd_k_jacobi_snapshot saves (restore = 0) / restores (1) the whole pressure plane, for rolling such a block back. */
int srcfd_jacobi_sums_ptr(srcfd_handle *h, uint64_t *ptr);
int srcfd_k_jacobi_snapshot(srcfd_handle *h, int restore);
int srcfd_k_solve_momentum(srcfd_handle *h, int k, int scheme, int32_t *sweeps, double *last_rms); /* LDC.py:248-290 */
/* One _implicit_solve (LDC.py:432-467 / BFS.py:622-673); residual and sweep counts via srcfd_download/srcfd_status. */
int srcfd_k_implicit_solve(srcfd_handle *h);

/* ---- slab domain decomposition of a large grid over several GPUs (BASELINE configs[3]; the reference has no such
 * path -- it is north_star's second way of sharding).  The plane is split along i into `world` contiguous slabs; a
 * slab is an ordinary handle created with sweep_order = SRCFD_ORDER_JACOBI for the LOCAL grid: nx = halo rows towards
 * rank-1 (0 for rank 0) + owned rows + halo rows towards rank+1 (0 for the last rank), dx/dy those of the whole domain,
 * host arrays (srcfd_upload/_download) hold global rows own_first-halo-1 .. own_last+halo+1.  srcfd_slab_configure tells
 * it its place; after that the srcfd_slab_* calls below replace srcfd_step / srcfd_k_solve_* for it.  They take an ARRAY
 * of handles -- the slabs this process drives: one per process in the one-process-per-GPU layout (peers mapped with
 * srcfd_slab_export + srcfd_slab_attach_ipc, the 64-byte blobs carried by any host channel), or several in one process
 * (srcfd_slab_attach_local).  Halo rows and residual sums travel as peer-memory stores from the producing kernel into
 * the consumers' mailboxes with sequence-number flags; no library collective, no host copy.  Results are those of the
 * single-domain JACOBI order bit for bit (fields, sweep counts), for any world size.
 * Not decomposed: QUICK's out-of-plane reads at an INFLOW boundary (hazard H4) see the local plane, not the far slab. */
#define SRCFD_SLAB_BLOB_BYTES 64
#define SRCFD_SLAB_MAX_WORLD 16
int srcfd_slab_configure(srcfd_handle *h, int world, int rank, int nx_global, int halo);
int srcfd_slab_export(srcfd_handle *h, void *blob, int blob_bytes);
int srcfd_slab_attach_ipc(srcfd_handle *h, int peer_rank, const void *blob);
int srcfd_slab_attach_local(srcfd_handle *h, int peer_rank, srcfd_handle *peer);
/* owned local rows (1-based, inclusive), exchanges done, bytes pushed to neighbours, blocks replayed after an overshoot */
int srcfd_slab_info(srcfd_handle *h, int32_t *own_row0, int32_t *own_row1, int64_t *exchanges, int64_t *halo_bytes,
                    int64_t *replays);
/* pressure solves run by the warp-streaming kernel / by the tile kernel (a rank falls back to tiles for a while when a
 * solve sent more than 0.01 % of its warp-steps to the IEEE division routine: denormal bands of a flow started from rest),
 * and the counts of the last solve */
int srcfd_slab_kernel_stats(srcfd_handle *h, int64_t *stream_solves, int64_t *tile_solves, int64_t *last_retries,
                            int64_t *last_warp_steps);
/* refresh the halo rows of plane k from the neighbours' owned rows */
int srcfd_slab_exchange(srcfd_handle *const *hs, int n, int k);
int srcfd_slab_solve_pressure(srcfd_handle *const *hs, int n, int32_t *sweeps, double *last_rms);            /* LDC.py:292-314 */
int srcfd_slab_solve_momentum(srcfd_handle *const *hs, int n, int k, int scheme, int32_t *sweeps, double *last_rms); /* LDC.py:248-290 */
/* n_outer x (_implicit_solve + _convergence_check + copy_new_to_old), LDC.py:408-419, 432-501; srcfd_status reports. */
int srcfd_slab_step(srcfd_handle *const *hs, int n, int64_t n_outer, const double crit[3]);

/* ---- batched small-grid solves: the coarse stage of the ML-accelerated workflow ---------------------------------
 * run_coarse_simulation (PyCFD_ML_accelerated.py:696-761 / bfs_ml_accelerated.py:893-977) builds a CFDSolver on the
 * lr_dim x lr_dim mesh (10x10), which zero-initialises (_initialize_fields) and runs solve() for up to 100 000 outer
 * iterations.  srcfd_coarse_solve_batch does that for n_cases independent cases in ONE launch: one CTA per case, the
 * whole state in shared memory, the reference sweep order (results bit-identical to CFDSolver.solve with
 * NUMBA_NUM_THREADS=1).  All cases share nx, ny and device; everything else (Re, dt, scheme, BCs, BFS inlet, relaxation)
 * is per case.  crit = n_cases x {u, v, p}.  Var / VarOld / Ff: (n_cases, 3|3|4, nx+2, ny+2) host arrays that receive
 * the final state (VarOld, Ff may be NULL); resume != 0 starts from their contents instead of zero fields.
 * hist (may be NULL): n_cases x hist_cap x 3 rms triplets sampled when count % 100 == 0.  *ms = device time of the
 * launch.  The grid must fit shared memory (srcfd_coarse_smem_bytes <= the device's opt-in limit; about 30x30). */
typedef struct srcfd_coarse_result {
    int64_t iterations;
    int32_t converged, nan_flag;     /* nan_flag: the reference raises ValueError at this iteration */
    double  rms[3];                  /* last sqrt(residual/(nx*ny))/dt triplet */
    int64_t total_sweeps[3];         /* inner sweeps executed (u, v, p) */
    int64_t n_hist;
    double  last_inner_rms[3];
    double  residual[3];             /* CFDSolver.residual after the last iteration */
    int32_t last_sweeps[3];
    int32_t reserved_;
} srcfd_coarse_result;
int srcfd_coarse_smem_bytes(int nx, int ny, uint64_t *bytes);
/* *fits = 1 when srcfd_coarse_solve_batch accepts an nx x ny grid on this device (its own size gate as a query). */
int srcfd_coarse_fits(int nx, int ny, int device, int *fits);
int srcfd_coarse_solve_batch(const srcfd_params *params, int n_cases, int64_t max_iterations, const double *crit,
                             int resume, double *Var, double *VarOld, double *Ff, srcfd_coarse_result *results,
                             double *hist, int64_t hist_cap, double *ms);

/* ---- introspection for benchmarks ---------------------------------------------------------- */
/* CUDA-event stopwatch on the handle's stream: start records an event, stop records a second one,
 * waits for it and returns the device time between them in milliseconds. */
int srcfd_timer_start(srcfd_handle *h);
int srcfd_timer_stop(srcfd_handle *h, double *ms);
/* Number of kernels this library has launched on the handle's stream since creation. */
/* SRCFD_TRACE=1 only: raw per-task timestamps (ns) of the last wavefront launch; tools/trace_gs2.py, tools/trace_gs3.py. */
int srcfd_trace_read(srcfd_handle *h, long long *out, int64_t n);
/* Debug builds: raw doubles stored behind the scratch plane. */
int srcfd_debug_read(srcfd_handle *h, double *out, int64_t n);
int srcfd_launch_count(srcfd_handle *h, int64_t *launches);
/* Device time (ms) and launches of the inner-solve kernels accumulated while timing is enabled
 * (CUDA events recorded around every inner-solve launch on the handle's stream). */
int srcfd_timing_enable(srcfd_handle *h, int enabled);
int srcfd_timing_read(srcfd_handle *h, double *pressure_ms, int64_t *pressure_launches,
                      double *momentum_ms, int64_t *momentum_launches);

/* ---- SR autoencoder inference (encoder_10 + decoder_400, sr-ae-conv.ipynb cell 162-169 / 277-287) ----
 * Replaces tf.keras load_model(...).predict at PyCFD_ML_accelerated.py:831-858.  float32, NHWC, kernels and
 * biases passed in their Keras layouts (Conv2D (kh,kw,Cin,Cout), Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out)). */
typedef struct srcfd_sr srcfd_sr;
const char *srcfd_sr_last_error(void);
int srcfd_sr_create(int device, srcfd_sr **out);
int srcfd_sr_destroy(srcfd_sr *h);
/* layers in order: conv2d, conv2d_1, dense, latent_vector */
int srcfd_sr_set_encoder(srcfd_sr *h, const float *const kernels[4], const float *const biases[4]);
/* layers in order: dense, conv2d_transpose, _1, _2, _3, _4, output_image_400 */
int srcfd_sr_set_decoder(srcfd_sr *h, const float *const kernels[7], const float *const biases[7]);
int srcfd_sr_encode(srcfd_sr *h, const float *x /* (B,10,10,1) */, int B, float *z /* (B,50) */);
int srcfd_sr_decode(srcfd_sr *h, const float *z /* (B,50) */, int B, float *out /* (B,400,400,1) */);
/* SuperResolutionAE.call (PyCFD_ML_accelerated.py:686-689) */
int srcfd_sr_predict(srcfd_sr *h, const float *x /* (B,10,10,1) */, int B, float *out /* (B,400,400,1) */);
/* ml_super_resolution's per-field pipeline (PyCFD_ML_accelerated.py:841-876 / bfs_ml_accelerated.py:1084-1134) for B
 * coarse fields in one call: [blend of the training statistics with the field's own mean/std when adaptive != 0,
 * bfs_ml_accelerated.py:1090-1100] -> standardize_with_stats -> encoder_10 -> decoder_400 -> inverse_standardize ->
 * NaN/Inf guard, all on the device.  x (B,10,10) float32; stats (B,4) float64 = {mean_lr, std_lr, mean_hr, std_hr} per
 * field; out (B,400,400) float32.  (The aspect-ratio spline resampling around it stays with the caller.) */
int srcfd_sr_super_resolve(srcfd_sr *h, const float *x, int B, const double *stats, int adaptive, double blend, float *out);
/* decoder on device-resident latents/outputs (cudaMalloc'ed by the caller); *ms = CUDA-event time of the batch */
int srcfd_sr_decode_device(srcfd_sr *h, uint64_t z_dev, int B, uint64_t out_dev, double *ms);
/* 0 = fp32 CUDA cores, 1 = bf16 tcgen05 tensor cores for the ConvT layers and the final conv (bf16 activations),
 * 2 = as 1 with the final conv on the CUDA-core tile kernel (what the tensor-core final conv is tested against),
 * 3 = split-operand tensor cores: fp32 activations, every product as three bf16 MMAs (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi),
 *     fp32 epilogues and final conv -- the accuracy of the fp32 path (1e-4 against the restatement) at tensor-core speed;
 *     THE DEFAULT of a new srcfd_sr context; its last two layers (ConvT 16 -> 8 and the final 3x3 conv, sr-ae-conv.ipynb
 *     cell 277-287) run as ONE kernel that keeps the 400x400x8 activation between them in shared memory;
 * 4 = as 3 with those two layers as separate launches (same bits; parity tests) */
int srcfd_sr_set_precision(srcfd_sr *h, int mode);
int srcfd_sr_tc_error(srcfd_sr *h, int *flag);
/* one tensor-core ConvT layer (1..4) in isolation, host fp32 in/out (operands rounded to bf16): parity tests */
int srcfd_sr_debug_convT_tc(srcfd_sr *h, int layer, const float *in, int B, float *out);
int srcfd_sr_launch_count(srcfd_sr *h, int64_t *launches);

#define SRCFD_OK 0
#define SRCFD_ERR_ARG 1
#define SRCFD_ERR_CUDA 2
#define SRCFD_ERR_NAN 3
#define SRCFD_ERR_DEADLOCK 4

#ifdef __cplusplus
}
#endif
#endif /* SRCFD_H */
