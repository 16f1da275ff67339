// srcfd.cu -- C ABI (include/srcfd.h) over the sm_100a kernels.  Single translation unit.
// Build: see sr-for-cfd_b200/csrc/Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/srcfd.h"
#include "common.cuh"
#include "cell_ops.cuh"
#include "tail_kernels.cuh"
#include "inner_solvers.cuh"
#include "inner_gs2.cuh"
#include "inner_gs3.cuh"
#include "jacobi_tb.cuh"
#include "slab.cuh"
#include "coarse_batch.cuh"

using namespace srcfd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SRCFD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define CKH(h)                                                                                     \
    do {                                                                                           \
        if (!(h)) return fail(SRCFD_ERR_ARG, "null handle");                                       \
        CK(cudaSetDevice((h)->dev));                                                               \
    } while (0)

struct EvPair { cudaEvent_t a, b; int kind; };
struct SlabState;               // slab_api.inl

struct srcfd_handle {
    srcfd_params p;
    Consts K;
    BcSpec bc;
    int dev = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    double *Var = nullptr, *VarOld = nullptr, *Ff = nullptr, *rhs = nullptr, *scratch = nullptr;
    bool ff_paired = false;     // Ff was produced by k_linear_interpolation (+ k_update_flux): W/S planes = negated E/N of the neighbour
    double *partials = nullptr, *res_partials = nullptr, *hist = nullptr;
    void* staging = nullptr;
    size_t staging_bytes = 0;
    int* prog = nullptr;
    Ctrl* ctrl = nullptr;       // device
    Ctrl* ctrl_host = nullptr;  // pinned mirror
    long long hist_cap = 0;
    int nbands = 1, band_rows = 1, wf_threads = 32;
    int grid_gs[3] = {0, 0, 0}, grid_sync = 0, tail_blocks = 0;
    size_t wf_smem = 0;
    size_t n_partials = 0;
    int64_t launches = 0;
    bool timing = false;
    std::vector<EvPair> ev_pending, ev_free;
    double t_ms[2] = {0, 0};
    int64_t t_n[2] = {0, 0};
    int spin_limit = 4000000;
    int gs_impl = 2;            // 2: K-sweep lock-step wavefront (inner_gs2.cuh), 1: one sweep per CTA
    Gs2Plan plan2[3];
    int grid_gs2[3] = {0, 0, 0};
    double* halo = nullptr;
    // second problem of a paired (u,v) momentum launch
    double *scratch2 = nullptr, *partials2 = nullptr, *halo2 = nullptr;
    double *sweeps1 = nullptr, *sweeps2 = nullptr;   // per-sweep result planes of a run's last group (two problems)
    int* prog2 = nullptr;
    bool pair_momentum = true;   // SRCFD_PAIR=0 disables
    bool ghosts_fresh = false;   // v ghost column known to equal -v(1,j): set by the BC passes, cleared by uploads
    // third-generation pressure solve (inner_gs3.cuh): full-height groups, diagonal streams
    bool gs3 = false;            // usable for this grid (SRCFD_GS3=0 disables)
    int gs3_K = 3, gs3_RP = 0, gs3_ND = 0, gs3_nbuf = 4, gs3_grid = 0;
    int gs3_stride = 512, gs3_kmax = 4;   // row slots per diagonal and largest group of the kernel instance in use
    size_t gs3_smem = 0;
    const void* gs3_fn = nullptr;
    uint4* gs3_ll = nullptr;
    double* gs3_rhsS = nullptr;
    unsigned long long* gs3_epoch = nullptr;
    // temporally blocked Jacobi pressure solve (jacobi_tb.cuh); SRCFD_JTB=0 falls back to one sweep per grid barrier
    int jtb_H = 0, jtb_grid = 0;
    size_t jtb_smem = 0;
    const void* jtb_fn = nullptr;
    double* jtb_partials = nullptr;
    double* jtb_sums = nullptr;         // [16 slots][8] per-sweep residual sums of single passes (device)
    unsigned* jtb_ticket = nullptr;     // last-CTA-done counter of the single-pass kernel
    bool jtb_ghosts_valid = false;      // boundary cells of the scratch plane match the pressure plane
    const void* jtb_pass_fn = nullptr;
    int jtb2_cols_env = 0;              // SRCFD_JTB2_COLS: 1 | 2 columns per lane of the streaming kernel (experiments; 0 = by size)
    int jtb2_rb_env = 0;                // SRCFD_JTB2_RB: rows per chunk of the streaming kernel (experiments)
    int jtb_impl = 2;                   // SRCFD_JTB_IMPL: 2 = warp-streaming kernel (k_jtb2_pass), 1 = shared-memory tile kernel
    double* jtb2_partials = nullptr;    // [4][units] per-(sweep, warp-unit) residual sums
    size_t jtb2_units_cap = 0;
    long long* trace = nullptr;   // SRCFD_TRACE=1: per-task timestamps of the last K-sweep launch
    size_t trace_n = 0;
    cudaEvent_t tm_a = nullptr, tm_b = nullptr;   // srcfd_timer_start/stop
    int inner_cap = 0;          // capacity of the per-sweep buffers (inner_max at creation)
    int guess_bias = 0;
    // environment knobs, read once at creation (getenv is neither cheap nor safe against a concurrent setenv)
    int gs3_k1max = 128, gs3_skip_idle = 1, gs3_pretouch = 1;
    bool maybe_stopped = false; // srcfd_step/solve may have left Ctrl.stop set on the device
    SlabState* slab = nullptr;  // non-null: this handle is one slab of a decomposed grid (srcfd_slab_configure)
};
static void slab_release(srcfd_handle* h);

// Derived constants, each with the reference's own expression (host compiled with -ffp-contract=off).
static Consts make_consts(const srcfd_params& p) {
    Consts K;
    K.nx = p.nx; K.ny = p.ny; K.pitch = p.ny + 2; K.plane = (long long)(p.nx + 2) * (p.ny + 2);
    K.dx = p.dx; K.dy = p.dy; K.volp = p.volp; K.dt = p.dt; K.nu = p.nu; K.rho = p.rho;
    K.dx2 = p.dx * p.dx; K.dy2 = p.dy * p.dy;
    K.ap_d = -p.volp * (2.0 / (p.dx * p.dx) + 2.0 / (p.dy * p.dy));
    K.volp_dt = p.volp / p.dt;
    K.rho_dt = p.rho / p.dt;
    K.neg_nu = -p.nu;
    K.neg_nu_ap_d = (-p.nu) * K.ap_d;
    K.dt_rho = p.dt / p.rho;
    K.mdt_rho = -p.dt / p.rho;
    K.two_dx = 2 * p.dx; K.two_dy = 2 * p.dy;
    return K;
}
static BcSpec make_bc(const srcfd_params& p) {
    BcSpec b;
    for (int k = 0; k < 3; ++k)
        for (int s = 0; s < 4; ++s) { b.types[k][s] = p.bc_types[k][s]; b.values[k][s] = p.bc_values[k][s]; }
    b.bfs = p.bfs_enabled; b.step_h = p.bfs_step_h; b.h = p.bfs_h; b.Ub = p.bfs_Ub;
    b.skip_lo = b.skip_hi = 0;
    return b;
}

static int check_params(const srcfd_params* p) {
    if (!p) return fail(SRCFD_ERR_ARG, "null params");
    if (p->nx < 1 || p->ny < 1) return fail(SRCFD_ERR_ARG, "nx, ny must be >= 1");
    if (p->scheme != SRCFD_SCHEME_UPWIND && p->scheme != SRCFD_SCHEME_QUICK) return fail(SRCFD_ERR_ARG, "bad scheme");
    if (p->sweep_order < 0 || p->sweep_order > 3) return fail(SRCFD_ERR_ARG, "bad sweep_order");
    if (p->inner_max < 1) return fail(SRCFD_ERR_ARG, "inner_max must be >= 1");
    if (!(p->sor_omega >= 0.0 && p->sor_omega < 2.0)) return fail(SRCFD_ERR_ARG, "sor_omega must be in [0, 2) (0 = off)");
    if (p->sweep_order == SRCFD_ORDER_RED_BLACK && p->scheme == SRCFD_SCHEME_QUICK)
        return fail(SRCFD_ERR_ARG, "red-black order is undefined for the 9-point QUICK stencil (same-colour second neighbours)");
    return SRCFD_OK;
}

template <int OP> static const void* gs_kernel() { return (const void*)k_solve_gs<OP>; }
template <int OP, int ORDER> static const void* sync_kernel() { return (const void*)k_solve_sync<OP, ORDER>; }

static const void* pick_sync(int op, int order) {
    if (order == SRCFD_ORDER_JACOBI) {
        if (op == OP_PRESSURE) return sync_kernel<OP_PRESSURE, 1>();
        if (op == OP_UPWIND) return sync_kernel<OP_UPWIND, 1>();
        return sync_kernel<OP_QUICK, 1>();
    }
    if (op == OP_PRESSURE) return sync_kernel<OP_PRESSURE, 2>();
    return sync_kernel<OP_UPWIND, 2>();
}
static const void* pick_gs(int op) {
    if (op == OP_PRESSURE) return gs_kernel<OP_PRESSURE>();
    if (op == OP_UPWIND) return gs_kernel<OP_UPWIND>();
    return gs_kernel<OP_QUICK>();
}

template <int OP> static const void* gs2_kernel() { return (const void*)k_solve_gs2<OP>; }
static const void* pick_gs2(int op) {
    if (op == OP_PRESSURE) return gs2_kernel<OP_PRESSURE>();
    if (op == OP_UPWIND) return gs2_kernel<OP_UPWIND>();
    return gs2_kernel<OP_QUICK>();
}
template <int OP> static void shape2(int& K, int& NB, int& MAXT, int& NAUX, int& AD) {
    K = Wf2Shape<OP>::K; NB = Wf2Shape<OP>::NB; MAXT = Wf2Shape<OP>::MAXT; NAUX = Wf2Shape<OP>::NAUX; AD = Wf2Shape<OP>::AD;
}

// Row bands for the K-sweep wavefront: (K+1) thread groups of RS slots each (+ service warps) must fit the
// CTA, and every band needs >= NB*K rows so that the redundant rows of the band above stay inside it.
// The opt-in dynamic shared-memory limit is a per-(device, kernel) attribute shared by every handle of the process:
// only ever raise it, or a handle created later for a smaller grid would break the launches of an earlier one.
static int raise_smem_limit(int dev, const void* fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> cur;
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[std::make_pair(dev, fn)];
    if (bytes > c) {
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        c = bytes;
    }
    return SRCFD_OK;
}

static int plan_gs2(srcfd_handle* h, int op) {
    int KT, NB, MAXT, NAUX, AD;
    if (op == OP_PRESSURE) shape2<OP_PRESSURE>(KT, NB, MAXT, NAUX, AD);
    else if (op == OP_UPWIND) shape2<OP_UPWIND>(KT, NB, MAXT, NAUX, AD);
    else shape2<OP_QUICK>(KT, NB, MAXT, NAUX, AD);
    const int nx = h->p.nx;
    Gs2Plan& P = h->plan2[op];
    int kmax = KT;
    if (const char* e = getenv(op == OP_PRESSURE ? "SRCFD_K_PRESSURE" : "SRCFD_K_MOMENTUM")) kmax = std::max(1, std::min(KT, atoi(e)));
    // Largest K (sweeps per group) whose banding works: every band needs >= NB*K rows so that the redundant rows
    // of the band above stay inside it, and (K+1) groups of RS slots + edge + service warps must fit the CTA.
    for (int K = kmax; K >= 1; --K) {
        const int edge = ((4 * K + 31) / 32) * 32;
        const int rs_max = ((MAXT - WF_SVC - edge) / (K + 1)) / 32 * 32;
        int rows_max = rs_max - NB * K - 2;
        if (rows_max < 1) continue;
        if (const char* e = getenv("SRCFD_BAND_ROWS")) { const int v = atoi(e); if (v >= NB * K && v < rows_max) rows_max = v; }
        bool ok = false;
        for (int nb = (nx + rows_max - 1) / rows_max; nb <= nx; ++nb) {
            const int br = (nx + nb - 1) / nb;
            const int nbe = (nx + br - 1) / br;
            const int last = nx - (nbe - 1) * br;
            if (br <= rows_max && (nbe == 1 || last >= NB * K)) { P.band_rows = br; P.nbands = nbe; ok = true; break; }
        }
        if (!ok) continue;
        P.K = K;
        P.RS = ((P.band_rows + NB * K + 2 + 31) / 32) * 32;
        P.ncomp = (K + 1) * P.RS + edge;
        P.nthreads = P.ncomp + WF_SVC;
        P.smem = sizeof(double) * ((size_t)(WF2_RING + 1) * P.ncomp + (size_t)NAUX * AD * P.RS);
        if (P.nthreads > MAXT) continue;
        int occ = 0;
        if (int rc = raise_smem_limit(h->dev, pick_gs2(op), P.smem)) return rc;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_gs2(op), P.nthreads, P.smem));
        if (occ < 1) continue;
        const int cap = h->p.max_ctas > 0 ? h->p.max_ctas : (1 << 30);
        h->grid_gs2[op] = std::max(1, std::min(cap, occ * h->num_sms));
        return SRCFD_OK;
    }
    return fail(SRCFD_ERR_ARG, "cannot band the grid for the wavefront kernel");
}

// The full-height pressure kernel needs one thread per row (+2 ghost rows) in a single CTA and tags with room for
// every boundary of a run; other grids keep the banded K-sweep kernel.
static int plan_gs3(srcfd_handle* h) {
    h->gs3 = false;
    if (const char* e = getenv("SRCFD_GS3")) if (atoi(e) == 0) return SRCFD_OK;
    const int nx = h->p.nx, ny = h->p.ny;
    const int RP = ((nx + 31) / 32) * 32 + 32;          // threads: one per row, plus the ghost warp
    if (h->inner_cap + 2 >= 4096) return SRCFD_OK;
    // two instances: up to 480 rows with groups of up to 4 sweeps (128 registers per thread), up to 990 rows with
    // one sweep per group (1024 threads leave 64 registers per thread)
    if (RP <= 512) { h->gs3_stride = 512; h->gs3_kmax = 4; }
    else if (RP <= 1024 && nx + 2 <= 1024 - 4) { h->gs3_stride = 1024; h->gs3_kmax = 1; h->gs3_nbuf = 2; }
    else return SRCFD_OK;
    h->gs3_K = std::min(h->gs3_K, h->gs3_kmax);
    if (const char* e = getenv("SRCFD_K3")) h->gs3_K = std::max(1, std::min(h->gs3_kmax, atoi(e)));
    if (const char* e = getenv("SRCFD_NBUF")) h->gs3_nbuf = std::max(2, atoi(e));
    h->gs3_RP = RP; h->gs3_ND = nx + ny + 2;
    const size_t per = sizeof(uint4) * (size_t)h->gs3_ND * h->gs3_stride;
    while (h->gs3_nbuf > 2 && per * h->gs3_nbuf > ((size_t)1 << 31)) --h->gs3_nbuf;
    if (per * h->gs3_nbuf > ((size_t)1 << 31)) return SRCFD_OK;
    h->gs3_smem = sizeof(double) * ((size_t)(3 * (h->gs3_kmax + 1) + WF3_RQ) * h->gs3_stride + 2 * (size_t)(ny + 2) + h->gs3_kmax * 32);
    if (h->gs3_smem > 200 * 1024) return SRCFD_OK;
    h->gs3_fn = RP <= 448 ? (const void*)k_solve_gs3<448, 512, 4> : RP <= 512 ? (const void*)k_solve_gs3<512, 512, 4>
                                                                              : (const void*)k_solve_gs3<1024, 1024, 1>;
    if (int rc = raise_smem_limit(h->dev, h->gs3_fn, h->gs3_smem)) return rc;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->gs3_fn, RP, h->gs3_smem));
    if (occ < 1) return SRCFD_OK;
    const int cap = h->p.max_ctas > 0 ? h->p.max_ctas : (1 << 30);
    h->gs3_grid = std::max(1, std::min(cap, h->num_sms));       // one group per SM at a time
    h->gs3 = true;
    return SRCFD_OK;
}

static int plan_jtb(srcfd_handle* h) {
    h->jtb_H = 0;
    if (const char* e = getenv("SRCFD_JTB")) if (atoi(e) == 0) return SRCFD_OK;
    const long long ncell = (long long)h->p.nx * h->p.ny;
    int H = ncell >= (1 << 20) ? 4 : 8;      // large planes are fp64-bound (less redundant halo work), small ones barrier-bound
    if (const char* e = getenv("SRCFD_JTB_H")) H = atoi(e) == 4 ? 4 : 8;
    h->jtb_fn = H == 4 ? (const void*)k_jacobi_tb<4> : (const void*)k_jacobi_tb<8>;
    h->jtb_smem = H == 4 ? JtbShape<4>::smem : JtbShape<8>::smem;
    h->jtb_pass_fn = H == 4 ? (const void*)k_jacobi_tb_pass<4> : (const void*)k_jacobi_tb_pass<8>;
    if (int rc = raise_smem_limit(h->dev, h->jtb_pass_fn, h->jtb_smem)) return rc;
    if (int rc = raise_smem_limit(h->dev, h->jtb_fn, h->jtb_smem)) return rc;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->jtb_fn, JTB_THREADS, h->jtb_smem));
    if (occ < 1) return SRCFD_OK;
    const long long tiles = (long long)((h->p.nx + JTB_TI - 1) / JTB_TI) * ((h->p.ny + JTB_TJ - 1) / JTB_TJ);
    const int cap = h->p.max_ctas > 0 ? h->p.max_ctas : (1 << 30);
    h->jtb_grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(cap, (long long)occ * h->num_sms), tiles));
    h->jtb_H = H;
    return SRCFD_OK;
}

static int plan_launches(srcfd_handle* h) {
    if (int rc = plan_gs3(h)) return rc;
    if (int rc = plan_jtb(h)) return rc;
    for (int op = 0; op < 3; ++op) if (int rc = plan_gs2(h, op)) return rc;
    const int nx = h->p.nx;
    h->nbands = (nx + WF_MAX_BAND - 1) / WF_MAX_BAND;
    h->band_rows = (nx + h->nbands - 1) / h->nbands;
    h->nbands = (nx + h->band_rows - 1) / h->band_rows;
    h->wf_threads = ((h->band_rows + 4 + 31) / 32) * 32 + WF_SVC;
    h->wf_smem = sizeof(double) * (8 * (size_t)(h->wf_threads + 4) + 4 + 32);
    const int cap = h->p.max_ctas > 0 ? h->p.max_ctas : (1 << 30);
    for (int op = 0; op < 3; ++op) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_gs(op), h->wf_threads, h->wf_smem));
        if (occ < 1) return fail(SRCFD_ERR_CUDA, "wavefront kernel does not fit on an SM");
        h->grid_gs[op] = std::max(1, std::min(cap, occ * h->num_sms));
    }
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_sync(OP_QUICK, SRCFD_ORDER_JACOBI), SYNC_THREADS, 0));
    if (occ < 1) return fail(SRCFD_ERR_CUDA, "sync kernel does not fit on an SM");
    const long long ncell = (long long)h->p.nx * h->p.ny;
    const long long want = (ncell + SYNC_THREADS - 1) / SYNC_THREADS;
    h->grid_sync = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(cap, (long long)occ * h->num_sms), want));
    h->tail_blocks = (int)want;
    return SRCFD_OK;
}

extern "C" {

int srcfd_abi_version(void) { return SRCFD_ABI_VERSION; }
const char* srcfd_last_error(void) { return g_err.c_str(); }

int srcfd_device_count(int* count) {
    if (!count) return fail(SRCFD_ERR_ARG, "null count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(SRCFD_ERR_CUDA, cudaGetErrorString(e)); }
    return SRCFD_OK;
}

int srcfd_destroy(srcfd_handle* h) {
    if (!h) return SRCFD_OK;
    cudaSetDevice(h->dev);
    if (h->stream) cudaStreamSynchronize(h->stream);
    slab_release(h);
    if (h->tm_a) cudaEventDestroy(h->tm_a);
    if (h->tm_b) cudaEventDestroy(h->tm_b);
    for (auto& e : h->ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : h->ev_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaFree(h->Var); cudaFree(h->VarOld); cudaFree(h->Ff); cudaFree(h->rhs); cudaFree(h->scratch);
    cudaFree(h->partials); cudaFree(h->res_partials); cudaFree(h->hist); cudaFree(h->prog); cudaFree(h->ctrl);
    cudaFree(h->staging); cudaFree(h->halo); cudaFree(h->trace);
    cudaFree(h->jtb_partials); cudaFree(h->jtb_sums); cudaFree(h->jtb_ticket); cudaFree(h->jtb2_partials);
    cudaFree(h->gs3_ll); cudaFree(h->gs3_rhsS); cudaFree(h->gs3_epoch);
    cudaFree(h->sweeps1); cudaFree(h->sweeps2);
    cudaFree(h->halo2); cudaFree(h->scratch2); cudaFree(h->partials2); cudaFree(h->prog2);
    if (h->ctrl_host) cudaFreeHost(h->ctrl_host);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SRCFD_OK;
}

int srcfd_create(const srcfd_params* params, srcfd_handle** out) {
    if (!out) return fail(SRCFD_ERR_ARG, "null out");
    *out = nullptr;
    if (int rc = check_params(params)) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SRCFD_ERR_CUDA, "no CUDA device: libsrcfd has no CPU fallback");
    if (params->device < 0 || params->device >= ndev) return fail(SRCFD_ERR_ARG, "bad device ordinal");
    srcfd_handle* h = new srcfd_handle();
    h->p = *params; h->K = make_consts(*params); h->bc = make_bc(*params); h->dev = params->device;
    h->inner_cap = params->inner_max;
    auto bail = [&](int rc) { std::string keep = g_err; srcfd_destroy(h); g_err = keep; return rc; };
#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(SRCFD_ERR_CUDA); } } while (0)
    CKB(cudaSetDevice(h->dev));
    cudaDeviceProp prop;
    CKB(cudaGetDeviceProperties(&prop, h->dev));
    if (!prop.cooperativeLaunch) { g_err = "device lacks cooperative launch"; return bail(SRCFD_ERR_CUDA); }
    h->num_sms = prop.multiProcessorCount;
    CKB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (const char* s = getenv("SRCFD_SPIN_LIMIT")) h->spin_limit = atoi(s);
    if (const char* s = getenv("SRCFD_GUESS_BIAS")) h->guess_bias = atoi(s);
    if (const char* s = getenv("SRCFD_GS_IMPL")) h->gs_impl = atoi(s);
    if (const char* s = getenv("SRCFD_K1MAX")) h->gs3_k1max = atoi(s);
    if (const char* s = getenv("SRCFD_SKIP_IDLE")) h->gs3_skip_idle = atoi(s);
    if (const char* s = getenv("SRCFD_PRETOUCH")) h->gs3_pretouch = atoi(s);
    if (int rc = plan_launches(h)) return bail(rc);
    const size_t P = (size_t)h->K.plane;
    const size_t pad = 4 * (size_t)h->K.pitch + 64;   // QUICK's flat over-reads stay inside the allocation
    CKB(cudaMalloc(&h->Var, sizeof(double) * (3 * P + pad)));
    CKB(cudaMalloc(&h->VarOld, sizeof(double) * (3 * P + pad)));
    CKB(cudaMalloc(&h->Ff, sizeof(double) * (4 * P + pad)));
    CKB(cudaMalloc(&h->rhs, sizeof(double) * (P + pad)));
    CKB(cudaMalloc(&h->scratch, sizeof(double) * (P + pad + 8192)));
    int gmax = std::max(h->grid_sync, std::max(h->grid_gs[0], std::max(h->grid_gs[1], h->grid_gs[2])));
    int maxbands = h->nbands;
    for (int op = 0; op < 3; ++op) { maxbands = std::max(maxbands, h->plan2[op].nbands); gmax = std::max(gmax, h->grid_gs2[op]); }
    h->n_partials = std::max((size_t)h->inner_cap * maxbands, (size_t)2 * gmax) + 64;
    if (getenv("SRCFD_TRACE")) {
        h->trace_n = std::max((size_t)(h->inner_cap / 4 + 2) * maxbands * 8, (size_t)8 * 1024 + 64);
        CKB(cudaMalloc(&h->trace, sizeof(long long) * h->trace_n));
        CKB(cudaMemsetAsync(h->trace, 0, sizeof(long long) * h->trace_n, h->stream));
    }
    const size_t halo_bytes = sizeof(double) * (size_t)2 * WF2_KMAX * maxbands * 2 * h->K.pitch + 64;
    CKB(cudaMalloc(&h->halo, halo_bytes));
    CKB(cudaMemsetAsync(h->halo, 0, halo_bytes, h->stream));
    CKB(cudaMalloc(&h->halo2, halo_bytes));
    CKB(cudaMemsetAsync(h->halo2, 0, halo_bytes, h->stream));
    CKB(cudaMalloc(&h->scratch2, sizeof(double) * (P + pad)));
    CKB(cudaMalloc(&h->sweeps1, sizeof(double) * (WF2_KMAX - 1) * (P + 8)));
    CKB(cudaMalloc(&h->sweeps2, sizeof(double) * (WF2_KMAX - 1) * (P + 8)));
    CKB(cudaMalloc(&h->partials2, sizeof(double) * h->n_partials));
    CKB(cudaMalloc(&h->prog2, sizeof(int) * ((size_t)h->inner_cap * maxbands + 64)));
    if (const char* e = getenv("SRCFD_PAIR")) h->pair_momentum = atoi(e) != 0;
    if (const char* e = getenv("SRCFD_JTB_IMPL")) h->jtb_impl = atoi(e);
    if (const char* e = getenv("SRCFD_JTB2_RB")) h->jtb2_rb_env = std::max(1, atoi(e));
    if (const char* e = getenv("SRCFD_JTB2_COLS")) h->jtb2_cols_env = atoi(e);
    h->jtb2_units_cap = (size_t)((h->p.ny + 23) / 24 + 1) * (size_t)((h->p.nx + 15) / 16 + 1);   // 24 = owned columns of a 32-column strip at 4 sweeps; chunks of >= 16 rows
    CKB(cudaMalloc(&h->jtb2_partials, sizeof(double) * 4 * h->jtb2_units_cap));
    if (h->jtb_H) { CKB(cudaMalloc(&h->jtb_partials, sizeof(double) * 2 * 8 * (size_t)(h->jtb_grid + 1))); CKB(cudaMalloc(&h->jtb_sums, sizeof(double) * 8 * 16)); CKB(cudaMemsetAsync(h->jtb_sums, 0, sizeof(double) * 8 * 16, h->stream));
                      CKB(cudaMalloc(&h->jtb_ticket, sizeof(unsigned))); CKB(cudaMemsetAsync(h->jtb_ticket, 0, sizeof(unsigned), h->stream)); }
    if (h->gs3) {
        const size_t llb = sizeof(uint4) * ((size_t)h->gs3_nbuf * h->gs3_ND + WF3_PAD_HI) * h->gs3_stride;
        CKB(cudaMalloc(&h->gs3_ll, llb));
        CKB(cudaMemsetAsync(h->gs3_ll, 0, llb, h->stream));
        const size_t rb = sizeof(double) * (size_t)(WF3_PAD_LO + h->gs3_ND + WF3_PAD_HI) * h->gs3_stride;
        CKB(cudaMalloc(&h->gs3_rhsS, rb));
        CKB(cudaMemsetAsync(h->gs3_rhsS, 0, rb, h->stream));
        CKB(cudaMalloc(&h->gs3_epoch, sizeof(unsigned long long)));
        const unsigned long long one = 1;
        CKB(cudaMemcpyAsync(h->gs3_epoch, &one, sizeof(one), cudaMemcpyHostToDevice, h->stream));
        CKB(cudaStreamSynchronize(h->stream));
    }
    CKB(cudaMalloc(&h->partials, sizeof(double) * h->n_partials));
    CKB(cudaMalloc(&h->prog, sizeof(int) * ((size_t)h->inner_cap * maxbands + 64)));
    CKB(cudaMalloc(&h->res_partials, sizeof(double) * 3 * (size_t)(h->tail_blocks + 1)));
    CKB(cudaMalloc(&h->ctrl, sizeof(Ctrl)));
    CKB(cudaMallocHost(&h->ctrl_host, sizeof(Ctrl)));
    CKB(cudaMemsetAsync(h->Var, 0, sizeof(double) * (3 * P + pad), h->stream));
    CKB(cudaMemsetAsync(h->VarOld, 0, sizeof(double) * (3 * P + pad), h->stream));
    CKB(cudaMemsetAsync(h->Ff, 0, sizeof(double) * (4 * P + pad), h->stream));
    CKB(cudaMemsetAsync(h->rhs, 0, sizeof(double) * (P + pad), h->stream));
    CKB(cudaMemsetAsync(h->scratch, 0, sizeof(double) * (P + pad), h->stream));
    CKB(cudaMemsetAsync(h->prog, 0, sizeof(int) * ((size_t)h->inner_cap * maxbands + 64), h->stream));
    memset(h->ctrl_host, 0, sizeof(Ctrl));
    h->ctrl_host->guess[0] = h->ctrl_host->guess[1] = std::min(8, h->p.inner_max);
    h->ctrl_host->guess[2] = h->p.inner_max;
    CKB(cudaMemcpyAsync(h->ctrl, h->ctrl_host, sizeof(Ctrl), cudaMemcpyHostToDevice, h->stream));
    CKB(cudaStreamSynchronize(h->stream));
#undef CKB
    *out = h;
    return SRCFD_OK;
}

int srcfd_set_params(srcfd_handle* h, const srcfd_params* params) {
    CKH(h);
    if (int rc = check_params(params)) return rc;
    if (params->nx != h->p.nx || params->ny != h->p.ny || params->device != h->p.device)
        return fail(SRCFD_ERR_ARG, "nx, ny and device are fixed at creation");
    if (params->inner_max > h->inner_cap) return fail(SRCFD_ERR_ARG, "inner_max cannot exceed its value at creation");
    const int keep_ctas = h->p.max_ctas;
    h->p = *params; h->p.max_ctas = keep_ctas;
    BcSpec nb = make_bc(*params);
    nb.skip_lo = h->bc.skip_lo; nb.skip_hi = h->bc.skip_hi;   // slab sides survive a parameter change
    if (memcmp(&nb, &h->bc, sizeof(BcSpec)) != 0) h->ghosts_fresh = false;   // new BCs: ghosts no longer known consistent
    h->K = make_consts(*params); h->bc = nb;
    return SRCFD_OK;
}

int srcfd_synchronize(srcfd_handle* h) { CKH(h); CK(cudaStreamSynchronize(h->stream)); return SRCFD_OK; }
int srcfd_stream(srcfd_handle* h, uint64_t* stream) {
    CKH(h);
    if (!stream) return fail(SRCFD_ERR_ARG, "null stream out");
    *stream = (uint64_t)(uintptr_t)h->stream;
    return SRCFD_OK;
}

int srcfd_upload(srcfd_handle* h, const double* Var, const double* VarOld, const double* Ff, const double* residual) {
    CKH(h);
    const size_t P = (size_t)h->K.plane;
    if (Var) { CK(cudaMemcpyAsync(h->Var, Var, sizeof(double) * 3 * P, cudaMemcpyHostToDevice, h->stream)); h->ghosts_fresh = false; h->jtb_ghosts_valid = false; }
    if (VarOld) CK(cudaMemcpyAsync(h->VarOld, VarOld, sizeof(double) * 3 * P, cudaMemcpyHostToDevice, h->stream));
    if (Ff) { CK(cudaMemcpyAsync(h->Ff, Ff, sizeof(double) * 4 * P, cudaMemcpyHostToDevice, h->stream)); h->ff_paired = false; }
    if (residual) CK(cudaMemcpyAsync((char*)h->ctrl + offsetof(Ctrl, residual), residual, sizeof(double) * 3, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // host buffers may be pageable and reused by the caller
    return SRCFD_OK;
}

int srcfd_download(srcfd_handle* h, double* Var, double* VarOld, double* Ff, double* residual) {
    CKH(h);
    const size_t P = (size_t)h->K.plane;
    if (Var) CK(cudaMemcpyAsync(Var, h->Var, sizeof(double) * 3 * P, cudaMemcpyDeviceToHost, h->stream));
    if (VarOld) CK(cudaMemcpyAsync(VarOld, h->VarOld, sizeof(double) * 3 * P, cudaMemcpyDeviceToHost, h->stream));
    if (Ff) CK(cudaMemcpyAsync(Ff, h->Ff, sizeof(double) * 4 * P, cudaMemcpyDeviceToHost, h->stream));
    if (residual) CK(cudaMemcpyAsync(residual, (char*)h->ctrl + offsetof(Ctrl, residual), sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SRCFD_OK;
}

int srcfd_host_alloc(uint64_t bytes, void** out) {
    if (!out || bytes == 0) return fail(SRCFD_ERR_ARG, "srcfd_host_alloc: null out or zero size");
    *out = nullptr;
    CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return SRCFD_OK;
}
int srcfd_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return SRCFD_OK;
}
int srcfd_device_ptrs(srcfd_handle* h, uint64_t* Var, uint64_t* VarOld, uint64_t* Ff) {
    CKH(h);
    if (Var) *Var = (uint64_t)(uintptr_t)h->Var;
    if (VarOld) *VarOld = (uint64_t)(uintptr_t)h->VarOld;
    if (Ff) { *Ff = (uint64_t)(uintptr_t)h->Ff; h->ff_paired = false; }   // the caller may write through the pointer
    return SRCFD_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// launch helpers (C++ linkage)
// ------------------------------------------------------------------------------------------------
#define LAUNCH_CHECK(h)                                                                            \
    do {                                                                                           \
        (h)->launches += 1;                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) return fail(SRCFD_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

static int l_copy_new_to_old(srcfd_handle* h) {
    const long long n = 3 * h->K.plane;
    const int blocks = (int)std::min<long long>((n + TAIL_THREADS - 1) / TAIL_THREADS, (long long)h->num_sms * 8);
    k_copy_new_to_old<<<blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->VarOld, n, h->ctrl);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_apply_bc(srcfd_handle* h, int k, int mode = 0, int nk = 1) {
    h->jtb_ghosts_valid = false;
    const int n = std::max(h->K.nx, h->K.ny);
    k_apply_bc<<<(n + 127) / 128, 128, 0, h->stream>>>(h->Var, k, h->K, h->bc, mode, h->ctrl, nk);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_linear_interpolation(srcfd_handle* h, bool with_rhs) {
    k_linear_interpolation<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->Ff, with_rhs ? h->rhs : nullptr, h->K, h->ctrl);
    h->ff_paired = true;
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_pressure_rhs(srcfd_handle* h) {
    k_pressure_rhs<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Ff, h->rhs, h->K, h->ctrl);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_update_flux(srcfd_handle* h) {
    k_update_flux<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->Ff, h->K, h->ctrl);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_under_relax(srcfd_handle* h, int k, double alpha, int k2 = -1, double alpha2 = 1.0) {
    k_under_relax<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->VarOld, k, alpha, h->K, h->ctrl, k2, alpha2);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_correct_velocity(srcfd_handle* h, bool assign = false) {
    k_correct_velocity<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->VarOld, h->res_partials, h->K, h->ctrl, 1, h->K.nx);
    LAUNCH_CHECK(h);
    k_residual_finish<<<1, 256, 0, h->stream>>>(h->res_partials, h->tail_blocks, h->ctrl, nullptr, assign ? 1 : 0);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int l_clear_stop(srcfd_handle* h) {
    if (!h->maybe_stopped) return SRCFD_OK;
    k_clear_stop<<<1, 1, 0, h->stream>>>(h->ctrl);
    LAUNCH_CHECK(h);
    h->maybe_stopped = false;
    return SRCFD_OK;
}

static int ev_begin(srcfd_handle* h, int kind, EvPair& ev) {
    if (h->ev_free.empty()) {
        CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
    } else { ev = h->ev_free.back(); h->ev_free.pop_back(); }
    ev.kind = kind;
    CK(cudaEventRecord(ev.a, h->stream));
    return SRCFD_OK;
}
static int ev_end(srcfd_handle* h, EvPair& ev) {
    CK(cudaEventRecord(ev.b, h->stream));
    h->ev_pending.push_back(ev);
    return SRCFD_OK;
}
static int ev_drain(srcfd_handle* h) {
    for (auto& ev : h->ev_pending) {
        CK(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        h->t_ms[ev.kind] += ms; h->t_n[ev.kind] += 1;
        h->ev_free.push_back(ev);
    }
    h->ev_pending.clear();
    return SRCFD_OK;
}

// One pass (nsw <= 4 sweeps) of the warp-streaming Jacobi kernel from plane src to plane dst; per-sweep sums of R^2 over
// rows [r0, r1] into sums[0..nsw).  The row chunk is sized so that every warp slot of the GPU gets about two units.
static int l_jtb2_pass(srcfd_handle* h, const JtbArgs& ja, const double* src, double* dst, int nsw, int r0, int r1, double* sums,
                       const int* done, unsigned long long* retries, long long* warp_steps) {
    if (nsw < 1 || nsw > 4) return fail(SRCFD_ERR_ARG, "jtb2: 1..4 sweeps per pass");
    const int slots2 = h->num_sms * JTB2_MINB * JTB2_WARPS;
    // columns per lane: 2 (64-column strips) unless the one-unit-per-slot chunks would be shorter than 48 rows -- then 1
    // (32-column strips: more units and half the registers, so more resident warps and longer chunks; measured on 4096
    // columns: 544 rows 157 GLUP/s against 118 for two columns and for the tile kernel, 1056 rows 197 against 186)
    int C = 2;
    {
        const int strips2 = (h->K.ny + (64 - 2 * nsw) - 1) / (64 - 2 * nsw);
        if ((h->K.nx + std::max(1, slots2 / strips2) - 1) / std::max(1, slots2 / strips2) < 48) C = 1;
        if (h->jtb2_cols_env == 1 || h->jtb2_cols_env == 2) C = h->jtb2_cols_env;
    }
    const int ctas_per_sm = C == 1 ? JTB2_MINB1 : JTB2_MINB;
    const int slots = h->num_sms * ctas_per_sm * JTB2_WARPS;
    Jtb2Geom g;
    g.own_cols = 32 * C - 2 * nsw;
    g.n_strips = (h->K.ny + g.own_cols - 1) / g.own_cols;
    // Rows per chunk: as many chunks as give ONE unit per warp slot (all units resident in a single round), at least 32 rows.
    // The kernel is very sensitive to this (measured, 4096 columns, 2368 slots, 74 strips -> 32 chunks): 4096 rows: 128 rows per
    // chunk 267 GLUP/s, 112 (1.16 units per slot: a second round) 176, 144 (0.9) 238; 2064 rows: 72 (0.9) 211, 64 (1.03) 198,
    // 56 156, 128 160; 1056 rows: 40 (0.84) 173, 48 155, 56 142.
    const int chunks0 = std::max(1, slots / g.n_strips);
    int RB = std::max(32, (h->K.nx + chunks0 - 1) / chunks0);
    if (h->jtb2_rb_env > 0) RB = h->jtb2_rb_env;
    g.RB = RB; g.n_chunks = (h->K.nx + RB - 1) / RB;
    const long long units = (long long)g.n_strips * g.n_chunks;
    if ((size_t)units > h->jtb2_units_cap) return fail(SRCFD_ERR_ARG, "jtb2: partial-sum buffer too small");
    const int grid = (int)std::max<long long>(1, std::min<long long>((units + JTB2_WARPS - 1) / JTB2_WARPS, (long long)h->num_sms * ctas_per_sm));
    const int cap = h->p.max_ctas > 0 ? h->p.max_ctas : (1 << 30);
    const dim3 gd(std::min(grid, cap));
#define JTB2_LAUNCH(NLv, Cv) k_jtb2_pass<NLv, Cv><<<gd, JTB2_THREADS, 0, h->stream>>>(ja, src, dst, g, r0, r1, h->jtb2_partials, sums, h->jtb_ticket, done, retries)
    if (C == 2) {
        switch (nsw) { case 1: JTB2_LAUNCH(1, 2); break; case 2: JTB2_LAUNCH(2, 2); break; case 3: JTB2_LAUNCH(3, 2); break; default: JTB2_LAUNCH(4, 2); break; }
    } else {
        switch (nsw) { case 1: JTB2_LAUNCH(1, 1); break; case 2: JTB2_LAUNCH(2, 1); break; case 3: JTB2_LAUNCH(3, 1); break; default: JTB2_LAUNCH(4, 1); break; }
    }
#undef JTB2_LAUNCH
    LAUNCH_CHECK(h);
    if (warp_steps) *warp_steps += units * (long long)(g.RB + 2 * nsw) * nsw;
    return SRCFD_OK;
}

// One inner solve: op in {OP_PRESSURE, OP_UPWIND, OP_QUICK} on plane k, counters in slot.
static int l_inner_solve(srcfd_handle* h, int op, int k, int slot, bool pair = false) {
    h->jtb_ghosts_valid = false;                            // the solves below use the scratch plane
    SolveArgs a;
    a.Var = h->Var; a.VarOld = h->VarOld; a.Ff = h->Ff; a.rhs = h->rhs; a.scratch = h->scratch;
    a.partials = h->partials; a.prog = h->prog; a.ctrl = h->ctrl; a.K = h->K;
    a.k = k; a.slot = slot; a.tol = h->p.inner_tol; a.max_iter = h->p.inner_max;
    a.nbands = h->nbands; a.band_rows = h->band_rows; a.spin_limit = h->spin_limit; a.guess_bias = h->guess_bias;
    a.omega = (h->p.sor_omega > 0.0 && op == OP_PRESSURE) ? h->p.sor_omega : 1.0;
    void* args[] = {&a};
    EvPair ev;
    if (h->timing) if (int rc = ev_begin(h, op == OP_PRESSURE ? 0 : 1, ev)) return rc;
    // RB_JACOBI: red-black (SOR) for the 5-point pressure stencil, Jacobi for the momentum stencils (QUICK is 9-point)
    const int order = h->p.sweep_order == SRCFD_ORDER_RB_JACOBI ? (op == OP_PRESSURE ? SRCFD_ORDER_RED_BLACK : SRCFD_ORDER_JACOBI)
                                                                : h->p.sweep_order;
    if (order == SRCFD_ORDER_GS_LEX && h->gs_impl == 2 && op == OP_PRESSURE && h->gs3 && !pair) {
        Gs3Args g3;
        g3.s = a; g3.K = h->gs3_K; g3.ND = h->gs3_ND; g3.nbuf = h->gs3_nbuf;
        g3.k1_max = h->gs3_k1max;
        g3.s.prog = h->trace ? h->prog : nullptr;      // this kernel has no progress flags: non-null only asks it to count polls
        g3.ll = h->gs3_ll; g3.rhsS = h->gs3_rhsS + (size_t)WF3_PAD_LO * h->gs3_stride; g3.partials = h->partials; g3.epoch = h->gs3_epoch; g3.trace = h->trace;
        g3.skip_idle = h->gs3_skip_idle;
        g3.pretouch = h->gs3_pretouch;
        void* args3[] = {&g3};
        CK(cudaLaunchCooperativeKernel(h->gs3_fn, dim3(h->gs3_grid), dim3(h->gs3_RP), args3, h->gs3_smem, h->stream));
    } else if (order == SRCFD_ORDER_GS_LEX && h->gs_impl == 2) {
        const Gs2Plan& P = h->plan2[op];
        Gs2Args ga;
        ga.s = a; ga.s.nbands = P.nbands; ga.s.band_rows = P.band_rows;
        ga.trace = h->trace;
        ga.band_rows = P.band_rows; ga.nbands = P.nbands; ga.RS = P.RS; ga.ncomp = P.ncomp; ga.K = P.K;
        ga.np = pair ? 2 : 1;
        ga.pr[0] = Gs2Prob{k, slot, h->scratch, h->partials, h->prog, h->halo, h->sweeps1};
        ga.pr[1] = Gs2Prob{1, 1, h->scratch2, h->partials2, h->prog2, h->halo2, h->sweeps2};   // pair: k = slot = 0 above, 1 here
        void* args2[] = {&ga};
        CK(cudaLaunchCooperativeKernel(pick_gs2(op), dim3(h->grid_gs2[op]), dim3(P.nthreads), args2, P.smem, h->stream));
    } else if (order == SRCFD_ORDER_GS_LEX) {
        CK(cudaLaunchCooperativeKernel(pick_gs(op), dim3(h->grid_gs[op]), dim3(h->wf_threads), args, h->wf_smem, h->stream));
    } else if (order == SRCFD_ORDER_JACOBI && op == OP_PRESSURE && h->jtb_H) {
        JtbArgs ja;
        ja.s = a; ja.partials = h->jtb_partials;
        void* argsj[] = {&ja};
        CK(cudaLaunchCooperativeKernel(h->jtb_fn, dim3(h->jtb_grid), dim3(JTB_THREADS), argsj, h->jtb_smem, h->stream));
    } else {
        if (order == SRCFD_ORDER_RED_BLACK && op == OP_QUICK)
            return fail(SRCFD_ERR_ARG, "red-black order is undefined for QUICK");
        CK(cudaLaunchCooperativeKernel(pick_sync(op, order), dim3(h->grid_sync), dim3(SYNC_THREADS), args, 0, h->stream));
    }
    h->launches += 1;
    if (h->timing) if (int rc = ev_end(h, ev)) return rc;
    return SRCFD_OK;
}

#define TRY(x) do { if (int rc_ = (x)) return rc_; } while (0)

// _implicit_solve: LDC.py:432-467 / BFS.py:622-673
static int l_implicit_solve(srcfd_handle* h) {
    // (the residual sums are ASSIGNED by k_residual_finish below, which is what zeroing them here and adding there amounts to)
    const int mop = h->p.scheme == SRCFD_SCHEME_QUICK ? OP_QUICK : OP_UPWIND;
    // The u and v momentum solves are independent (different planes, same read-only Ff): one paired wavefront
    // launch when the reference order is in use.  The BFS inlet pass for k = 0 also rewrites the v ghost column
    // (BFS.py:562) from the current v; that is a no-op exactly when the ghosts are fresh, so pair only then.
    // With stale ghosts the v solve must still see the refreshed column (applied up front, mode 4); only QUICK's u solve
    // also READS that column (flat over-read of row nx+2, hazard H4), so QUICK pairs only when the ghosts are fresh.
    const bool pair = h->pair_momentum && h->p.sweep_order == SRCFD_ORDER_GS_LEX && h->gs_impl == 2 &&
                      (!h->bc.bfs || h->ghosts_fresh || mop == OP_UPWIND);
    if (pair) {
        if (h->bc.bfs && !h->ghosts_fresh) TRY(l_apply_bc(h, 0, 4));
        TRY(l_inner_solve(h, mop, 0, 0, true));
        // relax u, BC u, relax v, BC v as two launches: the relaxations touch interior cells of their own plane only, and the one
        // cross-plane write of the BC passes (the k = 0 inlet pass resets the v ghost column, BFS.py:562) is overwritten by the
        // k = 1 pass that follows it in the same thread
        if (h->p.relax_enabled) TRY(l_under_relax(h, 0, h->p.relax[0], 1, h->p.relax[1]));
        TRY(l_apply_bc(h, 0, 0, 2));
    } else {
        for (int k = 0; k < 2; ++k) {
            TRY(l_inner_solve(h, mop, k, k));
            if (h->p.relax_enabled) TRY(l_under_relax(h, k, h->p.relax[k]));
            TRY(l_apply_bc(h, k));
        }
    }
    TRY(l_linear_interpolation(h, true));
    TRY(l_inner_solve(h, OP_PRESSURE, 2, 2));
    if (h->p.relax_enabled) TRY(l_under_relax(h, 2, h->p.relax[2]));
    TRY(l_apply_bc(h, 2));
    TRY(l_correct_velocity(h, true));
    TRY(l_apply_bc(h, 0, 0, 2));
    TRY(l_update_flux(h));
    h->ghosts_fresh = true;
    return SRCFD_OK;
}

static int fetch_ctrl(srcfd_handle* h) {
    CK(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SRCFD_OK;
}
static int push_ctrl(srcfd_handle* h) {
    CK(cudaMemcpyAsync(h->ctrl, h->ctrl_host, sizeof(Ctrl), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SRCFD_OK;
}
static int ctrl_verdict(srcfd_handle* h) {
    if (h->ctrl_host->deadlock) return fail(SRCFD_ERR_DEADLOCK, "wavefront dependence wait exceeded its spin limit");
    if (h->ctrl_host->nan_flag) return fail(SRCFD_ERR_NAN, "Solver failed: NaN/Inf in residuals");
    return SRCFD_OK;
}

extern "C" {

int srcfd_initialize_fields(srcfd_handle* h, int zero_first) {
    CKH(h); TRY(l_clear_stop(h));
    const size_t P = (size_t)h->K.plane;
    if (zero_first) {
        CK(cudaMemsetAsync(h->Var, 0, sizeof(double) * 3 * P, h->stream));
        CK(cudaMemsetAsync(h->VarOld, 0, sizeof(double) * 3 * P, h->stream));
        CK(cudaMemsetAsync(h->Ff, 0, sizeof(double) * 4 * P, h->stream));
    }
    for (int k = 0; k < 3; ++k) TRY(l_apply_bc(h, k));
    TRY(l_copy_new_to_old(h));
    TRY(l_linear_interpolation(h, false));
    h->ghosts_fresh = true;
    return SRCFD_OK;
}

int srcfd_set_fields(srcfd_handle* h, const void* fields, int is_float32) {
    CKH(h); TRY(l_clear_stop(h));
    if (!fields) return fail(SRCFD_ERR_ARG, "null fields");
    const size_t n = 3 * (size_t)h->K.nx * h->K.ny;
    const size_t bytes = n * (is_float32 ? sizeof(float) : sizeof(double));
    if (h->staging_bytes < bytes) {
        cudaFree(h->staging); h->staging = nullptr; h->staging_bytes = 0;
        CK(cudaMalloc(&h->staging, bytes));
        h->staging_bytes = bytes;
    }
    CK(cudaMemcpyAsync(h->staging, fields, bytes, cudaMemcpyHostToDevice, h->stream));
    dim3 grid((h->K.nx + 31) / 32, (h->K.ny + 31) / 32, 3), block(32, 8);
    if (is_float32) k_inject_fields<float><<<grid, block, 0, h->stream>>>(h->Var, (const float*)h->staging, h->K);
    else k_inject_fields<double><<<grid, block, 0, h->stream>>>(h->Var, (const double*)h->staging, h->K);
    LAUNCH_CHECK(h);
    TRY(srcfd_initialize_fields(h, 0));
    CK(cudaStreamSynchronize(h->stream));
    return SRCFD_OK;
}

int srcfd_step(srcfd_handle* h, int64_t n_outer, const double crit[3]) {
    CKH(h);
    if (!crit) return fail(SRCFD_ERR_ARG, "null crit");
    h->maybe_stopped = true;
    for (int64_t it = 0; it < n_outer; ++it) {
        TRY(l_implicit_solve(h));
        {   // _convergence_check + copy_new_to_old in one launch
            const long long n = 3 * h->K.plane;
            const int blocks = (int)std::min<long long>((n + TAIL_THREADS - 1) / TAIL_THREADS, (long long)h->num_sms * 8);
            k_check_and_copy<<<blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->VarOld, n, h->ctrl, h->hist, h->K, crit[0], crit[1], crit[2]);
            LAUNCH_CHECK(h);
        }
    }
    return SRCFD_OK;
}

int srcfd_status(srcfd_handle* h, int64_t* iterations, int32_t* converged, double rms[3],
                 int32_t last_sweeps[3], int64_t total_sweeps[3]) {
    CKH(h);
    TRY(fetch_ctrl(h));
    if (h->timing) TRY(ev_drain(h));
    const Ctrl& c = *h->ctrl_host;
    if (iterations) *iterations = c.iterations;
    if (converged) *converged = c.converged;
    for (int k = 0; k < 3; ++k) {
        if (rms) rms[k] = c.rms[k];
        if (last_sweeps) last_sweeps[k] = c.last_sweeps[k];
        if (total_sweeps) total_sweeps[k] = c.total_sweeps[k];
    }
    return ctrl_verdict(h);
}

int srcfd_reset_counters(srcfd_handle* h) {
    CKH(h);
    TRY(fetch_ctrl(h));
    Ctrl& c = *h->ctrl_host;
    c.stop = c.converged = c.nan_flag = c.deadlock = 0;
    h->maybe_stopped = false;
    c.iterations = 0; c.n_hist = 0;
    for (int k = 0; k < 3; ++k) { c.total_sweeps[k] = 0; c.last_sweeps[k] = 0; }
    return push_ctrl(h);
}

int srcfd_solve(srcfd_handle* h, int64_t max_iterations, const double crit[3], int64_t* iterations,
                double* seconds, double* hist, int64_t hist_cap, int64_t* n_hist) {
    CKH(h);
    if (!crit) return fail(SRCFD_ERR_ARG, "null crit");
    auto t0 = std::chrono::steady_clock::now();
    TRY(srcfd_reset_counters(h));
    const long long want_hist = max_iterations / 100 + 1;
    if (h->hist_cap < want_hist) {
        cudaFree(h->hist); h->hist = nullptr; h->hist_cap = 0;
        CK(cudaMalloc(&h->hist, sizeof(double) * 3 * want_hist));
        h->hist_cap = want_hist;
    }
    h->ctrl_host->hist_cap = h->hist_cap;
    TRY(push_ctrl(h));
    int64_t enq = 0;
    int rc = SRCFD_OK;
    while (enq < max_iterations) {
        const int64_t chunk = std::min<int64_t>(25, max_iterations - enq);
        TRY(srcfd_step(h, chunk, crit));
        enq += chunk;
        TRY(fetch_ctrl(h));
        rc = ctrl_verdict(h);
        if (rc || h->ctrl_host->stop) break;
    }
    if (h->timing) TRY(ev_drain(h));
    const Ctrl& c = *h->ctrl_host;
    if (iterations) *iterations = c.iterations;
    const int64_t nh = std::min<int64_t>(c.n_hist, hist_cap);
    if (hist && nh > 0) {
        CK(cudaMemcpyAsync(hist, h->hist, sizeof(double) * 3 * nh, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (n_hist) *n_hist = hist ? nh : 0;
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// ---- kernel-level entry points ----------------------------------------------------------------
int srcfd_k_copy_new_to_old(srcfd_handle* h) { CKH(h); TRY(l_clear_stop(h)); return l_copy_new_to_old(h); }
int srcfd_k_apply_bc(srcfd_handle* h, int k) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    if (k < 0 || k > 2) return fail(SRCFD_ERR_ARG, "k must be 0, 1 or 2");
    return l_apply_bc(h, k);
}
int srcfd_k_apply_bc_configured(srcfd_handle* h, int k) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    if (k < 0 || k > 2) return fail(SRCFD_ERR_ARG, "k must be 0, 1 or 2");
    return l_apply_bc(h, k, 1);
}
int srcfd_k_apply_bfs_inlet(srcfd_handle* h, int k) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    if (k < 0 || k > 2) return fail(SRCFD_ERR_ARG, "k must be 0, 1 or 2");
    return l_apply_bc(h, k, 2);
}
int srcfd_k_linear_interpolation(srcfd_handle* h) { CKH(h); TRY(l_clear_stop(h)); return l_linear_interpolation(h, false); }
int srcfd_k_update_flux(srcfd_handle* h) { CKH(h); TRY(l_clear_stop(h)); return l_update_flux(h); }
int srcfd_k_under_relax(srcfd_handle* h, int k, double alpha) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    if (k < 0 || k > 2) return fail(SRCFD_ERR_ARG, "k must be 0, 1 or 2");
    return l_under_relax(h, k, alpha);
}
int srcfd_k_correct_velocity(srcfd_handle* h, double residual_out[3]) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    TRY(l_correct_velocity(h));
    if (residual_out) return srcfd_download(h, nullptr, nullptr, nullptr, residual_out);
    return SRCFD_OK;
}
static int finish_inner(srcfd_handle* h, int slot, int32_t* sweeps, double* last_rms) {
    TRY(fetch_ctrl(h));
    if (h->timing) TRY(ev_drain(h));
    if (sweeps) *sweeps = h->ctrl_host->last_sweeps[slot];
    if (last_rms) *last_rms = h->ctrl_host->last_inner_rms[slot];
    return ctrl_verdict(h);
}
int srcfd_k_solve_pressure(srcfd_handle* h, int32_t* sweeps, double* last_rms) {
    CKH(h); TRY(l_clear_stop(h));
    TRY(l_pressure_rhs(h));
    TRY(l_inner_solve(h, OP_PRESSURE, 2, 2));
    return finish_inner(h, 2, sweeps, last_rms);
}
int srcfd_jacobi_pass_max(srcfd_handle* h, int* H) {
    CKH(h);
    if (!H) return fail(SRCFD_ERR_ARG, "null H");
    *H = h->jtb_H;
    return SRCFD_OK;
}
int srcfd_k_jacobi_pass(srcfd_handle* h, int nsweeps, int own_row0, int own_row1, int recompute_rhs, int commit, int slot,
                        double* sums) {
    CKH(h); TRY(l_clear_stop(h));
    if (!h->jtb_H) return fail(SRCFD_ERR_ARG, "the temporally blocked Jacobi kernel is disabled (SRCFD_JTB=0)");
    if (nsweeps < 1 || nsweeps > h->jtb_H) return fail(SRCFD_ERR_ARG, "nsweeps must be 1..srcfd_jacobi_pass_max()");
    if (own_row0 < 1 || own_row1 > h->p.nx || own_row0 > own_row1) return fail(SRCFD_ERR_ARG, "bad row range");
    if (slot < 0 || slot >= 16) return fail(SRCFD_ERR_ARG, "slot must be 0..15");
    if (recompute_rhs) TRY(l_pressure_rhs(h));
    JtbArgs ja;
    SolveArgs& a = ja.s;
    a.Var = h->Var; a.VarOld = h->VarOld; a.Ff = h->Ff; a.rhs = h->rhs; a.scratch = h->scratch;
    a.partials = h->partials; a.prog = h->prog; a.ctrl = h->ctrl; a.K = h->K;
    a.k = 2; a.slot = 2; a.tol = 0.0; a.max_iter = nsweeps;
    a.nbands = h->nbands; a.band_rows = h->band_rows; a.spin_limit = h->spin_limit; a.guess_bias = 0; a.omega = 1.0;
    ja.partials = h->jtb_partials;
    // boundary cells of the scratch plane must match the plane (the pass only writes interior cells)
    if (!h->jtb_ghosts_valid) {                             // they only change with the plane's own boundary cells
        k_jacobi_tb_ghosts<<<(std::max(h->K.nx, h->K.ny) + 2 + 127) / 128, 128, 0, h->stream>>>(a);
        LAUNCH_CHECK(h);
        h->launches += 1;
        h->jtb_ghosts_valid = true;
    }
    double* sums_dev = h->jtb_sums + 8 * slot;
    const double* src = h->Var + 2 * (size_t)h->K.plane;
    double* dst = h->scratch;
    const int* done = nullptr;
    unsigned long long* retries = nullptr;
    void* args[] = {&ja, &src, &dst, &nsweeps, &own_row0, &own_row1, &sums_dev, &h->jtb_ticket, &done, &retries};
    CK(cudaLaunchKernel(h->jtb_pass_fn, dim3(h->jtb_grid), dim3(JTB_THREADS), args, h->jtb_smem, h->stream));
    h->launches += 1;
    if (commit) {
        k_jacobi_tb_commit<<<std::max(1, h->tail_blocks / 4), 256, 0, h->stream>>>(a);
        LAUNCH_CHECK(h);
        h->launches += 1;
    }
    if (sums) {                                             // NULL: the caller reduces the device copy (srcfd_jacobi_sums_ptr)
        CK(cudaMemcpyAsync(sums, h->jtb_sums + 8 * slot, sizeof(double) * nsweeps, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return SRCFD_OK;
}
// whole pressure plane (halos and ghosts included) <-> the snapshot buffer: speculative blocks of passes roll back with it
int srcfd_k_jacobi_snapshot(srcfd_handle* h, int restore) {
    CKH(h);
    double* p = h->Var + 2 * (size_t)h->K.plane;
    if (restore) CK(cudaMemcpyAsync(p, h->scratch2, sizeof(double) * (size_t)h->K.plane, cudaMemcpyDeviceToDevice, h->stream));
    else CK(cudaMemcpyAsync(h->scratch2, p, sizeof(double) * (size_t)h->K.plane, cudaMemcpyDeviceToDevice, h->stream));
    return SRCFD_OK;
}
int srcfd_jacobi_sums_ptr(srcfd_handle* h, uint64_t* ptr) {
    CKH(h);
    if (!ptr) return fail(SRCFD_ERR_ARG, "null ptr");
    *ptr = (uint64_t)(uintptr_t)h->jtb_sums;
    return SRCFD_OK;
}
int srcfd_k_jacobi_commit(srcfd_handle* h) {
    CKH(h); TRY(l_clear_stop(h));
    SolveArgs a;
    a.Var = h->Var; a.VarOld = h->VarOld; a.Ff = h->Ff; a.rhs = h->rhs; a.scratch = h->scratch;
    a.partials = h->partials; a.prog = h->prog; a.ctrl = h->ctrl; a.K = h->K;
    a.k = 2; a.slot = 2; a.tol = 0.0; a.max_iter = 0;
    a.nbands = h->nbands; a.band_rows = h->band_rows; a.spin_limit = h->spin_limit; a.guess_bias = 0; a.omega = 1.0;
    k_jacobi_tb_commit<<<std::max(1, h->tail_blocks / 4), 256, 0, h->stream>>>(a);
    LAUNCH_CHECK(h);
    h->launches += 1;
    return SRCFD_OK;
}
int srcfd_k_solve_momentum(srcfd_handle* h, int k, int scheme, int32_t* sweeps, double* last_rms) {
    CKH(h); TRY(l_clear_stop(h));
    h->ghosts_fresh = false;
    if (k < 0 || k > 1) return fail(SRCFD_ERR_ARG, "momentum is solved for k = 0 (u) or 1 (v)");
    if (scheme != SRCFD_SCHEME_UPWIND && scheme != SRCFD_SCHEME_QUICK) return fail(SRCFD_ERR_ARG, "bad scheme");
    TRY(l_inner_solve(h, scheme == SRCFD_SCHEME_QUICK ? OP_QUICK : OP_UPWIND, k, k));
    return finish_inner(h, k, sweeps, last_rms);
}
int srcfd_k_implicit_solve(srcfd_handle* h) { CKH(h); TRY(l_clear_stop(h)); return l_implicit_solve(h); }

int srcfd_timer_start(srcfd_handle* h) {
    CKH(h);
    if (!h->tm_a) { CK(cudaEventCreate(&h->tm_a)); CK(cudaEventCreate(&h->tm_b)); }
    CK(cudaEventRecord(h->tm_a, h->stream));
    return SRCFD_OK;
}
int srcfd_timer_stop(srcfd_handle* h, double* ms) {
    CKH(h);
    if (!h->tm_a || !ms) return fail(SRCFD_ERR_ARG, "timer not started / null out");
    CK(cudaEventRecord(h->tm_b, h->stream));
    CK(cudaEventSynchronize(h->tm_b));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, h->tm_a, h->tm_b));
    *ms = f;
    return SRCFD_OK;
}
int srcfd_debug_read(srcfd_handle* h, double* out, int64_t n) {   // debug builds: raw doubles stored behind the scratch plane
    CKH(h);
    const int64_t slack = 4 * (int64_t)h->K.pitch + 8192;       // what srcfd_create allocated behind the plane (pad + 8192 - 64)
    if (!out || n < 0 || n > slack) return fail(SRCFD_ERR_ARG, "srcfd_debug_read: n exceeds the slack behind the scratch plane");
    CK(cudaMemcpy(out, h->scratch + h->K.plane + 64, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SRCFD_OK;
}
int srcfd_trace_read(srcfd_handle* h, long long* out, int64_t n) {   // SRCFD_TRACE=1 only
    CKH(h);
    if (!h->trace) return fail(SRCFD_ERR_ARG, "tracing is off (set SRCFD_TRACE=1 before creating the handle)");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->trace, sizeof(long long) * std::min<size_t>(n, h->trace_n), cudaMemcpyDeviceToHost));
    return SRCFD_OK;
}
int srcfd_launch_count(srcfd_handle* h, int64_t* launches) {
    if (!h || !launches) return fail(SRCFD_ERR_ARG, "null argument");
    *launches = h->launches;
    return SRCFD_OK;
}
int srcfd_timing_enable(srcfd_handle* h, int enabled) {
    CKH(h);
    if (h->timing) TRY(ev_drain(h));
    h->timing = enabled != 0;
    h->t_ms[0] = h->t_ms[1] = 0; h->t_n[0] = h->t_n[1] = 0;
    return SRCFD_OK;
}
int srcfd_timing_read(srcfd_handle* h, double* pressure_ms, int64_t* pressure_launches, double* momentum_ms,
                      int64_t* momentum_launches) {
    CKH(h);
    TRY(ev_drain(h));
    if (pressure_ms) *pressure_ms = h->t_ms[0];
    if (pressure_launches) *pressure_launches = h->t_n[0];
    if (momentum_ms) *momentum_ms = h->t_ms[1];
    if (momentum_launches) *momentum_launches = h->t_n[1];
    return SRCFD_OK;
}

}  // extern "C"

// ---- batched small-grid solves (coarse_batch.cuh) ---------------------------------------------------------------
extern "C" int srcfd_coarse_smem_bytes(int nx, int ny, uint64_t* bytes) {
    if (nx < 1 || ny < 1 || !bytes) return fail(SRCFD_ERR_ARG, "bad argument");
    const long long P = (long long)(nx + 2) * (ny + 2);
    const int M2 = (ny + 1) / 2, M3 = (ny + 2) / 3;
    const int ring2 = (nx - 1 + 2 * M2 - 1) / 2 + 2, ring3 = (nx - 1 + 3 * M3 - 1) / 3 + 2;
    const long long ring = std::max(ring2, ring3), stride = (long long)nx * M2;
    *bytes = (uint64_t)(12 * P + ring * stride) * sizeof(double);
    return SRCFD_OK;
}

// The size gate of srcfd_coarse_solve_batch as a query (same tests, same device attributes), so that callers need not copy it.
extern "C" int srcfd_coarse_fits(int nx, int ny, int device, int* fits) {
    if (!fits) return fail(SRCFD_ERR_ARG, "null fits");
    *fits = 0;
    uint64_t smem = 0;
    if (int rc = srcfd_coarse_smem_bytes(nx, ny, &smem)) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SRCFD_ERR_CUDA, "no CUDA device: libsrcfd has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(SRCFD_ERR_ARG, "bad device ordinal");
    CK(cudaSetDevice(device));
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, (const void*)k_coarse_solve));
    const int M2 = (ny + 1) / 2, threads = ((nx * M2 + 31) / 32) * 32 + 32;
    *fits = (smem + 2048 <= (uint64_t)optin && threads <= fa.maxThreadsPerBlock) ? 1 : 0;
    return SRCFD_OK;
}

extern "C" int srcfd_coarse_solve_batch(const srcfd_params* params, int n_cases, int64_t max_iterations, const double* crit,
                                        int resume, double* Var, double* VarOld, double* Ff, srcfd_coarse_result* results,
                                        double* hist, int64_t hist_cap, double* ms) {
    if (!params || n_cases < 1 || !crit || !Var || !results) return fail(SRCFD_ERR_ARG, "null argument");
    if (resume && (!VarOld || !Ff)) return fail(SRCFD_ERR_ARG, "resume needs Var, VarOld and Ff");
    if (max_iterations < 0 || hist_cap < 0) return fail(SRCFD_ERR_ARG, "negative count");
    const int nx = params[0].nx, ny = params[0].ny, dev = params[0].device;
    std::vector<CoarseCase> cases((size_t)n_cases);
    for (int c = 0; c < n_cases; ++c) {
        const srcfd_params& p = params[c];
        if (int rc = check_params(&p)) return rc;
        if (p.nx != nx || p.ny != ny || p.device != dev) return fail(SRCFD_ERR_ARG, "all cases of a batch share nx, ny and device");
        if (p.sweep_order != SRCFD_ORDER_GS_LEX) return fail(SRCFD_ERR_ARG, "the batched solver runs the reference sweep order only");
        CoarseCase& cs = cases[(size_t)c];
        cs.K = make_consts(p); cs.bc = make_bc(p);
        cs.scheme = p.scheme; cs.relax_enabled = p.relax_enabled;
        for (int k = 0; k < 3; ++k) { cs.relax[k] = p.relax[k]; cs.crit[k] = crit[3 * c + k]; }
        cs.inner_tol = p.inner_tol; cs.inner_max = p.inner_max; cs.max_iterations = max_iterations;
    }
    uint64_t smem = 0;
    if (int rc = srcfd_coarse_smem_bytes(nx, ny, &smem)) return rc;
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            return fail(SRCFD_ERR_CUDA, "no CUDA device: libsrcfd has no CPU fallback");
        if (dev < 0 || dev >= ndev) return fail(SRCFD_ERR_ARG, "bad device ordinal");
    }
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (smem + 2048 > (uint64_t)prop.sharedMemPerBlockOptin)
        return fail(SRCFD_ERR_ARG, "grid too large for the shared-memory resident batched solver (use srcfd_create/srcfd_solve)");
    const int M2 = (ny + 1) / 2, M3 = (ny + 2) / 3;
    const int ring = std::max((nx - 1 + 2 * M2 - 1) / 2 + 2, (nx - 1 + 3 * M3 - 1) / 3 + 2), stride = nx * M2;
    const int threads = ((stride + 31) / 32) * 32 + 32;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, (const void*)k_coarse_solve));
    if (threads > fa.maxThreadsPerBlock) return fail(SRCFD_ERR_ARG, "grid too large for one CTA per case");
    if (int rc = raise_smem_limit(dev, (const void*)k_coarse_solve, (size_t)smem)) return rc;

    const size_t P = (size_t)(nx + 2) * (ny + 2), nV = 3 * P * (size_t)n_cases, nF = 4 * P * (size_t)n_cases;
    CoarseCase* d_cases = nullptr; CoarseOut* d_out = nullptr;
    double *d_Var = nullptr, *d_VarOld = nullptr, *d_Ff = nullptr, *d_hist = nullptr;
    cudaStream_t st = nullptr; cudaEvent_t ea = nullptr, eb = nullptr;
    int rc = SRCFD_OK;
    auto body = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ea)); CK(cudaEventCreate(&eb));
        CK(cudaMalloc(&d_cases, sizeof(CoarseCase) * (size_t)n_cases));
        CK(cudaMalloc(&d_out, sizeof(CoarseOut) * (size_t)n_cases));
        CK(cudaMalloc(&d_Var, sizeof(double) * nV)); CK(cudaMalloc(&d_VarOld, sizeof(double) * nV));
        CK(cudaMalloc(&d_Ff, sizeof(double) * nF));
        if (hist && hist_cap > 0) CK(cudaMalloc(&d_hist, sizeof(double) * 3 * (size_t)hist_cap * (size_t)n_cases));
        CK(cudaMemcpyAsync(d_cases, cases.data(), sizeof(CoarseCase) * (size_t)n_cases, cudaMemcpyHostToDevice, st));
        if (resume) {
            CK(cudaMemcpyAsync(d_Var, Var, sizeof(double) * nV, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_VarOld, VarOld, sizeof(double) * nV, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_Ff, Ff, sizeof(double) * nF, cudaMemcpyHostToDevice, st));
        }
        CK(cudaEventRecord(ea, st));
        k_coarse_solve<<<n_cases, threads, (size_t)smem, st>>>(d_cases, d_Var, d_VarOld, d_Ff, d_out, d_hist,
                                                               (long long)hist_cap, ring, stride, resume ? 1 : 0);
        CK(cudaGetLastError());
        CK(cudaEventRecord(eb, st));
        std::vector<CoarseOut> outs((size_t)n_cases);
        CK(cudaMemcpyAsync(outs.data(), d_out, sizeof(CoarseOut) * (size_t)n_cases, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(Var, d_Var, sizeof(double) * nV, cudaMemcpyDeviceToHost, st));
        if (VarOld) CK(cudaMemcpyAsync(VarOld, d_VarOld, sizeof(double) * nV, cudaMemcpyDeviceToHost, st));
        if (Ff) CK(cudaMemcpyAsync(Ff, d_Ff, sizeof(double) * nF, cudaMemcpyDeviceToHost, st));
        if (d_hist) CK(cudaMemcpyAsync(hist, d_hist, sizeof(double) * 3 * (size_t)hist_cap * (size_t)n_cases, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (ms) { float f = 0.f; CK(cudaEventElapsedTime(&f, ea, eb)); *ms = f; }
        for (int c = 0; c < n_cases; ++c) {
            const CoarseOut& o = outs[(size_t)c];
            srcfd_coarse_result& r = results[c];
            r.iterations = o.iterations; r.converged = o.converged; r.nan_flag = o.nan_flag; r.n_hist = o.n_hist;
            for (int k = 0; k < 3; ++k) {
                r.rms[k] = o.rms[k]; r.total_sweeps[k] = o.total_sweeps[k]; r.last_inner_rms[k] = o.last_inner_rms[k];
                r.residual[k] = o.residual[k]; r.last_sweeps[k] = o.last_sweeps[k];
            }
        }
        return SRCFD_OK;
    };
    rc = body();
    if (d_cases) cudaFree(d_cases);
    if (d_out) cudaFree(d_out);
    if (d_Var) cudaFree(d_Var);
    if (d_VarOld) cudaFree(d_VarOld);
    if (d_Ff) cudaFree(d_Ff);
    if (d_hist) cudaFree(d_hist);
    if (ea) cudaEventDestroy(ea);
    if (eb) cudaEventDestroy(eb);
    if (st) cudaStreamDestroy(st);
    return rc;
}

#include "slab_api.inl"
