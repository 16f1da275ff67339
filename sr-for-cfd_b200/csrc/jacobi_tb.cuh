// jacobi_tb.cuh -- pressure solve in JACOBI order (opt-in), temporally blocked: H sweeps per pass over HBM.
//
// Same update and break rule as solve_pressure (LDC.py:292-314), with every cell of a sweep computed from the
// previous iterate (the oracle's ORDER_JACOBI restatement); results are bit-identical to that restatement.
//   * a pass moves the plane from buffer A to buffer B through shared memory ONCE and advances it by up to H
//     sweeps: a CTA loads a (TI + 2H) x (TJ + 2H) tile (owned cells + halo), relaxes it H times in a shared-memory
//     ping-pong -- the region that is still exact shrinks by one ring per sweep, except along domain edges, whose
//     ghost cells are constant -- and writes its TI x TJ owned cells.  DRAM traffic per cell update drops from 24 B
//     to (8 + 2*8*redundancy)/H B; the right-hand side tile is loaded once per pass;
//   * residual sums of each of the H sweeps are accumulated over owned cells only, per CTA in tile order, then
//     summed over CTAs in index order by every CTA after one grid-wide barrier per PASS (not per sweep);
//   * exact break: if a sweep before the last one of a pass meets the tolerance, the pass is simply repeated with
//     fewer sweeps -- buffer A is untouched by a pass.
#pragma once
#include "inner_gs3.cuh"

namespace srcfd {

constexpr int JTB_TI = 32, JTB_TJ = 64, JTB_THREADS = 256;

struct JtbArgs {
    SolveArgs s;
    double* partials;           // [2][H][gridDim] per-CTA residual sums
};

template <int H>
struct JtbShape {
    static constexpr int RI = JTB_TI + 2 * H, RJ = JTB_TJ + 2 * H;    // tile with halo
    static constexpr size_t smem = sizeof(double) * 3 * RI * RJ;      // two value buffers + right-hand side
};

// One pass: nsw (<= H) sweeps from plane src to plane dst; acc[t] += sum of R^2 of sweep t over this CTA's owned cells.
template <int H>
__device__ void jtb_pass(const SolveArgs& a, const double* __restrict__ src, double* __restrict__ dst, const int nsw,
                         double* sm, double* acc, const Gs3Div& D, const int res_r0, const int res_r1,
                         unsigned long long* __restrict__ retries = nullptr) {
    constexpr int RI = JtbShape<H>::RI, RJ = JtbShape<H>::RJ;
    const Consts& K = a.K;
    double* b0 = sm;
    double* b1 = sm + RI * RJ;
    double* rh = sm + 2 * RI * RJ;
    const int ti_n = (K.nx + JTB_TI - 1) / JTB_TI, tj_n = (K.ny + JTB_TJ - 1) / JTB_TJ;
    const double volp = K.volp;
    for (int tile = blockIdx.x; tile < ti_n * tj_n; tile += gridDim.x) {
        const int i0 = 1 + (tile / tj_n) * JTB_TI, j0 = 1 + (tile % tj_n) * JTB_TJ;     // first owned cell
        const int gi0 = i0 - H, gj0 = j0 - H;                                           // global index of tile cell (0,0)
        // clip the tile to the plane (ghost lines included)
        const int li0 = max(0, -gi0), li1 = min(RI - 1, K.nx + 1 - gi0);
        const int lj0 = max(0, -gj0), lj1 = min(RJ - 1, K.ny + 1 - gj0);
        __syncthreads();                                    // previous tile's readers are done with the buffers
        // all loads of a batch are issued before the first store: the tile arrives in one or two memory round trips
        constexpr int PER = (RI * RJ + JTB_THREADS - 1) / JTB_THREADS, BATCH = 6;
#pragma unroll
        for (int q0 = 0; q0 < PER; q0 += BATCH) {
            double v[BATCH], r[BATCH];
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                const int idx = threadIdx.x + (q0 + q) * JTB_THREADS;
                const int li = idx / RJ, lj = idx - li * RJ;
                v[q] = 0.0; r[q] = 0.0;
                if (q0 + q < PER && idx < RI * RJ && li >= li0 && li <= li1 && lj >= lj0 && lj <= lj1) {
                    const long long c = (long long)(gi0 + li) * K.pitch + (gj0 + lj);
                    v[q] = __ldcg(src + c);
                    r[q] = __ldg(a.rhs + c);
                }
            }
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                const int idx = threadIdx.x + (q0 + q) * JTB_THREADS;
                if (q0 + q < PER && idx < RI * RJ) { b0[idx] = v[q]; b1[idx] = v[q]; rh[idx] = r[q]; }   // b1 too: ghosts are never rewritten
            }
        }
        __syncthreads();
        double* cur = b0;
        double* nxt = b1;
        // column / row-segment of this thread (RJ columns x NSEG segments of SEGR rows; the few threads beyond idle)
        constexpr int NSEG = JTB_THREADS / RJ, SEGR = (RI + NSEG - 1) / NSEG;
        const int lj = threadIdx.x % RJ, sg = threadIdx.x / RJ;
        const int seg0 = sg * SEGR, seg1 = min(RI - 1, seg0 + SEGR - 1);
        const int gj = gj0 + lj;
        const bool col_ok = sg < NSEG && gj >= 1 && gj <= K.ny;                        // interior column
        const bool own_col = lj >= H && lj < H + JTB_TJ;
        for (int t = 1; t <= nsw; ++t) {
            // exact region of sweep t: shrinks by one ring per sweep from tile edges inside the domain; stays put
            // along edges that are domain boundaries (li0 > 0 etc. means the tile was clipped there)
            const int ri0 = (li0 > 0 ? li0 : t), ri1 = (li1 < RI - 1 ? li1 : RI - 1 - t);
            const int rj0 = (lj0 > 0 ? lj0 : t), rj1 = (lj1 < RJ - 1 ? lj1 : RJ - 1 - t);
            double part = 0.0;
            // thread = (column lj, row segment): walks down its column keeping (i-1,j) and (i,j) in registers, so a cell
            // costs 4 shared loads (i+1, j-1, j+1, rhs) and one store; four rows per trip give four independent chains
            if (col_ok && lj >= rj0 && lj <= rj1) {
                const int ra = max(max(ri0, seg0), 1 - gi0), rb = min(min(ri1, seg1), K.nx - gi0);    // interior rows only
                if (ra <= rb) {
                    int idx = ra * RJ + lj;
                    double im = cur[idx - RJ], c = cur[idx];
                    int li = ra;
                    for (; li + 3 <= rb; li += 4, idx += 4 * RJ) {          // four independent chains, one range flag
                        double v[6], nv[4], R[4];
                        v[0] = im; v[1] = c;
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q + 2] = cur[idx + (q + 1) * RJ];
                        bool bad = false;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            bool fail = false;
                            nv[q] = pressure_cell3z(v[q + 1], v[q + 2], v[q], cur[idx + q * RJ + 1], cur[idx + q * RJ - 1],
                                                   rh[idx + q * RJ], volp, D, R[q], fail);
                            bad = bad || fail;
                        }
                        if (__builtin_expect(bad, 0)) {
                            if (retries) atomicAdd(retries, 1ull);   // four-cell groups that took the IEEE routine (the host steers by it)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const double2 o = pressure_cell3_ieee(v[q + 1], v[q + 2], v[q], cur[idx + q * RJ + 1],
                                                                      cur[idx + q * RJ - 1], rh[idx + q * RJ], volp, D.dx2.b,
                                                                      D.dy2.b, D.apd.b);
                                nv[q] = o.x; R[q] = o.y;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            nxt[idx + q * RJ] = nv[q];
                            if (own_col && li + q >= H && li + q < H + JTB_TI && gi0 + li + q >= res_r0 && gi0 + li + q <= res_r1) part += R[q] * R[q];    // owned cell
                        }
                        im = v[4]; c = v[5];
                    }
                    for (; li <= rb; ++li, idx += RJ) {
                        const double ip = cur[idx + RJ];
                        double R;
                        bool fail = false;
                        double nv = pressure_cell3z(c, ip, im, cur[idx + 1], cur[idx - 1], rh[idx], volp, D, R, fail);
                        if (__builtin_expect(fail, 0)) {
                            const double2 o = pressure_cell3_ieee(c, ip, im, cur[idx + 1], cur[idx - 1], rh[idx], volp, D.dx2.b,
                                                                  D.dy2.b, D.apd.b);
                            nv = o.x; R = o.y;
                        }
                        nxt[idx] = nv;
                        if (own_col && li >= H && li < H + JTB_TI && gi0 + li >= res_r0 && gi0 + li <= res_r1) part += R * R;      // owned cell
                        im = c; c = ip;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < H; ++q)
                if (q == t - 1) acc[q] += part;             // static register index
            __syncthreads();
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
        // owned cells of the last sweep
        for (int idx = threadIdx.x; idx < JTB_TI * JTB_TJ; idx += JTB_THREADS) {
            const int oi = idx / JTB_TJ, oj = idx - oi * JTB_TJ;
            const int gi = i0 + oi, gj = j0 + oj;
            if (gi <= K.nx && gj <= K.ny) dst[(long long)gi * K.pitch + gj] = cur[(oi + H) * RJ + (oj + H)];
        }
    }
}

template <int H>
__global__ void __launch_bounds__(JTB_THREADS) k_jacobi_tb(JtbArgs ja) {
    cg::grid_group grid = cg::this_grid();
    const SolveArgs& a = ja.s;
    if (a.ctrl->stop) return;
    extern __shared__ double smem[];
    __shared__ double red[32];
    __shared__ double s_tot[H];
    const Consts& K = a.K;
    double* A = a.Var + (long long)a.k * K.plane;
    double* Bp = a.scratch;
    const long long ncell = (long long)K.nx * K.ny;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    for (long long t = gtid; t < K.plane; t += gsize) Bp[t] = A[t];      // the second buffer needs the ghost cells too
    Gs3Div D;
    D.dx2 = make_invdiv3(K.dx2); D.dy2 = make_invdiv3(K.dy2); D.apd = make_invdiv3(K.ap_d);
    grid.sync();
    const double* src = A;
    double* dst = Bp;
    int n = 0, pass = 0;
    double rms = 0.0;
    bool done = false;
    while (!done) {
        int nsw = min(H, a.max_iter - n);
        for (;;) {
            double acc[H];
#pragma unroll
            for (int t = 0; t < H; ++t) acc[t] = 0.0;
            jtb_pass<H>(a, src, dst, nsw, smem, acc, D, 1, K.nx);
            double* part = ja.partials + (size_t)(pass & 1) * H * gridDim.x;
            ++pass;
#pragma unroll
            for (int t = 0; t < H; ++t) {
                const double tot = block_sum(acc[t], red);
                if (threadIdx.x == 0) part[(size_t)t * gridDim.x + blockIdx.x] = tot;
            }
            grid.sync();
            for (int t = 0; t < nsw; ++t) {
                double s = 0.0;
                for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(part + (size_t)t * gridDim.x + b);
                const double all = block_sum(s, red);
                if (threadIdx.x == 0) s_tot[t] = all;
            }
            __syncthreads();
            int first = -1;
            for (int t = 0; t < nsw; ++t)
                if (sqrt(s_tot[t] / (double)ncell) < a.tol) { first = t; break; }
            __syncthreads();
            if (first >= 0 && first != nsw - 1) { nsw = first + 1; continue; }     // repeat the pass with fewer sweeps (src is intact)
            rms = sqrt(s_tot[nsw - 1] / (double)ncell);
            n += nsw;
            if (first >= 0 || n >= a.max_iter) done = true;
            break;
        }
        const double* t = src; src = dst; dst = const_cast<double*>(t);
    }
    if (src != A) {                                          // latest iterate lives in the scratch plane
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

// ---- one pass on its own (slab decomposition, srcfd/slab.py): no grid-wide barrier, so an ordinary launch -------------
// nsw (<= H) sweeps from the plane to the scratch plane; per-CTA residual sums of every sweep, counted over rows
// [res_r0, res_r1] only (a slab's halo rows are relaxed too, but belong to the neighbour).
// The last CTA to finish adds the per-CTA partial sums up, in CTA order (deterministic), into sums[0..nsw).
template <int H>
__global__ void __launch_bounds__(JTB_THREADS) k_jacobi_tb_pass(JtbArgs ja, const double* __restrict__ src, double* __restrict__ dst,
                                                                int nsw, int res_r0, int res_r1, double* __restrict__ sums,
                                                                unsigned* __restrict__ ticket, const int* __restrict__ done,
                                                                unsigned long long* __restrict__ retries) {
    const SolveArgs& a = ja.s;
    if (a.ctrl->stop) return;
    if (done && *(const volatile int*)done) return;          // slab solves: the break test was met in an earlier block
    extern __shared__ double smem[];
    __shared__ double red[32];
    const Consts& K = a.K;
    Gs3Div D;
    D.dx2 = make_invdiv3(K.dx2); D.dy2 = make_invdiv3(K.dy2); D.apd = make_invdiv3(K.ap_d);
    double acc[H];
#pragma unroll
    for (int t = 0; t < H; ++t) acc[t] = 0.0;
    jtb_pass<H>(a, src, dst, nsw, smem, acc, D, res_r0, res_r1, retries);
#pragma unroll
    for (int t = 0; t < H; ++t) {
        const double tot = block_sum(acc[t], red);
        if (threadIdx.x == 0) ja.partials[(size_t)t * gridDim.x + blockIdx.x] = tot;
    }
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        __threadfence();                                     // partials visible before the ticket
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if ((int)threadIdx.x < nsw) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(ja.partials + (size_t)threadIdx.x * gridDim.x + b);
            sums[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;                  // ready for the next pass (stream order)
    }
}
// boundary cells of the plane -> scratch plane (a pass writes interior cells only)
__global__ void k_jacobi_tb_ghosts(SolveArgs a) {
    if (a.ctrl->stop) return;
    const Consts& K = a.K;
    const double* A = a.Var + (long long)a.k * K.plane;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= K.ny + 1) { a.scratch[t] = A[t]; a.scratch[(long long)(K.nx + 1) * K.pitch + t] = A[(long long)(K.nx + 1) * K.pitch + t]; }
    if (t <= K.nx + 1) { a.scratch[(long long)t * K.pitch] = A[(long long)t * K.pitch]; a.scratch[(long long)t * K.pitch + K.ny + 1] = A[(long long)t * K.pitch + K.ny + 1]; }
}
// accept a pass: the interior of the scratch plane becomes the plane
__global__ void k_jacobi_tb_commit(SolveArgs a) {
    if (a.ctrl->stop) return;
    const Consts& K = a.K;
    double* A = a.Var + (long long)a.k * K.plane;
    const long long ncell = (long long)K.nx * K.ny;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < ncell; idx += (long long)gridDim.x * blockDim.x) {
        const long long c = (idx / K.ny + 1) * K.pitch + (idx % K.ny) + 1;
        A[c] = a.scratch[c];
    }
}

// =====================================================================================================================
// Second generation: warp-streaming temporal blocking (no shared memory, no block barrier).
//
// A WARP owns a strip of 64 columns (two per lane) and walks down a chunk of rows.  The NL time levels of a pass are a
// register pipeline: at step i the warp loads row i of the source plane and then, for t = 1..NL, computes row i-t of
// sweep t from the three most recent rows of sweep t-1, which it holds in registers (a 3-row window per level and
// column); the left/right neighbours are the lane's other column or one shuffle away.  Row i-NL of the last sweep is
// stored.  Nothing is staged in shared memory and no thread waits for another warp, so the per-sweep CTA barrier of
// the tile kernel above (37 % of its warp samples) is gone and latency is hidden by the other warps of the SM.
//   * validity: a lane within t columns of the strip edge holds garbage at sweep t (its neighbour was outside the
//     strip), so a strip owns its 64 - 2*NL middle columns; likewise a chunk streams NL rows above and below the rows
//     it owns.  Redundancy (64 / 56) x (RB + 8) / RB at NL = 4.
//   * boundary cells are constant during an inner solve (hazard H6): a cell that is not interior passes its value
//     from level to level unchanged, so ghost rows / columns need no special case and the out-of-plane lanes of an
//     edge strip simply hold zeros;
//   * per-sweep residual sums over owned interior cells: lane -> warp (xor shuffles) -> one partial per (sweep, unit),
//     added up in unit order by the last CTA to finish (fixed order: deterministic);
//   * arithmetic: pressure_cell3z, i.e. the same exact-reciprocal division sequence as above (IEEE fallback per level).
// DRAM traffic per pass: the plane and the right-hand side read once (x redundancy), the plane written once.
// =====================================================================================================================
#ifndef JTB2_MINB
#define JTB2_MINB 2
#endif
#ifndef JTB2_T
#define JTB2_T 256
#endif
#ifndef JTB2_PF
#define JTB2_PF 0
#endif
#ifndef JTB2_MINB1
#define JTB2_MINB1 2      // resident CTAs per SM of the one-column-per-lane variant (3 = 80 registers with spills: 97 against 157 GLUP/s on 544 rows)
#endif
constexpr int JTB2_THREADS = JTB2_T, JTB2_WARPS = JTB2_THREADS / 32;

// Second try for a pair of cells whose fast division missed its range test: zero numerators (fields at rest) are exact
// with a select (pressure_cell3z), anything else takes the IEEE routine.  Out of line: the steady-state loop only tests.
struct Jtb2Pair { double xA, RA, xB, RB; };
__device__ __noinline__ Jtb2Pair jtb2_retry_pair(double cA, double dnA, double upA, double cB, double dnB, double upB, double lft,
                                                 double rgt, double rhA, double rhB, double volp, const Gs3Div& D) {
    Jtb2Pair o;
    bool f1 = false, f2 = false;
    o.xA = pressure_cell3z(cA, dnA, upA, cB, lft, rhA, volp, D, o.RA, f1);
    o.xB = pressure_cell3z(cB, dnB, upB, rgt, cA, rhB, volp, D, o.RB, f2);
    if (f1) { const double2 t = pressure_cell3_ieee(cA, dnA, upA, cB, lft, rhA, volp, D.dx2.b, D.dy2.b, D.apd.b); o.xA = t.x; o.RA = t.y; }
    if (f2) { const double2 t = pressure_cell3_ieee(cB, dnB, upB, rgt, cA, rhB, volp, D.dx2.b, D.dy2.b, D.apd.b); o.xB = t.x; o.RB = t.y; }
    return o;
}

struct Jtb2Geom {
    int own_cols;     // 32*C - 2*NL
    int n_strips, n_chunks, RB;
};

struct Jtb2Lane {     // what a lane knows about its two columns
    bool inA, inB;    // inside the plane (ghost columns included): loads allowed
    bool intA, intB;  // interior column: the cell is relaxed (else its value passes through)
    bool ownA, ownB;  // interior and owned by this strip: counted and stored
};

// Steady-state step (every level's row is an interior row that has been streamed, three rows ahead exist): no range
// tests, and the 3-row windows / the 3 rows in flight rotate by NAME -- PH = step mod 3 is a template argument, so the
// slot indices are compile-time and nothing is moved.  Level L keeps row rho in slot (rho - i0 + L) mod 3, which makes
// "written this step" = PH, "centre" = PH+2, "row above" = PH+1 at every level.
// C = columns per lane: 2 (64-column strips, the default) or 1 (32-column strips: twice the units for thin slabs, where
// the two-column form cannot fill the warp slots with chunks long enough to pay for their lead-in rows).
template <int NL, int PH, int C>
__device__ __forceinline__ void jtb2_fast_step(int i, double (&wA)[NL][3], double (&wB)[NL][3], double (&rqA)[NL + 1], double (&rqB)[NL + 1],
                                               double (&acc)[NL], double (&nA)[3], double (&nB)[3], const double*& pA,
                                               const double*& qA, double*& oA, const int pitch, const Jtb2Lane& L, const double volp,
                                               const Gs3Div& D, const int ra, const int rb, const int sr0, const int sr1,
                                               const int i_valid, unsigned long long* __restrict__ retries) {
    constexpr int NEW = PH, UP = (PH + 1) % 3, CEN = (PH + 2) % 3;
    const double curA = nA[PH], curB = (C == 2) ? nB[PH] : 0.0;
    nA[PH] = L.inA ? __ldcg(pA + 3 * pitch) : 0.0;                    // row i+3 of the plane
    if (C == 2) nB[PH] = L.inB ? __ldcg(pA + 3 * pitch + 1) : 0.0;
#if JTB2_PF > 0
    if (L.inA) {                                                      // rows further ahead: into L2 only (no registers held)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pA + (3 + JTB2_PF) * pitch));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(qA + (1 + JTB2_PF) * pitch));
    }
#endif
    // right-hand side: rqX[t] = row i-1-t.  Row i is requested now and first used (as rqX[0]) one step from now.
#pragma unroll
    for (int t = NL - 1; t > 0; --t) { rqA[t] = rqA[t - 1]; if (C == 2) rqB[t] = rqB[t - 1]; }
    rqA[0] = rqA[NL];
    rqA[NL] = L.inA ? __ldg(qA) : 0.0;
    if (C == 2) { rqB[0] = rqB[NL]; rqB[NL] = L.inB ? __ldg(qA + 1) : 0.0; }
    pA += pitch; qA += pitch;
    wA[0][NEW] = curA;
    if (C == 2) wB[0][NEW] = curB;
    // All NL levels are computed without a branch between them (level t+1 needs level t only as its "row below", so most
    // of its arithmetic overlaps level t); a division that missed its range test is noticed ONCE, at the end of the step,
    // and the step is then redone level by level with the zero-safe / IEEE path -- every input of the redo is intact:
    // level 0 holds loaded rows, and each level's slot NEW is rewritten before the next level reads it.
    bool bad = false;
    double ss[NL];
#pragma unroll
    for (int t = 1; t <= NL; ++t) {
        const int r = i - t;
        const double cA = wA[t - 1][CEN], cB = (C == 2) ? wB[t - 1][CEN] : 0.0;
        const double lft = __shfl_up_sync(0xffffffffu, C == 2 ? cB : cA, 1);     // column jA - 1
        const double rgt = __shfl_down_sync(0xffffffffu, cA, 1);                 // column jB + 1 (C = 1: jA + 1)
        double RA, RB_ = 0.0;
        bool failA = false, failB = false;
        // plain range test: a widened test that lets exact +0 numerators pass (two integer instructions per division) was
        // measured at 170 against 267 GLUP/s on the same box -- fields with zeros or denormals are the tile kernel's job
        const double xA = pressure_cell3(cA, wA[t - 1][NEW], wA[t - 1][UP], C == 2 ? cB : rgt, lft, rqA[t - 1], volp, D, RA, failA);
        double xB = 0.0;
        if (C == 2) xB = pressure_cell3(cB, wB[t - 1][NEW], wB[t - 1][UP], rgt, cA, rqB[t - 1], volp, D, RB_, failB);
        // a miss only matters in an interior column (the out-of-plane lanes of an edge strip hold zeros and always miss)
        // and once the level is fed by streamed rows (the first 2t steps of a chunk compute lead-in garbage)
        bad = bad || (((failA && L.intA) || (C == 2 && failB && L.intB)) && i - i_valid >= 2 * t);
        const double nvA = L.intA ? xA : cA, nvB = L.intB ? xB : cB;
        ss[t - 1] = (C == 2) ? RA * RA + RB_ * RB_ : RA * RA;         // a lane owns both of its columns or neither
        if (t < NL) {
            wA[t][NEW] = nvA;
            if (C == 2) wB[t][NEW] = nvB;
        } else if ((unsigned)(r - ra) <= (unsigned)(rb - ra)) {
            if (L.ownA) oA[0] = nvA;
            if (C == 2 && L.ownB) oA[1] = nvB;
        }
    }
    if (__builtin_expect(__any_sync(0xffffffffu, bad), 0)) {
        if (retries && (threadIdx.x & 31) == 0) atomicAdd(retries, 1ull);   // the host steers by this count (see slab_api.inl)
#pragma unroll
        for (int t = 1; t <= NL; ++t) {
            const int r = i - t;
            const double cA = wA[t - 1][CEN], cB = (C == 2) ? wB[t - 1][CEN] : 0.0;
            const double lft = __shfl_up_sync(0xffffffffu, C == 2 ? cB : cA, 1);
            const double rgt = __shfl_down_sync(0xffffffffu, cA, 1);
            // C = 1: the "B cell" of the pair routine is fed the A cell's own operands and ignored
            const Jtb2Pair o = (C == 2) ? jtb2_retry_pair(cA, wA[t - 1][NEW], wA[t - 1][UP], cB, wB[t - 1][NEW], wB[t - 1][UP], lft, rgt,
                                                          rqA[t - 1], rqB[t - 1], volp, D)
                                        : jtb2_retry_pair(cA, wA[t - 1][NEW], wA[t - 1][UP], rgt, wA[t - 1][NEW], wA[t - 1][UP], lft, rgt,
                                                          rqA[t - 1], rqA[t - 1], volp, D);
            const double nvA = L.intA ? o.xA : cA, nvB = L.intB ? o.xB : cB;
            ss[t - 1] = (C == 2) ? o.RA * o.RA + o.RB * o.RB : o.RA * o.RA;
            if (t < NL) {
                wA[t][NEW] = nvA;
                if (C == 2) wB[t][NEW] = nvB;
            } else if ((unsigned)(r - ra) <= (unsigned)(rb - ra)) {
                if (L.ownA) oA[0] = nvA;
                if (C == 2 && L.ownB) oA[1] = nvB;
            }
        }
    }
#pragma unroll
    for (int t = 1; t <= NL; ++t)
        if ((unsigned)(i - t - sr0) <= (unsigned)(sr1 - sr0)) acc[t - 1] += ss[t - 1];     // warp-uniform row test
    oA += pitch;
}

template <int NL, int C>
__global__ void __launch_bounds__(JTB2_THREADS, (C == 1) ? JTB2_MINB1 : JTB2_MINB) k_jtb2_pass(JtbArgs ja, const double* __restrict__ src, double* __restrict__ dst,
                                                                       Jtb2Geom g, int res_r0, int res_r1, double* __restrict__ partials,
                                                                       double* __restrict__ sums, unsigned* __restrict__ ticket,
                                                                       const int* __restrict__ done,
                                                                       unsigned long long* __restrict__ retries) {
    const SolveArgs& a = ja.s;
    if (a.ctrl->stop) return;
    if (done && *(const volatile int*)done) return;
    __shared__ double red[32];
    const Consts& K = a.K;
    Gs3Div D;
    D.dx2 = make_invdiv3(K.dx2); D.dy2 = make_invdiv3(K.dy2); D.apd = make_invdiv3(K.ap_d);
    const double volp = K.volp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_units = g.n_strips * g.n_chunks;
    const int pitch = K.pitch;
    for (int u = blockIdx.x * JTB2_WARPS + warp; u < n_units; u += gridDim.x * JTB2_WARPS) {
        const int s = u % g.n_strips, ch = u / g.n_strips;
        const int jA = 1 + s * g.own_cols - NL + C * lane, jB = jA + 1;
        const int ra = 1 + ch * g.RB, rb = min(K.nx, ra + g.RB - 1);
        const int i_start = max(0, ra - NL), i_end = rb + NL;
        Jtb2Lane L;
        L.inA = jA >= 0 && jA <= K.ny + 1; L.inB = C == 2 && jB >= 0 && jB <= K.ny + 1;
        L.intA = jA >= 1 && jA <= K.ny; L.intB = C == 2 && jB >= 1 && jB <= K.ny;
        L.ownA = L.intA && C * lane >= NL && C * lane < NL + g.own_cols;
        L.ownB = L.intB && 2 * lane + 1 >= NL && 2 * lane + 1 < NL + g.own_cols;
        // a lane that owns exactly one of its two columns (odd NL, odd ny): the unit masks cell by cell in the generic steps
        const bool split = C == 2 && __any_sync(0xffffffffu, L.ownA != L.ownB);
        int sr0 = max(ra, res_r0), sr1 = min(rb, res_r1);             // rows whose residual this unit counts
        if (sr1 < sr0) sr0 = sr1 = 0x3fffffff;                        // none: keeps the unsigned range test of the steady-state step false
        const int i_valid = (ra - NL <= 0) ? -0x3fffffff : i_start;   // level t is fed by streamed rows from step i_valid + 2t on
        double wA[NL][3], wB[NL][3];                                  // window of level t: rows (newest-2, newest-1, newest)
        double rqA[NL + 1], rqB[NL + 1];                              // right-hand side of rows i-1 .. i-NL; [NL] = row i, in flight
        double acc[NL];
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            acc[t] = 0.0; rqA[t] = rqB[t] = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) wA[t][q] = wB[t][q] = 0.0;
        }
        rqA[NL] = rqB[NL] = 0.0;
        // plane rows are requested two steps ahead of their use (three in the steady state)
        const double* pA = src + (long long)i_start * pitch + jA;
        const double* qA = a.rhs + (long long)i_start * pitch + jA;
        double* oA = dst + (long long)(i_start - NL) * pitch + jA;
        double n0A = 0.0, n0B = 0.0, n1A = 0.0, n1B = 0.0;            // plane rows i, i+1 in flight
        if (i_start <= K.nx + 1) {
            if (L.inA) n0A = __ldcg(pA);
            if (L.inB) n0B = __ldcg(pA + 1);
        }
        if (i_start + 1 <= K.nx + 1) {
            if (L.inA) n1A = __ldcg(pA + pitch);
            if (L.inB) n1B = __ldcg(pA + pitch + 1);
        }
        // generic step: every range is tested (first / last rows of a chunk, plane edges, split units)
        auto slow_step = [&](const int i) {
            double curA = n0A, curB = n0B;
            n0A = n1A; n0B = n1B;
            n1A = n1B = 0.0;
            if (i + 2 <= K.nx + 1) {
                if (L.inA) n1A = __ldcg(pA + 2 * pitch);
                if (L.inB) n1B = __ldcg(pA + 2 * pitch + 1);
            }
#pragma unroll
            for (int t = NL - 1; t > 0; --t) { rqA[t] = rqA[t - 1]; rqB[t] = rqB[t - 1]; }
            rqA[0] = rqA[NL]; rqB[0] = rqB[NL];                       // rqX[t-1] = right-hand side of row i-t
            rqA[NL] = rqB[NL] = 0.0;
            if (i <= K.nx + 1) {
                if (L.inA) rqA[NL] = __ldg(qA);
                if (L.inB) rqB[NL] = __ldg(qA + 1);
            }
            pA += pitch; qA += pitch;
            wA[0][0] = wA[0][1]; wA[0][1] = wA[0][2]; wA[0][2] = curA;
            wB[0][0] = wB[0][1]; wB[0][1] = wB[0][2]; wB[0][2] = curB;
#pragma unroll
            for (int t = 1; t <= NL; ++t) {
                const int r = i - t;                                  // row computed at this level (warp-uniform)
                double nvA = wA[t - 1][1], nvB = wB[t - 1][1];        // not interior: the value passes through
                if (r >= i_start && r <= K.nx + 1) {
                    const double cA = wA[t - 1][1], cB = (C == 2) ? wB[t - 1][1] : 0.0;
                    const double lft = __shfl_up_sync(0xffffffffu, C == 2 ? cB : cA, 1);     // column jA - 1
                    const double rgt = __shfl_down_sync(0xffffffffu, cA, 1);                 // column jB + 1 (C = 1: jA + 1)
                    if (r >= 1 && r <= K.nx) {
                        double RA, RB_ = 0.0;
                        bool f1 = false, f2 = false;
                        const double jpA = (C == 2) ? cB : rgt;
                        double xA = pressure_cell3z(cA, wA[t - 1][2], wA[t - 1][0], jpA, lft, rqA[t - 1], volp, D, RA, f1);
                        double xB = 0.0;
                        if (C == 2) xB = pressure_cell3z(cB, wB[t - 1][2], wB[t - 1][0], rgt, cA, rqB[t - 1], volp, D, RB_, f2);
                        if (__builtin_expect((f1 && L.intA) || (f2 && L.intB), 0)) {
                            const double2 o1 = pressure_cell3_ieee(cA, wA[t - 1][2], wA[t - 1][0], jpA, lft, rqA[t - 1], volp, D.dx2.b, D.dy2.b, D.apd.b);
                            xA = o1.x; RA = o1.y;
                            if (C == 2) {
                                const double2 o2 = pressure_cell3_ieee(cB, wB[t - 1][2], wB[t - 1][0], rgt, cA, rqB[t - 1], volp, D.dx2.b, D.dy2.b, D.apd.b);
                                xB = o2.x; RB_ = o2.y;
                            }
                        }
                        if (L.intA) nvA = xA;
                        if (L.intB) nvB = xB;
                        if (r >= sr0 && r <= sr1) {
                            acc[t - 1] += L.ownA ? RA * RA : 0.0;
                            acc[t - 1] += L.ownB ? RB_ * RB_ : 0.0;
                        }
                    }
                }
                if (t < NL) {
                    wA[t][0] = wA[t][1]; wA[t][1] = wA[t][2]; wA[t][2] = nvA;
                    wB[t][0] = wB[t][1]; wB[t][1] = wB[t][2]; wB[t][2] = nvB;
                } else if (r >= ra && r <= rb) {
                    if (L.ownA) oA[0] = nvA;
                    if (L.ownB) oA[1] = nvB;
                }
            }
            oA += pitch;
        };
        // Steady-state steps need every level's row to be interior (1 <= i-NL, i-1 <= nx) and row i+3 to exist; rows above
        // the chunk's first streamed row only produce lead-in garbage, which is never stored, counted or retried, so an
        // interior chunk runs steady-state steps from its first row on.  Generic steps: ghost rows of the first and the
        // last chunk of the plane, the one or two rows left over by the unroll-by-3, and units with a split lane.
        int i = i_start;
        const int f_lo = (i_start == 0) ? NL + 1 : i_start, f_hi = min(K.nx - 2, i_end);
        bool steady_done = split;
        while (i <= i_end) {
            if (!steady_done && i >= f_lo) {
                steady_done = true;
                const int nfast = ((f_hi - i + 1) / 3) * 3;
                if (nfast > 0) {
                    double nA[3], nB[3];
                    nA[0] = n0A; nB[0] = n0B;
                    nA[1] = n1A; nB[1] = n1B;
                    nA[2] = L.inA ? __ldcg(pA + 2 * pitch) : 0.0;
                    nB[2] = L.inB ? __ldcg(pA + 2 * pitch + 1) : 0.0;
                    for (int m = 0; m < nfast; m += 3, i += 3) {
                        jtb2_fast_step<NL, 0, C>(i, wA, wB, rqA, rqB, acc, nA, nB, pA, qA, oA, pitch, L, volp, D, ra, rb, sr0, sr1, i_valid, retries);
                        jtb2_fast_step<NL, 1, C>(i + 1, wA, wB, rqA, rqB, acc, nA, nB, pA, qA, oA, pitch, L, volp, D, ra, rb, sr0, sr1, i_valid, retries);
                        jtb2_fast_step<NL, 2, C>(i + 2, wA, wB, rqA, rqB, acc, nA, nB, pA, qA, oA, pitch, L, volp, D, ra, rb, sr0, sr1, i_valid, retries);
                    }
                    n0A = nA[0]; n0B = nB[0];                          // rows i, i+1 for the generic steps (row i+2 is re-requested there)
                    n1A = nA[1]; n1B = nB[1];
                    continue;
                }
            }
            slow_step(i);
            ++i;
        }
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            double v = (split || L.ownA) ? acc[t] : 0.0;             // split units masked cell by cell; else a lane owns both or none
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) partials[(size_t)t * n_units + u] = v;
        }
    }
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int t = 0; t < NL; ++t) {
        double s = 0.0;
        for (int b = threadIdx.x; b < n_units; b += blockDim.x) s += __ldcg(partials + (size_t)t * n_units + b);
        const double all = block_sum(s, red);
        if (threadIdx.x == 0) sums[t] = all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

}  // namespace srcfd
