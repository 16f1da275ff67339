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
                         double* sm, double* acc, const Gs3Div& D, const int res_r0, const int res_r1) {
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
                                                                unsigned* __restrict__ ticket, const int* __restrict__ done) {
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
    jtb_pass<H>(a, src, dst, nsw, smem, acc, D, res_r0, res_r1);
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

}  // namespace srcfd
