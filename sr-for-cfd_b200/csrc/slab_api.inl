// slab_api.inl -- host side of the slab decomposition (slab.cuh), included at the end of srcfd.cu.
//
// A slab is an ordinary srcfd_handle created for the LOCAL grid (owned rows + halo rows, sweep_order = JACOBI) and
// then told its place with srcfd_slab_configure.  Every entry point takes an ARRAY of handles: one per rank that this
// process drives.  One process per GPU (the product layout) passes one handle and maps the other ranks' mailboxes
// with cudaIpc (srcfd_slab_export / srcfd_slab_attach_ipc); a single process may also drive several slabs itself
// (srcfd_slab_attach_local) -- on one device they then share one stream and are enqueued in lock step, which is how
// the protocol is tested on a single GPU without kernels that wait for each other.
//
// Control: the host enqueues blocks of sweeps speculatively and reads the device-side verdict once per inner solve
// in the steady state (once per "round"); nothing crosses PCIe between sweeps and no library collective is involved.

struct SlabState {
    int world = 1, rank = 0, nx_global = 0, halo = 0;
    int lo = 0, hi = 0, own0 = 1, own1 = 0;      // halo rows towards rank-1 / rank+1, owned local rows (inclusive)
    void* mail = nullptr;
    size_t mail_bytes = 0;
    void* peer_base[SLAB_MAX_WORLD] = {};
    bool peer_ipc[SLAB_MAX_WORLD] = {};
    SlabCtl* sc = nullptr;        // device
    SlabCtl* sc_host = nullptr;   // pinned
    double* sums = nullptr;       // [SLAB_NS] per-sweep sums of the running block (device)
    double* sweep_partials = nullptr;
    unsigned* tickets = nullptr;  // [0] push, [1] momentum sweep
    unsigned long long seq = 0;
    int guess[3] = {8, 8, 0};     // sweep counts of the last inner solves (u, v, p) and of the ones before: the prediction
    int guess_prev[3] = {0, 0, 0};
    int block_cap = 0;            // SRCFD_SLAB_BLOCK: cap on the sweeps per block (0 = as many as the halo allows)
    int64_t exchanges = 0;
    int64_t halo_bytes = 0;       // bytes this rank pushed to neighbours
    int64_t replays = 0;
    // The streaming pressure kernel (k_jtb2_pass) is the default; a field with a band of denormal values (a cavity
    // started from rest: the pressure front decays by ~1/4 per cell until it underflows) sends whole column strips to
    // the IEEE division routine on every step, and a strip is one warp's work for a whole chunk.  The kernel counts
    // those warp-steps; when they exceed 0.01 % in a solve (stragglers: the kernel waits for its slowest warp), the next solves use the tile kernel (k_jacobi_tb_pass,
    // same bits, small 2-D work units), which counts its own trips to the IEEE routine: the streaming kernel comes
    // back when a tile-kernel solve reports (almost) none.  A handle starts on tiles, i.e. its first solve is the probe.
    bool use_tiles = true;
    int stream_nl = 4;            // SRCFD_JTB2_NL (experiments): sweeps per pass of the streaming kernel
    bool force_stream = false;    // SRCFD_JTB2_FORCE (experiments): streaming kernel on thin slabs too
    bool four_faces = false;      // SRCFD_SLAB_FOUR_FACES (tests): momentum sweeps read the stored west-flux plane in every row
    int sweep2 = 1;               // momentum sweeps two per pass over HBM (k_slab_sweep2): bit 0 upwind, bit 1 QUICK.  Default: upwind
                                  // only (4096^2: 115 against 92 GLUP/s; QUICK 72 against 87: its 5-row windows, 14 % strip overlap and
                                  // spills cost more than the saved traffic).  SRCFD_SLAB_SWEEP2=0 none, =1 both
    double* sweep2_partials = nullptr;   // [units][2] partial sums of k_slab_sweep2
    int sweep2_slots[2] = {0, 0}; // resident warps of k_slab_sweep2<upwind / QUICK> on this device
    bool south_from_north = true;  // SRCFD_SLAB_SOUTH_FROM_NORTH=0: the one-sweep kernel streams the stored south-flux plane instead of taking the
                                  // negated north flux of the cell to the left (library-produced fluxes only; 4096^2: upwind 93.6 -> 98.5, QUICK 87.8 -> 89.9 GLUP/s)
    int sweep2_pf = 3;            // SRCFD_SWEEP2_PF: rows ahead that k_slab_sweep2 prefetches into L2 (0 = off; 4096^2 upwind: 109 GLUP/s off, 121 at 2-4, 113 at 12)
    int sweep2_chunks = 0;        // SRCFD_SWEEP2_CHUNKS (experiments): row chunks per strip, 0 = one unit per resident warp
    int64_t sweep2_passes = 0;
    long long warp_steps = 0;     // of the running solve
    int64_t tile_solves = 0, stream_solves = 0;
};

static void slab_release(srcfd_handle* h) {
    SlabState* S = h->slab;
    if (!S) return;
    for (int q = 0; q < SLAB_MAX_WORLD; ++q)
        if (S->peer_ipc[q] && S->peer_base[q]) cudaIpcCloseMemHandle(S->peer_base[q]);
    cudaFree(S->mail); cudaFree(S->sc); cudaFree(S->sums); cudaFree(S->sweep_partials); cudaFree(S->sweep2_partials); cudaFree(S->tickets);
    if (S->sc_host) cudaFreeHost(S->sc_host);
    delete S;
    h->slab = nullptr;
}

static SolveArgs slab_solve_args(srcfd_handle* h, int k, int slot) {
    SolveArgs a;
    a.Var = h->Var; a.VarOld = h->VarOld; a.Ff = h->Ff; a.rhs = h->rhs; a.scratch = h->scratch;
    a.partials = h->partials; a.prog = h->prog; a.ctrl = h->ctrl; a.K = h->K;
    a.k = k; a.slot = slot; a.tol = h->p.inner_tol; a.max_iter = h->p.inner_max;
    a.nbands = h->nbands; a.band_rows = h->band_rows; a.spin_limit = h->spin_limit; a.guess_bias = 0; a.omega = 1.0;
    return a;
}
static double* slab_buf(srcfd_handle* h, int k, int idx) {
    return idx == 0 ? h->Var + (size_t)k * h->K.plane : idx == 1 ? h->scratch : h->scratch2;
}

// All slabs of one process on one device share the first one's stream while a group call runs, so that the launches
// below execute in exactly the order they are enqueued (a gate then never has to wait for a later launch).
struct SlabGroup {
    srcfd_handle* const* hs;
    int n;
    std::vector<cudaStream_t> saved;
    bool shared = false;
    SlabGroup(srcfd_handle* const* hs_, int n_) : hs(hs_), n(n_) {
        bool same = n > 1;
        for (int i = 1; i < n; ++i) same = same && hs[i]->dev == hs[0]->dev;
        if (same) {
            shared = true;
            for (int i = 0; i < n; ++i) { cudaSetDevice(hs[i]->dev); cudaStreamSynchronize(hs[i]->stream); saved.push_back(hs[i]->stream); hs[i]->stream = hs[0]->stream; }
        }
    }
    ~SlabGroup() {
        if (shared) {
            cudaSetDevice(hs[0]->dev);
            cudaStreamSynchronize(hs[0]->stream);
            for (int i = 0; i < n; ++i) hs[i]->stream = saved[(size_t)i];
        }
    }
    srcfd_handle* operator[](int i) const { return hs[i]; }
};

static int slab_check_group(srcfd_handle* const* hs, int n) {
    if (!hs || n < 1) return fail(SRCFD_ERR_ARG, "no slab handles");
    for (int i = 0; i < n; ++i) {
        if (!hs[i] || !hs[i]->slab) return fail(SRCFD_ERR_ARG, "handle is not configured as a slab (srcfd_slab_configure)");
        const SlabState* S = hs[i]->slab;
        for (int q = 0; q < S->world; ++q)
            if (!S->peer_base[q]) return fail(SRCFD_ERR_ARG, "slab peers are not attached (srcfd_slab_attach_ipc / _local)");
        if (hs[i]->p.sweep_order != SRCFD_ORDER_JACOBI) return fail(SRCFD_ERR_ARG, "slab handles run the JACOBI order");
    }
    return SRCFD_OK;
}

// ---- one exchange: every local rank pushes, then every local rank gates ---------------------------------------------
struct SlabPlanes { int k0, i0, k1, i1, npl; };     // (plane, rotation-buffer index) of the rows that travel

static int slab_push(srcfd_handle* h, const SlabPlanes& pl, int nsums) {
    SlabState* S = h->slab;
    CK(cudaSetDevice(h->dev));
    SlabPushArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = S->sc; a.ctrl = h->ctrl;
    a.npl = S->world > 1 ? pl.npl : 0;
    a.src0 = pl.npl > 0 ? slab_buf(h, pl.k0, pl.i0) : nullptr;
    a.src1 = pl.npl > 1 ? slab_buf(h, pl.k1, pl.i1) : nullptr;
    a.pitch = h->K.pitch; a.halo = S->halo; a.own0 = S->own0; a.own1 = S->own1;
    a.rank = S->rank; a.world = S->world; a.seq = S->seq;
    a.sums_src = S->sums; a.nsums = nsums;
    a.has_lo = S->lo > 0; a.has_hi = S->hi > 0;
    if (a.has_lo) a.lo = slab_mail_at(S->peer_base[S->rank - 1]);
    if (a.has_hi) a.hi = slab_mail_at(S->peer_base[S->rank + 1]);
    for (int q = 0; q < S->world; ++q) {
        const SlabMail m = slab_mail_at(S->peer_base[q]);
        a.peer_sums[q] = m.sums; a.peer_sum_flag[q] = m.sum_flag;
    }
    a.ticket = S->tickets + 0;
    const long long n = (long long)a.npl * S->halo * h->K.pitch;
    const int grid = (int)std::max<long long>(1, std::min<long long>(h->num_sms, (n + SLAB_THREADS * 8 - 1) / (SLAB_THREADS * 8)));
    k_slab_push<<<grid, SLAB_THREADS, 0, h->stream>>>(a);
    LAUNCH_CHECK(h);
    S->halo_bytes += (int64_t)sizeof(double) * n * ((a.has_lo ? 1 : 0) + (a.has_hi ? 1 : 0));
    return SRCFD_OK;
}
static int slab_gate(srcfd_handle* h, const SlabPlanes& pl, int nsums, int mode, int block, int nsw, double tol) {
    SlabState* S = h->slab;
    CK(cudaSetDevice(h->dev));
    SlabGateArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = S->sc; a.ctrl = h->ctrl;
    a.npl = S->world > 1 ? pl.npl : 0;
    a.dst0 = pl.npl > 0 ? slab_buf(h, pl.k0, pl.i0) : nullptr;
    a.dst1 = pl.npl > 1 ? slab_buf(h, pl.k1, pl.i1) : nullptr;
    a.pitch = h->K.pitch; a.halo = S->halo; a.own0 = S->own0; a.own1 = S->own1;
    a.rank = S->rank; a.world = S->world; a.seq = S->seq;
    a.me = slab_mail_at(S->mail);
    a.has_lo = S->lo > 0; a.has_hi = S->hi > 0;
    a.nsums = nsums; a.mode = mode; a.block = block; a.nsw = nsw; a.tol = tol;
    a.ncell_global = (double)((long long)S->nx_global * (long long)h->K.ny);
    a.spin_limit = h->spin_limit;
    const long long n = (long long)a.npl * S->halo * h->K.pitch;
    const int grid = (int)std::max<long long>(1, std::min<long long>(h->num_sms, (n + SLAB_THREADS * 8 - 1) / (SLAB_THREADS * 8)));
    k_slab_gate<<<grid, SLAB_THREADS, 0, h->stream>>>(a);
    LAUNCH_CHECK(h);
    return SRCFD_OK;
}
static int slab_exchange(SlabGroup& G, const SlabPlanes& pl, int nsums, int mode, int block, int nsw, double tol) {
    for (int i = 0; i < G.n; ++i) { G[i]->slab->seq += 1; G[i]->slab->exchanges += 1; TRY(slab_push(G[i], pl, nsums)); }
    for (int i = 0; i < G.n; ++i) TRY(slab_gate(G[i], pl, nsums, mode, block, nsw, tol));
    return SRCFD_OK;
}

// ---- one block of sweeps from rotation buffer S; the result lands in buffer E (never S) -----------------------------
static int slab_run_block(srcfd_handle* h, int op, int k, int Sidx, int nsw, int& E) {
    SlabState* S = h->slab;
    CK(cudaSetDevice(h->dev));
    const int o1 = (Sidx + 1) % 3, o2 = (Sidx + 2) % 3;
    int src = Sidx, dst = o1;
    const int* done = &S->sc->done;
    if (op == OP_PRESSURE) {
        if (!h->jtb_H) return fail(SRCFD_ERR_ARG, "the temporally blocked Jacobi kernel is disabled (SRCFD_JTB=0)");
        JtbArgs ja;
        ja.s = slab_solve_args(h, 2, 2); ja.partials = h->jtb_partials;
        for (int t = 0; t < nsw;) {
            // Very thin slabs keep the tile kernel.  Measured per-rank compute on 4096 columns (tools/thin_slab_probe.py,
            // GLUP/s streaming / tiles): 288 rows 97 / 101, 544 rows 156 / 118 (one column per lane), 800 rows 184 / 127,
            // 1056 rows 197 / 137, 2064 rows 228 / 147, 4096 rows 267 / 152 -- the streaming kernel from ~12 rows per
            // two-column warp-slot chunk on (384 rows at 4096 columns).
            const int strips4 = (h->K.ny + 55) / 56, slots = h->num_sms * JTB2_MINB * JTB2_WARPS;
            const bool roomy = S->force_stream || h->K.nx >= 12 * std::max(1, slots / strips4);
            const bool stream = h->jtb_impl == 2 && !S->use_tiles && roomy;
            int m = std::min(stream ? S->stream_nl : h->jtb_H, nsw - t);
            const double* sp = slab_buf(h, k, src);
            double* dp = slab_buf(h, k, dst);
            double* sums = S->sums + t;
            int r0 = S->own0, r1 = S->own1;
            if (stream) {
                TRY(l_jtb2_pass(h, ja, sp, dp, m, r0, r1, sums, done, &S->sc->retries, &S->warp_steps));
            } else {
                unsigned long long* retries = &S->sc->retries;
                void* args[] = {&ja, &sp, &dp, &m, &r0, &r1, &sums, &h->jtb_ticket, &done, &retries};
                CK(cudaLaunchKernel(h->jtb_pass_fn, dim3(h->jtb_grid), dim3(JTB_THREADS), args, h->jtb_smem, h->stream));
                h->launches += 1;
                S->warp_steps += (long long)h->K.nx * h->K.ny * m / 4;   // four-cell groups relaxed
            }
            t += m; src = dst; dst = (dst == o1) ? o2 : o1;
        }
    } else {
        SolveArgs a = slab_solve_args(h, k, k);
        const bool paired = h->ff_paired && !S->four_faces;   // the library's own fluxes (k_slab_sweep)
        // two sweeps per pass over HBM (k_slab_sweep2); an odd sweep left over runs alone.  One warp per (strip of 32 columns
        // of which it owns 32-2NB, chunk of rows), about one unit per resident warp, chunks of at least 32 rows.
        const int NBv = op == OP_QUICK ? 2 : 1, wv = 32 - 2 * NBv;
        const int strips = (h->K.ny + wv - 1) / wv;
        const int slots = S->sweep2_slots[op == OP_QUICK ? 1 : 0];
        int nchunks = std::max(1, std::min(slots / strips, h->K.nx / 32));
        if (S->sweep2_chunks > 0) nchunks = std::min(S->sweep2_chunks, std::max(1, h->K.nx / 32));
        const int RB = (h->K.nx + nchunks - 1) / nchunks;
        nchunks = (h->K.nx + RB - 1) / RB;
        const int units = strips * nchunks;
        for (int t = 0; t < nsw;) {
            const double* sp = slab_buf(h, k, src);
            double* dp = slab_buf(h, k, dst);
            const bool two = (S->sweep2 & (op == OP_QUICK ? 2 : 1)) && slots > 0 && nsw - t >= 2;
            if (two) {
                const int grid2 = (units + SW2_THREADS / 32 - 1) / (SW2_THREADS / 32);
#define SLAB_SWEEP2(OPv, Pv) k_slab_sweep2<OPv, Pv><<<grid2, SW2_THREADS, 0, h->stream>>>(a, sp, dp, S->own0, S->own1, RB, strips, units, S->sweep2_pf, S->sweep2_partials, S->sums + t, S->tickets + 1, done)
                if (op == OP_UPWIND) { if (paired) SLAB_SWEEP2(OP_UPWIND, true); else SLAB_SWEEP2(OP_UPWIND, false); }
                else { if (paired) SLAB_SWEEP2(OP_QUICK, true); else SLAB_SWEEP2(OP_QUICK, false); }
#undef SLAB_SWEEP2
                S->sweep2_passes += 1;
            } else {
                const int gx = (h->K.ny + SLAB_THREADS - 1) / SLAB_THREADS;
                const dim3 grid(gx, std::max(1, std::min(h->K.nx, (h->num_sms * 8 + gx - 1) / gx)));
#define SLAB_SWEEP(OPv, Pv, P2v) k_slab_sweep<OPv, Pv, P2v><<<grid, SLAB_THREADS, 0, h->stream>>>(a, sp, dp, S->own0, S->own1, S->sweep_partials, S->sums + t, S->tickets + 1, done)
                const bool p2 = paired && S->south_from_north;
                if (op == OP_UPWIND) { if (p2) SLAB_SWEEP(OP_UPWIND, true, true); else if (paired) SLAB_SWEEP(OP_UPWIND, true, false); else SLAB_SWEEP(OP_UPWIND, false, false); }
                else { if (p2) SLAB_SWEEP(OP_QUICK, true, true); else if (paired) SLAB_SWEEP(OP_QUICK, true, false); else SLAB_SWEEP(OP_QUICK, false, false); }
#undef SLAB_SWEEP
            }
            LAUNCH_CHECK(h);
            t += two ? 2 : 1;
            src = dst; dst = (dst == o1) ? o2 : o1;
        }
    }
    E = src;
    return SRCFD_OK;
}

static int slab_sync_verdict(SlabGroup& G) {
    for (int i = 0; i < G.n; ++i) {
        srcfd_handle* h = G[i];
        CK(cudaSetDevice(h->dev));
        CK(cudaMemcpyAsync(h->slab->sc_host, h->slab->sc, sizeof(SlabCtl), cudaMemcpyDeviceToHost, h->stream));
    }
    for (int i = 0; i < G.n; ++i) {
        srcfd_handle* h = G[i];
        CK(cudaSetDevice(h->dev));
        CK(cudaStreamSynchronize(h->stream));
        if (h->slab->sc_host->deadlock) return fail(SRCFD_ERR_DEADLOCK, "slab exchange: a neighbour's rows or sums did not arrive within the spin limit");
    }
    return SRCFD_OK;
}
static int slab_begin(SlabGroup& G, bool new_solve = false) {
    for (int i = 0; i < G.n; ++i) {
        CK(cudaSetDevice(G[i]->dev));
        if (new_solve) G[i]->slab->warp_steps = 0;
        k_slab_begin<<<1, 1, 0, G[i]->stream>>>(G[i]->slab->sc, new_solve ? 1 : 0);
        LAUNCH_CHECK(G[i]);
    }
    return SRCFD_OK;
}

// One inner solve (solve_pressure / solve_momentum_*, LDC.py:248-314) on plane k in JACOBI order over all slabs.
static int slab_inner_solve(SlabGroup& G, int op, int k, int slot, int* sweeps_out, double* rms_out) {
    srcfd_handle* h0 = G[0];
    SlabState* S0 = h0->slab;
    const int world = S0->world, halo = S0->halo;
    const int NB = op == OP_QUICK ? 2 : 1;
    const int max_iter = h0->p.inner_max;
    const double tol = h0->p.inner_tol;
    int SB;
    if (op == OP_PRESSURE) {
        const int H = 4;                                     // block sizes do not depend on which pressure kernel a rank runs
        SB = world > 1 ? (halo >= H ? (halo / H) * H : halo) : (SLAB_NS / H) * H;
    } else {
        SB = world > 1 ? std::min(SLAB_NS, (halo - 1) / NB) : SLAB_NS;
    }
    SB = std::min(SB, SLAB_NS);
    if (S0->block_cap > 0) SB = std::min(SB, S0->block_cap);
    if (SB < 1) return fail(SRCFD_ERR_ARG, "slab halo too thin for this stencil");
    TRY(slab_begin(G, true));
    for (int i = 0; i < G.n; ++i) {
        srcfd_handle* h = G[i];
        CK(cudaSetDevice(h->dev));
        k_slab_ghosts<<<(std::max(h->K.nx, h->K.ny) + 2 + 127) / 128, 128, 0, h->stream>>>(slab_buf(h, k, 0), h->scratch, h->scratch2, h->K, h->ctrl);
        LAUNCH_CHECK(h);
        h->jtb_ghosts_valid = false;
    }
    struct Rec { int S, E, nsw, n0; };
    std::vector<Rec> recs;
    int cur = 0, n = 0, b = 0;
    // Speculation.  The sweep count of an inner solve drifts slowly from one outer iteration to the next, so it is predicted
    // from the last two counts; full-size blocks run up to a little before the prediction, then SMALL blocks (one pass / two
    // momentum sweeps) up to a little past it: the block that meets the tolerance is then short, so its replay is at most a
    // few sweeps, and a count that grew is still covered without a host round trip (launches queued behind the hit are
    // no-ops).  Only a badly wrong prediction costs a second round, whose blocks double in size.
    const int last = S0->guess[slot], prev = S0->guess_prev[slot];
    int pred = last > 0 ? last : max_iter;
    if (last > 0 && prev > 0) pred = std::max(std::max(1, last / 2), std::min(2 * last, last + (last - prev)));
    pred = std::min(pred, max_iter);
    const int unit = std::min(SB, op == OP_PRESSURE ? 4 : 2);
    const int lo = op == OP_PRESSURE ? 8 : 0, hi = op == OP_PRESSURE ? 16 : 4;
    const bool at_cap = last <= 0 || last >= max_iter;        // the solve ran into the cap last time: whole blocks to the end
    int want = at_cap ? max_iter : std::min(max_iter, pred + hi);
    int fine = unit;
    int n_final = 0, final_buf = 0;
    double rms_final = 0.0;
    for (;;) {
        while (n < want) {
            int nsw;
            if (at_cap) nsw = std::min(SB, max_iter - n);
            else if (n < pred - lo) nsw = std::min(SB, pred - lo - n);
            else nsw = std::min(std::min(fine, SB), max_iter - n);
            int E = cur;
            for (int i = 0; i < G.n; ++i) TRY(slab_run_block(G[i], op, k, cur, nsw, E));
            const SlabPlanes pl = {k, E, 0, 0, 1};
            TRY(slab_exchange(G, pl, nsw, 1, b, nsw, tol));
            recs.push_back(Rec{cur, E, nsw, n});
            cur = E; n += nsw; ++b;
        }
        TRY(slab_sync_verdict(G));
        const SlabCtl& c = *S0->sc_host;
        if (c.hit) {
            const Rec& r = recs[(size_t)c.hit_block];
            rms_final = c.rms;
            n_final = r.n0 + c.hit_sweep + 1;
            if (c.hit_sweep == r.nsw - 1) {
                final_buf = r.E;
            } else {                                         // overshoot: replay the block from its intact start buffer
                TRY(slab_begin(G));
                int E = r.S;
                for (int i = 0; i < G.n; ++i) { TRY(slab_run_block(G[i], op, k, r.S, c.hit_sweep + 1, E)); G[i]->slab->replays += 1; }
                final_buf = E;
            }
            break;
        }
        if (n >= max_iter || h0->ctrl_host->stop) { n_final = n; final_buf = cur; rms_final = c.rms; break; }
        fine = std::min(SB, fine * 2);                       // the prediction was short: larger steps from here on
        want = std::min(max_iter, n + std::max(4 * fine, n / 4));
    }
    TRY(slab_begin(G));                                      // done = 0 again: the refresh below must run
    for (int i = 0; i < G.n; ++i) {
        srcfd_handle* h = G[i];
        CK(cudaSetDevice(h->dev));
        if (op == OP_PRESSURE && h->jtb_impl == 2) {         // which pressure kernel the NEXT solves of this rank use (see SlabState)
            SlabState* S = h->slab;
            const double frac = S->warp_steps > 0 ? (double)S->sc_host->retries / (double)S->warp_steps : 0.0;
            if (!S->use_tiles) { S->stream_solves += 1; if (frac > 1e-4 && !S->force_stream) S->use_tiles = true; }
            else { S->tile_solves += 1; if (frac < 1e-6) S->use_tiles = false; }
        }
        if (final_buf != 0)
            CK(cudaMemcpyAsync(slab_buf(h, k, 0), slab_buf(h, k, final_buf), sizeof(double) * (size_t)h->K.plane, cudaMemcpyDeviceToDevice, h->stream));
        k_slab_finish_inner<<<1, 1, 0, h->stream>>>(h->ctrl, slot, n_final, rms_final);
        LAUNCH_CHECK(h);
        h->slab->guess_prev[slot] = h->slab->guess[slot];
        h->slab->guess[slot] = n_final;
    }
    const SlabPlanes pl = {k, 0, 0, 0, 1};
    TRY(slab_exchange(G, pl, 0, 0, 0, 0, 0.0));              // halo rows of the accepted iterate
    if (sweeps_out) *sweeps_out = n_final;
    if (rms_out) *rms_out = rms_final;
    return SRCFD_OK;
}

// _implicit_solve (LDC.py:432-467 / BFS.py:622-673) over the slabs, JACOBI order.
static int slab_implicit_solve(SlabGroup& G) {
    const int mop = G[0]->p.scheme == SRCFD_SCHEME_QUICK ? OP_QUICK : OP_UPWIND;
    auto each = [&](auto&& f) -> int {
        for (int i = 0; i < G.n; ++i) { CK(cudaSetDevice(G[i]->dev)); if (int rc = f(G[i])) return rc; }
        return SRCFD_OK;
    };
    for (int k = 0; k < 2; ++k) {
        TRY(slab_inner_solve(G, mop, k, k, nullptr, nullptr));
        TRY(each([&](srcfd_handle* h) -> int {
            if (h->p.relax_enabled) TRY(l_under_relax(h, k, h->p.relax[k]));
            return l_apply_bc(h, k);
        }));
    }
    TRY(each([&](srcfd_handle* h) -> int { return l_linear_interpolation(h, true); }));
    TRY(slab_inner_solve(G, OP_PRESSURE, 2, 2, nullptr, nullptr));
    TRY(each([&](srcfd_handle* h) -> int {
        if (h->p.relax_enabled) TRY(l_under_relax(h, 2, h->p.relax[2]));
        TRY(l_apply_bc(h, 2));
        SlabState* S = h->slab;
        k_correct_velocity<<<h->tail_blocks, TAIL_THREADS, 0, h->stream>>>(h->Var, h->VarOld, h->res_partials, h->K, h->ctrl, S->own0, S->own1);
        LAUNCH_CHECK(h);
        k_residual_finish<<<1, 256, 0, h->stream>>>(h->res_partials, h->tail_blocks, h->ctrl, S->sums, 0);
        LAUNCH_CHECK(h);
        return SRCFD_OK;
    }));
    const SlabPlanes uv = {0, 0, 1, 0, 2};
    TRY(slab_exchange(G, uv, 3, 2, 0, 0, 0.0));              // corrected u, v rows + the three residual sums (LDC.py:326-328)
    TRY(each([&](srcfd_handle* h) -> int {
        TRY(l_apply_bc(h, 0));
        TRY(l_apply_bc(h, 1));
        TRY(l_update_flux(h));
        h->ghosts_fresh = true;
        return SRCFD_OK;
    }));
    return SRCFD_OK;
}

extern "C" {

int srcfd_slab_configure(srcfd_handle* h, int world, int rank, int nx_global, int halo) {
    CKH(h);
    if (h->slab) return fail(SRCFD_ERR_ARG, "handle is already configured as a slab");
    if (world < 1 || world > SLAB_MAX_WORLD || rank < 0 || rank >= world) return fail(SRCFD_ERR_ARG, "bad world/rank");
    if (h->p.sweep_order != SRCFD_ORDER_JACOBI) return fail(SRCFD_ERR_ARG, "slab handles run the JACOBI order");
    const int lo = rank > 0 ? halo : 0, hi = rank < world - 1 ? halo : 0;
    const int n_own = h->p.nx - lo - hi;
    if (world > 1 && halo < 3) return fail(SRCFD_ERR_ARG, "halo must be >= 3 rows");
    if (n_own < std::max(1, world > 1 ? halo : 1)) return fail(SRCFD_ERR_ARG, "local grid thinner than its halos (nx = halo_lo + owned + halo_hi, owned >= halo)");
    if (nx_global < n_own) return fail(SRCFD_ERR_ARG, "bad nx_global");
    SlabState* S = new SlabState();
    h->slab = S;
    S->world = world; S->rank = rank; S->nx_global = nx_global; S->halo = world > 1 ? halo : 0;
    S->lo = lo; S->hi = hi; S->own0 = lo + 1; S->own1 = lo + n_own;
    S->guess[2] = h->p.inner_max;
    if (const char* e = getenv("SRCFD_SLAB_BLOCK")) S->block_cap = atoi(e);
    if (const char* e = getenv("SRCFD_SLAB_FOUR_FACES")) S->four_faces = atoi(e) != 0;
    if (const char* e = getenv("SRCFD_JTB2_NL")) S->stream_nl = std::max(1, std::min(4, atoi(e)));
    if (const char* e = getenv("SRCFD_JTB2_FORCE")) { S->force_stream = atoi(e) != 0; if (S->force_stream) S->use_tiles = false; }   // no probe solve either
    S->mail_bytes = slab_mail_bytes(std::max(1, S->halo), h->K.pitch);
    auto bail = [&](int rc) { std::string keep = g_err; slab_release(h); g_err = keep; return rc; };
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(SRCFD_ERR_CUDA); } } while (0)
    CKS(cudaMalloc(&S->mail, S->mail_bytes));
    CKS(cudaMemsetAsync(S->mail, 0, S->mail_bytes, h->stream));
    CKS(cudaMalloc(&S->sc, sizeof(SlabCtl)));
    CKS(cudaMemsetAsync(S->sc, 0, sizeof(SlabCtl), h->stream));
    CKS(cudaMallocHost(&S->sc_host, sizeof(SlabCtl)));
    memset(S->sc_host, 0, sizeof(SlabCtl));
    CKS(cudaMalloc(&S->sums, sizeof(double) * SLAB_NS));
    CKS(cudaMemsetAsync(S->sums, 0, sizeof(double) * SLAB_NS, h->stream));
    CKS(cudaMalloc(&S->sweep_partials, sizeof(double) * ((size_t)((h->K.ny + SLAB_THREADS - 1) / SLAB_THREADS) * h->K.nx + 1)));
    if (const char* e = getenv("SRCFD_SLAB_SWEEP2")) S->sweep2 = atoi(e) != 0 ? 3 : 0;
    if (const char* e = getenv("SRCFD_SWEEP2_CHUNKS")) S->sweep2_chunks = atoi(e);
    if (const char* e = getenv("SRCFD_SWEEP2_PF")) S->sweep2_pf = std::max(0, atoi(e));
    if (const char* e = getenv("SRCFD_SLAB_SOUTH_FROM_NORTH")) S->south_from_north = atoi(e) != 0;
    {   // k_slab_sweep2: resident warps per device (unit count of a pass) and room for two partial sums per unit
        int occ_u = 0, occ_q = 0;
        CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, k_slab_sweep2<OP_UPWIND, true>, SW2_THREADS, 0));
        CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_q, k_slab_sweep2<OP_QUICK, true>, SW2_THREADS, 0));
        S->sweep2_slots[0] = h->num_sms * occ_u * (SW2_THREADS / 32);
        S->sweep2_slots[1] = h->num_sms * occ_q * (SW2_THREADS / 32);
        const size_t units_max = (size_t)((h->K.ny + 27) / 28) * (size_t)(h->K.nx / 32 + 1);
        CKS(cudaMalloc(&S->sweep2_partials, sizeof(double) * 2 * units_max));
    }
    CKS(cudaMalloc(&S->tickets, sizeof(unsigned) * 4));
    CKS(cudaMemsetAsync(S->tickets, 0, sizeof(unsigned) * 4, h->stream));
    CKS(cudaStreamSynchronize(h->stream));
#undef CKS
    S->peer_base[rank] = S->mail;
    h->bc.skip_lo = lo > 0; h->bc.skip_hi = hi > 0;
    return SRCFD_OK;
}

int srcfd_slab_kernel_stats(srcfd_handle* h, int64_t* stream_solves, int64_t* tile_solves, int64_t* last_retries, int64_t* last_warp_steps) {
    CKH(h);
    if (!h->slab) return fail(SRCFD_ERR_ARG, "not a slab");
    if (stream_solves) *stream_solves = h->slab->stream_solves;
    if (tile_solves) *tile_solves = h->slab->tile_solves;
    if (last_retries) *last_retries = (int64_t)h->slab->sc_host->retries;
    if (last_warp_steps) *last_warp_steps = h->slab->warp_steps;
    return SRCFD_OK;
}

int srcfd_slab_info(srcfd_handle* h, int32_t* own_row0, int32_t* own_row1, int64_t* exchanges, int64_t* halo_bytes, int64_t* replays) {
    CKH(h);
    if (!h->slab) return fail(SRCFD_ERR_ARG, "not a slab");
    if (own_row0) *own_row0 = h->slab->own0;
    if (own_row1) *own_row1 = h->slab->own1;
    if (exchanges) *exchanges = h->slab->exchanges;
    if (halo_bytes) *halo_bytes = h->slab->halo_bytes;
    if (replays) *replays = h->slab->replays;
    return SRCFD_OK;
}

int srcfd_slab_export(srcfd_handle* h, void* blob, int blob_bytes) {
    CKH(h);
    if (!h->slab) return fail(SRCFD_ERR_ARG, "not a slab");
    if (!blob || blob_bytes < (int)sizeof(cudaIpcMemHandle_t)) return fail(SRCFD_ERR_ARG, "blob must hold SRCFD_SLAB_BLOB_BYTES bytes");
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, h->slab->mail));
    memcpy(blob, &hd, sizeof(hd));
    return SRCFD_OK;
}

int srcfd_slab_attach_ipc(srcfd_handle* h, int peer_rank, const void* blob) {
    CKH(h);
    SlabState* S = h->slab;
    if (!S) return fail(SRCFD_ERR_ARG, "not a slab");
    if (peer_rank < 0 || peer_rank >= S->world || peer_rank == S->rank || !blob) return fail(SRCFD_ERR_ARG, "bad peer rank / blob");
    if (S->peer_base[peer_rank]) return fail(SRCFD_ERR_ARG, "peer already attached");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, blob, sizeof(hd));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    S->peer_base[peer_rank] = p; S->peer_ipc[peer_rank] = true;
    return SRCFD_OK;
}

int srcfd_slab_attach_local(srcfd_handle* h, int peer_rank, srcfd_handle* peer) {
    CKH(h);
    SlabState* S = h->slab;
    if (!S || !peer || !peer->slab) return fail(SRCFD_ERR_ARG, "not a slab");
    if (peer_rank < 0 || peer_rank >= S->world || peer_rank == S->rank) return fail(SRCFD_ERR_ARG, "bad peer rank");
    if (peer->slab->rank != peer_rank || peer->slab->world != S->world || peer->slab->halo != S->halo || peer->K.pitch != h->K.pitch)
        return fail(SRCFD_ERR_ARG, "peer slab does not match (rank, world, halo, ny)");
    if (peer->dev != h->dev) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->dev, peer->dev));
        if (!can) return fail(SRCFD_ERR_CUDA, "devices cannot access each other's memory");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SRCFD_ERR_CUDA, cudaGetErrorString(e));
        cudaGetLastError();
    }
    S->peer_base[peer_rank] = peer->slab->mail; S->peer_ipc[peer_rank] = false;
    return SRCFD_OK;
}

int srcfd_slab_exchange(srcfd_handle* const* hs, int n, int k) {
    TRY(slab_check_group(hs, n));
    if (k < 0 || k > 2) return fail(SRCFD_ERR_ARG, "k must be 0, 1 or 2");
    SlabGroup G(hs, n);
    TRY(slab_begin(G));
    const SlabPlanes pl = {k, 0, 0, 0, 1};
    TRY(slab_exchange(G, pl, 0, 0, 0, 0, 0.0));
    return slab_sync_verdict(G);
}

int srcfd_slab_solve_pressure(srcfd_handle* const* hs, int n, int32_t* sweeps, double* last_rms) {
    TRY(slab_check_group(hs, n));
    SlabGroup G(hs, n);
    for (int i = 0; i < n; ++i) { CK(cudaSetDevice(G[i]->dev)); TRY(l_clear_stop(G[i])); TRY(l_pressure_rhs(G[i])); }
    int sw = 0; double rms = 0.0;
    TRY(slab_inner_solve(G, OP_PRESSURE, 2, 2, &sw, &rms));
    TRY(slab_sync_verdict(G));
    if (sweeps) *sweeps = sw;
    if (last_rms) *last_rms = rms;
    return SRCFD_OK;
}

int srcfd_slab_solve_momentum(srcfd_handle* const* hs, int n, int k, int scheme, int32_t* sweeps, double* last_rms) {
    TRY(slab_check_group(hs, n));
    if (k < 0 || k > 1) return fail(SRCFD_ERR_ARG, "momentum is solved for k = 0 (u) or 1 (v)");
    if (scheme != SRCFD_SCHEME_UPWIND && scheme != SRCFD_SCHEME_QUICK) return fail(SRCFD_ERR_ARG, "bad scheme");
    SlabGroup G(hs, n);
    for (int i = 0; i < n; ++i) { CK(cudaSetDevice(G[i]->dev)); TRY(l_clear_stop(G[i])); G[i]->ghosts_fresh = false; }
    int sw = 0; double rms = 0.0;
    TRY(slab_inner_solve(G, scheme == SRCFD_SCHEME_QUICK ? OP_QUICK : OP_UPWIND, k, k, &sw, &rms));
    TRY(slab_sync_verdict(G));
    if (sweeps) *sweeps = sw;
    if (last_rms) *last_rms = rms;
    return SRCFD_OK;
}

// n_outer x (_implicit_solve + _convergence_check + copy_new_to_old); stops early on convergence / NaN like srcfd_step.
int srcfd_slab_step(srcfd_handle* const* hs, int n, int64_t n_outer, const double crit[3]) {
    TRY(slab_check_group(hs, n));
    if (!crit) return fail(SRCFD_ERR_ARG, "null crit");
    SlabGroup G(hs, n);
    for (int i = 0; i < n; ++i) G[i]->maybe_stopped = true;
    for (int64_t it = 0; it < n_outer; ++it) {
        TRY(slab_implicit_solve(G));
        for (int i = 0; i < n; ++i) {
            srcfd_handle* h = G[i];
            CK(cudaSetDevice(h->dev));
            Consts Kg = h->K;
            Kg.nx = h->slab->nx_global;                      // rms = sqrt(residual / (Nx*Ny)) / dt over the whole domain
            k_convergence_check<<<1, 1, 0, h->stream>>>(h->ctrl, h->hist, Kg, crit[0], crit[1], crit[2]);
            LAUNCH_CHECK(h);
            TRY(l_copy_new_to_old(h));
        }
        bool stop = false;
        for (int i = 0; i < n; ++i) {
            CK(cudaSetDevice(G[i]->dev));
            TRY(fetch_ctrl(G[i]));
            if (int rc = ctrl_verdict(G[i])) return rc;
            stop = stop || G[i]->ctrl_host->stop;
        }
        if (stop) break;
    }
    return SRCFD_OK;
}

}  // extern "C"
