// 2D engine: context, Krylov solver, Newton driver, time loop, adjoint sweep, PGD iteration, C ABI.
// Host control flow lives here; all arithmetic is in the kernels of vch2d_kernels.cuh / vch2d_tiles.cuh / vch_dct.cuh / vch_fft16.cuh.
#include "vch2d_tiles.cuh"
#include <algorithm>
#include <functional>
#include <chrono>
#include <cstdlib>

namespace vch {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }

}  // namespace vch

using namespace vch;

struct vch2d_ctx {
    vch2d_params prm;
    Geo g;
    Phys ph;
    int device = 0;
    cudaStream_t stream = nullptr;   // private work stream: every kernel of the library runs here (capturable)
    cudaStream_t user = nullptr;     // caller's stream (vch2d_set_stream); ordered with `stream` by events per call
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    int use_graphs = 1;
    struct SolveGraph { bool adj; const double* a; const double* rphi; const double* rmu; cudaGraph_t g; cudaGraphExec_t exec; };
    std::vector<SolveGraph> graphs;
    LaunchLog log;
    double krylov_tol = 1e-11;
    double krylov_first_tol = 1e-6;   // time loop only: relative tolerance of the FIRST linear solve of a Newton solve (0 = krylov_tol)
    int krylov_maxit = 200;
    int floor_aware = 1;         // fp64-resolution-aware Newton stop
    int bicg6 = 1;               // 6-launch BiCGStab iteration (x/r update deferred into the next row transform, scalars from the second
                                 // epilogue; measured +6.6 % it/s).  0 (VCH_BICG6=0, slab mode, dense-table grids): the 7-launch form
    int pdl = 0;                 // programmatic dependent launch for every kernel of the work stream (vch_common.cuh); VCH_PDL=0|1
    int half_exit = 1;           // forward BiCGStab may stop after the first half of an iteration (VCH_NO_HALF_EXIT=1 disables).
                                 // Not used for the adjoint: its tolerance is on the TRUE residual, whose error is ~1e3 larger;
                                 // the second half of its last iteration is what keeps the gradient at ~1e-12 of the reference's
                                 // (measured: r moves by 6.6e-9 with the exit, for 6.6 % of the adjoint sweep's time)
    int debug = 0;
    int tiled = 0;               // shared-memory tile kernels for the stencils (vch2d_tiles.cuh); not in slab mode.  VCH_TILED=0 disables
    int fused_solve = 0;         // tiled + radix-16 transforms + 6-launch iteration: the Schur right-hand side is formed by the first row
                                 // transform of the solve, the BiCGStab start by the last one of P^-1 b (forward) / by the adjoint rhs
                                 // kernel, the closing x update by the dmu kernel — 3 launches fewer per linear solve
    Tiling tl;
    const double* cur_rphi = nullptr;   // fused forward solve: R_phi / R_mu the solve's first kernel reads
    const double* cur_rmu = nullptr;
    DctPlan dct;
    // work vectors (n doubles each)
    DevBuf phi, mu, phit, mut, w0, w1, cphi, cmu, Rphi, Rmu, a, RphiT, RmuT, aT;
    DevBuf kb, kx, kr, kr0, kp, kv, ks, kt, ktmp, kq, dmu;
    DevBuf adj_p[2], adj_q[2], adj_r[2], mu_old;
    DevBuf ck_seg, ck_r;         // checkpointed sweep: phi levels of the segment being recomputed, ring of r levels
    DevBuf stage[7];             // device staging of the host-buffer PGD iteration, kept across calls (no 50 GB malloc/free per step)
    long long stream_budget = 0; // bytes of device staging the host-buffer path may use; 0 = whatever fits (vch2d_set_stream_budget)
    std::vector<double> r_host;  // host copy of the gradient trajectory for the bounded-memory path when the caller gives no r_out
    RedBuf red;                  // partials for grid reductions (+ the Comm header the kernels read)
    // slab mode (one rank of a row-slab decomposition of a square grid; see vch_common.cuh "slab mode")
    bool slab = false;
    Comm cm;                     // rank 0 of 1 unless slab
    Arena arena;                 // every work vector lives here (ghost margins of 2 rows), peers address it over NVLink
    int rows_lo = 0;             // rows owned by each rank but the last (the last owns rows_lo + 1)
    std::vector<void*> peer_maps;   // cudaIpcOpenMemHandle mappings
    bool attached = false;
    DevBuf small;                // small device vectors: weights, out4
    unsigned int* ticket = nullptr;
    Scal* sc = nullptr;          // device scalars
    Scal* sc_host = nullptr;     // pinned mirror
    double* out4 = nullptr;      // device, 8 doubles
    double* out4_host = nullptr; // pinned

    // Newton update dphi after newton_linear_solve: the fused dmu kernel writes the closed iterate into kt (see dmu_close_tile_kernel)
    double* dphi() const { return fused_solve ? kt.p : kx.p; }
    int eb() const { return (int)((g.n + 255) / 256); }
    int rb() const { return red_blocks(g.n); }
};

namespace {

#define LAUNCH(c, kern, grid, block, ...)                       \
    do {                                                        \
        (c)->log.begin(#kern, (c)->stream);                     \
        launch_pdl((c)->pdl != 0, kern, dim3(grid), dim3(block), 0, (c)->stream, __VA_ARGS__); \
        (c)->log.end((c)->stream);                              \
    } while (0)

// Device scalars -> pinned host mirror, written by a kernel straight into the (device-mapped) pinned page.  Not a
// cudaMemcpyAsync: a 200-byte D2H memcpy is served by the D2H copy engine, where it queues behind the 270 MB trajectory chunks
// of the streamed host-buffer path — once per Newton iteration (measured with a background D2H stream: 58 ms instead of
// 1.3 ms per time step).
__global__ void publish_scalars_kernel(const Scal* __restrict__ src, Scal* __restrict__ dst_host) {
    pdl_enter();
    static_assert(sizeof(Scal) % 8 == 0, "Scal is copied in 8-byte words");
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(dst_host);
    for (int i = threadIdx.x; i < (int)(sizeof(Scal) / 8); i += blockDim.x) d[i] = s[i];
}

// Waits for scalars that the last kernel on the stream has published itself (residual_kernel with publish = sc_host).
void wait_scalars(vch2d_ctx* c) {
    VCH_CUDA(cudaStreamSynchronize(c->stream));
    if (c->sc_host->comm_err) throw Error(VCH_E_COMM, "slab mode: a peer rank did not arrive within the wait limit (ranks out of step or a peer failed)");
}
void fetch_scalars(vch2d_ctx* c) {
    LAUNCH(c, publish_scalars_kernel, 1, 32, c->sc, c->sc_host);
    wait_scalars(c);
}


// Device-to-device field copy by a kernel (see copy_kernel).
void dev_copy(vch2d_ctx* c, double* dst, const double* src, size_t count) {
    if (dst == src || count == 0) return;
    LAUNCH(c, copy_kernel, red_blocks((long long)count), kRedThreads, src, dst, (long long)count);
}

// Slab mode: ship the first / last `rows` owned rows of one or two fields into the neighbours' ghost rows; a barrier inside
// the kernel guarantees the rows have landed before anything reads them (lead = 1 adds a barrier in front, see
// halo_push_kernel).  No-op outside slab mode.
void halo_push(vch2d_ctx* c, const double* f0, const double* f1, int rows, const int* done = nullptr, int lead = 0) {
    if (!c->slab) return;
    for (const double* f : {f0, f1})
        if (f && (f < c->arena.base + kArenaHeader || f >= c->arena.base + c->arena.cap))
            throw Error(VCH_E_ARG, "halo_push: field is not in the slab arena");
    LAUNCH(c, halo_push_kernel, 1, 1024, c->cm, f0, f1, rows, c->g.no, c->rows_lo, c->g.ni, done, lead);
}
// Copy of a ghosted work vector including its ghost rows.
void copy_ghosted(vch2d_ctx* c, double* dst, const double* src) {
    const size_t m = c->slab ? 2 * (size_t)c->g.ni : 0;
    dev_copy(c, dst - m, src - m, (size_t)c->g.n + 2 * m);
}
// kernels per BiCGStab iteration / per solve prologue (accounting of launches inside solve graphs)
template <bool ADJ> int iter_launches(const vch2d_ctx* c) { return (c->bicg6 ? 6 : 7) + (c->slab ? 4 : 0); }
// FWD: P^-1 b (3 kernels) + init;  ADJ (right-preconditioned): init + the closing x = P^-1 y (3 kernels); slab mode adds the
// barriers of one preconditioner application (2 + the trailing one)
template <bool ADJ> int prologue_launches(const vch2d_ctx* c) {
    if (c->fused_solve) return ADJ ? 4 : 3;      // FWD: P^-1 b with the start fused;  ADJ: closing update + x = P^-1 y (start: adjoint rhs kernel)
    return 4 + (c->slab ? 3 : 0) + (c->bicg6 ? 1 : 0);
}

// Orders the library's private stream after the caller's stream on entry and the caller's stream after ours on exit,
// so callers see ordinary stream semantics (torch.cuda.Event on their stream brackets our kernels).
struct StreamScope {
    vch2d_ctx* c;
    explicit StreamScope(vch2d_ctx* ctx) : c(ctx) {
        VCH_CUDA(cudaSetDevice(c->device));
        VCH_CUDA(cudaEventRecord(c->ev_in, c->user));
        VCH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
    }
    ~StreamScope() {
        cudaEventRecord(c->ev_out, c->stream);
        cudaStreamWaitEvent(c->user, c->ev_out, 0);
    }
};

// One BiCGStab iteration, stencil-free for both operators (7 launches):
//   Forward Schur operator (ADJ = false), LEFT-preconditioned:  A = P - L diag(a - abar), so
//     P^-1 A x = x + DCT^-1[(lambda/sym) DCT((a - abar) x)]      (coefficient multiply fused into the row prologue)
//   Adjoint operator (ADJ = true), RIGHT-preconditioned:        A = P - diag(a - abar) L, so
//     A P^-1 y = y + (a - abar) DCT^-1[(lambda/sym) DCT(y)]      (coefficient multiply fused into the row epilogue);
//     the iteration runs on y with the TRUE residual b - A x, and x = P^-1 y closes the solve.
//   rows[p = r + beta q (; (a-abar)p)] -> cols[lambda/sym] -> rows[(a-abar)z + p, (r0,v)] -> rows[s = r - alpha v ...] ->
//   cols -> rows[... + s, (t,s),(t,t)] -> x/r/q update with (r,r),(r0,r).
template <bool ADJ>
void enqueue_bicg_iteration(vch2d_ctx* c, const double* a, const SymbolArgs& sy, cudaGraphConditionalHandle cond, int use_cond) {
    const long long n = c->g.n;
    const int rb = c->rb();
    const int* done = &c->sc->done;
    const double* pro_a = ADJ ? nullptr : a;     // multiply before the transform (forward) ...
    const double* epi_a = ADJ ? a : nullptr;     // ... or after it (adjoint)
    if (c->bicg6) {   // rows[x/r update of the previous iteration, p] cols rows[v, alpha] rows[s] cols rows[t, omega, rho, (r,r), stop]
        DotEpilogue e1{1, c->kr0.p, c->sc, c->red.part, c->ticket, c->kp.p, epi_a, (c->half_exit && !ADJ) ? c->kr.p : nullptr};
        e1.cond = cond; e1.use_cond = use_cond;
        RowPrologue p3{3, c->kr.p, c->kv.p, pro_a, c->kp.p, c->sc};
        p3.s = c->ks.p; p3.t = c->kt.p; p3.x = c->kx.p; p3.rw = c->kr.p;
        c->dct.apply(c->stream, c->kr.p, c->kv.p, sy, done, e1, p3, 1);
        DotEpilogue e2{4, c->ks.p, c->sc, c->red.part, c->ticket, c->ks.p, epi_a, c->kr0.p};
        e2.cond = cond; e2.use_cond = use_cond;
        c->dct.apply(c->stream, c->kr.p, c->kt.p, sy, done, e2, RowPrologue{2, c->kr.p, c->kv.p, pro_a, c->ks.p, c->sc}, 1);
        (void)n; (void)rb;
        return;
    }
    c->dct.apply(c->stream, c->kr.p, c->kv.p, sy, done, DotEpilogue{1, c->kr0.p, c->sc, c->red.part, c->ticket, c->kp.p, epi_a, (c->half_exit && !ADJ) ? c->kr.p : nullptr},
                 RowPrologue{1, c->kr.p, c->kq.p, pro_a, c->kp.p, c->sc}, 1);
    c->dct.apply(c->stream, c->kr.p, c->kt.p, sy, done, DotEpilogue{2, c->ks.p, c->sc, c->red.part, c->ticket, c->ks.p, epi_a},
                 RowPrologue{2, c->kr.p, c->kv.p, pro_a, c->ks.p, c->sc}, 1);
    LAUNCH(c, bicg_x_kernel, rb, kRedThreads, c->kx.p, c->kr.p, c->kp.p, c->ks.p, c->kt.p, c->kr0.p, c->kv.p, c->kq.p, n, c->sc,
           c->red.part, c->ticket, cond, use_cond);
}

// Start of a solve: FWD r = P^-1 b; ADJ r = b.  Then r0 = r, x = q = 0 and the norms.
template <bool ADJ>
void enqueue_bicg_prologue(vch2d_ctx* c, const SymbolArgs& sy, cudaGraphConditionalHandle cond, int use_cond) {
    const long long n = c->g.n;
    if (c->fused_solve) {
        if (ADJ) return;     // adj_rhs_tile_kernel has set r = r0 = b, x = 0 and the scalars
        RowPrologue p4; p4.mode = 4; p4.r = c->cur_rphi; p4.qv = c->cur_rmu; p4.k_in = c->g.ihi2; p4.k_out = c->g.iho2;
        DotEpilogue e5; e5.mode = 5; e5.sc = c->sc; e5.part = c->red.part; e5.ticket = c->ticket; e5.out2 = c->kr0.p; e5.zero = c->kx.p;
        e5.cond = cond; e5.use_cond = use_cond;
        c->dct.apply(c->stream, c->cur_rphi, c->kr.p, sy, nullptr, e5, p4);     // r = r0 = P^-1 (L R_phi - R_mu), x = 0, ||r||^2
        return;
    }
    if (!ADJ) c->dct.apply(c->stream, c->kb.p, c->kr.p, sy, nullptr);
    LAUNCH(c, bicg_init_kernel, c->rb(), kRedThreads, ADJ ? c->kb.p : c->kr.p, c->kr.p, c->kr0.p, c->kx.p, n, c->sc,
           c->red.part, c->ticket, cond, use_cond);
}
// End of a solve: ADJ x = P^-1 y.
template <bool ADJ>
void enqueue_bicg_epilogue(vch2d_ctx* c, const SymbolArgs& sy) {
    if (c->fused_solve && !ADJ) return;     // dmu_close_tile_kernel applies the closing update
    if (c->bicg6) LAUNCH(c, bicg_close_kernel, c->rb(), kRedThreads, c->kx.p, c->kp.p, c->ks.p, c->g.n, c->sc, c->red.part, c->ticket);
    if (ADJ) c->dct.apply(c->stream, c->kx.p, c->kx.p, sy, nullptr);
}

// Whole linear solve as ONE CUDA graph: [P^-1 b, init] -> WHILE(not converged){ BiCGStab iteration } — the loop condition
// is set on the device by the last kernel of each iteration (cudaGraphSetConditional), so the host never polls.
template <bool ADJ>
cudaGraphExec_t solve_graph(vch2d_ctx* c, const double* a) {
    const double* key_rp = (c->fused_solve && !ADJ) ? c->cur_rphi : nullptr;
    const double* key_rm = (c->fused_solve && !ADJ) ? c->cur_rmu : nullptr;
    for (auto& g : c->graphs)
        if (g.adj == ADJ && g.a == a && g.rphi == key_rp && g.rmu == key_rm) return g.exec;
    if (c->graphs.size() >= 12) {
        for (auto& g : c->graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.g); }
        c->graphs.clear();
    }
    SymbolArgs sy{0.0, 0.0, &c->sc->abar, 0.0, &c->sc->c0};
    const long long count0 = c->log.count;
    cudaGraph_t graph;
    VCH_CUDA(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle cond;
    VCH_CUDA(cudaGraphConditionalHandleCreate(&cond, graph, 1, cudaGraphCondAssignDefault));
    // prologue nodes
    std::vector<cudaGraphNode_t> leaf;
    cudaGraph_t tmp;
    if (!(c->fused_solve && ADJ)) {   // (fused adjoint solve: no prologue, the graph starts with the WHILE node)
        VCH_CUDA(cudaStreamBeginCaptureToGraph(c->stream, graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        enqueue_bicg_prologue<ADJ>(c, sy, cond, prologue_launches<ADJ>(c));
        cudaStreamCaptureStatus st; const cudaGraphNode_t* deps = nullptr; size_t ndeps = 0;
        VCH_CUDA(cudaStreamGetCaptureInfo(c->stream, &st, nullptr, nullptr, &deps, &ndeps));
        leaf.assign(deps, deps + ndeps);
        VCH_CUDA(cudaStreamEndCapture(c->stream, &tmp));
    }
    // WHILE node
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t wnode;
    VCH_CUDA(cudaGraphAddNode(&wnode, graph, leaf.empty() ? nullptr : leaf.data(), leaf.size(), &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    VCH_CUDA(cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    enqueue_bicg_iteration<ADJ>(c, a, sy, cond, iter_launches<ADJ>(c));
    VCH_CUDA(cudaStreamEndCapture(c->stream, &tmp));
    if (ADJ || (c->bicg6 && !c->fused_solve)) {   // closing x = P^-1 y (and, 6-launch form, the last x update), after the loop
        VCH_CUDA(cudaStreamBeginCaptureToGraph(c->stream, graph, &wnode, nullptr, 1, cudaStreamCaptureModeThreadLocal));
        enqueue_bicg_epilogue<ADJ>(c, sy);
        VCH_CUDA(cudaStreamEndCapture(c->stream, &tmp));
    }
    cudaGraphExec_t exec;
    VCH_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    c->log.count = count0;            // capture launches nothing
    c->graphs.push_back({ADJ, a, key_rp, key_rm, graph, exec});
    return exec;
}

// Left-preconditioned BiCGStab on  P^-1 A x = P^-1 b  with A = c0 I - {L diag(a) | diag(a) L} + c2 L^2 and
// P the same operator with a replaced by the device scalar abar.  Scalars stay on the device.
//   graph path (default): one graph launch, no host polling; iteration statistics accumulate in Scal.
//   polled path (profiling, VCH_NO_GRAPHS=1): kernels gated on sc->done, enqueued in batches, host polls per batch.
// Result in c->kx.  Returns the iteration count, or -1 when it is not known yet (graph path, read later from Scal).
template <bool ADJ>
int krylov_solve(vch2d_ctx* c, const double* b, const double* a, vch_stats* st) {   // coefficients: sc->c0 / sc->c2, set by the rhs kernel
    const long long n = c->g.n;
    if (b && b != c->kb.p) dev_copy(c, c->kb.p, b, (size_t)n);      // b == nullptr (fused solve): the right-hand side is already in place
    if (c->use_graphs && !c->log.profiling) {
        cudaGraphExec_t exec = solve_graph<ADJ>(c, a);
        VCH_CUDA(cudaGraphLaunch(exec, c->stream));
        return -1;
    }
    SymbolArgs sy{0.0, 0.0, &c->sc->abar, 0.0, &c->sc->c0};
    enqueue_bicg_prologue<ADJ>(c, sy, (cudaGraphConditionalHandle)0, 0);
    int launched = 0, batch = 2;
    while (true) {
        for (int k = 0; k < batch; ++k) enqueue_bicg_iteration<ADJ>(c, a, sy, (cudaGraphConditionalHandle)0, 0);
        launched += batch;
        fetch_scalars(c);
        if (c->sc_host->done || launched >= c->krylov_maxit) break;
        batch = (launched < 8) ? 2 : 4;
    }
    enqueue_bicg_epilogue<ADJ>(c, sy);
    VCH_CUDA(cudaGetLastError());
    if (!c->sc_host->done) {   // mirror the device-side accounting of the graph path
        long long one = c->sc_host->stalls + 1;
        VCH_CUDA(cudaMemcpyAsync(&c->sc->stalls, &one, sizeof(one), cudaMemcpyHostToDevice, c->stream));
        long long adj1 = c->sc_host->stalls_adj + 1;
        if (ADJ) VCH_CUDA(cudaMemcpyAsync(&c->sc->stalls_adj, &adj1, sizeof(adj1), cudaMemcpyHostToDevice, c->stream));
        VCH_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (c->sc_host->nonfinite) throw Error(VCH_E_NONFINITE, "non-finite value in Krylov solve");
    (void)st;
    return c->sc_host->iters;
}

// Device-side solver counters -> vch_stats (call after a fetch_scalars).
struct StatMark { long long its, solves, stalls, launches, glaunches, halves, stalls_adj; };
StatMark stat_mark(vch2d_ctx* c) {
    fetch_scalars(c);
    VCH_CUDA(cudaMemsetAsync(&c->sc->iters_max, 0, sizeof(int), c->stream));   // per-call maximum
    // Sticky state of an EARLIER call must not leak into this one: a non-finite event (NaN control, overflowing trial step)
    // raised VCH_E_NONFINITE there; contexts are cached and reused by the Python drop-ins, so the flag is cleared at every
    // API entry (with the half-step marker a solve aborted mid-iteration could have left behind).
    VCH_CUDA(cudaMemsetAsync(&c->sc->nonfinite, 0, sizeof(int), c->stream));
    VCH_CUDA(cudaMemsetAsync(&c->sc->half, 0, sizeof(int), c->stream));
    return {c->sc_host->iters_total, c->sc_host->solves, c->sc_host->stalls, c->log.count, c->sc_host->g_launches,
            c->sc_host->half_exits, c->sc_host->stalls_adj};
}
void stat_collect(vch2d_ctx* c, const StatMark& m0, vch_stats* st) {
    fetch_scalars(c);
    if (c->sc_host->nonfinite) throw Error(VCH_E_NONFINITE, "non-finite value in Krylov solve");
    if (!st) return;
    const long long its = c->sc_host->iters_total - m0.its, solves = c->sc_host->solves - m0.solves;
    st->krylov_iterations += its;
    st->newton_linear_solves += solves;
    st->krylov_stalls += c->sc_host->stalls - m0.stalls;
    st->krylov_stalls_adjoint += c->sc_host->stalls_adj - m0.stalls_adj;
    st->krylov_half_exits += c->sc_host->half_exits - m0.halves;
    if (c->sc_host->stalls > m0.stalls)    // never silent: the direct solver this replaces cannot stall
        fprintf(stderr, "[vch_b200] warning: %lld linear solve(s) stopped above the Krylov tolerance (max_iter %d); see vch_stats.krylov_stalls\n",
                (long long)(c->sc_host->stalls - m0.stalls), c->krylov_maxit);
    st->krylov_max_iterations = std::max<long long>(st->krylov_max_iterations, c->sc_host->iters_max);
    // kernels inside solve graphs are not seen by the launch log: 4 prologue kernels per solve + 11 per iteration
    st->kernel_launches += (c->log.count - m0.launches) + (c->sc_host->g_launches - m0.glaunches);
}

// The forward Newton loop re-checks the true residual after every linear solve, so a stalled BiCGStab there only costs an
// iteration.  The adjoint recurrence has no such outer check: a solve that stopped above tolerance leaves an inaccurate p_n
// that propagates into every earlier level and into the gradient.  The reference's direct solve cannot fail this way, so it
// is an error here (VCH_E_KRYLOV -> KrylovStall in Python), raised after the call's outputs have been written.
void require_adjoint_converged(vch2d_ctx* c, const StatMark& m0) {
    const long long bad = c->sc_host->stalls_adj - m0.stalls_adj;
    if (bad > 0)
        throw Error(VCH_E_KRYLOV, "adjoint sweep: " + std::to_string(bad) + " linear solve(s) stopped above the Krylov tolerance "
                                  "(raise max_iter with vch2d_set_krylov); the gradient would be inaccurate");
}

// publish: the kernel's last block also writes the scalars into the pinned mirror (follow with wait_scalars, not fetch_scalars)
// next_tol (tiled kernels): relative tolerance of the linear solve that would follow this evaluation; its coefficients and
// tolerance are then set by the residual kernel itself (the round-1 path sets them in schur_rhs_kernel)
__global__ void set_solve_kernel(Scal* sc, double c0, double c2, double tol2, int adj) {
    pdl_enter();
    sc->c0 = c0; sc->c2 = c2; sc->tol2 = tol2; sc->adj = adj;
}
void eval_residual(vch2d_ctx* c, const double* phi, const double* mu, double* Rphi, double* Rmu, double* a, double dt,
                   bool publish = false, double next_tol = 0.0) {
    if (c->tiled) {
        LAUNCH(c, residual_tile_kernel, c->tl.grid(), kTileThreads, phi, mu, c->cphi.p, c->cmu.p, Rphi, Rmu, a, c->g, c->tl, c->ph, dt,
               c->sc, c->red.part, c->ticket, publish ? c->sc_host : (Scal*)nullptr, 1.0 / dt, 0.5 * c->ph.kappa, next_tol * next_tol);
        return;
    }
    LAUNCH(c, residual_kernel, c->rb(), kRedThreads, phi, mu, c->cphi.p, c->cmu.p, Rphi, Rmu, a, c->g, c->ph, dt, c->sc,
           c->red.part, c->ticket, publish ? c->sc_host : (Scal*)nullptr);
}

// Solve J [dphi; dmu] = -[Rphi; Rmu] by Schur reduction:  (1/dt I - L(diag(a) - kappa/2 L)) dphi = -Rmu + L Rphi,
// dmu = 2 (a dphi - kappa/2 L dphi + Rphi).   dphi -> c->kx, dmu -> c->dmu.  phi may be null (no ceiling minima).
void newton_linear_solve(vch2d_ctx* c, const double* Rphi, const double* Rmu, const double* a, const double* phi,
                         double dt, vch_stats* st, double rel_tol = 0.0, const double* mu = nullptr, double* phit = nullptr,
                         double* mut = nullptr, bool scalars_set = false) {
    if (!(rel_tol > 0.0)) rel_tol = c->krylov_tol;
    if (c->fused_solve) {
        // rows[b = L R_phi - R_mu] cols rows[r = r0 = P^-1 b, x = 0, ||r||^2] WHILE{6 kernels} | dmu kernel with the closing x update.
        // scalars_set: the residual kernel in front has already written c0 / c2 / tol2 for this solve
        if (!scalars_set) LAUNCH(c, set_solve_kernel, 1, 1, c->sc, 1.0 / dt, 0.5 * c->ph.kappa, rel_tol * rel_tol, 0);
        c->cur_rphi = Rphi; c->cur_rmu = Rmu;
        krylov_solve<false>(c, nullptr, a, st);
        LAUNCH(c, dmu_close_tile_kernel, c->tl.grid(), kTileThreads, (const double*)c->kx.p, c->kt.p, c->kp.p, c->ks.p, a, Rphi, phi, c->dmu.p, c->g, c->tl, c->ph,
               c->sc, c->red.part, c->ticket, mu, (phi && mu) ? phit : nullptr, (phi && mu) ? mut : nullptr, c->ph.tau / dt, 1);
        return;
    }
    halo_push(c, Rphi, nullptr, 1);
    LAUNCH(c, schur_rhs_kernel, c->eb(), 256, Rphi, Rmu, c->kb.p, c->g, c->sc, 1.0 / dt, 0.5 * c->ph.kappa, rel_tol * rel_tol);
    krylov_solve<false>(c, c->kb.p, a, st);
    if (c->tiled) {
        LAUNCH(c, dmu_close_tile_kernel, c->tl.grid(), kTileThreads, (const double*)c->kx.p, (double*)nullptr, c->kp.p, c->ks.p, a, Rphi, phi, c->dmu.p, c->g, c->tl, c->ph,
               c->sc, c->red.part, c->ticket, mu, (phi && mu) ? phit : nullptr, (phi && mu) ? mut : nullptr, c->ph.tau / dt, 0);
        return;
    }
    halo_push(c, c->kx.p, nullptr, 1);
    LAUNCH(c, dmu_ceiling_kernel, c->rb(), kRedThreads, c->kx.p, a, Rphi, phi, c->dmu.p, c->g, c->ph, c->sc, c->red.part,
           c->ticket, mu, (phi && mu) ? phit : nullptr, (phi && mu) ? mut : nullptr, c->ph.tau / dt);
}

// One Newton solve (Forward2_solver.py:323-427).  Inputs: device phi_old, mu_old, w_old, w_new.
// Result left in c->phi / c->mu.  hist receives ||R|| per iteration.
//
// inexact_first (time loop): the FIRST linear solve of the Newton solve runs to krylov_first_tol (1e-6) instead of krylov_tol
// (1e-11).  Newton from the reference's initial guess always takes a second iteration here (||R|| ~ 1e4..1e8 -> 1e-2 -> 1e-8),
// and that iteration re-solves to krylov_tol from wherever the first one landed: the accepted iterate moves by ~1e-15
// relative per step (<= 2e-12 over a 1000-step trajectory, measured with the NumPy prototype in scripts/krylov_forcing_study.py)
// while the BiCGStab iterations of the forward sweep drop by 22 %.  Iteration counts of Newton itself do not change.
// wstep (time loop, tiled kernels): the set-up kernel first forms w_new = solve_w(w_old, u_n, u_{n+1}) itself (solve_w_kernel fused).
struct WStep { const double* un; const double* un1; double gdt; double* w_new; };
void newton_step(vch2d_ctx* c, const double* phi_old, const double* mu_old, const double* w_old, const double* w_new,
                 double dt, std::vector<double>* hist, vch_stats* st, bool inexact_first = false, const WStep* wstep = nullptr) {
    const long long n = c->g.n;
    const int eb = c->eb();
    // tiled kernels: the first iterate is read where it lies (a trajectory level in the time loop) — no copy into the work vector
    const bool in_place = c->tiled && phi_old != c->phi.p && phi_old != c->phit.p;
    if (!in_place) dev_copy(c, c->phi.p, phi_old, (size_t)n);
    if (c->slab && mu_old != c->mu_old.p) {   // test-level entry: bring mu_old into a ghosted work vector
        c->mu_old.alloc(n);
        dev_copy(c, c->mu_old.p, mu_old, (size_t)n);
        mu_old = c->mu_old.p;
        halo_push(c, c->phi.p, c->mu_old.p, 1);
    } else halo_push(c, c->phi.p, nullptr, 1);
    if (c->tiled)
        LAUNCH(c, step_setup_tile_kernel, c->tl.grid(), kTileThreads, in_place ? phi_old : c->phi.p, mu_old, w_old, wstep ? wstep->w_new : const_cast<double*>(w_new),
               wstep ? wstep->un : (const double*)nullptr, wstep ? wstep->un1 : (const double*)nullptr, wstep ? wstep->gdt : 0.0,
               c->cphi.p, c->cmu.p, c->mu.p, c->g, c->tl, c->ph, dt);
    else
        LAUNCH(c, step_setup_kernel, eb, 256, c->phi.p, mu_old, w_old, w_new, c->cphi.p, c->cmu.p, c->mu.p, c->g, c->ph, dt);
    halo_push(c, c->mu.p, nullptr, 1);
    const double* phi = in_place ? phi_old : c->phi.p;
    double *mu = c->mu.p, *phit = c->phit.p, *mut = c->mut.p;
    double *Rp = c->Rphi.p, *Rm = c->Rmu.p, *a = c->a.p, *RpT = c->RphiT.p, *RmT = c->RmuT.p, *aT = c->aT.p;
    // tolerance of the linear solve that follows an evaluation (the tiled residual kernel sets the solver scalars)
    const double tol_first = (inexact_first && c->krylov_first_tol > c->krylov_tol) ? c->krylov_first_tol : c->krylov_tol;
    eval_residual(c, phi, mu, Rp, Rm, a, dt, true, tol_first);
    wait_scalars(c);
    if (st) st->newton_residual_evals += 1;
    double normR = std::sqrt(c->sc_host->res2);
    const double tol = 1e-6, eta = 1e-4;
    const int max_iter = 500;
    // Resolution of R_mu in fp64: mu is stored to eps|mu| and L amplifies that by ~(1/hx^2 + 1/hy^2), so the MEASURED ||R||_2
    // cannot go below floor ~ eps (1/hx^2+1/hy^2) ||mu||_2.  On the reference's grids (<= 512^2) floor < 1e-6 and nothing below
    // fires; at >= 1024^2 it is ABOVE the reference's absolute tolerance 1e-6, where the reference's loop would spin to max_iter
    // on rounding noise.  In that regime (1.5 floor >= tol) the reference's own criterion ||R|| < tol is applied to the residual
    // with its rounding noise removed: the system is linear in (phi, mu) except for c1 l(phi), so after a full Newton step the
    // true residual is the nonlinear remainder c1 [l(phi+dphi) - l(phi) - l'(phi) dphi] (dmu_ceiling_kernel, no Laplacian in it)
    // plus what the linear solve left (<= 2e3 x its relative tolerance x ||R_prev||: the preconditioned residual under-estimates
    // the raw one by up to 1.8e3 on phase-separated states).  Unlike a threshold on the noisy measured norm (round 1: stop at
    // ||R|| <= 1.5 floor — 2 % of the steps sat within rounding of that threshold, and every flipped decision moved the trajectory
    // by ~1e-9), this decision depends on quantities that are defined to ~1e-9 relative, so all solver settings, rank counts and
    // the oracle take the same decisions.  Safety net: stop once Newton fails to halve the measured ||R|| inside 50x of the floor.
    auto floor_est = [&] { return 2.220446049250313e-16 * (c->g.ihi2 + c->g.iho2) * std::sqrt(c->sc_host->mu2); };
    double floor_now = floor_est(), dbg_prev = 0.0, true_est = INFINITY;
    int k_done = 0;
    for (int k = 0; k < max_iter; ++k) {
        k_done = k;
        if (hist) hist->push_back(normR);
        if (st) st->last_newton_residual = normR;
        if (!std::isfinite(normR)) throw Error(VCH_E_NONFINITE, "non-finite Newton residual");
        if (normR < tol) break;
        if (c->floor_aware && 1.5 * floor_now >= tol && normR < 50.0 * floor_now && true_est < tol) break;
        const double normR_prev = normR;
        dbg_prev = normR;
        const double solve_tol = (inexact_first && k == 0 && c->krylov_first_tol > c->krylov_tol) ? c->krylov_first_tol : c->krylov_tol;
        // speculative full step: almost always alpha = 1 is both allowed by the ceiling and accepted by Armijo, so the
        // trial iterate (formed by the dmu kernel) and its residual are enqueued before the host has seen the ceiling
        // -> ONE sync per Newton iteration
        newton_linear_solve(c, Rp, Rm, a, phi, dt, st, solve_tol, mu, phit, mut, c->tiled != 0);
        halo_push(c, phit, mut, 1);
        eval_residual(c, phit, mut, RpT, RmT, aT, dt, true, c->krylov_tol);
        wait_scalars(c);
        if (c->sc_host->nonfinite) throw Error(VCH_E_NONFINITE, "non-finite value in Krylov solve");
        if (st) st->newton_residual_evals += 1;
        double amax = 2.0;
        if (std::isfinite(c->sc_host->ceil_pos)) amax = std::min(amax, 0.9 * c->sc_host->ceil_pos);
        if (std::isfinite(c->sc_host->ceil_neg)) amax = std::min(amax, 0.9 * c->sc_host->ceil_neg);
        if (!std::isfinite(amax) || amax <= 0.0) amax = 1.0;
        double alpha = std::min(1.0, amax);
        const double rem_full = std::sqrt(c->sc_host->rem2), lin_left = 2e3 * solve_tol * normR_prev;
        const bool full_step = (alpha == 1.0);
        double best = INFINITY, best_alpha = 0.0;
        bool accepted = false;
        bool have_trial = (alpha == 1.0);      // the speculative evaluation is the first trial
        for (int ls = 0; ls < 12; ++ls) {
            if (!have_trial) {
                LAUNCH(c, trial_kernel, eb, 256, phi, mu, c->dphi(), c->dmu.p, phit, mut, n, alpha);
                halo_push(c, phit, mut, 1);
                eval_residual(c, phit, mut, RpT, RmT, aT, dt, true, c->krylov_tol);
                wait_scalars(c);
                if (st) st->newton_residual_evals += 1;
            }
            have_trial = false;
            const double nt = std::sqrt(c->sc_host->res2);
            if (nt < best) { best = nt; best_alpha = alpha; }
            if (nt <= (1.0 - eta * alpha) * normR) {
                accepted = true;
                normR = nt;
                break;
            }
            alpha *= 0.5;
        }
        // noise-free estimate of the accepted iterate's residual: only for the full step (a damped step keeps (1 - alpha) R)
        true_est = (accepted && full_step && alpha == 1.0) ? std::sqrt(rem_full * rem_full + lin_left * lin_left) : INFINITY;
        if (!accepted) {
            if (best < normR) {   // fall back to the best trial (re-evaluated: same arithmetic, same values)
                LAUNCH(c, trial_kernel, eb, 256, phi, mu, c->dphi(), c->dmu.p, phit, mut, n, best_alpha);
                halo_push(c, phit, mut, 1);
                eval_residual(c, phit, mut, RpT, RmT, aT, dt, true, c->krylov_tol);
                wait_scalars(c);
                normR = std::sqrt(c->sc_host->res2);
                accepted = true;
            }
        }
        if (accepted) {
            // the old iterate's buffer becomes the next trial buffer — unless it is the caller's array (in_place, first accept)
            double* freed = (phi == phi_old && in_place) ? c->phi.p : const_cast<double*>(phi);
            phi = phit; phit = freed;
            std::swap(mu, mut);
            std::swap(Rp, RpT); std::swap(Rm, RmT); std::swap(a, aT);
            floor_now = floor_est();
        }
        if (c->floor_aware && normR >= tol && normR > 0.5 * normR_prev && normR < 50.0 * floor_now) {
            if (hist) hist->push_back(normR);
            if (st) { st->last_newton_residual = normR; }
            break;   // stalled at fp64 resolution
        }
    }
    if (c->debug) fprintf(stderr, "[vch2d] newton: %d its, |R| = %.3e, floor_est = %.3e, before the last solve %.3e\n", k_done, normR, floor_now, dbg_prev);
    // leave the result in c->phi / c->mu
    if (phi == phi_old && in_place) dev_copy(c, c->phi.p, phi_old, (size_t)n);     // no step was taken (already converged)
    else if (phi != c->phi.p) { std::swap(c->phi.p, c->phit.p); std::swap(c->phi.n, c->phit.n); std::swap(c->phi.view, c->phit.view); }
    if (mu != c->mu.p) { std::swap(c->mu.p, c->mut.p); std::swap(c->mu.n, c->mut.n); std::swap(c->mu.view, c->mut.view); }
    if (Rp != c->Rphi.p) {
        std::swap(c->Rphi.p, c->RphiT.p); std::swap(c->Rmu.p, c->RmuT.p); std::swap(c->a.p, c->aT.p);
    }
}

// Post-Newton clip + interior mass correction into `dst` (Forward2_solver.py:562-577).
void post_step(vch2d_ctx* c, const double* phi_new, double* dst) {
    const double hxhy = c->prm.hx * c->prm.hy;
    LAUNCH(c, clip_mass_kernel, c->rb(), kRedThreads, phi_new, dst, c->g, c->ph, hxhy, c->sc, 0, c->red.part, c->ticket);
    LAUNCH(c, mass_shift_kernel, c->eb(), 256, dst, c->g, c->ph, c->prm.Lx * c->prm.Ly, c->sc);
}

using LevelHook = std::function<void(int)>;
// level -> device address, for trajectories that live in chunk rings (bounded-memory streaming); unset entries fall back
// to the dense arrays passed alongside
struct LevelMap {
    std::function<double*(int)> hist, Q, r, u, hist_new;
};

// Time steps s0 .. s1-1 of the forward solve (Forward2_solver.py:542-585).  On entry HN(s0) holds phi_{s0}, c->mu_old holds mu_{s0}
// and c->w0 holds w_{s0}; sc->mass0 has been set from level 0 (forward_begin).  Level s+1 is written to HN(s+1).
// after_level(k): called once level k of phi_hist has been enqueued (streaming D2H of the trajectory hooks in here).
void forward_steps(vch2d_ctx* c, int s0, int s1, const std::function<double*(int)>& HN, const std::function<const double*(int)>& U,
                   bool have_u, int u_rows, const double* dt_steps, double* mu_hist, double* w_hist, vch_stats* st,
                   const LevelHook& after_level, const LevelHook& before_step) {
    const long long n = c->g.n;
    const int eb = c->eb();
    double* mu_old = c->mu_old.p;
    for (int s = s0; s < s1; ++s) {
        const double dt = dt_steps[s];
        if (before_step) before_step(s);     // streaming path: makes sure control rows s and s+1 exist
        const double* un = nullptr; const double* un1 = nullptr;
        if (have_u && s < u_rows - 1) { un = U(s); un1 = U(s + 1); }
        const double* phi_old = HN(s);
        if (c->tiled) {
            const WStep ws{un, un1, c->prm.gamma / dt, c->w1.p};
            newton_step(c, phi_old, mu_old, c->w0.p, c->w1.p, dt, nullptr, st, true, &ws);
        } else {
            LAUNCH(c, solve_w_kernel, eb, 256, c->w0.p, un, un1, c->w1.p, n, c->prm.gamma / dt);
            newton_step(c, phi_old, mu_old, c->w0.p, c->w1.p, dt, nullptr, st, true);
        }
        post_step(c, c->phi.p, HN(s + 1));
        if (c->tiled) {   // the accepted mu becomes mu_old by exchanging the two work vectors
            std::swap(c->mu_old.p, c->mu.p); std::swap(c->mu_old.n, c->mu.n); std::swap(c->mu_old.view, c->mu.view);
            mu_old = c->mu_old.p;
        } else copy_ghosted(c, mu_old, c->mu.p);      // slab mode: the ghost rows of the accepted iterate travel along
        std::swap(c->w0.p, c->w1.p);
        if (mu_hist) dev_copy(c, mu_hist + (size_t)s * n, mu_old, (size_t)n);
        if (w_hist) dev_copy(c, w_hist + (size_t)s * n, c->w0.p, (size_t)n);
        if (after_level) after_level(s + 1);
    }
}

// Start of a forward solve from level 0 (already stored at lvl0): w_0 = 0, mu_0 = initialize_mu(phi_0, 0) (Forward2_solver.py:520),
// and the reference mass of the interior mass correction (:565).
void forward_begin(vch2d_ctx* c, const double* lvl0, bool set_state) {
    const long long n = c->g.n;
    const int eb = c->eb();
    c->mu_old.alloc(n);
    if (set_state) {
        VCH_CUDA(cudaMemsetAsync(c->w0.p, 0, n * sizeof(double), c->stream));
        const double* phi_first = lvl0;
        if (c->slab) {   // the stencil needs ghost rows: work on a ghosted copy of level 0
            dev_copy(c, c->phi.p, lvl0, (size_t)n);
            halo_push(c, c->phi.p, nullptr, 1);
            phi_first = c->phi.p;
        }
        LAUNCH(c, mu_init_kernel, eb, 256, phi_first, c->w0.p, c->mu_old.p, c->g, c->ph);
        halo_push(c, c->mu_old.p, nullptr, 1);
    }
    LAUNCH(c, clip_mass_kernel, c->rb(), kRedThreads, lvl0, (double*)nullptr, c->g, c->ph, c->prm.hx * c->prm.hy, c->sc, 1,
           c->red.part, c->ticket);
}

void forward_dev(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int n_steps, const double* dt_steps,
                 double* phi_hist, double* mu_hist, double* w_hist, vch_stats* st, const LevelHook& after_level = nullptr,
                 const LevelHook& before_step = nullptr, const LevelMap* lm = nullptr) {
    const long long n = c->g.n;
    auto HN = [&](int k) -> double* { return (lm && lm->hist_new) ? lm->hist_new(k) : phi_hist + (size_t)k * n; };
    auto U = [&](int k) -> const double* { return (lm && lm->u) ? lm->u(k) : u + (size_t)k * n; };
    dev_copy(c, HN(0), phi0, (size_t)n);
    forward_begin(c, HN(0), true);
    forward_steps(c, 0, n_steps, HN, U, u || (lm && lm->u), u_rows, dt_steps, mu_hist, w_hist, st, after_level, before_step);
    VCH_CUDA(cudaGetLastError());
}

// need_level(k): called before level k of phi_hist / phiQ is first read (streaming H2D hooks in here; levels descend).
// k_hi / k_lo: only the levels k_hi (terminal solve, when k_hi is the last level) .. k_lo are processed; the rolling p, q (and r)
// of level k_hi must then still be in the context's rings from the call that produced them (checkpointed sweep).
void adjoint_dev(vch2d_ctx* c, const double* phi_hist, int levels, const double* t_hist, double b1, double b2,
                 const double* phiQ, const double* phiT, double* p_out, double* q_out, double* r_out, vch_stats* st,
                 const LevelHook& need_level = nullptr, const LevelMap* lm = nullptr, int k_hi = -1, int k_lo = 0) {
    const long long n = c->g.n;
    const int eb = c->eb();
    auto H = [&](int k) -> const double* { return (lm && lm->hist) ? lm->hist(k) : phi_hist + (size_t)k * n; };
    const bool haveQ = phiQ || (lm && lm->Q);
    auto Q = [&](int k) -> const double* { return !haveQ ? nullptr : ((lm && lm->Q) ? lm->Q(k) : phiQ + (size_t)k * n); };
    for (int k = 0; k < 2; ++k) { c->adj_p[k].alloc(n); c->adj_q[k].alloc(n); c->adj_r[k].alloc(n); }
    // slab mode: p and q feed stencils, so they always live in the ghosted ring and are copied out when asked for
    auto slot = [&](double* out, DevBuf (&ring)[2], int lvl) { return out ? out + (size_t)lvl * n : ring[lvl & 1].p; };
    double* const p_user = c->slab ? p_out : nullptr; double* const q_user = c->slab ? q_out : nullptr;
    if (c->slab) { p_out = nullptr; q_out = nullptr; }
    auto copy_out = [&](int lvl, const double* pv, const double* qv) {
        if (p_user) dev_copy(c, p_user + (size_t)lvl * n, pv, (size_t)n);
        if (q_user) dev_copy(c, q_user + (size_t)lvl * n, qv, (size_t)n);
    };
    const int M = levels - 1;
    if (k_hi < 0) k_hi = M;
    auto rslot = [&](int lvl) -> double* { return (lm && lm->r) ? lm->r(lvl) : slot(r_out, c->adj_r, lvl); };
    if (k_hi == M) {
        if (need_level) need_level(M);
        double* pM = slot(p_out, c->adj_p, M); double* qM = slot(q_out, c->adj_q, M); double* rM = rslot(M);
        LAUNCH(c, adj_terminal_rhs_kernel, eb, 256, H(M), phiT, c->kb.p, n, b2);
        SymbolArgs sy{1.0, 0.0, nullptr, c->ph.tau, nullptr};
        c->dct.apply(c->stream, c->kb.p, pM, sy, nullptr);          // (I - tau L) p_M = b2 (phi_M - phi_T): exact in the DCT basis
        halo_push(c, pM, nullptr, 1);
        if (c->tiled) LAUNCH(c, adj_qr_tile_kernel, c->tl.grid(), kTileThreads, (const double*)pM, pM, (const double*)nullptr, (const double*)nullptr, qM, rM, c->g, c->tl, 0.0, 0.0);
        else LAUNCH(c, adj_qr_kernel, eb, 256, pM, (const double*)nullptr, (const double*)nullptr, qM, rM, c->g, 0.0, 0.0);
        halo_push(c, qM, nullptr, 1);
        copy_out(M, pM, qM);
    }
    for (int k = k_hi - 1; k >= k_lo; --k) {
        const double dt = t_hist[k + 1] - t_hist[k];
        if (need_level) need_level(k);     // before the level's addresses are taken: ring slots may be recycled in here
        double *p0 = slot(p_out, c->adj_p, k), *q0 = slot(q_out, c->adj_q, k), *r0 = rslot(k);
        const double *p1 = slot(p_out, c->adj_p, k + 1), *q1 = slot(q_out, c->adj_q, k + 1), *r1 = rslot(k + 1);
        if (dt <= 1e-14) {   // backward2_solver.py:214-216
            if (c->slab) { copy_ghosted(c, p0, p1); copy_ghosted(c, q0, q1); }
            else {
                dev_copy(c, p0, p1, (size_t)n);
                dev_copy(c, q0, q1, (size_t)n);
            }
            dev_copy(c, r0, r1, (size_t)n);
            copy_out(k, p0, q0);
            continue;
        }
        const double* f1 = H(k + 1); const double* f0 = H(k);
        const double den = c->ph.gamma + 0.5 * dt;
        if (c->tiled) {
            // rhs kernel (+ the BiCGStab start when the solve is fused) | solve graph | q/r recurrence (+ the copy of the solution into p_k)
            const int fs = c->fused_solve;
            LAUNCH(c, adj_rhs_tile_kernel, c->tl.grid(), kTileThreads, p1, q1, f1, f0, Q(k + 1), Q(k), fs ? c->kr.p : c->kb.p, c->kr0.p, c->kx.p,
                   c->a.p, c->g, c->tl, c->ph, dt, b1, c->sc, c->red.part, c->ticket, c->krylov_tol * c->krylov_tol, fs);
            krylov_solve<true>(c, fs ? (const double*)nullptr : c->kb.p, c->a.p, st);
            LAUNCH(c, adj_qr_tile_kernel, c->tl.grid(), kTileThreads, (const double*)c->kx.p, p0, q1, r1, q0, r0, c->g, c->tl,
                   (c->ph.gamma - 0.5 * dt) / den, 0.5 * dt / den);
            copy_out(k, p0, q0);
            continue;
        }
        LAUNCH(c, adj_rhs_kernel, c->rb(), kRedThreads, p1, q1, f1, f0, Q(k + 1), Q(k), c->kb.p, c->a.p, c->g, c->ph, dt, b1, c->sc,
               c->red.part, c->ticket, c->krylov_tol * c->krylov_tol);
        krylov_solve<true>(c, c->kb.p, c->a.p, st);
        dev_copy(c, p0, c->kx.p, (size_t)n);
        halo_push(c, p0, nullptr, 1);
        LAUNCH(c, adj_qr_kernel, eb, 256, p0, q1, r1, q0, r0, c->g, (c->ph.gamma - 0.5 * dt) / den, 0.5 * dt / den);
        halo_push(c, q0, nullptr, 1);
        copy_out(k, p0, q0);
    }
    VCH_CUDA(cudaGetLastError());
}

// np.trapz weights for abscissae x: w_i = (x_{i+1}-x_{i-1})/2 with one-sided ends.
std::vector<double> trapz_weights(const double* x, int n) {
    std::vector<double> w(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) { const double h = 0.5 * (x[i + 1] - x[i]); w[i] += h; w[i + 1] += h; }
    return w;
}

struct CostWeights { double *wt, *wx, *wy; };
// Uploads the np.trapz weights (time, x, y) into c->small.  Synchronises the work stream (pageable host vectors).
CostWeights cost_weights(vch2d_ctx* c, int levels, const double* x, const double* y, const double* t) {
    const int nx1 = c->g.nx1, ny1 = c->g.ny1;   // slab mode: x is the GLOBAL abscissa vector, this rank integrates rows [o0, o0 + nx1)
    std::vector<double> w = trapz_weights(t, levels), wxg = trapz_weights(x, c->g.nxg), wy = trapz_weights(y, ny1);
    std::vector<double> wx(wxg.begin() + c->g.o0, wxg.begin() + c->g.o0 + nx1);
    c->small.alloc((size_t)levels + nx1 + ny1);
    double *dwt = c->small.p, *dwx = dwt + levels, *dwy = dwx + nx1;
    VCH_CUDA(cudaMemcpyAsync(dwt, w.data(), levels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VCH_CUDA(cudaMemcpyAsync(dwx, wx.data(), nx1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VCH_CUDA(cudaMemcpyAsync(dwy, wy.data(), ny1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VCH_CUDA(cudaStreamSynchronize(c->stream));
    return {dwt, dwx, dwy};
}
void cost_finish(vch2d_ctx* c, double b1, double b2, double b3, double ksp, double* J_out);

void cost_dev(vch2d_ctx* c, const double* phi_hist, const double* u, const double* phiQ, const double* phiT, int levels,
              const double* x, const double* y, const double* t, double b1, double b2, double b3, double ksp, double* J_out) {
    const CostWeights cw = cost_weights(c, levels, x, y, t);
    const long long total = (long long)levels * c->g.n;
    LAUNCH(c, cost_kernel, red_blocks(total), kRedThreads, phi_hist, u, phiQ, phiT, levels, c->g.nx1, c->g.ny1, cw.wt, cw.wx, cw.wy,
           c->out4, c->red.part, c->ticket, -2, 0);
    cost_finish(c, b1, b2, b3, ksp, J_out);
}
// Reads the four raw integrals from c->out4 and forms J, J1..J4.
void cost_finish(vch2d_ctx* c, double b1, double b2, double b3, double ksp, double* J_out) {
    VCH_CUDA(cudaMemcpyAsync(c->out4_host, c->out4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VCH_CUDA(cudaStreamSynchronize(c->stream));
    const double J1 = 0.5 * b1 * c->out4_host[0], J2 = 0.5 * b2 * c->out4_host[1], J3 = 0.5 * b3 * c->out4_host[2],
                 J4 = ksp * c->out4_host[3];
    J_out[0] = J1 + J2 + J3 + J4; J_out[1] = J1; J_out[2] = J2; J_out[3] = J3; J_out[4] = J4;
    for (int k = 0; k < 4; ++k) J_out[5 + k] = c->out4_host[k];   // raw integrals: the driver's monitoring norms (D1) come for free
}

struct SmallScratch {   // context-free reductions (vch_grad_prox / vch_kkt_counts / vch_solve_w)
    RedBuf red; double* part = nullptr; unsigned int* ticket = nullptr; double* out = nullptr; double* out_host = nullptr;
    void ensure() {
        if (part) return;
        red.alloc(8 * kRedBlocksMax, Comm());
        part = red.part;
        VCH_CUDA(cudaMalloc(&ticket, sizeof(unsigned int)));
        VCH_CUDA(cudaMemset(ticket, 0, sizeof(unsigned int)));
        VCH_CUDA(cudaMalloc(&out, 8 * sizeof(double)));
        VCH_CUDA(cudaMallocHost(&out_host, 8 * sizeof(double)));
    }
};
thread_local SmallScratch g_scratch;
thread_local long long g_free_launches = 0;

void check_params(const vch2d_params* p) {
    VCH_REQUIRE(p != nullptr, VCH_E_ARG, "null params");
    VCH_REQUIRE(p->Nx >= 2 && p->Ny >= 2, VCH_E_SHAPE, "Nx, Ny must be >= 2");
    VCH_REQUIRE(p->hx > 0 && p->hy > 0, VCH_E_ARG, "hx, hy must be positive");
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

const char* vch_last_error(void) { return g_last_error.c_str(); }
int vch_version(void) { return 100; }
int vch_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int create_ctx(const vch2d_params* p, int device, int rank, int nranks, vch2d_ctx** out) {
    return guarded([&] {
        check_params(p);
        VCH_REQUIRE(out != nullptr, VCH_E_ARG, "null out");
        VCH_REQUIRE(vch_device_count() > device, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        VCH_CUDA(cudaSetDevice(device));
        const bool slab = nranks > 1;
        if (slab) {
            const int N = p->Nx;
            VCH_REQUIRE(p->Nx == p->Ny, VCH_E_SHAPE, "slab mode needs a square grid (Nx == Ny)");
            VCH_REQUIRE((N & (N - 1)) == 0 && N >= 32 && N <= 4096, VCH_E_SHAPE, "slab mode needs N = 2^k, 32 <= N <= 4096");
            VCH_REQUIRE(nranks == 2 || nranks == 4 || nranks == 8, VCH_E_ARG, "slab mode: 2, 4 or 8 ranks");
            VCH_REQUIRE(rank >= 0 && rank < nranks && N / nranks >= 8, VCH_E_ARG, "slab mode: bad rank / fewer than 8 rows per rank");
        }
        auto* c = new vch2d_ctx();
        c->prm = *p; c->device = device;
        c->debug = getenv("VCH_DEBUG") ? atoi(getenv("VCH_DEBUG")) : 0;
        if (getenv("VCH_NEWTON_STRICT")) c->floor_aware = 0;
        if (getenv("VCH_NO_HALF_EXIT")) c->half_exit = 0;
        c->bicg6 = (getenv("VCH_BICG6") ? atoi(getenv("VCH_BICG6")) : 1);
        c->pdl = (getenv("VCH_PDL") ? atoi(getenv("VCH_PDL")) : 0) && !slab;   // slab mode: cross-rank waits inside kernels, keep full serialization
        if (getenv("VCH_KRYLOV_FIRST_RTOL")) { const double t = atof(getenv("VCH_KRYLOV_FIRST_RTOL")); if (t >= 0 && t < 1) c->krylov_first_tol = t; }
        if (getenv("VCH_KRYLOV_RTOL")) { const double t = atof(getenv("VCH_KRYLOV_RTOL")); if (t > 0 && t < 1) c->krylov_tol = t; }
        Geo& g = c->g;
        g.ni = p->Nx + 1; g.no = p->Ny + 1; g.nx1 = p->Nx + 1; g.ny1 = p->Ny + 1;
        g.nxg = g.nx1; g.o0 = 0;
        g.ihi2 = 1.0 / (p->hx * p->hx); g.iho2 = 1.0 / (p->hy * p->hy);
        DctSlab sl;
        if (slab) {   // rows [o0, o0 + no) of the (N+1) x (N+1) grid; every rank but the last owns N / nranks rows
            const int N = p->Nx, rw = N / nranks;
            c->slab = true; c->rows_lo = rw;
            c->cm.rank = rank; c->cm.nranks = nranks;
            g.o0 = rank * rw; g.no = rw + (rank == nranks - 1 ? 1 : 0);
            g.nx1 = g.no; g.nxg = N + 1;
            g.glo = rank > 0; g.ghi = rank < nranks - 1;
            sl.nloc = g.no; sl.o0 = g.o0; sl.wloc = g.no; sl.col0 = g.o0;
            sl.shift = 0; while ((1 << sl.shift) < rw) ++sl.shift;
            sl.p1 = (rw + 1 + 7) & ~7;      // owned columns (rw or rw + 1) rounded up to whole 8-column groups of the column kernel; padding stays zero
        }
        g.n = (long long)g.ni * g.no;
        c->ph = Phys{p->tau, p->gamma, p->c1, p->c2, p->kappa, 1.0 - p->delta_sep, std::max(1e-8, 0.5 * p->delta_sep),
                     1.0 - p->delta_sep * p->delta_sep};
        const size_t n = (size_t)g.n;
        std::vector<DevBuf*> fields = {&c->phi, &c->mu, &c->phit, &c->mut, &c->w0, &c->w1, &c->cphi, &c->cmu, &c->Rphi, &c->Rmu, &c->a,
                                       &c->RphiT, &c->RmuT, &c->aT, &c->kb, &c->kx, &c->kr, &c->kr0, &c->kp, &c->kv, &c->ks, &c->kt,
                                       &c->ktmp, &c->kq, &c->dmu};
        if (slab) {
            // identical arena layout on every rank: slots are sized for the largest slab (rows_lo + 1 rows)
            for (int k = 0; k < 2; ++k) { fields.push_back(&c->adj_p[k]); fields.push_back(&c->adj_q[k]); fields.push_back(&c->adj_r[k]); }
            fields.push_back(&c->mu_old);
            const size_t nmax = (size_t)(c->rows_lo + 1) * g.ni, margin = 2 * (size_t)g.ni;
            const size_t t1 = (size_t)(p->Nx + 1) * sl.p1;
            c->arena.create(kArenaHeader + fields.size() * (nmax + 2 * margin + 64) + t1 + 256);
            for (DevBuf* b : fields) c->arena.view(*b, nmax, margin);
            sl.T1 = c->arena.carve(t1, 0);
            c->cm.peer[rank] = c->arena.base;
        } else {
            for (DevBuf* b : fields) b->alloc(n);
        }

        VCH_CUDA(cudaMalloc(&c->ticket, sizeof(unsigned int)));
        VCH_CUDA(cudaMemset(c->ticket, 0, sizeof(unsigned int)));
        VCH_CUDA(cudaMalloc(&c->sc, sizeof(Scal)));
        VCH_CUDA(cudaMemset(c->sc, 0, sizeof(Scal)));
        VCH_CUDA(cudaHostAlloc(&c->sc_host, sizeof(Scal), cudaHostAllocMapped));   // written by publish_scalars_kernel (UVA: same pointer)
        VCH_CUDA(cudaMalloc(&c->out4, 8 * sizeof(double)));
        VCH_CUDA(cudaMallocHost(&c->out4_host, 8 * sizeof(double)));
        if (slab) {
            VCH_CUDA(cudaMalloc(&c->cm.seq, 2 * sizeof(unsigned long long)));
            VCH_CUDA(cudaMemset(c->cm.seq, 0, 2 * sizeof(unsigned long long)));
            c->cm.err = &c->sc->comm_err;
            sl.cm = c->cm;
            c->dct.init_slab(p->Nx + 1, p->hy, p->hx, &c->log, sl);
            if (!(c->dct.lean && c->dct.inner.log2L >= 8)) c->bicg6 = 0;   // the round-1 kernels carry the 6-launch modes only without transposition
        } else {
            c->dct.init(g.no, g.ni, p->hy, p->hx, &c->log);
            c->dct.pdl = c->pdl != 0;
            if (!(c->dct.inner.fft && c->dct.outer.fft)) c->bicg6 = 0;   // the dense-table path has no fused mode 3 / 4
        }
        c->tl = make_tiling(g);
        c->tiled = !slab && !(getenv("VCH_TILED") && atoi(getenv("VCH_TILED")) == 0);
        c->fused_solve = c->tiled && c->bicg6 && c->dct.lean && c->dct.inner.fft && c->dct.outer.fft &&
                         !(getenv("VCH_FUSED_SOLVE") && atoi(getenv("VCH_FUSED_SOLVE")) == 0);
        c->red.alloc(8 * (size_t)std::max(c->dct.max_grid(), c->tl.grid()), c->cm);
        Scal init{}; init.tol2 = c->krylov_tol * c->krylov_tol; init.maxit = c->krylov_maxit;
        VCH_CUDA(cudaMemcpy(c->sc, &init, sizeof(Scal), cudaMemcpyHostToDevice));
        VCH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        VCH_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
        VCH_CUDA(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
        if (getenv("VCH_NO_GRAPHS")) c->use_graphs = 0;
        *out = c;
        return VCH_OK;
    });
}

int vch2d_create(const vch2d_params* p, int device, vch2d_ctx** out) { return create_ctx(p, device, 0, 1, out); }

// ---- slab mode: one rank of a row-slab decomposition (one process per GPU; peers are reached through CUDA IPC)
int vch2d_slab_create(const vch2d_params* p, int device, int rank, int nranks, vch2d_ctx** out) {
    if (nranks < 2) { set_last_error("slab mode needs at least 2 ranks (use vch2d_create)"); return VCH_E_ARG; }
    return create_ctx(p, device, rank, nranks, out);
}

int vch2d_slab_rows(vch2d_ctx* c, int* row0, int* nrows) {
    return guarded([&] {
        VCH_REQUIRE(c && row0 && nrows, VCH_E_ARG, "null argument");
        *row0 = c->g.o0; *nrows = c->g.no;
        return VCH_OK;
    });
}

int vch2d_slab_ipc_handle(vch2d_ctx* c, void* handle_out) {
    return guarded([&] {
        VCH_REQUIRE(c && c->slab && handle_out, VCH_E_ARG, "not a slab context");
        VCH_CUDA(cudaSetDevice(c->device));
        cudaIpcMemHandle_t h;
        VCH_CUDA(cudaIpcGetMemHandle(&h, c->arena.base));
        static_assert(sizeof(h) == VCH_IPC_HANDLE_BYTES, "IPC handle size");
        std::memcpy(handle_out, &h, sizeof(h));
        return VCH_OK;
    });
}

int vch2d_slab_attach(vch2d_ctx* c, const void* handles) {
    return guarded([&] {
        VCH_REQUIRE(c && c->slab && handles && !c->attached, VCH_E_ARG, "slab_attach: not a slab context / already attached");
        VCH_CUDA(cudaSetDevice(c->device));
        for (int r = 0; r < c->cm.nranks; ++r) {
            if (r == c->cm.rank) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char*)handles + (size_t)r * VCH_IPC_HANDLE_BYTES, sizeof(h));
            void* ptr = nullptr;
            VCH_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            c->peer_maps.push_back(ptr);
            c->cm.peer[r] = (double*)ptr;
        }
        c->red.set_comm(c->cm);
        c->dct.slab.cm = c->cm;
        c->attached = true;
        return VCH_OK;
    });
}

// Exercises the three cross-rank primitives: barrier, neighbour halo push, reduction (sum / min / max of rank + 1).
// out (host, 5): sum, min, max, lower ghost value (rank of the neighbour below + 1, or 0), upper ghost value.
__global__ void slab_selftest_kernel(double* field, int n, double val, double* out3, double* part, unsigned int* ticket) {
    pdl_enter();
    double v[3] = {0.0, INFINITY, -INFINITY};
    if (blockIdx.x == 0 && threadIdx.x == 0) { v[0] = val; v[1] = val; v[2] = val; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) field[i] = val;
    const int op[3] = {0, 1, 2};
    double tot[3];
    if (grid_reduce<3>(v, op, part, ticket, tot) && threadIdx.x == 0) { out3[0] = tot[0]; out3[1] = tot[1]; out3[2] = tot[2]; }
}
int vch2d_slab_selftest(vch2d_ctx* c, double* out5) {
    return guarded([&] {
        VCH_REQUIRE(c && c->slab && c->attached && out5, VCH_E_ARG, "slab_selftest: attach first");
        StreamScope scope(c);
        const int n = (int)c->g.n, ni = c->g.ni;
        LAUNCH(c, slab_selftest_kernel, 8, 256, c->ktmp.p, n, (double)(c->cm.rank + 1), c->out4, c->red.part, c->ticket);
        halo_push(c, c->ktmp.p, nullptr, 2, nullptr, 1);
        double ghost[2] = {0.0, 0.0};
        if (c->g.glo) VCH_CUDA(cudaMemcpyAsync(&ghost[0], c->ktmp.p - 2 * ni, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (c->g.ghi) VCH_CUDA(cudaMemcpyAsync(&ghost[1], c->ktmp.p + n + 2 * ni - 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VCH_CUDA(cudaMemcpyAsync(c->out4_host, c->out4, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        fetch_scalars(c);
        for (int k = 0; k < 3; ++k) out5[k] = c->out4_host[k];
        out5[3] = ghost[0]; out5[4] = ghost[1];
        return VCH_OK;
    });
}

void vch2d_destroy(vch2d_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& g : c->graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.g); }
    cudaEventDestroy(c->ev_in); cudaEventDestroy(c->ev_out); cudaStreamDestroy(c->stream);
    c->dct.destroy();
    for (void* m : c->peer_maps) cudaIpcCloseMemHandle(m);
    if (c->cm.seq) cudaFree(c->cm.seq);
    c->arena.destroy();
    cudaFree(c->ticket); cudaFree(c->sc); cudaFreeHost(c->sc_host); cudaFree(c->out4); cudaFreeHost(c->out4_host);
    delete c;
}

int vch2d_set_stream(vch2d_ctx* c, void* s) {
    return guarded([&] { VCH_REQUIRE(c, VCH_E_ARG, "null ctx"); c->user = (cudaStream_t)s; return VCH_OK; });
}

int vch2d_set_krylov(vch2d_ctx* c, double rel_tol, int max_iter) {
    return guarded([&] {
        VCH_REQUIRE(c && rel_tol > 0 && max_iter > 0, VCH_E_ARG, "bad Krylov settings");
        c->krylov_tol = rel_tol; c->krylov_maxit = max_iter;
        const double t2 = rel_tol * rel_tol;
        VCH_CUDA(cudaSetDevice(c->device));
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        VCH_CUDA(cudaMemcpy(&c->sc->tol2, &t2, sizeof(double), cudaMemcpyHostToDevice));
        VCH_CUDA(cudaMemcpy(&c->sc->maxit, &max_iter, sizeof(int), cudaMemcpyHostToDevice));
        return VCH_OK;
    });
}

int vch2d_set_krylov_first(vch2d_ctx* c, double rel_tol) {
    return guarded([&] {
        VCH_REQUIRE(c && rel_tol >= 0 && rel_tol < 1, VCH_E_ARG, "bad first-solve tolerance");
        c->krylov_first_tol = rel_tol;
        return VCH_OK;
    });
}

int vch2d_set_stream_budget(vch2d_ctx* c, long long bytes) {
    return guarded([&] {
        VCH_REQUIRE(c && bytes >= 0, VCH_E_ARG, "bad stream budget");
        if (bytes != c->stream_budget) for (auto& b : c->stage) b.release();
        c->stream_budget = bytes;
        return VCH_OK;
    });
}

int vch2d_set_newton(vch2d_ctx* c, int floor_aware) {
    return guarded([&] { VCH_REQUIRE(c, VCH_E_ARG, "null ctx"); c->floor_aware = floor_aware ? 1 : 0; return VCH_OK; });
}

long long vch2d_launch_count(vch2d_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (cudaMemcpyAsync(c->sc_host, c->sc, sizeof(Scal), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return c->log.count;
    cudaStreamSynchronize(c->stream);
    return c->log.count + c->sc_host->g_launches;   // + kernels that ran inside solve graphs
}

int vch2d_profile(vch2d_ctx* c, int enable) {
    return guarded([&] {
        VCH_REQUIRE(c, VCH_E_ARG, "null ctx");
        VCH_CUDA(cudaSetDevice(c->device));
        if (enable) c->log.report();   // drop stale records
        c->log.profiling = enable != 0;
        return VCH_OK;
    });
}

int vch2d_profile_report(vch2d_ctx* c, char* names, int names_cap, double* ms, long long* counts, int cap, int* n_out) {
    return guarded([&] {
        VCH_REQUIRE(c && names && ms && counts && n_out, VCH_E_ARG, "null argument");
        VCH_CUDA(cudaSetDevice(c->device));
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        auto rows = c->log.report();
        std::string joined;
        int k = 0;
        for (auto& r : rows) {
            if (k >= cap) break;
            if (k) joined += ";";
            joined += r.name; ms[k] = r.ms; counts[k] = r.n; ++k;
        }
        VCH_REQUIRE((int)joined.size() < names_cap, VCH_E_ARG, "names buffer too small");
        std::memcpy(names, joined.c_str(), joined.size() + 1);
        *n_out = k;
        return VCH_OK;
    });
}

int vch2d_apply_laplacian(vch2d_ctx* c, const double* v, double* out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && v && out, VCH_E_SHAPE, "apply_laplacian: null array");
        StreamScope scope(c);
        Stager st(c->stream, mem);
        const double* dv = st.in(v, c->g.n); double* dout = st.out(out, c->g.n);
        if (c->slab) {   // ghosted copy + halo rows from the neighbours
            VCH_CUDA(cudaMemcpyAsync(c->ktmp.p, dv, c->g.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            halo_push(c, c->ktmp.p, nullptr, 1, nullptr, 1);
            dv = c->ktmp.p;
        }
        LAUNCH(c, lap_kernel, c->eb(), 256, dv, dout, c->g, 1.0);
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch2d_initialize_mu(vch2d_ctx* c, const double* phi, const double* w, double* mu_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi && w && mu_out, VCH_E_SHAPE, "initialize_mu: null array");
        StreamScope scope(c);
        Stager st(c->stream, mem);
        const double *dp = st.in(phi, c->g.n), *dw = st.in(w, c->g.n); double* dm = st.out(mu_out, c->g.n);
        if (c->slab) {
            VCH_CUDA(cudaMemcpyAsync(c->ktmp.p, dp, c->g.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            halo_push(c, c->ktmp.p, nullptr, 1, nullptr, 1);
            dp = c->ktmp.p;
        }
        LAUNCH(c, mu_init_kernel, c->eb(), 256, dp, dw, dm, c->g, c->ph);
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch_solve_w(void* stream, long long count, const double* w_old, double dt, double gamma, const double* u_n,
                const double* u_np1, double* w_new_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(count > 0 && w_old && w_new_out, VCH_E_SHAPE, "solve_w: bad arguments");
        VCH_REQUIRE(vch_device_count() > 0, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        cudaStream_t s = (cudaStream_t)stream;
        Stager st(s, mem);
        const double *dw = st.in(w_old, count), *da = st.in(u_n, count), *db = st.in(u_np1, count);
        double* dout = st.out(w_new_out, count);
        solve_w_kernel<<<(int)std::min<long long>((count + 255) / 256, 1 << 20), 256, 0, s>>>(dw, da, db, dout, count, gamma / dt);
        ++g_free_launches;
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch2d_residual(vch2d_ctx* c, const double* phi_new, const double* phi_old, const double* mu_new, const double* mu_old,
                   const double* w_new, const double* w_old, double dt, double* Rphi_out, double* Rmu_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi_new && phi_old && mu_new && mu_old && w_new && w_old && Rphi_out && Rmu_out, VCH_E_SHAPE,
                    "residual: null array");
        VCH_REQUIRE(!c->slab, VCH_E_ARG, "residual: test-level entry, not available in slab mode");
        StreamScope scope(c);
        const long long n = c->g.n;
        Stager st(c->stream, mem);
        const double *a = st.in(phi_new, n), *b = st.in(phi_old, n), *m1 = st.in(mu_new, n), *m0 = st.in(mu_old, n),
                     *w1 = st.in(w_new, n), *w0 = st.in(w_old, n);
        double *rp = st.out(Rphi_out, n), *rm = st.out(Rmu_out, n);
        LAUNCH(c, residual_full_kernel, c->eb(), 256, a, b, m1, m0, w1, w0, rp, rm, c->g, c->ph, dt);
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch2d_jacobian_solve(vch2d_ctx* c, const double* phi, double dt, const double* Rphi, const double* Rmu,
                         double* dphi_out, double* dmu_out, int* its_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi && Rphi && Rmu && dphi_out && dmu_out, VCH_E_SHAPE, "jacobian_solve: null array");
        StreamScope scope(c);
        const long long n = c->g.n;
        Stager st(c->stream, mem);
        const double *dp = st.in(phi, n), *rp = st.in(Rphi, n), *rm = st.in(Rmu, n);
        double *o1 = st.out(dphi_out, n), *o2 = st.out(dmu_out, n);
        LAUNCH(c, jac_diag_kernel, c->rb(), kRedThreads, dp, c->a.p, c->g, c->ph, dt, c->sc, c->red.part, c->ticket);
        vch_stats s{};
        const StatMark mark0 = stat_mark(c);
        if (c->slab) {   // R_phi feeds a stencil: it has to live in a ghosted work vector
            VCH_CUDA(cudaMemcpyAsync(c->Rphi.p, rp, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            rp = c->Rphi.p;
        }
        newton_linear_solve(c, rp, rm, c->a.p, nullptr, dt, &s);
        VCH_CUDA(cudaMemcpyAsync(o1, c->dphi(), n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(o2, c->dmu.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        stat_collect(c, mark0, &s);
        if (its_out) *its_out = (int)s.krylov_iterations;
        st.finish();
        VCH_REQUIRE(s.krylov_stalls == 0, VCH_E_KRYLOV, "jacobian_solve: BiCGStab stopped above tolerance");
        return VCH_OK;
    });
}

int vch2d_newton(vch2d_ctx* c, const double* phi_old, const double* mu_old, const double* w_old, const double* w_new,
                 double dt, double* phi_new_out, double* mu_new_out, double* res_hist, int hist_cap, int* n_hist,
                 vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi_old && mu_old && w_old && w_new && phi_new_out && mu_new_out, VCH_E_SHAPE, "newton: null array");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        Stager st(c->stream, mem);
        const double *p0 = st.in(phi_old, n), *m0 = st.in(mu_old, n), *w0 = st.in(w_old, n), *w1 = st.in(w_new, n);
        double *po = st.out(phi_new_out, n), *mo = st.out(mu_new_out, n);
        std::vector<double> hist;
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        newton_step(c, p0, m0, w0, w1, dt, &hist, s);
        VCH_CUDA(cudaMemcpyAsync(po, c->phi.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(mo, c->mu.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        st.finish();
        if (n_hist) *n_hist = (int)hist.size();
        if (res_hist) for (int i = 0; i < (int)hist.size() && i < hist_cap; ++i) res_hist[i] = hist[i];
        stat_collect(c, mark0, s);
        return VCH_OK;
    });
}

int vch2d_forward(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int n_steps, const double* dt_steps,
                  double* phi_hist_out, double* mu_hist_out, double* w_hist_out, vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi0 && phi_hist_out && dt_steps && n_steps >= 0, VCH_E_SHAPE, "forward: bad arguments");
        VCH_REQUIRE(!u || u_rows >= 1, VCH_E_SHAPE, "forward: control needs at least one row");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        Stager st(c->stream, mem);
        const double* dphi0 = st.in(phi0, n);
        const double* du = st.in(u, (size_t)u_rows * n);
        double* dh = st.out(phi_hist_out, (size_t)(n_steps + 1) * n);
        double* dm = st.out(mu_hist_out, (size_t)n_steps * n);
        double* dw = st.out(w_hist_out, (size_t)n_steps * n);
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        forward_dev(c, dphi0, du, u_rows, n_steps, dt_steps, dh, dm, dw, s);
        st.finish();
        stat_collect(c, mark0, s);
        return VCH_OK;
    });
}

int vch2d_adjoint(vch2d_ctx* c, const double* phi_hist, int levels, const double* t_hist, double b1, double b2,
                  const double* phiQ, const double* phiT, double* p_out, double* q_out, double* r_out, vch_stats* stats,
                  int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi_hist && t_hist && r_out && levels >= 1, VCH_E_SHAPE, "adjoint: bad arguments");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        const size_t tot = (size_t)levels * n;
        Stager st(c->stream, mem);
        const double *dh = st.in(phi_hist, tot), *dq = st.in(phiQ, tot), *dT = st.in(phiT, n);
        double *po = st.out(p_out, tot), *qo = st.out(q_out, tot), *ro = st.out(r_out, tot);
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        adjoint_dev(c, dh, levels, t_hist, b1, b2, dq, dT, po, qo, ro, s);
        st.finish();
        stat_collect(c, mark0, s);
        require_adjoint_converged(c, mark0);
        return VCH_OK;
    });
}

int vch2d_cost(vch2d_ctx* c, const double* phi_hist, const double* u, const double* phiQ, const double* phiT, int levels,
               const double* x, const double* y, const double* t_hist, double b1, double b2, double b3, double ksp,
               double* J_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi_hist && x && y && t_hist && J_out && levels >= 1, VCH_E_SHAPE, "cost: bad arguments");
        StreamScope scope(c);
        const long long n = c->g.n;
        const size_t tot = (size_t)levels * n;
        Stager st(c->stream, mem);
        const double *dh = st.in(phi_hist, tot), *du = st.in(u, tot), *dq = st.in(phiQ, tot), *dT = st.in(phiT, n);
        cost_dev(c, dh, du, dq, dT, levels, x, y, t_hist, b1, b2, b3, ksp, J_out);
        st.finish();
        return VCH_OK;
    });
}

int vch_grad_prox(void* stream, long long count, const double* u, const double* r, double b3, double alpha, double ksp,
                  double umin, double umax, double* grad_out, double* u_new_out, double* red_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(count > 0 && u && r && u_new_out, VCH_E_SHAPE, "grad_prox: bad arguments");
        VCH_REQUIRE(vch_device_count() > 0, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        cudaStream_t s = (cudaStream_t)stream;
        g_scratch.ensure();
        Stager st(s, mem);
        const double *du = st.in(u, count), *dr = st.in(r, count);
        double *dg = st.out(grad_out, count), *dn = st.out(u_new_out, count);
        grad_prox_kernel<<<red_blocks(count), kRedThreads, 0, s>>>(du, dr, dg, dn, count, b3, alpha, ksp, umin, umax,
                                                                    g_scratch.out, g_scratch.part, g_scratch.ticket);
        ++g_free_launches;
        VCH_CUDA(cudaGetLastError());
        VCH_CUDA(cudaMemcpyAsync(g_scratch.out_host, g_scratch.out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
        st.finish();
        VCH_CUDA(cudaStreamSynchronize(s));
        if (red_out) for (int k = 0; k < 4; ++k) red_out[k] = g_scratch.out_host[k];
        return VCH_OK;
    });
}

int vch_kkt_counts(void* stream, long long count, const double* u, const double* r, double ksp, double tol,
                   long long* counts_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(count > 0 && u && r && counts_out, VCH_E_SHAPE, "kkt_counts: bad arguments");
        VCH_REQUIRE(vch_device_count() > 0, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        cudaStream_t s = (cudaStream_t)stream;
        g_scratch.ensure();
        Stager st(s, mem);
        const double *du = st.in(u, count), *dr = st.in(r, count);
        kkt_kernel<<<red_blocks(count), kRedThreads, 0, s>>>(du, dr, count, ksp, tol, g_scratch.out, g_scratch.part,
                                                              g_scratch.ticket);
        ++g_free_launches;
        VCH_CUDA(cudaGetLastError());
        VCH_CUDA(cudaMemcpyAsync(g_scratch.out_host, g_scratch.out, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        st.finish();
        VCH_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < 3; ++k) counts_out[k] = (long long)(g_scratch.out_host[k] + 0.5);
        return VCH_OK;
    });
}

int vch_free_energy(void* stream, int n0, int n1, const double* phi, const double* w, double kappa, double c1, double c2,
                    double h1, double h0, double eps, double* E_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(n0 >= 1 && n1 >= 1 && phi && E_out, VCH_E_SHAPE, "free_energy: bad arguments");
        VCH_REQUIRE(vch_device_count() > 0, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        cudaStream_t s = (cudaStream_t)stream;
        g_scratch.ensure();
        const long long count = (long long)n0 * n1;
        Stager st(s, mem);
        const double *dp = st.in(phi, count), *dw = st.in(w, count);
        energy_kernel<<<red_blocks(count), kRedThreads, 0, s>>>(dp, dw, n0, n1, c1, c2, eps, g_scratch.out, g_scratch.part, g_scratch.ticket);
        ++g_free_launches;
        VCH_CUDA(cudaGetLastError());
        VCH_CUDA(cudaMemcpyAsync(g_scratch.out_host, g_scratch.out, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
        st.finish();
        VCH_CUDA(cudaStreamSynchronize(s));
        const double* o = g_scratch.out_host;     // h1: spacing along the contiguous axis (x), h0: along the other (y)
        E_out[0] = kappa / (2.0 * h1) * o[0] * h0 + kappa / (2.0 * h0) * o[1] * h1 + h1 * h0 * o[2] - (w ? h1 * h0 * o[3] : 0.0);
        return VCH_OK;
    });
}

// Host-buffer path of vch2d_pgd_iteration: PCIe traffic is streamed under the compute instead of bracketing it.
//   H2D on a copy stream: phi_T, then (phi_hist, phi_Q) chunks from the LAST level down — the order in which the adjoint
//   sweep consumes them — then u; the sweep waits per chunk on events.  D2H: u_new right after the prox, phi_hist_new chunk
//   by chunk while the forward solve is still producing later levels.  Full-duplex PCIe, both directions overlap compute
//   when the caller's buffers are pinned (pageable buffers still work: the copies just stop overlapping).
static void pgd_iteration_host_streamed(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps, const double* x,
                                        const double* y, const double* u, const double* phi_hist, const double* phiQ,
                                        const double* phiT, double b1, double b2, double b3, double ksp, double umin,
                                        double umax, double alpha, double* u_new_out, double* phi_hist_out, double* r_out,
                                        double* J_out, double* red_out, vch_stats* s) {
    const long long n = c->g.n;
    const size_t fb = (size_t)n * sizeof(double), tot = (size_t)levels * n;
    DevBuf &du = c->stage[0], &dh = c->stage[1], &dq = c->stage[2], &dT = c->stage[3], &dun = c->stage[4], &dhn = c->stage[5],
           &dr = c->stage[6];
    du.alloc(tot); dh.alloc(tot); dun.alloc(tot); dhn.alloc(tot); dr.alloc(tot);
    if (phiQ) dq.alloc(tot);
    if (phiT) dT.alloc(n);
    cudaStream_t cp;
    VCH_CUDA(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
    // levels per chunk (269 MB at 1024^2).  Measured with 8 / 32 / 128: no difference (e2e 0.4935 / 0.4942 / 0.4863 it/s); what the
    // host-buffer path loses against the device-resident one (~0.11 s of 1.9 s) is the kernels slowing down while the copy engines
    // move the trajectories through the memory system (48 ms of it vanish when the D2H copies are left out), not waiting.
    const int CH = getenv("VCH_STREAM_CHUNK") ? std::max(1, atoi(getenv("VCH_STREAM_CHUNK"))) : 32;
    const int nch = (levels + CH - 1) / CH;
    std::vector<cudaEvent_t> ev_in(nch), ev_out(nch);
    cudaEvent_t ev_u, ev_prox, ev_adj;
    for (auto* v : {&ev_in, &ev_out}) for (auto& e : *v) VCH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&ev_u, &ev_prox, &ev_adj}) VCH_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    auto cleanup = [&] {
        cudaStreamSynchronize(cp); cudaStreamSynchronize(c->stream);
        for (auto* v : {&ev_in, &ev_out}) for (auto& e : *v) cudaEventDestroy(e);
        cudaEventDestroy(ev_u); cudaEventDestroy(ev_prox); cudaEventDestroy(ev_adj);
        cudaStreamDestroy(cp);
    };
    cudaEvent_t tph[4] = {nullptr, nullptr, nullptr, nullptr};           // VCH_DEBUG: phase times on the work stream
    if (c->debug) for (auto& e : tph) cudaEventCreate(&e);
    const auto host_t0 = std::chrono::steady_clock::now();
    try {
        // ---- enqueue every H2D copy up front, in consumption order
        if (c->debug) cudaEventRecord(tph[0], c->stream);
        VCH_CUDA(cudaEventRecord(c->ev_in, c->stream));                 // buffers above were allocated; order cp after entry
        VCH_CUDA(cudaStreamWaitEvent(cp, c->ev_in, 0));
        if (phiT) VCH_CUDA(cudaMemcpyAsync(dT.p, phiT, fb, cudaMemcpyHostToDevice, cp));
        for (int j = nch - 1; j >= 0; --j) {
            const size_t lo = (size_t)j * CH, cnt = std::min<size_t>(CH, levels - lo);
            VCH_CUDA(cudaMemcpyAsync(dh.p + lo * n, phi_hist + lo * n, cnt * fb, cudaMemcpyHostToDevice, cp));
            if (phiQ) VCH_CUDA(cudaMemcpyAsync(dq.p + lo * n, phiQ + lo * n, cnt * fb, cudaMemcpyHostToDevice, cp));
            VCH_CUDA(cudaEventRecord(ev_in[j], cp));
        }
        std::vector<cudaEvent_t> ev_uc(nch);
        for (auto& e : ev_uc) VCH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int j = 0; j < nch; ++j) {                     // u in ASCENDING chunks: consumed just in time by the forward sweep
            const size_t lo = (size_t)j * CH, cnt = std::min<size_t>(CH, levels - lo);
            VCH_CUDA(cudaMemcpyAsync(du.p + lo * n, u + lo * n, cnt * fb, cudaMemcpyHostToDevice, cp));
            VCH_CUDA(cudaEventRecord(ev_uc[j], cp));
        }
        // ---- (1) adjoint sweep, waiting chunk by chunk
        int waited = nch;
        auto need = [&](int k) {
            const int j = k / CH;
            while (waited > j) { --waited; VCH_CUDA(cudaStreamWaitEvent(c->stream, ev_in[waited], 0)); }
        };
        adjoint_dev(c, dh.p, levels, t_hist, b1, b2, phiQ ? dq.p : nullptr, phiT ? dT.p : nullptr, nullptr, nullptr, dr.p, s, need);
        need(0);
        if (c->debug) cudaEventRecord(tph[1], c->stream);
        VCH_CUDA(cudaEventRecord(ev_adj, c->stream));
        if (r_out) {
            VCH_CUDA(cudaStreamWaitEvent(cp, ev_adj, 0));
            VCH_CUDA(cudaMemcpyAsync(r_out, dr.p, tot * sizeof(double), cudaMemcpyDeviceToHost, cp));
        }
        // ---- (2)+(3) gradient/prox chunk by chunk just ahead of the forward sweep; u_new and the trajectory stream out behind it
        VCH_CUDA(cudaMemsetAsync(c->out4 + 4, 0, 4 * sizeof(double), c->stream));
        int proxed = 0;                                      // chunks of u_new already produced
        auto prox_upto = [&](int level) {
            const int jmax = std::min(nch - 1, level / CH);
            while (proxed <= jmax) {
                const int j = proxed++;
                const size_t lo = (size_t)j * CH, cnt = std::min<size_t>(CH, levels - lo);
                VCH_CUDA(cudaStreamWaitEvent(c->stream, ev_uc[j], 0));
                LAUNCH(c, grad_prox_kernel, red_blocks((long long)(cnt * n)), kRedThreads, du.p + lo * n, dr.p + lo * n, (double*)nullptr,
                       dun.p + lo * n, (long long)(cnt * n), b3, alpha, ksp, umin, umax, c->out4 + 4, c->red.part, c->ticket, 1);
                VCH_CUDA(cudaEventRecord(ev_prox, c->stream));
                VCH_CUDA(cudaStreamWaitEvent(cp, ev_prox, 0));
                VCH_CUDA(cudaMemcpyAsync(u_new_out + lo * n, dun.p + lo * n, cnt * fb, cudaMemcpyDeviceToHost, cp));
            }
        };
        auto before = [&](int step) { prox_upto(step + 1); };
        auto after = [&](int k) {
            if ((k + 1) % CH == 0 || k == levels - 1) {
                const int j = k / CH;
                const size_t lo = (size_t)j * CH, cnt = (size_t)k + 1 - lo;
                VCH_CUDA(cudaEventRecord(ev_out[j], c->stream));
                VCH_CUDA(cudaStreamWaitEvent(cp, ev_out[j], 0));
                VCH_CUDA(cudaMemcpyAsync(phi_hist_out + lo * n, dhn.p + lo * n, cnt * fb, cudaMemcpyDeviceToHost, cp));
            }
        };
        forward_dev(c, dh.p, dun.p, levels, levels - 1, dt_steps, dhn.p, nullptr, nullptr, s, after, before);
        prox_upto(levels - 1);
        if (c->debug) cudaEventRecord(tph[2], c->stream);
        for (auto& e : ev_uc) cudaEventDestroy(e);
        // ---- (4) cost
        cost_dev(c, dhn.p, dun.p, phiQ ? dq.p : nullptr, phiT ? dT.p : nullptr, levels, x, y, t_hist, b1, b2, b3, ksp, J_out);
        if (red_out) {
            VCH_CUDA(cudaMemcpyAsync(c->out4_host + 4, c->out4 + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            VCH_CUDA(cudaStreamSynchronize(c->stream));
            for (int k = 0; k < 4; ++k) red_out[k] = c->out4_host[4 + k];
        }
        if (c->debug) cudaEventRecord(tph[3], c->stream);
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        const auto host_t1 = std::chrono::steady_clock::now();
        VCH_CUDA(cudaStreamSynchronize(cp));
        if (c->debug) {
            float a = 0, f = 0, k = 0;
            cudaEventElapsedTime(&a, tph[0], tph[1]); cudaEventElapsedTime(&f, tph[1], tph[2]); cudaEventElapsedTime(&k, tph[2], tph[3]);
            const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - host_t0).count();
            const double work = std::chrono::duration<double>(host_t1 - host_t0).count();
            fprintf(stderr, "[vch2d] streamed host path: adjoint %.3f s, prox+forward %.3f s, cost %.3f s on the work stream; work stream done "
                            "at %.3f s, copies drained at %.3f s\n", a * 1e-3, f * 1e-3, k * 1e-3, work, wall);
        }
    } catch (...) { cleanup(); if (c->debug) for (auto& e : tph) cudaEventDestroy(e); throw; }
    cleanup();
    if (c->debug) for (auto& e : tph) cudaEventDestroy(e);
}

// Bounded-memory variant of the host-buffer path: no trajectory ever exists on the device in full.  Every array is
// walked in chunks of CH levels through a ring of 3 chunk slots, in the order the sweeps consume / produce it:
//   adjoint sweep (levels descending):  phi_hist, phi_Q chunks in;  r chunks out to the host
//   prox + forward sweep (ascending):   u, r chunks in -> u_new chunk (out, and input of the time steps) -> phi_hist_new
//                                       chunk out, with the cost integrals accumulated per chunk (phi_Q chunks in again)
// so trajectories larger than HBM (2048^2 x 1000 on one GPU, 4096^2 x 3000 on eight) only need host memory and PCIe time.
// A slot is recycled once its last consumer (kernel or copy) has been passed by an event; three slots suffice because
// every stage touches at most two neighbouring chunks at a time.
static void pgd_iteration_host_bounded(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps, const double* x,
                                       const double* y, const double* u, const double* phi_hist, const double* phiQ,
                                       const double* phiT, double b1, double b2, double b3, double ksp, double umin,
                                       double umax, double alpha, double* u_new_out, double* phi_hist_out, double* r_out,
                                       double* J_out, double* red_out, vch_stats* s, size_t budget_bytes) {
    const long long n = c->g.n;
    const size_t fb = (size_t)n * sizeof(double), tot = (size_t)levels * n;
    const int D = 3;
    const int nrings = phiQ ? 6 : 5;
    int CH = (int)std::min<size_t>((size_t)levels, std::max<size_t>(2, budget_bytes / ((size_t)nrings * D * fb)));
    const int nch = (levels + CH - 1) / CH;
    DevBuf &rU = c->stage[0], &rH = c->stage[1], &rQ = c->stage[2], &dT = c->stage[3], &rUN = c->stage[4], &rHN = c->stage[5],
           &rR = c->stage[6];
    const size_t ring = (size_t)D * CH * n;
    rU.alloc(ring); rH.alloc(ring); rUN.alloc(ring); rHN.alloc(ring); rR.alloc(ring);
    if (phiQ) rQ.alloc(ring);
    if (phiT) dT.alloc(n);
    auto at = [&](DevBuf& b, int level) { const int j = level / CH; return b.p + ((size_t)(j % D) * CH + (level - j * CH)) * n; };
    auto lo_of = [&](int j) { return (size_t)j * CH; };
    auto cnt_of = [&](int j) { return std::min<size_t>(CH, levels - lo_of(j)); };
    double* rh = r_out;
    if (!rh) { c->r_host.resize(tot); rh = c->r_host.data(); }

    cudaStream_t cp;
    VCH_CUDA(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
    enum { InA, FreeA, Rdone, Rsaved, InB, Prox, UNsaved, InQ, Cost, HNsaved, NKINDS };
    std::vector<cudaEvent_t> ev((size_t)NKINDS * nch);
    for (auto& e : ev) VCH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaEvent_t ev_misc[2];
    for (auto& e : ev_misc) VCH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    auto E = [&](int kind, int j) { return ev[(size_t)kind * nch + j]; };
    auto rec = [&](int kind, int j, cudaStream_t st) { VCH_CUDA(cudaEventRecord(E(kind, j), st)); };
    auto wait = [&](cudaStream_t st, int kind, int j) { VCH_CUDA(cudaStreamWaitEvent(st, E(kind, j), 0)); };
    auto cleanup = [&] {
        cudaStreamSynchronize(cp); cudaStreamSynchronize(c->stream);
        for (auto& e : ev) cudaEventDestroy(e);
        for (auto& e : ev_misc) cudaEventDestroy(e);
        cudaStreamDestroy(cp);
    };
    try {
        VCH_CUDA(cudaEventRecord(c->ev_in, c->stream));
        VCH_CUDA(cudaStreamWaitEvent(cp, c->ev_in, 0));
        if (phiT) {
            VCH_CUDA(cudaMemcpyAsync(dT.p, phiT, fb, cudaMemcpyHostToDevice, cp));
            VCH_CUDA(cudaEventRecord(ev_misc[0], cp));
            VCH_CUDA(cudaStreamWaitEvent(c->stream, ev_misc[0], 0));
        }
        // ------------------------------------------------------------------ (1) adjoint sweep, levels descending
        int nextLoadA = nch - 1, nextFreeA = nch - 1, waitedA = nch, rDone = nch, rEntered = nch;
        auto loadA = [&](int j) {
            if (j + D < nch) wait(cp, FreeA, j + D);
            VCH_CUDA(cudaMemcpyAsync(at(rH, (int)lo_of(j)), phi_hist + lo_of(j) * n, cnt_of(j) * fb, cudaMemcpyHostToDevice, cp));
            if (phiQ) VCH_CUDA(cudaMemcpyAsync(at(rQ, (int)lo_of(j)), phiQ + lo_of(j) * n, cnt_of(j) * fb, cudaMemcpyHostToDevice, cp));
            rec(InA, j, cp);
        };
        auto save_r = [&](int j) {   // chunk j of r is complete on the device
            rec(Rdone, j, c->stream); wait(cp, Rdone, j);
            VCH_CUDA(cudaMemcpyAsync(rh + lo_of(j) * n, at(rR, (int)lo_of(j)), cnt_of(j) * fb, cudaMemcpyDeviceToHost, cp));
            rec(Rsaved, j, cp);
        };
        auto needA = [&](int k) {    // called before level k (which reads levels k and k + 1) is processed
            const int jk = k / CH;
            while (nextFreeA >= 0 && lo_of(nextFreeA) > (size_t)k + 1) { rec(FreeA, nextFreeA, c->stream); --nextFreeA; }
            while (rDone - 1 > jk) { --rDone; save_r(rDone); }
            while (nextLoadA >= 0 && nextLoadA >= jk - 1) { loadA(nextLoadA); --nextLoadA; }
            while (waitedA > jk) { --waitedA; wait(c->stream, InA, waitedA); }
            while (rEntered > jk) { --rEntered; if (rEntered + D < nch) wait(c->stream, Rsaved, rEntered + D); }
        };
        LevelMap lmA;
        lmA.hist = [&](int k) { return at(rH, k); };
        if (phiQ) lmA.Q = [&](int k) { return at(rQ, k); };
        lmA.r = [&](int k) { return at(rR, k); };
        adjoint_dev(c, nullptr, levels, t_hist, b1, b2, nullptr, phiT ? dT.p : nullptr, nullptr, nullptr, nullptr, s, needA, &lmA);
        while (rDone > 0) { --rDone; save_r(rDone); }
        VCH_CUDA(cudaEventRecord(ev_misc[1], c->stream));           // phase (2) re-uses the Q and r rings
        VCH_CUDA(cudaStreamWaitEvent(cp, ev_misc[1], 0));

        // ------------------------------------------------------------------ (2) prox + forward sweep + cost, levels ascending
        const CostWeights cw = cost_weights(c, levels, x, y, t_hist);
        VCH_CUDA(cudaMemsetAsync(c->out4, 0, 8 * sizeof(double), c->stream));
        int nextLoadB = 0, proxed = 0, nextLoadQ = 0, hnEntered = 0;
        auto loadB = [&](int j) {
            if (j - D >= 0) wait(cp, Prox, j - D);
            VCH_CUDA(cudaMemcpyAsync(at(rU, (int)lo_of(j)), u + lo_of(j) * n, cnt_of(j) * fb, cudaMemcpyHostToDevice, cp));
            VCH_CUDA(cudaMemcpyAsync(at(rR, (int)lo_of(j)), rh + lo_of(j) * n, cnt_of(j) * fb, cudaMemcpyHostToDevice, cp));
            rec(InB, j, cp);
        };
        auto loadQ = [&](int j) {
            if (!phiQ) return;
            if (j - D >= 0) wait(cp, Cost, j - D);
            VCH_CUDA(cudaMemcpyAsync(at(rQ, (int)lo_of(j)), phiQ + lo_of(j) * n, cnt_of(j) * fb, cudaMemcpyHostToDevice, cp));
            rec(InQ, j, cp);
        };
        auto prox_upto = [&](int level) {
            const int jmax = std::min(nch - 1, level / CH);
            while (proxed <= jmax) {
                const int j = proxed++;
                while (nextLoadB < nch && nextLoadB <= j + 1) { loadB(nextLoadB); ++nextLoadB; }
                wait(c->stream, InB, j);
                if (j - D >= 0) wait(c->stream, UNsaved, j - D);
                const long long cnt = (long long)cnt_of(j) * n;
                LAUNCH(c, grad_prox_kernel, red_blocks(cnt), kRedThreads, at(rU, (int)lo_of(j)), at(rR, (int)lo_of(j)), (double*)nullptr,
                       at(rUN, (int)lo_of(j)), cnt, b3, alpha, ksp, umin, umax, c->out4 + 4, c->red.part, c->ticket, 1);
                rec(Prox, j, c->stream); wait(cp, Prox, j);
                VCH_CUDA(cudaMemcpyAsync(u_new_out + lo_of(j) * n, at(rUN, (int)lo_of(j)), cnt_of(j) * fb, cudaMemcpyDeviceToHost, cp));
                rec(UNsaved, j, cp);
            }
        };
        auto before = [&](int step) {
            prox_upto(step + 1);
            const int j = (step + 1) / CH;       // slot that receives level step + 1
            while (hnEntered <= j) { if (hnEntered - D >= 0) wait(c->stream, HNsaved, hnEntered - D); ++hnEntered; }
        };
        auto after = [&](int k) {                // level k of the new trajectory has been enqueued
            if ((k + 1) % CH != 0 && k != levels - 1) return;
            const int j = k / CH;
            while (nextLoadQ < nch && nextLoadQ <= j + 1) { loadQ(nextLoadQ); ++nextLoadQ; }
            if (phiQ) wait(c->stream, InQ, j);
            const int cnt = (int)cnt_of(j);
            LAUNCH(c, cost_kernel, red_blocks((long long)cnt * n), kRedThreads, at(rHN, (int)lo_of(j)), at(rUN, (int)lo_of(j)),
                   phiQ ? at(rQ, (int)lo_of(j)) : (const double*)nullptr, phiT ? dT.p : (const double*)nullptr, cnt, c->g.nx1, c->g.ny1,
                   cw.wt + lo_of(j), cw.wx, cw.wy, c->out4, c->red.part, c->ticket, (k == levels - 1) ? cnt - 1 : -1, 1);
            rec(Cost, j, c->stream); wait(cp, Cost, j);
            VCH_CUDA(cudaMemcpyAsync(phi_hist_out + lo_of(j) * n, at(rHN, (int)lo_of(j)), cnt_of(j) * fb, cudaMemcpyDeviceToHost, cp));
            rec(HNsaved, j, cp);
        };
        loadQ(0); nextLoadQ = 1;
        LevelMap lmB;
        lmB.u = [&](int k) { return at(rUN, k); };
        lmB.hist_new = [&](int k) { return at(rHN, k); };
        hnEntered = 1;                                               // level 0 goes into a fresh slot
        forward_dev(c, at(rH, 0), nullptr, levels, levels - 1, dt_steps, nullptr, nullptr, nullptr, s, after, before, &lmB);
        prox_upto(levels - 1);
        if (levels == 1) after(0);
        cost_finish(c, b1, b2, b3, ksp, J_out);
        if (red_out) {
            VCH_CUDA(cudaMemcpyAsync(c->out4_host + 4, c->out4 + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            VCH_CUDA(cudaStreamSynchronize(c->stream));
            for (int k = 0; k < 4; ++k) red_out[k] = c->out4_host[4 + k];
        }
        VCH_CUDA(cudaStreamSynchronize(cp));
        VCH_CUDA(cudaStreamSynchronize(c->stream));
    } catch (...) { cleanup(); throw; }
    cleanup();
}

// ---------------------------------------------------------------------------------------------- checkpointed sweeps
// The state trajectory exists only as checkpoints (phi, mu, w) every S levels.  The adjoint sweep walks the segments from the
// last to the first: each is recomputed forward from its checkpoint into a buffer of S+1 levels (the same kernels in the same
// order as the sweep that produced the checkpoints, deterministic reductions: bit-identical levels), then its adjoint levels
// follow, then the gradient/prox of those levels — so neither phi_hist nor (unless asked for) r ever exists in full.  The forward
// sweep under the new control writes the new checkpoints and accumulates the cost integrals segment by segment.  Price: one
// extra forward sweep per iteration.  Memory: 3 x 2 x (M/S + 1) + 2 (S + 1) levels instead of 3 (M + 1).
static int ckpt_count(int levels, int S) { return (levels - 1 + S - 1) / S + 1; }
static int ckpt_level(int j, int levels, int S) { return std::min(j * S, levels - 1); }

static void ckpt_load_state(vch2d_ctx* c, const double* ck_mu_j, const double* ck_w_j) {
    const size_t n = (size_t)c->g.n;
    c->mu_old.alloc(n);
    dev_copy(c, c->mu_old.p, ck_mu_j, n);
    halo_push(c, c->mu_old.p, nullptr, 1);
    dev_copy(c, c->w0.p, ck_w_j, n);
}

// Forward sweep from level 0 that keeps only checkpoints.  seg_done(j, s0, s1, seg): levels s0..s1 of segment j are in seg[0..s1-s0].
static void forward_ckpt_dev(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int levels, const double* dt_steps, int S,
                             double* ck_phi, double* ck_mu, double* ck_w, vch_stats* st,
                             const std::function<void(int, int, int, const double*)>& seg_done = nullptr) {
    const size_t n = (size_t)c->g.n;
    const int M = levels - 1, ncp = ckpt_count(levels, S);
    c->ck_seg.alloc((size_t)(S + 1) * n);
    double* seg = c->ck_seg.p;
    dev_copy(c, seg, phi0, n);
    forward_begin(c, seg, true);
    dev_copy(c, ck_phi, seg, n); dev_copy(c, ck_mu, c->mu_old.p, n); dev_copy(c, ck_w, c->w0.p, n);
    for (int j = 0; j + 1 < ncp; ++j) {
        const int s0 = j * S, s1 = std::min(s0 + S, M);
        auto HN = [&](int k) -> double* { return seg + (size_t)(k - s0) * n; };
        auto U = [&](int k) -> const double* { return u + (size_t)k * n; };
        forward_steps(c, s0, s1, HN, U, u != nullptr, u_rows, dt_steps, nullptr, nullptr, st, nullptr, nullptr);
        dev_copy(c, ck_phi + (size_t)(j + 1) * n, HN(s1), n);
        dev_copy(c, ck_mu + (size_t)(j + 1) * n, c->mu_old.p, n);
        dev_copy(c, ck_w + (size_t)(j + 1) * n, c->w0.p, n);
        if (seg_done) seg_done(j, s0, s1, seg);
        if (j + 2 < ncp) dev_copy(c, seg, HN(s1), n);          // first level of the next segment
    }
    if (ncp == 1 && seg_done) seg_done(0, 0, 0, seg);
    VCH_CUDA(cudaGetLastError());
}

static void pgd_iteration_ckpt_dev(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps, const double* x,
                                   const double* y, const double* u, const double* ck_phi, const double* ck_mu, const double* ck_w,
                                   int S, const double* phiQ, const double* phiT, double b1, double b2, double b3, double ksp,
                                   double umin, double umax, double alpha, double* u_new, double* ck_phi_out, double* ck_mu_out,
                                   double* ck_w_out, double* r_out, double* J_out, double* red_out, vch_stats* st) {
    const size_t n = (size_t)c->g.n;
    const int M = levels - 1, ncp = ckpt_count(levels, S);
    c->ck_seg.alloc((size_t)(S + 1) * n);
    if (!r_out) c->ck_r.alloc((size_t)(S + 1) * n);
    double* seg = c->ck_seg.p;
    auto R = [&](int k) -> double* { return r_out ? r_out + (size_t)k * n : c->ck_r.p + (size_t)(k % (S + 1)) * n; };
    forward_begin(c, ck_phi, false);                             // reference mass of the OLD trajectory (its level 0)
    VCH_CUDA(cudaMemsetAsync(c->out4 + 4, 0, 4 * sizeof(double), c->stream));
    auto prox_level = [&](int k) {
        LAUNCH(c, grad_prox_kernel, red_blocks((long long)n), kRedThreads, u + (size_t)k * n, R(k), (double*)nullptr, u_new + (size_t)k * n,
               (long long)n, b3, alpha, ksp, umin, umax, c->out4 + 4, c->red.part, c->ticket, 1);
    };
    if (ncp == 1) {   // a single level: terminal adjoint only
        LevelMap lm; lm.hist = [&](int) { return const_cast<double*>(ck_phi); }; lm.r = R;
        if (phiQ) lm.Q = [&](int k) { return const_cast<double*>(phiQ) + (size_t)k * n; };
        adjoint_dev(c, nullptr, levels, t_hist, b1, b2, nullptr, phiT, nullptr, nullptr, nullptr, st, nullptr, &lm, M, M);
        prox_level(M);
    }
    for (int j = ncp - 2; j >= 0; --j) {
        const int s0 = j * S, s1 = std::min(s0 + S, M);
        auto HN = [&](int k) -> double* { return seg + (size_t)(k - s0) * n; };
        auto U = [&](int k) -> const double* { return u + (size_t)k * n; };
        // (1a) recompute the segment from its checkpoint
        dev_copy(c, seg, ck_phi + (size_t)j * n, n);
        ckpt_load_state(c, ck_mu + (size_t)j * n, ck_w + (size_t)j * n);
        forward_steps(c, s0, s1, HN, U, true, levels, dt_steps, nullptr, nullptr, st, nullptr, nullptr);
        // (1b) its adjoint levels (the rolling p, q, r of level s1 are still in the rings / in R(s1))
        LevelMap lm;
        lm.hist = [&](int k) { return HN(k); };
        if (phiQ) lm.Q = [&](int k) { return const_cast<double*>(phiQ) + (size_t)k * n; };
        lm.r = R;
        adjoint_dev(c, nullptr, levels, t_hist, b1, b2, nullptr, phiT, nullptr, nullptr, nullptr, st, nullptr, &lm, s1, s0);
        // (2) gradient + prox of the levels whose r is now final
        for (int k = (s1 == M ? M : s1 - 1); k >= s0; --k) prox_level(k);
    }
    // (3) forward solve under the new control: new checkpoints, cost integrals accumulated per segment       GD2_configured.py:309-312
    const CostWeights cw = cost_weights(c, levels, x, y, t_hist);
    VCH_CUDA(cudaMemsetAsync(c->out4, 0, 4 * sizeof(double), c->stream));
    auto seg_cost = [&](int j, int s0, int s1, const double* sg) {
        const bool last = (s1 == M);
        const int cnt = (s1 - s0) + (last ? 1 : 0);              // levels s0 .. s1-1, plus the terminal level in the last segment
        if (cnt <= 0) return;
        LAUNCH(c, cost_kernel, red_blocks((long long)cnt * (long long)n), kRedThreads, sg, u_new + (size_t)s0 * n,
               phiQ ? phiQ + (size_t)s0 * n : (const double*)nullptr, phiT, cnt, c->g.nx1, c->g.ny1, cw.wt + s0, cw.wx, cw.wy, c->out4,
               c->red.part, c->ticket, last ? cnt - 1 : -1, 1);
        (void)j;
    };
    forward_ckpt_dev(c, ck_phi, u_new, levels, levels, dt_steps, S, ck_phi_out, ck_mu_out, ck_w_out, st, seg_cost);
    cost_finish(c, b1, b2, b3, ksp, J_out);
    if (red_out) {
        VCH_CUDA(cudaMemcpyAsync(c->out4_host + 4, c->out4 + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 4; ++k) red_out[k] = c->out4_host[4 + k];
    }
    VCH_CUDA(cudaStreamSynchronize(c->stream));
}

int vch2d_forward_ckpt(vch2d_ctx* c, const double* phi0, const double* u, int u_rows, int n_steps, const double* dt_steps,
                       int ckpt_stride, double* ck_phi_out, double* ck_mu_out, double* ck_w_out, vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && phi0 && dt_steps && n_steps >= 0 && ckpt_stride >= 1 && ck_phi_out && ck_mu_out && ck_w_out, VCH_E_SHAPE,
                    "forward_ckpt: bad arguments");
        VCH_REQUIRE(!u || u_rows >= 1, VCH_E_SHAPE, "forward_ckpt: control needs at least one row");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        const int levels = n_steps + 1, ncp = ckpt_count(levels, ckpt_stride);
        Stager st(c->stream, mem);
        const double* dphi0 = st.in(phi0, n);
        const double* du = st.in(u, (size_t)u_rows * n);
        double *cp = st.out(ck_phi_out, (size_t)ncp * n), *cm = st.out(ck_mu_out, (size_t)ncp * n), *cwv = st.out(ck_w_out, (size_t)ncp * n);
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        forward_ckpt_dev(c, dphi0, du, u_rows, levels, dt_steps, ckpt_stride, cp, cm, cwv, s);
        st.finish();
        stat_collect(c, mark0, s);
        return VCH_OK;
    });
}

int vch2d_pgd_iteration_ckpt(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps, const double* x, const double* y,
                             const double* u, const double* ck_phi, const double* ck_mu, const double* ck_w, int ckpt_stride,
                             const double* phiQ, const double* phiT, double b1, double b2, double b3, double ksp, double umin,
                             double umax, double alpha, double* u_new_out, double* ck_phi_out, double* ck_mu_out, double* ck_w_out,
                             double* r_out, double* J_out, double* red_out, vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && levels >= 1 && t_hist && dt_steps && x && y && u && ck_phi && ck_mu && ck_w && ckpt_stride >= 1 && u_new_out &&
                    ck_phi_out && ck_mu_out && ck_w_out && J_out, VCH_E_SHAPE, "pgd_iteration_ckpt: bad arguments");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        const size_t tot = (size_t)levels * n, ctot = (size_t)ckpt_count(levels, ckpt_stride) * n;
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        Stager st(c->stream, mem);
        const double *du = st.in(u, tot), *cp = st.in(ck_phi, ctot), *cm = st.in(ck_mu, ctot), *cwv = st.in(ck_w, ctot),
                     *dq = st.in(phiQ, tot), *dT = st.in(phiT, n);
        double *dun = st.out(u_new_out, tot), *cpo = st.out(ck_phi_out, ctot), *cmo = st.out(ck_mu_out, ctot), *cwo = st.out(ck_w_out, ctot),
               *dr = st.out(r_out, tot);
        pgd_iteration_ckpt_dev(c, levels, t_hist, dt_steps, x, y, du, cp, cm, cwv, ckpt_stride, dq, dT, b1, b2, b3, ksp, umin, umax, alpha,
                               dun, cpo, cmo, cwo, dr, J_out, red_out, s);
        st.finish();
        stat_collect(c, mark0, s);
        require_adjoint_converged(c, mark0);
        return VCH_OK;
    });
}

int vch2d_pgd_iteration(vch2d_ctx* c, int levels, const double* t_hist, const double* dt_steps, const double* x,
                        const double* y, const double* u, const double* phi_hist, const double* phiQ, const double* phiT,
                        double b1, double b2, double b3, double ksp, double umin, double umax, double alpha,
                        double* u_new_out, double* phi_hist_out, double* r_out, double* J_out, double* red_out,
                        vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && levels >= 2 && t_hist && dt_steps && x && y && u && phi_hist && u_new_out && phi_hist_out && J_out,
                    VCH_E_SHAPE, "pgd_iteration: bad arguments");
        StreamScope scope(c);
        const long long n = c->g.n;
        const StatMark mark0 = stat_mark(c);
        const size_t tot = (size_t)levels * n;
        vch_stats local{}; vch_stats* s = stats ? stats : &local;
        if (mem == VCH_MEM_HOST) {
            // full device staging (5-6 trajectories) when it fits, chunk rings otherwise or when a budget was set
            size_t freeb = 0, totalb = 0, held = 0;
            VCH_CUDA(cudaMemGetInfo(&freeb, &totalb));
            for (auto& b : c->stage) held += b.n * sizeof(double);
            const size_t full = ((size_t)(phiQ ? 6 : 5) * tot + (size_t)n) * sizeof(double);
            const bool bounded = c->stream_budget > 0 || full > (size_t)(0.92 * (double)(freeb + held));
            if (bounded) {
                if (c->stream_budget <= 0)      // the full-size staging of an earlier, smaller call is of no use here
                    for (auto& b : c->stage) if (b.n > 0 && b.n * sizeof(double) > (size_t)(0.05 * (double)totalb)) b.release();
                VCH_CUDA(cudaMemGetInfo(&freeb, &totalb));
                const size_t budget = c->stream_budget > 0 ? (size_t)c->stream_budget : (size_t)(0.6 * (double)freeb);
                pgd_iteration_host_bounded(c, levels, t_hist, dt_steps, x, y, u, phi_hist, phiQ, phiT, b1, b2, b3, ksp, umin, umax,
                                           alpha, u_new_out, phi_hist_out, r_out, J_out, red_out, s, budget);
            } else {
                pgd_iteration_host_streamed(c, levels, t_hist, dt_steps, x, y, u, phi_hist, phiQ, phiT, b1, b2, b3, ksp, umin, umax,
                                            alpha, u_new_out, phi_hist_out, r_out, J_out, red_out, s);
            }
            stat_collect(c, mark0, s);
            require_adjoint_converged(c, mark0);
            return VCH_OK;
        }
        const double *du = u, *dh = phi_hist, *dq = phiQ, *dT = phiT;
        double *dun = u_new_out, *dhn = phi_hist_out, *dr = r_out;
        DevBuf rscratch;
        if (!dr) { rscratch.alloc(tot); dr = rscratch.p; }
        // (1) adjoint sweep over the stored trajectory                       GD2_configured.py:299
        adjoint_dev(c, dh, levels, t_hist, b1, b2, dq, dT, nullptr, nullptr, dr, s);
        // (2) gradient + soft-threshold prox + box, with the driver's norms  GD2_configured.py:304-305, :375
        LAUNCH(c, grad_prox_kernel, red_blocks((long long)tot), kRedThreads, du, dr, (double*)nullptr, dun, (long long)tot, b3,
               alpha, ksp, umin, umax, c->out4 + 4, c->red.part, c->ticket, 0);
        // (3) forward solve under the new control                            GD2_configured.py:309
        forward_dev(c, dh, dun, levels, levels - 1, dt_steps, dhn, nullptr, nullptr, s);
        // (4) cost functional                                                GD2_configured.py:312
        cost_dev(c, dhn, dun, dq, dT, levels, x, y, t_hist, b1, b2, b3, ksp, J_out);
        if (red_out) {
            VCH_CUDA(cudaMemcpyAsync(c->out4_host + 4, c->out4 + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            VCH_CUDA(cudaStreamSynchronize(c->stream));
            for (int k = 0; k < 4; ++k) red_out[k] = c->out4_host[4 + k];
        }
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        stat_collect(c, mark0, s);
        require_adjoint_converged(c, mark0);
        return VCH_OK;
    });
}

}  // extern "C"
