// Shared device/host helpers for the vch_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <utility>
#include <algorithm>
#include "../../include/vch_b200.h"

namespace vch {

// ---------------------------------------------------------------- error plumbing
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);

#define VCH_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw vch::Error(VCH_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

#define VCH_REQUIRE(cond, code, msg)                                                            \
    do { if (!(cond)) throw vch::Error(code, msg); } while (0)

// Wraps the body of every extern "C" entry point.
template <class F>
static inline int guarded(F&& f) {
    try { return f(); }
    catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error(e.what()); return VCH_E_CUDA; }
}

// ---------------------------------------------------------------- device memory
struct DevBuf {
    double* p = nullptr;
    size_t n = 0;
    bool view = false;          // carved out of an Arena (slab mode): not owned
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        VCH_CUDA(cudaMalloc(&p, count * sizeof(double)));
        n = count;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; n = 0; view = false; }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// Device staging buffers of the host-buffer entry points are recycled through a small per-thread pool instead of a
// cudaMalloc / cudaFree pair per array per call (every test-level drop-in call pays for them; cudaFree also synchronises
// the whole device).  At most kStagePoolBytes stay cached; larger requests are freed on return as before.
constexpr size_t kStagePoolBytes = (size_t)2 << 30;
struct StagePool {
    struct Item { double* p; size_t bytes; int dev; };
    std::vector<Item> free_list;
    size_t cached = 0;
    static int cur_dev() { int d = 0; cudaGetDevice(&d); return d; }
    double* get(size_t bytes) {
        int best = -1;
        const int dev = cur_dev();
        for (int i = 0; i < (int)free_list.size(); ++i)
            if (free_list[i].dev == dev && free_list[i].bytes >= bytes && free_list[i].bytes <= 2 * bytes + 4096 &&
                (best < 0 || free_list[i].bytes < free_list[best].bytes)) best = i;
        if (best >= 0) {
            double* p = free_list[best].p;
            cached -= free_list[best].bytes;
            free_list.erase(free_list.begin() + best);
            return p;
        }
        double* d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) {       // make room and retry once
            cudaGetLastError();
            clear();
            VCH_CUDA(cudaMalloc(&d, bytes));
        }
        return d;
    }
    void put(double* p, size_t bytes) {
        if (bytes > kStagePoolBytes / 2 || cached + bytes > kStagePoolBytes || free_list.size() >= 32) { cudaFree(p); return; }
        free_list.push_back({p, bytes, cur_dev()}); cached += bytes;
    }
    void clear() { for (auto& it : free_list) cudaFree(it.p); free_list.clear(); cached = 0; }
    ~StagePool() { /* process exit: the driver reclaims the memory */ }
};
inline StagePool& stage_pool() { static thread_local StagePool pool; return pool; }

// Stages caller arrays when mem == VCH_MEM_HOST; passes device pointers through otherwise.
struct Stager {
    cudaStream_t s;
    int mem;
    struct Own { double* p; size_t bytes; };
    std::vector<Own> owned;
    struct Out { double* host; double* dev; size_t n; };
    std::vector<Out> outs;
    bool drained = true;
    Stager(cudaStream_t st, int m) : s(st), mem(m) {}
    double* take(size_t n) {
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(double);
        double* d = stage_pool().get(bytes);
        owned.push_back({d, bytes});
        drained = false;
        return d;
    }
    const double* in(const double* a, size_t n) {
        if (!a || mem == VCH_MEM_DEVICE) return a;
        double* d = take(n);
        VCH_CUDA(cudaMemcpyAsync(d, a, n * sizeof(double), cudaMemcpyHostToDevice, s));
        return d;
    }
    double* out(double* a, size_t n) {
        if (!a || mem == VCH_MEM_DEVICE) return a;
        double* d = take(n);
        outs.push_back({a, d, n});
        return d;
    }
    void finish() {
        for (auto& o : outs) VCH_CUDA(cudaMemcpyAsync(o.host, o.dev, o.n * sizeof(double), cudaMemcpyDeviceToHost, s));
        VCH_CUDA(cudaStreamSynchronize(s));
        outs.clear();
        drained = true;
    }
    ~Stager() {
        if (!owned.empty() && !drained) cudaStreamSynchronize(s);     // error path: nothing may still be using the buffers
        for (auto& o : owned) stage_pool().put(o.p, o.bytes);
    }
};

// ---------------------------------------------------------------- launch accounting / per-kernel event timing
// Every kernel launch of the library goes through a LaunchLog.  With profiling on, each launch is bracketed by a
// pair of CUDA events on the launching stream; report() folds them into per-kernel totals (bench.py's roofline leg).
struct LaunchLog {
    long long count = 0;
    bool profiling = false;
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; VCH_CUDA(cudaEventCreate(&e)); return e;
    }
    void begin(const char* name, cudaStream_t s) {
        ++count;
        if (!profiling) return;
        Rec r{name, get(), get()};
        VCH_CUDA(cudaEventRecord(r.a, s));
        recs.push_back(r);
    }
    void end(cudaStream_t s) {
        if (!profiling) return;
        VCH_CUDA(cudaEventRecord(recs.back().b, s));
    }
    struct Row { std::string name; double ms = 0; long long n = 0; };
    std::vector<Row> report() {
        std::vector<Row> rows;
        for (auto& r : recs) {
            VCH_CUDA(cudaEventSynchronize(r.b));
            float ms = 0.f; VCH_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
            Row* row = nullptr;
            for (auto& x : rows) if (x.name == r.name) { row = &x; break; }
            if (!row) { rows.push_back(Row{r.name, 0.0, 0}); row = &rows.back(); }
            row->ms += ms; row->n += 1;
            pool.push_back(r.a); pool.push_back(r.b);
        }
        recs.clear();
        return rows;
    }
    ~LaunchLog() { for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } for (auto e : pool) cudaEventDestroy(e); }
};

// Device-resident scalars.  Host reads them through a pinned mirror.
struct Scal {
    double rho, rho_new, alpha, omega;
    double r0v, ts, tt, rr, thr2, bnorm2;
    double res2, amin, amax, abar, mu2;
    double ceil_pos, ceil_neg;
    double rem2;              // || c1 [l(phi + dphi) - l(phi) - l'(phi) dphi] ||^2: the nonlinear remainder of a full Newton step (see newton_step)
    double mass, wint, mass0;
    double tol2;
    double c0, c2;            // operator coefficients of the current solve (read by the graph's kernels: one graph serves every dt)
    int done, iters, nonfinite, maxit;
    long long iters_total, solves, stalls;   // accumulated on the device (graph-driven solves are never polled)
    long long g_launches;                    // kernels that ran inside solve graphs (not seen by the host launch log)
    long long half_exits;                    // solves that stopped after the first half of a BiCGStab iteration (||s|| <= tol ||b||)
    int iters_max, comm_err;                 // comm_err: a bounded cross-rank wait expired (slab mode)
    int half, adj;                           // half = 1: converged at the half step, bicg_x_kernel only applies x += alpha p
                                             // adj = 1: the current solve belongs to the adjoint sweep (set by adj_rhs_kernel)
    long long stalls_adj;                    // stalled solves of the adjoint sweep (no outer residual check guards them: an error)
};


// ---------------------------------------------------------------- slab mode: peer-memory communication (one process per GPU)
// Every rank owns one cudaMalloc'd ARENA with an identical layout; ranks exchange CUDA-IPC handles once and from then
// on address each other's arena directly over NVLink (peer[r] + the same offset).  Three primitives, all device-side:
//   * xrank_reduce  - inside grid_reduce: the last block of a reduction kernel writes its totals into every peer's slot,
//                     waits for the peers' totals and combines them in rank order (deterministic, no extra launch);
//   * xbar_kernel   - one-CTA barrier across ranks (monotone sequence flags in peer memory);
//   * halo_push_kernel - writes this rank's boundary rows into the neighbours' ghost rows.
// Waits are bounded (kSpinLimitNs): a timeout raises *err and lets the kernel finish, so a rank that lost its peers
// fails loudly instead of hanging the GPU.
constexpr int kMaxRanks = 8;
constexpr int kRedK = 8;                                   // values per cross-rank reduction (max K of grid_reduce)
constexpr size_t kSlotOff = 0;                             // [2 parities][kMaxRanks][kRedK + 1 (tag)] doubles
constexpr size_t kFlagOff = 2 * kMaxRanks * (kRedK + 1);   // [kMaxRanks] barrier flags (unsigned long long)
constexpr size_t kArenaHeader = 256;                       // doubles reserved at the start of every arena
constexpr unsigned long long kSpinLimitNs = 10ull * 1000ull * 1000ull * 1000ull;

struct Comm {
    int rank = 0, nranks = 1;
    double* peer[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long* seq = nullptr;   // local device counters: [0] reductions, [1] barriers
    int* err = nullptr;                  // local device flag: a bounded wait expired
};
constexpr size_t kCommHeaderBytes = 256;   // the partials buffer of grid_reduce is preceded by a Comm (see comm_of)
static_assert(sizeof(Comm) <= kCommHeaderBytes, "Comm must fit the header of the partials buffer");

__device__ __forceinline__ const Comm* comm_of(const double* part) {
    return reinterpret_cast<const Comm*>(reinterpret_cast<const char*>(part) - kCommHeaderBytes);
}
__device__ __forceinline__ unsigned long long global_ns() {
#ifdef VCH_CPU_EMU   // tests/emu: kernels compiled for the host by g++ (no PTX)
    return 0ull;
#else
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
#endif
}
// Spins until *flag >= want (monotone flags) or the limit expires.  Returns false on timeout.
__device__ __forceinline__ bool spin_until(volatile unsigned long long* flag, unsigned long long want, int* err, bool exact) {
    if (err && *reinterpret_cast<volatile int*>(err)) return false;   // an earlier wait already failed: do not wait again
    const unsigned long long t0 = global_ns();
    unsigned int polls = 0;
    while (true) {
        const unsigned long long v = *flag;
        if (exact ? (v == want) : (v >= want)) return true;
        if ((++polls & 1023u) == 0u && global_ns() - t0 > kSpinLimitNs) { if (err) *err = 1; return false; }
    }
}

// Partials buffer with its Comm header (device memory).  `part` is what kernels receive.
struct RedBuf {
    char* base = nullptr;
    double* part = nullptr;
    void alloc(size_t doubles, const Comm& cm) {
        release();
        VCH_CUDA(cudaMalloc(&base, kCommHeaderBytes + doubles * sizeof(double)));
        part = reinterpret_cast<double*>(base + kCommHeaderBytes);
        set_comm(cm);
    }
    void set_comm(const Comm& cm) { VCH_CUDA(cudaMemcpy(base, &cm, sizeof(Comm), cudaMemcpyHostToDevice)); }
    void release() { if (base) cudaFree(base); base = nullptr; part = nullptr; }
    ~RedBuf() { release(); }
    RedBuf() = default;
    RedBuf(const RedBuf&) = delete;
    RedBuf& operator=(const RedBuf&) = delete;
};

// One cudaMalloc per rank, carved identically on every rank (sizes must not depend on the rank).
struct Arena {
    double* base = nullptr;
    size_t cap = 0, used = 0;    // doubles
    void create(size_t doubles) {
        VCH_CUDA(cudaMalloc(&base, doubles * sizeof(double)));
        VCH_CUDA(cudaMemset(base, 0, doubles * sizeof(double)));
        cap = doubles; used = kArenaHeader;
    }
    // `margin` doubles of ghost storage on both sides of the returned pointer
    double* carve(size_t count, size_t margin) {
        size_t start = (used + margin + 31) & ~size_t(31);          // 256-byte aligned payload
        if (start + count + margin > cap) throw Error(VCH_E_ARG, "slab arena exhausted");
        used = start + count + margin;
        return base + start;
    }
    void view(DevBuf& b, size_t count, size_t margin) { b.release(); b.p = carve(count, margin); b.n = count; b.view = true; }
    void destroy() { if (base) cudaFree(base); base = nullptr; cap = used = 0; }
};

// ---------------------------------------------------------------- programmatic dependent launch
// The hot path is a chain of 10-25 us kernels, each depending on the one before.  Launched with programmatic stream
// serialization, kernel N+1 is scheduled while kernel N is still running (its CTAs become resident as soon as every CTA
// of N has passed pdl_enter) and parks in griddepcontrol.wait until N has completed and flushed its memory — the launch
// latency and CTA ramp-up move under the predecessor instead of sitting between the two.  EVERY kernel of this library
// begins with pdl_enter(): the wait makes the predecessor's writes (and, transitively, everything before it) visible, so
// nothing is read or written ahead of time; in a kernel launched without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
#ifndef VCH_CPU_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

#ifndef VCH_CPU_EMU
template <typename... KArgs, typename... Args>
inline void launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    if (pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    }
    VCH_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}
#else   // tests/emu: the same funnel runs the kernel on the CPU emulator (vch_emu::launch_kernel, cuda_emu.h)
template <typename... KArgs, typename... Args>
inline void launch_pdl(bool, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args&&... args) {
    vch_emu::launch_kernel(kern, grid.x, block.x, smem, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------- launch geometry
constexpr int kSMs = 148;              // B200
constexpr int kRedThreads = 256;
constexpr int kRedBlocksMax = kSMs * 4;  // grid-stride reductions: 4 resident CTAs of 256 threads per SM (8/SM measured slower)

static inline int red_blocks(long long n) {
    long long b = (n + kRedThreads - 1) / kRedThreads;
    if (b < 1) b = 1;
    return (int)(b > kRedBlocksMax ? kRedBlocksMax : b);
}

// ---------------------------------------------------------------- deterministic grid reductions
// Each block reduces K values with warp shuffles + one smem stage, writes its partials to
// part[k*gridDim.x + blockIdx.x]; the last block to arrive (ticket) re-reduces all partials in a
// fixed order so the result does not depend on block scheduling.  OP: 0 = sum, 1 = min, 2 = max.
template <int OP> __device__ __forceinline__ double red_op(double a, double b) {
    if (OP == 0) return a + b;
    if (OP == 1) return fmin(a, b);
    return fmax(a, b);
}
template <int OP> __device__ __forceinline__ double red_identity() {
    if (OP == 0) return 0.0;
    if (OP == 1) return INFINITY;
    return -INFINITY;
}
template <int OP> __device__ __forceinline__ double warp_red(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int OP> __device__ __forceinline__ double block_red(double v, double* sh /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_red<OP>(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : red_identity<OP>();
    if (wid == 0) v = warp_red<OP>(v);
    return v;   // valid in warp 0 (all lanes)
}

// Cross-rank stage of a grid reduction (thread 0 of the last block).  Slots are double-buffered by sequence parity: a
// rank cannot be two reductions ahead of a peer, because finishing reduction s+1 needs that peer's s+1 contribution.
template <int K>
__device__ __forceinline__ void xrank_reduce(double (&tot)[K], const int (&op)[K], const Comm& cm) {
    static_assert(K <= kRedK, "too many values for a cross-rank reduction");
    const unsigned long long s = ++cm.seq[0];
    const size_t mine = kSlotOff + ((size_t)(s & 1ull) * kMaxRanks + cm.rank) * (kRedK + 1);
    for (int r = 0; r < cm.nranks; ++r) {
        volatile double* slot = cm.peer[r] + mine;
#pragma unroll
        for (int k = 0; k < K; ++k) slot[k] = tot[k];
    }
    __threadfence_system();
    for (int r = 0; r < cm.nranks; ++r)
        *reinterpret_cast<volatile unsigned long long*>(cm.peer[r] + mine + kRedK) = s;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = (op[k] == 0) ? 0.0 : (op[k] == 1 ? INFINITY : -INFINITY);
    for (int r = 0; r < cm.nranks; ++r) {
        volatile double* slot = cm.peer[cm.rank] + kSlotOff + ((size_t)(s & 1ull) * kMaxRanks + r) * (kRedK + 1);
        if (!spin_until(reinterpret_cast<volatile unsigned long long*>(slot + kRedK), s, cm.err, true)) break;
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double x = slot[k];
            acc[k] = (op[k] == 0) ? acc[k] + x : (op[k] == 1 ? fmin(acc[k], x) : fmax(acc[k], x));
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = acc[k];
}

// All K values through the block at once: one shared-memory stage and two barriers for the whole set (round 1 ran K block
// reductions back to back — 2K barriers in every block and, in the last block, K dependent rounds of partial loads; measured
// in the fused row epilogues: ~9 us on top of the plain transform).  The combination order per value is unchanged (strided
// per-thread accumulation, xor-shuffle tree, warp partials in warp order), so results are bit-identical to round 1.
template <int K>
__device__ __forceinline__ void block_red_multi(double (&v)[K], const int (&op)[K], double (*sh)[32] /* [K][32] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = (op[k] == 0) ? warp_red<0>(v[k]) : (op[k] == 1 ? warp_red<1>(v[k]) : warp_red<2>(v[k]));
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sh[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = (threadIdx.x < nw) ? sh[k][threadIdx.x] : ((op[k] == 0) ? 0.0 : (op[k] == 1 ? INFINITY : -INFINITY));
            v[k] = (op[k] == 0) ? warp_red<0>(x) : (op[k] == 1 ? warp_red<1>(x) : warp_red<2>(x));
        }
    }
}

// Returns true (block-uniform) in the last block; there tot[k] holds the grid-wide result (valid in thread 0).
template <int K>
__device__ __forceinline__ bool grid_reduce(double (&v)[K], const int (&op)[K], double* part, unsigned int* ticket,
                                            double (&tot)[K]) {
    __shared__ double sh[K][32];
    __shared__ bool is_last;
    // slab mode flag, fetched up front so that its latency hides under the block reduction instead of trailing the kernel
    int nranks = 1;
    if (threadIdx.x == 0) nranks = comm_of(part)->nranks;
    block_red_multi<K>(v, op, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[(size_t)k * gridDim.x + blockIdx.x] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = (op[k] == 0) ? 0.0 : (op[k] == 1 ? INFINITY : -INFINITY);
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        double x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = __ldcg(&part[(size_t)k * gridDim.x + b]);     // K independent loads in flight
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] = (op[k] == 0) ? a[k] + x[k] : (op[k] == 1 ? fmin(a[k], x[k]) : fmax(a[k], x[k]));
    }
    block_red_multi<K>(a, op, sh);
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = a[k];
    if (threadIdx.x == 0) {
        *ticket = 0u;   // re-arm for the next launch on this stream
        if (nranks > 1) xrank_reduce<K>(tot, op, *comm_of(part));
    }
    return true;
}

// Mirror (even) reflection of an index into [0, n-1] for ghost offsets up to 2 (Neumann: v[-1] = v[1]).
__device__ __forceinline__ int mirror(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}


// One-CTA barrier across ranks: every rank announces its arrival in all peers' flag arrays and waits for everyone.
// Stream order makes everything this rank enqueued before the barrier (including its peer writes) complete first.
static __global__ void xbar_kernel(Comm cm, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    __shared__ unsigned long long s;
    if (threadIdx.x == 0) s = ++cm.seq[1];
    __syncthreads();
    const int r = threadIdx.x;
    if (r < cm.nranks) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(cm.peer[r] + kFlagOff + cm.rank) = s;
        spin_until(reinterpret_cast<volatile unsigned long long*>(cm.peer[cm.rank] + kFlagOff + r), s, cm.err, false);
    }
    __syncthreads();
    __threadfence_system();
}

// Barrier body for one CTA (thread r talks to rank r).  Must be called by every thread of the block.
__device__ __forceinline__ void xbar_block(const Comm& cm, unsigned long long* s_shared) {
    __syncthreads();
    if (threadIdx.x == 0) *s_shared = ++cm.seq[1];
    __syncthreads();
    const int r = threadIdx.x;
    if (r < cm.nranks) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(cm.peer[r] + kFlagOff + cm.rank) = *s_shared;
        spin_until(reinterpret_cast<volatile unsigned long long*>(cm.peer[cm.rank] + kFlagOff + r), *s_shared, cm.err, false);
    }
    __syncthreads();
    __threadfence_system();
}

// ONE CTA: [optional barrier] -> write the first / last `rows` owned rows of up to two fields into the neighbours' ghost
// rows -> barrier (the rows have landed).  Fields live at the same arena offset on every rank; ranks below the last own
// `rows_lo` rows each.
// The leading barrier (lead = 1) protects ghost rows that a neighbour might still be reading.  The solver does not need it:
// every kernel that reads ghost rows is followed — before the next push into the same buffer — by a cross-rank reduction
// (grid_reduce) on every rank, and a rank can only pass reduction R after every peer has reached it, i.e. after the peers'
// earlier kernels (stream order) have finished.  Callers outside that pattern pass lead = 1.
static __global__ void __launch_bounds__(1024) halo_push_kernel(Comm cm, const double* f0, const double* f1, int rows, int nloc,
                                                                int rows_lo, int ni, const int* __restrict__ done, int lead) {
    pdl_enter();
    if (done && *done) return;
    __shared__ unsigned long long s;
    if (lead) xbar_block(cm, &s);
    const int cnt = rows * ni;
    for (int w = 0; w < 2; ++w) {
        const double* f = w ? f1 : f0;
        if (!f) continue;
        const size_t off = (size_t)(f - cm.peer[cm.rank]);
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            if (cm.rank > 0)               // my first rows -> upper ghost rows of the rank below
                cm.peer[cm.rank - 1][off + (size_t)rows_lo * ni + e] = f[e];
            if (cm.rank < cm.nranks - 1)   // my last rows -> lower ghost rows [-rows, 0) of the rank above
                *(cm.peer[cm.rank + 1] + (long long)off - cnt + e) = f[(long long)(nloc - rows) * ni + e];
        }
    }
    xbar_block(cm, &s);
}

}  // namespace vch
