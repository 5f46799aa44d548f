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
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        VCH_CUDA(cudaMalloc(&p, count * sizeof(double)));
        n = count;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// Stages caller arrays when mem == VCH_MEM_HOST; passes device pointers through otherwise.
struct Stager {
    cudaStream_t s;
    int mem;
    std::vector<double*> owned;
    struct Out { double* host; double* dev; size_t n; };
    std::vector<Out> outs;
    Stager(cudaStream_t st, int m) : s(st), mem(m) {}
    const double* in(const double* a, size_t n) {
        if (!a || mem == VCH_MEM_DEVICE) return a;
        double* d; VCH_CUDA(cudaMalloc(&d, n * sizeof(double))); owned.push_back(d);
        VCH_CUDA(cudaMemcpyAsync(d, a, n * sizeof(double), cudaMemcpyHostToDevice, s));
        return d;
    }
    double* out(double* a, size_t n) {
        if (!a || mem == VCH_MEM_DEVICE) return a;
        double* d; VCH_CUDA(cudaMalloc(&d, n * sizeof(double))); owned.push_back(d);
        outs.push_back({a, d, n});
        return d;
    }
    void finish() {
        for (auto& o : outs) VCH_CUDA(cudaMemcpyAsync(o.host, o.dev, o.n * sizeof(double), cudaMemcpyDeviceToHost, s));
        VCH_CUDA(cudaStreamSynchronize(s));
        outs.clear();
    }
    ~Stager() { for (auto d : owned) cudaFree(d); }
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
    double mass, wint, mass0;
    double tol2;
    double c0, c2;            // operator coefficients of the current solve (read by the graph's kernels: one graph serves every dt)
    int done, iters, nonfinite, maxit;
    long long iters_total, solves, stalls;   // accumulated on the device (graph-driven solves are never polled)
    long long g_launches;                    // kernels that ran inside solve graphs (not seen by the host launch log)
    int iters_max, pad;
};

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

// Returns true (block-uniform) in the last block; there tot[k] holds the grid-wide result (valid in thread 0).
template <int K>
__device__ __forceinline__ bool grid_reduce(double (&v)[K], const int (&op)[K], double* part, unsigned int* ticket,
                                            double (&tot)[K]) {
    __shared__ double sh[32];
    __shared__ bool is_last;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double r = (op[k] == 0) ? block_red<0>(v[k], sh) : (op[k] == 1) ? block_red<1>(v[k], sh) : block_red<2>(v[k], sh);
        if (threadIdx.x == 0) part[(size_t)k * gridDim.x + blockIdx.x] = r;
    }
    __threadfence();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = (op[k] == 0) ? 0.0 : (op[k] == 1 ? INFINITY : -INFINITY);
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            double x = __ldcg(&part[(size_t)k * gridDim.x + b]);
            a = (op[k] == 0) ? a + x : (op[k] == 1 ? fmin(a, x) : fmax(a, x));
        }
        tot[k] = (op[k] == 0) ? block_red<0>(a, sh) : (op[k] == 1) ? block_red<1>(a, sh) : block_red<2>(a, sh);
    }
    if (threadIdx.x == 0) *ticket = 0u;   // re-arm for the next launch on this stream
    return true;
}

// Mirror (even) reflection of an index into [0, n-1] for ghost offsets up to 2 (Neumann: v[-1] = v[1]).
__device__ __forceinline__ int mirror(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

}  // namespace vch
