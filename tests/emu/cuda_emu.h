// Minimal CPU emulation of the CUDA execution model — TEST INFRASTRUCTURE ONLY (tests/emu).
//
// Purpose: run the library's kernels (csrc/*.cuh, compiled by g++ with -DVCH_CPU_EMU) on the host so that kernel-level
// logic — index maps, fused prologues/epilogues, flag protocols, reductions — can be checked without a GPU, and so that
// experimental kernel variants can be diffed bit for bit against the variant that has been verified on a B200.
// It models what those kernels use and nothing more:
//   * one block at a time, its threads as cooperatively scheduled fibers (ucontext), switched at __syncthreads and at
//     warp shuffles (so barrier placement errors deadlock or misread exactly as they would diverge on the device);
//   * threadIdx / blockIdx / blockDim / gridDim, static and dynamic shared memory (blocks run one after the other, so
//     `__shared__` variables are function statics), __shfl_xor_sync within 32-lane warps, atomicAdd, fences (no-ops),
//     __ldcg, the round-to-nearest fp64 intrinsics.
// It does NOT model memory ordering, races between blocks, or timing.
#pragma once
#include <cuda_runtime.h>
#include <ucontext.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#undef __shared__
#define __shared__ static
#ifndef __grid_constant__
#define __grid_constant__
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __maxnreg__
#define __maxnreg__(...)
#endif

namespace vch_emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
inline Dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

struct Fiber {
    ucontext_t ctx;
    bool done = false;
    unsigned tid = 0;
};
constexpr size_t kStackBytes = 96 * 1024;
// fiber stacks are reused across launches (thread t of every block runs on stack t; allocated once, never cleared)
inline char* fiber_stack(unsigned t) {
    static std::vector<char*> pool;
    while (pool.size() <= t) pool.push_back(static_cast<char*>(std::malloc(kStackBytes)));
    return pool[t];
}

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = -1;
    // block barrier
    unsigned bar_arrived = 0, bar_gen = 0;
    // warp shuffle exchange: per warp 32 slots, arrival counter and generation
    std::vector<double> shfl_val;
    std::vector<unsigned> shfl_arrived, shfl_gen;
    std::function<void()> body;
    std::vector<char> dyn_smem;
    unsigned live = 0;             // threads that have not returned from the kernel yet
};
inline Block* g_block = nullptr;

inline void* dynamic_smem() { return g_block->dyn_smem.data(); }

inline void yield() {
    Block* b = g_block;
    Fiber& f = b->fibers[b->current];
    swapcontext(&f.ctx, &b->sched);
}

inline void fiber_entry() {
    Block* b = g_block;
    b->body();
    b->fibers[b->current].done = true;
    b->live -= 1;
    yield();
}

inline unsigned live_threads(Block* b) { return b->live; }

// All threads of the block must arrive (threads that have already returned from the kernel are not waited for, which is
// what the hardware does as well).
inline void syncthreads() {
    Block* b = g_block;
    const unsigned gen = b->bar_gen;
    b->bar_arrived += 1;
    while (b->bar_gen == gen) {
        if (b->bar_arrived >= live_threads(b)) { b->bar_arrived = 0; b->bar_gen += 1; break; }
        yield();
    }
}

inline double shfl_xor(double v, int lane_mask) {
    Block* b = g_block;
    const unsigned tid = b->fibers[b->current].tid, warp = tid >> 5, lane = tid & 31u;
    const unsigned nthreads = (unsigned)b->fibers.size();
    const unsigned wsize = (warp * 32u + 32u <= nthreads) ? 32u : (nthreads - warp * 32u);
    // phase 1: publish, wait for the whole warp
    b->shfl_val[warp * 32u + lane] = v;
    unsigned gen = b->shfl_gen[warp];
    b->shfl_arrived[warp] += 1;
    while (b->shfl_gen[warp] == gen) {
        if (b->shfl_arrived[warp] >= wsize) { b->shfl_arrived[warp] = 0; b->shfl_gen[warp] += 1; break; }
        yield();
    }
    const unsigned src = lane ^ (unsigned)lane_mask;
    // A lane that was never launched (block size not a multiple of 32) reads as 0 — what the B200 returns in the GPU suite
    // for the 136-thread row transform of the 32^2 grid (formally undefined; the library only sums across such warps).
    const double r = (src < wsize) ? b->shfl_val[warp * 32u + src] : 0.0;
    // phase 2: everybody has read before anybody publishes again
    gen = b->shfl_gen[warp];
    b->shfl_arrived[warp] += 1;
    while (b->shfl_gen[warp] == gen) {
        if (b->shfl_arrived[warp] >= wsize) { b->shfl_arrived[warp] = 0; b->shfl_gen[warp] += 1; break; }
        yield();
    }
    return r;
}

// Runs `body` for every thread of every block of the grid (blocks in order, x-dimension only — all the library uses).
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
    g_gridDim = Dim3{grid, 1, 1};
    g_blockDim = Dim3{block, 1, 1};
    for (unsigned bid = 0; bid < grid; ++bid) {
        Block blk;
        blk.body = body;
        blk.fibers.resize(block);
        blk.live = block;
        blk.dyn_smem.assign(smem_bytes + 64, 0);
        const unsigned nwarps = (block + 31u) / 32u;
        blk.shfl_val.assign(nwarps * 32u, 0.0);
        blk.shfl_arrived.assign(nwarps, 0u);
        blk.shfl_gen.assign(nwarps, 0u);
        g_block = &blk;
        g_blockIdx = Dim3{bid, 0, 0};
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = blk.fibers[t];
            f.tid = t;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = fiber_stack(t);
            f.ctx.uc_stack.ss_size = kStackBytes;
            f.ctx.uc_link = &blk.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        unsigned long long spins = 0;
        while (live_threads(&blk) > 0) {
            bool progressed = false;
            for (unsigned t = 0; t < block; ++t) {
                Fiber& f = blk.fibers[t];
                if (f.done) continue;
                blk.current = (int)t;
                g_threadIdx = Dim3{t, 0, 0};
                swapcontext(&blk.sched, &f.ctx);
                progressed = true;
            }
            if (!progressed) break;
            if (++spins > 100000000ull) { fprintf(stderr, "cuda_emu: block %u does not terminate (barrier mismatch?)\n", bid); abort(); }
        }
        g_block = nullptr;
    }
}

// Kernel launch by function pointer (the library's launch_pdl funnel): arguments are converted to the kernel's parameter types
// and copied, as a launch does; trailing parameters the call site leaves to their C++ defaults are value-initialised (every
// such default in the library is 0).
template <typename Tuple, typename... KArgs, size_t... I>
inline void call_with_tuple(void (*kern)(KArgs...), Tuple& t, std::index_sequence<I...>) { kern(std::get<I>(t)...); }
template <size_t I, typename Tuple> inline void fill_from(Tuple&) {}
template <size_t I, typename Tuple, typename A, typename... Rest>
inline void fill_from(Tuple& t, A&& a, Rest&&... rest) {
    std::get<I>(t) = static_cast<std::tuple_element_t<I, Tuple>>(a);
    fill_from<I + 1>(t, std::forward<Rest>(rest)...);
}
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, Args&&... args) {
    static_assert(sizeof...(Args) <= sizeof...(KArgs), "too many kernel arguments");
    std::tuple<std::remove_cv_t<std::remove_reference_t<KArgs>>...> t{};
    fill_from<0>(t, std::forward<Args>(args)...);
    launch(grid, block, smem, [&] { call_with_tuple(kern, t, std::index_sequence_for<KArgs...>{}); });
}

template <typename F> inline cudaError_t func_set_attribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

}  // namespace vch_emu

// ---- the names kernels use
#define threadIdx vch_emu::g_threadIdx
#define blockIdx vch_emu::g_blockIdx
#define blockDim vch_emu::g_blockDim
#define gridDim vch_emu::g_gridDim

using std::isfinite;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline void __syncthreads() { vch_emu::syncthreads(); }
inline double __shfl_xor_sync(unsigned, double v, int m) { return vch_emu::shfl_xor(v, m); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p = o + v; return o; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __drcp_rn(double a) { volatile double r = 1.0 / a; return r; }
typedef unsigned long long cudaGraphConditionalHandle_emu;
inline void cudaGraphSetConditional(cudaGraphConditionalHandle, unsigned int) {}
