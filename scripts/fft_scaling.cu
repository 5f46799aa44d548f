// Scratch probe (not part of the product): is the row / column transform latency- or throughput-bound?  Times the kernels
// in a CUDA graph for 1, 2, 3, 4 CTAs per SM (number of lines varied), N = 1024.
#define VCH_FFT16_TIMING 1
#include "../sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200/csrc/vch_dct.cuh"
#include <functional>
namespace vch { static thread_local std::string g_err; void set_last_error(const std::string& m) { g_err = m; } }
using namespace vch;
int main() {
    const int N = 1024, n1 = N + 1; const size_t n = (size_t)n1 * n1;
    LaunchLog log; DctPlan plan; plan.init(n1, n1, 1.0 / N, 1.0 / N, &log);
    double *a, *b; cudaMalloc(&a, n * 8 + 4096); cudaMalloc(&b, (size_t)n1 * plan.pitch * 8 + 4096);
    std::vector<double> h(n); for (size_t i = 0; i < n; ++i) h[i] = sin(0.001 * i) + 0.3 * cos(0.37 * i);
    cudaMemcpy(a, h.data(), n * 8, cudaMemcpyHostToDevice); cudaMemset(b, 0, (size_t)n1 * plan.pitch * 8);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int K = 20, reps = 20;
    auto bench = [&](const char* name, int lines, std::function<void()> launch) {
        cudaGraph_t gr; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < K; ++k) launch();
        cudaStreamEndCapture(s, &gr); cudaGraphInstantiate(&ge, gr, 0);
        for (int w = 0; w < 2; ++w) cudaGraphLaunch(ge, s);
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-18s lines %5d : %7.2f us  %s\n", name, lines, 1e3 * ms / (reps * K), cudaGetErrorString(cudaGetLastError()));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(gr);
    };
    const SymbolArgs nosym{1.0, 0.0, nullptr, 0.0, nullptr};
    const SymbolArgs sy{100.0, 5e-5, nullptr, 7.0, nullptr};
    const double norm = 1.0 / (4.0 * N * N);
    const int P = plan.pitch;
    for (int lines : {2, 74, 148, 296, 592, 888, 1025}) {
        const int rppb = 1, rthreads = 256, rgrid = (lines + 1) / 2;
        bench("rows plain", lines, [&] {
            dct_fft_kernel<11, false, 512><<<rgrid, rthreads, dct_smem(plan.inner, rppb), s>>>(a, b, lines, n1, n1, 1, P, 1, rppb, plan.inner.tw, nullptr, nullptr,
                nosym, 1.0, 0, RowPrologue(), DotEpilogue(), nullptr, Scatter());
        });
    }
    for (int cols : {4, 296, 592, 888, 1025}) {
        const int cppb = 2, cthreads = 512, cgrid = ((cols + 1) / 2 + cppb - 1) / cppb;
        bench("cols solve", cols, [&] {
            dct_fft_kernel<11, true, 512><<<cgrid, cthreads, dct_smem(plan.outer, cppb), s>>>(b, b, cols, n1, 1, P, 1, P, cppb, plan.outer.tw, plan.inner.lam, plan.outer.lam,
                sy, norm, 0, RowPrologue(), DotEpilogue(), nullptr, Scatter());
        });
    }
    for (int cols : {296, 592, 1025}) {   // column solve with 2 columns per CTA (256 threads)
        const int cppb = 1, cthreads = 256, cgrid = ((cols + 1) / 2 + cppb - 1) / cppb;
        bench("cols solve ppb1", cols, [&] {
            dct_fft_kernel<11, true, 512><<<cgrid, cthreads, dct_smem(plan.outer, cppb), s>>>(b, b, cols, n1, 1, P, 1, P, cppb, plan.outer.tw, plan.inner.lam, plan.outer.lam,
                sy, norm, 0, RowPrologue(), DotEpilogue(), nullptr, Scatter());
        });
    }
    bench("full apply (old)", 1025, [&] { plan.lean = false; plan.apply(s, a, a, sy, nullptr); });
    // ---- radix-16 kernels (vch_fft16.cuh)
    using G = F16<11>;
    for (int lines : {2, 148, 296, 592, 888, 1025}) {
        const int grid = ((lines + 1) / 2 + G::fpb - 1) / G::fpb;
        bench("rows16 plain", lines, [&] {
            rows16_kernel<11, 0, 0, false><<<grid, G::rthreads, G::rows_smem_plain, s>>>(a, b, lines, n1, P, plan.inner.tw16, RowPrologue(), DotEpilogue(), nullptr, Scatter());
        });
    }
    for (int cols : {8, 1025}) {
        const int grid = (cols + 2 * G::cp - 1) / (2 * G::cp);
        bench("cols16 solve", cols, [&] {
            cols16_kernel<11><<<grid, G::cthreads, G::cols_smem_bytes, s>>>(b, P, cols, plan.outer.tw16, plan.inner.lam, plan.outer.lam, sy, norm, 0, nullptr);
        });
    }
    {
        CUtensorMap tm;
        const bool ok = cols16_tensor_map(&tm, 11, b, P, n1);
        cudaFuncSetAttribute(cols16_tma_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F16T<11>::smem_bytes);
        for (int cols : {8, 1025}) {
            const int grid = (cols + 2 * G::cp - 1) / (2 * G::cp);
            if (ok) bench("cols16 tma solve", cols, [&] {
                cols16_tma_kernel<11><<<grid, G::cthreads, F16T<11>::smem_bytes, s>>>(tm, cols, plan.outer.tw16, plan.inner.lam, plan.outer.lam, sy, norm, 0, nullptr);
            });
        }
    }
    bench("full apply (fft16)", 1025, [&] { plan.lean = true; plan.apply(s, a, a, sy, nullptr); });
    {   // fused modes, in graph
        double* w[8]; for (auto& p : w) { cudaMalloc(&p, n * 8); cudaMemcpy(p, h.data(), n * 8, cudaMemcpyHostToDevice); }
        RedBuf red; red.alloc(8 * 4096, Comm());
        unsigned int* ticket; cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
        Scal* sc; cudaMalloc(&sc, sizeof(Scal)); Scal hs{}; hs.alpha = 0.7; hs.omega = 0.9; hs.rho = 1.0; hs.rho_new = 0.8; hs.abar = 7.0; hs.iters = 1; hs.thr2 = 0.0; hs.rr = 1.0; hs.maxit = 1 << 30;
        cudaMemcpy(sc, &hs, sizeof(Scal), cudaMemcpyHostToDevice);
        const int grid = ((n1 + 1) / 2 + G::fpb - 1) / G::fpb;
        RowPrologue p2{2, w[0], w[1], w[2], w[3], sc};
        RowPrologue p3{3, w[0], w[1], w[2], w[3], sc}; p3.s = w[4]; p3.t = w[5]; p3.x = w[6]; p3.rw = w[0];
        bench("rows16 pro2 mul", 1025, [&] { rows16_kernel<11, 2, 0, true><<<grid, G::rthreads, G::rows_smem_staged, s>>>(a, b, n1, n1, P, plan.inner.tw16, p2, DotEpilogue(), nullptr, Scatter()); });
        bench("rows16 pro3 mul", 1025, [&] { rows16_kernel<11, 3, 0, true><<<grid, G::rthreads, G::rows_smem_staged, s>>>(a, b, n1, n1, P, plan.inner.tw16, p3, DotEpilogue(), nullptr, Scatter()); });
        DotEpilogue e1{1, w[1], sc, red.part, ticket, w[2], nullptr, w[3]};
        DotEpilogue e4{4, w[1], sc, red.part, ticket, w[1], nullptr, w[3]};
        bench("rows16 epi1", 1025, [&] { rows16_kernel<11, 0, 1, false><<<grid, G::rthreads, G::rows_smem_plain, s>>>(b, w[7], n1, P, n1, plan.inner.tw16, RowPrologue(), e1, nullptr, Scatter()); });
        bench("rows16 epi4", 1025, [&] { rows16_kernel<11, 0, 4, false><<<grid, G::rthreads, G::rows_smem_plain, s>>>(b, w[7], n1, P, n1, plan.inner.tw16, RowPrologue(), e4, nullptr, Scatter()); });
        DotEpilogue e1n = e1; e1n.rvec = nullptr;
        bench("rows16 epi1 no r", 1025, [&] { rows16_kernel<11, 0, 1, false><<<grid, G::rthreads, G::rows_smem_plain, s>>>(b, w[7], n1, P, n1, plan.inner.tw16, RowPrologue(), e1n, nullptr, Scatter()); });
    }
    {   // phase stamps of the column solve (block 0, thread 0), one launch on an idle GPU and one with the full grid
        for (int cols : {8, 1025}) {
            const int grid = (cols + 2 * F16<11>::cp - 1) / (2 * F16<11>::cp);
            for (int w = 0; w < 10; ++w)
                cols16_kernel<11><<<grid, F16<11>::cthreads, F16<11>::cols_smem_bytes, s>>>(b, P, cols, plan.outer.tw16, plan.inner.lam, plan.outer.lam, sy, norm, 0, nullptr);
            cudaStreamSynchronize(s);
            long long st[32]; cudaMemcpyFromSymbol(st, vch_dbg_clock, sizeof(st));
            printf("cols16 phases (%d cols), cycles: stage-in %lld | load+first %lld | middle %lld | last %lld | factor+sync %lld | inverse fft %lld | stage-out write %lld | store %lld | total %lld\n",
                   cols, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[7] - st[6], st[8] - st[7], st[8] - st[0]);
        }
    }
    printf("done %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
