// Scratch probe (not part of the product): is the row / column transform latency- or throughput-bound?  Times the kernels
// in a CUDA graph for 1, 2, 3, 4 CTAs per SM (number of lines varied), N = 1024.
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
    bench("full apply", 1025, [&] { plan.apply(s, a, a, sy, nullptr); });
    printf("done %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
