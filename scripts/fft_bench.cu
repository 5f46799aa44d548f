// Scratch micro-benchmark for the DCT kernels (not part of the product): times rows / cols-solve kernels in a loop on
// a warm L2, for variants selected with -D flags.   nvcc -O3 -arch=sm_100a scripts/fft_bench.cu -o /tmp/fft_bench
#ifdef WITH_CUFFT
#include <cufft.h>
#endif
#include "../sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200/csrc/vch_dct.cuh"
namespace vch { static thread_local std::string g_err; void set_last_error(const std::string& m) { g_err = m; } }
using namespace vch;
int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1024, reps = 200;
    const int n1 = N + 1; const size_t n = (size_t)n1 * n1;
    LaunchLog log; DctPlan plan; plan.init(n1, n1, 1.0 / N, 1.0 / N, &log);
    double *a, *b; cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
    std::vector<double> h(n); for (size_t i = 0; i < n; ++i) h[i] = sin(0.001 * i) + 0.3 * cos(0.37 * i);
    cudaMemcpy(a, h.data(), n * 8, cudaMemcpyHostToDevice);
    SymbolArgs sy{100.0, 5e-5, nullptr, 7.0, nullptr};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 5; ++w) plan.apply(0, a, b, sy, nullptr);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) plan.apply(0, a, b, sy, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("N=%d full precond apply: %.2f us\n", N, 1e3 * ms / reps);
    log.profiling = true;
    for (int r = 0; r < 50; ++r) plan.apply(0, a, b, sy, nullptr);
    cudaDeviceSynchronize();
    for (auto& row : log.report()) printf("  %-22s %8.2f us avg over %lld\n", row.name.c_str(), 1e3 * row.ms / row.n, row.n);
#ifdef WITH_CUFFT
    {   // library yardstick: the same amount of transform work done by cuFFT (Z2Z, length 2N, (N+1)/2 lines per batch),
        // contiguous lines (our row kernel) and strided lines (our column kernel, forward only)
        const int Lf = 2 * N, batch = (n1 + 1) / 2;
        cufftDoubleComplex *ci, *co; cudaMalloc(&ci, sizeof(cufftDoubleComplex) * (size_t)Lf * batch); cudaMalloc(&co, sizeof(cufftDoubleComplex) * (size_t)Lf * batch);
        cudaMemset(ci, 0, sizeof(cufftDoubleComplex) * (size_t)Lf * batch);
        cufftHandle pr, pc; int len[1] = {Lf};
        cufftPlanMany(&pr, 1, len, len, 1, Lf, len, 1, Lf, CUFFT_Z2Z, batch);
        cufftPlanMany(&pc, 1, len, len, batch, 1, len, batch, 1, CUFFT_Z2Z, batch);
        for (int which = 0; which < 2; ++which) {
            cufftHandle pl = which ? pc : pr;
            for (int w = 0; w < 5; ++w) cufftExecZ2Z(pl, ci, co, CUFFT_FORWARD);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; ++r) cufftExecZ2Z(pl, ci, co, CUFFT_FORWARD);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("cuFFT Z2Z len %d x %d lines, %s: %.2f us per batch\n", Lf, batch, which ? "strided" : "contiguous", 1e3 * ms / reps);
        }
        // real-to-complex of the even extension is what a library user would actually call for a DCT-I
        cufftHandle pd; cufftPlanMany(&pd, 1, len, len, 1, Lf, len, 1, Lf / 2 + 1, CUFFT_D2Z, n1);
        double* ri; cudaMalloc(&ri, sizeof(double) * (size_t)Lf * n1); cudaMemset(ri, 0, sizeof(double) * (size_t)Lf * n1);
        cufftDoubleComplex* ro; cudaMalloc(&ro, sizeof(cufftDoubleComplex) * (size_t)(Lf / 2 + 1) * n1);
        for (int w = 0; w < 5; ++w) cufftExecD2Z(pd, ri, ro);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cufftExecD2Z(pd, ri, ro);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cuFFT D2Z len %d x %d lines (even extension materialised): %.2f us per batch\n", Lf, n1, 1e3 * ms / reps);
    }
#endif
    // checksum
    cudaMemcpy(h.data(), b, n * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (size_t i = 0; i < n; i += 97) s += h[i] * (1 + (i % 7));
    printf("checksum %.12e  err=%s\n", s, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
