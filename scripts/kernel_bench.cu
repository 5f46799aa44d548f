// Scratch micro-benchmark (not part of the product): in-graph time of the elementwise / stencil / reduction kernels of
// csrc/vch2d_kernels.cuh at 1024^2, L2-warm, K launches per graph.  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include "../sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200/csrc/vch2d_tiles.cuh"
#include <functional>
namespace vch { static thread_local std::string g_err; void set_last_error(const std::string& m) { g_err = m; } }
using namespace vch;

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1024;
    const int n1 = N + 1; const long long n = (long long)n1 * n1;
    Geo g{}; g.no = n1; g.ni = n1; g.nx1 = n1; g.ny1 = n1; g.n = n; g.iho2 = (double)N * N; g.ihi2 = (double)N * N; g.nxg = n1;
    Phys ph{0.05, 10.0, 0.75, 1.0, 1e-4, 0.99, 5e-3, 1.0 - 1e-4};
    const int NB = 16;
    double* buf[NB];
    std::vector<double> h(n);
    for (int b = 0; b < NB; ++b) {
        cudaMalloc(&buf[b], n * 8);
        for (long long i = 0; i < n; ++i) h[i] = 0.3 * sin(0.001 * i + b) + 0.2 * cos(0.37 * i);
        cudaMemcpy(buf[b], h.data(), n * 8, cudaMemcpyHostToDevice);
    }
    RedBuf red; red.alloc(8 * 4096, Comm());
    unsigned int* ticket; cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
    Scal* sc; cudaMalloc(&sc, sizeof(Scal)); cudaMemset(sc, 0, sizeof(Scal));
    Scal hs{}; hs.alpha = 0.7; hs.omega = 0.9; hs.rho = 1.0; hs.rho_new = 0.8; hs.mass = 1.0; hs.mass0 = 0.9; hs.wint = 1.0;
    cudaMemcpy(sc, &hs, sizeof(Scal), cudaMemcpyHostToDevice);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int K = 20, reps = 20;
    const int eb = (int)((n + 255) / 256), rb = red_blocks(n);
    auto bench = [&](const char* name, double bytes_per_node, std::function<void()> launch) {
        cudaGraph_t gr; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < K; ++k) launch();
        cudaStreamEndCapture(s, &gr); cudaGraphInstantiate(&ge, gr, 0);
        for (int w = 0; w < 2; ++w) cudaGraphLaunch(ge, s);
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = 1e3 * ms / (reps * K);
        printf("%-28s %7.2f us   %7.1f GB/s  (%g B/node)  %s\n", name, us, bytes_per_node * n / us * 1e-3, bytes_per_node, cudaGetErrorString(cudaGetLastError()));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(gr);
    };
    double** b = buf;
    bench("copy_kernel", 16, [&] { copy_kernel<<<rb, 256, 0, s>>>(b[0], b[1], n); });
    bench("solve_w_kernel", 32, [&] { solve_w_kernel<<<eb, 256, 0, s>>>(b[0], b[1], b[2], b[3], n, 1000.0); });
    bench("step_setup_kernel", 56, [&] { step_setup_kernel<<<eb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], g, ph, 1e-2); });
    bench("residual_kernel", 56, [&] { residual_kernel<<<rb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], g, ph, 1e-2, sc, red.part, ticket, nullptr); });
    bench("schur_rhs_kernel", 24, [&] { schur_rhs_kernel<<<eb, 256, 0, s>>>(b[0], b[1], b[2], g, sc, 100.0, 5e-5, 1e-22); });
    bench("bicg_init_kernel", 24, [&] { bicg_init_kernel<<<rb, 256, 0, s>>>(b[0], b[0], b[1], b[2], n, sc, red.part, ticket, 0, 0); });
    bench("bicg_x_kernel", 72, [&] { cudaMemsetAsync(&sc->done, 0, 4, s); bicg_x_kernel<<<rb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], n, sc, red.part, ticket, 0, 0); });
    bench("bicg_close_kernel", 32, [&] { bicg_close_kernel<<<rb, 256, 0, s>>>(b[0], b[1], b[2], n, sc, red.part, ticket); });
    bench("dmu_ceiling_kernel", 64, [&] { dmu_ceiling_kernel<<<rb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], g, ph, sc, red.part, ticket, b[5], b[6], b[7], 5.0); });
    bench("trial_kernel", 48, [&] { trial_kernel<<<eb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], n, 0.5); });
    bench("clip_mass_kernel", 16, [&] { clip_mass_kernel<<<rb, 256, 0, s>>>(b[0], b[1], g, ph, 1e-6, sc, 0, red.part, ticket); });
    bench("mass_shift_kernel", 16, [&] { mass_shift_kernel<<<eb, 256, 0, s>>>(b[0], g, ph, 1.0, sc); });
    bench("adj_rhs_kernel", 64, [&] { adj_rhs_kernel<<<rb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], g, ph, 1e-2, 5.0, sc, red.part, ticket, 1e-22); });
    bench("adj_qr_kernel", 40, [&] { adj_qr_kernel<<<eb, 256, 0, s>>>(b[0], b[1], b[2], b[3], b[4], g, 0.9, 0.1); });
    bench("grad_prox_kernel(1 level)", 24, [&] { grad_prox_kernel<<<rb, 256, 0, s>>>(b[0], b[1], nullptr, b[2], n, 1e-4, 50.0, 1e-4, -1.0, 1.0, (double*)sc, red.part, ticket, 0); });
    const Tiling tl = make_tiling(g); const int tg = tl.grid();
    bench("step_setup_tile(+solve_w)", 88, [&] { step_setup_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[1], b[2], b[3], b[8], b[9], 1000.0, b[4], b[5], b[6], g, tl, ph, 1e-2); });
    bench("residual_tile_kernel", 56, [&] { residual_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], g, tl, ph, 1e-2, sc, red.part, ticket, nullptr, 100.0, 5e-5, 1e-22); });
    bench("dmu_close_tile (close=0)", 64, [&] { dmu_close_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[12], b[8], b[9], b[1], b[2], b[3], b[4], g, tl, ph, sc, red.part, ticket, b[5], b[6], b[7], 5.0, 0); });
    { Scal h2 = hs; h2.iters = 3; cudaMemcpy(sc, &h2, sizeof(Scal), cudaMemcpyHostToDevice); }
    bench("dmu_close_tile (close=1)", 88, [&] { dmu_close_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[12], b[8], b[9], b[1], b[2], b[3], b[4], g, tl, ph, sc, red.part, ticket, b[5], b[6], b[7], 5.0, 1); });
    bench("adj_rhs_tile (init=1)", 80, [&] { adj_rhs_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[10], b[11], b[7], g, tl, ph, 1e-2, 5.0, sc, red.part, ticket, 1e-22, 1); });
    bench("adj_qr_tile (+copy)", 48, [&] { adj_qr_tile_kernel<<<tg, kTileThreads, 0, s>>>(b[0], b[5], b[1], b[2], b[3], b[4], g, tl, 0.9, 0.1); });
    printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
