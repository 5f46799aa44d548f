// CPU run of one Newton linear solve and one adjoint step exactly as csrc/vch2d.cu enqueues them (residual -> Schur right-hand
// side -> P^-1 b -> BiCGStab start -> iterations until the device-side done flag -> delta-mu / step ceiling / trial iterate; adjoint
// right-hand side -> right-preconditioned BiCGStab -> closing P^-1 y), under tests/emu/cuda_emu.h.  TEST INFRASTRUCTURE.
//   usage: krylov_emu_harness <log2L: 6|7> <rel_tol> <half_exit: 0|1> <out.bin>
#define VCH_CPU_EMU 1
#define VCH_CPU_EMU_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "../../sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200/csrc/vch2d_kernels.cuh"

#include <cstdint>
#include <random>

using namespace vch;

static FILE* g_out = nullptr;
static void dump(const char* name, const double* p, size_t n) {
    char tag[32] = {0};
    snprintf(tag, sizeof(tag), "%s", name);
    uint64_t cnt = n;
    fwrite(tag, 1, 32, g_out); fwrite(&cnt, 8, 1, g_out); fwrite(p, 8, n, g_out);
}

template <typename F> static void launch(int grid, int block, F&& f) { vch_emu::launch((unsigned)grid, (unsigned)block, 0, f); }

template <int LG>
struct Plan {   // the all-FFT path of DctPlan::apply (csrc/vch_dct.cuh)
    static constexpr int Lf = 1 << LG, N = Lf / 2, n = N + 1, tpf = Lf / 8;
    int pitch = (n + 3) & ~3;
    std::vector<double2> tw;
    std::vector<double> lam, t1;
    explicit Plan(double h) : tw(Lf), lam(n), t1((size_t)n * ((n + 3) & ~3), 0.0) {
        for (int m = 0; m < Lf; ++m) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * m / Lf;
            tw[m] = make_double2((double)cosl(a), (double)sinl(a));
        }
        for (int k = 0; k < n; ++k) {
            const long double s = sinl(3.14159265358979323846264338327950288L * k / (2.0L * N));
            lam[k] = (double)(4.0L * s * s / ((long double)h * h));
        }
    }
    static size_t smem(int ppb) { return sizeof(double2) * ((size_t)ppb * (Lf + (Lf >> 3) + 1) + kTwLo + (Lf >> 5)); }
    static int rows_ppb(int lines) { int p = 256 / tpf; if (p < 1) p = 1; const int pairs = (lines + 1) / 2; return p > pairs ? pairs : p; }
    static int cols_ppb(int ncols) {
        int p = 512 / tpf; if (p < 2 && tpf <= 256) p = 2; if (p < 1) p = 1; if (p > 8) p = 8;
        const int pairs = (ncols + 1) / 2; return p > pairs ? pairs : p;
    }
    void apply(const double* in, double* out, const SymbolArgs& sy, const int* done, const DotEpilogue& epi, const RowPrologue& pro, int scale_mode) {
        const SymbolArgs nosym{1.0, 0.0, nullptr, 0.0, nullptr};
        const Scatter sct;
        const int rp = rows_ppb(n), rt = rp * tpf, rg = ((n + 1) / 2 + rp - 1) / rp;
        const int cp = cols_ppb(n), ct = cp * tpf, cg = ((n + 1) / 2 + cp - 1) / cp;
        const double norm = 1.0 / (4.0 * (double)N * (double)N);
        double* t = t1.data();
        vch_emu::launch(rg, rt, smem(rp), [&] {
            dct_fft_kernel<LG, false, 512>(in, t, n, n, n, 1, pitch, 1, rp, tw.data(), nullptr, nullptr, nosym, 1.0, 0, pro, DotEpilogue(), done, sct);
        });
        vch_emu::launch(cg, ct, smem(cp), [&] {
            dct_fft_kernel<LG, true, 512>(t, t, n, n, 1, pitch, 1, pitch, cp, tw.data(), lam.data(), lam.data(), sy, norm, scale_mode,
                                          RowPrologue(), DotEpilogue(), done, sct);
        });
        vch_emu::launch(rg, rt, smem(rp), [&] {
            dct_fft_kernel<LG, false, 512>(t, out, n, n, pitch, 1, n, 1, rp, tw.data(), nullptr, nullptr, nosym, 1.0, 0, RowPrologue(), epi, done, sct);
        });
    }
};

template <int LG>
static void run(double rel_tol, int half_exit) {
    constexpr int N = (1 << LG) / 2, n1 = N + 1;
    const long long n = (long long)n1 * n1;
    const double h = 1.0 / N, dt = 1e-2;
    Plan<LG> P(h);
    Geo g; g.no = n1; g.ni = n1; g.nx1 = n1; g.ny1 = n1; g.n = n; g.iho2 = 1.0 / (h * h); g.ihi2 = 1.0 / (h * h); g.nxg = n1;
    Phys ph{0.05, 10.0, 0.75, 1.0, 1e-4, 0.99, 5e-3, 1.0 - 1e-4};
    std::mt19937_64 rng(99 + LG);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<double> phi(n), mu(n), cphi(n), cmu(n);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n1; ++j) {
            const double x = i * h, y = j * h;
            phi[(size_t)i * n1 + j] = 0.85 * tanh(3.0 * sin(4 * M_PI * x) * cos(2 * M_PI * y)) + 0.02 * nd(rng);
            mu[(size_t)i * n1 + j] = 0.3 * cos(2 * M_PI * x) * cos(M_PI * y) + 0.01 * nd(rng);
            cphi[(size_t)i * n1 + j] = -4.0 * phi[(size_t)i * n1 + j] + 0.05 * nd(rng);
            cmu[(size_t)i * n1 + j] = -100.0 * phi[(size_t)i * n1 + j] + 0.5 * nd(rng);
        }
    dump("phi", phi.data(), n); dump("mu", mu.data(), n); dump("cphi", cphi.data(), n); dump("cmu", cmu.data(), n);

    std::vector<char> redbuf(kCommHeaderBytes + 8 * 1024 * sizeof(double), 0);
    { Comm cm; memcpy(redbuf.data(), &cm, sizeof(Comm)); }
    double* part = reinterpret_cast<double*>(redbuf.data() + kCommHeaderBytes);
    unsigned int ticket = 0;
    Scal sc, mirror; memset(&sc, 0, sizeof(sc)); memset(&mirror, 0xff, sizeof(mirror));
    sc.maxit = 200;
    const int rb = red_blocks(n), eb = (int)((n + 255) / 256);
    const cudaGraphConditionalHandle cond = 0;

    std::vector<double> Rphi(n), Rmu(n), a(n), kb(n), kr(n), kr0(n), kp(n, NAN), kv(n, NAN), ks(n, NAN), kt(n, NAN), kq(n, NAN), kx(n, NAN);
    // ---- forward: residual, Schur right-hand side, solve
    launch(rb, kRedThreads, [&] { residual_kernel(phi.data(), mu.data(), cphi.data(), cmu.data(), Rphi.data(), Rmu.data(), a.data(), g, ph, dt, &sc, part, &ticket, &mirror); });
    dump("Rphi", Rphi.data(), n); dump("Rmu", Rmu.data(), n); dump("a", a.data(), n);
    { double s[6] = {sc.res2, sc.amin, sc.amax, sc.abar, sc.mu2, (double)(memcmp(&sc, &mirror, sizeof(Scal)) == 0)}; dump("res_scal", s, 6); }
    launch(eb, 256, [&] { schur_rhs_kernel(Rphi.data(), Rmu.data(), kb.data(), g, &sc, 1.0 / dt, 0.5 * ph.kappa, rel_tol * rel_tol); });
    dump("b", kb.data(), n);
    const SymbolArgs sy{0.0, 0.0, &sc.abar, 0.0, &sc.c0};
    P.apply(kb.data(), kr.data(), sy, nullptr, DotEpilogue(), RowPrologue(), 0);
    launch(rb, kRedThreads, [&] { bicg_init_kernel(kr.data(), kr.data(), kr0.data(), kx.data(), n, &sc, part, &ticket, cond, 0); });
    int launched = 0;
    while (!sc.done && launched < 60) {
        P.apply(kr.data(), kv.data(), sy, &sc.done, DotEpilogue{1, kr0.data(), &sc, part, &ticket, kp.data(), nullptr, half_exit ? kr.data() : nullptr},
                RowPrologue{1, kr.data(), kq.data(), a.data(), kp.data(), &sc}, 1);
        P.apply(kr.data(), kt.data(), sy, &sc.done, DotEpilogue{2, ks.data(), &sc, part, &ticket, ks.data(), nullptr},
                RowPrologue{2, kr.data(), kv.data(), a.data(), ks.data(), &sc}, 1);
        launch(rb, kRedThreads, [&] { bicg_x_kernel(kx.data(), kr.data(), kp.data(), ks.data(), kt.data(), kr0.data(), kv.data(), kq.data(), n, &sc, part, &ticket, cond, 0); });
        ++launched;
    }
    // one more (gated) iteration after convergence must not change anything — the polled path enqueues such iterations
    std::vector<double> x_before = kx;
    P.apply(kr.data(), kv.data(), sy, &sc.done, DotEpilogue{1, kr0.data(), &sc, part, &ticket, kp.data(), nullptr, half_exit ? kr.data() : nullptr},
            RowPrologue{1, kr.data(), kq.data(), a.data(), kp.data(), &sc}, 1);
    launch(rb, kRedThreads, [&] { bicg_x_kernel(kx.data(), kr.data(), kp.data(), ks.data(), kt.data(), kr0.data(), kv.data(), kq.data(), n, &sc, part, &ticket, cond, 0); });
    { double s[8] = {(double)sc.iters, (double)sc.done, (double)sc.half_exits, (double)sc.half, (double)sc.solves, (double)launched,
                     (double)(memcmp(x_before.data(), kx.data(), n * sizeof(double)) == 0), (double)sc.nonfinite}; dump("fwd_scal", s, 8); }
    dump("dphi", kx.data(), n);
    std::vector<double> dmu(n), phit(n), mut(n);
    launch(rb, kRedThreads, [&] { dmu_ceiling_kernel(kx.data(), a.data(), Rphi.data(), phi.data(), dmu.data(), g, ph, &sc, part, &ticket, mu.data(), phit.data(), mut.data(), ph.tau / dt); });
    dump("dmu", dmu.data(), n); dump("phit", phit.data(), n); dump("mut", mut.data(), n);
    { double s[2] = {sc.ceil_pos, sc.ceil_neg}; dump("ceil", s, 2); }

    // ---- adjoint step: right-hand side (also sets c0 = 1, c2 = dt/2, the tolerance), right-preconditioned solve, x = P^-1 y
    std::vector<double> p1(n), q1(n), phi1(n), rhs(n), aa(n);
    for (long long i = 0; i < n; ++i) { p1[i] = 2.0 * phi[i] + 0.1 * nd(rng); phi1[i] = phi[i] + 0.01 * nd(rng); }
    launch(eb, 256, [&] { lap_kernel(p1.data(), q1.data(), g, -1.0); });          // q1 = -L p1
    launch(rb, kRedThreads, [&] { adj_rhs_kernel(p1.data(), q1.data(), phi1.data(), phi.data(), nullptr, nullptr, rhs.data(), aa.data(), g, ph, dt, 5.0, &sc, part, &ticket, 1e-22); });
    dump("p1", p1.data(), n); dump("q1", q1.data(), n); dump("phi1", phi1.data(), n); dump("adj_rhs", rhs.data(), n); dump("adj_a", aa.data(), n);
    launch(rb, kRedThreads, [&] { bicg_init_kernel(rhs.data(), kr.data(), kr0.data(), kx.data(), n, &sc, part, &ticket, cond, 0); });
    launched = 0;
    while (!sc.done && launched < 60) {
        P.apply(kr.data(), kv.data(), sy, &sc.done, DotEpilogue{1, kr0.data(), &sc, part, &ticket, kp.data(), aa.data(), nullptr},
                RowPrologue{1, kr.data(), kq.data(), nullptr, kp.data(), &sc}, 1);
        P.apply(kr.data(), kt.data(), sy, &sc.done, DotEpilogue{2, ks.data(), &sc, part, &ticket, ks.data(), aa.data()},
                RowPrologue{2, kr.data(), kv.data(), nullptr, ks.data(), &sc}, 1);
        launch(rb, kRedThreads, [&] { bicg_x_kernel(kx.data(), kr.data(), kp.data(), ks.data(), kt.data(), kr0.data(), kv.data(), kq.data(), n, &sc, part, &ticket, cond, 0); });
        ++launched;
    }
    P.apply(kx.data(), kx.data(), sy, nullptr, DotEpilogue(), RowPrologue(), 0);
    { double s[4] = {(double)sc.iters, (double)sc.done, (double)sc.half_exits, sc.abar}; dump("adj_scal", s, 4); }
    dump("adj_p", kx.data(), n);
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s <6|7> <rel_tol> <half_exit> <out.bin>\n", argv[0]); return 2; }
    g_out = fopen(argv[4], "wb");
    if (!g_out) return 3;
    const int lg = atoi(argv[1]);
    const double tol = atof(argv[2]);
    const int he = atoi(argv[3]);
    if (lg == 6) run<6>(tol, he);
    else if (lg == 7) run<7>(tol, he);
    else return 4;
    fclose(g_out);
    return 0;
}
