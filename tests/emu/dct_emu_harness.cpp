// CPU run of the DCT / BiCGStab-fused transform kernels (csrc/vch_dct.cuh) under tests/emu/cuda_emu.h — TEST INFRASTRUCTURE.
// Built by tests/test_kernel_emulation.py; the binary runs the same
// seeded cases and writes every output array to a file.  The test compares the default build with NumPy/SciPy (so the kernels'
// logic is pinned on the CPU) and the experimental build with the default build.
//   usage: dct_emu_harness <log2L: 6|7> <out.bin>
#define VCH_CPU_EMU 1
#define VCH_CPU_EMU_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "../../sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200/csrc/vch_dct.cuh"

#include <cstdint>
#include <random>
#include <string>

using namespace vch;

static FILE* g_out = nullptr;
static void dump(const char* name, const double* p, size_t n) {
    char tag[32] = {0};
    snprintf(tag, sizeof(tag), "%s", name);
    uint64_t cnt = n;
    fwrite(tag, 1, 32, g_out); fwrite(&cnt, 8, 1, g_out); fwrite(p, 8, n, g_out);
}

struct Red {   // partials buffer with its Comm header, ticket, scalars
    std::vector<char> buf;
    double* part;
    unsigned int ticket = 0;
    Scal sc;
    Red() : buf(kCommHeaderBytes + 8 * 4096 * sizeof(double), 0) {
        Comm cm;                       // rank 0 of 1
        memcpy(buf.data(), &cm, sizeof(Comm));
        part = reinterpret_cast<double*>(buf.data() + kCommHeaderBytes);
        memset(&sc, 0, sizeof(sc));
    }
};

template <int LG>
struct Plan {
    static constexpr int Lf = 1 << LG, N = Lf / 2, n = N + 1, tpf = Lf / 8;
    int pitch = (n + 3) & ~3;
    std::vector<double2> tw;
    std::vector<double> lam, t1;
    double h = 1.0 / N;
    Plan() : tw(Lf), lam(n), t1((size_t)n * ((n + 3) & ~3), 0.0) {
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
    void rows(const double* in, double* out, int in_ls, int out_ls, const RowPrologue& pro, const DotEpilogue& epi, const int* done) {
        const int ppb = rows_ppb(n), threads = ppb * tpf, grid = ((n + 1) / 2 + ppb - 1) / ppb;
        const SymbolArgs nosym{1.0, 0.0, nullptr, 0.0, nullptr};
        const Scatter sct;
        vch_emu::launch(grid, threads, smem(ppb), [&] {
            dct_fft_kernel<LG, false, 512>(in, out, n, n, in_ls, 1, out_ls, 1, ppb, tw.data(), nullptr, nullptr, nosym, 1.0, 0, pro, epi, done, sct);
        });
    }
    void cols(double* buf, const SymbolArgs& sy, int scale_mode, const int* done) {
        const int ppb = cols_ppb(n), threads = ppb * tpf, grid = ((n + 1) / 2 + ppb - 1) / ppb;
        const double norm = 1.0 / (4.0 * (double)N * (double)N);
        const Scatter sct;
        vch_emu::launch(grid, threads, smem(ppb), [&] {
            dct_fft_kernel<LG, true, 512>(buf, buf, n, n, 1, pitch, 1, pitch, ppb, tw.data(), lam.data(), lam.data(), sy, norm, scale_mode,
                                          RowPrologue(), DotEpilogue(), done, sct);
        });
    }
    // DctPlan::apply, all-FFT path
    void apply(const double* in, double* out, const SymbolArgs& sy, const int* done, const DotEpilogue& epi, const RowPrologue& pro, int scale_mode) {
        rows(in, t1.data(), n, pitch, pro, DotEpilogue(), done);
        cols(t1.data(), sy, scale_mode, done);
        rows(t1.data(), out, pitch, n, RowPrologue(), epi, done);
    }
};

template <int LG>
static void run() {
    Plan<LG> P;
    const int n = P.n; const size_t nn = (size_t)n * n;
    std::mt19937_64 rng(1234 + LG);
    std::normal_distribution<double> nd(0.0, 1.0);
    auto rnd = [&](size_t c) { std::vector<double> v(c); for (auto& x : v) x = nd(rng); return v; };
    std::vector<double> in = rnd(nn), r = rnd(nn), q = rnd(nn), r0 = rnd(nn), a = rnd(nn), out(nn), w(nn), out2(nn);
    for (auto& x : a) x = 7.0 + 0.5 * x;
    dump("in", in.data(), nn); dump("r", r.data(), nn); dump("q", q.data(), nn); dump("r0", r0.data(), nn); dump("a", a.data(), nn);
    dump("lam", P.lam.data(), n);
    Red red;
    int done = 0;

    // A. plain forward row transform (pitched output) and the exact constant-coefficient solve P^-1 in
    {
        std::vector<double> t((size_t)n * P.pitch, 0.0);
        P.rows(in.data(), t.data(), n, P.pitch, RowPrologue(), DotEpilogue(), nullptr);
        dump("A_rows", t.data(), t.size());
        const SymbolArgs sy{100.0, 5e-5, nullptr, 7.5, nullptr};
        P.apply(in.data(), out.data(), sy, nullptr, DotEpilogue(), RowPrologue(), 0);
        dump("A_solve", out.data(), nn);
        P.apply(in.data(), out.data(), sy, nullptr, DotEpilogue(), RowPrologue(), 1);
        dump("A_lamsolve", out.data(), nn);
    }
    // B. forward BiCGStab operator application 1: prologue mode 1 (p = r + beta q, x = (a - abar) p), epilogue mode 1 with r
    auto scal_init = [&](double rho_new, double rho, double alpha, double omega) {
        memset(&red.sc, 0, sizeof(Scal));
        red.sc.rho_new = rho_new; red.sc.rho = rho; red.sc.alpha = alpha; red.sc.omega = omega; red.sc.abar = 7.25;
        red.sc.c0 = 100.0; red.sc.c2 = 5e-5; red.sc.rr = 3.0; red.sc.thr2 = 1e-22; red.sc.tol2 = 1e-22;
    };
    const SymbolArgs sy{0.0, 0.0, &red.sc.abar, 0.0, &red.sc.c0};
    {
        scal_init(1.7, 0.9, 0.6, 1.3);
        std::fill(w.begin(), w.end(), -7.0);
        P.apply(r.data(), out.data(), sy, &done, DotEpilogue{1, r0.data(), &red.sc, red.part, &red.ticket, w.data(), nullptr, r.data()},
                RowPrologue{1, r.data(), q.data(), a.data(), w.data(), &red.sc}, 1);
        dump("B_p", w.data(), nn); dump("B_v", out.data(), nn);
        double s[6] = {red.sc.alpha, red.sc.r0v, red.sc.rho, (double)red.sc.done, (double)red.sc.half, 0.0};
        dump("B_scal", s, 6);
    }
    // C. operator application 2: prologue mode 2 (s = r - alpha v), epilogue mode 2
    {
        scal_init(1.7, 0.9, 0.6, 1.3);
        std::fill(w.begin(), w.end(), -7.0);
        P.apply(r.data(), out.data(), sy, &done, DotEpilogue{2, w.data(), &red.sc, red.part, &red.ticket, w.data(), nullptr},
                RowPrologue{2, r.data(), q.data(), a.data(), w.data(), &red.sc}, 1);
        dump("C_s", w.data(), nn); dump("C_t", out.data(), nn);
        double s[6] = {red.sc.omega, red.sc.ts, red.sc.tt, (double)red.sc.done, 0.0, 0.0};
        dump("C_scal", s, 6);
    }
    // D. first iteration: beta = 0 (alpha = 0) with a poisoned q — nothing of q may reach p or v
    {
        scal_init(1.7, 1.0, 0.0, 1.0);
        std::vector<double> qnan(nn, std::nan(""));
        P.apply(r.data(), out.data(), sy, &done, DotEpilogue{1, r0.data(), &red.sc, red.part, &red.ticket, w.data(), nullptr, r.data()},
                RowPrologue{1, r.data(), qnan.data(), a.data(), w.data(), &red.sc}, 1);
        dump("D_p", w.data(), nn); dump("D_v", out.data(), nn);
    }
    // E. adjoint (right-preconditioned) form: no multiply in the prologue, (a - abar) in the epilogue
    {
        scal_init(1.7, 0.9, 0.6, 1.3);
        P.apply(r.data(), out.data(), sy, &done, DotEpilogue{1, r0.data(), &red.sc, red.part, &red.ticket, w.data(), a.data(), nullptr},
                RowPrologue{1, r.data(), q.data(), nullptr, w.data(), &red.sc}, 1);
        dump("E_p", w.data(), nn); dump("E_v", out.data(), nn);
        double s[6] = {red.sc.alpha, red.sc.r0v, 0, 0, 0, 0};
        dump("E_scal", s, 6);
        P.apply(r.data(), out2.data(), sy, &done, DotEpilogue{2, w.data(), &red.sc, red.part, &red.ticket, w.data(), a.data()},
                RowPrologue{2, r.data(), out.data(), nullptr, w.data(), &red.sc}, 1);
        dump("E_s", w.data(), nn); dump("E_t", out2.data(), nn);
        double s2[6] = {red.sc.omega, red.sc.ts, red.sc.tt, 0, 0, 0};
        dump("E_scal2", s2, 6);
    }
    // F. half-step exit: v = r / alpha_expected makes s = r - alpha v tiny; (r,r) inside the trust region
    {
        scal_init(2.0, 1.0, 1.0, 1.0);
        // operator with a == abar is the identity: v = p = r (beta = 0), alpha = rho_new / (r0, r)
        std::vector<double> aconst(nn, 7.25);
        double rr = 0.0, r0r = 0.0;
        for (size_t i = 0; i < nn; ++i) { rr += r[i] * r[i]; }
        red.sc.alpha = 0.0;            // beta = 0
        // choose r0 = r so that alpha = rho_new/(r,r); set rho_new = (r,r) -> alpha = 1 -> s = 0
        red.sc.rho_new = rr; red.sc.rr = rr; red.sc.thr2 = 1e-3 * rr;      // (r,r) <= 1e6 thr2: the formula is trusted
        (void)r0r;
        done = 0; red.sc.done = 0;
        P.apply(r.data(), out.data(), sy, &red.sc.done, DotEpilogue{1, r.data(), &red.sc, red.part, &red.ticket, w.data(), nullptr, r.data()},
                RowPrologue{1, r.data(), q.data(), aconst.data(), w.data(), &red.sc}, 1);
        double s[6] = {red.sc.alpha, (double)red.sc.done, (double)red.sc.half, red.sc.r0v / rr, 0, 0};
        dump("F_scal", s, 6);
        // with done set, the next application must leave its outputs untouched
        std::vector<double> keep(nn, 42.0), keepw(nn, 43.0);
        P.apply(r.data(), keep.data(), sy, &red.sc.done, DotEpilogue{2, keepw.data(), &red.sc, red.part, &red.ticket, keepw.data(), nullptr},
                RowPrologue{2, r.data(), out.data(), aconst.data(), keepw.data(), &red.sc}, 1);
        dump("F_keep", keep.data(), nn); dump("F_keepw", keepw.data(), nn);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <6|7> <out.bin>\n", argv[0]); return 2; }
    g_out = fopen(argv[2], "wb");
    if (!g_out) return 3;
    const int lg = atoi(argv[1]);
    if (lg == 6) run<6>();
    else if (lg == 7) run<7>();
    else return 4;
    fclose(g_out);
    return 0;
}
