// 1D engine (batched ensembles).  One CTA owns one control problem and keeps its whole state in shared memory:
// the time loop, the Newton iteration, the line search and the linear solves all run inside a single kernel, so an
// ensemble of B problems is ONE launch with no host round trips.  In 1D the Schur-reduced Newton operator
// (1/dt) I - L (diag(d) - kappa/2 L) and the adjoint operator I - (tau + dt/2 f'') L + dt/2 L^2 are pentadiagonal,
// so the reference's dense LAPACK solves (Forward_solver.py:185, backward_solver.py:116) become banded eliminations.
#include "vch_common.cuh"
#include <algorithm>

using namespace vch;

struct vch1d_ctx {
    vch1d_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;   // private non-blocking work stream (the legacy default stream serialises against every
                                     // other blocking stream of the process, e.g. NCCL's)
    cudaStream_t user = nullptr;     // caller's stream; ordered with `stream` by events at entry/exit of every call
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    long long launches = 0;
    // scratch of the entry points, kept across calls (grow-only): a cudaMalloc / cudaFree pair per call costs a device-wide
    // synchronisation and, for the 100 MB adjoint scratch of a 1024-problem ensemble, occasional 0.3 s stalls
    vch::DevBuf s_hist, s_dts, s_p, s_q, s_small;
};

namespace {
struct Scope1D {
    vch1d_ctx* c;
    explicit Scope1D(vch1d_ctx* ctx) : c(ctx) {
        VCH_CUDA(cudaSetDevice(c->device));
        VCH_CUDA(cudaEventRecord(c->ev_in, c->user));
        VCH_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in, 0));
    }
    ~Scope1D() {
        cudaEventRecord(c->ev_out, c->stream);
        cudaStreamWaitEvent(c->user, c->ev_out, 0);
    }
};
}  // namespace

namespace {

struct P1 {
    int n;                       // nodes
    double a;                    // 1/h^2
    double h, Lx, tau, gamma, c1, c2, kappa, lim, eps_log;
};

__device__ __forceinline__ double lap1(const double* v, int i, int n, double a) {
    const double c = v[i];
    const double l = v[i > 0 ? i - 1 : 1], r = v[i < n - 1 ? i + 1 : n - 2];
    return ((r - c) + (l - c)) * a;
}
// entry (i, j) of the mirror-ghost Neumann matrix (Forward_solver.py:64-76)
__device__ __forceinline__ double Lij(int i, int j, int n, double a) {
    if (j < 0 || j >= n) return 0.0;
    if (i == j) return -2.0 * a;
    if (i == 0) return (j == 1) ? 2.0 * a : 0.0;
    if (i == n - 1) return (j == n - 2) ? 2.0 * a : 0.0;
    return (j == i - 1 || j == i + 1) ? a : 0.0;
}
__device__ __forceinline__ double flog1(double phi, double eps) {
    const double s = fmin(fmax(phi, -1.0 + eps), 1.0 - eps);
    return log((1.0 + s) / (1.0 - s));
}
__device__ __forceinline__ double fpp1(double phi, double c1, double c2) {
    const double s = fmin(fmax(phi, -1.0 + 1e-8), 1.0 - 1e-8);
    return 2.0 * c1 / (1.0 - s * s) - 2.0 * c2;
}

// Block-wide reduction for blockDim <= 256 (threads_for), result in every thread, TWO barriers: the kernels below are barrier-bound
// (ncu: the barrier is their first stall reason) and spent four per reduction (block_red's two + a broadcast through shared memory).
// The warp partials are combined by every thread itself, in the order of block_red's second-stage xor-shuffle tree (o = 16, 8 meet
// only identity lanes; 4, 2, 1 pair partials i and i ^ o), so the result has the same bits as before.  Measured on the 1024-problem
// ensemble: 6.05 -> 5.99 ms per iteration — with 7 CTAs per SM the other CTAs already cover most of a barrier wait.
template <int OP> __device__ double bred(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_red<OP>(v);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = (i < nw) ? sh[i] : red_identity<OP>();
        x[i] = red_op<OP>(red_op<OP>(x[i], red_identity<OP>()), red_identity<OP>());
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = red_op<OP>(x[i], x[i + 4]);
    x[0] = red_op<OP>(x[0], x[2]); x[1] = red_op<OP>(x[1], x[3]);
    const double r = red_op<OP>(x[0], x[1]);
    __syncthreads();                       // sh is free for the next reduction
    return r;
}

// Pentadiagonal solve by block parallel cyclic reduction, executed by the whole CTA.
// Bands e2,e1,d0,f1,f2 (sub-sub .. super-super; entries that would fall outside the matrix are zero) and rhs b -> x.
// Rows are paired into 2x2 blocks (x_{2i}, x_{2i+1}), which makes the matrix block-tridiagonal; every reduction step
// eliminates the couplings at distance s in all block rows at once:
//     al = -A_i B_{i-s}^-1,  ga = -C_i B_{i+s}^-1
//     A_i <- al A_{i-s},  C_i <- ga C_{i+s},  B_i <- B_i + al C_{i-s} + ga A_{i+s},  d_i <- d_i + al d_{i-s} + ga d_{i+s}
// and after ceil(log2(nb)) steps x_i = B_i^-1 d_i.  Round 1 ran a Thomas-type elimination on ONE thread of the CTA (~13 000
// dependent instructions, ~45 us per Newton iteration, measured: 20 ms for the forward sweep of the 1024-problem ensemble);
// this takes 7 steps of ~150 instructions at N = 128.  No pivoting in either; against LAPACK's pivoted LU on the Newton and
// adjoint matrices of the default problem the solutions agree to 1e-13 / 6e-11 (cond 2e4 / 2e7; NumPy prototype of the same
// recurrences; the adjoint sweep, whose q = -L p amplifies the error of p by 1/h^2, adds one step of iterative refinement —
// penta_refine — which brings it back to the level of the pivoted solve).  Scratch: two buffers of 14*nb doubles, nb = (n+1)/2,
// structure-of-arrays; Y must be distinct from everything, X may overlap the bands and b (they are copied into Y first) and x.
__device__ __forceinline__ void inv2(const double* m, double* r) {     // m, r: row-major 2x2
    const double idet = 1.0 / (m[0] * m[3] - m[1] * m[2]);
    r[0] = m[3] * idet; r[1] = -m[1] * idet; r[2] = -m[2] * idet; r[3] = m[0] * idet;
}
__device__ __forceinline__ void mm2(const double* a, const double* b, double* c) {   // c = a b
    c[0] = a[0] * b[0] + a[1] * b[2]; c[1] = a[0] * b[1] + a[1] * b[3];
    c[2] = a[2] * b[0] + a[3] * b[2]; c[3] = a[2] * b[1] + a[3] * b[3];
}
__device__ void penta_pcr(const double* e2, const double* e1, const double* d0, const double* f1, const double* f2, const double* b,
                          double* X, double* Y, double* x, int n) {
    const int nb = (n + 1) >> 1;
    // record of block row i: field k (0..3 A, 4..7 B, 8..11 C, 12..13 d) at buf[k * nb + i]
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const int r0 = 2 * i, r1 = r0 + 1;
        const bool has1 = r1 < n;
        double rec[14] = {e2[r0], e1[r0], 0.0, has1 ? e2[r1] : 0.0,
                          d0[r0], has1 ? f1[r0] : 0.0, has1 ? e1[r1] : 0.0, has1 ? d0[r1] : 1.0,
                          has1 ? f2[r0] : 0.0, 0.0, has1 ? f1[r1] : 0.0, has1 ? f2[r1] : 0.0,
                          b[r0], has1 ? b[r1] : 0.0};
#pragma unroll
        for (int k = 0; k < 14; ++k) Y[k * nb + i] = rec[k];
    }
    __syncthreads();
    double* src = Y; double* dst = X;
    for (int sft = 1; sft < nb; sft <<= 1) {
        for (int i = threadIdx.x; i < nb; i += blockDim.x) {
            double A[4], B[4], Cm[4], d[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) { A[k] = src[k * nb + i]; B[k] = src[(4 + k) * nb + i]; Cm[k] = src[(8 + k) * nb + i]; }
            d[0] = src[12 * nb + i]; d[1] = src[13 * nb + i];
            double An[4] = {0.0, 0.0, 0.0, 0.0}, Cn[4] = {0.0, 0.0, 0.0, 0.0};
            const int lo = i - sft, hi = i + sft;
            if (lo >= 0) {
                double Bl[4], Bi[4], al[4], Al[4], Cl[4], t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { Al[k] = src[k * nb + lo]; Bl[k] = src[(4 + k) * nb + lo]; Cl[k] = src[(8 + k) * nb + lo]; }
                inv2(Bl, Bi); mm2(A, Bi, al);
#pragma unroll
                for (int k = 0; k < 4; ++k) al[k] = -al[k];
                mm2(al, Al, An); mm2(al, Cl, t);
#pragma unroll
                for (int k = 0; k < 4; ++k) B[k] += t[k];
                const double dl0 = src[12 * nb + lo], dl1 = src[13 * nb + lo];
                d[0] += al[0] * dl0 + al[1] * dl1; d[1] += al[2] * dl0 + al[3] * dl1;
            }
            if (hi < nb) {
                double Bh[4], Bi[4], ga[4], Ah[4], Ch[4], t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { Ah[k] = src[k * nb + hi]; Bh[k] = src[(4 + k) * nb + hi]; Ch[k] = src[(8 + k) * nb + hi]; }
                inv2(Bh, Bi); mm2(Cm, Bi, ga);
#pragma unroll
                for (int k = 0; k < 4; ++k) ga[k] = -ga[k];
                mm2(ga, Ch, Cn); mm2(ga, Ah, t);
#pragma unroll
                for (int k = 0; k < 4; ++k) B[k] += t[k];
                const double dh0 = src[12 * nb + hi], dh1 = src[13 * nb + hi];
                d[0] += ga[0] * dh0 + ga[1] * dh1; d[1] += ga[2] * dh0 + ga[3] * dh1;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { dst[k * nb + i] = An[k]; dst[(4 + k) * nb + i] = B[k]; dst[(8 + k) * nb + i] = Cn[k]; }
            dst[12 * nb + i] = d[0]; dst[13 * nb + i] = d[1];
        }
        __syncthreads();
        double* t = src; src = dst; dst = t;
    }
    // decoupled: x_i = B_i^-1 d_i.  The results go through registers and a barrier because x may live inside X.
    double xr[4][2];
    int cnt = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x, ++cnt) {
        double B[4], Bi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) B[k] = src[(4 + k) * nb + i];
        inv2(B, Bi);
        const double d0v = src[12 * nb + i], d1v = src[13 * nb + i];
        xr[cnt][0] = Bi[0] * d0v + Bi[1] * d1v; xr[cnt][1] = Bi[2] * d0v + Bi[3] * d1v;
    }
    __syncthreads();
    cnt = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x, ++cnt) {
        x[2 * i] = xr[cnt][0];
        if (2 * i + 1 < n) x[2 * i + 1] = xr[cnt][1];
    }
    __syncthreads();
}

// One step of iterative refinement for penta_pcr: x += A^-1 (b - A x) with the residual in fp64.  res and dx: scratch vectors.
__device__ void penta_refine(const double* e2, const double* e1, const double* d0, const double* f1, const double* f2, const double* b,
                             double* X, double* Y, double* x, double* res, double* dx, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = b[i] - d0[i] * x[i];
        if (i >= 1) acc -= e1[i] * x[i - 1];
        if (i >= 2) acc -= e2[i] * x[i - 2];
        if (i + 1 < n) acc -= f1[i] * x[i + 1];
        if (i + 2 < n) acc -= f2[i] * x[i + 2];
        res[i] = acc;
    }
    __syncthreads();
    penta_pcr(e2, e1, d0, f1, f2, res, X, Y, dx, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] += dx[i];
    __syncthreads();
}

struct Sm {   // shared-memory carve-up: 22 arrays of n doubles + the second scratch buffer of the pentadiagonal solver
    double *phi, *mu, *phi0, *mu0, *w0, *w1, *cphi, *cmu, *Rphi, *Rmu, *d, *dphi, *dmu, *phit, *mut;
    double *e2, *e1, *d0, *f1, *f2, *b, *tmp;
    // penta_pcr scratch (two buffers of 14 * ((n + 1) / 2) <= 8 n doubles for n >= 8): Y behind the 22 arrays; X inside them —
    // forward Newton: the 9 arrays phit..tmp, free during a solve (the bands are copied into Y before X is written);
    // adjoint sweep: the 8 arrays w1..dmu, which it does not use (its bands must survive for the refinement step).
    // 30 KB per CTA at N = 128: 7 CTAs per SM, so the 1024-problem ensemble is a single wave on 148 SMs.
    double *pcrY, *pcrX_fwd, *pcrX_adj;
    __device__ void carve(double* base, int n) {
        double** f[] = {&phi, &mu, &phi0, &mu0, &w0, &w1, &cphi, &cmu, &Rphi, &Rmu, &d, &dphi, &dmu, &phit, &mut,
                        &e2, &e1, &d0, &f1, &f2, &b, &tmp};
        for (int k = 0; k < 22; ++k) *f[k] = base + (size_t)k * n;
        pcrY = base + (size_t)22 * n;
        const bool roomy = n >= 8;                     // tiny grids: 14 nb > 8 n, both X buffers move behind Y
        pcrX_fwd = roomy ? phit : pcrY + 14 * ((n + 1) / 2);
        pcrX_adj = roomy ? w1 : pcrY + 14 * ((n + 1) / 2);
    }
};
constexpr int kSmArrays = 22;

__device__ double residual1(const Sm& s, const double* phi, const double* mu, double* Rphi, double* Rmu, const P1& p,
                            double dt, double* sh) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        const double f = phi[i];
        const double rp = p.tau / dt * f - 0.5 * p.kappa * lap1(phi, i, p.n, p.a) + p.c1 * flog1(f, p.eps_log) - 0.5 * mu[i] + s.cphi[i];
        const double rm = f / dt - 0.5 * lap1(mu, i, p.n, p.a) + s.cmu[i];
        Rphi[i] = rp; Rmu[i] = rm;
        acc += rp * rp + rm * rm;
    }
    return sqrt(bred<0>(acc, sh));
}

// Newton solve on the state in shared memory (Forward_solver.py:139-235).  phi0/mu0/w0/w1 given; result in phi/mu.
// status: 0 converged / max-iter, 1 line-search failure (the reference returns the last iterate), 3 non-finite.
__device__ int newton1(Sm& s, const P1& p, double dt, double* hist, int hist_cap, int* n_hist, double* sh) {
    const int n = p.n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double f = s.phi0[i], m = s.mu0[i];
        s.cphi[i] = -p.tau * f / dt - 0.5 * p.kappa * lap1(s.phi0, i, n, p.a) - 2.0 * p.c2 * f - 0.5 * m - 0.5 * (s.w1[i] + s.w0[i]);
        s.cmu[i] = -f / dt - 0.5 * lap1(s.mu0, i, n, p.a);
        s.phi[i] = f; s.mu[i] = m;
    }
    __syncthreads();
    int nh = 0, status = 0;
    double normR = residual1(s, s.phi, s.mu, s.Rphi, s.Rmu, p, dt, sh);
    for (int k = 0; k < 50; ++k) {
        if (hist && nh < hist_cap && threadIdx.x == 0) hist[nh] = normR;
        ++nh;
        // (the reference's DEBUG mass-defect print, Forward_solver.py:166-170, has no effect on the iterates; a non-finite residual —
        // what its sum over R_mu would reveal — shows in the norm)
        if (k % 10 == 0 && !isfinite(normR)) { status = 3; break; }
        if (normR < 1e-6) break;
        // Schur system: (1/dt) I - L (diag(d) - kappa/2 L),  rhs = -Rmu + L Rphi
        for (int i = threadIdx.x; i < n; i += blockDim.x) s.d[i] = p.tau / dt + 2.0 * p.c1 / (1.0 - s.phi[i] * s.phi[i]);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double band[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int j = i - 2 + q;
                double acc = 0.0;
                if (j >= 0 && j < n) {
                    for (int m = i - 1; m <= i + 1; ++m) {
                        if (m < 0 || m >= n) continue;
                        const double lim_ = Lij(i, m, n, p.a);
                        const double kmj = ((m == j) ? s.d[m] : 0.0) - 0.5 * p.kappa * Lij(m, j, n, p.a);
                        acc += lim_ * kmj;
                    }
                    acc = ((i == j) ? 1.0 / dt : 0.0) - acc;
                }
                band[q] = acc;
            }
            s.e2[i] = band[0]; s.e1[i] = band[1]; s.d0[i] = band[2]; s.f1[i] = band[3]; s.f2[i] = band[4];
            s.b[i] = lap1(s.Rphi, i, n, p.a) - s.Rmu[i];
        }
        __syncthreads();
        penta_pcr(s.e2, s.e1, s.d0, s.f1, s.f2, s.b, s.pcrX_fwd, s.pcrY, s.dphi, n);
        double ap = INFINITY, an = INFINITY;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double dp = s.dphi[i];
            s.dmu[i] = 2.0 * (s.d[i] * dp - 0.5 * p.kappa * lap1(s.dphi, i, n, p.a) + s.Rphi[i]);
            if (dp > 0.0) ap = fmin(ap, (p.lim - s.phi[i]) / dp);
            else if (dp < 0.0) an = fmin(an, (-p.lim - s.phi[i]) / dp);
        }
        double amax = bred<1>(fmin(ap, an), sh);      // min over both ceilings in one reduction
        if (!isfinite(amax) || amax <= 0.0) amax = 1.0;
        double alpha = fmin(1.0, 0.9 * amax);
        bool accepted = false;
        for (int ls = 0; ls < 12; ++ls) {
            double mx = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double f = s.phi[i] + alpha * s.dphi[i];
                s.phit[i] = f; s.mut[i] = s.mu[i] + alpha * s.dmu[i];
                mx = fmax(mx, fabs(f));
                if (!(fabs(f) < p.lim)) mx = INFINITY;   // also catches NaN
            }
            mx = bred<2>(mx, sh);
            if (mx < p.lim) {
                const double nt = residual1(s, s.phit, s.mut, s.d0, s.f1, p, dt, sh);   // trial residuals into scratch bands
                if (nt <= (1.0 - 1e-3 * alpha) * normR) {
                    for (int i = threadIdx.x; i < n; i += blockDim.x) {
                        s.phi[i] = s.phit[i]; s.mu[i] = s.mut[i]; s.Rphi[i] = s.d0[i]; s.Rmu[i] = s.f1[i];
                    }
                    __syncthreads();
                    normR = nt; accepted = true;
                    break;
                }
            }
            alpha *= 0.5;
        }
        if (!accepted) { status = 1; break; }
    }
    if (n_hist && threadIdx.x == 0) *n_hist = nh;
    __syncthreads();
    return status;
}

// Occupancy of the two time-loop kernels (one CTA per problem).  Uncapped they take 106 / 112 registers and ran with 160 threads
// (one per node): 3 CTAs per SM, 444 at a time, THREE waves for the 1024-problem ensemble.  They are barrier-bound (block cyclic
// reduction over (n + 1) / 2 block rows), not register-hungry: with one thread per BLOCK ROW (96 threads at N = 128, threads_for) and
// 80 registers (12 - 64 bytes of spills) 7 CTAs fit an SM — 1036 at a time, the ensemble is ONE wave.  Measured on B200, ms per
// ensemble iteration (profiles/r02_1d_occupancy.txt): 8.79 uncapped / 160 threads, 7.21 with 96 registers, 7.97 with 96 threads alone,
// 6.06 with both at 80 registers (169 000 problem-iterations/s instead of 116 000).
#ifndef VCH_1D_REGS
#define VCH_1D_REGS __maxnreg__(80)
#endif
// Whole forward solve of one problem per CTA (Forward_solver.py:286-386).
__global__ void VCH_1D_REGS forward1d_kernel(P1 p, const double* __restrict__ phi_init, const double* __restrict__ u, int u_rows,
                                 int n_steps, const double* __restrict__ dts, double* __restrict__ phi_hist,
                                 double* __restrict__ mu_hist, double* __restrict__ w_hist, int* __restrict__ status_out) {
    extern __shared__ double smem[];
    __shared__ double sh[32];
    Sm s; s.carve(smem, p.n);
    const int n = p.n, prob = blockIdx.x;
    const double* ub = u ? u + (size_t)prob * u_rows * n : nullptr;
    double* hist = phi_hist + (size_t)prob * (n_steps + 2) * n;
    double m0 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double f = phi_init[(size_t)prob * n + i];
        s.phi0[i] = f; s.w0[i] = 0.0;
        hist[i] = f; hist[n + i] = f;                       // level 0 stored twice (:329-336)
        m0 += p.h * ((i == 0 || i == n - 1) ? 0.5 : 1.0) * f;
    }
    m0 = bred<0>(m0, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x)       // mu_0 = initialize_mu(phi_0, 0)  (:82-86, :324)
        s.mu0[i] = -p.kappa * lap1(s.phi0, i, n, p.a) + p.c1 * flog1(s.phi0[i], p.eps_log) - 2.0 * p.c2 * s.phi0[i] - s.w0[i];
    __syncthreads();
    int status = 0;
    for (int step = 0; step < n_steps; ++step) {
        const double dt = dts[step], gdt = p.gamma / dt;
        const int r0 = step, r1 = (step < u_rows - 1) ? step + 1 : step;       // last row repeats (:347-353)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double un = ub ? ub[(size_t)min(r0, u_rows - 1) * n + i] : 0.0;
            const double un1 = ub ? ub[(size_t)min(r1, u_rows - 1) * n + i] : 0.0;
            s.w1[i] = __ddiv_rn(__dadd_rn(__dmul_rn(gdt - 0.5, s.w0[i]), __dmul_rn(0.5, __dadd_rn(un1, un))), gdt + 0.5);
        }
        __syncthreads();
        const int st = newton1(s, p, dt, nullptr, 0, nullptr, sh);
        if (st == 3) { status = VCH_E_NONFINITE; break; }
        double mass = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double f = fmin(fmax(s.phi[i], -p.lim), p.lim);
            s.phi[i] = f;
            mass += p.h * ((i == 0 || i == n - 1) ? 0.5 : 1.0) * f;
        }
        mass = bred<0>(mass, sh);
        const double shift = (mass - m0) / p.Lx;               // uniform shift, no re-clip (:364-366)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double f = s.phi[i] - shift;
            s.phi0[i] = f; s.mu0[i] = s.mu[i]; s.w0[i] = s.w1[i];
            hist[(size_t)(step + 2) * n + i] = f;
            if (mu_hist) mu_hist[((size_t)prob * n_steps + step) * n + i] = s.mu[i];
            if (w_hist) w_hist[((size_t)prob * n_steps + step) * n + i] = s.w1[i];
        }
        __syncthreads();
    }
    if (status_out && threadIdx.x == 0) status_out[prob] = status;
}

__global__ void newton1d_kernel(P1 p, const double* __restrict__ phi_old, const double* __restrict__ mu_old,
                                const double* __restrict__ w_old, const double* __restrict__ w_new, double dt,
                                double* __restrict__ phi_new, double* __restrict__ mu_new, double* __restrict__ hist,
                                int hist_cap, int* __restrict__ n_hist, int* __restrict__ status_out) {
    extern __shared__ double smem[];
    __shared__ double sh[32];
    Sm s; s.carve(smem, p.n);
    const int n = p.n, prob = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s.phi0[i] = phi_old[(size_t)prob * n + i]; s.mu0[i] = mu_old[(size_t)prob * n + i];
        s.w0[i] = w_old[(size_t)prob * n + i]; s.w1[i] = w_new[(size_t)prob * n + i];
    }
    __syncthreads();
    const int st = newton1(s, p, dt, hist ? hist + (size_t)prob * hist_cap : nullptr, hist_cap, n_hist ? n_hist + prob : nullptr, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        phi_new[(size_t)prob * n + i] = s.phi[i]; mu_new[(size_t)prob * n + i] = s.mu[i];
    }
    if (status_out && threadIdx.x == 0) status_out[prob] = (st == 3) ? VCH_E_NONFINITE : VCH_OK;
}

__global__ void residual1d_kernel(P1 p, const double* __restrict__ phi, const double* __restrict__ phi0,
                                  const double* __restrict__ mu, const double* __restrict__ mu0,
                                  const double* __restrict__ w1, const double* __restrict__ w0, double dt,
                                  double* __restrict__ Rphi, double* __restrict__ Rmu, long long total) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long prob = idx / p.n; const int i = (int)(idx - prob * p.n);
        const double *f = phi + prob * p.n, *f0 = phi0 + prob * p.n, *m = mu + prob * p.n, *m0 = mu0 + prob * p.n;
        Rphi[idx] = p.tau * (f[i] - f0[i]) / dt - 0.5 * p.kappa * (lap1(f, i, p.n, p.a) + lap1(f0, i, p.n, p.a))
                    + (p.c1 * flog1(f[i], p.eps_log) - 2.0 * p.c2 * f0[i]) - 0.5 * (m[i] + m0[i]) - 0.5 * (w1[idx] + w0[idx]);
        Rmu[idx] = (f[i] - f0[i]) / dt - 0.5 * (lap1(m, i, p.n, p.a) + lap1(m0, i, p.n, p.a));
    }
}

__global__ void mu_init1d_kernel(P1 p, const double* __restrict__ phi, const double* __restrict__ w, double* __restrict__ mu,
                                 long long total) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long prob = idx / p.n; const int i = (int)(idx - prob * p.n);
        const double* f = phi + prob * p.n;
        mu[idx] = -p.kappa * lap1(f, i, p.n, p.a) + p.c1 * flog1(f[i], p.eps_log) - 2.0 * p.c2 * f[i] - w[idx];
    }
}

// Whole adjoint sweep of one problem per CTA (backward_solver.py:72-125).
__global__ void VCH_1D_REGS adjoint1d_kernel(P1 p, const double* __restrict__ phi_hist, int levels, const double* __restrict__ t,
                                 const double* __restrict__ b1v, const double* __restrict__ b2v,
                                 const double* __restrict__ phiQ, const double* __restrict__ phiT,
                                 double* __restrict__ pO, double* __restrict__ qO, double* __restrict__ rO) {
    extern __shared__ double smem[];
    Sm s; s.carve(smem, p.n);
    const int n = p.n, prob = blockIdx.x;
    const size_t base = (size_t)prob * levels * n;
    const double* F = phi_hist + base;
    const double* Q = phiQ ? phiQ + base : nullptr;
    const double* Tt = phiT ? phiT + (size_t)prob * n : nullptr;
    double *P = pO + base, *Qo = qO + base, *R = rO + base;
    const double b1 = b1v[prob], b2 = b2v[prob];
    for (size_t e = threadIdx.x; e < (size_t)levels * n; e += blockDim.x) { P[e] = 0.0; Qo[e] = 0.0; R[e] = 0.0; }
    // terminal: (I - tau L) p_M = b2 (phi_M - phi_T): tridiagonal, solved as a pentadiagonal with empty outer bands
    const int M = levels - 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s.e2[i] = 0.0; s.f2[i] = 0.0;
        s.e1[i] = -p.tau * Lij(i, i - 1, n, p.a); s.f1[i] = -p.tau * Lij(i, i + 1, n, p.a); s.d0[i] = 1.0 + 2.0 * p.tau * p.a;
        s.b[i] = b2 * (F[(size_t)M * n + i] - (Tt ? Tt[i] : 0.0));
    }
    __syncthreads();
    penta_pcr(s.e2, s.e1, s.d0, s.f1, s.f2, s.b, s.pcrX_adj, s.pcrY, s.phi, n);   // s.phi := p_{k+1}
    penta_refine(s.e2, s.e1, s.d0, s.f1, s.f2, s.b, s.pcrX_adj, s.pcrY, s.phi, s.phi0, s.mu0, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s.mu[i] = -lap1(s.phi, i, n, p.a);      // s.mu := q_{k+1}
        s.w0[i] = 0.0;                           // s.w0 := r_{k+1}
        P[(size_t)M * n + i] = s.phi[i]; Qo[(size_t)M * n + i] = s.mu[i];
    }
    __syncthreads();
    for (int k = M - 1; k >= 0; --k) {
        const double dt = t[k + 1] - t[k];
        if (dt <= 0.0) continue;                 // level stays zero; p_{k+1} for the next level is then the zero row
        const double hdt = 0.5 * dt;
        // p1/q1/r1 are the stored rows k+1 (zeros if that level was skipped), exactly as the reference reads them
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            s.phi[i] = P[(size_t)(k + 1) * n + i]; s.mu[i] = Qo[(size_t)(k + 1) * n + i]; s.w0[i] = R[(size_t)(k + 1) * n + i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double f1 = F[(size_t)(k + 1) * n + i], f0 = F[(size_t)k * n + i];
            const double src = hdt * b1 * ((f0 - (Q ? Q[(size_t)k * n + i] : 0.0)) + (f1 - (Q ? Q[(size_t)(k + 1) * n + i] : 0.0)));
            // B p1 = p1 - tau L p1 - dt/2 L^2 p1 + dt/2 f''(phi1) L p1, evaluated from p1 itself (q rows may be zero-skipped)
            const double lp = lap1(s.phi, i, n, p.a);
            s.tmp[i] = lp;
            s.b[i] = s.phi[i] - p.tau * lp + hdt * fpp1(f1, p.c1, p.c2) * lp + src;
            const double av = p.tau + hdt * fpp1(f0, p.c1, p.c2);
            double band[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int j = i - 2 + q;
                double acc = 0.0;
                if (j >= 0 && j < n) {
                    double l2 = 0.0;
                    for (int m = i - 1; m <= i + 1; ++m) if (m >= 0 && m < n) l2 += Lij(i, m, n, p.a) * Lij(m, j, n, p.a);
                    acc = ((i == j) ? 1.0 : 0.0) - av * Lij(i, j, n, p.a) + hdt * l2;
                }
                band[q] = acc;
            }
            s.e2[i] = band[0]; s.e1[i] = band[1]; s.d0[i] = band[2]; s.f1[i] = band[3]; s.f2[i] = band[4];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) s.b[i] -= hdt * lap1(s.tmp, i, n, p.a);
        __syncthreads();
        penta_pcr(s.e2, s.e1, s.d0, s.f1, s.f2, s.b, s.pcrX_adj, s.pcrY, s.phit, n);
        penta_refine(s.e2, s.e1, s.d0, s.f1, s.f2, s.b, s.pcrX_adj, s.pcrY, s.phit, s.phi0, s.mu0, n);
        const double den = p.gamma + hdt, fb = (p.gamma - hdt) / den, fs = hdt / den;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double qv = -lap1(s.phit, i, n, p.a);
            P[(size_t)k * n + i] = s.phit[i]; Qo[(size_t)k * n + i] = qv;
            R[(size_t)k * n + i] = fb * s.w0[i] + fs * (qv + s.mu[i]);
        }
        __syncthreads();
    }
}

// J1..J4 per problem with np.trapezoid weights (cost_and_function.py:55-75).
__global__ void cost1d_kernel(int n, int levels, const double* __restrict__ phi, const double* __restrict__ u,
                              const double* __restrict__ Q, const double* __restrict__ phiT,
                              const double* __restrict__ wx, const double* __restrict__ wt,
                              const double* __restrict__ weights, double* __restrict__ J) {
    __shared__ double sh[32];
    const int prob = blockIdx.x;
    const size_t base = (size_t)prob * levels * n;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    for (size_t e = threadIdx.x; e < (size_t)levels * n; e += blockDim.x) {
        const int tl = (int)(e / n), i = (int)(e - (size_t)tl * n);
        const double w = wt[tl] * wx[i];
        const double f = phi[base + e];
        const double d = f - (Q ? Q[base + e] : 0.0);
        v0 += w * d * d;
        if (tl == levels - 1) { const double dd = f - (phiT ? phiT[(size_t)prob * n + i] : 0.0); v1 += wx[i] * dd * dd; }
        if (u) { const double uv = u[base + e]; v2 += w * uv * uv; v3 += w * fabs(uv); }
    }
    v0 = bred<0>(v0, sh); v1 = bred<0>(v1, sh); v2 = bred<0>(v2, sh); v3 = bred<0>(v3, sh);
    if (threadIdx.x == 0) {
        const double* w = weights + (size_t)prob * 4;
        const double J1 = 0.5 * w[0] * v0, J2 = 0.5 * w[1] * v1, J3 = 0.5 * w[2] * v2, J4 = w[3] * v3;
        double* o = J + (size_t)prob * 5;
        o[0] = J1 + J2 + J3 + J4; o[1] = J1; o[2] = J2; o[3] = J3; o[4] = J4;
    }
}

__global__ void grad_prox1d_kernel(long long per, const double* __restrict__ u, const double* __restrict__ r,
                                   const double* __restrict__ par, double* __restrict__ un, double* __restrict__ red) {
    __shared__ double sh[32];
    const int prob = blockIdx.x;
    const double* q = par + (size_t)prob * 6;
    const double b3 = q[0], alpha = q[1], ksp = q[2], umin = q[3], umax = q[4], thr = alpha * ksp;
    const size_t base = (size_t)prob * per;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    for (long long e = threadIdx.x; e < per; e += blockDim.x) {
        const double uv = u[base + e];
        const double y = __dsub_rn(uv, __dmul_rn(alpha, __dadd_rn(r[base + e], __dmul_rn(b3, uv))));
        const double m = fmax(__dsub_rn(fabs(y), thr), 0.0);
        double sv = (y > 0.0) ? m : ((y < 0.0) ? -m : 0.0);
        sv = fmin(fmax(sv, umin), umax);
        un[base + e] = sv;
        const double d = sv - uv;
        v0 += d * d; v1 += uv * uv; v2 += (sv != 0.0) ? 1.0 : 0.0; v3 += (sv == umin || sv == umax) ? 1.0 : 0.0;
    }
    v0 = bred<0>(v0, sh); v1 = bred<0>(v1, sh); v2 = bred<0>(v2, sh); v3 = bred<0>(v3, sh);
    if (threadIdx.x == 0) { double* o = red + (size_t)prob * 4; o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; }
}

P1 make_p1(const vch1d_params& q) {
    P1 p;
    p.n = q.N + 1; p.a = 1.0 / (q.h * q.h); p.h = q.h; p.Lx = q.Lx; p.tau = q.tau; p.gamma = q.gamma; p.c1 = q.c1; p.c2 = q.c2;
    p.kappa = q.kappa; p.lim = 1.0 - q.delta_sep; p.eps_log = std::max(1e-8, 0.5 * q.delta_sep);
    return p;
}
// one thread per block row of the cyclic reduction (the node loops take two trips), whole warps, 64 .. 256
int threads_for(int n) { int t = (((n + 1) / 2 + 31) / 32) * 32; return std::min(std::max(t, 64), 256); }
size_t smem_for(int n) {          // 22 state arrays + the solver's scratch buffer Y (and X for tiny grids, see Sm)
    const size_t nb = (size_t)(n + 1) / 2;
    return ((size_t)kSmArrays * n + (n >= 8 ? 14 : 28) * nb) * sizeof(double);
}

std::vector<double> trapz_w(const double* x, int n) {
    std::vector<double> w(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) { const double h = 0.5 * (x[i + 1] - x[i]); w[i] += h; w[i + 1] += h; }
    return w;
}

struct IntOut {   // device int array mirrored to a host int array
    int* d = nullptr; int* h; size_t n;
    IntOut(int* host, size_t count) : h(host), n(count) { if (h) VCH_CUDA(cudaMalloc(&d, n * sizeof(int))); }
    void fetch(cudaStream_t s) { if (h) { VCH_CUDA(cudaMemcpyAsync(h, d, n * sizeof(int), cudaMemcpyDeviceToHost, s)); VCH_CUDA(cudaStreamSynchronize(s)); } }
    ~IntOut() { if (d) cudaFree(d); }
};

void prep(vch1d_ctx* c, const void* kernel, int n) {
    VCH_CUDA(cudaSetDevice(c->device));
    VCH_REQUIRE(smem_for(n) <= 220 * 1024, VCH_E_SHAPE, "1D grid too large for the shared-memory resident solver (N+1 <= 1280)");
    VCH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(n)));
}

}  // namespace

extern "C" {

int vch1d_create(const vch1d_params* p, int device, vch1d_ctx** out) {
    return guarded([&] {
        VCH_REQUIRE(p && out, VCH_E_ARG, "null argument");
        VCH_REQUIRE(p->N >= 2 && p->h > 0, VCH_E_SHAPE, "N must be >= 2 and h positive");
        VCH_REQUIRE(vch_device_count() > device, VCH_E_CUDA, "no CUDA device: vch_b200 has no CPU fallback");
        VCH_CUDA(cudaSetDevice(device));
        auto* c = new vch1d_ctx();
        c->prm = *p; c->device = device;
        VCH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        VCH_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
        VCH_CUDA(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
        *out = c;
        return VCH_OK;
    });
}
void vch1d_destroy(vch1d_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(c->ev_in); cudaEventDestroy(c->ev_out); cudaStreamDestroy(c->stream);
    delete c;
}
int vch1d_set_stream(vch1d_ctx* c, void* s) {
    return guarded([&] { VCH_REQUIRE(c, VCH_E_ARG, "null ctx"); c->user = (cudaStream_t)s; return VCH_OK; });
}
long long vch1d_launch_count(vch1d_ctx* c) { return c ? c->launches : 0; }

int vch1d_residual(vch1d_ctx* c, int batch, const double* phi_new, const double* phi_old, const double* mu_new,
                   const double* mu_old, const double* w_new, const double* w_old, double dt, double* Rphi_out,
                   double* Rmu_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi_new && phi_old && mu_new && mu_old && w_new && w_old && Rphi_out && Rmu_out,
                    VCH_E_SHAPE, "residual: bad arguments");
        VCH_CUDA(cudaSetDevice(c->device));
        const P1 p = make_p1(c->prm);
        const long long tot = (long long)batch * p.n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *a = st.in(phi_new, tot), *b = st.in(phi_old, tot), *m1 = st.in(mu_new, tot), *m0 = st.in(mu_old, tot),
                     *w1 = st.in(w_new, tot), *w0 = st.in(w_old, tot);
        double *rp = st.out(Rphi_out, tot), *rm = st.out(Rmu_out, tot);
        residual1d_kernel<<<red_blocks(tot), 256, 0, c->stream>>>(p, a, b, m1, m0, w1, w0, dt, rp, rm, tot);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch1d_initialize_mu(vch1d_ctx* c, int batch, const double* phi, const double* w, double* mu_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi && w && mu_out, VCH_E_SHAPE, "initialize_mu: bad arguments");
        VCH_CUDA(cudaSetDevice(c->device));
        const P1 p = make_p1(c->prm);
        const long long tot = (long long)batch * p.n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *a = st.in(phi, tot), *b = st.in(w, tot);
        double* o = st.out(mu_out, tot);
        mu_init1d_kernel<<<red_blocks(tot), 256, 0, c->stream>>>(p, a, b, o, tot);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch1d_newton(vch1d_ctx* c, int batch, const double* phi_old, const double* mu_old, const double* w_old,
                 const double* w_new, double dt, double* phi_new_out, double* mu_new_out, double* res_hist, int hist_cap,
                 int* n_hist, int* status_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi_old && mu_old && w_old && w_new && phi_new_out && mu_new_out, VCH_E_SHAPE,
                    "newton: bad arguments");
        const P1 p = make_p1(c->prm);
        prep(c, (const void*)newton1d_kernel, p.n);
        const size_t tot = (size_t)batch * p.n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *p0 = st.in(phi_old, tot), *m0 = st.in(mu_old, tot), *w0 = st.in(w_old, tot), *w1 = st.in(w_new, tot);
        double *po = st.out(phi_new_out, tot), *mo = st.out(mu_new_out, tot);
        DevBuf& hist = c->s_hist; if (res_hist) hist.alloc((size_t)batch * hist_cap);
        IntOut nh(n_hist, batch), so(status_out, batch);
        newton1d_kernel<<<batch, threads_for(p.n), smem_for(p.n), c->stream>>>(p, p0, m0, w0, w1, dt, po, mo, res_hist ? hist.p : (double*)nullptr, hist_cap, nh.d, so.d);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        if (res_hist) VCH_CUDA(cudaMemcpyAsync(res_hist, hist.p, (size_t)batch * hist_cap * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        st.finish();
        nh.fetch(c->stream); so.fetch(c->stream);
        if (status_out) for (int b = 0; b < batch; ++b) VCH_REQUIRE(status_out[b] != VCH_E_NONFINITE, VCH_E_NONFINITE,
                                                                   "Non-finite mass_defect; check phi bounds/log regularization.");
        return VCH_OK;
    });
}

int vch1d_forward(vch1d_ctx* c, int batch, const double* phi0, const double* u, int u_rows, int n_steps,
                  const double* dt_steps, double* phi_hist_out, double* mu_hist_out, double* w_hist_out, int* status_out,
                  vch_stats* stats, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi0 && dt_steps && phi_hist_out && n_steps >= 0, VCH_E_SHAPE, "forward: bad arguments");
        VCH_REQUIRE(!u || u_rows >= 1, VCH_E_SHAPE, "forward: control needs at least one row");
        const P1 p = make_p1(c->prm);
        prep(c, (const void*)forward1d_kernel, p.n);
        const size_t n = p.n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double* d0 = st.in(phi0, (size_t)batch * n);
        const double* du = st.in(u, (size_t)batch * u_rows * n);
        double* dh = st.out(phi_hist_out, (size_t)batch * (n_steps + 2) * n);
        double* dm = st.out(mu_hist_out, (size_t)batch * n_steps * n);
        double* dw = st.out(w_hist_out, (size_t)batch * n_steps * n);
        DevBuf& dts = c->s_dts; dts.alloc(std::max(1, n_steps));
        VCH_CUDA(cudaMemcpyAsync(dts.p, dt_steps, n_steps * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        std::vector<int> local_status(batch, 0);
        IntOut so(status_out ? status_out : local_status.data(), batch);
        forward1d_kernel<<<batch, threads_for(p.n), smem_for(p.n), c->stream>>>(p, d0, du, u_rows, n_steps, dts.p, dh, dm, dw, so.d);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        st.finish();
        so.fetch(c->stream);
        if (stats) stats->kernel_launches += 1;
        for (int b = 0; b < batch; ++b)
            VCH_REQUIRE(so.h[b] != VCH_E_NONFINITE, VCH_E_NONFINITE, "Non-finite mass_defect; check phi bounds/log regularization.");
        return VCH_OK;
    });
}

int vch1d_adjoint(vch1d_ctx* c, int batch, const double* phi_hist, int levels, const double* t_hist, const double* b1,
                  const double* b2, const double* phiQ, const double* phiT, double* p_out, double* q_out, double* r_out,
                  int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi_hist && t_hist && b1 && b2 && levels >= 1 && r_out, VCH_E_SHAPE, "adjoint: bad arguments");
        const P1 p = make_p1(c->prm);
        prep(c, (const void*)adjoint1d_kernel, p.n);
        const size_t n = p.n, tot = (size_t)batch * levels * n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *dh = st.in(phi_hist, tot), *dq = st.in(phiQ, tot), *dT = st.in(phiT, (size_t)batch * n);
        double *po = st.out(p_out, tot), *qo = st.out(q_out, tot), *ro = st.out(r_out, tot);
        DevBuf &ptmp = c->s_p, &qtmp = c->s_q;
        if (!po) { ptmp.alloc(tot); po = ptmp.p; }
        if (!qo) { qtmp.alloc(tot); qo = qtmp.p; }
        DevBuf& small = c->s_small; small.alloc((size_t)levels + 2 * batch);
        double *dt_ = small.p, *db1 = dt_ + levels, *db2 = db1 + batch;
        VCH_CUDA(cudaMemcpyAsync(dt_, t_hist, levels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(db1, b1, batch * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(db2, b2, batch * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        adjoint1d_kernel<<<batch, threads_for(p.n), smem_for(p.n), c->stream>>>(p, dh, levels, dt_, db1, db2, dq, dT, po, qo, ro);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        st.finish();
        return VCH_OK;
    });
}

int vch1d_cost(vch1d_ctx* c, int batch, const double* phi_hist, const double* u, const double* phiQ, const double* phiT,
               int levels, const double* x, const double* t_hist, const double* weights, double* J_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && phi_hist && x && t_hist && weights && J_out && levels >= 1, VCH_E_SHAPE, "cost: bad arguments");
        VCH_CUDA(cudaSetDevice(c->device));
        const int n = c->prm.N + 1;
        const size_t tot = (size_t)batch * levels * n;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *dh = st.in(phi_hist, tot), *du = st.in(u, tot), *dq = st.in(phiQ, tot), *dT = st.in(phiT, (size_t)batch * n);
        std::vector<double> wx = trapz_w(x, n), wt = trapz_w(t_hist, levels);
        DevBuf& small = c->s_small; small.alloc((size_t)n + levels + 9 * (size_t)batch);
        double *dwx = small.p, *dwt = dwx + n, *dwe = dwt + levels, *dJ = dwe + 4 * (size_t)batch;
        VCH_CUDA(cudaMemcpyAsync(dwx, wx.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(dwt, wt.data(), levels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        VCH_CUDA(cudaMemcpyAsync(dwe, weights, 4 * (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        cost1d_kernel<<<batch, 256, 0, c->stream>>>(n, levels, dh, du, dq, dT, dwx, dwt, dwe, dJ);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        VCH_CUDA(cudaMemcpyAsync(J_out, dJ, 5 * (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        st.finish();
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        return VCH_OK;
    });
}

int vch1d_grad_prox(vch1d_ctx* c, int batch, long long per_problem, const double* u, const double* r, const double* par,
                    double* u_new_out, double* red_out, int mem) {
    return guarded([&] {
        VCH_REQUIRE(c && batch >= 1 && per_problem >= 1 && u && r && par && u_new_out, VCH_E_SHAPE, "grad_prox: bad arguments");
        VCH_CUDA(cudaSetDevice(c->device));
        const size_t tot = (size_t)batch * per_problem;
        Scope1D scope(c);
        Stager st(c->stream, mem);
        const double *du = st.in(u, tot), *dr = st.in(r, tot);
        double* dn = st.out(u_new_out, tot);
        DevBuf& small = c->s_small; small.alloc(10 * (size_t)batch);
        double *dpar = small.p, *dred = dpar + 6 * (size_t)batch;
        VCH_CUDA(cudaMemcpyAsync(dpar, par, 6 * (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        grad_prox1d_kernel<<<batch, 256, 0, c->stream>>>(per_problem, du, dr, dpar, dn, dred);
        ++c->launches;
        VCH_CUDA(cudaGetLastError());
        if (red_out) VCH_CUDA(cudaMemcpyAsync(red_out, dred, 4 * (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        st.finish();
        VCH_CUDA(cudaStreamSynchronize(c->stream));
        return VCH_OK;
    });
}

}  // extern "C"
