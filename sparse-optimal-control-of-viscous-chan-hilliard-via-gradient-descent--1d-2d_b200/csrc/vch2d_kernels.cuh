// 2D kernels: Neumann stencils, Newton residual, Schur / adjoint operators, BiCGStab vector updates,
// post-step mass correction, adjoint recurrences.  All fp64, HBM/L2-bound; reductions are deterministic
// (vch_common.cuh).  Fields are flat arrays of n = no*ni doubles; the Laplacian acts on the (no, ni) view
// with the contiguous axis ni = Nx+1 (spacing hx) and the strided axis no = Ny+1 (spacing hy) — exactly the
// operator kron(I_{Ny+1}, L_x) + kron(L_y, I_{Nx+1}) of the reference (Forward2_solver.py:125-137).
#pragma once
#include "vch_common.cuh"
#include "vch_dct.cuh"

namespace vch {

struct Geo {
    int no, ni;          // Laplacian view (slab mode: no = rows owned by this rank)
    int nx1, ny1;        // array view (Nx+1, Ny+1) used by the trapezoid weights (slab mode: nx1 = owned rows)
    long long n;
    double iho2, ihi2;   // 1/h_outer^2, 1/h_inner^2
    // slab mode (square grids, rows [o0, o0+no) of nxg): ghost rows below (glo) / above (ghi) the owned rows hold the
    // neighbour rank's boundary rows, so the outer-axis mirror rule applies only at the global boundary
    int glo = 0, ghi = 0, o0 = 0, nxg = 0;
};

struct Phys {
    double tau, gamma, c1, c2, kappa, lim /* 1-delta_sep */, eps_log /* max(1e-8, delta_sep/2) */, dsq /* 1-delta_sep^2 */;
};

__device__ __forceinline__ double lap_g(const double* __restrict__ v, int o, int i, const Geo& g) {
    const double c = v[(long long)o * g.ni + i];
    const int im = (i > 0) ? i - 1 : 1, ip = (i < g.ni - 1) ? i + 1 : g.ni - 2;
    const int om = (o > 0 || g.glo) ? o - 1 : 1, op = (o < g.no - 1 || g.ghi) ? o + 1 : g.no - 2;
    const double a = (v[(long long)o * g.ni + ip] - c) + (v[(long long)o * g.ni + im] - c);
    const double b = (v[(long long)op * g.ni + i] - c) + (v[(long long)om * g.ni + i] - c);
    return a * g.ihi2 + b * g.iho2;
}

__device__ __forceinline__ double flory_log(double phi, double eps) {
    const double s = fmin(fmax(phi, -1.0 + eps), 1.0 - eps);
    return log((1.0 + s) / (1.0 - s));
}

// ---------------------------------------------------------------------------------- elementwise / stencil
__global__ void lap_kernel(const double* __restrict__ v, double* __restrict__ out, Geo g, double scale) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        out[idx] = scale * lap_g(v, o, i, g);
    }
}

// Field copy on the SMs.  The per-level copies of the sweeps must not go through cudaMemcpyAsync: device-to-device memcpys
// are served by a copy engine, where they queue behind the 270 MB host<->device chunks of the streamed host-buffer path and
// stall the work stream for milliseconds (measured: adjoint sweep 0.55 s instead of 0.33 s).
__global__ void copy_kernel(const double* __restrict__ src, double* __restrict__ dst, long long n) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
        dst[idx] = src[idx];
}

__global__ void solve_w_kernel(const double* __restrict__ w0, const double* __restrict__ un, const double* __restrict__ un1,
                               double* __restrict__ w1, long long n, double gdt) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double a = un ? un[idx] : 0.0, b = un1 ? un1[idx] : 0.0;
        // no FMA contraction: bit-identical to the reference's NumPy expression (its tests ask rtol 1e-15)
        w1[idx] = __ddiv_rn(__dadd_rn(__dmul_rn(gdt - 0.5, w0[idx]), __dmul_rn(0.5, __dadd_rn(b, a))), gdt + 0.5);
    }
}

__global__ void mu_init_kernel(const double* __restrict__ phi, const double* __restrict__ w, double* __restrict__ mu, Geo g, Phys p) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double f = phi[idx];
        mu[idx] = -p.kappa * lap_g(phi, o, i, g) + p.c1 * flory_log(f, p.eps_log) - 2.0 * p.c2 * f - w[idx];
    }
}

// Per-step constants of the Newton residual and the reference's initial guess for mu (Forward2_solver.py:351):
//   cphi = -tau phi0/dt - kappa/2 L phi0 - 2 c2 phi0 - mu0/2 - (w1+w0)/2,   cmu = -phi0/dt - 1/2 L mu0
__global__ void step_setup_kernel(const double* __restrict__ phi0, const double* __restrict__ mu0,
                                  const double* __restrict__ w0, const double* __restrict__ w1,
                                  double* __restrict__ cphi, double* __restrict__ cmu, double* __restrict__ mu_guess,
                                  Geo g, Phys p, double dt) {
    pdl_enter();
    const double idt = 1.0 / dt;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double f = phi0[idx], m = mu0[idx], wa = w0[idx], wb = w1[idx];
        const double lf = lap_g(phi0, o, i, g), lm = lap_g(mu0, o, i, g);
        cphi[idx] = -p.tau * f * idt - 0.5 * p.kappa * lf - 2.0 * p.c2 * f - 0.5 * m - 0.5 * (wb + wa);
        cmu[idx] = -f * idt - 0.5 * lm;
        if (mu_guess) mu_guess[idx] = -p.kappa * lf + p.c1 * flory_log(f, p.eps_log) - 2.0 * p.c2 * f - wb;
    }
}

// Device scalars -> mapped pinned host mirror (8-byte words; visible to the host after the stream synchronises).  Called by
// ALL threads of the last block of a reduction kernel, after thread 0 has written its totals: one word per thread, so the
// copy costs one L2 round trip instead of sizeof(Scal)/8 dependent ones (a single-thread loop added 6.5 us to the kernel).
__device__ __forceinline__ void publish_scalars(const Scal* src, Scal* dst_host) {
    static_assert(sizeof(Scal) % 8 == 0 && sizeof(Scal) / 8 <= 64, "Scal is copied in 8-byte words by the first 64 threads");
    __threadfence_block();
    __syncthreads();                       // thread 0's scalar stores are visible to the block
    if (threadIdx.x < sizeof(Scal) / 8) {
        const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(src) + threadIdx.x);
        reinterpret_cast<volatile unsigned long long*>(dst_host)[threadIdx.x] = v;
        __threadfence_system();
    }
}

// [R_phi; R_mu], the Jacobian diagonal a = tau/dt + 2c1/(1 - min(phi^2, 1-delta^2)) and ||R||^2, min a, max a.
// __launch_bounds__(256, 4): the grid is 4 CTAs per SM (red_blocks); at 67 registers only 3 fit and a quarter of the blocks ran as a
// second, nearly empty wave
__global__ void __launch_bounds__(kRedThreads, 4) residual_kernel(const double* __restrict__ phi, const double* __restrict__ mu,
                                const double* __restrict__ cphi, const double* __restrict__ cmu,
                                double* __restrict__ Rphi, double* __restrict__ Rmu, double* __restrict__ a,
                                Geo g, Phys p, double dt, Scal* sc, double* part, unsigned int* ticket, Scal* publish) {
    pdl_enter();
    const double idt = 1.0 / dt, tdt = p.tau / dt;
    double v[4] = {0.0, INFINITY, -INFINITY, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double f = phi[idx], m = mu[idx];
        const double rp = tdt * f - 0.5 * p.kappa * lap_g(phi, o, i, g) + p.c1 * flory_log(f, p.eps_log) - 0.5 * m + cphi[idx];
        const double rm = f * idt - 0.5 * lap_g(mu, o, i, g) + cmu[idx];
        const double d = tdt + 2.0 * p.c1 / (1.0 - fmin(f * f, p.dsq));
        Rphi[idx] = rp; Rmu[idx] = rm;
        if (a) a[idx] = d;
        v[0] += rp * rp + rm * rm;
        v[1] = fmin(v[1], d); v[2] = fmax(v[2], d);
        v[3] += m * m;
    }
    const int op[4] = {0, 1, 2, 0};
    double tot[4];
    if (grid_reduce<4>(v, op, part, ticket, tot)) {      // block-uniform: the whole last block enters
        if (threadIdx.x == 0) {
            sc->res2 = tot[0]; sc->amin = tot[1]; sc->amax = tot[2]; sc->abar = sqrt(tot[1] * tot[2]); sc->mu2 = tot[3];
            if (!isfinite(tot[0])) sc->nonfinite = 1;
        }
        // Newton's host-side decisions follow every residual evaluation: the last block writes the scalars straight into
        // the mapped pinned mirror (saves the separate publish launch in front of the host's stream synchronisation)
        if (publish) publish_scalars(sc, publish);
    }
}

// Reference-form residual for the test-level entry point (all six fields given explicitly).
__global__ void residual_full_kernel(const double* __restrict__ phi, const double* __restrict__ phi0,
                                     const double* __restrict__ mu, const double* __restrict__ mu0,
                                     const double* __restrict__ w1, const double* __restrict__ w0,
                                     double* __restrict__ Rphi, double* __restrict__ Rmu, Geo g, Phys p, double dt) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double f = phi[idx], f0 = phi0[idx];
        Rphi[idx] = p.tau * (f - f0) / dt - 0.5 * p.kappa * (lap_g(phi, o, i, g) + lap_g(phi0, o, i, g))
                    + (p.c1 * flory_log(f, p.eps_log) - 2.0 * p.c2 * f0) - 0.5 * (mu[idx] + mu0[idx]) - 0.5 * (w1[idx] + w0[idx]);
        Rmu[idx] = (f - f0) / dt - 0.5 * (lap_g(mu, o, i, g) + lap_g(mu0, o, i, g));
    }
}

// Jacobian diagonal only (test-level jacobian_solve).
__global__ void jac_diag_kernel(const double* __restrict__ phi, double* __restrict__ a, Geo g, Phys p, double dt,
                                Scal* sc, double* part, unsigned int* ticket) {
    pdl_enter();
    double v[2] = {INFINITY, -INFINITY};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const double f = phi[idx];
        const double d = p.tau / dt + 2.0 * p.c1 / (1.0 - fmin(f * f, p.dsq));
        a[idx] = d; v[0] = fmin(v[0], d); v[1] = fmax(v[1], d);
    }
    const int op[2] = {1, 2};
    double tot[2];
    if (grid_reduce<2>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->amin = tot[0]; sc->amax = tot[1]; sc->abar = sqrt(tot[0] * tot[1]);
    }
}

// Right-hand side of the Schur-reduced Newton system: b = -R_mu + L R_phi.
__global__ void schur_rhs_kernel(const double* __restrict__ Rphi, const double* __restrict__ Rmu, double* __restrict__ b, Geo g,
                                 Scal* sc, double c0, double c2, double tol2) {
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc->c0 = c0; sc->c2 = c2; sc->tol2 = tol2; sc->adj = 0; }   // coefficients and (squared) relative tolerance of the solve that follows
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        b[idx] = lap_g(Rphi, o, i, g) - Rmu[idx];
    }
}

// ---------------------------------------------------------------------------------- BiCGStab vector kernels
// cond/use_cond: CUDA-graph WHILE handle of the enclosing solve graph (device-driven Krylov loop); ignored when use_cond = 0.
__global__ void bicg_init_kernel(const double* r_in, double* r /* may equal r_in */, double* __restrict__ r0,
                                 double* __restrict__ x, long long n, Scal* sc,
                                 double* part, unsigned int* ticket, cudaGraphConditionalHandle cond, int use_cond) {
    pdl_enter();
    // r = r0 = r_in, x = 0, (r,r).  p, v and q are NOT cleared: the first iteration forms p = r + beta*q with beta = 0
    // (alpha starts at 0; the prologue skips the q term for a zero coefficient, so stale values cannot leak in), v is
    // written by the first operator application before anything reads it, q by the first x/r update.
    double acc[1] = {0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double rv = r_in[idx];
        if (r != r_in) r[idx] = rv;
        r0[idx] = rv; x[idx] = 0.0;
        acc[0] += rv * rv;
    }
    const int op[1] = {0};
    double tot[1];
    if (grid_reduce<1>(acc, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->bnorm2 = tot[0]; sc->rr = tot[0]; sc->rho_new = tot[0];
        sc->rho = 1.0; sc->alpha = 0.0; sc->omega = 1.0;
        sc->thr2 = sc->tol2 * tot[0];
        sc->iters = 0; sc->half = 0;
        sc->done = (tot[0] == 0.0 || !isfinite(tot[0])) ? 1 : 0;
        if (!isfinite(tot[0])) sc->nonfinite = 1;
        sc->solves += 1;
        if (use_cond) { sc->g_launches += use_cond; cudaGraphSetConditional(cond, sc->done ? 0u : 1u); }
    }
}

// x += alpha p + omega s, r = s - omega t, q = p - omega v (the next iteration's fused prologue forms p = r + beta q without an
// in-place hazard), with (r,r) and (r0,r); the p- and s-updates and the other dot products live in the DCT row kernels.
__global__ void bicg_x_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                              const double* __restrict__ s, const double* __restrict__ t, const double* __restrict__ r0,
                              const double* __restrict__ v, double* __restrict__ q, long long n, Scal* sc, double* part,
                              unsigned int* ticket, cudaGraphConditionalHandle cond, int use_cond) {
    pdl_enter();
    const int half = sc->half;
    if (sc->done && !half) return;
    const double al = sc->alpha, om = sc->omega;
    if (half) {   // the solve converged at the half step (dots_finish): x += alpha p closes it; s, t were not formed
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
            x[idx] += al * p[idx];
        double one[1] = {0.0};
        const int op1[1] = {0};
        double t1[1];
        if (grid_reduce<1>(one, op1, part, ticket, t1) && threadIdx.x == 0) {   // last block: every block has read sc->half
            sc->half = 0; sc->iters += 1; sc->iters_total += 1; sc->half_exits += 1;
            if (sc->iters > sc->iters_max) sc->iters_max = sc->iters;
            if (use_cond) { sc->g_launches += use_cond; cudaGraphSetConditional(cond, 0u); }
        }
        return;
    }
    double acc[2] = {0.0, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double sv = s[idx], pv = p[idx];
        if (q) q[idx] = pv - om * v[idx];
        x[idx] += al * pv + om * sv;
        const double rv = sv - om * t[idx];
        r[idx] = rv;
        acc[0] += rv * rv; acc[1] += r0[idx] * rv;
    }
    const int op[2] = {0, 0};
    double tot[2];
    if (grid_reduce<2>(acc, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->rr = tot[0]; sc->rho_new = tot[1]; sc->iters += 1; sc->iters_total += 1;
        const bool bad = !isfinite(tot[0]) || !isfinite(tot[1]);
        if (bad) sc->nonfinite = 1;
        if (bad || tot[0] <= sc->thr2) sc->done = 1;
        if (sc->iters > sc->iters_max) sc->iters_max = sc->iters;
        if (use_cond) {
            sc->g_launches += use_cond;
            const bool stop = sc->done || sc->iters >= sc->maxit;
            if (stop && !sc->done) { sc->stalls += 1; if (sc->adj) sc->stalls_adj += 1; }
            cudaGraphSetConditional(cond, stop ? 0u : 1u);
        }
    }
}

// 6-launch iteration: the iterate lags one update behind (the x/r update lives in the next iteration's first row
// transform); this kernel applies the last one after the loop:  x += alpha p + omega s, or x += alpha p after a half-step exit.
// A solve that ended without an iteration (zero right-hand side) leaves x = 0.
__global__ void bicg_close_kernel(double* __restrict__ x, const double* __restrict__ p, const double* __restrict__ s, long long n,
                                  Scal* sc, double* part, unsigned int* ticket) {
    pdl_enter();
    const int half = sc->half;
    if (sc->iters == 0 && !half) return;
    const double al = sc->alpha, om = half ? 0.0 : sc->omega;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
        x[idx] += half ? al * p[idx] : al * p[idx] + om * s[idx];
    double one[1] = {0.0};
    const int op1[1] = {0};
    double t1[1];
    if (grid_reduce<1>(one, op1, part, ticket, t1) && threadIdx.x == 0 && half) {   // last block: every block has read sc->half
        sc->half = 0; sc->iters += 1; sc->iters_total += 1; sc->half_exits += 1;
        if (sc->iters > sc->iters_max) sc->iters_max = sc->iters;
    }
}

// ---------------------------------------------------------------------------------- Newton step pieces
// dmu = 2 (a dphi - kappa/2 L dphi + R_phi) and the step ceiling minima (Forward2_solver.py:377-391).
__global__ void dmu_ceiling_kernel(const double* __restrict__ dphi, const double* __restrict__ a,
                                   const double* __restrict__ Rphi, const double* __restrict__ phi,
                                   double* __restrict__ dmu, Geo g, Phys p, Scal* sc, double* part, unsigned int* ticket,
                                   const double* __restrict__ mu, double* __restrict__ phit, double* __restrict__ mut, double tau_dt) {
    pdl_enter();
    // phit/mut (optional, need phi and mu): the full-step trial iterate phi + dphi, mu + dmu — what trial_kernel(alpha = 1)
    // computes — so the speculative first Armijo trial costs no launch of its own
    // v[2]: squared norm of the nonlinear remainder of the FULL step.  The residual is linear in (phi, mu) except for the term
    // c1 l(phi), l = log((1+s)/(1-s)), so after a full Newton step with an exact linear solve the new residual IS
    // c1 [l(phi + dphi) - l(phi) - l'(phi) dphi] (first component) — computable without the Laplacians whose rounding noise
    // (eps / h^2) hides it in the measured residual on fine grids.  newton_step uses it for its stop test there.
    double v[3] = {INFINITY, INFINITY, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double d = dphi[idx];
        const double av = a[idx];
        const double dm = 2.0 * (av * d - 0.5 * p.kappa * lap_g(dphi, o, i, g) + Rphi[idx]);
        dmu[idx] = dm;
        if (phi) {
            const double f = phi[idx];
            if (phit) { phit[idx] = f + 1.0 * d; mut[idx] = mu[idx] + 1.0 * dm; }
            if (d > 0.0) v[0] = fmin(v[0], (p.lim - f) / d);
            else if (d < 0.0) v[1] = fmin(v[1], (-p.lim - f) / d);
            // a = tau/dt + c1 l'(phi)  =>  c1 l'(phi) dphi = (a - tau/dt) dphi
            const double rem = p.c1 * (flory_log(f + d, p.eps_log) - flory_log(f, p.eps_log)) - (av - tau_dt) * d;
            v[2] += rem * rem;
        }
    }
    const int op[3] = {1, 1, 0};
    double tot[3];
    if (grid_reduce<3>(v, op, part, ticket, tot) && threadIdx.x == 0) { sc->ceil_pos = tot[0]; sc->ceil_neg = tot[1]; sc->rem2 = tot[2]; }
}

__global__ void trial_kernel(const double* __restrict__ phi, const double* __restrict__ mu, const double* __restrict__ dphi,
                             const double* __restrict__ dmu, double* __restrict__ phit, double* __restrict__ mut,
                             long long n, double alpha) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        phit[idx] = phi[idx] + alpha * dphi[idx];
        mut[idx] = mu[idx] + alpha * dmu[idx];
    }
}

// Weighted mass of clip(phi) and the interior weight (Forward2_solver.py:562-571).  WRITE_CLIP stores the clipped field.
__global__ void clip_mass_kernel(const double* __restrict__ phin, double* __restrict__ phic, Geo g, Phys p, double hxhy,
                                 Scal* sc, int set_mass0, double* part, unsigned int* ticket) {
    pdl_enter();
    double v[2] = {0.0, 0.0};
    const double thr = p.lim - 5e-3;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(idx / g.ny1), iy = (int)(idx - (long long)ix * g.ny1);
        const int ixg = ix + g.o0;
        const double w = hxhy * ((ixg == 0 || ixg == g.nxg - 1) ? 0.5 : 1.0) * ((iy == 0 || iy == g.ny1 - 1) ? 0.5 : 1.0);
        double f = phin[idx];
        if (!set_mass0) { f = fmin(fmax(f, -p.lim), p.lim); phic[idx] = f; }
        v[0] += w * f;
        if (fabs(f) < thr) v[1] += w;
    }
    const int op[2] = {0, 0};
    double tot[2];
    if (grid_reduce<2>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        if (set_mass0) sc->mass0 = tot[0];
        else { sc->mass = tot[0]; sc->wint = tot[1]; }
    }
}

// Interior-only mass shift (uniform shift + re-clip when there is no interior), Forward2_solver.py:566-577.
__global__ void mass_shift_kernel(double* __restrict__ phi, Geo g, Phys p, double area, const Scal* __restrict__ sc) {
    pdl_enter();
    const double err = sc->mass - sc->mass0;
    if (!(fabs(err) > 1e-16)) return;
    const double wint = sc->wint, thr = p.lim - 5e-3;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        double f = phi[idx];
        if (wint > 0.0) { if (fabs(f) < thr) phi[idx] = f - err / wint; }
        else phi[idx] = fmin(fmax(f - err / area, -p.lim), p.lim);
    }
}

// ---------------------------------------------------------------------------------- adjoint sweep pieces
__device__ __forceinline__ double fpp_dev(double phi, double c1, double c2) {
    const double s = fmin(fmax(phi, -1.0 + 1e-8), 1.0 - 1e-8);
    return 2.0 * c1 / (1.0 - s * s) - 2.0 * c2;
}

__global__ void adj_terminal_rhs_kernel(const double* __restrict__ phiM, const double* __restrict__ phiT,
                                        double* __restrict__ b, long long n, double b2) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
        b[idx] = b2 * (phiM[idx] - (phiT ? phiT[idx] : 0.0));
}

// rhs = B(phi_{n+1}) p_{n+1} + src with L p_{n+1} = -q_{n+1} already known (backward2_solver.py:200-203, :222-226):
//   rhs = p1 + tau q1 + dt/2 L q1 - dt/2 f''(phi1) q1 + dt/2 b1 ((phi0 - Q0) + (phi1 - Q1))
// and the coefficient a = tau + dt/2 f''(phi0) of the implicit operator A(phi_n) with its min/max.
__global__ void adj_rhs_kernel(const double* __restrict__ p1, const double* __restrict__ q1,
                               const double* __restrict__ phi1, const double* __restrict__ phi0,
                               const double* __restrict__ Q1, const double* __restrict__ Q0,
                               double* __restrict__ rhs, double* __restrict__ a, Geo g, Phys p, double dt, double b1,
                               Scal* sc, double* part, unsigned int* ticket, double tol2) {
    pdl_enter();
    double v[2] = {INFINITY, -INFINITY};
    const double hdt = 0.5 * dt;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double f1 = phi1[idx], f0 = phi0[idx], qv = q1[idx];
        const double src = hdt * b1 * ((f0 - (Q0 ? Q0[idx] : 0.0)) + (f1 - (Q1 ? Q1[idx] : 0.0)));
        rhs[idx] = p1[idx] + p.tau * qv + hdt * lap_g(q1, o, i, g) - hdt * fpp_dev(f1, p.c1, p.c2) * qv + src;
        const double av = p.tau + hdt * fpp_dev(f0, p.c1, p.c2);
        a[idx] = av; v[0] = fmin(v[0], av); v[1] = fmax(v[1], av);
    }
    const int op[2] = {1, 2};
    double tot[2];
    if (grid_reduce<2>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->amin = tot[0]; sc->amax = tot[1];
        sc->abar = (tot[0] > 0.0) ? sqrt(tot[0] * tot[1]) : 0.5 * (tot[0] + tot[1]);
        sc->c0 = 1.0; sc->c2 = hdt; sc->tol2 = tol2; sc->adj = 1;              // coefficients / tolerance of the solve that follows
    }
}

// q0 = -L p0 ;  r0 = fb r1 + fs (q0 + q1)      (backward2_solver.py:233-242).  r1 == nullptr: terminal level (r = 0).
__global__ void adj_qr_kernel(const double* __restrict__ p0, const double* __restrict__ q1, const double* __restrict__ r1,
                              double* __restrict__ q0, double* __restrict__ r0, Geo g, double fb, double fs) {
    pdl_enter();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < g.n; idx += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(idx / g.ni), i = (int)(idx - (long long)o * g.ni);
        const double qv = -lap_g(p0, o, i, g);
        q0[idx] = qv;
        if (r0) r0[idx] = r1 ? fb * r1[idx] + fs * (qv + q1[idx]) : 0.0;
    }
}

// ---------------------------------------------------------------------------------- cost / prox / KKT reductions
// J sums over (levels, Nx+1, Ny+1) with np.trapz weights wt[t] wx[i] wy[j] (cost2_and_function.py:80-108).
__global__ void cost_kernel(const double* __restrict__ phi, const double* __restrict__ u, const double* __restrict__ Q,
                            const double* __restrict__ phiT, int levels, int nx1, int ny1,
                            const double* __restrict__ wt, const double* __restrict__ wx, const double* __restrict__ wy,
                            double* out4, double* part, unsigned int* ticket, int term_level = -2, int accumulate = 0) {
    pdl_enter();
    // term_level: index (within the levels given) of the terminal level carrying the J2 term; -2 = the last one (whole
    // trajectory in one launch), -1 = none (a chunk that does not contain the final time).  accumulate: add to out4.
    // one warp per array row (t, ix): the level / row indices and their weights are per-row scalars, the lanes stride along y —
    // no division and two weight loads less per element than the round-1 flat loop (2.4 TB/s)
    const long long n = (long long)nx1 * ny1, rows = (long long)levels * nx1;
    if (term_level == -2) term_level = levels - 1;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
        const int t = (int)(row / nx1), ix = (int)(row - (long long)t * nx1);
        const double wxv = wx[ix], wtv = wt[t];
        const bool term = (t == term_level);
        const double* __restrict__ pf = phi + row * ny1;
        const double* __restrict__ pq = Q ? Q + row * ny1 : nullptr;
        const double* __restrict__ pu = u ? u + row * ny1 : nullptr;
        const double* __restrict__ pt = (term && phiT) ? phiT + (long long)ix * ny1 : nullptr;
        for (int iy = lane; iy < ny1; iy += 32) {
            const double ws = wxv * wy[iy], w = wtv * ws;
            const double f = pf[iy];
            const double e = f - (pq ? pq[iy] : 0.0);
            v[0] += w * e * e;
            if (term) { const double d = f - (pt ? pt[iy] : 0.0); v[1] += ws * d * d; }
            if (pu) { const double uv = pu[iy]; v[2] += w * uv * uv; v[3] += w * fabs(uv); }
        }
    }
    (void)n;
    const int op[4] = {0, 0, 0, 0};
    double tot[4];
    if (grid_reduce<4>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        if (accumulate) { out4[0] += tot[0]; out4[1] += tot[1]; out4[2] += tot[2]; out4[3] += tot[3]; }
        else { out4[0] = tot[0]; out4[1] = tot[1]; out4[2] = tot[2]; out4[3] = tot[3]; }
    }
}

// g = r + b3 u ; v = u - alpha g ; soft threshold alpha*kappa ; box.  Also ||u_new-u||^2, ||u||^2, support and bound counts.
__global__ void grad_prox_kernel(const double* __restrict__ u, const double* __restrict__ r, double* __restrict__ grad,
                                 double* __restrict__ un, long long n, double b3, double alpha, double ksp, double umin,
                                 double umax, double* out4, double* part, unsigned int* ticket, int accumulate = 0) {
    pdl_enter();
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const double thr = alpha * ksp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double uv = u[idx];
        // explicit round-to-nearest mul/add (no FMA) so the support pattern matches NumPy's evaluation order bit for bit
        const double gv = __dadd_rn(r[idx], __dmul_rn(b3, uv));
        if (grad) grad[idx] = gv;
        const double y = __dsub_rn(uv, __dmul_rn(alpha, gv));
        const double m = fmax(__dsub_rn(fabs(y), thr), 0.0);
        double s = (y > 0.0) ? m : ((y < 0.0) ? -m : 0.0);
        s = fmin(fmax(s, umin), umax);
        un[idx] = s;
        const double d = s - uv;
        v[0] += d * d; v[1] += uv * uv;
        v[2] += (s != 0.0) ? 1.0 : 0.0;
        v[3] += (s == umin || s == umax) ? 1.0 : 0.0;
    }
    const int op[4] = {0, 0, 0, 0};
    double tot[4];
    if (grid_reduce<4>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        if (accumulate) { out4[0] += tot[0]; out4[1] += tot[1]; out4[2] += tot[2]; out4[3] += tot[3]; }   // chunked launches
        else { out4[0] = tot[0]; out4[1] = tot[1]; out4[2] = tot[2]; out4[3] = tot[3]; }
    }
}

// Discrete free energy (Forward2_solver.py:256-319): forward-difference gradient sums along both axes, trapezoid-weighted
// bulk term c1[(1+s)ln(1+s) + (1-s)ln(1-s)] - c2 s^2 with s = clip(phi, +-(1-eps)), optional coupling sum w phi.
// phi is (n0, n1) row-major.  out4 = { sum dphi_x^2, sum dphi_y^2, sum wts psi, sum wts w phi }.
__global__ void energy_kernel(const double* __restrict__ phi, const double* __restrict__ w, int n0, int n1, double c1, double c2,
                              double eps, double* out4, double* part, unsigned int* ticket) {
    pdl_enter();
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const long long n = (long long)n0 * n1;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / n1), c = (int)(idx - (long long)r * n1);
        const double f = phi[idx];
        if (c + 1 < n1) { const double d = phi[idx + 1] - f; v[0] += d * d; }
        if (r + 1 < n0) { const double d = phi[idx + n1] - f; v[1] += d * d; }
        const double wt = ((r == 0 || r == n0 - 1) ? 0.5 : 1.0) * ((c == 0 || c == n1 - 1) ? 0.5 : 1.0);
        const double s = fmin(fmax(f, -1.0 + eps), 1.0 - eps);
        v[2] += wt * (c1 * ((1.0 + s) * log(1.0 + s) + (1.0 - s) * log(1.0 - s)) - c2 * s * s);
        if (w) v[3] += wt * w[idx] * f;
    }
    const int op[4] = {0, 0, 0, 0};
    double tot[4];
    if (grid_reduce<4>(v, op, part, ticket, tot) && threadIdx.x == 0) { out4[0] = tot[0]; out4[1] = tot[1]; out4[2] = tot[2]; out4[3] = tot[3]; }
}

__global__ void kkt_kernel(const double* __restrict__ u, const double* __restrict__ r, long long n, double ksp, double tol,
                           double* out3, double* part, unsigned int* ticket) {
    pdl_enter();
    double v[3] = {0.0, 0.0, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const bool a = fabs(u[idx]) < tol, b = fabs(r[idx]) <= ksp;
        v[0] += a ? 1.0 : 0.0; v[1] += b ? 1.0 : 0.0; v[2] += (a == b) ? 1.0 : 0.0;
    }
    const int op[3] = {0, 0, 0};
    double tot[3];
    if (grid_reduce<3>(v, op, part, ticket, tot) && threadIdx.x == 0) { out3[0] = tot[0]; out3[1] = tot[1]; out3[2] = tot[2]; }
}

}  // namespace vch
