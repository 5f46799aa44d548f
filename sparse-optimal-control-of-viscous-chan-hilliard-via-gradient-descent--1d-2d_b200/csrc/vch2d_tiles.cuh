// Tiled 2D stencil kernels (round 2): the Newton residual, the Newton update (dmu + step ceilings), the per-step constants and the
// adjoint right-hand side / recurrences on shared-memory halo tiles, with the neighbouring vector kernels fused in.
//
// Why: the round-1 stencil kernels (vch2d_kernels.cuh: 1-D grid-stride loop, a 64-bit division per node, five dependent scalar
// global loads per stencilled field) ran at 0.27-0.45 of the copy roofline, and the forward Newton iteration spent 5 launches
// (~100 us at 1024^2) outside the Krylov loop.  Here
//   * one CTA owns a TH x TW tile (<= 32 x 64 nodes, balanced so that no sliver tiles exist: 17 x 33 = 561 CTAs at 1025^2, one wave
//     at 4 CTAs / SM); row / column come from the tile origin — no division per node;
//   * every stencilled field is staged ONCE per tile (+ 1 halo, Neumann mirror applied in the load) with coalesced loads that are
//     all in flight together, and the 5-point stencil reads shared memory (conflict-free: a warp reads 32 consecutive doubles);
//   * each thread owns 8 nodes of one column (rows ty + 4k), so its global loads / stores are coalesced 512-byte row segments;
//   * fused neighbours: solve_w into the step set-up; the solver coefficients into the residual; the closing BiCGStab update
//     x += alpha p + omega s into the dmu kernel (the update is applied while the tile of dphi is staged); the BiCGStab start
//     (r = r0 = b, x = 0, ||b||^2) into the adjoint right-hand side; the copy of the solution into the adjoint q/r recurrence.
// The arithmetic per node is the round-1 kernels' (same expressions, same operation order inside the Laplacian), so results agree
// to rounding of the reductions.  Slab mode keeps the round-1 kernels (ghost rows live in the arena margins, not in a mirror rule).
#pragma once
#include "vch2d_kernels.cuh"

namespace vch {

#ifndef VCH_CPU_EMU
constexpr int kTW = 64, kTH = 32, kTileThreads = 256;
#else
constexpr int kTW = 16, kTH = 16, kTileThreads = 256;   // tests/emu: small tiles, so that the small validation grids span several tiles
#endif
constexpr int kTR = kTileThreads / kTW;             // tile rows covered by one pass of the CTA (4)
constexpr int kTP = kTH / kTR;                      // nodes per thread (8): rows ty + kTR k of column tx
constexpr int kHW = kTW + 2, kHH = kTH + 2;         // halo tile in shared memory

struct Tiling {
    int ntx = 1, nty = 1;
    int grid() const { return ntx * nty; }
};
static inline Tiling make_tiling(const Geo& g) { return Tiling{(g.ni + kTW - 1) / kTW, (g.no + kTH - 1) / kTH}; }

struct Tile { int r0, c0, th, tw; };
__device__ __forceinline__ Tile tile_of(const Geo& g, const Tiling& tl) {
    const int tj = (int)blockIdx.x / tl.ntx, ti = (int)blockIdx.x - tj * tl.ntx;
    Tile t;
    t.r0 = (int)((long long)tj * g.no / tl.nty); t.th = (int)((long long)(tj + 1) * g.no / tl.nty) - t.r0;
    t.c0 = (int)((long long)ti * g.ni / tl.ntx); t.tw = (int)((long long)(ti + 1) * g.ni / tl.ntx) - t.c0;
    return t;
}

// sm[(r + 1) * kHW + (c + 1)] = f(node index of tile node (r, c)), r in [-1, th], c in [-1, tw]; nodes outside the grid are the
// mirror images (Neumann ghost rule of lap_g: v[-1] = v[1], v[n] = v[n-2]).  Thread (tx, ty) stages column tx of the rows
// ty + kTR k - 1 (k = 0 .. kFillK-1: the column index and its mirror are computed once, the row offset advances by a constant),
// the first 2 (th + 2) threads add the two halo columns.  Trip counts are compile-time so that all loads are issued before the
// first shared-memory store — the kernels are instruction-issue bound (ncu: 64 % issue-active at 46 % warps active), so the
// address arithmetic per staged element matters as much as the loads themselves.
constexpr int kFillK = (kHH + kTR - 1) / kTR;
template <class F>
__device__ __forceinline__ void tile_fill(double* __restrict__ sm, const Tile& t, const Geo& g, F f) {
    const int tx = (int)threadIdx.x & (kTW - 1), ty = (int)threadIdx.x / kTW;
    const int rows = t.th + 2;
    const bool col_ok = tx < t.tw;
    const size_t gc = (size_t)(t.c0 + tx);                      // interior columns need no mirror
    double v[kFillK];
#pragma unroll
    for (int k = 0; k < kFillK; ++k) {
        const int r = ty + kTR * k;
        if (col_ok && r < rows) v[k] = f((size_t)mirror(t.r0 - 1 + r, g.no) * g.ni + gc);
    }
    // halo columns: thread h < 2 rows -> row h >> 1, left (h even) or right (h odd) halo column
    const int h = (int)threadIdx.x;
    double vh = 0.0;
    const bool halo = h < 2 * rows;
    if (halo) vh = f((size_t)mirror(t.r0 - 1 + (h >> 1), g.no) * g.ni + mirror((h & 1) ? t.c0 + t.tw : t.c0 - 1, g.ni));
#pragma unroll
    for (int k = 0; k < kFillK; ++k) {
        const int r = ty + kTR * k;
        if (col_ok && r < rows) sm[r * kHW + tx + 1] = v[k];
    }
    if (halo) sm[(h >> 1) * kHW + ((h & 1) ? t.tw + 1 : 0)] = vh;
}

// lap_g on the staged tile: same operand order as lap_g (inner axis first: (east - c) + (west - c); then (row+1 - c) + (row-1 - c))
__device__ __forceinline__ double lap_s(const double* __restrict__ sm, int ci, const Geo& g) {
    const double c = sm[ci];
    const double a = (sm[ci + 1] - c) + (sm[ci - 1] - c);
    const double b = (sm[ci + kHW] - c) + (sm[ci - kHW] - c);
    return a * g.ihi2 + b * g.iho2;
}

// for k in 0..kTP-1: tile row r = ty + kTR k, column tx; body(k, node index, shared-memory centre index) for nodes inside the tile
#define VCH_TILE_NODES(t, g)                                                                                   \
    const int tx = (int)threadIdx.x & (kTW - 1), ty = (int)threadIdx.x / kTW;                                   \
    const bool col_ok = tx < (t).tw;                                                                           \
    const size_t node0 = (size_t)((t).r0 + ty) * (g).ni + (t).c0 + tx;                                         \
    const int ci0 = (ty + 1) * kHW + tx + 1

// ---------------------------------------------------------------------------------- per-step constants (+ solve_w)
// w1 = solve_w(w0, u_n, u_{n+1}) when gdt > 0 (solve_w_kernel's expression), then step_setup_kernel's cphi / cmu / mu guess.
__global__ void __launch_bounds__(kTileThreads, 4)
step_setup_tile_kernel(const double* __restrict__ phi0, const double* __restrict__ mu0, const double* __restrict__ w0,
                       double* __restrict__ w1, const double* __restrict__ un, const double* __restrict__ un1, double gdt,
                       double* __restrict__ cphi, double* __restrict__ cmu, double* __restrict__ mu_guess, Geo g, Tiling tl, Phys p, double dt) {
    pdl_enter();
    __shared__ double sf[kHH * kHW], sm_[kHH * kHW];
    const Tile t = tile_of(g, tl);
    tile_fill(sf, t, g, [&](size_t i) { return phi0[i]; });
    tile_fill(sm_, t, g, [&](size_t i) { return mu0[i]; });
    __syncthreads();
    VCH_TILE_NODES(t, g);
    const double idt = 1.0 / dt;
    if (!col_ok) return;
    double wa[kTP], wb[kTP];
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int r = ty + kTR * k;
        if (r < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            wa[k] = w0[idx];
            if (gdt > 0.0) {
                const double a = un ? un[idx] : 0.0, b = un1 ? un1[idx] : 0.0;
                wb[k] = __ddiv_rn(__dadd_rn(__dmul_rn(gdt - 0.5, wa[k]), __dmul_rn(0.5, __dadd_rn(b, a))), gdt + 0.5);
            } else wb[k] = w1[idx];
        }
    }
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int r = ty + kTR * k;
        if (r < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            const int ci = ci0 + kTR * k * kHW;
            const double f = sf[ci], m = sm_[ci];
            const double lf = lap_s(sf, ci, g), lm = lap_s(sm_, ci, g);
            if (gdt > 0.0) w1[idx] = wb[k];
            cphi[idx] = -p.tau * f * idt - 0.5 * p.kappa * lf - 2.0 * p.c2 * f - 0.5 * m - 0.5 * (wb[k] + wa[k]);
            cmu[idx] = -f * idt - 0.5 * lm;
            if (mu_guess) mu_guess[idx] = -p.kappa * lf + p.c1 * flory_log(f, p.eps_log) - 2.0 * p.c2 * f - wb[k];
        }
    }
}

// ---------------------------------------------------------------------------------- Newton residual
// residual_kernel on tiles.  tol2 > 0: also the coefficients / tolerance of the linear solve that follows (schur_rhs_kernel set
// them; the Schur right-hand side itself is formed by the first row transform of the solve, RowPrologue mode 4).
__global__ void __launch_bounds__(kTileThreads, 4)
residual_tile_kernel(const double* __restrict__ phi, const double* __restrict__ mu, const double* __restrict__ cphi,
                     const double* __restrict__ cmu, double* __restrict__ Rphi, double* __restrict__ Rmu, double* __restrict__ a,
                     Geo g, Tiling tl, Phys p, double dt, Scal* sc, double* part, unsigned int* ticket, Scal* publish,
                     double c0, double c2, double tol2) {
    pdl_enter();
    __shared__ double sf[kHH * kHW], sm_[kHH * kHW];
    const Tile t = tile_of(g, tl);
    tile_fill(sf, t, g, [&](size_t i) { return phi[i]; });
    tile_fill(sm_, t, g, [&](size_t i) { return mu[i]; });
    VCH_TILE_NODES(t, g);
    __syncthreads();
    const double idt = 1.0 / dt, tdt = p.tau / dt;
    double v[4] = {0.0, INFINITY, -INFINITY, 0.0};
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int r = ty + kTR * k;
        if (col_ok && r < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            const int ci = ci0 + kTR * k * kHW;
            const double f = sf[ci], m = sm_[ci];
            const double rp = tdt * f - 0.5 * p.kappa * lap_s(sf, ci, g) + p.c1 * flory_log(f, p.eps_log) - 0.5 * m + cphi[idx];
            const double rm = f * idt - 0.5 * lap_s(sm_, ci, g) + cmu[idx];
            const double d = tdt + 2.0 * p.c1 / (1.0 - fmin(f * f, p.dsq));
            Rphi[idx] = rp; Rmu[idx] = rm;
            if (a) a[idx] = d;
            v[0] += rp * rp + rm * rm;
            v[1] = fmin(v[1], d); v[2] = fmax(v[2], d);
            v[3] += m * m;
        }
    }
    const int op[4] = {0, 1, 2, 0};
    double tot[4];
    if (grid_reduce<4>(v, op, part, ticket, tot)) {
        if (threadIdx.x == 0) {
            sc->res2 = tot[0]; sc->amin = tot[1]; sc->amax = tot[2]; sc->abar = sqrt(tot[1] * tot[2]); sc->mu2 = tot[3];
            if (!isfinite(tot[0])) sc->nonfinite = 1;
            if (tol2 > 0.0) { sc->c0 = c0; sc->c2 = c2; sc->tol2 = tol2; sc->adj = 0; }
        }
        if (publish) publish_scalars(sc, publish);
    }
}

// ---------------------------------------------------------------------------------- Newton update
// bicg_close_kernel + dmu_ceiling_kernel: dphi = x + alpha p + omega s (x + alpha p after a half-step exit, x itself when the
// solve ended without an iteration) is formed while its tile is staged and written to xout — NOT back to x: neighbouring tiles
// read x in their halos, an in-place update would race with them; then dmu, the trial iterate, the step-ceiling minima and the
// nonlinear remainder exactly as dmu_ceiling_kernel.  close = 0: dphi = x as it is, nothing is written to xout.
__global__ void __launch_bounds__(kTileThreads, 4)
dmu_close_tile_kernel(const double* __restrict__ x, double* __restrict__ xout, const double* __restrict__ pk, const double* __restrict__ sk,
                      const double* __restrict__ a, const double* __restrict__ Rphi, const double* __restrict__ phi,
                      double* __restrict__ dmu, Geo g, Tiling tl, Phys p, Scal* sc, double* part, unsigned int* ticket,
                      const double* __restrict__ mu, double* __restrict__ phit, double* __restrict__ mut, double tau_dt, int close) {
    pdl_enter();
    __shared__ double sd[kHH * kHW];
    const Tile t = tile_of(g, tl);
    const int half = close ? sc->half : 0, its = close ? sc->iters : 0;
    const double al = close ? sc->alpha : 0.0, om = close ? sc->omega : 0.0;
    const int mode = !close ? 0 : (half ? 1 : (its > 0 ? 2 : 0));        // 0: dphi = x;  1: x + alpha p;  2: x + alpha p + omega s
    if (mode == 2) tile_fill(sd, t, g, [&](size_t i) { return x[i] + (al * pk[i] + om * sk[i]); });
    else if (mode == 1) tile_fill(sd, t, g, [&](size_t i) { return x[i] + al * pk[i]; });
    else tile_fill(sd, t, g, [&](size_t i) { return x[i]; });
    VCH_TILE_NODES(t, g);
    __syncthreads();
    double v[3] = {INFINITY, INFINITY, 0.0};
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int r = ty + kTR * k;
        if (col_ok && r < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            const int ci = ci0 + kTR * k * kHW;
            const double d = sd[ci], av = a[idx];
            if (close) xout[idx] = d;
            const double dm = 2.0 * (av * d - 0.5 * p.kappa * lap_s(sd, ci, g) + Rphi[idx]);
            dmu[idx] = dm;
            if (phi) {
                const double f = phi[idx];
                if (phit) { phit[idx] = f + 1.0 * d; mut[idx] = mu[idx] + 1.0 * dm; }
                if (d > 0.0) v[0] = fmin(v[0], (p.lim - f) / d);
                else if (d < 0.0) v[1] = fmin(v[1], (-p.lim - f) / d);
                const double rem = p.c1 * (flory_log(f + d, p.eps_log) - flory_log(f, p.eps_log)) - (av - tau_dt) * d;
                v[2] += rem * rem;
            }
        }
    }
    const int op[3] = {1, 1, 0};
    double tot[3];
    if (grid_reduce<3>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->ceil_pos = tot[0]; sc->ceil_neg = tot[1]; sc->rem2 = tot[2];
        if (half) {   // what bicg_close_kernel's last block does (every block has read sc->half / iters before its ticket)
            sc->half = 0; sc->iters += 1; sc->iters_total += 1; sc->half_exits += 1;
            if (sc->iters > sc->iters_max) sc->iters_max = sc->iters;
        }
    }
}

// ---------------------------------------------------------------------------------- adjoint sweep
// adj_rhs_kernel on tiles; init = 1: the start of the right-preconditioned BiCGStab solve fused in — r = r0 = rhs, x = 0, ||rhs||^2 and
// the scalars bicg_init_kernel sets (the solve graph then starts with its WHILE node).
__global__ void __launch_bounds__(kTileThreads, 4)
adj_rhs_tile_kernel(const double* __restrict__ p1, const double* __restrict__ q1, const double* __restrict__ phi1,
                    const double* __restrict__ phi0, const double* __restrict__ Q1, const double* __restrict__ Q0,
                    double* __restrict__ r, double* __restrict__ r0, double* __restrict__ x, double* __restrict__ a, Geo g, Tiling tl,
                    Phys p, double dt, double b1, Scal* sc, double* part, unsigned int* ticket, double tol2, int init) {
    pdl_enter();
    __shared__ double sq[kHH * kHW];
    const Tile t = tile_of(g, tl);
    tile_fill(sq, t, g, [&](size_t i) { return q1[i]; });
    VCH_TILE_NODES(t, g);
    const double hdt = 0.5 * dt;
    __syncthreads();
    double v[3] = {INFINITY, -INFINITY, 0.0};
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int rr_ = ty + kTR * k;
        if (col_ok && rr_ < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            const int ci = ci0 + kTR * k * kHW;
            const double qv = sq[ci], f1 = phi1[idx], f0 = phi0[idx];
            const double src = hdt * b1 * ((f0 - (Q0 ? Q0[idx] : 0.0)) + (f1 - (Q1 ? Q1[idx] : 0.0)));
            const double rhs = p1[idx] + p.tau * qv + hdt * lap_s(sq, ci, g) - hdt * fpp_dev(f1, p.c1, p.c2) * qv + src;
            const double av = p.tau + hdt * fpp_dev(f0, p.c1, p.c2);
            r[idx] = rhs; a[idx] = av;
            if (init) { r0[idx] = rhs; x[idx] = 0.0; }
            v[0] = fmin(v[0], av); v[1] = fmax(v[1], av); v[2] += rhs * rhs;
        }
    }
    const int op[3] = {1, 2, 0};
    double tot[3];
    if (grid_reduce<3>(v, op, part, ticket, tot) && threadIdx.x == 0) {
        sc->amin = tot[0]; sc->amax = tot[1];
        sc->abar = (tot[0] > 0.0) ? sqrt(tot[0] * tot[1]) : 0.5 * (tot[0] + tot[1]);
        sc->c0 = 1.0; sc->c2 = hdt; sc->tol2 = tol2; sc->adj = 1;
        if (init) solve_init_scalars(sc, tot[2], (cudaGraphConditionalHandle)0, 0);
    }
}

// adj_qr_kernel on tiles; pin is the solution of the level's solve (the Krylov iterate), copied to p0 on the way (p0 == pin: no copy).
__global__ void __launch_bounds__(kTileThreads, 4)
adj_qr_tile_kernel(const double* pin, double* p0, const double* __restrict__ q1, const double* __restrict__ r1,
                   double* __restrict__ q0, double* __restrict__ r0, Geo g, Tiling tl, double fb, double fs) {
    pdl_enter();
    __shared__ double sp[kHH * kHW];
    const Tile t = tile_of(g, tl);
    tile_fill(sp, t, g, [&](size_t i) { return pin[i]; });
    VCH_TILE_NODES(t, g);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTP; ++k) {
        const int r = ty + kTR * k;
        if (col_ok && r < t.th) {
            const size_t idx = node0 + (size_t)(kTR * k) * g.ni;
            const int ci = ci0 + kTR * k * kHW;
            const double qv = -lap_s(sp, ci, g);
            if (p0 != pin) p0[idx] = sp[ci];
            q0[idx] = qv;
            if (r0) r0[idx] = r1 ? fb * r1[idx] + fs * (qv + q1[idx]) : 0.0;
        }
    }
}

}  // namespace vch
