// DCT-I fast solve: out = V f(Lambda)^-1 V^-1 in, where V diagonalises the mirror-ghost Neumann Laplacian
// (eigenvectors cos(pi j k / N), eigenvalues -(4/h^2) sin^2(pi k / 2N)) and
// f(lambda) = c0 + abar*lambda + c2*lambda^2 is the constant-coefficient symbol of the Schur / adjoint operator.
//
// A line of N+1 reals is transformed through the FFT of its even extension (length Lf = 2N).  Two lines share one
// complex FFT (line a -> real part, line b -> imaginary part): the transform of a real-even sequence is real, so
// Re/Im of the result are the two DCT-I outputs and no post-twiddle pass is needed.
//
// The FFT is a Stockham autosort transform, radix 8 with one radix-8/4/2 last pass, 8 points per thread in registers:
//   * the FIRST pass reads its 8 points straight from global memory (the even extension is formed by index
//     reflection in the load), so the input is never staged;
//   * middle passes go through a padded shared-memory array (index i -> i + i/8: the stride-8 stores of pass 1
//     become conflict-free for 16-byte accesses);
//   * the LAST pass hands its natural-order outputs to an epilogue: global store (rows), symbol division + mirrored
//     store for the inverse transform (fused column solve), optionally dot products for BiCGStab;
//   * twiddles w^r are built from 3 table loads (w, w^2, w^4) and 4 complex multiplies.
// Rows (contiguous lines) and columns (stride-ni lines, 2*ppb adjacent columns per CTA so every 32-byte sector that
// is fetched is fully used) share one kernel template.  N must be a power of two >= 32 for this path; other N use the
// dense-table kernels at the bottom (small validation grids).
#pragma once
#include "vch_common.cuh"
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <algorithm>

namespace vch {

struct DctAxis {
    int n = 0;            // nodes = N+1
    bool fft = false;     // N is a power of two, 32 <= N <= 4096
    int Lf = 0, log2L = 0;
    double2* tw = nullptr;    // Lf twiddles exp(-2 pi i m / Lf)
    double2* tw16 = nullptr;  // per-pass twiddle tables of the radix-16 kernels (vch_fft16.cuh)
    double* denseT = nullptr; // n*n, denseT[j*n + k] = 2 c_j cos(pi j k / N)   (transposed for coalescing)
    double* lam = nullptr;    // n eigenvalues of -L1d (>= 0)
};

struct SymbolArgs {
    double c0, c2;
    const double* abar_ptr;   // device scalar (may be null -> abar_const)
    double abar_const;
    const double* coef_ptr;   // device {c0, c2} (may be null -> the immediate values above)
};

// Optional epilogue of the last row transform: add a vector (out = z + addend) and BiCGStab dot products over `out`.
struct DotEpilogue {
    int mode = 0;                 // 0 none, 1: (other, out) -> alpha = rho_new / dot, rho = rho_new ; 2: (out, other), (out, out) -> omega
                                  // 4: mode 2 plus (s,s), (r0,s), (r0,t) -> omega, rho_new, (r,r), stop decision (6-launch iteration)
    const double* other = nullptr;
    Scal* sc = nullptr;
    double* part = nullptr;
    unsigned int* ticket = nullptr;
    const double* addend = nullptr;
    // right-preconditioned form (adjoint operator A P^-1 = I - diag(a - abar) L P^-1): out = addend + (mul_a - abar) * z
    const double* mul_a = nullptr;
    // mode 1 only: the current residual r.  With (r, out) the kernel also knows ||s||^2 of s = r - alpha*out and can end the
    // solve at the half step (dots_finish)
    const double* rvec = nullptr;
    // 6-launch iteration (measured on B200: -6.6 % per PGD iteration against the 7-launch form).  mode 4: second epilogue that
    // also takes (s,s), (r0,s), (r0,t) — rvec carries r0 — and derives rho_new = (r0,s) - omega (r0,t), (r,r) = (s,s) - 2 omega (t,s)
    // + omega^2 (t,t), the stop decision and the loop conditional, which the x/r update kernel produces in the 7-launch form.
    cudaGraphConditionalHandle cond = 0;
    int use_cond = 0;
    // mode 5 (start of a forward solve fused into the last row transform of P^-1 b): out = r, out2 = r0 (same values),
    // zero = x (cleared), (r,r) -> the scalars bicg_init_kernel sets.  Radix-16 kernels only.
    double* out2 = nullptr;
    double* zero = nullptr;
};

// Start-of-solve scalars from ||r||^2 (what bicg_init_kernel's last block does).
__device__ __forceinline__ void solve_init_scalars(Scal* sc, double rr, cudaGraphConditionalHandle cond, int use_cond) {
    sc->bnorm2 = rr; sc->rr = rr; sc->rho_new = rr;
    sc->rho = 1.0; sc->alpha = 0.0; sc->omega = 1.0;
    sc->thr2 = sc->tol2 * rr;
    sc->iters = 0; sc->half = 0;
    sc->done = (rr == 0.0 || !isfinite(rr)) ? 1 : 0;
    if (!isfinite(rr)) sc->nonfinite = 1;
    sc->solves += 1;
    if (use_cond) { sc->g_launches += use_cond; cudaGraphSetConditional(cond, sc->done ? 0u : 1u); }
}

// Last block of a dot-product epilogue: BiCGStab scalars from the grid-wide sums tot = {(other,out), (out,out), (r,out)}.
// Half-step exit (mode 1): ||s||^2 = (r,r) - 2 alpha (r,v) + alpha^2 (v,v) for s = r - alpha v.  The three-term formula
// carries a rounding error of ~1e-13 (r,r), so it is only trusted once (r,r) is within 1e6 of the threshold (a half step
// gains 1e-2..1e-4 on ||.||^2); otherwise the iteration simply continues as before.  When it fires, done = half = 1: the
// remaining transform kernels of the iteration return at once and bicg_x_kernel applies x += alpha p.
__device__ __forceinline__ void dots_finish(const DotEpilogue& epi, const double (&tot)[3]) {
    Scal* sc = epi.sc;
    if (epi.mode == 1) {
        const double al = sc->rho_new / tot[0];
        sc->r0v = tot[0]; sc->alpha = al; sc->rho = sc->rho_new;
        if (epi.rvec) {
            const double ss = sc->rr - 2.0 * al * tot[2] + al * al * tot[1];
            if (isfinite(ss) && ss <= sc->thr2 && sc->rr <= 1e6 * sc->thr2) {
                sc->half = 1; sc->done = 1;
                // 6-launch iteration: the rest of the loop body returns on `done`, nobody else would clear the loop conditional
                if (epi.use_cond) { sc->g_launches += epi.use_cond; cudaGraphSetConditional(epi.cond, 0u); }
            }
        }
    } else {
        sc->ts = tot[0]; sc->tt = tot[1]; sc->omega = (tot[1] > 0.0) ? tot[0] / tot[1] : 0.0;
    }
}
// tot = {(s,t), (t,t), (r0,t), (s,s), (r0,s)}: everything bicg_x_kernel's reduction delivered, without forming r.
// The three-term (r,r) is trusted for the stop test only once (s,s) is within 1e6 of the threshold (see dots_finish).
__device__ __forceinline__ void dots_finish6(const DotEpilogue& epi, const double (&tot)[5]) {
    Scal* sc = epi.sc;
    const double ts = tot[0], tt = tot[1], r0t = tot[2], ss = tot[3], r0s = tot[4];
    const double om = (tt > 0.0) ? ts / tt : 0.0;
    double rr = ss - 2.0 * om * ts + om * om * tt;
    if (rr < 0.0) rr = 0.0;
    const double rho_new = r0s - om * r0t;
    sc->ts = ts; sc->tt = tt; sc->omega = om; sc->rr = rr; sc->rho_new = rho_new;
    sc->iters += 1; sc->iters_total += 1;
    if (sc->iters > sc->iters_max) sc->iters_max = sc->iters;
    const bool bad = !isfinite(rr) || !isfinite(rho_new);
    if (bad) sc->nonfinite = 1;
    if (bad || (rr <= sc->thr2 && ss <= 1e6 * sc->thr2)) sc->done = 1;
    if (epi.use_cond) {
        sc->g_launches += epi.use_cond;
        const bool stop = sc->done || sc->iters >= sc->maxit;
        if (stop && !sc->done) { sc->stalls += 1; if (sc->adj) sc->stalls_adj += 1; }
        cudaGraphSetConditional(epi.cond, stop ? 0u : 1u);
    }
}

// Optional prologue of the first row transform (fused BiCGStab vector update + coefficient multiply):
//   mode 0: x = in
//   mode 1: p = r + beta*q   (written to w), x = (a - abar) * p        beta = (rho_new/rho)(alpha/omega)
//   mode 2: s = r - alpha*v  (written to w), x = (a - abar) * s
//   mode 3 (6-launch iteration): the deferred update of the previous iteration, then p:  r = s - omega t (written to rw),
//           x += alpha p + omega s,  p = r + beta (p - omega v) (in place in w),  x_in = (a - abar) p;  first iteration: p = r
struct RowPrologue {
    int mode = 0;
    const double* r = nullptr;
    const double* qv = nullptr;
    const double* a = nullptr;      // nullptr: x = w (no coefficient multiply; right-preconditioned form)
    double* w = nullptr;
    const Scal* sc = nullptr;
    const double* s = nullptr;      // mode 3: s, t of the previous iteration, its v (in qv), the iterate x and the residual buffer
    const double* t = nullptr;
    double* x = nullptr;
    double* rw = nullptr;
    // mode 4 (start of a forward Newton solve, radix-16 kernels only): x_in = L r - qv with the 5-point Neumann Laplacian
    // (r = R_phi, qv = R_mu: the Schur right-hand side, schur_rhs_kernel) evaluated on the fly; k_in / k_out = 1/h^2 along / across the lines
    double k_in = 0.0, k_out = 0.0;
};

// Slab mode: the transposes between the row and the column transforms are done by the kernels' own stores, straight
// into the peers' buffers over NVLink (no all-to-all pass).
//   mode 1 (forward row kernel, STORE): element kk of local line l -> peer r = min(kk >> shift, nr-1):
//           T1_r[(base + l) * pitch + kk - (r << shift)]      (contiguous chunks of N/nr doubles: efficient remote writes)
//   mode 3 (inverse row kernel, LOAD):  the same addresses are read back after the owners' in-place column solves
//           (contiguous remote reads); the kernel's own output is an ordinary local store.
struct Scatter {
    int mode = 0, shift = 0, nr = 1, base = 0, pitch = 0;
    size_t off = 0;                       // arena offset (doubles) of the destination buffer, identical on every rank
    double* peer[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct DctSlab {
    bool on = false;
    Comm cm;
    int nloc = 0, o0 = 0;        // owned rows [o0, o0 + nloc)
    int wloc = 0, col0 = 0;      // owned columns of the transposed layout [col0, col0 + wloc)
    int shift = 0;               // log2(N / nranks): rows per rank = columns per rank (the last rank has one more)
    int p1 = 0;                  // pitch of T1 (all global rows x owned columns)
    double* T1 = nullptr;
};

struct DctPlan {
    DctAxis inner, outer;     // inner = contiguous axis (length ni), outer = strided axis (length no)
    int ni = 0, no = 0, pitch = 0;   // pitch of tmp1 (even, so column pairs are 16-byte aligned)
    DevBuf tmp1, tmp2;
    LaunchLog* log = nullptr;
    bool pdl = false;         // launch the transform kernels with programmatic stream serialization (see launch_pdl)
    bool lean = true;         // radix-16 kernels of vch_fft16.cuh where they apply (VCH_FFT16=0: the round-1 kernels everywhere)
    DctSlab slab;             // slab mode: no = global rows; the plan transforms the owned rows / columns only
    CUtensorMap tmap;         // tensor map of tmp1 (slab mode: T1) for the TMA column kernel (vch_fft16.cuh)
    bool have_tmap = false;
    void init(int no_, int ni_, double h_outer, double h_inner, LaunchLog* launch_log);
    void init_slab(int n_global, double h_outer, double h_inner, LaunchLog* launch_log, const DctSlab& sl);
    void barrier(cudaStream_t s, const int* done);
    void destroy();
    int max_grid() const;
    // out = P^-1 in   (in may equal out); epi = optional fused dots on `out`
    // scale_mode 0: divide by the symbol (P^-1);  1: multiply by lambda/symbol (-P^-1 L, the identity-plus-correction
    // form of P^-1 A for the forward Schur operator: P^-1 A x = x + DCT^-1[(lambda/sym) DCT((a - abar) x)])
    void apply(cudaStream_t s, const double* in, double* out, const SymbolArgs& sym, const int* done_flag,
               const DotEpilogue& epi = DotEpilogue(), const RowPrologue& pro = RowPrologue(), int scale_mode = 0);
};

// ------------------------------------------------------------------------------------------------ device side
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ int padi(int i) { return i + (i >> 3); }

template <int R> __device__ __forceinline__ void dft(double2 (&v)[R]);
template <> __device__ __forceinline__ void dft<2>(double2 (&v)[2]) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void dft<4>(double2 (&v)[4]) {
    double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    double2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2); v[1] = cadd(t1, t3); v[2] = csub(t0, t2); v[3] = csub(t1, t3);
}
template <> __device__ __forceinline__ void dft<8>(double2 (&v)[8]) {
    double2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    dft<4>(e); dft<4>(o);
    const double h = 0.70710678118654752440;
    double2 o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * (1-i)/sqrt2
    double2 o2 = mul_mi(o[2]);                                                     // * (-i)
    double2 o3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * (-1-i)/sqrt2
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}

// Twiddle tables in shared memory: w(m) = exp(-2 pi i m / Lf) = hi[m >> 5] * lo[m & 31]  (32 + Lf/32 entries instead of a
// global table of Lf entries: no global loads inside the passes).  lo is stored padded (i -> i + i/8) so that the
// stride-4 reads of the third pass are conflict-free.
struct TwTab { const double2* lo; const double2* hi; };
constexpr int kTwLo = 32 + 4;   // padded length of the low table

__device__ __forceinline__ double2 tw_get(const TwTab& T, int m) {
    const int l = m & 31;
    return cmul(T.hi[m >> 5], T.lo[l + (l >> 3)]);
}

// Multiply v[r] by w^r, w = exp(-2 pi i k twstep / Lf); w from the tables, powers by squaring / multiplying.
template <int R>
__device__ __forceinline__ void twiddle(double2 (&v)[R], int k, int twstep, const TwTab& T) {
#ifdef VCH_FFT_NOTW
    return;
#endif
    if (k == 0) return;
    const double2 w1 = tw_get(T, k * twstep);
    v[1] = cmul(v[1], w1);
    if (R >= 4) {
        const double2 w2 = cmul(w1, w1);
        const double2 w3 = cmul(w1, w2);
        v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
        if (R == 8) {
            const double2 w4 = cmul(w2, w2);
            v[4] = cmul(v[4], w4); v[5] = cmul(v[5], cmul(w1, w4));
            v[6] = cmul(v[6], cmul(w2, w4)); v[7] = cmul(v[7], cmul(w3, w4));
        }
    }
}

// All FFT geometry is a compile-time function of LOG2L (Lf = 2^LOG2L = 2N), so shared-memory offsets fold into
// immediates: with Lf/8 and Ns multiples of 8,  padi(x + r*S) = padi(x) + r*(S + S/8).
template <int LOG2L> struct FftGeom {
    static constexpr int Lf = 1 << LOG2L, tpf = Lf / 8;
    static constexpr int n8 = LOG2L / 3, rem = LOG2L % 3;
    static constexpr int mids = (rem == 0) ? n8 - 2 : n8 - 1;     // radix-8 passes strictly between first and last
    static constexpr int lastR = (rem == 0) ? 8 : (rem == 2 ? 4 : 2);
    static constexpr int lastNs = Lf / lastR;
    static constexpr int ld = Lf + Lf / 8 + 1;
};

// Middle radix-8 pass through padded shared memory: load, twiddle, DFT, barrier, store, barrier.
template <int LOG2L, int NS>
__device__ __forceinline__ void fft_mid_pass(double2* data, int t, const TwTab& tw) {
    using G = FftGeom<LOG2L>;
    double2 v[8];
    const int k = t & (NS - 1);
    const double2* src = data + padi(t);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = src[r * (G::tpf + G::tpf / 8)];
    twiddle<8>(v, k, G::Lf / (NS * 8), tw);
    dft<8>(v);
    __syncthreads();
    double2* dst = data + padi((t - k) * 8 + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) dst[r * (NS + NS / 8)] = v[r];
    __syncthreads();
}

// First pass on 8 register values (Ns = 1: no twiddles): DFT and autosort store, then a barrier.
__device__ __forceinline__ void fft_first_pass_store(double2* data, double2 (&v)[8], int t) {
    dft<8>(v);
    double2* dst = data + 9 * t;            // padi(8t + r) = 9t + r
#pragma unroll
    for (int r = 0; r < 8; ++r) dst[r] = v[r];
    __syncthreads();
}

template <int LOG2L>
__device__ __forceinline__ void fft_middle(double2* data, int t, const TwTab& tw) {
    using G = FftGeom<LOG2L>;
#ifndef VCH_FFT_NOMID
    if constexpr (G::mids >= 1) fft_mid_pass<LOG2L, 8>(data, t, tw);
    if constexpr (G::mids >= 2) fft_mid_pass<LOG2L, 64>(data, t, tw);
    if constexpr (G::mids >= 3) fft_mid_pass<LOG2L, 512>(data, t, tw);
#endif
}

// Last pass (radix 8 / 4 / 2): outputs stay in registers; z[i] has natural index t + FftOut<LOG2L>::off(i).
template <int LOG2L> struct FftOut {
    using G = FftGeom<LOG2L>;
    // i = m*R + r  ->  offset m*tpf + r*Ns  (j0 = j because j < Ns in the last pass)
    __host__ __device__ static constexpr int off(int i) { return (i / G::lastR) * G::tpf + (i % G::lastR) * G::lastNs; }
    // position of z[i] among the inputs {t + q*tpf} of a following first pass: q = off / tpf
    __host__ __device__ static constexpr int q(int i) { return off(i) / G::tpf; }
};

template <int LOG2L>
__device__ __forceinline__ void fft_last_pass(const double2* data, int t, const TwTab& tw, double2 (&z)[8]) {
    using G = FftGeom<LOG2L>;
    constexpr int R = G::lastR, NB = 8 / R, stride = G::Lf / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * G::tpf;
        double2 v[R];
        const double2* src = data + padi(j);
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[r * (stride + stride / 8)];
        twiddle<R>(v, j, 1, tw);             // k = j, twstep = Lf / (Ns R) = 1
        dft<R>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) z[m * R + r] = v[r];
    }
}

// ---- Fused BiCGStab vector work around the row transforms.
// Round-1 SASS showed the in-line prologue as 16 dependent groups of 3 loads + 1 store per thread (the stores to w may alias
// r/q/a as far as the compiler knows, so the next group's loads waited behind them) and the BiCGStab scalars reloaded and
// re-divided for every point.  The helpers below take restrict-qualified PARAMETERS (struct members carry no such promise), read
// the scalars once and issue the 4-8 loads of a point (both lines) back to back.  Measured on B200 (1024^2): fused prologue
// 17.2 -> 15.1 us (the plain transform takes 14.8), first epilogue 20.5 -> 19.6 us, +4.1 % PGD iterations/s.
// First-pass inputs of a row pair with the fused BiCGStab vector update (RowPrologue modes 1 and 2):
//   w = r + coef*q (coef = 0: w = r, q is not read — it may hold anything, bicg_init_kernel does not clear it),
//   stored once (the mirror images e > N are recomputed, not stored), x = (a - abar)*w or w.
// A function of its own so that the pointers are restrict-qualified PARAMETERS: the kernel's struct members carry no such
// promise, and without it every store to w ordered the loads of the next point behind it (8 dependent L2 round trips).
// r, q|v, a and w are distinct work vectors (enqueue_bicg_iteration).
template <int LOG2L, bool USEQ, bool HASA>
__device__ __forceinline__ void row_prologue_load_impl(double2 (&v)[8], const double* __restrict__ pr, const double* __restrict__ pq,
                                                       const double* __restrict__ pa, double* __restrict__ pw, double coef, double abar,
                                                       bool va, bool vb, size_t base_a, size_t base_b, int t, int N, int in_es) {
    using G = FftGeom<LOG2L>;
    // an absent line (the odd line out of the last pair, idle FFT slots of the last CTA) is loaded from line 0 and discarded: no branch
    // separates the loads, so the 4-6 loads of a point (both lines) are issued back to back and, the stores being known not to
    // alias them, the next point's loads can follow at once
    const size_t ba = va ? base_a : 0, bb = vb ? base_b : 0;   // line 0 always exists
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int e = t + r * G::tpf;
        const size_t o = (size_t)((e <= N ? e : G::Lf - e) * in_es);
        const double ra = pr[ba + o], rb = pr[bb + o];
        double qa = 0.0, qb = 0.0, aa = 0.0, ab = 0.0;
        if (USEQ) { qa = pq[ba + o]; qb = pq[bb + o]; }
        if (HASA) { aa = pa[ba + o]; ab = pa[bb + o]; }
        const double wa = USEQ ? ra + coef * qa : ra, wb = USEQ ? rb + coef * qb : rb;
        if (e <= N) { if (va) pw[ba + o] = wa; if (vb) pw[bb + o] = wb; }
        const double xa = HASA ? (aa - abar) * wa : wa, xb = HASA ? (ab - abar) * wb : wb;
        v[r] = make_double2(va ? xa : 0.0, vb ? xb : 0.0);
    }
}
template <int LOG2L>
__device__ __forceinline__ void row_prologue_load(double2 (&v)[8], const double* pr, const double* pq, const double* pa, double* pw,
                                                  double coef, double abar, bool va, bool vb, size_t base_a, size_t base_b, int t,
                                                  int N, int in_es) {
    // coef == 0 (first iteration of a solve: beta = 0): q is not read — bicg_init_kernel does not clear it
    if (coef != 0.0) {
        if (pa) row_prologue_load_impl<LOG2L, true, true>(v, pr, pq, pa, pw, coef, abar, va, vb, base_a, base_b, t, N, in_es);
        else row_prologue_load_impl<LOG2L, true, false>(v, pr, pq, pa, pw, coef, abar, va, vb, base_a, base_b, t, N, in_es);
    } else {
        if (pa) row_prologue_load_impl<LOG2L, false, true>(v, pr, pq, pa, pw, coef, abar, va, vb, base_a, base_b, t, N, in_es);
        else row_prologue_load_impl<LOG2L, false, false>(v, pr, pq, pa, pw, coef, abar, va, vb, base_a, base_b, t, N, in_es);
    }
}


// Same idea for the fused row epilogue (out = (mul_a - abar) z + addend, BiCGStab dot products): with in-line code every store
// to `out` held back the loads of the next point.  Here the up to 8 loads of a point (both lines: addend, other, r, a) are issued
// together; `out` is distinct from the vectors that are read (enqueue_bicg_iteration).  Absent lines read line 0 and are
// discarded.  mode 4 also accumulates (other, other) and (rvec, other) — (s,s) and (r0,s) of the 6-launch iteration.
template <int LOG2L>
__device__ __forceinline__ void row_epilogue_store(const double2 (&z)[8], double* __restrict__ out, const double* __restrict__ addend,
                                                   const double* __restrict__ other, const double* __restrict__ rvec,
                                                   const double* __restrict__ mul_a, double eabar, int mode, bool va, bool vb,
                                                   size_t base_a, size_t base_b, int t, int N, int out_es,
                                                   double& acc1, double& acc2, double& acc3, double& acc4, double& acc5) {
    const size_t ba = va ? base_a : 0, bb = vb ? base_b : 0;
    const bool has_add = addend != nullptr, has_mul = mul_a != nullptr, has_r = rvec != nullptr && mode != 0, has_o = mode != 0;
    const bool m4 = mode == 4;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int kk = t + FftOut<LOG2L>::off(q);
        if (kk <= N) {
            const size_t off = (size_t)(kk * out_es);
            const double ma = has_mul ? mul_a[ba + off] : 0.0, mb = has_mul ? mul_a[bb + off] : 0.0;
            const double da = has_add ? addend[ba + off] : 0.0, db = has_add ? addend[bb + off] : 0.0;
            const double oa = has_o ? other[ba + off] : 0.0, ob = has_o ? other[bb + off] : 0.0;
            const double ra = has_r ? rvec[ba + off] : 0.0, rb = has_r ? rvec[bb + off] : 0.0;
            const double zx = has_mul ? (ma - eabar) * z[q].x : z[q].x, zy = has_mul ? (mb - eabar) * z[q].y : z[q].y;
            const double xo = has_add ? zx + da : zx, yo = has_add ? zy + db : zy;
            if (va) {
                out[ba + off] = xo;
                if (has_o) { acc1 += oa * xo; acc2 += xo * xo; if (has_r) acc3 += ra * xo; }
                if (m4) { acc4 += oa * oa; acc5 += ra * oa; }
            }
            if (vb) {
                out[bb + off] = yo;
                if (has_o) { acc1 += ob * yo; acc2 += yo * yo; if (has_r) acc3 += rb * yo; }
                if (m4) { acc4 += ob * ob; acc5 += rb * ob; }
            }
        }
    }
}

// One CTA = ppb complex FFTs (2*ppb lines).  Line l, element e lives at base[l*line_stride + e*elem_stride].
//   SOLVE = false: out = DCT-I(in) per line (unnormalised "FFT of the even extension").
//   SOLVE = true : fused  forward transform -> divide by symbol -> inverse transform  (column solve, in place).
#ifndef VCH_FFT_MINB
#define VCH_FFT_MINB (1024 / MAXT)
#endif
// XM: slab transposition mode of the row kernel, compile-time (0 none, 1 transposing store, 3 gathering load; see Scatter)
template <int LOG2L, bool SOLVE, int MAXT, int XM = 0>
__global__ void __launch_bounds__(MAXT, VCH_FFT_MINB)
dct_fft_kernel(const double* in, double* out, int nlines, int n, int in_ls, int in_es, int out_ls, int out_es, int ppb,
               const double2* __restrict__ twg, const double* __restrict__ lam_line, const double* __restrict__ lam_elem,
               SymbolArgs sy, double norm, int scale_mode, RowPrologue pro, DotEpilogue epi, const int* __restrict__ done,
               const __grid_constant__ Scatter sct) {   // __grid_constant__: sct.peer[r] is indexed straight from the constant bank
    pdl_enter();
    using G = FftGeom<LOG2L>;
    constexpr int Lf = G::Lf, tpf = G::tpf, ld = G::ld;
    if (done && *done) return;
#ifdef VCH_CPU_EMU
    double2* sm = reinterpret_cast<double2*>(vch_emu::dynamic_smem());
#else
    extern __shared__ double2 sm[];
#endif
    const int N = n - 1;
    const int f = threadIdx.x / tpf, t = threadIdx.x - f * tpf;
    // twiddle tables first (visible after the first-pass barrier), FFT buffers behind them
    double2* tlo = sm;
    double2* thi = sm + kTwLo;
    for (int i = threadIdx.x; i < 32 + (Lf >> 5); i += blockDim.x) {
        if (i < 32) tlo[i + (i >> 3)] = twg[i];
        else thi[i - 32] = twg[(i - 32) << 5];
    }
    const TwTab tw{tlo, thi};
    double2* data = sm + kTwLo + (Lf >> 5) + (size_t)f * ld;
    const int la = 2 * (blockIdx.x * ppb + f), lb = la + 1;
    const bool va = la < nlines, vb = lb < nlines;
    const double* pa = in + (size_t)la * in_ls;
    const double* pb = in + (size_t)lb * in_ls;

    // ---- first pass straight from global memory (even extension by index reflection).
    // SOLVE (columns): the two lines of a pair are adjacent doubles of a pitched buffer whose pitch is even, so one
    // 16-byte access moves both.
    double2 v[8];
    if (!SOLVE && XM == 3) {
        // slab mode: gather the row pair from the column owners' buffers.  Staged through shared memory with 16-byte loads so
        // that every element crosses NVLink once (the direct path below re-reads each element for its mirror image, which
        // is free from L1/L2 but doubles remote traffic).  The staging area is the FFT's own data buffer.
        double* stg = reinterpret_cast<double*>(data);           // line a at [0, N], line b at [N + 2, 2N + 2]
        for (int k = 2 * t; k <= N; k += 2 * tpf) {
            int rr = k >> sct.shift; if (rr >= sct.nr) rr = sct.nr - 1;
            const double* src = sct.peer[rr] + sct.off + (size_t)(sct.base + la) * sct.pitch + (k - (rr << sct.shift));
            if (k < N) {
                const double2 a2 = va ? *reinterpret_cast<const double2*>(src) : make_double2(0.0, 0.0);
                const double2 b2 = vb ? *reinterpret_cast<const double2*>(src + sct.pitch) : make_double2(0.0, 0.0);
                stg[k] = a2.x; stg[k + 1] = a2.y; stg[N + 2 + k] = b2.x; stg[N + 3 + k] = b2.y;
            } else {
                stg[k] = va ? src[0] : 0.0; stg[N + 2 + k] = vb ? src[sct.pitch] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = t + r * tpf;
            const int ee = (e <= N) ? e : Lf - e;
            v[r] = make_double2(stg[ee], stg[N + 2 + ee]);
        }
        __syncthreads();                                          // staging is overwritten by the first-pass store
    } else
    if (!SOLVE && XM == 0 && (pro.mode == 3 || pro.mode == 2)) {
        // 6-launch iteration.  mode 3: deferred x/r update of the previous iteration + the new p; mode 2: s = r - alpha v.  Every
        // element is touched by exactly one thread (so p and x can be updated in place); the transform inputs go through shared
        // memory, from where the first pass reads them with the even extension (as in the slab gather above) — half the global
        // loads of the recomputing prologue.
        const Scal* sc = pro.sc;
        const bool first = sc->iters == 0, m2 = pro.mode == 2;
        const double al = sc->alpha, om = sc->omega, abar = sc->abar;
        const double beta = (first || m2) ? 0.0 : (sc->rho_new / sc->rho) * (al / om);
        double* stg = reinterpret_cast<double*>(data);            // line a at [0, N], line b at [N + 2, 2N + 2]
        for (int k = t; k <= N; k += tpf) {
#pragma unroll
            for (int L = 0; L < 2; ++L) {
                const bool valid = L ? vb : va;
                double xin = 0.0;
                if (valid) {
                    const size_t idx = (size_t)(L ? lb : la) * in_ls + k;
                    double pn;
                    if (m2) pn = pro.r[idx] - al * pro.qv[idx];
                    else if (first) pn = pro.r[idx];
                    else {
                        const double sv = pro.s[idx], tv = pro.t[idx], pv = pro.w[idx];
                        const double rn = sv - om * tv;
                        pro.x[idx] += al * pv + om * sv;
                        pro.rw[idx] = rn;
                        pn = rn + beta * (pv - om * pro.qv[idx]);
                    }
                    pro.w[idx] = pn;
                    xin = pro.a ? (pro.a[idx] - abar) * pn : pn;
                }
                stg[(L ? N + 2 : 0) + k] = xin;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = t + r * tpf;
            const int ee = (e <= N) ? e : Lf - e;
            v[r] = make_double2(stg[ee], stg[N + 2 + ee]);
        }
        __syncthreads();
    } else if (!SOLVE && pro.mode != 0) {
        // 7-launch iteration (slab mode, and the reference form the 6-launch one is tested against): p = r + beta q / s = r - alpha v
        // recomputed where the mirror image is needed, written once
        const Scal* sc = pro.sc;
        const double coef = (pro.mode == 1) ? (sc->rho_new / sc->rho) * (sc->alpha / sc->omega) : -sc->alpha;
        row_prologue_load<LOG2L>(v, pro.r, pro.qv, pro.a, pro.w, coef, sc->abar, va, vb, (size_t)la * in_ls, (size_t)lb * in_ls, t, N, in_es);
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = t + r * tpf;
            const int off = (e <= N ? e : Lf - e) * in_es;
            if (SOLVE) {
                v[r] = va ? *reinterpret_cast<const double2*>(pa + off) : make_double2(0.0, 0.0);
            } else {
                v[r].x = va ? pa[off] : 0.0;
                v[r].y = vb ? pb[off] : 0.0;
            }
        }
    }
    fft_first_pass_store(data, v, t);
    fft_middle<LOG2L>(data, t, tw);
    double2 z[8];
    fft_last_pass<LOG2L>(data, t, tw, z);

    if (SOLVE) {
        // Forward outputs -> multiply by the spectral factor -> inverse transform.  The last pass leaves every thread
        // with exactly the 8 spectrum entries {t + q*tpf} that the first pass of the next transform consumes, so the
        // hand-over stays in registers (entries beyond N are the thread's own copy of the mirror image; the spectrum of
        // a real-even line is even up to rounding).
        const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
        const double sc0 = sy.coef_ptr ? sy.coef_ptr[0] : sy.c0, sc2 = sy.coef_ptr ? sy.coef_ptr[1] : sy.c2;
        const double lla = va ? lam_line[la] : 0.0, llb = vb ? lam_line[lb] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = t + FftOut<LOG2L>::off(i);
            const int kf = (kk <= N) ? kk : Lf - kk;
            const double le = lam_elem[kf];
            const double s1 = le + lla, s2 = le + llb;
            double f1 = norm / (sc0 + s1 * (abar + sc2 * s1)), f2 = norm / (sc0 + s2 * (abar + sc2 * s2));
            if (scale_mode == 1) { f1 *= s1; f2 *= s2; }
            v[FftOut<LOG2L>::q(i)] = make_double2(z[i].x * f1, z[i].y * f2);   // static permutation into first-pass order
        }
        __syncthreads();                         // every thread has finished reading the forward data
        fft_first_pass_store(data, v, t);
        fft_middle<LOG2L>(data, t, tw);
        fft_last_pass<LOG2L>(data, t, tw, z);
    }

    double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0, acc4 = 0.0, acc5 = 0.0;
    const double eabar = epi.mul_a ? epi.sc->abar : 0.0;
    if (SOLVE || XM == 1) {
        double* qa = out + (size_t)la * out_ls;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int kk = t + FftOut<LOG2L>::off(q);
            if (kk <= N) {
                if (!SOLVE) {   // slab mode: transposing store into the owning peer's buffer
                    int r = kk >> sct.shift; if (r >= sct.nr) r = sct.nr - 1;
                    const int kl = kk - (r << sct.shift);
                    double* dst = sct.peer[r] + sct.off;
                    if (va) dst[(size_t)(sct.base + la) * sct.pitch + kl] = z[q].x;
                    if (vb) dst[(size_t)(sct.base + lb) * sct.pitch + kl] = z[q].y;
                } else {
                    if (va) *reinterpret_cast<double2*>(qa + kk * out_es) = make_double2(z[q].x, vb ? z[q].y : 0.0);
                }
            }
        }
    } else {
        row_epilogue_store<LOG2L>(z, out, epi.addend, epi.other, epi.rvec, epi.mul_a, eabar, epi.mode, va, vb, (size_t)la * out_ls,
                                  (size_t)lb * out_ls, t, N, out_es, acc1, acc2, acc3, acc4, acc5);
    }
    if (epi.mode == 4) {     // block-uniform: every thread of every CTA takes part in the reduction
        double vals[5] = {acc1, acc2, acc3, acc4, acc5};
        const int op[5] = {0, 0, 0, 0, 0};
        double tot[5];
        if (grid_reduce<5>(vals, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) dots_finish6(epi, tot);
    } else if (epi.mode) {
        double vals[3] = {acc1, acc2, acc3};
        const int op[3] = {0, 0, 0};
        double tot[3];
        if (grid_reduce<3>(vals, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) dots_finish(epi, tot);
    }
}

}  // namespace vch
#include "vch_fft16.cuh"
namespace vch {

// Dense-table fallbacks for N that is not a power of two (small validation grids).
__global__ void dct_rows_dense_kernel(const double* __restrict__ in, double* __restrict__ out, int lines, int n,
                                      const double* __restrict__ Tt, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)lines * n) return;
    const int l = (int)(idx / n), k = (int)(idx - (long long)l * n);
    const double* row = in + (size_t)l * n;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc = fma(Tt[(size_t)j * n + k], row[j], acc);
    out[idx] = acc;
}
__global__ void dct_cols_dense_kernel(const double* __restrict__ in, double* __restrict__ out, int no, int ni,
                                      const double* __restrict__ Tt, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)no * ni) return;
    const int k = (int)(idx / ni), c = (int)(idx - (long long)k * ni);
    double acc = 0.0;
    for (int o = 0; o < no; ++o) acc = fma(Tt[(size_t)o * no + k], in[(size_t)o * ni + c], acc);
    out[idx] = acc;
}
__global__ void dct_scale_kernel(double* __restrict__ d, int no, int ni, const double* __restrict__ lam_o,
                                 const double* __restrict__ lam_i, SymbolArgs sy, double norm,
                                 const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)no * ni) return;
    const int k = (int)(idx / ni), c = (int)(idx - (long long)k * ni);
    const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
    const double sc0 = sy.coef_ptr ? sy.coef_ptr[0] : sy.c0, sc2 = sy.coef_ptr ? sy.coef_ptr[1] : sy.c2;
    const double l = lam_o[k] + lam_i[c];
    d[idx] *= norm / (sc0 + l * (abar + sc2 * l));
}
// Dense-path stand-ins for the fused prologue / epilogue / spectral factor.
__global__ void dct_prologue_kernel(RowPrologue pro, double* __restrict__ x, long long n, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const Scal* sc = pro.sc;
    const double coef = (pro.mode == 1) ? (sc->rho_new / sc->rho) * (sc->alpha / sc->omega) : -sc->alpha;
    const double abar = sc->abar;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double w = (coef != 0.0) ? pro.r[idx] + coef * pro.qv[idx] : pro.r[idx];
        pro.w[idx] = w;
        x[idx] = pro.a ? (pro.a[idx] - abar) * w : w;
    }
}
__global__ void dct_addend_kernel(double* __restrict__ outv, const double* __restrict__ addend, const double* __restrict__ mul_a,
                                  const Scal* __restrict__ sc, long long n, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const double abar = mul_a ? sc->abar : 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double z = mul_a ? (mul_a[idx] - abar) * outv[idx] : outv[idx];
        outv[idx] = addend ? z + addend[idx] : z;
    }
}
__global__ void dct_lambda_kernel(double* __restrict__ d, int no, int ni, const double* __restrict__ lam_o,
                                  const double* __restrict__ lam_i, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)no * ni) return;
    const int k = (int)(idx / ni), c = (int)(idx - (long long)k * ni);
    d[idx] *= lam_o[k] + lam_i[c];
}
// Stand-alone BiCGStab dots for the paths without the fused epilogue.
__global__ void dct_dots_kernel(const double* __restrict__ outv, DotEpilogue epi, long long n, const int* __restrict__ done) {
    pdl_enter();
    if (done && *done) return;
    double vals[3] = {0.0, 0.0, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const double zz = outv[idx];
        vals[0] += epi.other[idx] * zz; vals[1] += zz * zz;
        if (epi.rvec) vals[2] += epi.rvec[idx] * zz;
    }
    const int op[3] = {0, 0, 0};
    double tot[3];
    if (grid_reduce<3>(vals, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) dots_finish(epi, tot);
}

// ------------------------------------------------------------------------------------------------ host side
#ifndef VCH_CPU_EMU_KERNELS_ONLY   // (the kernel-level CPU emulation harnesses of tests/emu drive the kernels above directly)
static inline void dct_axis_init(DctAxis& ax, int n, double h) {
    const int N = n - 1;
    ax.n = n;
    ax.fft = (N >= 32) && ((N & (N - 1)) == 0) && (N <= 4096);
    std::vector<double> lam(n);
    for (int k = 0; k < n; ++k) {
        const long double s = sinl(3.14159265358979323846264338327950288L * k / (2.0L * N));
        lam[k] = (double)(4.0L * s * s / ((long double)h * h));
    }
    VCH_CUDA(cudaMalloc(&ax.lam, n * sizeof(double)));
    VCH_CUDA(cudaMemcpy(ax.lam, lam.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    if (ax.fft) {
        ax.Lf = 2 * N;
        ax.log2L = 0;
        while ((1 << ax.log2L) < ax.Lf) ++ax.log2L;
        std::vector<double2> tw(ax.Lf);
        for (int m = 0; m < ax.Lf; ++m) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * m / ax.Lf;
            tw[m] = make_double2((double)cosl(a), (double)sinl(a));
        }
        VCH_CUDA(cudaMalloc(&ax.tw, ax.Lf * sizeof(double2)));
        VCH_CUDA(cudaMemcpy(ax.tw, tw.data(), ax.Lf * sizeof(double2), cudaMemcpyHostToDevice));
        const std::vector<double2> t16 = fft16_twiddles(ax.log2L);
        VCH_CUDA(cudaMalloc(&ax.tw16, t16.size() * sizeof(double2)));
        VCH_CUDA(cudaMemcpy(ax.tw16, t16.data(), t16.size() * sizeof(double2), cudaMemcpyHostToDevice));
    } else {
        std::vector<double> T((size_t)n * n);
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                const long double cj = (j == 0 || j == N) ? 0.5L : 1.0L;
                T[(size_t)j * n + k] = (double)(2.0L * cj * cosl(3.14159265358979323846264338327950288L * j * k / N));
            }
        VCH_CUDA(cudaMalloc(&ax.denseT, T.size() * sizeof(double)));
        VCH_CUDA(cudaMemcpy(ax.denseT, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
}

static inline int dct_rows_ppb(const DctAxis& ax, int lines) {
    const int tpf = ax.Lf >> 3;
    int ppb = 256 / tpf; if (ppb < 1) ppb = 1;
    const int pairs = (lines + 1) / 2;
    if (ppb > pairs) ppb = pairs;
    return ppb;
}
static inline int dct_cols_ppb(const DctAxis& ax, int ncols) {
    const int tpf = ax.Lf >> 3;
#ifdef VCH_FFT_COLS_PPB1
    int ppb = 256 / tpf;
#else
    int ppb = 512 / tpf;
    if (ppb < 2 && tpf <= 256) ppb = 2;       // >= 4 adjacent columns so fetched 32-byte sectors are fully used (CTA <= 512 threads)
#endif
    if (ppb < 1) ppb = 1;
    if (ppb > 8) ppb = 8;
    const int pairs = (ncols + 1) / 2;
    if (ppb > pairs) ppb = pairs;
    return ppb;
}
static inline size_t dct_smem(const DctAxis& ax, int ppb) {
    return sizeof(double2) * ((size_t)ppb * (ax.Lf + (ax.Lf >> 3) + 1) + kTwLo + (ax.Lf >> 5));
}

inline void DctPlan::init(int no_, int ni_, double h_outer, double h_inner, LaunchLog* launch_log) {
    no = no_; ni = ni_; log = launch_log;
    dct_axis_init(inner, ni, h_inner);
    dct_axis_init(outer, no, h_outer);
    pitch = (ni + 7) & ~7;        // whole column groups of the column kernels (8 columns per CTA); padding columns stay zero
    tmp1.alloc((size_t)no * pitch);
    VCH_CUDA(cudaMemset(tmp1.p, 0, (size_t)no * pitch * sizeof(double)));
    tmp2.alloc((size_t)no * ni);
    const int big = 200 * 1024;
#define VCH_FFT_ATTR(LG)                                                                                                                        \
    VCH_CUDA(cudaFuncSetAttribute(dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024)>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)); \
    VCH_CUDA(cudaFuncSetAttribute(dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024), 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)); \
    VCH_CUDA(cudaFuncSetAttribute(dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024), 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)); \
    VCH_CUDA(cudaFuncSetAttribute(dct_fft_kernel<LG, true, ((1 << LG) / 8 <= 512 ? 512 : 1024)>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    for (const DctAxis* ax : {&inner, &outer}) {
        if (!ax->fft) continue;
        switch (ax->log2L) {
            case 6: VCH_FFT_ATTR(6) break; case 7: VCH_FFT_ATTR(7) break; case 8: VCH_FFT_ATTR(8) break; case 9: VCH_FFT_ATTR(9) break;
            case 10: VCH_FFT_ATTR(10) break; case 11: VCH_FFT_ATTR(11) break; case 12: VCH_FFT_ATTR(12) break; case 13: VCH_FFT_ATTR(13) break;
            default: break;
        }
    }
#undef VCH_FFT_ATTR
#define VCH_F16_ATTR(LG)                                                                   \
    case LG:                                                                               \
        rows16_set_attributes<LG, 0>();                                                    \
        fft16_attr(cols16_kernel<LG>, F16<LG>::cols_smem_bytes);                           \
        if (LG >= 8) { rows16_set_attributes<(LG >= 8 ? LG : 8), 1>(); rows16_set_attributes<(LG >= 8 ? LG : 8), 3>(); } \
        break;
    if (inner.fft && outer.fft) {
        if (getenv("VCH_FFT16")) lean = atoi(getenv("VCH_FFT16")) != 0;
        for (const DctAxis* ax : {&inner, &outer}) {
            switch (ax->log2L) {
                VCH_F16_ATTR(6) VCH_F16_ATTR(7) VCH_F16_ATTR(8) VCH_F16_ATTR(9) VCH_F16_ATTR(10) VCH_F16_ATTR(11) VCH_F16_ATTR(12) VCH_F16_ATTR(13)
                default: break;
            }
        }
    } else lean = false;
#undef VCH_F16_ATTR
    if (lean) {
        have_tmap = cols16_tensor_map(&tmap, outer.log2L, tmp1.p, pitch, no);
#ifndef VCH_CPU_EMU
#define VCH_F16T_ATTR(LG) case LG: if (F16T<LG>::use) fft16_attr(cols16_tma_kernel<LG>, F16T<LG>::smem_bytes); break;
        switch (outer.log2L) { VCH_F16T_ATTR(9) VCH_F16T_ATTR(10) VCH_F16T_ATTR(11) VCH_F16T_ATTR(12) VCH_F16T_ATTR(13) default: break; }
#undef VCH_F16T_ATTR
#endif
    }
}

// Slab mode: square global grid n_global x n_global (N = n_global - 1 a power of two), T1/T2 carved from the arena by the caller.
inline void DctPlan::init_slab(int n_global, double h_outer, double h_inner, LaunchLog* launch_log, const DctSlab& sl) {
    init(n_global, n_global, h_outer, h_inner, launch_log);
    if (!(inner.fft && outer.fft)) throw Error(VCH_E_SHAPE, "slab mode needs N = 2^k, 32 <= N <= 4096");
    tmp1.release(); tmp2.release();
    slab = sl; slab.on = true;
    have_tmap = lean && cols16_tensor_map(&tmap, outer.log2L, slab.T1, slab.p1, no);
}
inline void DctPlan::barrier(cudaStream_t s, const int* done) {
    log->begin("xbar", s);
    xbar_kernel<<<1, 32, 0, s>>>(slab.cm, done);
    log->end(s);
}

inline int DctPlan::max_grid() const {
    int g = kRedBlocksMax;
    if (inner.fft) { const int ppb = dct_rows_ppb(inner, no); g = std::max(g, ((no + 1) / 2 + ppb - 1) / ppb); }
    return g;
}

inline void DctPlan::apply(cudaStream_t s, const double* in, double* out, const SymbolArgs& sym, const int* done,
                           const DotEpilogue& epi, const RowPrologue& pro, int scale_mode) {
    const long long n = (long long)no * ni;
    const int eb = (int)((n + 255) / 256);
    const double norm = 1.0 / (4.0 * (double)(ni - 1) * (double)(no - 1));
    const SymbolArgs nosym{1.0, 0.0, nullptr, 0.0, nullptr};
    const bool all_fft = inner.fft && outer.fft;
    // tmp1 is pitched only when both axes use the FFT kernels (the dense kernels address dense arrays)
    const int P = all_fft ? pitch : ni;
    double* t1 = tmp1.p;
    double* t2 = tmp2.p;
    auto fft_launch = [&](bool solve, const DctAxis& ax, int grid, int threads, size_t smem, const double* a, double* b, int nl,
                          int nn, int ils, int ies, int ols, int oes, const double* ll, const double* le, const SymbolArgs& sy,
                          double nrm, int smode, const RowPrologue& pr, const DotEpilogue& ep, const Scatter& sc8 = Scatter()) {
        const int ppb_ = threads / (ax.Lf >> 3);
        const bool pd = pdl;
#define VCH_FFT_ARGS a, b, nl, nn, ils, ies, ols, oes, ppb_, ax.tw, ll, le, sy, nrm, smode, pr, ep, done, sc8
#define VCH_FFT_CASE(LG)                                                                                                      \
        case LG:                                                                                                              \
            if (solve) launch_pdl(pd, dct_fft_kernel<LG, true, ((1 << LG) / 8 <= 512 ? 512 : 1024)>, grid, threads, smem, s, VCH_FFT_ARGS); \
            else if (sc8.mode == 1) launch_pdl(pd, dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024), 1>, grid, threads, smem, s, VCH_FFT_ARGS); \
            else if (sc8.mode == 3) launch_pdl(pd, dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024), 3>, grid, threads, smem, s, VCH_FFT_ARGS); \
            else launch_pdl(pd, dct_fft_kernel<LG, false, ((1 << LG) / 8 <= 512 ? 512 : 1024)>, grid, threads, smem, s, VCH_FFT_ARGS); \
            break;
        switch (ax.log2L) {
            VCH_FFT_CASE(6) VCH_FFT_CASE(7) VCH_FFT_CASE(8) VCH_FFT_CASE(9) VCH_FFT_CASE(10) VCH_FFT_CASE(11) VCH_FFT_CASE(12)
            VCH_FFT_CASE(13)
            default: throw Error(VCH_E_ARG, "unsupported FFT length");
        }
#undef VCH_FFT_CASE
#undef VCH_FFT_ARGS
    };
    if (slab.on && lean && inner.log2L >= 8 && (pro.mode == 0 || pro.mode == 2 || pro.mode == 3) && (epi.mode == 0 || epi.mode == 1 || epi.mode == 4)) {
        // slab mode with the radix-16 kernels: same protocol as below (rows scattered into the column owners' T1 | barrier | column
        // solve in place on the owned columns | barrier | rows gathered back), same barriers
        const DctSlab& sl = slab;
        Scatter s1; s1.mode = 1; s1.shift = sl.shift; s1.nr = sl.cm.nranks; s1.base = sl.o0; s1.pitch = sl.p1;
        s1.off = (size_t)(sl.T1 - sl.cm.peer[sl.cm.rank]);
        for (int r = 0; r < sl.cm.nranks; ++r) s1.peer[r] = sl.cm.peer[r];
        Scatter s3 = s1; s3.mode = 3;
#define VCH_F16_SLAB(CALL)                                                                                        \
        switch (inner.log2L) {                                                                                    \
            case 8: { constexpr int LG = 8; CALL; } break;   case 9: { constexpr int LG = 9; CALL; } break;       \
            case 10: { constexpr int LG = 10; CALL; } break; case 11: { constexpr int LG = 11; CALL; } break;     \
            case 12: { constexpr int LG = 12; CALL; } break; case 13: { constexpr int LG = 13; CALL; } break;     \
            default: throw Error(VCH_E_ARG, "unsupported FFT length");                                            \
        }
        log->begin(pro.mode ? "rows16_pro" : "rows16", s);
        VCH_F16_SLAB((rows16_forward<LG, 1>(false, s, in, sl.T1, sl.nloc, ni, sl.p1, inner.tw16, pro, done, s1)))
        log->end(s);
        barrier(s, done);
        log->begin("cols16_solve", s);
        VCH_F16_SLAB((cols16_solve<LG>(false, s, sl.T1, sl.p1, sl.wloc, outer.tw16, inner.lam + sl.col0, outer.lam, sym, norm, scale_mode, done, have_tmap ? &tmap : nullptr)))
        log->end(s);
        barrier(s, done);
        log->begin(epi.mode == 1 ? "rows16_epi1" : (epi.mode == 4 ? "rows16_epi4" : "rows16"), s);
        VCH_F16_SLAB((rows16_inverse<LG, 3>(false, s, sl.T1, out, sl.nloc, sl.p1, ni, inner.tw16, epi, done, s3)))
        log->end(s);
        if (!epi.mode) barrier(s, done);
#undef VCH_F16_SLAB
        VCH_CUDA(cudaGetLastError());
        return;
    }
    if (slab.on) {
        // rows of the owned slab, stored transposed into the column owners' T1 | barrier | fused solve in place on the owned
        // columns (all global rows) | barrier | inverse rows, gathered back from the owners' T1.  Re-use of T1 by the next
        // application is safe because a cross-rank synchronisation always follows: the fused dot-product reduction of the
        // last kernel, or the explicit trailing barrier when there is none.
        const DctSlab& sl = slab;
        const int rp = dct_rows_ppb(inner, sl.nloc), rt = rp * (inner.Lf >> 3), rg = ((sl.nloc + 1) / 2 + rp - 1) / rp;
        const int cp = dct_cols_ppb(outer, sl.wloc), ct = cp * (outer.Lf >> 3), cg = ((sl.wloc + 1) / 2 + cp - 1) / cp;
        Scatter s1; s1.mode = 1; s1.shift = sl.shift; s1.nr = sl.cm.nranks; s1.base = sl.o0; s1.pitch = sl.p1;
        s1.off = (size_t)(sl.T1 - sl.cm.peer[sl.cm.rank]);
        for (int r = 0; r < sl.cm.nranks; ++r) s1.peer[r] = sl.cm.peer[r];
        Scatter s3 = s1; s3.mode = 3;
        log->begin(pro.mode ? "dct_rows_fft_pro" : "dct_rows_fft", s);
        fft_launch(false, inner, rg, rt, dct_smem(inner, rp), in, sl.T1, sl.nloc, ni, ni, 1, sl.p1, 1, nullptr, nullptr, nosym, 1.0, 0, pro, DotEpilogue(), s1);
        log->end(s);
        barrier(s, done);
        log->begin("dct_cols_fft_solve", s);
        fft_launch(true, outer, cg, ct, dct_smem(outer, cp), sl.T1, sl.T1, sl.wloc, no, 1, sl.p1, 1, sl.p1, inner.lam + sl.col0, outer.lam, sym, norm, scale_mode, RowPrologue(), DotEpilogue());
        log->end(s);
        barrier(s, done);
        log->begin(epi.mode == 1 ? "dct_rows_fft_epi1" : (epi.mode == 2 ? "dct_rows_fft_epi2" : "dct_rows_fft"), s);
        fft_launch(false, inner, rg, rt, dct_smem(inner, rp), sl.T1, out, sl.nloc, ni, sl.p1, 1, ni, 1, nullptr, nullptr, nosym, 1.0, 0, RowPrologue(), epi, s3);
        log->end(s);
        if (!epi.mode) barrier(s, done);
        VCH_CUDA(cudaGetLastError());
        return;
    }
    const int rppb = inner.fft ? dct_rows_ppb(inner, no) : 0, rthreads = inner.fft ? rppb * (inner.Lf >> 3) : 0;
    const int rgrid = inner.fft ? ((no + 1) / 2 + rppb - 1) / rppb : 0;
    const int cppb = outer.fft ? dct_cols_ppb(outer, ni) : 0, cthreads = outer.fft ? cppb * (outer.Lf >> 3) : 0;
    const int cgrid = outer.fft ? ((ni + 1) / 2 + cppb - 1) / cppb : 0;

    if (all_fft && lean && (pro.mode == 0 || pro.mode == 2 || pro.mode == 3 || pro.mode == 4) &&
        (epi.mode == 0 || epi.mode == 1 || epi.mode == 4 || epi.mode == 5)) {
        // radix-16 kernels (vch_fft16.cuh), one instantiation per fused mode: rows (prologue) -> column solve -> rows (epilogue)
        const Scatter nosct;
#define VCH_F16_SWITCH(AX, CALL)                                                                                  \
        switch ((AX).log2L) {                                                                                     \
            case 6: { constexpr int LG = 6; CALL; } break;   case 7: { constexpr int LG = 7; CALL; } break;       \
            case 8: { constexpr int LG = 8; CALL; } break;   case 9: { constexpr int LG = 9; CALL; } break;       \
            case 10: { constexpr int LG = 10; CALL; } break; case 11: { constexpr int LG = 11; CALL; } break;     \
            case 12: { constexpr int LG = 12; CALL; } break; case 13: { constexpr int LG = 13; CALL; } break;     \
            default: throw Error(VCH_E_ARG, "unsupported FFT length");                                            \
        }
        log->begin(pro.mode == 4 ? "rows16_schur" : (pro.mode ? "rows16_pro" : "rows16"), s);
        VCH_F16_SWITCH(inner, (rows16_forward<LG, 0>(pdl, s, in, t1, no, ni, P, inner.tw16, pro, done, nosct)))
        log->end(s);
        log->begin("cols16_solve", s);
        VCH_F16_SWITCH(outer, (cols16_solve<LG>(pdl, s, t1, P, ni, outer.tw16, inner.lam, outer.lam, sym, norm, scale_mode, done, have_tmap ? &tmap : nullptr)))
        log->end(s);
        log->begin(epi.mode == 1 ? "rows16_epi1" : (epi.mode == 4 ? "rows16_epi4" : (epi.mode == 5 ? "rows16_init" : "rows16")), s);
        VCH_F16_SWITCH(inner, (rows16_inverse<LG, 0>(pdl, s, t1, out, no, P, ni, inner.tw16, epi, done, nosct)))
        log->end(s);
        VCH_CUDA(cudaGetLastError());
        return;
    }
    if (pro.mode == 4 || epi.mode == 5) throw Error(VCH_E_ARG, "DctPlan::apply: prologue mode 4 / epilogue mode 5 need the radix-16 kernels");
    if (all_fft) {
        // rows (prologue fused) -> fused column solve in place on the pitched buffer -> rows (addend + dots fused)
        log->begin(pro.mode ? "dct_rows_fft_pro" : "dct_rows_fft", s);
        fft_launch(false, inner, rgrid, rthreads, dct_smem(inner, rppb), in, t1, no, ni, ni, 1, P, 1, nullptr, nullptr, nosym, 1.0, 0, pro, DotEpilogue());
        log->end(s);
        log->begin("dct_cols_fft_solve", s);
        fft_launch(true, outer, cgrid, cthreads, dct_smem(outer, cppb), t1, t1, ni, no, 1, P, 1, P, inner.lam, outer.lam, sym, norm, scale_mode, RowPrologue(), DotEpilogue());
        log->end(s);
        log->begin(epi.mode == 1 ? "dct_rows_fft_epi1" : (epi.mode == 2 ? "dct_rows_fft_epi2" : "dct_rows_fft"), s);
        fft_launch(false, inner, rgrid, rthreads, dct_smem(inner, rppb), t1, out, no, ni, P, 1, ni, 1, nullptr, nullptr, nosym, 1.0, 0, RowPrologue(), epi);
        log->end(s);
        VCH_CUDA(cudaGetLastError());
        return;
    }
    // ---- mixed / dense path (validation grids): prologue, transforms, spectral factor and epilogue as separate kernels
    const double* src = in;
    if (pro.mode) {
        log->begin("dct_prologue", s); dct_prologue_kernel<<<red_blocks(n), 256, 0, s>>>(pro, t2, n, done); log->end(s);
        src = t2;
    }
    auto rows = [&](const double* a, double* b) {     // a != b
        if (inner.fft) { log->begin("dct_rows_fft", s); fft_launch(false, inner, rgrid, rthreads, dct_smem(inner, rppb), a, b, no, ni, ni, 1, ni, 1, nullptr, nullptr, nosym, 1.0, 0, RowPrologue(), DotEpilogue()); }
        else { log->begin("dct_rows_dense", s); dct_rows_dense_kernel<<<eb, 256, 0, s>>>(a, b, no, ni, inner.denseT, done); }
        log->end(s);
    };
    auto cols = [&](const double* a, double* b) {     // a != b
        if (outer.fft) { log->begin("dct_cols_fft", s); fft_launch(false, outer, cgrid, cthreads, dct_smem(outer, cppb), a, b, ni, no, 1, ni, 1, ni, nullptr, nullptr, nosym, 1.0, 0, RowPrologue(), DotEpilogue()); }
        else { log->begin("dct_cols_dense", s); dct_cols_dense_kernel<<<eb, 256, 0, s>>>(a, b, no, ni, outer.denseT, done); }
        log->end(s);
    };
    rows(src, t1);
    cols(t1, t2);
    log->begin("dct_scale", s); dct_scale_kernel<<<eb, 256, 0, s>>>(t2, no, ni, outer.lam, inner.lam, sym, norm, done); log->end(s);
    if (scale_mode == 1) { log->begin("dct_lambda", s); dct_lambda_kernel<<<eb, 256, 0, s>>>(t2, no, ni, outer.lam, inner.lam, done); log->end(s); }
    cols(t2, t1);
    rows(t1, t2);
    VCH_CUDA(cudaMemcpyAsync(out, t2, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (epi.addend || epi.mul_a) { log->begin("dct_addend", s); dct_addend_kernel<<<red_blocks(n), 256, 0, s>>>(out, epi.addend, epi.mul_a, epi.sc, n, done); log->end(s); }
    if (epi.mode) { log->begin("dct_dots", s); dct_dots_kernel<<<red_blocks(n), kRedThreads, 0, s>>>(out, epi, n, done); log->end(s); }
    VCH_CUDA(cudaGetLastError());
}

inline void DctPlan::destroy() {
    for (DctAxis* ax : {&inner, &outer}) {
        if (ax->tw) cudaFree(ax->tw);
        if (ax->tw16) cudaFree(ax->tw16);
        ax->tw16 = nullptr;
        if (ax->denseT) cudaFree(ax->denseT);
        if (ax->lam) cudaFree(ax->lam);
        ax->tw = nullptr; ax->denseT = nullptr; ax->lam = nullptr;
    }
    tmp1.release(); tmp2.release();
}
#endif   // VCH_CPU_EMU_KERNELS_ONLY

}  // namespace vch
