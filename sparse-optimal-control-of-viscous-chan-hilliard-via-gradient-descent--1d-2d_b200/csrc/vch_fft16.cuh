// Lean DCT-I transform kernels (round 2): radix-16 Stockham FFT, 16 points per thread, everything a compile-time function of
// the FFT length, one kernel instantiation per fused mode.
//
// Why a second kernel family: measured on B200 (profiles/r02_fft_scaling_before.txt, r02_latency_probe.txt) the round-1 kernel
// (vch_dct.cuh, dct_fft_kernel) is INSTRUCTION-ISSUE bound, not HBM-, FP64- or launch-bound: 1300-1850 instructions per thread
// per 2048-point FFT against ~560 of FP64 work (runtime strides and line counts, per-element validity branches, twiddles
// rebuilt from factored tables, one kernel body carrying every prologue / epilogue mode — with spills under its 64-register
// cap); a single CTA needs 4.7 us for its load -> 4 passes -> store chain and every further CTA on the same SM adds 1.8 us.
// Kernel boundaries inside a CUDA graph cost 0.7 us, so nothing is gained by fusing further — the instructions have to go:
//   * 16 points per thread, passes 16 x 16 x {2,4,8,16}: 2 shared-memory exchanges per 2048-point FFT instead of 3;
//   * N, the thread count, the pass structure, the padded shared-memory offsets and the even-extension index reflection are
//     compile-time (the mirrored index of input slot r >= 8 is N - t - (r-8)*tpf: no comparison per element);
//   * twiddles are read from per-pass tables (w^(k r) stored r-major, so a warp reads consecutive entries), not rebuilt;
//   * the fused BiCGStab prologues (6-launch iteration: RowPrologue modes 2 and 3) touch every element once and hand the
//     transform input over through shared memory (even extension applied on the read), the epilogues (DotEpilogue modes 1
//     and 4) are separate instantiations: no mode branch, no spill, restrict-qualified pointers;
//   * the column solve stages its strided 32-byte row segments (4 adjacent columns per CTA) through shared memory in and
//     out, computes the spectral factor for the N+1 distinct spectrum entries only (the mirrored half of the inverse
//     transform's input is read back from shared memory) with one reciprocal per entry.
// The round-1 kernel remains for slab mode (transposing stores / gathering loads over NVLink), for the 7-launch iteration
// (VCH_BICG6=0) and nothing else; grids whose N is not a power of two use the dense-table kernels as before.
#pragma once

namespace vch {

#ifdef VCH_FFT16_TIMING     // scripts/fft_scaling.cu only: clock64 stamps of block 0 / thread 0 at the phase boundaries
__device__ long long vch_dbg_clock[32];
#define VCH_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) vch_dbg_clock[i] = clock64(); } while (0)
#else
#define VCH_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ int padi16(int i) { return i + (i >> 4); }

// L1 is cold at every kernel start, so the first twiddle load of each pass would cost an exposed L2 round trip (measured in
// the column solve: ~2500 cycles for a middle pass that issues in ~700).  Every thread prefetches the table lines it is going to
// read while its first global loads are in flight.
__device__ __forceinline__ void prefetch_l1(const void* p) {
#ifndef VCH_CPU_EMU
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}


template <int LOG2L> struct F16 {
    static_assert(LOG2L >= 6 && LOG2L <= 13, "FFT length 2N with 32 <= N <= 4096");
    static constexpr int Lf = 1 << LOG2L, N = Lf / 2, tpf = Lf / 16;        // tpf: threads per FFT (16 points each)
    static constexpr int rem = LOG2L % 4, n16 = LOG2L / 4;
    static constexpr int lastR = rem == 0 ? 16 : (1 << rem);                  // radix of the last pass
    static constexpr int npass = rem == 0 ? n16 : n16 + 1;
    static constexpr int mids = npass - 2;                                    // radix-16 passes strictly between first and last
    static constexpr int lastNs = Lf / lastR;
    static constexpr int NB = 16 / lastR;                                     // butterflies per thread in the last pass
    static constexpr int ld = Lf + Lf / 16;                                   // padded FFT buffer, double2 entries
    static constexpr bool fold = (tpf % 16) == 0;                             // padded offsets fold into immediates
    static constexpr int tw_mid1 = 0, tw_mid2 = 15 * 16;                        // table offsets (double2 entries) of the middle passes
    static constexpr int tw_last_off = mids == 0 ? 0 : (mids == 1 ? 15 * 16 : 15 * 16 + 15 * 256);
    static constexpr int tw_total = tw_last_off + (lastR - 1) * lastNs;
    // row kernels: fpb line pairs per CTA (>= 128 threads);  column kernel: cp column pairs per CTA
    static constexpr int fpb = tpf >= 128 ? 1 : 128 / tpf;
    static constexpr int rthreads = fpb * tpf;
    // cp: as many adjacent columns per CTA as shared memory allows — the strided row segments of the pitched buffer are what
    // limits the column kernel (measured: one 32-byte request per ~3.2 cycles per SM), and a segment of 2*cp doubles is one request
    static constexpr int cp = LOG2L <= 11 ? 4 : (LOG2L == 12 ? 2 : 1);
    static constexpr int cthreads = cp * tpf;
    static constexpr int sst = N + 2;                                         // stage stride per column pair (N + 2 = 2 mod 8: conflict-free staging stores)
    static constexpr int rminb = 512 / rthreads > 4 ? 4 : (512 / rthreads > 0 ? 512 / rthreads : 1);
    static constexpr int cminb = 512 / cthreads > 4 ? 4 : (512 / cthreads > 0 ? 512 / cthreads : 1);
    static constexpr size_t rows_smem_plain = sizeof(double2) * (size_t)fpb * ld;
    static constexpr size_t rows_smem_staged = sizeof(double2) * (size_t)fpb * (ld + N + 1);
    static constexpr size_t cols_smem_bytes = sizeof(double2) * (size_t)cp * (ld + sst);
};

// Barrier among the tpf threads of ONE FFT (named barrier 1 + f) when they are whole warps, else the CTA barrier.  The FFT
// groups of a CTA share nothing between the staging steps, so they need not march in lockstep: with their own barriers one group's
// butterflies overlap another group's shared-memory traffic.  (CPU emulation: every group executes the same barrier sequence,
// so the CTA barrier is a valid stand-in.)
template <int LOG2L, bool GROUP>
__device__ __forceinline__ void fft16_sync(int f) {
#ifndef VCH_CPU_EMU
    if (GROUP) { asm volatile("bar.sync %0, %1;" ::"r"(f + 1), "n"(F16<LOG2L>::tpf) : "memory"); return; }
#endif
    (void)f;
    __syncthreads();
}

template <int LOG2L>
__device__ __forceinline__ void fft16_prefetch_twiddles(const double2* __restrict__ tw, int t) {
    using G = F16<LOG2L>;
    // a 128-byte line holds 8 entries: one prefetch per line, spread over the threads of the FFT
    constexpr int lines = (G::tw_total + 7) / 8;
    for (int l = t; l < lines; l += G::tpf) prefetch_l1(tw + 8 * l);
}

// first-pass input slot r of thread t  <->  element t + r*tpf of the even extension  <->  stored element idx16(t, r) <= N
template <int LOG2L> __device__ __forceinline__ int idx16(int t, int r) {
    using G = F16<LOG2L>;
    return r < 8 ? t + r * G::tpf : G::N - t - (r - 8) * G::tpf;
}

// ---- 16-point DFT (forward sign), 4 x 4:  n = c + 4m, k = r + 4s:  w16^(nk) = w4^(mr) w16^(cr) w4^(cs)
template <> __device__ __forceinline__ void dft<16>(double2 (&v)[16]) {
    const double C = 0.92387953251128675613, S = 0.38268343236508977173, h = 0.70710678118654752440;
    double2 a[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double2 x[4] = {v[c], v[c + 4], v[c + 8], v[c + 12]};
        dft<4>(x);
#pragma unroll
        for (int r = 0; r < 4; ++r) a[c][r] = x[r];
    }
    // multiply a[c][r] by w16^(c r), w16 = exp(-i pi/8)
    auto w1 = [&](double2 z) { return make_double2(z.x * C + z.y * S, z.y * C - z.x * S); };      // (C, -S)
    auto w2 = [&](double2 z) { return make_double2(h * (z.x + z.y), h * (z.y - z.x)); };          // (h, -h)
    auto w3 = [&](double2 z) { return make_double2(z.x * S + z.y * C, z.y * S - z.x * C); };      // (S, -C)
    auto w6 = [&](double2 z) { return make_double2(h * (z.y - z.x), -h * (z.x + z.y)); };         // (-h, -h)
    auto w9 = [&](double2 z) { return make_double2(-(z.x * C + z.y * S), z.x * S - z.y * C); };   // (-C, S)
    a[1][1] = w1(a[1][1]); a[1][2] = w2(a[1][2]); a[1][3] = w3(a[1][3]);
    a[2][1] = w2(a[2][1]); a[2][2] = mul_mi(a[2][2]); a[2][3] = w6(a[2][3]);
    a[3][1] = w3(a[3][1]); a[3][2] = w6(a[3][2]); a[3][3] = w9(a[3][3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double2 y[4] = {a[0][r], a[1][r], a[2][r], a[3][r]};
        dft<4>(y);
#pragma unroll
        for (int s = 0; s < 4; ++s) v[r + 4 * s] = y[s];
    }
}

// ---- passes.  Stockham autosort: butterfly j of a radix-R pass with Ns = product of the earlier radices reads
// j + r*(Lf/R), multiplies by w^(k r Lf/(Ns R)), k = j mod Ns, and writes (j - k) R + k + r Ns.
template <int LOG2L, bool GROUP = false>
__device__ __forceinline__ void fft16_first_store(double2* data, double2 (&v)[16], int t, int f = 0) {
    dft<16>(v);
    double2* dst = data + 17 * t;                 // padi16(16 t + r) = 17 t + r
#pragma unroll
    for (int r = 0; r < 16; ++r) dst[r] = v[r];
    fft16_sync<LOG2L, GROUP>(f);
}

// Twiddles w^r, r = 1..R-1, of one butterfly: the powers 1, 2, 4, 8 come from the per-pass table (r-major, so a warp reads
// consecutive entries), the rest are products of at most three of them.  Loading all of them cost more than it saved: the
// kernels are bound by the load/store pipe (ncu: FP64 pipe ~50 %), and every table entry is a 16-byte L1 access.
template <int R>
__device__ __forceinline__ void fft16_twiddles(double2 (&w)[R - 1], const double2* __restrict__ tab, int ns, int k) {
    w[0] = __ldg(&tab[k]);
    if (R >= 4) { w[1] = __ldg(&tab[ns + k]); w[2] = cmul(w[0], w[1]); }
    if (R >= 8) { w[3] = __ldg(&tab[3 * ns + k]); w[4] = cmul(w[0], w[3]); w[5] = cmul(w[1], w[3]); w[6] = cmul(w[2], w[3]); }
    if (R >= 16) {
        w[7] = __ldg(&tab[7 * ns + k]);
#pragma unroll
        for (int r = 0; r < 7; ++r) w[8 + r] = cmul(w[r], w[7]);
    }
}

template <int LOG2L, int NS, bool GROUP>
__device__ __forceinline__ void fft16_mid_pass(double2* data, int t, const double2* __restrict__ twp, int f) {
    using G = F16<LOG2L>;
    static_assert(G::fold, "a middle pass exists only for Lf >= 512");
    double2 v[16], w[15];
    const int k = t & (NS - 1);
    fft16_twiddles<16>(w, twp, NS, k);
    const double2* src = data + padi16(t);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = src[r * (G::tpf + G::tpf / 16)];
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], w[r - 1]);
    dft<16>(v);
    fft16_sync<LOG2L, GROUP>(f);
    double2* dst = data + padi16((t - k) * 16 + k);
#pragma unroll
    for (int r = 0; r < 16; ++r) dst[r * (NS + NS / 16)] = v[r];
    fft16_sync<LOG2L, GROUP>(f);
}

template <int LOG2L, bool GROUP = false>
__device__ __forceinline__ void fft16_middle(double2* data, int t, const double2* __restrict__ tw, int f = 0) {
    using G = F16<LOG2L>;
    if constexpr (G::mids >= 1) fft16_mid_pass<LOG2L, 16, GROUP>(data, t, tw + G::tw_mid1, f);
    if constexpr (G::mids >= 2) fft16_mid_pass<LOG2L, 256, GROUP>(data, t, tw + G::tw_mid2, f);
}

// Last pass: output slot q = m + r*NB of thread t is element t + q*tpf in natural order — the same convention as the
// first-pass input slots, so a following transform (column solve) takes the registers as they are.
template <int LOG2L>
__device__ __forceinline__ void fft16_last_pass(const double2* data, int t, const double2* __restrict__ tw, double2 (&z)[16]) {
    using G = F16<LOG2L>;
    constexpr int R = G::lastR, NB = G::NB;
    const double2* twl = tw + G::tw_last_off;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * G::tpf;
        double2 v[R], w[R - 1];
        fft16_twiddles<R>(w, twl, G::lastNs, j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = m + r * NB;
            v[r] = G::fold ? data[padi16(t) + q * (G::tpf + G::tpf / 16)] : data[padi16(t + q * G::tpf)];
        }
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r - 1]);
        dft<R>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) z[m + r * NB] = v[r];
    }
}

// ---- fused BiCGStab prologues of the 6-launch iteration.  Every element of the two lines is touched by exactly one thread;
// the transform input goes to stage[e] = (line a, line b).  Pointers are restrict-qualified parameters: each vector that is
// written is distinct from every vector that is only read (enqueue_bicg_iteration); x and p are updated in place.
template <int LOG2L, bool MUL>
__device__ __forceinline__ void pro16_mode2(double2* __restrict__ stage, const double* __restrict__ r, const double* __restrict__ vv,
                                            const double* __restrict__ a, double* __restrict__ s, double al, double abar,
                                            size_t ba, size_t bb, bool va, bool vb, int t) {
    using G = F16<LOG2L>;
    auto elem = [&](int e) {       // s = r - alpha v;  x_in = (a - abar) s
        const double ra = r[ba + e], rb = r[bb + e], qa = vv[ba + e], qb = vv[bb + e];
        const double aa = MUL ? a[ba + e] : 0.0, ab = MUL ? a[bb + e] : 0.0;
        const double sa = ra - al * qa, sb = rb - al * qb;
        if (va) s[ba + e] = sa;
        if (vb) s[bb + e] = sb;
        stage[e] = make_double2(va ? (MUL ? (aa - abar) * sa : sa) : 0.0, vb ? (MUL ? (ab - abar) * sb : sb) : 0.0);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) elem(t + q * G::tpf);
    if (t == 0) elem(G::N);
}
template <int LOG2L, bool MUL>
__device__ __forceinline__ void pro16_mode3_first(double2* __restrict__ stage, const double* __restrict__ r, const double* __restrict__ a,
                                                  double* __restrict__ p, double abar, size_t ba, size_t bb, bool va, bool vb, int t) {
    using G = F16<LOG2L>;
    auto elem = [&](int e) {       // first iteration of a solve: p = r (x = 0 already)
        const double pa = r[ba + e], pb = r[bb + e];
        const double aa = MUL ? a[ba + e] : 0.0, ab = MUL ? a[bb + e] : 0.0;
        if (va) p[ba + e] = pa;
        if (vb) p[bb + e] = pb;
        stage[e] = make_double2(va ? (MUL ? (aa - abar) * pa : pa) : 0.0, vb ? (MUL ? (ab - abar) * pb : pb) : 0.0);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) elem(t + q * G::tpf);
    if (t == 0) elem(G::N);
}
template <int LOG2L, bool MUL>
__device__ __forceinline__ void pro16_mode3(double2* __restrict__ stage, const double* __restrict__ s, const double* __restrict__ tt,
                                            const double* __restrict__ vv, const double* __restrict__ a, double* __restrict__ p,
                                            double* __restrict__ x, double* __restrict__ rw, double al, double om, double beta,
                                            double abar, size_t ba, size_t bb, bool va, bool vb, int t) {
    using G = F16<LOG2L>;
    auto elem = [&](int e) {       // r = s - omega t;  x += alpha p + omega s;  p = r + beta (p - omega v);  x_in = (a - abar) p
        const double sa = s[ba + e], sb = s[bb + e], ta = tt[ba + e], tb = tt[bb + e];
        const double pa = p[ba + e], pb = p[bb + e], qa = vv[ba + e], qb = vv[bb + e];
        const double xa = x[ba + e], xb = x[bb + e];
        const double aa = MUL ? a[ba + e] : 0.0, ab = MUL ? a[bb + e] : 0.0;
        const double ra = sa - om * ta, rb = sb - om * tb;
        const double na = ra + beta * (pa - om * qa), nb = rb + beta * (pb - om * qb);
        if (va) { x[ba + e] = xa + (al * pa + om * sa); rw[ba + e] = ra; p[ba + e] = na; }
        if (vb) { x[bb + e] = xb + (al * pb + om * sb); rw[bb + e] = rb; p[bb + e] = nb; }
        stage[e] = make_double2(va ? (MUL ? (aa - abar) * na : na) : 0.0, vb ? (MUL ? (ab - abar) * nb : nb) : 0.0);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) elem(t + q * G::tpf);
    if (t == 0) elem(G::N);
}

// Start of a forward Newton solve: the Schur right-hand side b = L R_phi - R_mu (schur_rhs_kernel's formula, same operation
// order as lap_g) formed while loading, so that kernel and the round trip of b through memory disappear.  Lines la, lb of an
// nlines x (N+1) field; the line neighbours of la / lb are lb / la themselves plus one more line each.
template <int LOG2L>
__device__ __forceinline__ void pro16_mode4(double2* __restrict__ stage, const double* __restrict__ rp, const double* __restrict__ rm,
                                            double k_in, double k_out, int la, int lb, int nlines, int ls, bool vb, int t) {
    using G = F16<LOG2L>;
    const int lbb = vb ? lb : la;                                                   // absent second line: staged as zero
    const size_t ba = (size_t)la * ls, bb = (size_t)lbb * ls;
    const size_t bam = (size_t)(la > 0 ? la - 1 : 1) * ls;                          // line below la
    const size_t bap = (size_t)(la < nlines - 1 ? la + 1 : nlines - 2) * ls;        // line above la (= lb when present)
    const size_t bbm = (size_t)(lbb > 0 ? lbb - 1 : 1) * ls;
    const size_t bbp = (size_t)(lbb < nlines - 1 ? lbb + 1 : nlines - 2) * ls;
    auto elem = [&](int e) {
        const int em = e > 0 ? e - 1 : 1, ep = e < G::N ? e + 1 : G::N - 1;
        const double ca = rp[ba + e], cb = rp[bb + e];
        const double a_in = (rp[ba + ep] - ca) + (rp[ba + em] - ca), a_out = (rp[bap + e] - ca) + (rp[bam + e] - ca);
        const double b_in = (rp[bb + ep] - cb) + (rp[bb + em] - cb), b_out = (rp[bbp + e] - cb) + (rp[bbm + e] - cb);
        stage[e] = make_double2((a_in * k_in + a_out * k_out) - rm[ba + e], vb ? (b_in * k_in + b_out * k_out) - rm[bb + e] : 0.0);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) elem(t + q * G::tpf);
    if (t == 0) elem(G::N);
}

// ---- fused epilogue: out = [(mul_a - abar)] z + addend and the BiCGStab dot products (DotEpilogue modes 1 and 4)
template <int LOG2L, int EPI, bool MUL, int XM>
__device__ __forceinline__ void epi16_store(const double2 (&z)[16], double* __restrict__ out, const double* __restrict__ addend,
                                            const double* __restrict__ other, const double* __restrict__ rvec,
                                            const double* __restrict__ mul_a, double abar, size_t ba, size_t bb, bool va, bool vb,
                                            int t, double (&acc)[5], const Scatter& sct, int la, double* __restrict__ out2 = nullptr,
                                            double* __restrict__ zero = nullptr) {
    using G = F16<LOG2L>;
    const bool has_r = EPI == 4 || rvec != nullptr;
    auto elem = [&](int e, double2 zz) {
        if (EPI == 0 && XM == 1) {   // slab mode: transposing store into the buffer of the rank that owns column e
            int r = e >> sct.shift; if (r >= sct.nr) r = sct.nr - 1;
            double* dst = sct.peer[r] + sct.off + (size_t)(sct.base + la) * sct.pitch + (e - (r << sct.shift));
            if (va) dst[0] = zz.x;
            if (vb) dst[sct.pitch] = zz.y;
            return;
        }
        if (EPI == 0) {
            if (va) out[ba + e] = zz.x;
            if (vb) out[bb + e] = zz.y;
            return;
        }
        if (EPI == 5) {   // start of a solve: r = r0 = z, x = 0, (r,r)
            if (va) { out[ba + e] = zz.x; out2[ba + e] = zz.x; zero[ba + e] = 0.0; acc[0] += zz.x * zz.x; }
            if (vb) { out[bb + e] = zz.y; out2[bb + e] = zz.y; zero[bb + e] = 0.0; acc[0] += zz.y * zz.y; }
            return;
        }
        const double da = addend[ba + e], db = addend[bb + e], oa = other[ba + e], ob = other[bb + e];
        const double ma = MUL ? mul_a[ba + e] : 0.0, mb = MUL ? mul_a[bb + e] : 0.0;
        const double ra = has_r ? rvec[ba + e] : 0.0, rb = has_r ? rvec[bb + e] : 0.0;
        const double xo = (MUL ? (ma - abar) * zz.x : zz.x) + da, yo = (MUL ? (mb - abar) * zz.y : zz.y) + db;
        if (va) {
            out[ba + e] = xo;
            acc[0] += oa * xo; acc[1] += xo * xo; acc[2] += ra * xo;
            if (EPI == 4) { acc[3] += oa * oa; acc[4] += ra * oa; }
        }
        if (vb) {
            out[bb + e] = yo;
            acc[0] += ob * yo; acc[1] += yo * yo; acc[2] += rb * yo;
            if (EPI == 4) { acc[3] += ob * ob; acc[4] += rb * ob; }
        }
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) elem(t + q * G::tpf, z[q]);
    if (t == 0) elem(G::N, z[8]);
}

// ---- row kernel: one CTA = fpb line pairs; line l at base + l*ls, contiguous elements.
//   PRO 0: x = in;  2: s = r - alpha v;  3: deferred x/r update + new p;  4: x = L R_phi - R_mu (Schur right-hand side)  (RowPrologue);
//       MUL: multiply by (a - abar) on this side
//   EPI 0: plain store;  1 / 4: addend + dot products;  5: start of a solve (r = r0 = out, x = 0, (r,r))  (DotEpilogue)
//   XM (slab mode, Scatter in vch_dct.cuh): 1 = the plain store goes transposed into the column owners' buffers over NVLink;
//       3 = the input is gathered from the column owners' buffers (every element crosses NVLink once, staged like a prologue)
template <int LOG2L, int PRO, int EPI, bool MUL, int XM = 0>
__global__ void __launch_bounds__(F16<LOG2L>::rthreads, F16<LOG2L>::rminb)
rows16_kernel(const double* __restrict__ in, double* __restrict__ out, int nlines, int in_ls, int out_ls,
              const double2* __restrict__ tw, RowPrologue pro, DotEpilogue epi, const int* __restrict__ done,
              const __grid_constant__ Scatter sct) {
    pdl_enter();
    using G = F16<LOG2L>;
    if (done && *done) {
        // A solve that is already finished when its loop body starts (zero right-hand side with the start fused into the kernel
        // in front of the graph): the body's last kernel must still clear the WHILE condition, nobody else will.
        if (EPI == 4 && epi.use_cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(epi.cond, 0u);
        return;
    }
#ifdef VCH_CPU_EMU
    double2* sm = reinterpret_cast<double2*>(vch_emu::dynamic_smem());
#else
    extern __shared__ double2 sm[];
#endif
    const int f = G::fpb == 1 ? 0 : threadIdx.x / G::tpf, t = threadIdx.x - f * G::tpf;
    double2* data = sm + (size_t)f * G::ld;
    const int la = 2 * (blockIdx.x * G::fpb + f), lb = la + 1;
    const bool va = la < nlines, vb = lb < nlines;
    // Absent lines (the odd line out — N + 1 lines is always odd — and idle FFT slots of the last CTA) load line 0 so that no branch
    // separates the loads, are never written, and enter the transform as ZEROS.  Not as what was loaded: the deferred update of
    // mode 3 rewrites x and p of line 0 in place in another CTA at the same time, the value read here depends on which CTA runs
    // first, and the imaginary input of the shared complex FFT leaks into the real line at rounding level — results were
    // reproducible only to ~1e-16 per application (measured: scripts/determinism_probe.py; 1e-11 after a few time steps at 1024^2).
    const size_t ia = (size_t)(va ? la : 0) * in_ls, ib = (size_t)(vb ? lb : 0) * in_ls;
    double2 v[16];
    if (PRO == 0 && XM == 3) {
        double2* stage = sm + (size_t)G::fpb * G::ld + (size_t)f * (G::N + 1);
        auto fetch = [&](int e) {
            int r = e >> sct.shift; if (r >= sct.nr) r = sct.nr - 1;
            const double* src = sct.peer[r] + sct.off + (size_t)(sct.base + (va ? la : 0)) * sct.pitch + (e - (r << sct.shift));
            stage[e] = make_double2(src[0], vb ? src[sct.pitch] : 0.0);
        };
#pragma unroll
        for (int q = 0; q < 8; ++q) fetch(t + q * G::tpf);
        if (t == 0) fetch(G::N);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = stage[idx16<LOG2L>(t, r)];
    } else if (PRO == 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r) { const int e = idx16<LOG2L>(t, r); v[r] = make_double2(va ? in[ia + e] : 0.0, vb ? in[ib + e] : 0.0); }
    } else {
        double2* stage = sm + (size_t)G::fpb * G::ld + (size_t)f * (G::N + 1);
        const Scal* sc = pro.sc;
        const double al = PRO == 4 ? 0.0 : sc->alpha, om = PRO == 4 ? 0.0 : sc->omega, abar = PRO == 4 ? 0.0 : sc->abar;
        if (PRO == 4) pro16_mode4<LOG2L>(stage, pro.r, pro.qv, pro.k_in, pro.k_out, va ? la : 0, lb, nlines, in_ls, vb, t);
        else if (PRO == 2) pro16_mode2<LOG2L, MUL>(stage, pro.r, pro.qv, pro.a, pro.w, al, abar, ia, ib, va, vb, t);
        else if (sc->iters == 0) pro16_mode3_first<LOG2L, MUL>(stage, pro.r, pro.a, pro.w, abar, ia, ib, va, vb, t);
        else pro16_mode3<LOG2L, MUL>(stage, pro.s, pro.t, pro.qv, pro.a, pro.w, pro.x, pro.rw, al, om, (sc->rho_new / sc->rho) * (al / om),
                                      abar, ia, ib, va, vb, t);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = stage[idx16<LOG2L>(t, r)];
    }
    fft16_first_store<LOG2L>(data, v, t);
    fft16_middle<LOG2L>(data, t, tw);
    double2 z[16];
    fft16_last_pass<LOG2L>(data, t, tw, z);

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const size_t oa = (size_t)(va ? la : 0) * out_ls, ob = (size_t)(vb ? lb : 0) * out_ls;
    epi16_store<LOG2L, EPI, MUL, XM>(z, out, epi.addend, epi.other, epi.rvec, epi.mul_a, MUL && EPI ? epi.sc->abar : 0.0, oa, ob, va, vb, t, acc,
                                     sct, la, epi.out2, epi.zero);
    if (EPI == 5) {
        double vals[1] = {acc[0]};
        const int op[1] = {0};
        double tot[1];
        if (grid_reduce<1>(vals, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) solve_init_scalars(epi.sc, tot[0], epi.cond, epi.use_cond);
    } else if (EPI == 4) {          // every thread of every CTA takes part in the reduction
        const int op[5] = {0, 0, 0, 0, 0};
        double tot[5];
        if (grid_reduce<5>(acc, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) dots_finish6(epi, tot);
    } else if (EPI == 1) {
        double vals[3] = {acc[0], acc[1], acc[2]};
        const int op[3] = {0, 0, 0};
        double tot[3];
        if (grid_reduce<3>(vals, op, epi.part, epi.ticket, tot) && threadIdx.x == 0) dots_finish(epi, tot);
    }
}

// ---- column solve, in place on the pitched buffer: forward transform -> spectral factor -> inverse transform.
// One CTA = cp column pairs = 2*cp adjacent columns (64-byte row segments for cp = 4; the buffer's pitch is a multiple of 8 and
// its padding columns hold zeros, so no column is ever absent).  A pair (c, c+1) is one complex FFT: Re = column c.
template <int LOG2L>
__global__ void __launch_bounds__(F16<LOG2L>::cthreads, F16<LOG2L>::cminb)
cols16_kernel(double* __restrict__ buf, int pitch, int ncols, const double2* __restrict__ tw, const double* __restrict__ lam_col,
              const double* __restrict__ lam_row, SymbolArgs sy, double norm, int scale_mode, const int* __restrict__ done) {
    pdl_enter();
    using G = F16<LOG2L>;
    constexpr int CP = G::cp, N = G::N;
    if (done && *done) return;
#ifdef VCH_CPU_EMU
    double2* sm = reinterpret_cast<double2*>(vch_emu::dynamic_smem());
#else
    extern __shared__ double2 sm[];
#endif
    const int f = CP == 1 ? 0 : threadIdx.x / G::tpf, t = threadIdx.x - f * G::tpf;
    double2* data = sm + (size_t)f * G::ld;
    double2* stage_all = sm + (size_t)CP * G::ld;                 // [CP][sst]
    double2* stage = stage_all + (size_t)f * G::sst;
    const int c0 = 2 * CP * blockIdx.x;                           // first column of this CTA
    double* base = buf + c0;
    VCH_STAMP(0);
    // stage in: CP consecutive threads fetch the CP double2 of one row segment (2*CP doubles, contiguous), every segment once;
    // the trip count is compile-time so that all loads are in flight together
    constexpr int NSEG = (N + 1) * CP, NIT = (NSEG + G::cthreads - 1) / G::cthreads;
    // Every CTA walks its strip in a different row order (rotation by blockIdx): with all CTAs starting at row 0 the whole grid
    // requests pieces of the SAME few rows at the same time, i.e. a handful of L2 slices serve all SMs.
#ifndef VCH_COLS_NO_ROT
    const int rot = (int)(((long long)blockIdx.x * (N + 1)) / gridDim.x);
#else
    const int rot = 0;
#endif
    auto seg_row = [&](int idx) { int r = idx / CP + rot; return r > N ? r - (N + 1) : r; };
    {
        double2 p[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int idx = threadIdx.x + i * G::cthreads;
            if (idx < NSEG) p[i] = *reinterpret_cast<const double2*>(base + (size_t)seg_row(idx) * pitch + 2 * (idx % CP));
        }
        fft16_prefetch_twiddles<LOG2L>(tw, t);
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int idx = threadIdx.x + i * G::cthreads;
            if (idx < NSEG) stage_all[(idx % CP) * G::sst + seg_row(idx)] = p[i];
        }
    }
    // eigenvalues of the 9 spectrum entries this thread scales (slots 0..7, slot 8 for thread 0), fetched early
    double lrow[9];
#pragma unroll
    for (int q = 0; q < 8; ++q) lrow[q] = lam_row[t + q * G::tpf];
    lrow[8] = lam_row[N];
    __syncthreads();
    VCH_STAMP(1);
    double2 v[16], z[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = stage[idx16<LOG2L>(t, r)];
    constexpr bool GS = CP > 1 && (G::tpf % 32) == 0;           // per-FFT barriers between the CTA-wide staging steps
    fft16_first_store<LOG2L, GS>(data, v, t, f);                 // (its barrier also ends this group's reads of its stage)
    VCH_STAMP(2);
    fft16_middle<LOG2L, GS>(data, t, tw, f);
    VCH_STAMP(3);
    fft16_last_pass<LOG2L>(data, t, tw, z);
    VCH_STAMP(4);
    {   // spectral factor on the N+1 distinct entries (slots 0..7, and slot 8 of thread 0); the mirrored half of the next
        // transform's input is read back from the stage, which also makes the spectrum exactly even
        const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
        const double k0 = sy.coef_ptr ? sy.coef_ptr[0] : sy.c0, k2 = sy.coef_ptr ? sy.coef_ptr[1] : sy.c2;
        const int ca = c0 + 2 * f, cb = ca + 1;
        const double lca = lam_col[ca < ncols ? ca : ncols - 1], lcb = lam_col[cb < ncols ? cb : ncols - 1];
        auto elem = [&](int e, double le, double2 zz) {
            const double s1 = le + lca, s2 = le + lcb;
            const double d1 = k0 + s1 * (abar + k2 * s1), d2 = k0 + s2 * (abar + k2 * s2);
            const double inv = norm * __drcp_rn(d1 * d2);        // one reciprocal for both columns
            double f1 = inv * d2, f2 = inv * d1;
            if (scale_mode == 1) { f1 *= s1; f2 *= s2; }
            stage[e] = make_double2(zz.x * f1, zz.y * f2);
        };
#pragma unroll
        for (int q = 0; q < 8; ++q) elem(t + q * G::tpf, lrow[q], z[q]);
        if (t == 0) elem(N, lrow[8], z[8]);
    }
    fft16_sync<LOG2L, GS>(f);          // also: every thread of the group has finished reading `data` in the last pass
    VCH_STAMP(5);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = stage[idx16<LOG2L>(t, r)];
    fft16_first_store<LOG2L, GS>(data, v, t, f);
    fft16_middle<LOG2L, GS>(data, t, tw, f);
    fft16_last_pass<LOG2L>(data, t, tw, z);
    VCH_STAMP(6);
    {   // Padding columns (>= ncols) are written back as exact zeros.  A padding column shares its complex FFT with the last real
        // column; the split of the two real transforms is exact only up to rounding, so it would otherwise collect ~1e-16 of its
        // neighbour, keep it (no row kernel writes there) and feed ~1e-32 of it back on the next application: harmless, but
        // results would depend on the context's history and a zero right-hand side would not give an exactly zero solution.
        const bool za = c0 + 2 * f >= ncols, zb = c0 + 2 * f + 1 >= ncols;
#pragma unroll
        for (int q = 0; q < 9; ++q) { if (za) z[q].x = 0.0; if (zb) z[q].y = 0.0; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) stage[t + q * G::tpf] = z[q];
    if (t == 0) stage[N] = z[8];
    __syncthreads();
    VCH_STAMP(7);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int idx = threadIdx.x + i * G::cthreads;
        if (idx < NSEG) *reinterpret_cast<double2*>(base + (size_t)seg_row(idx) * pitch + 2 * (idx % CP)) = stage_all[(idx % CP) * G::sst + seg_row(idx)];
    }
    VCH_STAMP(8);
}

// ---- column solve with TMA staging (sm_100a; not compiled for the CPU emulation).
// Same transform chain as cols16_kernel; what changes is how the strip of 2*cp columns travels between the pitched buffer and shared
// memory.  Measured (profiles/r02_fft_scaling_*.txt): the per-thread 16-byte loads of 16*cp-byte row segments cost 9050 of the
// kernel's 28200 cycles at 1024^2 with the full grid (3700 alone) — half-line requests exhaust the SM's outstanding-load slots long
// before its bandwidth — and the stores another 2400.  Here one thread issues nb bulk tensor copies (boxes of BR rows x 2*cp
// columns, cp.async.bulk.tensor.2d) that land on an mbarrier, and the result leaves the same way; the hardware swizzle of the
// tensor map (16-byte chunk index XOR row bits: SWIZZLE_64B for 64-byte rows, SWIZZLE_32B for 32-byte rows) makes the
// column-pair reads and writes of the stage conflict-free without padding.  Rows beyond N (the last box) are zero-filled on
// load and clipped on store.
#ifndef VCH_CPU_EMU
template <int LOG2L> struct F16T {
    using G = F16<LOG2L>;
    static constexpr int CP = G::cp, CPB = CP == 4 ? 2 : (CP == 2 ? 1 : 0);
    static constexpr int rows = G::N + 1;
    static constexpr int nb = (rows + 255) / 256;                              // boxes per strip (box height <= 256)
    static constexpr int BR = (((rows + nb - 1) / nb) + 7) & ~7;               // box height: multiple of 8 rows -> box bases stay swizzle-aligned
    static constexpr int row_bytes = 16 * CP;
    static constexpr unsigned box_bytes = (unsigned)BR * row_bytes;
    static constexpr size_t stage_bytes = (size_t)nb * box_bytes;              // multiple of 128
    static constexpr size_t data_off = (stage_bytes + 1023) & ~size_t(1023);
    static constexpr size_t bar_off = data_off + sizeof(double2) * (size_t)CP * G::ld;
    static constexpr size_t smem_bytes = bar_off + 16;
    static constexpr bool use = LOG2L >= 9;                                    // N >= 256; smaller strips keep the plain kernel
    // byte offset of entry e of column pair f in the swizzled stage
    __device__ static __forceinline__ unsigned off(int e, int f) {
        return (unsigned)e * row_bytes + (unsigned)((f ^ ((e >> (3 - CPB)) & (CP - 1))) << 4);
    }
};
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int LOG2L>
__global__ void __launch_bounds__(F16<LOG2L>::cthreads, 1)
cols16_tma_kernel(const __grid_constant__ CUtensorMap tmap, int ncols, const double2* __restrict__ tw, const double* __restrict__ lam_col,
                  const double* __restrict__ lam_row, SymbolArgs sy, double norm, int scale_mode, const int* __restrict__ done) {
    pdl_enter();
    using G = F16<LOG2L>;
    using T = F16T<LOG2L>;
    constexpr int CP = G::cp, N = G::N;
    if (done && *done) return;
    extern __shared__ __align__(1024) unsigned char smraw[];
    unsigned char* stage = smraw;
    double2* data_all = reinterpret_cast<double2*>(smraw + T::data_off);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smraw + T::bar_off);
    const int f = CP == 1 ? 0 : threadIdx.x / G::tpf, t = threadIdx.x - f * G::tpf;
    double2* data = data_all + (size_t)f * G::ld;
    const int c0 = 2 * CP * blockIdx.x;
    const unsigned bar_a = smem_u32(bar), stage_a = smem_u32(stage);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((unsigned)T::stage_bytes) : "memory");
#pragma unroll
        for (int k = 0; k < T::nb; ++k) {
            const int b = (k + (int)blockIdx.x) % T::nb;          // CTAs start on different boxes: fewer CTAs on the same rows at once
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(stage_a + b * T::box_bytes), "l"(&tmap), "r"(c0), "r"(b * T::BR), "r"(bar_a) : "memory");
        }
    }
    fft16_prefetch_twiddles<LOG2L>(tw, t);
    double lrow[9];
#pragma unroll
    for (int q = 0; q < 8; ++q) lrow[q] = lam_row[t + q * G::tpf];
    lrow[8] = lam_row[N];
    __syncthreads();                                    // the barrier is initialised before anybody polls it
    {
        unsigned ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar_a) : "memory");
    }
    auto st = [&](int e) -> double2& { return *reinterpret_cast<double2*>(stage + T::off(e, f)); };
    double2 v[16], z[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = st(idx16<LOG2L>(t, r));
    constexpr bool GS = CP > 1 && (G::tpf % 32) == 0;
    fft16_first_store<LOG2L, GS>(data, v, t, f);
    fft16_middle<LOG2L, GS>(data, t, tw, f);
    fft16_last_pass<LOG2L>(data, t, tw, z);
    {
        const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
        const double k0 = sy.coef_ptr ? sy.coef_ptr[0] : sy.c0, k2 = sy.coef_ptr ? sy.coef_ptr[1] : sy.c2;
        const int ca = c0 + 2 * f, cb = ca + 1;
        const double lca = lam_col[ca < ncols ? ca : ncols - 1], lcb = lam_col[cb < ncols ? cb : ncols - 1];
        auto elem = [&](int e, double le, double2 zz) {
            const double s1 = le + lca, s2 = le + lcb;
            const double d1 = k0 + s1 * (abar + k2 * s1), d2 = k0 + s2 * (abar + k2 * s2);
            const double inv = norm * __drcp_rn(d1 * d2);
            double f1 = inv * d2, f2 = inv * d1;
            if (scale_mode == 1) { f1 *= s1; f2 *= s2; }
            st(e) = make_double2(zz.x * f1, zz.y * f2);
        };
#pragma unroll
        for (int q = 0; q < 8; ++q) elem(t + q * G::tpf, lrow[q], z[q]);
        if (t == 0) elem(N, lrow[8], z[8]);
    }
    fft16_sync<LOG2L, GS>(f);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = st(idx16<LOG2L>(t, r));
    fft16_first_store<LOG2L, GS>(data, v, t, f);
    fft16_middle<LOG2L, GS>(data, t, tw, f);
    fft16_last_pass<LOG2L>(data, t, tw, z);
    {   // padding columns leave as exact zeros (see cols16_kernel)
        const bool za = c0 + 2 * f >= ncols, zb = c0 + 2 * f + 1 >= ncols;
#pragma unroll
        for (int q = 0; q < 9; ++q) { if (za) z[q].x = 0.0; if (zb) z[q].y = 0.0; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) st(t + q * G::tpf) = z[q];
    if (t == 0) st(N) = z[8];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of the stage -> visible to the bulk copies
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < T::nb; ++b)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&tmap), "r"(c0), "r"(b * T::BR), "r"(stage_a + b * T::box_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
#endif   // !VCH_CPU_EMU

// host: launch / attribute helpers.  XM = 0: single GPU; 1 / 3: slab mode (forward rows scatter, inverse rows gather).
#ifndef VCH_CPU_EMU_KERNELS_ONLY
template <typename K>
static inline void fft16_attr(K kern, size_t bytes) {
    VCH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}
template <int LG, int XM>
static inline void rows16_set_attributes() {
    using G = F16<LG>;
    if (XM != 3) {
        fft16_attr(rows16_kernel<LG, 0, 0, false, XM>, G::rows_smem_plain);
        fft16_attr(rows16_kernel<LG, 2, 0, false, XM>, G::rows_smem_staged); fft16_attr(rows16_kernel<LG, 2, 0, true, XM>, G::rows_smem_staged);
        fft16_attr(rows16_kernel<LG, 3, 0, false, XM>, G::rows_smem_staged); fft16_attr(rows16_kernel<LG, 3, 0, true, XM>, G::rows_smem_staged);
        if (XM == 0) fft16_attr(rows16_kernel<LG, 4, 0, false, 0>, G::rows_smem_staged);
    }
    if (XM != 1) {
        constexpr size_t sm = XM == 3 ? G::rows_smem_staged : G::rows_smem_plain;
        if (XM == 3) fft16_attr(rows16_kernel<LG, 0, 0, false, XM>, sm);
        fft16_attr(rows16_kernel<LG, 0, 1, false, XM>, sm); fft16_attr(rows16_kernel<LG, 0, 1, true, XM>, sm);
        fft16_attr(rows16_kernel<LG, 0, 4, false, XM>, sm); fft16_attr(rows16_kernel<LG, 0, 4, true, XM>, sm);
        if (XM == 0) fft16_attr(rows16_kernel<LG, 0, 5, false, 0>, sm);
    }
}
// forward rows: `lines` lines of `in` (prologue fused) -> out (XM = 1: scattered through sct)
template <int LG, int XM>
static inline void rows16_forward(bool pd, cudaStream_t s, const double* in, double* out, int lines, int in_ls, int out_ls,
                                  const double2* tw, const RowPrologue& pro, const int* done, const Scatter& sct) {
    using G = F16<LG>;
    const int grid = ((lines + 1) / 2 + G::fpb - 1) / G::fpb;
    const bool mul = pro.mode != 0 && pro.mode != 4 && pro.a != nullptr;
    const DotEpilogue none;
#define VCH_R16(PRO, MUL, SM) launch_pdl(pd, rows16_kernel<LG, PRO, 0, MUL, XM>, grid, G::rthreads, SM, s, in, out, lines, in_ls, out_ls, tw, pro, none, done, sct)
    if (pro.mode == 0) VCH_R16(0, false, G::rows_smem_plain);
    else if (pro.mode == 4) {
        if constexpr (XM == 0) launch_pdl(pd, rows16_kernel<LG, 4, 0, false, 0>, grid, G::rthreads, G::rows_smem_staged, s, in, out, lines, in_ls, out_ls, tw, pro, none, done, sct);
        else throw Error(VCH_E_ARG, "prologue mode 4 is not available in slab mode");
    }
    else if (pro.mode == 2 && mul) VCH_R16(2, true, G::rows_smem_staged);
    else if (pro.mode == 2) VCH_R16(2, false, G::rows_smem_staged);
    else if (mul) VCH_R16(3, true, G::rows_smem_staged);
    else VCH_R16(3, false, G::rows_smem_staged);
#undef VCH_R16
}
// inverse rows: in (XM = 3: gathered through sct) -> out with the fused epilogue
template <int LG, int XM>
static inline void rows16_inverse(bool pd, cudaStream_t s, const double* in, double* out, int lines, int in_ls, int out_ls,
                                  const double2* tw, const DotEpilogue& epi, const int* done, const Scatter& sct) {
    using G = F16<LG>;
    const int grid = ((lines + 1) / 2 + G::fpb - 1) / G::fpb;
    const bool mul = epi.mode != 0 && epi.mul_a != nullptr;
    constexpr size_t sm = XM == 3 ? G::rows_smem_staged : G::rows_smem_plain;
    const RowPrologue none;
#define VCH_R16(EPI, MUL) launch_pdl(pd, rows16_kernel<LG, 0, EPI, MUL, XM>, grid, G::rthreads, sm, s, in, out, lines, in_ls, out_ls, tw, none, epi, done, sct)
    if (epi.mode == 0) VCH_R16(0, false);
    else if (epi.mode == 5) {
        if constexpr (XM == 0) launch_pdl(pd, rows16_kernel<LG, 0, 5, false, 0>, grid, G::rthreads, sm, s, in, out, lines, in_ls, out_ls, tw, none, epi, done, sct);
        else throw Error(VCH_E_ARG, "epilogue mode 5 is not available in slab mode");
    }
    else if (epi.mode == 1 && mul) VCH_R16(1, true);
    else if (epi.mode == 1) VCH_R16(1, false);
    else if (mul) VCH_R16(4, true);
    else VCH_R16(4, false);
#undef VCH_R16
}
// tmap: tensor map of the pitched buffer (cols16_tensor_map) or nullptr -> the plain kernel
template <int LG>
static inline void cols16_solve(bool pd, cudaStream_t s, double* buf, int pitch, int ncols, const double2* tw, const double* lam_col,
                                const double* lam_row, const SymbolArgs& sym, double norm, int scale_mode, const int* done,
                                const CUtensorMap* tmap = nullptr) {
    using G = F16<LG>;
    const int grid = (ncols + 2 * G::cp - 1) / (2 * G::cp);
#ifndef VCH_CPU_EMU
    if constexpr (F16T<LG>::use) {
        if (tmap) {
            launch_pdl(pd, cols16_tma_kernel<LG>, grid, G::cthreads, F16T<LG>::smem_bytes, s, *tmap, ncols, tw, lam_col, lam_row, sym, norm, scale_mode, done);
            return;
        }
    }
#endif
    launch_pdl(pd, cols16_kernel<LG>, grid, G::cthreads, G::cols_smem_bytes, s, buf, pitch, ncols, tw, lam_col, lam_row, sym, norm, scale_mode, done);
}

// Tensor map of a pitched buffer of `rows` rows (row stride pitch doubles, pitch a multiple of 2*cp) for the strips of the TMA column
// kernel: boxes of BR rows x 2*cp columns, swizzle matched to the row width.  Returns false when TMA staging does not apply (short
// columns, CPU emulation, driver entry point missing): the caller then passes nullptr to cols16_solve.
static inline bool cols16_tensor_map(CUtensorMap* out, int log2L, double* buf, int pitch, int rows) {
#ifdef VCH_CPU_EMU
    (void)out; (void)log2L; (void)buf; (void)pitch; (void)rows;
    return false;
#else
    if (log2L < 9 || log2L > 13) return false;
    if (getenv("VCH_COLS_TMA") && atoi(getenv("VCH_COLS_TMA")) == 0) return false;
    int cp = 4, br = 0;
    switch (log2L) {
        case 9: cp = F16<9>::cp; br = F16T<9>::BR; break;    case 10: cp = F16<10>::cp; br = F16T<10>::BR; break;
        case 11: cp = F16<11>::cp; br = F16T<11>::BR; break; case 12: cp = F16<12>::cp; br = F16T<12>::BR; break;
        default: cp = F16<13>::cp; br = F16T<13>::BR; break;
    }
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * cp), (cuuint32_t)br};
    const cuuint32_t est[2] = {1, 1};
    const CUtensorMapSwizzle sw = cp == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : (cp == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, buf, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
#endif
}
#endif

// host: per-pass twiddle tables for the length Lf = 2^log2L, layout as F16<LOG2L>::tw_* describes
static inline std::vector<double2> fft16_twiddles(int log2L) {
    const int Lf = 1 << log2L, rem = log2L % 4, n16 = log2L / 4;
    const int lastR = rem == 0 ? 16 : (1 << rem), npass = rem == 0 ? n16 : n16 + 1, mids = npass - 2, lastNs = Lf / lastR;
    std::vector<double2> tab;
    auto w = [&](long long m) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)(m % Lf) / Lf;
        return make_double2((double)cosl(a), (double)sinl(a));
    };
    int ns = 16;
    for (int p = 1; p <= mids; ++p, ns *= 16)
        for (int r = 1; r < 16; ++r)
            for (int k = 0; k < ns; ++k) tab.push_back(w((long long)k * r * (Lf / (ns * 16))));
    for (int r = 1; r < lastR; ++r)
        for (int k = 0; k < lastNs; ++k) tab.push_back(w((long long)k * r));
    return tab;
}

}  // namespace vch
