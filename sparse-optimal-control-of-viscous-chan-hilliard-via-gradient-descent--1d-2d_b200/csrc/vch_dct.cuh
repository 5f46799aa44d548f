// DCT-I fast solve: out = V f(Lambda)^-1 V^-1 in, where V diagonalises the mirror-ghost Neumann Laplacian
// (eigenvectors cos(pi j k / N), eigenvalues -(4/h^2) sin^2(pi k / 2N)) and
// f(lambda) = c0 + abar*lambda + c2*lambda^2 is the constant-coefficient symbol of the Schur / adjoint operator.
//
// A line of N+1 reals is transformed through the FFT of its even extension (length 2N).  Two lines share one
// complex FFT (line a -> real part, line b -> imaginary part): the transform of a real-even sequence is real,
// so Re/Im of the result are the two DCT-I outputs and no post-twiddle pass is needed.  The FFT itself is an
// in-place shared-memory Stockham transform, radix 8 with one radix-4/2 tail pass, 8 points per thread in
// registers, twiddles from a precomputed table.  Power-of-two N only; other N use the dense-table kernels below.
#pragma once
#include "vch_common.cuh"

namespace vch {

struct DctAxis {
    int n = 0;            // nodes = N+1
    bool fft = false;     // N is a power of two >= 4
    int Lf = 0, log2L = 0;
    double2* tw = nullptr;    // Lf twiddles exp(-2 pi i m / Lf)
    double* denseT = nullptr; // n*n, denseT[j*n + k] = 2 c_j cos(pi j k / N)   (transposed for coalescing)
    double* lam = nullptr;    // n eigenvalues of -L1d (>= 0)
};

struct SymbolArgs {
    double c0, c2;
    const double* abar_ptr;   // device scalar (may be null -> abar_const)
    double abar_const;
};

struct DctPlan {
    DctAxis inner, outer;     // inner = contiguous axis (length ni), outer = strided axis (length no)
    int ni = 0, no = 0;
    DevBuf tmp1, tmp2;
    LaunchLog* log = nullptr;
    void init(int no_, int ni_, double h_outer, double h_inner, LaunchLog* launch_log);
    void destroy();
    // out = P^-1 in   (in may equal out)
    void apply(cudaStream_t s, const double* in, double* out, const SymbolArgs& sym, const int* done_flag);
};

// ------------------------------------------------------------------------------------------------ device side
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

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

// One Stockham pass of radix R over the FFT stored in `data` (length Lf), executed by tpf = Lf/8 threads.
template <int R>
__device__ __forceinline__ void fft_pass(double2* data, int Lf, int Ns, int t, int tpf, const double2* __restrict__ tw) {
    constexpr int NB = 8 / R;
    const int stride = Lf / R;
    const int twstep = Lf / (Ns * R);
    double2 v[NB][R];
    int kk[NB], jj[NB];
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * tpf;
        const int k = j & (Ns - 1);
        jj[m] = j; kk[m] = k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double2 x = data[j + r * stride];
            if (r > 0 && k > 0) x = cmul(x, __ldg(&tw[r * k * twstep]));
            v[m][r] = x;
        }
        dft<R>(v[m]);
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j0 = (jj[m] - kk[m]) * R + kk[m];
#pragma unroll
        for (int r = 0; r < R; ++r) data[j0 + r * Ns] = v[m][r];
    }
    __syncthreads();
}

__device__ __forceinline__ void fft_inplace(double2* data, int Lf, int log2L, int t, int tpf, const double2* __restrict__ tw) {
    int Ns = 1;
    const int n8 = log2L / 3, rem = log2L - 3 * n8;
    for (int p = 0; p < n8; ++p) { fft_pass<8>(data, Lf, Ns, t, tpf, tw); Ns *= 8; }
    if (rem == 2) fft_pass<4>(data, Lf, Ns, t, tpf, tw);
    else if (rem == 1) fft_pass<2>(data, Lf, Ns, t, tpf, tw);
}

// Lines along the contiguous axis.  Each CTA owns ppb line pairs; blockDim = ppb * Lf/8.
__global__ void dct_rows_fft_kernel(const double* __restrict__ in, double* __restrict__ out, int lines, int n,
                                    int Lf, int log2L, int ppb, const double2* __restrict__ tw,
                                    const int* __restrict__ done) {
    if (done && *done) return;
    extern __shared__ double2 sm[];
    const int N = n - 1, ld = Lf + 1;
    const int pair0 = blockIdx.x * ppb;
    for (int e = threadIdx.x; e < ppb * n; e += blockDim.x) {
        const int f = e / n, j = e - f * n;
        const int a = 2 * (pair0 + f), b = a + 1;
        double2 z;
        z.x = (a < lines) ? in[(size_t)a * n + j] : 0.0;
        z.y = (b < lines) ? in[(size_t)b * n + j] : 0.0;
        sm[f * ld + j] = z;
        if (j > 0 && j < N) sm[f * ld + 2 * N - j] = z;
    }
    __syncthreads();
    const int tpf = Lf >> 3;
    const int f = threadIdx.x / tpf, t = threadIdx.x - f * tpf;
    fft_inplace(sm + f * ld, Lf, log2L, t, tpf, tw);
    for (int e = threadIdx.x; e < ppb * n; e += blockDim.x) {
        const int ff = e / n, k = e - ff * n;
        const int a = 2 * (pair0 + ff), b = a + 1;
        const double2 z = sm[ff * ld + k];
        if (a < lines) out[(size_t)a * n + k] = z.x;
        if (b < lines) out[(size_t)b * n + k] = z.y;
    }
}

// Lines along the strided axis, fused forward transform -> divide by the symbol -> inverse transform.
// Each CTA owns 2*ppb adjacent columns so global accesses are 16*ppb-byte segments.
__global__ void dct_cols_fft_solve_kernel(double* d, int no, int ni,
                                          int Lf, int log2L, int ppb, const double2* __restrict__ tw,
                                          const double* __restrict__ lam_o, const double* __restrict__ lam_i,
                                          SymbolArgs sy, double norm, const int* __restrict__ done) {
    if (done && *done) return;
    extern __shared__ double2 sm[];
    const int N = no - 1, ld = Lf + 1, cw = 2 * ppb;
    const int c0 = blockIdx.x * cw;
    double* smd = reinterpret_cast<double*>(sm);
    for (int e = threadIdx.x; e < no * cw; e += blockDim.x) {
        const int o = e / cw, cc = e - o * cw, c = c0 + cc;
        const double x = (c < ni) ? d[(size_t)o * ni + c] : 0.0;
        const int f = cc >> 1, part = cc & 1;
        smd[2 * (f * ld + o) + part] = x;
        if (o > 0 && o < N) smd[2 * (f * ld + 2 * N - o) + part] = x;
    }
    __syncthreads();
    const int tpf = Lf >> 3;
    const int f = threadIdx.x / tpf, t = threadIdx.x - f * tpf;
    fft_inplace(sm + f * ld, Lf, log2L, t, tpf, tw);
    const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
    for (int e = threadIdx.x; e < ppb * no; e += blockDim.x) {
        const int ff = e / no, k = e - ff * no;
        const int ca = c0 + 2 * ff, cb = ca + 1;
        double2 z = sm[ff * ld + k];
        const double lo = lam_o[k];
        const double la = lo + ((ca < ni) ? lam_i[ca] : 0.0), lb = lo + ((cb < ni) ? lam_i[cb] : 0.0);
        z.x *= norm / (sy.c0 + la * (abar + sy.c2 * la));
        z.y *= norm / (sy.c0 + lb * (abar + sy.c2 * lb));
        sm[ff * ld + k] = z;
        if (k > 0 && k < N) sm[ff * ld + 2 * N - k] = z;
    }
    __syncthreads();
    fft_inplace(sm + f * ld, Lf, log2L, t, tpf, tw);
    for (int e = threadIdx.x; e < no * cw; e += blockDim.x) {
        const int o = e / cw, cc = e - o * cw, c = c0 + cc;
        if (c < ni) d[(size_t)o * ni + c] = smd[2 * ((cc >> 1) * ld + o) + (cc & 1)];
    }
}

// Dense-table fallbacks for N that is not a power of two (small validation grids).
__global__ void dct_rows_dense_kernel(const double* __restrict__ in, double* __restrict__ out, int lines, int n,
                                      const double* __restrict__ Tt, const int* __restrict__ done) {
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
    if (done && *done) return;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)no * ni) return;
    const int k = (int)(idx / ni), c = (int)(idx - (long long)k * ni);
    const double abar = sy.abar_ptr ? *sy.abar_ptr : sy.abar_const;
    const double l = lam_o[k] + lam_i[c];
    d[idx] *= norm / (sy.c0 + l * (abar + sy.c2 * l));
}

// ------------------------------------------------------------------------------------------------ host side
static inline void dct_axis_init(DctAxis& ax, int n, double h) {
    const int N = n - 1;
    ax.n = n;
    ax.fft = (N >= 4) && ((N & (N - 1)) == 0) && (2 * N <= 8192);
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
    int ppb = 512 / tpf; if (ppb < 1) ppb = 1; if (ppb > 8) ppb = 8;
    const int pairs = (ncols + 1) / 2;
    if (ppb > pairs) ppb = pairs;
    return ppb;
}

inline void DctPlan::init(int no_, int ni_, double h_outer, double h_inner, LaunchLog* launch_log) {
    no = no_; ni = ni_; log = launch_log;
    dct_axis_init(inner, ni, h_inner);
    dct_axis_init(outer, no, h_outer);
    tmp1.alloc((size_t)no * ni);
    tmp2.alloc((size_t)no * ni);
    VCH_CUDA(cudaFuncSetAttribute(dct_rows_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VCH_CUDA(cudaFuncSetAttribute(dct_cols_fft_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

inline void DctPlan::apply(cudaStream_t s, const double* in, double* out, const SymbolArgs& sym, const int* done) {
    const long long n = (long long)no * ni;
    const int eb = (int)((n + 255) / 256);
    const double norm = 1.0 / (4.0 * (double)(ni - 1) * (double)(no - 1));
    auto rows = [&](const double* a, double* b) {
        if (inner.fft) {
            const int ppb = dct_rows_ppb(inner, no), tpf = inner.Lf >> 3;
            const int grid = ((no + 1) / 2 + ppb - 1) / ppb;
            const size_t smem = sizeof(double2) * (size_t)ppb * (inner.Lf + 1);
            log->begin("dct_rows_fft", s);
            dct_rows_fft_kernel<<<grid, ppb * tpf, smem, s>>>(a, b, no, ni, inner.Lf, inner.log2L, ppb, inner.tw, done);
        } else {
            log->begin("dct_rows_dense", s);
            dct_rows_dense_kernel<<<eb, 256, 0, s>>>(a, b, no, ni, inner.denseT, done);
        }
        log->end(s);
    };
    if (outer.fft) {
        // rows (in -> tmp1), fused column solve in place on tmp1, rows (tmp1 -> out)
        double* t1 = tmp1.p;
        rows(in, t1);
        const int ppb = dct_cols_ppb(outer, ni), tpf = outer.Lf >> 3;
        const int grid = ((ni + 1) / 2 + ppb - 1) / ppb;
        const size_t smem = sizeof(double2) * (size_t)ppb * (outer.Lf + 1);
        log->begin("dct_cols_fft_solve", s);
        dct_cols_fft_solve_kernel<<<grid, ppb * tpf, smem, s>>>(t1, no, ni, outer.Lf, outer.log2L, ppb, outer.tw,
                                                               outer.lam, inner.lam, sym, norm, done);
        log->end(s);
        if (inner.fft) rows(t1, out); else { rows(t1, tmp2.p); VCH_CUDA(cudaMemcpyAsync(out, tmp2.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s)); }
    } else {
        double *t1 = tmp1.p, *t2 = tmp2.p;
        rows(in, t1);
        log->begin("dct_cols_dense", s); dct_cols_dense_kernel<<<eb, 256, 0, s>>>(t1, t2, no, ni, outer.denseT, done); log->end(s);
        log->begin("dct_scale", s); dct_scale_kernel<<<eb, 256, 0, s>>>(t2, no, ni, outer.lam, inner.lam, sym, norm, done); log->end(s);
        log->begin("dct_cols_dense", s); dct_cols_dense_kernel<<<eb, 256, 0, s>>>(t2, t1, no, ni, outer.denseT, done); log->end(s);
        if (inner.fft) rows(t1, out);
        else { rows(t1, t2); VCH_CUDA(cudaMemcpyAsync(out, t2, n * sizeof(double), cudaMemcpyDeviceToDevice, s)); }
    }
    VCH_CUDA(cudaGetLastError());
}

inline void DctPlan::destroy() {
    for (DctAxis* ax : {&inner, &outer}) {
        if (ax->tw) cudaFree(ax->tw);
        if (ax->denseT) cudaFree(ax->denseT);
        if (ax->lam) cudaFree(ax->lam);
        ax->tw = nullptr; ax->denseT = nullptr; ax->lam = nullptr;
    }
    tmp1.release(); tmp2.release();
}

}  // namespace vch
