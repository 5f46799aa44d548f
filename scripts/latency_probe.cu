// Scratch probe (not part of the product): what does a phase boundary cost on B200?
//   A  chain of empty kernels in a CUDA graph                      -> launch gap
//   B  chain of 8.4 MB field copies (L2-resident), one kernel each -> kernel boundary + ramp + drain
//   C  the same copies inside ONE cooperative kernel, grid.sync() between phases
//   D  grid.sync() alone;  E  hand-written atomic barrier alone
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/latency_probe.cu -o scripts/latency_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 1000) *p = 1; }
__global__ void __launch_bounds__(256, 4) copy_kernel(const double2* __restrict__ a, double2* __restrict__ b, long long n2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256, 4) coop_copy(double2* a, double2* b, long long n2, int phases, int do_copy) {
    cg::grid_group g = cg::this_grid();
    for (int p = 0; p < phases; ++p) {
        if (do_copy) {
            const double2* s = (p & 1) ? b : a; double2* d = (p & 1) ? a : b;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
        }
        g.sync();
    }
}
// sense-reversing barrier on one counter: last arriver bumps the generation
__device__ __forceinline__ void my_barrier(unsigned int* count, volatile unsigned int* gen, unsigned int nb) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int g0 = *gen;
        __threadfence();
        if (atomicAdd(count, 1u) == nb - 1) { *count = 0; __threadfence(); *gen = g0 + 1; }
        else while (*gen == g0) { }
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256, 4) coop_copy2(double2* a, double2* b, long long n2, int phases, int do_copy, unsigned int* count, unsigned int* gen) {
    for (int p = 0; p < phases; ++p) {
        if (do_copy) {
            const double2* s = (p & 1) ? b : a; double2* d = (p & 1) ? a : b;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
        }
        my_barrier(count, gen, gridDim.x);
    }
}

int main() {
    const long long n = 1025LL * 1025LL, n2 = n / 2;
    double2 *a, *b; CK(cudaMalloc(&a, n * 8 + 64)); CK(cudaMalloc(&b, n * 8 + 64));
    CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
    unsigned int* bar; CK(cudaMalloc(&bar, 8)); CK(cudaMemset(bar, 0, 8));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int K = 70, reps = 50;
    float ms;
    for (int which = 0; which < 3; ++which) {
        const int grids[3] = {1, 592, 592};
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < K; ++k) {
            if (which < 2) empty_kernel<<<grids[which], 256, 0, s>>>(nullptr);
            else copy_kernel<<<592, 256, 0, s>>>((k & 1) ? b : a, (k & 1) ? a : b, n2);
        }
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, s));
        cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.2f us per kernel\n", which == 0 ? "A  graph, empty kernel 1 CTA      " : which == 1 ? "A' graph, empty kernel 592 CTAs   " : "B  graph, 8.4 MB copy kernel      ", 1e3 * ms / (reps * K));
    }
    // plain stream launches of the copy kernel
    for (int w = 0; w < 20; ++w) copy_kernel<<<592, 256, 0, s>>>(a, b, n2);
    cudaEventRecord(e0, s);
    for (int r = 0; r < reps * K; ++r) copy_kernel<<<592, 256, 0, s>>>((r & 1) ? b : a, (r & 1) ? a : b, n2);
    cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("B' stream, 8.4 MB copy kernel     : %.2f us per kernel\n", 1e3 * ms / (reps * K));
    for (int do_copy = 1; do_copy >= 0; --do_copy) {
        int phases = K; long long nn = n2;
        void* args[] = {&a, &b, &nn, &phases, &do_copy};
        for (int w = 0; w < 3; ++w) CK(cudaLaunchCooperativeKernel((void*)coop_copy, dim3(592), dim3(256), args, 0, s));
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) CK(cudaLaunchCooperativeKernel((void*)coop_copy, dim3(592), dim3(256), args, 0, s));
        cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.2f us per phase\n", do_copy ? "C  cooperative, copy + grid.sync  " : "D  cooperative, grid.sync only    ", 1e3 * ms / (reps * K));
        unsigned int* cnt = bar; unsigned int* gen = bar + 1;
        void* args2[] = {&a, &b, &nn, &phases, &do_copy, &cnt, &gen};
        for (int w = 0; w < 3; ++w) CK(cudaLaunchCooperativeKernel((void*)coop_copy2, dim3(592), dim3(256), args2, 0, s));
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) CK(cudaLaunchCooperativeKernel((void*)coop_copy2, dim3(592), dim3(256), args2, 0, s));
        cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.2f us per phase\n", do_copy ? "C' cooperative, copy + own barrier" : "E  cooperative, own barrier only  ", 1e3 * ms / (reps * K));
    }
    // 148-CTA variants of the barrier
    for (int nb : {148, 296}) {
        int phases = K, do_copy = 0; long long nn = n2;
        void* args[] = {&a, &b, &nn, &phases, &do_copy};
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps; ++r) CK(cudaLaunchCooperativeKernel((void*)coop_copy, dim3(nb), dim3(256), args, 0, s));
        cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("D  grid.sync only, %d CTAs: %.2f us per phase\n", nb, 1e3 * ms / (reps * K));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
