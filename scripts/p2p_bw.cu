// Scratch: achievable NVLink bandwidth from SM load/store kernels between two GPUs of one process (peer access).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/p2p_bw.cu -o scripts/p2p_bw && ./scripts/p2p_bw
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <typename T> __global__ void copy_k(const T* __restrict__ src, T* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd)); if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = 64ull << 20;
    double *a0, *a1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, bytes)); CK(cudaMemset(a1, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMemset(a0, 0, bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t mb : {8ull, 64ull}) {
        const size_t nb = mb << 20;
        for (int grid : {148, 592, 2368}) {
            for (int mode = 0; mode < 4; ++mode) {
                const bool wr = mode < 2, wide = mode & 1;
                const void* s = wr ? (void*)a0 : (void*)a1; void* d = wr ? (void*)a1 : (void*)a0;
                auto run = [&] {
                    if (wide) copy_k<double2><<<grid, 256>>>((const double2*)s, (double2*)d, nb / 16);
                    else copy_k<double><<<grid, 256>>>((const double*)s, (double*)d, nb / 8);
                };
                for (int w = 0; w < 3; ++w) run();
                cudaEventRecord(e0); for (int r = 0; r < 20; ++r) run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("%2zu MB grid %4d %s %2d B/thread: %7.1f us  %6.1f GB/s\n", mb, grid, wr ? "remote WRITE" : "remote READ ", wide ? 16 : 8,
                       1e3 * ms / 20, nb / (ms / 20 * 1e-3) / 1e9);
            }
        }
    }
    // cudaMemcpyPeer-style DMA for reference
    cudaEventRecord(e0); for (int r = 0; r < 20; ++r) cudaMemcpyAsync(a1, a0, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("copy engine 64 MB: %.1f GB/s   err=%s\n", bytes / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
