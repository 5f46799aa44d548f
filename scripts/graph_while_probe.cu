// Scratch probe (not part of the product): what does one trip of a CUDA-graph WHILE node cost on B200?
//   A  flat graph of ITERS*6 kernels   B  WHILE node whose body is the same 6 kernels, ITERS trips
//   C  graph launch latency: graphs of 6 kernels launched one after the other (the polled alternative)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/graph_while_probe.cu -o scripts/graph_while_probe
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void __launch_bounds__(256) work(const double2* __restrict__ a, double2* __restrict__ b, long long n2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256) work_last(const double2* __restrict__ a, double2* __restrict__ b, long long n2, int* counter, int iters, cudaGraphConditionalHandle h) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { int c = ++(*counter); if (c >= iters) { *counter = 0; cudaGraphSetConditional(h, 0); } else cudaGraphSetConditional(h, 1); }
}
int main() {
    const long long n2 = 1025LL * 1025LL / 2;
    double2 *a, *b; CK(cudaMalloc(&a, n2 * 16 + 64)); CK(cudaMalloc(&b, n2 * 16 + 64)); CK(cudaMemset(a, 0, n2 * 16)); CK(cudaMemset(b, 0, n2 * 16));
    int* cnt; CK(cudaMalloc(&cnt, 4)); CK(cudaMemset(cnt, 0, 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ITERS = 50, K = 6, reps = 20; float ms;
    {   // A flat
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < ITERS * K; ++i) work<<<592, 256, 0, s>>>((i & 1) ? b : a, (i & 1) ? a : b, n2);
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 2; ++w) CK(cudaGraphLaunch(ge, s));
        cudaEventRecord(e0, s); for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, s)); cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("A flat graph          : %.2f us per 6-kernel group\n", 1e3 * ms / (reps * ITERS));
    }
    {   // B WHILE
        cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
        cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
        cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
        cudaGraph_t body = p.conditional.phGraph_out[0];
        CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < K - 1; ++i) work<<<592, 256, 0, s>>>((i & 1) ? b : a, (i & 1) ? a : b, n2);
        work_last<<<592, 256, 0, s>>>(b, a, n2, cnt, ITERS, h);
        CK(cudaStreamEndCapture(s, nullptr));
        cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 2; ++w) CK(cudaGraphLaunch(ge, s));
        cudaEventRecord(e0, s); for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, s)); cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("B WHILE node, %d trips : %.2f us per trip (6 kernels)\n", ITERS, 1e3 * ms / (reps * ITERS));
    }
    {   // C many small graph launches
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < K; ++i) work<<<592, 256, 0, s>>>((i & 1) ? b : a, (i & 1) ? a : b, n2);
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 5; ++w) CK(cudaGraphLaunch(ge, s));
        cudaEventRecord(e0, s); for (int r = 0; r < reps * ITERS; ++r) CK(cudaGraphLaunch(ge, s)); cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("C one graph launch per group: %.2f us per 6-kernel group\n", 1e3 * ms / (reps * ITERS));
    }
    {   // D plain stream launches
        for (int w = 0; w < 20; ++w) work<<<592, 256, 0, s>>>(a, b, n2);
        cudaEventRecord(e0, s); for (int r = 0; r < reps * ITERS * K; ++r) work<<<592, 256, 0, s>>>((r & 1) ? b : a, (r & 1) ? a : b, n2); cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("D stream launches     : %.2f us per 6-kernel group\n", 1e3 * ms / (reps * ITERS));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
