// Scratch probe (not part of the product): GPU idle time caused by ONE host synchronisation per Newton iteration.
// Pattern of newton_step: [residual kernel publishes scalars] -> host sync -> host decides -> cudaGraphLaunch(solve graph with a
// WHILE node) + 2 kernel launches.  Compares the wall time per iteration with and without the synchronisation in the loop.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/sync_gap_probe.cu -o scripts/sync_gap_probe
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void __launch_bounds__(256) work(const double2* __restrict__ a, double2* __restrict__ b, long long n2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256) work_pub(const double2* __restrict__ a, double2* __restrict__ b, long long n2, volatile double* host) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { host[0] = 1.0; __threadfence_system(); }
}
__global__ void __launch_bounds__(256) work_last(const double2* __restrict__ a, double2* __restrict__ b, long long n2, int* counter, int iters, cudaGraphConditionalHandle h) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { int c = ++(*counter); if (c >= iters) { *counter = 0; cudaGraphSetConditional(h, 0); } else cudaGraphSetConditional(h, 1); }
}
int main() {
    const long long n2 = 1025LL * 1025LL / 2;
    double2 *a, *b; CK(cudaMalloc(&a, n2 * 16 + 64)); CK(cudaMalloc(&b, n2 * 16 + 64)); CK(cudaMemset(a, 0, n2 * 16)); CK(cudaMemset(b, 0, n2 * 16));
    int* cnt; CK(cudaMalloc(&cnt, 4)); CK(cudaMemset(cnt, 0, 4));
    double* host; CK(cudaHostAlloc(&host, 64, cudaHostAllocMapped));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, ev; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    const int TRIPS = 5, K = 6, reps = 300; float ms;
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    // prologue (3 kernels) -> WHILE { 6 kernels }
    CK(cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 3; ++i) work<<<592, 256, 0, s>>>(a, b, n2);
    cudaStreamCaptureStatus st; const cudaGraphNode_t* deps = nullptr; size_t nd = 0;
    CK(cudaStreamGetCaptureInfo(s, &st, nullptr, nullptr, &deps, &nd));
    cudaGraphNode_t leaf = deps[0]; cudaGraph_t tmp; CK(cudaStreamEndCapture(s, &tmp));
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, &leaf, 1, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < K - 1; ++i) work<<<592, 256, 0, s>>>((i & 1) ? b : a, (i & 1) ? a : b, n2);
    work_last<<<592, 256, 0, s>>>(b, a, n2, cnt, TRIPS, h);
    CK(cudaStreamEndCapture(s, nullptr));
    cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
    auto iter = [&](int mode) {   // 0: no sync; 1: stream sync before the graph launch; 2: event sync (event right after the publishing kernel)
        work_pub<<<592, 256, 0, s>>>(a, b, n2, host);      // "residual kernel"
        if (mode == 1) cudaStreamSynchronize(s);
        if (mode == 2) { cudaEventRecord(ev, s); cudaEventSynchronize(ev); }
        if (mode == 3) { cudaEventRecord(ev, s); for (int i = 0; i < 3; ++i) work<<<592, 256, 0, s>>>(a, b, n2); cudaEventSynchronize(ev); }   // speculative prologue under the wait
        cudaGraphLaunch(ge, s);                             // solve
        work<<<592, 256, 0, s>>>(a, b, n2);                 // "dmu"
    };
    for (int mode = 0; mode < 4; ++mode) {
        for (int w = 0; w < 20; ++w) iter(mode);
        CK(cudaStreamSynchronize(s));
        cudaEventRecord(e0, s); for (int r = 0; r < reps; ++r) iter(mode); cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d (%s): %.2f us per Newton iteration (%d kernels%s)\n", mode,
               mode == 0 ? "no host sync" : mode == 1 ? "cudaStreamSynchronize" : mode == 2 ? "event sync" : "event sync + 3 speculative kernels under it",
               1e3 * ms / reps, 2 + 3 + K * TRIPS, mode == 3 ? " + 3" : "");
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
