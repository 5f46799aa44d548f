#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* counter, cudaGraphConditionalHandle h) {
    int c = atomicAdd(counter, 1) + 1;
    if (c >= 5) cudaGraphSetConditional(h, 0);
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; 
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("addnode: %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    // capture body into the conditional body graph
    cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    body<<<1,1,0,s>>>(d, h);
    cudaStreamEndCapture(s, nullptr);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int hc; cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost); printf("counter=%d err=%s\n", hc, cudaGetErrorString(cudaGetLastError()));
}
