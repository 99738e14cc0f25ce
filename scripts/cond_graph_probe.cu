#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *ctr, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0) {
    int v = atomicAdd(ctr, 1);
    cudaGraphSetConditional(h, v < 9 ? 1u : 0u);
  }
}
int main() {
  int *ctr; cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  auto e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
  printf("addnode %s\n", cudaGetErrorString(e));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  printf("begin %s\n", cudaGetErrorString(e));
  body<<<1, 32, 0, st>>>(ctr, h);
  body<<<1, 32, 0, st>>>(ctr, h);
  e = cudaStreamEndCapture(st, nullptr);
  printf("end %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
  printf("inst %s\n", cudaGetErrorString(e));
  e = cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int v; cudaMemcpy(&v, ctr, 4, cudaMemcpyDeviceToHost);
  printf("launch %s ctr=%d\n", cudaGetErrorString(cudaGetLastError()), v);
}
