// Probe (GPU box): per-iteration cost of a device-side WHILE graph loop vs host-enqueued launches.
// nvcc -arch=sm_100a -o /tmp/graph_loop_probe scripts/graph_loop_probe.cu && /tmp/graph_loop_probe
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k_count(int *ctr, int limit, cudaGraphConditionalHandle h, int use) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int v = atomicAdd(ctr, 1);
    if (use) cudaGraphSetConditional(h, v + 1 < limit ? 1u : 0u);
  }
}
__global__ void k_empty(int *p) {
  if (p == nullptr && threadIdx.x == 1234) printf("x");
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int run_while(int body_kernels, int grid, int block, int iters) {
  int *ctr; CK(cudaMalloc(&ctr, 4)); CK(cudaMemset(ctr, 0, 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  k_count<<<grid, block, 0, st>>>(ctr, iters, h, 1);
  for (int i = 1; i < body_kernels; i++) k_empty<<<grid, block, 0, st>>>(ctr);
  CK(cudaStreamEndCapture(st, nullptr));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));  // warm
  CK(cudaMemset(ctr, 0, 4));
  double t0 = now();
  CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
  double dt = now() - t0;
  int v; CK(cudaMemcpy(&v, ctr, 4, cudaMemcpyDeviceToHost));
  printf("WHILE graph: %d kernels/body grid %d x %d: %d iterations, %.2f us/iteration\n", body_kernels, grid, block, v, dt / v * 1e6);
  return 0;
}
int run_host(int body_kernels, int grid, int block, int iters) {
  int *ctr; CK(cudaMalloc(&ctr, 4)); CK(cudaMemset(ctr, 0, 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraphConditionalHandle h{};
  for (int w = 0; w < 2; w++) {
    double t0 = now();
    for (int it = 0; it < iters; it++) {
      k_count<<<grid, block, 0, st>>>(ctr, iters, h, 0);
      for (int i = 1; i < body_kernels; i++) k_empty<<<grid, block, 0, st>>>(ctr);
    }
    CK(cudaStreamSynchronize(st));
    double dt = now() - t0;
    if (w) printf("host launches: %d kernels/iter grid %d x %d: %.2f us/iteration\n", body_kernels, grid, block, dt / iters * 1e6);
  }
  // a plain (non-conditional) graph of `unroll` iterations, relaunched
  const int unroll = 16;
  cudaGraph_t g; CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int it = 0; it < unroll; it++) {
    k_count<<<grid, block, 0, st>>>(ctr, iters, h, 0);
    for (int i = 1; i < body_kernels; i++) k_empty<<<grid, block, 0, st>>>(ctr);
  }
  CK(cudaStreamEndCapture(st, &g));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  for (int w = 0; w < 2; w++) {
    double t0 = now();
    for (int it = 0; it < iters / unroll; it++) CK(cudaGraphLaunch(ex, st));
    CK(cudaStreamSynchronize(st));
    double dt = now() - t0;
    if (w) printf("plain graph x%d ticks relaunched: %d kernels/iter grid %d x %d: %.2f us/iteration\n", unroll, body_kernels, grid, block, dt / (iters / unroll * unroll) * 1e6);
  }
  return 0;
}
int main() {
  for (int bk : {1, 3, 8}) {
    if (run_while(bk, 1, 32, 20000)) return 1;
    if (run_while(bk, 148, 512, 20000)) return 1;
  }
  for (int bk : {3, 8}) {
    if (run_host(bk, 1, 32, 20000)) return 1;
    if (run_host(bk, 148, 512, 20000)) return 1;
  }
  return 0;
}
