// Floor of a dependent kernel node inside a CUDA graph on this device: chains of N trivial kernels (1 CTA / 148 CTAs,
// with and without a dependent global load + store), replayed as a graph.  Context for DESIGN.md section 4: a node of
// the S2 V-cycle costs ~6 us; this prints what an empty node costs.  nvcc -O3 -arch=sm_100a -o bench_graphnode ...
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_empty(double* p) {}
__global__ void k_touch(double* p) { p[blockIdx.x * 256 + threadIdx.x] += 1.0; }
__global__ void k_chain(double* p) {  // two dependent global round trips + store, like metadata -> data -> result
  int i = (int)p[blockIdx.x * 256 + threadIdx.x + 65536] & 1023;
  p[blockIdx.x * 256 + threadIdx.x] = p[i + 131072] + 1.0;
}
template <class K>
float run(K kern, int grid, int n, double* d) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) kern<<<grid, 256, 0, st>>>(d);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0, st);
    for (int q = 0; q < 10; ++q) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  return best * 1e3f / (10 * n);
}
int main() {
  double* d;
  cudaMalloc(&d, 8 << 20);
  cudaMemset(d, 0, 8 << 20);
  printf("us per dependent graph node (chain of 200 kernels, 10 replays)\n%6s %10s %10s %10s\n", "grid", "empty", "touch", "2 loads");
  for (int grid : {1, 8, 148, 592, 2368})
    printf("%6d %10.2f %10.2f %10.2f\n", grid, run(k_empty, grid, 200, d), run(k_touch, grid, 200, d), run(k_chain, grid, 200, d));
  return 0;
}
