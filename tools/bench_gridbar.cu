// Micro-benchmark behind the design of mega_grid_barrier (csrc/mega.cuh): cost per grid-wide barrier of a persistent
// kernel on this device, for several implementations and grid sizes.  Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o bench_gridbar bench_gridbar.cu ; run: ./bench_gridbar
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;

__device__ __forceinline__ void red_release(unsigned long long* p) {
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void red_relaxed(unsigned long long* p) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// variant 0: release-red + acquire polls (mega.cuh)       1: + __threadfence before       2: relaxed polls + fence after
// variant 3: cooperative groups grid.sync()               4: no barrier at all (loop overhead)
// variant 5: two-level: one arrival per SM-group of CTAs is not possible without knowing placement -> flags per CTA:
//            CTA 0 gathers, others poll a single release word (fan-in through a counter, fan-out through a flag)
// work != 0: every CTA also writes a line and reads a neighbour's line after the barrier (store -> barrier -> load)
template <int VARIANT>
__global__ void __launch_bounds__(256) k_bar(unsigned long long* sync, int nphase, double* data, int work) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long target = 0;
  double acc = 0.0;
  const int nb = gridDim.x;
  for (int p = 0; p < nphase; ++p) {
    if (work) data[(size_t)blockIdx.x * 256 + threadIdx.x] = acc + p;
    target += nb;
    if (VARIANT == 3) {
      grid.sync();
    } else if (VARIANT != 4) {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (VARIANT == 0) {
          red_release(sync);
          while (ld_acquire(sync) < target) {}
        } else if (VARIANT == 1) {
          __threadfence();
          red_release(sync);
          while (ld_acquire(sync) < target) {}
        } else if (VARIANT == 2) {
          red_release(sync);
          while (ld_relaxed(sync) < target) {}
          __threadfence();
        } else if (VARIANT == 5) {
          // fan-in: counter; fan-out: CTA 0 publishes the phase number in a separate word
          if (blockIdx.x == 0) {
            red_release(sync);
            while (ld_acquire(sync) < target) {}
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(sync + 16), "l"((unsigned long long)(p + 1)) : "memory");
          } else {
            red_release(sync);
            while (ld_acquire(sync + 16) < (unsigned long long)(p + 1)) {}
          }
        }
      }
      __syncthreads();
    }
    if (work) acc += data[(size_t)((blockIdx.x + 1) % nb) * 256 + threadIdx.x];
  }
  if (work && acc == 12345.678) printf("x");
}

template <int V>
float run(int grid, int nphase, unsigned long long* d_sync, double* d_data, int work) {
  cudaMemset(d_sync, 0, 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  void* args[] = {&d_sync, &nphase, &d_data, &work};
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(d_sync, 0, 256);
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k_bar<V>, dim3(grid), dim3(256), args, 0, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return best * 1e3f / nphase;  // us per phase
}

int main() {
  unsigned long long* d_sync;
  double* d_data;
  cudaMalloc(&d_sync, 256);
  cudaMalloc(&d_data, (size_t)1024 * 256 * 8);
  cudaMemset(d_data, 0, (size_t)1024 * 256 * 8);
  const int nphase = 2000;
  printf("us per phase (256 threads per CTA, %d phases)\n", nphase);
  printf("%6s %5s | %9s %9s %9s %9s %9s %9s\n", "grid", "work", "rel+acq", "+fence", "rlx+fence", "cg.sync", "none", "fan-out");
  for (int work = 0; work < 2; ++work)
    for (int grid : {1, 8, 74, 148, 296, 444, 592}) {
      printf("%6d %5d | %9.3f %9.3f %9.3f %9.3f %9.3f %9.3f\n", grid, work, run<0>(grid, nphase, d_sync, d_data, work),
             run<1>(grid, nphase, d_sync, d_data, work), run<2>(grid, nphase, d_sync, d_data, work),
             run<3>(grid, nphase, d_sync, d_data, work), run<4>(grid, nphase, d_sync, d_data, work),
             run<5>(grid, nphase, d_sync, d_data, work));
    }
  return 0;
}
