// bar_bench.cu — microbenchmark of the env-group kernel's software barrier on B200 (tools/micro; not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bar_bench bar_bench.cu && ./bar_bench
// G groups of S co-resident CTAs (cooperative launch) run N barriers back to back; variants:
//   0  __syncthreads + [thread 0: fence, atomicAdd, ld.acquire spin, fence] + __syncthreads   (epi_grp.cuh today)
//   1  same with red.release.gpu / ld.acquire.gpu and no explicit fences
//   2  per-CTA flags: st.release own flag, one warp polls all S flags (no atomics)
//   3  variant 0, with every thread storing 16 floats to global memory before each barrier (fence drains them)
//   4  __syncthreads only (floor)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_rel(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int V>
__global__ void k(unsigned* bars, unsigned* flags, float* scratch, int S, int N, long long* cycles) {
  const int g = blockIdx.x / S, c = blockIdx.x % S;
  unsigned* bar = bars + g * 32;
  unsigned* fl = flags + (size_t)g * 1024;
  unsigned epoch = 0;
  float acc = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) {
    if (V == 3) {
#pragma unroll
      for (int q = 0; q < 16; q++) scratch[((size_t)blockIdx.x * 16 + q) * blockDim.x + threadIdx.x] = acc + q;
    }
    if (V == 4) { __syncthreads(); continue; }
    __syncthreads();
    epoch += S;
    if (V == 0 || V == 3) {
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(ld_acq(bar) - epoch) < 0) {}
        __threadfence();
      }
    } else if (V == 1) {
      if (threadIdx.x == 0) {
        red_rel(bar);
        while ((int)(ld_acq(bar) - epoch) < 0) {}
      }
    } else if (V == 2) {
      if (threadIdx.x == 0) st_rel(fl + c, (unsigned)(i + 1));
      if (threadIdx.x < 32) {
        for (int q = threadIdx.x; q < S; q += 32)
          while ((int)(ld_acq(fl + q) - (unsigned)(i + 1)) < 0) {}
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc < 0) scratch[0] = acc;
}

template <int V>
float run(int G, int S, int threads, int N, unsigned* bars, unsigned* flags, float* scratch, long long* cyc) {
  cudaMemset(bars, 0, 4 * 32 * 64);
  cudaMemset(flags, 0, 4 * 1024 * 64);
  void* args[] = {&bars, &flags, &scratch, &S, &N, &cyc};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k<V>, dim3(G * S), dim3(threads), args, 0, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / N;  // us per barrier
}

int main() {
  unsigned *bars, *flags; float* scratch; long long* cyc;
  cudaMalloc(&bars, 4 * 32 * 64); cudaMalloc(&flags, 4 * 1024 * 64);
  cudaMalloc(&scratch, (size_t)296 * 16 * 512 * 4); cudaMalloc(&cyc, 8 * 1024);
  const int N = 20000;
  struct Cfg { int G, S, T; } cfgs[] = {{4, 37, 448}, {4, 74, 224}, {2, 74, 448}, {1, 148, 448}, {74, 2, 416}, {148, 1, 448}};
  for (auto c : cfgs) {
    printf("G=%d S=%d threads=%d:", c.G, c.S, c.T);
    printf(" v0 fence+atomic %.3f us", run<0>(c.G, c.S, c.T, N, bars, flags, scratch, cyc));
    printf(" | v1 red.release %.3f us", run<1>(c.G, c.S, c.T, N, bars, flags, scratch, cyc));
    printf(" | v2 flags %.3f us", run<2>(c.G, c.S, c.T, N, bars, flags, scratch, cyc));
    printf(" | v3 v0+stores %.3f us", run<3>(c.G, c.S, c.T, N, bars, flags, scratch, cyc));
    printf(" | v4 syncthreads %.3f us\n", run<4>(c.G, c.S, c.T, N, bars, flags, scratch, cyc));
  }
  return 0;
}
