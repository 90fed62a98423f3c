// Fake <cuda_runtime.h> for the HOST EMULATION build of epipolicy_rl_b200/csrc (tests only).
// With -DEPI_HOST_EMU -Itests/emu the product sources compile with plain g++: every "device" function
// runs on the CPU with a one-thread team, device memory is malloc. This lets the CPU test suite cover
// the EpiDesc -> DevPlan compiler, the C-ABI semantics and the kernel arithmetic without a GPU, and
// compare them with the oracle. It is NOT a product path: nothing in epipolicy_rl_b200/ loads it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__ static

struct EmuDim3 { int x = 0, y = 0, z = 0; };
static EmuDim3 threadIdx, blockIdx;
static EmuDim3 blockDim{1, 1, 1}, gridDim{1, 1, 1};

template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }   // build with -ffp-contract=off
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
inline void __syncthreads() {}
inline void __syncwarp() {}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { size_t sharedMemPerBlockOptin = 232448, sharedMemPerMultiprocessor = 233472; int multiProcessorCount = 148; };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
