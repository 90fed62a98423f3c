// Fake <cuda_runtime.h> for the HOST EMULATION build of epipolicy_rl_b200/csrc (tests only).
// With -DEPI_HOST_EMU -Itests/emu the product sources compile with plain g++: every "device" function
// runs on the CPU, device memory is malloc. The team kernels run with a one-thread team; the cell kernel
// (epi_cell.cuh) runs each warp as `warp_width` host threads with a pthread barrier as __syncwarp, so the
// lane-level code (shared-memory staging, votes, reductions) is executed exactly as written. This lets the CPU test suite cover
// the EpiDesc -> DevPlan compiler, the C-ABI semantics and the kernel arithmetic without a GPU, and
// compare them with the oracle. It is NOT a product path: nothing in epipolicy_rl_b200/ loads it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <functional>
#include <vector>

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
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
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
inline double __ddiv_rn(double a, double b) { return a / b; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
namespace emu {
static int warp_width = 1;                                // lanes per emulated warp (set per launch)
static thread_local pthread_barrier_t* warp_barrier = nullptr;
static thread_local int lane_id = 0;
static thread_local int* vote_buf = nullptr;
static thread_local unsigned long long* xchg_buf = nullptr;   // one 8-byte slot per thread of the CTA (shuffles)
static thread_local pthread_barrier_t* group_barrier = nullptr;  // all threads of a CTA group (epi_grp.cuh)
inline void group_arrive_wait() { if (group_barrier) pthread_barrier_wait(group_barrier); }
}  // namespace emu
inline void __syncwarp() { if (emu::warp_barrier) pthread_barrier_wait(emu::warp_barrier); }
inline void __syncthreads() { if (emu::warp_barrier) pthread_barrier_wait(emu::warp_barrier); }
// butterfly exchange inside 32-lane warps of the emulated CTA (all threads of the CTA call it together); threads
// whose partner lies beyond the CTA keep their own value. Without a lock-stepped CTA: the identity.
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
  if (!emu::warp_barrier || !emu::xchg_buf) return v;
  static_assert(sizeof(T) <= 8, "emulated shuffle: at most 8 bytes");
  unsigned long long w = 0;
  memcpy(&w, &v, sizeof(T));
  emu::xchg_buf[emu::lane_id] = w;
  pthread_barrier_wait(emu::warp_barrier);
  const int partner = emu::lane_id ^ m;
  T r = v;
  if (partner < emu::warp_width) { unsigned long long x = emu::xchg_buf[partner]; memcpy(&r, &x, sizeof(T)); }
  pthread_barrier_wait(emu::warp_barrier);
  return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
  if (!emu::warp_barrier || !emu::xchg_buf) return v;
  unsigned long long w = 0;
  memcpy(&w, &v, sizeof(T));
  emu::xchg_buf[emu::lane_id] = w;
  pthread_barrier_wait(emu::warp_barrier);
  const int partner = (emu::lane_id & ~31) + src;
  T r = v;
  if (partner < emu::warp_width) { unsigned long long x = emu::xchg_buf[partner]; memcpy(&r, &x, sizeof(T)); }
  pthread_barrier_wait(emu::warp_barrier);
  return r;
}
inline int __any_sync(unsigned, int pred) {
  if (!emu::warp_barrier) return pred;
  emu::vote_buf[emu::lane_id] = pred;
  pthread_barrier_wait(emu::warp_barrier);
  int r = 0;
  for (int i = 0; i < emu::warp_width; i++) r |= emu::vote_buf[i];
  pthread_barrier_wait(emu::warp_barrier);
  return r;
}
inline int __syncthreads_or(int pred) { return __any_sync(0xffffffffu, pred); }
namespace emu {
// run `n_warps` warps one after the other, each as `width` lock-stepped host threads
inline void run_warps(int n_warps, int width, size_t smem_bytes, const std::function<void(char*, int, int)>& body) {
  warp_width = width;
  std::vector<char> smem(smem_bytes + 64);
  std::vector<int> votes(width);
  for (int w = 0; w < n_warps; w++) {
    memset(smem.data(), 0, smem.size());
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)width);
    struct Arg { const std::function<void(char*, int, int)>* body; char* smem; int lane, warp; pthread_barrier_t* bar; int* votes; };
    std::vector<Arg> args(width);
    std::vector<pthread_t> th(width);
    char* base = (char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int l = 0; l < width; l++) {
      args[l] = Arg{&body, base, l, w, &bar, votes.data()};
      pthread_create(&th[l], nullptr, [](void* p) -> void* {
        Arg* a = (Arg*)p;
        warp_barrier = a->bar; lane_id = a->lane; vote_buf = a->votes;
        (*a->body)(a->smem, a->lane, a->warp);
        warp_barrier = nullptr;
        return nullptr;
      }, &args[l]);
    }
    for (int l = 0; l < width; l++) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&bar);
  }
}
// run `n_groups` CTA groups one after the other; the S CTAs of a group run TOGETHER, each as `width` lock-stepped host
// threads with its own shared memory and CTA barrier, plus one barrier over all threads of the group (the group
// barrier of epi_grp.cuh). body(smem, tid, block) with block = group * S + cta.
inline void run_groups(int n_groups, int S, int width, size_t smem_bytes, const std::function<void(char*, int, int)>& body) {
  warp_width = width;
  for (int g = 0; g < n_groups; g++) {
    std::vector<std::vector<char>> smem(S, std::vector<char>(smem_bytes + 64, 0));
    std::vector<std::vector<int>> votes(S, std::vector<int>(width));
    std::vector<std::vector<unsigned long long>> xch(S, std::vector<unsigned long long>(width));
    std::vector<pthread_barrier_t> bars(S);
    pthread_barrier_t gbar;
    pthread_barrier_init(&gbar, nullptr, (unsigned)(S * width));
    for (int c = 0; c < S; c++) pthread_barrier_init(&bars[c], nullptr, (unsigned)width);
    struct Arg { const std::function<void(char*, int, int)>* body; char* smem; int lane, block; pthread_barrier_t *bar, *gbar; int* votes; unsigned long long* xch; };
    std::vector<Arg> args(S * width);
    std::vector<pthread_t> th(S * width);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1 << 20);
    for (int c = 0; c < S; c++) {
      char* base = (char*)(((uintptr_t)smem[c].data() + 15) & ~(uintptr_t)15);
      for (int l = 0; l < width; l++) {
        args[c * width + l] = Arg{&body, base, l, g * S + c, &bars[c], &gbar, votes[c].data(), xch[c].data()};
        pthread_create(&th[c * width + l], &attr, [](void* p) -> void* {
          Arg* a = (Arg*)p;
          warp_barrier = a->bar; group_barrier = a->gbar; lane_id = a->lane; vote_buf = a->votes; xchg_buf = a->xch;
          (*a->body)(a->smem, a->lane, a->block);
          warp_barrier = nullptr; group_barrier = nullptr; xchg_buf = nullptr;
          return nullptr;
        }, &args[c * width + l]);
      }
    }
    for (auto& t : th) pthread_join(t, nullptr);
    pthread_attr_destroy(&attr);
    for (int c = 0; c < S; c++) pthread_barrier_destroy(&bars[c]);
    pthread_barrier_destroy(&gbar);
  }
}
}  // namespace emu

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { size_t sharedMemPerBlockOptin = 232448, sharedMemPerMultiprocessor = 233472; int multiProcessorCount = 148; int cooperativeLaunch = 1; };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
enum { cudaHostAllocMapped = 2 };
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = (T*)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
