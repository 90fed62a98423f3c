// epi_fdiv.cuh — division and the controller's power in the two precision modes.
// fp64 parity mode: IEEE. fp32 throughput mode: the fast hardware sequences (MUFU.RCP / LG2 / EX2, ~2 ulp):
// the 1e-4 tolerance of that mode leaves room, an IEEE float division is ~10 dependent instructions plus a
// slow-path branch, and the error-norm code of the RK45 controller alone holds ~150 of them.
#pragma once
namespace epi {
__device__ __forceinline__ double fdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double ctrl_pow(double, double x, double p) { return pow(x, p); }
#ifdef EPI_HOST_EMU
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ double ctrl_pow(float, double x, double p) { return (double)powf((float)x, (float)p); }
#else
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double ctrl_pow(float, double x, double p) { return (double)__powf((float)x, (float)p); }
#endif
}  // namespace epi
