// epi_fdiv.cuh — division and the controller's power in the two precision modes.
// fp64 parity mode: IEEE. fp32 throughput mode: the fast hardware sequences (MUFU.RCP / LG2 / EX2, ~2 ulp):
// the 1e-4 tolerance of that mode leaves room, an IEEE float division is ~10 dependent instructions plus a
// slow-path branch, and the error-norm code of the RK45 controller alone holds ~150 of them.
#pragma once
namespace epi {
// packed pair of fp32 multiply-adds: one FFMA2 on sm_100a (Blackwell's 2-wide fp32 instruction), two fmaf on the host
#ifdef EPI_HOST_EMU
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#else
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif
__device__ __forceinline__ double fdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double ctrl_pow(double, double x, double p) { return pow(x, p); }
// a cost component's term of the scaled RMS norm, x / sc, and the norm's square root: double in the parity mode;
// in the fp32 mode the controller only needs them to float accuracy
__device__ __forceinline__ double norm_term(double, double x, double sc) { return x / sc; }
__device__ __forceinline__ double ctrl_sqrt(double, double x) { return sqrt(x); }
// reciprocal shared by several quotients with one denominator: fp32 mode only (the fp64 mode divides, like the
// reference); callers test `sizeof(real) == 8`
__device__ __forceinline__ double frcp(double x) { return 1.0 / x; }
#ifdef EPI_HOST_EMU
__device__ __forceinline__ double norm_term(float, double x, double sc) { return (double)((float)x / (float)sc); }
__device__ __forceinline__ double ctrl_sqrt(float, double x) { return (double)sqrtf((float)x); }
__device__ __forceinline__ float frcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ double ctrl_pow(float, double x, double p) { return (double)powf((float)x, (float)p); }
#else
__device__ __forceinline__ float frcp(float x);
__device__ __forceinline__ double norm_term(float, double x, double sc) { return (double)((float)x * frcp((float)sc)); }
__device__ __forceinline__ double ctrl_sqrt(float, double x) { return (double)__fsqrt_rn((float)x); }
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }  // MUFU.RCP + FMUL, ~2 ulp
__device__ __forceinline__ double ctrl_pow(float, double x, double p) { return (double)__powf((float)x, (float)p); }
#endif
}  // namespace epi
