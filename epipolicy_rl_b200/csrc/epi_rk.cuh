// epi_rk.cuh — pieces shared by the lane-per-cell kernels (epi_cell.cuh, epi_env.cuh): float copies of the
// Dormand-Prince tableau, element types of the working-set columns / per-env tables, a 2-vector.
#pragma once
#include "epi_kernels.cuh"
#include "epi_fdiv.cuh"

namespace epi {
namespace rk {

// float copies of the tableau for the fp32 mode (no F2F in the stage algebra)
__constant__ float RK_A32[6][5] = {
    {0, 0, 0, 0, 0},
    {(float)(1.0 / 5), 0, 0, 0, 0},
    {(float)(3.0 / 40), (float)(9.0 / 40), 0, 0, 0},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0, 0},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656)}};
__constant__ float RK_B32[7] = {(float)(35.0 / 384), 0, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84), 0};
__constant__ float RK_E32[7] = {(float)(-71.0 / 57600), 0, (float)(71.0 / 16695), (float)(-71.0 / 1920), (float)(17253.0 / 339200), (float)(-22.0 / 525), (float)(1.0 / 40)};
__device__ __forceinline__ float rkA(float, int s, int j) { return RK_A32[s][j]; }
__device__ __forceinline__ double rkA(double, int s, int j) { return RK_A[s][j]; }
__device__ __forceinline__ float rkB(float, int j) { return RK_B32[j]; }
__device__ __forceinline__ double rkB(double, int j) { return RK_B[j]; }
__device__ __forceinline__ float rkE(float, int j) { return RK_E32[j]; }
__device__ __forceinline__ double rkE(double, int j) { return RK_E[j]; }

// stage-input coefficients by phase of the RK45 state machine (epi_cell.cuh: rk45_day): y = X + hh * sum_j CO[phase][j] K_j.
// phase 0 / 8: y = X; phase 1: X + h0 K_0; phases 2..6: row `phase - 1` of A; phase 7: B.
#define EPI_CO_ROWS(T)                                                                                             \
  {{0, 0, 0, 0, 0, 0},                                                                                             \
   {1, 0, 0, 0, 0, 0},                                                                                             \
   {(T)(1.0 / 5), 0, 0, 0, 0, 0},                                                                                  \
   {(T)(3.0 / 40), (T)(9.0 / 40), 0, 0, 0, 0},                                                                     \
   {(T)(44.0 / 45), (T)(-56.0 / 15), (T)(32.0 / 9), 0, 0, 0},                                                      \
   {(T)(19372.0 / 6561), (T)(-25360.0 / 2187), (T)(64448.0 / 6561), (T)(-212.0 / 729), 0, 0},                      \
   {(T)(9017.0 / 3168), (T)(-355.0 / 33), (T)(46732.0 / 5247), (T)(49.0 / 176), (T)(-5103.0 / 18656), 0},          \
   {(T)(35.0 / 384), 0, (T)(500.0 / 1113), (T)(125.0 / 192), (T)(-2187.0 / 6784), (T)(11.0 / 84)},                 \
   {0, 0, 0, 0, 0, 0}}
__constant__ float RK_CO32[9][6] = EPI_CO_ROWS(float);
__constant__ double RK_CO64[9][6] = EPI_CO_ROWS(double);
// per-phase scalars: K slot written, number of stage terms, t_eval = t + PH_TC * hh, weights of the evaluation's
// flows in the B / E sums, how the flow sums are updated (0 none, 1 add, 2 set), whether the flow column is kept
__constant__ int PH_STAGE[9] = {0, 1, 1, 2, 3, 4, 5, 6, 0};
__constant__ int PH_NJ[9] = {0, 1, 1, 2, 3, 4, 5, 6, 0};
__constant__ double PH_TC[9] = {0, 1, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1, 1, 0};
// (stage 1 has B = E = 0: its flows enter neither sum)
__constant__ int PH_ACC[9] = {0, 0, 0, 1, 1, 1, 1, 1, 2};
__constant__ int PH_KEEP[9] = {1, 1, 0, 0, 0, 0, 0, 1, 0};
#define EPI_PH_BW(T) {0, 0, (T)0, (T)(500.0 / 1113), (T)(125.0 / 192), (T)(-2187.0 / 6784), (T)(11.0 / 84), 0, (T)(35.0 / 384)}
#define EPI_PH_EW(T) {0, 0, (T)0, (T)(71.0 / 16695), (T)(-71.0 / 1920), (T)(17253.0 / 339200), (T)(-22.0 / 525), (T)(1.0 / 40), (T)(-71.0 / 57600)}
__constant__ float PH_BW32[9] = EPI_PH_BW(float);
__constant__ double PH_BW64[9] = EPI_PH_BW(double);
__constant__ float PH_EW32[9] = EPI_PH_EW(float);
__constant__ double PH_EW64[9] = EPI_PH_EW(double);
__device__ __forceinline__ float phBW(float, int ph) { return PH_BW32[ph]; }
__device__ __forceinline__ double phBW(double, int ph) { return PH_BW64[ph]; }
__device__ __forceinline__ float phEW(float, int ph) { return PH_EW32[ph]; }
__device__ __forceinline__ double phEW(double, int ph) { return PH_EW64[ph]; }
__device__ __forceinline__ float rkCO(float, int ph, int j) { return RK_CO32[ph][j]; }
__device__ __forceinline__ double rkCO(double, int ph, int j) { return RK_CO64[ph][j]; }

// per-lane columns and per-env tables: element types
#define T_X double
#define T_XF real
#define T_ys real
#define T_K real
#define T_fl real
#define T_FBE real2
#define T_accY real
#define T_accYn real
#define T_fls real
#define T_tdc real
#define T_cmA real
#define T_cmB real
#define T_mvY float
#define T_mvD float
#define T_ysS double
#define T_KS double
#define T_red double
#define T_mult_y float
#define T_mult_d float
#define T_cpv float
#define T_act uint8_t
#define T_live uint8_t
#define T_elive uint8_t
#define T_expr double
#define T_sel double
#define T_mpy float
#define T_mpg float
#define T_mpt0 float
#define T_mpFy float
#define T_mpFg float
#define T_coefS real
#define T_Tnum float
#define T_Tden real
#define T_fi_y float
#define T_fi_gap float
#define T_ft_y float
#define T_ft_gap float
#define T_cost double
#define T_crY double
#define T_crD double
#define T_mob_y float
#define T_mob_g float
#define T_mxr float
#define T_mxc float


template <class T> struct Pair { T x, y; };
template <> struct __align__(8) Pair<float> { float x, y; };
template <> struct __align__(16) Pair<double> { double x, y; };
#define real2 Pair<real>


}  // namespace rk
}  // namespace epi
