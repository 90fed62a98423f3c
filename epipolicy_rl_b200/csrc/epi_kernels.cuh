// epi_kernels.cuh — the fused simulated-day kernel (sm_100a).
//
// One TEAM of threads owns one environment for a whole launch (1..n_days simulated days):
//   * WARP team  (small scenarios: every shipped JSON)  — several envs per CTA, the env's complete
//     RK45 working set (state, 7 stage derivatives, cumulative accumulators, parameter tables) lives
//     in shared memory; team barriers are __syncwarp().
//   * CTA team   (larger scenarios) — one env per CTA; small tables in shared memory, the
//     locale-proportional arrays in shared memory when they fit, else in an L2-resident global scratch.
//
// Per day, fused in one pass (reference call stack in SURVEY.md §3.2/§3.3):
//   prologue  = Executor.execute + get_parameter_from_delta + get_gap_parameter
//               (core/executor.py:272-284, obj/parameter.py:53-78)  -> per-env class multipliers,
//               facility tables, move table, cost rates
//   rk45_day  = scipy solve_ivp(RK45) controller (rk.py:86-176, common.py:63-134) around
//   rhs       = scipy_fun/forward/compute_delta_observable (core/core_optimize.py:165-222) with the
//               parameter interpolation of obj/parameter.py:80-92
//   epilogue  = reward = -sum(delta cumulative_cost) (core/epidemic.py:115), day counter, done flag.
//
// Precision: parameters, multipliers, interpolation weights and the
// mixing factor ft*fi*fi^T are float32 with the reference's exact rounding points (__fmul_rn /
// __fadd_rn so nvcc cannot contract them into FMAs); state and flows are `real` (double in the
// fp64 parity mode, float in the fp32 throughput mode); cumulative cost and the step-size controller
// are double in both modes.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "epi_plan.h"

#define EPI_MAX_NG 8

namespace epi {

// Dormand-Prince tableau (scipy/integrate/_ivp/rk.py:538-552)
__constant__ double RK_C[7] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1, 1};
__constant__ double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
__constant__ double RK_B[7] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0};
__constant__ double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

template <bool CTA>
struct Team {
  int lane, T;
  double* red;  // CTA: >= 33 doubles of shared scratch
  __device__ __forceinline__ void sync() const {
    if (CTA) __syncthreads(); else __syncwarp();
  }
  // all lanes receive the bit-identical sum (needed: the result steers uniform control flow)
  __device__ __forceinline__ double sum(double v) const {
#ifdef EPI_HOST_EMU
    return v;  // one-thread team (tests/emu)
#endif
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (CTA) {
      int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
      if ((threadIdx.x & 31) == 0) red[w] = v;
      __syncthreads();
      v = 0;
      for (int i = 0; i < nw; i++) v += red[i];
      __syncthreads();
    }
    return v;
  }
};

template <class real>
struct Ws {
  // small region
  float *mult_y, *mult_d, *cpv;
  uint8_t* act;
  double *expr, *sel;
  float* mpt0;   // [n_lc,P,G] parameters at facility 0, time t
  real* coefS;   // [NG,n_lc,F,G]
  float* Tnum;   // [n_lc,F,G,G]
  real* Tden;    // [n_lc,F,G,G] (only when has_ft_ops)
  float *fi_y, *fi_gap, *ft_y, *ft_gap;
  float *mob_y, *mob_g, *mxr, *mxc;  // mobility multipliers present (P.mob_dyn): mode matrices (y, gap), combined at time t
  // big region
  double* X;  // state is integrated in double in BOTH modes (see rk45_day); stages are `real`
  real *ys, *K, *accY, *accB, *accE, *accF, *accF2, *flows, *alive, *Xw, *Ag, *Wg, *s, *r;
  float *mvY, *mvD;
  double *cost, *crY, *crD;
  // fp32 mode only: stage value and the 7 stage derivatives of every move slot's SOURCE compartment in
  // double. A clamped source decays like exp(-t) to 1e-46 and below while the reference still splits the
  // day's doses in proportion to those values (executor.py:181-183); float would freeze them at 0.
  double *ysS, *KS;
  double* xnv;  // general move entries (P.xmove): the clamped amount of every entry in the current evaluation
};

template <class real>
__device__ __forceinline__ void bind_ws(const DevPlan& P, char* small, char* big, Ws<real>& w) {
  w.mult_y = (float*)(small + P.o_mult_y); w.mult_d = (float*)(small + P.o_mult_d);
  w.cpv = (float*)(small + P.o_cpv); w.act = (uint8_t*)(small + P.o_act);
  w.expr = (double*)(small + P.o_expr); w.sel = (double*)(small + P.o_sel);
  w.mpt0 = (float*)(small + P.o_mpt0); w.coefS = (real*)(small + P.o_coefS);
  w.Tnum = (float*)(small + P.o_Tnum); w.Tden = (real*)(small + P.o_Tden);
  w.fi_y = (float*)(small + P.o_fi_y); w.fi_gap = (float*)(small + P.o_fi_gap);
  w.ft_y = (float*)(small + P.o_ft_y); w.ft_gap = (float*)(small + P.o_ft_gap);
  w.mob_y = (float*)(small + P.o_mob_y); w.mob_g = (float*)(small + P.o_mob_g);
  w.mxr = (float*)(small + P.o_mxr); w.mxc = (float*)(small + P.o_mxc);
  w.X = (double*)(big + P.o_X); w.ys = (real*)(big + P.o_ys); w.K = (real*)(big + P.o_K);
  w.accY = (real*)(big + P.o_accY); w.accB = (real*)(big + P.o_accB); w.accE = (real*)(big + P.o_accE);
  w.accF = (real*)(big + P.o_accF); w.accF2 = (real*)(big + P.o_accF2); w.flows = (real*)(big + P.o_flows);
  w.alive = (real*)(big + P.o_alive); w.Xw = (real*)(big + P.o_Xw); w.Ag = (real*)(big + P.o_Ag);
  w.Wg = (real*)(big + P.o_Wg); w.s = (real*)(big + P.o_s); w.r = (real*)(big + P.o_r);
  w.mvY = (float*)(big + P.o_mvY); w.mvD = (float*)(big + P.o_mvD);
  w.cost = (double*)(big + P.o_cost); w.crY = (double*)(big + P.o_crY); w.crD = (double*)(big + P.o_crD);
  w.ysS = (double*)(big + P.o_ysS); w.KS = (double*)(big + P.o_KS); w.xnv = (double*)(big + P.o_xnv);
}

// f32 linear interpolation with the reference's rounding points (obj/parameter.py:85-89)
__device__ __forceinline__ float lerp_rn(float y, float gap, float tf) { return __fadd_rn(y, __fmul_rn(gap, tf)); }

// model parameter (lc,p,f,g) at interpolation weight tf: default * class multipliers of yesterday/today
__device__ __forceinline__ float mp_at(const DevPlan& P, const float* mult_y, const float* mult_d, int idx, float tf) {
  float m0 = __ldg(P.mp0 + idx);
  int cls = __ldg(P.mp_cls + idx);
  if (cls == 0) return m0;
  float y = __fmul_rn(m0, mult_y[cls]), d = __fmul_rn(m0, mult_d[cls]);
  return lerp_rn(y, __fsub_rn(d, y), tf);
}

// ---- mobility multipliers (MOB_MUL ops; core/executor.py:121-127 -> CooMatrix.pairwise_multiply, coo.py:112-117)
// One row (mode m, group g, origin row r) of the env's mode matrix for yesterday's and today's class multipliers:
// origin (*) delta, the row rescaled to the ORIGIN's row sum; a row multiplied to nothing becomes the identity row
// scaled by that sum (get_normalized_matrix, matrix/coo.py:173-188: sums in double, cells float). Writes y and the
// float gap = today - yesterday (get_gap_parameter, obj/parameter.py:68-78).
__device__ inline void mob_row(const DevPlan& P, const float* mult_y, const float* mult_d, int m, int g, int r, float* oy,
                               float* og) {
  const int off = P.mode_off[m], nnz_m = P.mode_off[m + 1] - off, base = P.G * off + g * nnz_m;
  const int k0 = P.mode_indptr[m * (P.L + 1) + r], k1 = P.mode_indptr[m * (P.L + 1) + r + 1];
  double osum = 0, sy = 0, sd = 0;
  for (int k = k0; k < k1; k++) {
    const float org = P.mob_org[base + k];
    const int cls = P.mob_cls[base + k];
    osum += (double)org;
    sy += (double)__fmul_rn(org, mult_y[cls]);
    sd += (double)__fmul_rn(org, mult_d[cls]);
  }
  for (int k = k0; k < k1; k++) {
    const float org = P.mob_org[base + k];
    const int cls = P.mob_cls[base + k];
    const bool diag = P.mode_indices[off + k] == r;
    const float a = sy > 1e-15 ? (float)__dmul_rn(__ddiv_rn((double)__fmul_rn(org, mult_y[cls]), sy), osum) : (diag ? (float)osum : 0.0f);
    const float b = sd > 1e-15 ? (float)__dmul_rn(__ddiv_rn((double)__fmul_rn(org, mult_d[cls]), sd), osum) : (diag ? (float)osum : 0.0f);
    oy[base + k] = a;
    og[base + k] = __fsub_rn(b, a);
  }
}

// combined mobility of union-pattern entry k (CSR order) and group g at interpolation weight tf: every mode matrix
// interpolated in float (parameter.py:85), then the f32 sum over the modes of bias * value (core_optimize.py:69-92)
__device__ __forceinline__ float mob_comb(const DevPlan& P, const float* mob_y, const float* mob_g, int g, int k, float tf) {
  float v = 0.0f;
  for (int m = 0; m < P.M; m++) {
    const int km = P.comb_src[m * P.nnz + k];
    if (km < 0) continue;
    const int off = P.mode_off[m], i = P.G * off + g * (P.mode_off[m + 1] - off) + km;
    v = __fadd_rn(v, __fmul_rn(P.mode_bias[m], lerp_rn(mob_y[i], mob_g[i], tf)));
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// right-hand side  (core_optimize.py:165-222)
// ------------------------------------------------------------------------------------------
template <class real, bool CTA>
__device__ void rhs(const DevPlan& P, const Ws<real>& w, const Team<CTA>& tm, double t, int ex_bits,
                    const real* __restrict__ y, real* __restrict__ Kout, int stage, real* accStore, real bw, real ew,
                    bool accumulate) {
  const int G = P.G, F = P.F, LG = P.LG, nnz = P.nnz, NG = P.NG, n_lc = P.n_lc, PP = P.P;
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83

  // ---- time-t tables (one per locale class, not per locale) + phase A
  for (int i = tm.lane; i < n_lc * PP * G; i += tm.T) {
    int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
    w.mpt0[i] = mp_at(P, w.mult_y, w.mult_d, ((lc * PP + p) * F + 0) * G + g, tf);
  }
  for (int i = tm.lane; i < n_lc * F * G * G; i += tm.T) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G);  // lv = lc*F + v
    float ft = lerp_rn(w.ft_y[lv * G + u], w.ft_gap[lv * G + u], tf);
    int a = (lv * G + u0) * G + u, b = (lv * G + u) * G + u0;
    float fa = lerp_rn(w.fi_y[a], w.fi_gap[a], tf), fb = lerp_rn(w.fi_y[b], w.fi_gap[b], tf);
    w.Tnum[i] = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (P.has_ft_ops) w.Tden[i] = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + b);
  }
  for (int i = tm.lane; i < NG * n_lc * F * G; i += tm.T) {
    int u0 = i % G, v = (i / G) % F, lc = (i / (G * F)) % n_lc, k = i / (G * F * n_lc);
    int lv = lc * F + v;
    real c = (real)lerp_rn(w.ft_y[lv * G + u0], w.ft_gap[lv * G + u0], tf);
    c *= (real)mp_at(P, w.mult_y, w.mult_d, ((lc * PP + P.ig_p0[k]) * F + v) * G + u0, tf);
    for (int q = P.ig_np_off[k]; q < P.ig_np_off[k + 1]; q++)
      c *= (real)mp_at(P, w.mult_y, w.mult_d, ((lc * PP + P.ig_np[q]) * F + v) * G + u0, tf);
    for (int q = P.ig_dp_off[k]; q < P.ig_dp_off[k + 1]; q++) {
      float m = mp_at(P, w.mult_y, w.mult_d, ((lc * PP + P.ig_dp[q]) * F + v) * G + u0, tf);
      if (fabsf(m) > 0) c /= (real)m;
    }
    w.coefS[i] = c;
  }
  if (P.mob_dyn)
    for (int i = tm.lane; i < G * nnz; i += tm.T) {
      const int g = i / nnz, k = i - g * nnz;
      const float v = mob_comb(P, w.mob_y, w.mob_g, g, k, tf);
      w.mxr[i] = v;
      w.mxc[g * nnz + __ldg(P.csc_of_csr + k)] = v;
    }
  // phase A: alive matrix (N excludes death and hospitalized, :28-41,:168) and weighted infectious fields
  for (int i = tm.lane; i < LG; i += tm.T) {
    real a = 0;
    for (int c = 0; c < P.C; c++) a += y[c * LG + i] * (real)__ldg(P.alive_coef + c);
    w.alive[i] = a;
    for (int k = 0; k < NG; k++) {
      real x = 0;
      for (int q = P.ig_w_off[k]; q < P.ig_w_off[k + 1]; q++) x += (real)__ldg(P.ig_w + q) * y[__ldg(P.ig_w_c2 + q) * LG + i];
      w.Xw[k * LG + i] = x;
    }
  }
  tm.sync();
  // ---- phase B: CSC gather over source locales (compute_foi_entry inner loop, :105-111)
  for (int i = tm.lane; i < LG; i += tm.T) {
    int j = i / G, u = i % G;
    real a = 0, xs[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) xs[k] = 0;
    for (int q = __ldg(P.csc_indptr + j); q < __ldg(P.csc_indptr + j + 1); q++) {
      int src = __ldg(P.csc_indices + q) * G + u;
      real val = (real)(P.mob_dyn ? w.mxc[u * nnz + q] : __ldg(P.mx_csc + u * nnz + q));
      a += w.alive[src] * val;
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) xs[k] += w.Xw[k * LG + src] * val;
    }
    w.Ag[i] = a;
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++)
      if (k < NG) w.Wg[k * LG + i] = xs[k];
  }
  tm.sync();
  // ---- phase C: G x G facility mixing -> force of infection -> s[k][j,u0] (:112-124, :146-157)
  for (int i = tm.lane; i < LG; i += tm.T) {
    int j = i / G, u0 = i % G, lc = __ldg(P.lclass + j);
    real sk[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) sk[k] = 0;
    for (int v = 0; v < F; v++) {
      int tb = ((lc * F + v) * G + u0) * G;
      real den = 0;
      if (P.has_ft_ops) {
        for (int u = 0; u < G; u++) den += w.Ag[j * G + u] * w.Tden[tb + u];
      } else if (sizeof(real) == 8) {
        for (int u = 0; u < G; u++) den += w.Ag[j * G + u] * (real)__ldg(P.Tden_const64 + tb + u);
      } else {
        for (int u = 0; u < G; u++) den += w.Ag[j * G + u] * (real)__ldg(P.Tden_const + tb + u);
      }
      if (den <= (real)1e-15) den = 1;
      real inv = (real)1 / den;
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) {
          real num = 0;
          for (int u = 0; u < G; u++) num += w.Wg[k * LG + j * G + u] * (real)w.Tnum[tb + u];
          real foi = sizeof(real) == 8 ? num / den : num * inv;
          sk[k] += w.coefS[((k * n_lc + lc) * F + v) * G + u0] * foi;
        }
    }
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++)
      if (k < NG) w.s[k * LG + i] = sk[k];
  }
  tm.sync();
  // ---- phase D: CSR gather over destination locales (:144-158)
  for (int i = tm.lane; i < LG; i += tm.T) {
    int i0 = i / G, u0 = i % G;
    real rs[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) rs[k] = 0;
    for (int q = __ldg(P.csr_indptr + i0); q < __ldg(P.csr_indptr + i0 + 1); q++) {
      int dst = __ldg(P.csr_indices + q) * G + u0;
      real val = (real)(P.mob_dyn ? w.mxr[u0 * nnz + q] : __ldg(P.mx_csr + u0 * nnz + q));
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) rs[k] += w.s[k * LG + dst] * val;
    }
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++)
      if (k < NG) w.r[k * LG + i] = rs[k];
  }
  tm.sync();
  // ---- phase E: flows of regular edges (compute_fraction, :43-65) and infectious groups
  const int nE = P.E;
  for (int it = tm.lane; it < (nE + NG) * LG; it += tm.T) {
    int src = it / LG, cell = it % LG;
    real f;
    if (src < nE) {
      int lc = __ldg(P.lclass + cell / G), g = cell % G;
      const int* ep = P.ep + __ldg(P.ep_off + src);
      auto val = [&](int fac) -> real {
        int kind = fac >> 24, idx = fac & 0xffffff;
        if (kind == 0) return (real)w.mpt0[(lc * PP + idx) * G + g];  // facility 0 only (:48)
        if (kind == 1) return w.alive[cell];
        return y[idx * LG + cell];
      };
      int n = __ldg(ep++);
      real numer = 1;
      for (int q = 0; q < n; q++) numer *= val(__ldg(ep++));
      int ns = __ldg(ep++);
      real denom = 0;
      for (int sidx = 0; sidx < ns; sidx++) {
        real sv = (real)__ldg(P.ep_coef + __ldg(ep++));
        int nf = __ldg(ep++);
        for (int q = 0; q < nf; q++) sv *= val(__ldg(ep++));
        denom += sv;
      }
      f = denom > 0 ? numer / denom : numer;
    } else {
      int k = src - nE;
      f = y[__ldg(P.ig_c0 + k) * LG + cell] * w.r[k * LG + cell];
    }
    w.flows[it] = f;
  }
  tm.sync();
  // ---- phase F: gather flows into dX (scatter maps of obj/edge.py:76-111 inverted)
  // The reference forms next = X + dX, applies the clamped moves to `next`, then returns next - X
  // (core_optimize.py:196-208). For the compartments a move touches this round trip is mirrored
  // literally: when a move is clamped the reference gets dX[c1] = 0 - X EXACTLY, and the near-exhausted
  // susceptible compartment is so stiff that an ulp-level deviation here grows to 1e-2 within the day.
  for (int it = tm.lane; it < P.CLG; it += tm.T) {
    int c = it / LG, cell = it % LG;
    real d = 0;
    for (int q = __ldg(P.xg_off + c); q < __ldg(P.xg_off + c + 1); q++)
      d += w.flows[__ldg(P.xg_src + q) * LG + cell] * (real)__ldg(P.xg_coef + q);
    Kout[it] = d;
  }
  // ---- phase G with general move entries (a MOVE op with targets in other locales): source cell by source cell, the
  // reference's sequential clamp over the cell's outgoing entries (next[c1,l1,g] shrinks entry by entry,
  // core_optimize.py:196-208); then every target cell adds what arrived, in list order. The engine admits only
  // entry lists whose result does not depend on the order between ops (build_plan).
  if (P.xmove) {
    tm.sync();
    for (int it = tm.lane; it < P.n_slot * LG; it += tm.T) {
      w.flows[(nE + NG) * LG + it] = 0;
      if (P.has_src64) w.KS[(size_t)stage * P.n_slot * LG + it] = (double)Kout[P.slot_c1[it / LG] * LG + it % LG];
    }
    tm.sync();
    for (int s_ = tm.lane; s_ < P.n_src; s_ += tm.T) {
      const int e0 = P.src_off[s_], e1 = P.src_off[s_ + 1], mo = P.src_mo[s_];
      if (e0 == e1 || !((ex_bits >> mo) & 0x10001)) continue;
      const bool has_y = (ex_bits >> mo) & 1, has_d = (ex_bits >> (16 + mo)) & 1;
      const int i1 = P.src_idx[s_], cell = i1 % LG, sl1 = P.src_sl[s_];
      const double y1 = P.has_src64 ? w.ysS[sl1 * LG + cell] : (double)y[i1];
      double n1 = y1 + (double)Kout[i1];
      for (int e = e0; e < e1; e++) {
        const float my = has_y ? w.mvY[e] : 0.0f, md = has_d ? w.mvD[e] : 0.0f;
        const float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        const double nv = (double)v < n1 ? (double)v : n1;
        n1 -= nv;
        w.xnv[e] = nv;
        if (P.ent_l1[e] == P.ent_l2[e]) w.flows[(nE + NG + P.ent_sl[e]) * LG + cell] = (real)nv;  // cumulative_move: same locale only
      }
      const double k1 = n1 - y1;
      Kout[i1] = (real)k1;
      if (P.has_src64) w.KS[((size_t)stage * P.n_slot + sl1) * LG + cell] = k1;
    }
    tm.sync();
    for (int t_ = tm.lane; t_ < P.n_tgt; t_ += tm.T) {
      const int i2 = P.tgt_idx[t_], cell = i2 % LG, sl2 = P.tgt_sl[t_];
      const double y2 = (P.has_src64 && sl2 >= 0) ? w.ysS[sl2 * LG + cell] : (double)y[i2];
      double n2 = y2 + (double)Kout[i2];
      bool any = false;
      for (int q = P.tgt_off[t_]; q < P.tgt_off[t_ + 1]; q++) {
        const int e = P.tgt_ent[q];
        if ((ex_bits >> P.ent_mo[e]) & 0x10001) { n2 += w.xnv[e]; any = true; }
      }
      if (any) {
        Kout[i2] = (real)(n2 - y2);
        if (P.has_src64 && sl2 >= 0) w.KS[((size_t)stage * P.n_slot + sl2) * LG + cell] = n2 - y2;
      }
    }
  } else
  // ---- phase G: clamped moves; each cell is owned by one thread, slots in op order
  if (P.n_slot > 0) {
    tm.sync();
    for (int cell = tm.lane; cell < LG; cell += tm.T) {
      for (int sl = 0; sl < P.n_slot; sl++) {
        real nv = 0;
        if (__ldg(P.slot_cover + sl * LG + cell) && ((ex_bits >> sl) & 0x10001)) {
          float my = ((ex_bits >> sl) & 1) ? w.mvY[sl * LG + cell] : 0.0f;
          float md = ((ex_bits >> (16 + sl)) & 1) ? w.mvD[sl * LG + cell] : 0.0f;
          float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
          int i1 = P.slot_c1[sl] * LG + cell, i2 = P.slot_c2[sl] * LG + cell;
          // next = X + dX, clamp, dX = next - X, in double in BOTH modes: a clamped move must give
          // dX[c1] = -X[c1] to full relative precision even when X[c1] << dX[c1]
          double y1 = P.has_src64 ? w.ysS[sl * LG + cell] : (double)y[i1], y2 = (double)y[i2];
          double n1 = y1 + (double)Kout[i1], n2 = y2 + (double)Kout[i2];
          double nvd = (double)v < n1 ? (double)v : n1;
          double k1 = (n1 - nvd) - y1;
          Kout[i1] = (real)k1;
          Kout[i2] = (real)((n2 + nvd) - y2);
          nv = (real)nvd;
          if (P.has_src64) w.KS[(stage * P.n_slot + sl) * LG + cell] = k1;
        } else if (P.has_src64) {
          w.KS[(stage * P.n_slot + sl) * LG + cell] = (double)Kout[P.slot_c1[sl] * LG + cell];
        }
        w.flows[(nE + NG + sl) * LG + cell] = nv;
      }
    }
  }
  tm.sync();
  // ---- phase H: live cumulative components (cumulative_move / gain / lost)
  for (int it = tm.lane; it < P.n_acc * LG; it += tm.T) {
    int a = it / LG, cell = it % LG;
    real d = 0;
    for (int q = __ldg(P.ag_off + a); q < __ldg(P.ag_off + a + 1); q++)
      d += w.flows[__ldg(P.ag_src + q) * LG + cell] * (real)__ldg(P.ag_coef + q);
    if (accStore) accStore[it] = d;
    if (accumulate) {
      w.accB[it] += bw * d;
      w.accE[it] += ew * d;
    }
  }
  tm.sync();
}

// ------------------------------------------------------------------------------------------
// day prologue: the intervention executor + parameter build, on the compiled op table
// ------------------------------------------------------------------------------------------
__device__ inline double eval_expr(const DevPlan& P, int e, const float* cpv, const double* sel) {
  double st[16];
  int sp = 0;
  for (int k = P.expr_off[e]; k < P.expr_off[e + 1]; k++) {
    int op = P.tok_op[k];
    double r;
    if (op == 0) r = P.tok_val[k];
    else if (op == 1) r = (double)cpv[P.tok_arg[k]];
    else if (op == 2) r = sel[P.tok_arg[k]];
    else if (op == 7) r = -st[--sp];
    else {
      double b = st[--sp], a = st[--sp];
      r = op == 3 ? __dadd_rn(a, b) : op == 4 ? __dsub_rn(a, b) : op == 5 ? __dmul_rn(a, b) : a / b;
    }
    if (P.tok_f32[k]) r = (double)(float)r;
    st[sp++] = r;
  }
  return st[0];
}

// is op `o` live today?  (its intervention is active and it belongs to the program variant in use)
__device__ __forceinline__ bool op_live(const DevPlan& P, const uint8_t* act, int variant, int o) {
  int itv = P.op_itv[o];
  return act[itv] && P.op_prog[o] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
}

// facility_interaction row of one side (multipliers m): product with the class multipliers, rows
// renormalised only if their f32 sum is > 1 or < 0 (get_normalized_facility_interaction,
// matrix/dense.py:21-32; numba's np.sum(axis=3) accumulates sequentially in f32)
__device__ inline void facility_row_fi(const DevPlan& P, const float* m, int row, float* out) {
  const int G = P.G;
  float s = 0.0f;
  for (int g2 = 0; g2 < G; g2++) s = __fadd_rn(s, __fmul_rn(P.fi0[row * G + g2], m[P.fi_cls[row * G + g2]]));
  bool norm = s > 1.0f || s < 0.0f;
  for (int g2 = 0; g2 < G; g2++) {
    float v = __fmul_rn(P.fi0[row * G + g2], m[P.fi_cls[row * G + g2]]);
    out[g2] = norm ? __fdiv_rn(v, s) : v;
  }
}

template <class real, bool CTA>
__device__ int prologue(const DevPlan& P, const Ws<real>& w, const Team<CTA>& tm, int variant, double* Xprev_g,
                        int ex_y_bits) {
  const int LG = P.LG, G = P.G, L = P.L;
  // 1. select sums (executor.py:346-359, ['Value'].sum())
  for (int s = 0; s < P.n_sel; s++) {
    double part = 0;
    int mode = P.sel_mode[s];
    for (int it = tm.lane; it < P.CLG; it += tm.T) {
      int c = it / LG, cell = it % LG;
      if (P.sel_c[s * P.C + c] && P.sel_l[s * L + cell / G] && P.sel_g[s * G + cell % G]) {
        double v;
        if (mode == 0) v = w.X[it];
        else if (mode == 1) v = P.X0[it];
        else v = w.X[it] - Xprev_g[P.chg_slot_of_comp[c] * LG + cell];
        part += v;
      }
    }
    double tot = tm.sum(part);
    if (tm.lane == 0) w.sel[s] = tot;
  }
  tm.sync();
  // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
  for (int it = tm.lane; it < P.CLG; it += tm.T) {
    int c = it / LG, sl = P.chg_slot_of_comp[c];
    if (sl >= 0) Xprev_g[sl * LG + it % LG] = w.X[it];
  }
  // 2. expressions
  for (int e = tm.lane; e < P.n_expr; e += tm.T) {
    int itv = P.expr_itv[e];
    bool live = w.act[itv] && P.expr_prog[e] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
    w.expr[e] = live ? eval_expr(P, e, w.cpv, w.sel) : 0.0;
  }
  tm.sync();
  // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
  for (int k = tm.lane; k < P.n_cls; k += tm.T) {
    float m = 1.0f;
    for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
      int o = P.cls_op[q];
      if (op_live(P, w.act, variant, o)) m = __fmul_rn(m, (float)w.expr[P.op_expr[o]]);
    }
    w.mult_d[k] = m;
  }
  tm.sync();
  // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
  for (int row = tm.lane; row < P.n_lc * P.F * G; row += tm.T) {
    float a[32], b[32];  // G <= 32 (checked on the host)
    facility_row_fi(P, w.mult_y, row, a);
    facility_row_fi(P, w.mult_d, row, b);
    for (int g2 = 0; g2 < G; g2++) {
      w.fi_y[row * G + g2] = a[g2];
      w.fi_gap[row * G + g2] = __fsub_rn(b[g2], a[g2]);
    }
  }
  for (int it = tm.lane; it < P.n_lc * G; it += tm.T) {
    int lc = it / G, g = it % G;
    float a[32], b[32];  // F <= 32 (checked on the host)
    // facility_timespent normalised over F (get_normalized_facility_timespent, dense.py:6-19)
    {
      const int F = P.F;
      float s = 0.0f, s2 = 0.0f;
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        a[f] = __fmul_rn(P.ft0[i], w.mult_y[P.ft_cls[i]]);
        b[f] = __fmul_rn(P.ft0[i], w.mult_d[P.ft_cls[i]]);
        s = __fadd_rn(s, a[f]);
        s2 = __fadd_rn(s2, b[f]);
      }
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        float ya = s > 1e-15f ? __fdiv_rn(a[f], s) : (f == 0 ? 1.0f : 0.0f);
        float yb = s2 > 1e-15f ? __fdiv_rn(b[f], s2) : (f == 0 ? 1.0f : 0.0f);
        w.ft_y[i] = ya;
        w.ft_gap[i] = __fsub_rn(yb, ya);
      }
    }
  }
  if (P.mob_dyn)  // the mode matrices of yesterday and today -> (y, gap)
    for (int it = tm.lane; it < P.M * G * L; it += tm.T) mob_row(P, w.mult_y, w.mult_d, it / (G * L), (it / L) % G, it % L, w.mob_y, w.mob_g);
  // 5. move table of today (nb_move, executor.py:166-234)
  for (int it = tm.lane; it < P.mv_len; it += tm.T) w.mvD[it] = 0.0f;
  int ex_d = 0;
  tm.sync();
  if (P.xmove) {
    // general move entries: all three branches of nb_move. Bit mo of the existence mask: the op added its entries today
    for (int mo = 0; mo < P.n_mv_op; mo++) {
      const int o = P.mv_op[mo];
      if (!op_live(P, w.act, variant, o)) continue;  // uniform across the team
      const double value = w.expr[P.op_expr[o]];
      double part = 0;
      for (int s_ = P.mo_src_off[mo] + tm.lane; s_ < P.mo_src_off[mo + 1]; s_ += tm.T) part += w.X[P.src_idx[s_]];
      const double depart_sum = tm.sum(part);
      if (depart_sum > 0) {
        for (int s_ = P.mo_src_off[mo] + tm.lane; s_ < P.mo_src_off[mo + 1]; s_ += tm.T) {
          const int e0 = P.src_off[s_], e1 = P.src_off[s_ + 1];
          if (e0 == e1) continue;
          const double depart = w.X[P.src_idx[s_]] / depart_sum * value;
          double arrive = 0;
          for (int e = e0; e < e1; e++) arrive += w.X[P.ent_c2[e] * LG + P.ent_l2[e] * G + P.ent_g[e]];
          for (int e = e0; e < e1; e++) {
            const double amt = arrive > 0 ? w.X[P.ent_c2[e] * LG + P.ent_l2[e] * G + P.ent_g[e]] / arrive * depart
                                          : 1.0 / (e1 - e0) * depart;
            w.mvD[e] = (float)(0.0 + amt);  // CooMatrix.element_add into a fresh f32 cell (coo.py:52-57)
          }
        }
        ex_d |= 1 << mo;
      }
      tm.sync();
    }
  } else
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    if (!op_live(P, w.act, variant, o)) continue;  // uniform across the team
    double value = w.expr[P.op_expr[o]];
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    double part = 0;
    for (int it = tm.lane; it < nc1 * LG; it += tm.T) {
      int cell = it % LG;
      if (P.mv_g[mo * G + cell % G] && P.mv_l1[mo * L + cell / G]) part += (double)w.X[c1s[it / LG] * LG + cell];
    }
    double depart_sum = tm.sum(part);
    if (depart_sum > 0) {
      for (int it = tm.lane; it < nc1 * LG; it += tm.T) {
        int cell = it % LG, c1 = c1s[it / LG];
        if (!(P.mv_g[mo * G + cell % G] && P.mv_l1[mo * L + cell / G])) continue;
        double depart = (double)w.X[c1 * LG + cell] / depart_sum * value;
        double arrive = 0;
        for (int q = 0; q < nc2; q++) arrive += (double)w.X[c2s[q] * LG + cell];
        for (int q = 0; q < nc2; q++) {
          double amt = arrive > 0 ? (double)w.X[c2s[q] * LG + cell] / arrive * depart : 1.0 / nc2 * depart;
          int sl = P.slot_of_pair[c1 * P.C + c2s[q]];
          float* cellp = &w.mvD[sl * LG + cell];
          *cellp = (float)((double)*cellp + amt);  // CooMatrix.element_add: f32 cell (coo.py:52-57)
        }
      }
      for (int q1 = 0; q1 < nc1; q1++)
        for (int q = 0; q < nc2; q++) ex_d |= 1 << P.slot_of_pair[c1s[q1] * P.C + c2s[q]];
    }
    tm.sync();
  }
  // 6. cost rates of today (nb_add, executor.py:236-243)
  for (int it = tm.lane; it < P.I * L; it += tm.T) {
    int i = it / L, l = it % L;
    double c = 0;
    for (int ao = 0; ao < P.n_add_op; ao++) {
      int o = P.add_op[ao];
      if (P.add_i[ao * P.I + i] && P.add_l[ao * L + l] && op_live(P, w.act, variant, o))
        c += w.expr[P.op_expr[o]] / (double)P.add_parts[ao];
    }
    w.crD[it] = c;
  }
  tm.sync();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: scipy RK45 (rk.py / common.py restated), returns the day's reward
// ------------------------------------------------------------------------------------------
template <class real, bool CTA>
__device__ double rk45_day(const DevPlan& P, const Ws<real>& w, const Team<CTA>& tm, int ex_bits, int* trace,
                           double* last_err_out, double* dbg) {
  int n_att = 0;
  const double rtol = 1e-3, atol = 1e-6;
  const int CLG = P.CLG, nA = P.n_acc * P.LG, nC = P.I * P.L;
  const double inv_n = 1.0 / (double)P.n_flat;
  real* K0 = w.K;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0;
  double last_err = 0;
  double t = 0.0;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };

  const int nS = P.has_src64 ? P.n_slot * P.LG : 0;  // double-carried source cells (fp32 mode)
  for (int i = tm.lane; i < CLG; i += tm.T) w.ys[i] = (real)w.X[i];
  for (int i = tm.lane; i < nS; i += tm.T) w.ysS[i] = w.X[P.slot_c1[i / P.LG] * P.LG + i % P.LG];
  tm.sync();
  rhs<real, CTA>(P, w, tm, 0.0, ex_bits, w.ys, K0, 0, w.accF, 0, 0, false);
  nfev++;
  // ---- select_initial_step (common.py:68-134)
  double h_abs;
  {
    double d0 = 0, d1 = 0;
    for (int i = tm.lane; i < CLG; i += tm.T) {
      double yv = (double)w.X[i], sc = atol + fabs(yv) * rtol;
      double a = yv / sc, b = (double)K0[i] / sc;
      d0 += a * a; d1 += b * b;
    }
    for (int i = tm.lane; i < nA; i += tm.T) {
      double yv = (double)w.accY[i], sc = atol + fabs(yv) * rtol;
      double a = yv / sc, b = (double)w.accF[i] / sc;
      d0 += a * a; d1 += b * b;
    }
    for (int i = tm.lane; i < nC; i += tm.T) {
      double yv = w.cost[i], sc = atol + fabs(yv) * rtol;
      double a = yv / sc, b = w.crY[i] / sc;  // q(0) = 0
      d0 += a * a; d1 += b * b;
    }
    d0 = sqrt(tm.sum(d0) * inv_n);
    d1 = sqrt(tm.sum(d1) * inv_n);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = fmin(h0, 1.0);
    for (int i = tm.lane; i < CLG; i += tm.T) w.ys[i] = (real)(w.X[i] + h0 * (double)K0[i]);
    for (int i = tm.lane; i < nS; i += tm.T) w.ysS[i] = w.X[P.slot_c1[i / P.LG] * P.LG + i % P.LG] + h0 * w.KS[i];
    tm.sync();
    real* K1 = w.K + CLG;
    rhs<real, CTA>(P, w, tm, h0, ex_bits, w.ys, K1, 1, w.accF2, 0, 0, false);
    nfev++;
    double d2 = 0, q0 = qd(h0);
    for (int i = tm.lane; i < CLG; i += tm.T) {
      double sc = atol + fabs((double)w.X[i]) * rtol, a = ((double)K1[i] - (double)K0[i]) / sc;
      d2 += a * a;
    }
    for (int i = tm.lane; i < nA; i += tm.T) {
      double sc = atol + fabs((double)w.accY[i]) * rtol, a = ((double)w.accF2[i] - (double)w.accF[i]) / sc;
      d2 += a * a;
    }
    for (int i = tm.lane; i < nC; i += tm.T) {
      double sc = atol + fabs(w.cost[i]) * rtol, a = ((w.crD[i] - w.crY[i]) * q0) / sc;
      d2 += a * a;
    }
    d2 = sqrt(tm.sum(d2) * inv_n) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 0.2);
    h_abs = fmin(fmin(100 * h0, h1), 1.0);
  }
  // ---- step loop (rk.py:111-176)
  int guard = 0;
  while (t < 1.0 && status == 0) {
    double min_step = 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t);  // nextafter(t, inf), t >= 0
    if (h_abs < min_step) h_abs = min_step;
    bool accepted = false, rejected = false;
    double t_new = t, h = 0;
    while (!accepted) {
      if (h_abs < min_step || ++guard > 10000 || !(h_abs == h_abs)) { status = -1; break; }
      h = h_abs;
      t_new = t + h;
      if (t_new - 1.0 > 0) t_new = 1.0;
      h = t_new - t;
      h_abs = fabs(h);
      const real hr = (real)h;
      for (int i = tm.lane; i < nA; i += tm.T) {
        real f = w.accF[i];
        w.accB[i] = (real)RK_B[0] * f;
        w.accE[i] = (real)RK_E[0] * f;
      }
      // stages 1..5 (rk_step, rk.py:14-71)
#pragma unroll 1
      for (int s = 1; s < 6; s++) {
        for (int i = tm.lane; i < CLG; i += tm.T) {
          real dy = 0;
          for (int j = 0; j < s; j++) dy += w.K[j * CLG + i] * (real)RK_A[s][j];
          w.ys[i] = (real)(w.X[i] + (double)dy * h);
        }
        for (int i = tm.lane; i < nS; i += tm.T) {
          double dy = 0;
          for (int j = 0; j < s; j++) dy += w.KS[j * nS + i] * RK_A[s][j];
          w.ysS[i] = w.X[P.slot_c1[i / P.LG] * P.LG + i % P.LG] + dy * h;
        }
        tm.sync();
        rhs<real, CTA>(P, w, tm, t + RK_C[s] * h, ex_bits, w.ys, w.K + s * CLG, s, nullptr, (real)RK_B[s], (real)RK_E[s], true);
      }
      for (int i = tm.lane; i < CLG; i += tm.T) {
        real dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) dy += w.K[j * CLG + i] * (real)RK_B[j];
        w.ys[i] = (real)(w.X[i] + h * (double)dy);  // y_new
      }
      for (int i = tm.lane; i < nS; i += tm.T) {
        double dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) dy += w.KS[j * nS + i] * RK_B[j];
        w.ysS[i] = w.X[P.slot_c1[i / P.LG] * P.LG + i % P.LG] + h * dy;
      }
      tm.sync();
      rhs<real, CTA>(P, w, tm, t + h, ex_bits, w.ys, w.K + 6 * CLG, 6, w.accF2, 0, (real)RK_E[6], true);
      nfev += 6;
      // error norm over ALL n_flat components (common.py:63-65, rk.py:146-148)
      double e2 = 0;
      for (int i = tm.lane; i < CLG; i += tm.T) {
        double sc = atol + fmax(fabs(w.X[i]), fabs((double)w.ys[i])) * rtol;
        real e = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) e += w.K[j * CLG + i] * (real)RK_E[j];
        double a = (double)e * h / sc;
        e2 += a * a;
      }
      for (int i = tm.lane; i < nA; i += tm.T) {
        double yv = (double)w.accY[i], yn = yv + h * (double)w.accB[i];
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = (double)w.accE[i] * h / sc;
        e2 += a * a;
      }
      double qs[7];
#pragma unroll
      for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
      for (int i = tm.lane; i < nC; i += tm.T) {
        double cy = w.crY[i], cg = w.crD[i] - cy, sb = 0, se = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          double k = cy + cg * qs[j];
          sb += k * RK_B[j];
          se += k * RK_E[j];
        }
        double yv = w.cost[i], yn = yv + h * sb;
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = se * h / sc;
        e2 += a * a;
      }
      double err = sqrt(tm.sum(e2) * inv_n);
      last_err = err;
      if (dbg && tm.lane == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
      n_att++;
      if (err < 1) {
        double factor = err == 0 ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
        if (rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        accepted = true;
      } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
        rejected = true;
        nrej++;
      }
      if (accepted) {
        // commit: y = y_new, f = f_new
        // the state itself is advanced in double so that float rounding of X (ulp(6e4) = 4e-3) does not
        // pile up over hundreds of days; in the fp64 mode this is the same arithmetic as y_new above
        for (int i = tm.lane; i < CLG; i += tm.T) {
          real dy = 0;
#pragma unroll
          for (int j = 0; j < 6; j++) dy += w.K[j * CLG + i] * (real)RK_B[j];
          w.X[i] = w.X[i] + h * (double)dy;
          w.K[i] = w.K[6 * CLG + i];
        }
        if (nS) tm.sync();
        for (int i = tm.lane; i < nS; i += tm.T) {
          w.X[P.slot_c1[i / P.LG] * P.LG + i % P.LG] = w.ysS[i];  // = X + h * sum_j B_j KS_j in double
          w.KS[i] = w.KS[6 * nS + i];
        }
        for (int i = tm.lane; i < nA; i += tm.T) {
          w.accY[i] += hr * w.accB[i];
          w.accF[i] = w.accF2[i];
        }
        for (int i = tm.lane; i < nC; i += tm.T) {
          double cy = w.crY[i], cg = w.crD[i] - cy, sb = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
          w.cost[i] += h * sb;
        }
      }
      tm.sync();
    }
    if (status != 0) break;
    nsteps++;
    t = t_new;
  }
  if (tm.lane == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
  return 0.0;
}

// ------------------------------------------------------------------------------------------
// the launch: EpiEnv.step for a batch (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
struct StepArgs {
  const float* actions;   // [N,A] device, or null
  const float* cpv;       // [N,NCP] device full control-parameter vectors (plan mode), or null
  const uint8_t* active;  // [N,I] device (plan mode), or null => all interventions active
  int variant;            // 0: programs for locale regex '*' (EpiEnv), 1: the schedule's day-0 programs
  int n_days;
  int init_only;          // 1: run the prologue only and store the result as "yesterday" (make_setting)
  int N;
  void* obs_out;          // real [N,CLG] or null
  double* reward_out;     // [N] or null
  uint8_t* done_out;      // [N] or null
  // second destination of the same results (epi_bind_mirror): the learner rank's receive buffer, mapped peer memory
  // over NVLink - the epilogue's stores ARE the gather, overlapped with the envs still being stepped
  void* obs_out2;
  double* reward_out2;
  uint8_t* done_out2;
  char* big_scratch;      // n_teams * big_bytes when !big_in_smem
  char* small_scratch;    // n_teams * small_bytes when !small_in_smem
  double* dbg_attempts;   // optional [N,64,2]: (h, error_norm) of each attempted RK step of the last day
  double* cr_scratch;     // env kernel: n_ctas * I*L doubles (today's cost rate)
  // env-group kernel (epi_grp.cuh): per-group scratch, per-CTA scratch, per-group arrive counters (zeroed per launch)
  char* grp_scratch;
  char* grp_cta_scratch;
  unsigned* grp_bar;
  // schedule mode (epi_run_plan; Epidemic.run, core/epidemic.py:118-134, for a batch of what-if schedules in ONE launch):
  // day d of the launch takes its control parameters / Act flags from row d (`cpv` / `active` point at row 0) and
  // emits the day's end state and reward
  const float* cpv_days;       // [n_days,N,NCP] or null
  const uint8_t* active_days;  // [n_days,N,I] or null (all interventions have an Act every day)
  void* obs_days;              // real [n_days,N,CLG] or null
  double* reward_days;         // [n_days,N] or null
  double* x_days;              // [n_days,N,CLG] the state in double (reporter input), or null
  void* acc_days;              // real [n_days,N,n_acc*LG] live cumulative components at the end of the day, or null
  float* mult_days;            // [n_days,N,n_cls] the day's class multipliers, or null
};

// one team, one env at a time: load state, prologue + RK45 for each day, store
template <class real, bool CTA>
__device__ void step_team(const DevPlan& P, const DevState& S, const StepArgs& a, const Team<CTA>& tm, int team,
                          int n_teams, char* small, char* big) {
  Ws<real> w;
  bind_ws<real>(P, small, big, w);
  const int CLG = P.CLG, LG = P.LG, nA = P.n_acc * LG, nC = P.I * P.L, nM = P.mv_len;

  for (int env = team; env < a.N; env += n_teams) {
    // ---- load the env's persistent state
    double* gX = S.X + (size_t)env * CLG;
    real* gAcc = (real*)S.acc + (size_t)env * nA;
    double* gXprev = S.Xprev + (size_t)env * P.n_chg * LG;
    int* meta = S.meta + (size_t)env * 8;
    for (int i = tm.lane; i < CLG; i += tm.T) w.X[i] = gX[i];
    for (int i = tm.lane; i < nA; i += tm.T) w.accY[i] = gAcc[i];
    for (int i = tm.lane; i < nC; i += tm.T) {
      w.cost[i] = S.cost[(size_t)env * nC + i];
      w.crY[i] = S.crY[(size_t)env * nC + i];
    }
    for (int i = tm.lane; i < nM; i += tm.T) w.mvY[i] = S.mvY[(size_t)env * nM + i];
    for (int i = tm.lane; i < P.n_cls; i += tm.T) w.mult_y[i] = S.multY[(size_t)env * P.n_cls + i];
    // control parameters (epipolicy_environment.py:40-54) and active flags
    for (int k = tm.lane; k < P.NCP; k += tm.T) {
      float v;
      if (a.cpv) v = a.cpv[(size_t)env * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[(size_t)env * P.A + s] : P.cp_fixed[k]; }
      w.cpv[k] = v;
    }
    for (int i = tm.lane; i < P.I; i += tm.T)
      w.act[i] = a.active ? a.active[(size_t)env * P.I + i] : (a.init_only ? P.init_active[i] : 1);
    int t_day = meta[0], ex_y = meta[1];
    tm.sync();

    double reward = 0;
    int done = t_day >= P.horizon;
    for (int day = 0; day < (a.init_only ? 1 : a.n_days); day++) {
      if (day > 0 && a.cpv_days) {  // schedule mode: today's row
        const size_t row = (size_t)day * a.N + env;
        for (int k = tm.lane; k < P.NCP; k += tm.T) w.cpv[k] = a.cpv_days[row * P.NCP + k];
        for (int i = tm.lane; i < P.I; i += tm.T) w.act[i] = a.active_days ? a.active_days[row * P.I + i] : 1;
        tm.sync();
      }
      int ex_bits = prologue<real, CTA>(P, w, tm, a.variant, gXprev, ex_y);
      double c0 = 0;
      if (!a.init_only) {
        for (int i = tm.lane; i < nC; i += tm.T) c0 += w.cost[i];
        rk45_day<real, CTA>(P, w, tm, ex_bits, meta, S.last_err + env, a.dbg_attempts ? a.dbg_attempts + (size_t)env * 128 : nullptr);
        double c1 = 0;
        for (int i = tm.lane; i < nC; i += tm.T) c1 += w.cost[i];
        // reward = -sum(cumulative_cost_next - cumulative_cost_current)  (epidemic.py:115)
        const double dr = tm.sum(c1 - c0);
        reward -= dr;
        t_day++;
        if (a.reward_days && tm.lane == 0) a.reward_days[(size_t)day * a.N + env] = -dr;
        const size_t row = (size_t)day * a.N + env;
        if (a.obs_days)
          for (int i = tm.lane; i < CLG; i += tm.T) ((real*)a.obs_days)[row * CLG + i] = (real)w.X[i];
        if (a.x_days)
          for (int i = tm.lane; i < CLG; i += tm.T) a.x_days[row * CLG + i] = w.X[i];
        if (a.acc_days)
          for (int i = tm.lane; i < nA; i += tm.T) ((real*)a.acc_days)[row * nA + i] = w.accY[i];
        if (a.mult_days)
          for (int i = tm.lane; i < P.n_cls; i += tm.T) a.mult_days[row * P.n_cls + i] = w.mult_d[i];
      }
      // today's parameters become yesterday's (State(t+1, next_obs, dest_parameter), epidemic.py:89)
      for (int i = tm.lane; i < P.n_cls; i += tm.T) w.mult_y[i] = w.mult_d[i];
      for (int i = tm.lane; i < nM; i += tm.T) w.mvY[i] = w.mvD[i];
      for (int i = tm.lane; i < nC; i += tm.T) w.crY[i] = w.crD[i];
      ex_y = (ex_bits >> 16) & 0xffff;
      tm.sync();
      done = t_day >= P.horizon;
      if (done) break;
    }
    // ---- store
    for (int i = tm.lane; i < CLG; i += tm.T) {
      double x = w.X[i];
      gX[i] = x;
      if (a.obs_out) ((real*)a.obs_out)[(size_t)env * CLG + i] = (real)x;
    }
    for (int i = tm.lane; i < nA; i += tm.T) gAcc[i] = w.accY[i];
    for (int i = tm.lane; i < nC; i += tm.T) {
      S.cost[(size_t)env * nC + i] = w.cost[i];
      S.crY[(size_t)env * nC + i] = w.crY[i];
    }
    for (int i = tm.lane; i < nM; i += tm.T) S.mvY[(size_t)env * nM + i] = w.mvY[i];
    for (int i = tm.lane; i < P.n_cls; i += tm.T) S.multY[(size_t)env * P.n_cls + i] = w.mult_y[i];
    if (tm.lane == 0) {
      meta[0] = t_day; meta[1] = ex_y;
      if (a.reward_out) a.reward_out[env] = reward;
      if (a.done_out) a.done_out[env] = (uint8_t)done;
    }
    tm.sync();
  }
}

#ifndef EPI_HOST_EMU
template <class real, bool CTA>
__global__ void __launch_bounds__(CTA ? 512 : 128)
step_kernel(const DevPlan P, const DevState S, const StepArgs a) {
  extern __shared__ __align__(16) char smem[];
  __shared__ double red[34];
  Team<CTA> tm;
  int team, n_teams;
  char* sm_base = smem;
  if (CTA) {
    tm.lane = threadIdx.x; tm.T = blockDim.x; team = blockIdx.x; n_teams = gridDim.x;
  } else {
    int wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    tm.lane = threadIdx.x & 31; tm.T = 32; team = blockIdx.x * wpb + wib; n_teams = gridDim.x * wpb;
    sm_base = smem + (size_t)wib * (size_t)((P.small_in_smem ? P.small_bytes : 0) + (P.big_in_smem ? P.big_bytes : 0));
  }
  tm.red = red;
  char* small = P.small_in_smem ? sm_base : a.small_scratch + (size_t)team * P.small_bytes;
  char* big = P.big_in_smem ? sm_base + (P.small_in_smem ? P.small_bytes : 0) : a.big_scratch + (size_t)team * P.big_bytes;
  step_team<real, CTA>(P, S, a, tm, team, n_teams, small, big);
}
#else
// host emulation (tests/emu): the same team code, one thread, heap workspace
template <class real, bool CTA>
void step_emulated(const DevPlan& P, const DevState& S, const StepArgs& a) {
  Team<CTA> tm;
  tm.lane = 0; tm.T = 1; tm.red = nullptr;
  char* small = (char*)calloc((size_t)P.small_bytes + 16, 1);
  char* big = (char*)calloc((size_t)P.big_bytes + 16, 1);
  step_team<real, CTA>(P, S, a, tm, 0, 1, small, big);
  free(small);
  free(big);
}
#endif

// reset: copy the detached initial record (computed once by an init_only launch) into the masked envs
template <class real>
__global__ void reset_kernel(const DevPlan P, const DevState S, const DevState init, const uint8_t* mask, int N,
                             void* obs_out) {
  const int CLG = P.CLG, LG = P.LG, nA = P.n_acc * LG, nC = P.I * P.L, nM = P.mv_len, nP = P.n_chg * LG;
  for (int env = blockIdx.x; env < N; env += gridDim.x) {
    if (mask && !mask[env]) continue;
    for (int i = threadIdx.x; i < CLG; i += blockDim.x) {
      double x = P.X0[i];
      S.X[(size_t)env * CLG + i] = x;
      if (obs_out) ((real*)obs_out)[(size_t)env * CLG + i] = (real)x;
    }
    for (int i = threadIdx.x; i < nA; i += blockDim.x) ((real*)S.acc)[(size_t)env * nA + i] = 0;
    for (int i = threadIdx.x; i < nC; i += blockDim.x) {
      S.cost[(size_t)env * nC + i] = 0;
      S.crY[(size_t)env * nC + i] = init.crY[i];
    }
    for (int i = threadIdx.x; i < nM; i += blockDim.x) S.mvY[(size_t)env * nM + i] = init.mvY[i];
    for (int i = threadIdx.x; i < nP; i += blockDim.x) {
      int c = -1;  // row -> compartment
      for (int cc = 0; cc < P.C; cc++) if (P.chg_slot_of_comp[cc] == i / LG) c = cc;
      S.Xprev[(size_t)env * nP + i] = P.X0[c * LG + i % LG];
    }
    for (int i = threadIdx.x; i < P.n_cls; i += blockDim.x) S.multY[(size_t)env * P.n_cls + i] = init.multY[i];
    for (int i = threadIdx.x; i < 8; i += blockDim.x) S.meta[(size_t)env * 8 + i] = i == 1 ? init.meta[1] : 0;
    if (threadIdx.x == 0) S.last_err[env] = 0;
  }
}

}  // namespace epi
