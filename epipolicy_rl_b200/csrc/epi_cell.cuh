// epi_cell.cuh — the "cell" kernel: one LANE per (env, locale, group) cell (sm_100a).
//
// Mapping (small scenarios, L*G <= 32: every scenario the reference ships):
//   * a warp carries EPW = 32 / (L*G) environments; lane -> (env slot, locale j, group u);
//   * each lane runs the whole per-cell program serially — every regular edge, the scatter into dX,
//     the cumulative-flow sums, the RK45 stage algebra — so the interpreted tables (edge programs,
//     gather lists) are walked with WARP-UNIFORM control flow and their cost is shared by 32 cells;
//   * the only cross-lane traffic is the force of infection: the CSC/CSR mobility gathers over the
//     env's locales and the G x G facility mixing over its groups, staged through shared memory,
//     plus one shared-memory reduction per RK45 attempt for the error norm;
//   * per-lane RK45 working set (state, 7 stage derivatives, flow sums) lives in shared memory as
//     [word][lane] columns (bank-conflict free); per-env tables (parameters at time t, facility
//     tables, cost rates) in a per-slot shared block.
// Envs in one warp run their own step-size sequences; the loop is warp-uniform ("any env still
// stepping") and lanes of finished envs idle through the barriers.
//
// The cumulative_move/gain/lost components the scipy error norm needs are linear in
// the per-stage flows: d acc_a = sum_q coef_q * flow[src_q]. Instead of 5 accumulator arrays the lane
// keeps FB = sum_s B_s flow_s and FE = sum_s E_s flow_s and gathers them once per attempt. A rejected
// attempt (0-2 per 365-day episode) re-evaluates the RHS at (t, y) to restore the stage-0 flows.
//
// Reference call stack and precision map: see epi_kernels.cuh (same arithmetic, same rounding points).
#pragma once
#include "epi_kernels.cuh"

namespace epi {
namespace cell {

#ifdef EPI_HOST_EMU
#define EPI_W (emu::warp_width)
#else
#define EPI_W 32
#endif
#define EPI_FULL 0xffffffffu
// dimensions: compile-time constants in a scenario-specialised build (generated header, see gen_spec)
#ifdef EPI_SPEC
#define DIM(x) (spec::x)
#define SLOT_C1(i) (spec::slot_c1[i])
#define SLOT_C2(i) (spec::slot_c2[i])
#else
#define DIM(x) (P.x)
#define SLOT_C1(i) (P.slot_c1[i])
#define SLOT_C2(i) (P.slot_c2[i])
#endif

template <class real>
struct Cx {
  int lane, slot, cell, j, u, lc, l0;  // l0: first lane of my env
  bool valid;                          // lane carries a cell of a slot < EPW
  // per-lane columns: element i of lane ln at base[i * EPI_W + ln]
  double* X; real *ys, *K, *fl, *FB, *FE, *cmA, *cmB;
  float *mvY, *mvD;
  double *ysS, *KS, *red;
  // per-env (slot) tables
  float *mult_y, *mult_d, *cpv;
  uint8_t* act;
  double *expr, *sel;
  float *mpy, *mpg, *mpt0;   // [n_lc,P,G] facility-0 parameters: yesterday, gap, at time t
  float *mpFy, *mpFg;        // [n_igp,n_lc,F,G] parameters used by infectious groups
  real* coefS;               // [NG,n_lc,F,G]
  float* Tnum;               // [n_lc,F,G,G]
  real* Tden;
  float *fi_y, *fi_gap, *ft_y, *ft_gap;
  double *cost, *crY, *crD;  // [I*L]
};

#define LA(arr, i) (c.arr[(i) * EPI_W + c.lane])
#define LO(arr, i, ln) (c.arr[(i) * EPI_W + (ln)])

template <class real>
__device__ __forceinline__ void bind(const DevPlan& P, char* warp_base, int lane, Cx<real>& c) {
  const CellLayout& Y = P.cl;
  const int LG = DIM(LG);
  c.lane = lane;
  c.slot = lane / LG;
  c.cell = lane - c.slot * LG;
  c.valid = c.slot < Y.epw;
  if (!c.valid) { c.slot = 0; c.cell = 0; }  // harmless aliases; every store of an invalid lane is predicated off
  c.j = c.cell / DIM(G);
  c.u = c.cell - c.j * DIM(G);
  c.lc = __ldg(P.lclass + c.j);
  c.l0 = c.slot * LG;
  c.X = (double*)(warp_base + Y.l_X); c.ys = (real*)(warp_base + Y.l_ys); c.K = (real*)(warp_base + Y.l_K);
  c.fl = (real*)(warp_base + Y.l_fl); c.FB = (real*)(warp_base + Y.l_FB); c.FE = (real*)(warp_base + Y.l_FE);
  c.cmA = (real*)(warp_base + Y.l_cmA); c.cmB = (real*)(warp_base + Y.l_cmB);
  c.mvY = (float*)(warp_base + Y.l_mvY); c.mvD = (float*)(warp_base + Y.l_mvD);
  c.ysS = (double*)(warp_base + Y.l_ysS); c.KS = (double*)(warp_base + Y.l_KS); c.red = (double*)(warp_base + Y.l_red);
  char* e = warp_base + Y.e_base + (int64_t)c.slot * Y.e_bytes;
  c.mult_y = (float*)(e + Y.e_mult_y); c.mult_d = (float*)(e + Y.e_mult_d); c.cpv = (float*)(e + Y.e_cpv);
  c.act = (uint8_t*)(e + Y.e_act); c.expr = (double*)(e + Y.e_expr); c.sel = (double*)(e + Y.e_sel);
  c.mpy = (float*)(e + Y.e_mpy); c.mpg = (float*)(e + Y.e_mpg); c.mpt0 = (float*)(e + Y.e_mpt0);
  c.mpFy = (float*)(e + Y.e_mpFy); c.mpFg = (float*)(e + Y.e_mpFg);
  c.coefS = (real*)(e + Y.e_coefS); c.Tnum = (float*)(e + Y.e_Tnum); c.Tden = (real*)(e + Y.e_Tden);
  c.fi_y = (float*)(e + Y.e_fi_y); c.fi_gap = (float*)(e + Y.e_fi_gap);
  c.ft_y = (float*)(e + Y.e_ft_y); c.ft_gap = (float*)(e + Y.e_ft_gap);
  c.cost = (double*)(e + Y.e_cost); c.crY = (double*)(e + Y.e_crY); c.crD = (double*)(e + Y.e_crD);
}

// sum over the lanes of my env, in lane order; every lane of the env gets the bit-identical result
template <class real>
__device__ __forceinline__ double team_sum(const DevPlan& P, const Cx<real>& c, double v, bool on) {
  c.red[c.lane] = on ? v : 0.0;
  __syncwarp();
  double s = 0;
  if (on)
    for (int i = 0; i < DIM(LG); i++) s += c.red[c.l0 + i];
  __syncwarp();
  return s;
}

// d acc_a from a flow column (cumulative_move / gain / lost derivative, obj/edge.py:76-111 inverted)
template <class real>
__device__ __forceinline__ real acc_gather(const DevPlan& P, const Cx<real>& c, const real* col, int a) {
  real d = 0;
  for (int q = __ldg(P.ag_off + a); q < __ldg(P.ag_off + a + 1); q++)
    d += col[__ldg(P.ag_src + q) * EPI_W + c.lane] * (real)__ldg(P.ag_coef + q);
  return d;
}

// all accumulator derivatives of a flow column: generated straight-line code under EPI_SPEC
#ifdef EPI_SPEC
#define ACC_DECL(name, col) real name[spec::n_acc > 0 ? spec::n_acc : 1]; spec::acc((col) + c.lane, name)
#define ACC_AT(name, col, a) (name[a])
#else
#define ACC_DECL(name, col)
#define ACC_AT(name, col, a) acc_gather<real>(P, c, col, a)
#endif

// ------------------------------------------------------------------------------------------
// my env's time-t tables (obj/parameter.py:80-92), shared by its lanes. When today's multipliers equal
// yesterday's (days 2..7 of a gym step) every gap is zero and the tables are filled once per day.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void fill_tables(const DevPlan& P, const Cx<real>& c, float tf) {
  const int G = DIM(G), F = DIM(F), LG = DIM(LG), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P);
  for (int i = c.cell; i < n_lc * PP * G; i += LG) c.mpt0[i] = lerp_rn(c.mpy[i], c.mpg[i], tf);
  for (int i = c.cell; i < n_lc * F * G * G; i += LG) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G);  // lv = lc*F + v
    float ft = lerp_rn(c.ft_y[lv * G + u], c.ft_gap[lv * G + u], tf);
    int a = (lv * G + u0) * G + u, b = (lv * G + u) * G + u0;
    float fa = lerp_rn(c.fi_y[a], c.fi_gap[a], tf), fb = lerp_rn(c.fi_y[b], c.fi_gap[b], tf);
    c.Tnum[i] = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (DIM(has_ft_ops)) c.Tden[i] = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + b);
  }
  const int FG = F * G, nFG = n_lc * FG;
  for (int i = c.cell; i < NG * nFG; i += LG) {
    int k = i / nFG, r = i - k * nFG;  // r = (lc*F + v)*G + u0
    real cf = (real)lerp_rn(c.ft_y[r], c.ft_gap[r], tf);
    int pi = __ldg(P.ig_p0i + k);
    cf *= (real)lerp_rn(c.mpFy[pi * nFG + r], c.mpFg[pi * nFG + r], tf);
    for (int q = __ldg(P.ig_np_off + k); q < __ldg(P.ig_np_off + k + 1); q++) {
      pi = __ldg(P.ig_npi + q);
      cf *= (real)lerp_rn(c.mpFy[pi * nFG + r], c.mpFg[pi * nFG + r], tf);
    }
    for (int q = __ldg(P.ig_dp_off + k); q < __ldg(P.ig_dp_off + k + 1); q++) {
      pi = __ldg(P.ig_dpi + q);
      float m = lerp_rn(c.mpFy[pi * nFG + r], c.mpFg[pi * nFG + r], tf);
      if (fabsf(m) > 0) cf /= (real)m;
    }
    c.coefS[i] = cf;
  }
}

// ------------------------------------------------------------------------------------------
// right-hand side for my cell (core_optimize.py:165-222). acc_mode: 0 none, 1 FB += bw*flow, FE += ew*flow,
// 2 FB = bw*flow, FE = ew*flow. `dyn`: my env's tables depend on t today. `keep_fl`: the caller reads the
// flow column afterwards (the generic path always writes it).
// Under EPI_SPEC the per-cell program (alive, weighted sums, edges, scatter, flow sums) is straight-line code
// generated from the plan (epi_engine.cu: gen_spec) working on registers; otherwise it is interpreted.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ void rhs(const DevPlan& P, const Cx<real>& c, bool on, bool dyn, double t, int ex_bits, real* __restrict__ Kout,
                    int stage, real bw, real ew, int acc_mode, bool keep_fl) {
  const int G = DIM(G), F = DIM(F), LG = DIM(LG), nnz = DIM(nnz), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P),
            C = DIM(C), nE = DIM(E);
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83
  real alive = 0;
#ifdef EPI_SPEC
  real y[spec::C], xw[spec::NG > 0 ? spec::NG : 1];
#endif
  if (on) {
    if (dyn) fill_tables<real>(P, c, tf);
    // ---- phase A: alive (N excludes death and hospitalized, :28-41,:168) and weighted infectious sums
#ifdef EPI_SPEC
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = LA(ys, k);
    spec::phaseA(y, alive, xw);
    LA(cmA, 0) = alive;
#pragma unroll
    for (int k = 0; k < NG; k++) LA(cmA, 1 + k) = xw[k];
#else
    for (int k = 0; k < C; k++) alive += LA(ys, k) * (real)__ldg(P.alive_coef + k);
    LA(cmA, 0) = alive;
    for (int k = 0; k < NG; k++) {
      real x = 0;
      for (int q = __ldg(P.ig_w_off + k); q < __ldg(P.ig_w_off + k + 1); q++)
        x += (real)__ldg(P.ig_w + q) * LA(ys, __ldg(P.ig_w_c2 + q));
      LA(cmA, 1 + k) = x;
    }
#endif
  }
  __syncwarp();
  // ---- phase B: CSC gather over source locales (compute_foi_entry inner loop, :105-111)
  if (on && NG > 0) {
    real a = 0, xs[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) xs[k] = 0;
    for (int q = __ldg(P.csc_indptr + c.j); q < __ldg(P.csc_indptr + c.j + 1); q++) {
      int src = c.l0 + __ldg(P.csc_indices + q) * G + c.u;
      real val = (real)__ldg(P.mx_csc + c.u * nnz + q);
      a += LO(cmA, 0, src) * val;
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) xs[k] += LO(cmA, 1 + k, src) * val;
    }
    LA(cmB, 0) = a;
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++)
      if (k < NG) LA(cmB, 1 + k) = xs[k];
  }
  __syncwarp();
  // ---- phase C: G x G facility mixing -> force of infection -> s[k] for my (j, u0)  (:112-124, :146-157)
  if (on && NG > 0) {
    real sk[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) sk[k] = 0;
    const int row = c.l0 + c.j * G;  // lanes of my locale
    for (int v = 0; v < F; v++) {
      int tb = ((c.lc * F + v) * G + c.u) * G;
      real den = 0;
      if (DIM(has_ft_ops)) {
        for (int u = 0; u < G; u++) den += LO(cmB, 0, row + u) * c.Tden[tb + u];
      } else if (sizeof(real) == 8) {
        for (int u = 0; u < G; u++) den += LO(cmB, 0, row + u) * (real)__ldg(P.Tden_const64 + tb + u);
      } else {
        for (int u = 0; u < G; u++) den += LO(cmB, 0, row + u) * (real)__ldg(P.Tden_const + tb + u);
      }
      if (den <= (real)1e-15) den = 1;
      real inv = (real)1 / den;
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) {
          real num = 0;
          for (int u = 0; u < G; u++) num += LO(cmB, 1 + k, row + u) * (real)c.Tnum[tb + u];
          real foi = sizeof(real) == 8 ? num / den : num * inv;
          sk[k] += c.coefS[((k * n_lc + c.lc) * F + v) * G + c.u] * foi;
        }
    }
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++)
      if (k < NG) LA(cmA, 1 + k) = sk[k];  // cmA[1..] is free again: phase B has been read out
  }
  __syncwarp();
  if (on) {
    // ---- phase D: CSR gather over destination locales (:144-158) -> the group flows
    real rs[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) rs[k] = 0;
    if (NG > 0) {
      for (int q = __ldg(P.csr_indptr + c.j); q < __ldg(P.csr_indptr + c.j + 1); q++) {
        int dst = c.l0 + __ldg(P.csr_indices + q) * G + c.u;
        real val = (real)__ldg(P.mx_csr + c.u * nnz + q);
#pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) rs[k] += LO(cmA, 1 + k, dst) * val;
      }
    }
    const float* mp = c.mpt0 + c.lc * PP * G + c.u;  // parameters from facility 0 only (:48)
#ifdef EPI_SPEC
    // ---- phases E-G on registers: flows, scatter into dX, clamped moves
    real f[spec::NSRC], d[spec::C];
    spec::flows(y, alive, mp, f);
#pragma unroll
    for (int k = 0; k < NG; k++) f[nE + k] = y[spec::ig_c0[k]] * rs[k];
    spec::dx(f, d);
#pragma unroll
    for (int sl = 0; sl < spec::n_slot; sl++) {
      real nv = 0;
      const int c1 = spec::slot_c1[sl], c2 = spec::slot_c2[sl];
      if (__ldg(P.slot_cover + sl * LG + c.cell) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? LA(mvY, sl) : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? LA(mvD, sl) : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double y1 = spec::has_src64 ? LA(ysS, sl) : (double)y[c1], y2 = (double)y[c2];
        double n1 = y1 + (double)d[c1], n2 = y2 + (double)d[c2];
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - y1;
        d[c1] = (real)k1;
        d[c2] = (real)((n2 + nvd) - y2);
        nv = (real)nvd;
        if (spec::has_src64) LA(KS, stage * spec::n_slot + sl) = k1;
      } else if (spec::has_src64) {
        LA(KS, stage * spec::n_slot + sl) = (double)d[c1];
      }
      f[nE + NG + sl] = nv;
    }
#pragma unroll
    for (int k = 0; k < C; k++) Kout[k * EPI_W + c.lane] = d[k];
    if (acc_mode == 1) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) {
        LA(FB, i) += bw * f[i];
        LA(FE, i) += ew * f[i];
      }
    } else if (acc_mode == 2) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) {
        LA(FB, i) = bw * f[i];
        LA(FE, i) = ew * f[i];
      }
    }
    if (keep_fl) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) LA(fl, i) = f[i];
    }
#else
    for (int k = 0; k < NG; k++) LA(fl, nE + k) = LA(ys, __ldg(P.ig_c0 + k)) * rs[k];
    // ---- phase E: regular edges (compute_fraction, :43-65)
    const int* ep = P.ep;
    for (int e = 0; e < nE; e++) {
      auto val = [&](int fac) -> real {
        int kind = fac >> 24, idx = fac & 0xffffff;
        if (kind == 0) return (real)mp[idx * G];
        if (kind == 1) return alive;
        return LA(ys, idx);
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
      LA(fl, e) = denom > 0 ? numer / denom : numer;
    }
    // ---- phase F: gather flows into dX
    for (int k = 0; k < C; k++) {
      real d = 0;
      for (int q = __ldg(P.xg_off + k); q < __ldg(P.xg_off + k + 1); q++)
        d += LA(fl, __ldg(P.xg_src + q)) * (real)__ldg(P.xg_coef + q);
      Kout[k * EPI_W + c.lane] = d;
    }
    // ---- phase G: clamped moves, slots in op order. next = X + dX, clamp, dX = next - X in double in BOTH
    // modes: a clamped move must give dX[c1] = -X[c1] to full relative precision (core_optimize.py:196-208)
    for (int sl = 0; sl < P.n_slot; sl++) {
      real nv = 0;
      const int c1 = SLOT_C1(sl), c2 = SLOT_C2(sl);
      if (__ldg(P.slot_cover + sl * LG + c.cell) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? LA(mvY, sl) : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? LA(mvD, sl) : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double y1 = P.has_src64 ? LA(ysS, sl) : (double)LA(ys, c1), y2 = (double)LA(ys, c2);
        double n1 = y1 + (double)Kout[c1 * EPI_W + c.lane], n2 = y2 + (double)Kout[c2 * EPI_W + c.lane];
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - y1;
        Kout[c1 * EPI_W + c.lane] = (real)k1;
        Kout[c2 * EPI_W + c.lane] = (real)((n2 + nvd) - y2);
        nv = (real)nvd;
        if (P.has_src64) LA(KS, stage * P.n_slot + sl) = k1;
      } else if (P.has_src64) {
        LA(KS, stage * P.n_slot + sl) = (double)Kout[c1 * EPI_W + c.lane];
      }
      LA(fl, nE + NG + sl) = nv;
    }
    // ---- running flow sums for the cumulative components
    if (acc_mode == 1) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = LA(fl, i);
        LA(FB, i) += bw * f;
        LA(FE, i) += ew * f;
      }
    } else if (acc_mode == 2) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = LA(fl, i);
        LA(FB, i) = bw * f;
        LA(FE, i) = ew * f;
      }
    }
#endif
  }
  __syncwarp();  // the env tables are rewritten by the next evaluation
}

// ------------------------------------------------------------------------------------------
// day prologue for my env (Executor.execute + get_parameter_from_delta + get_gap_parameter)
// ------------------------------------------------------------------------------------------
template <class real>
__device__ int prologue(const DevPlan& P, const Cx<real>& c, bool on, int variant, double* Xprev_g, int ex_y_bits,
                        bool& dyn) {
  const int LG = DIM(LG), G = DIM(G), L = DIM(L), C = DIM(C), F = DIM(F), PP = DIM(P);
  // 1. select sums (executor.py:346-359, ['Value'].sum())
  for (int s = 0; s < P.n_sel; s++) {
    double part = 0;
    if (on && P.sel_l[s * L + c.j] && P.sel_g[s * G + c.u]) {
      int mode = P.sel_mode[s];
      for (int k = 0; k < C; k++)
        if (P.sel_c[s * C + k]) {
          if (mode == 0) part += LA(X, k);
          else if (mode == 1) part += P.X0[k * LG + c.cell];
          else part += LA(X, k) - Xprev_g[P.chg_slot_of_comp[k] * LG + c.cell];
        }
    }
    double tot = team_sum(P, c, part, on);
    if (on && c.cell == 0) c.sel[s] = tot;
  }
  __syncwarp();
  if (on) {
    // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
    for (int k = 0; k < C; k++) {
      int sl = P.chg_slot_of_comp[k];
      if (sl >= 0) Xprev_g[sl * LG + c.cell] = LA(X, k);
    }
    // 2. expressions
    for (int e = c.cell; e < P.n_expr; e += LG) {
      int itv = P.expr_itv[e];
      bool live = c.act[itv] && P.expr_prog[e] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
      c.expr[e] = live ? eval_expr(P, e, c.cpv, c.sel) : 0.0;
    }
  }
  __syncwarp();
  if (on) {
    // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
    for (int k = c.cell; k < P.n_cls; k += LG) {
      float m = 1.0f;
      for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
        int o = P.cls_op[q];
        if (op_live(P, c.act, variant, o)) m = __fmul_rn(m, (float)c.expr[P.op_expr[o]]);
      }
      c.mult_d[k] = m;
    }
  }
  __syncwarp();
  {
    // do today's multipliers differ from yesterday's? (if not, every parameter gap is zero)
    double diff = 0;
    if (on)
      for (int k = c.cell; k < P.n_cls; k += LG) diff += c.mult_d[k] != c.mult_y[k] ? 1.0 : 0.0;
    dyn = team_sum(P, c, diff, on) > 0;
  }
  int ex_d = 0;
  if (on) {
    // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
    for (int row = c.cell; row < P.n_lc * F * G; row += LG) {
      float a[32], b[32];  // G <= 32 (checked on the host)
      facility_row_fi(P, c.mult_y, row, a);
      facility_row_fi(P, c.mult_d, row, b);
      for (int g2 = 0; g2 < G; g2++) {
        c.fi_y[row * G + g2] = a[g2];
        c.fi_gap[row * G + g2] = __fsub_rn(b[g2], a[g2]);
      }
    }
    for (int it = c.cell; it < P.n_lc * G; it += LG) {
      int lc = it / G, g = it % G;
      float a[32], b[32];  // F <= 32 (checked on the host)
      // facility_timespent normalised over F (get_normalized_facility_timespent, dense.py:6-19)
      float s = 0.0f, s2 = 0.0f;
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        a[f] = __fmul_rn(P.ft0[i], c.mult_y[P.ft_cls[i]]);
        b[f] = __fmul_rn(P.ft0[i], c.mult_d[P.ft_cls[i]]);
        s = __fadd_rn(s, a[f]);
        s2 = __fadd_rn(s2, b[f]);
      }
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        float ya = s > 1e-15f ? __fdiv_rn(a[f], s) : (f == 0 ? 1.0f : 0.0f);
        float yb = s2 > 1e-15f ? __fdiv_rn(b[f], s2) : (f == 0 ? 1.0f : 0.0f);
        c.ft_y[i] = ya;
        c.ft_gap[i] = __fsub_rn(yb, ya);
      }
    }
    // model parameters: default * class multiplier of yesterday / today (parameter.py:62,:72)
    for (int i = c.cell; i < P.n_lc * PP * G; i += LG) {
      int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
      int idx = ((lc * PP + p) * F + 0) * G + g;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, c.mult_y[cls]), d = __fmul_rn(m0, c.mult_d[cls]);
      c.mpy[i] = y;
      c.mpg[i] = __fsub_rn(d, y);
    }
    const int nFG = P.n_lc * F * G;
    for (int i = c.cell; i < P.n_igp * nFG; i += LG) {
      int pi = i / nFG, r = i - pi * nFG, lc = r / (F * G), fg = r - lc * F * G;
      int idx = (lc * PP + __ldg(P.igp + pi)) * F * G + fg;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, c.mult_y[cls]), d = __fmul_rn(m0, c.mult_d[cls]);
      c.mpFy[i] = y;
      c.mpFg[i] = __fsub_rn(d, y);
    }
    // 5. move table of today (nb_move, executor.py:166-234; different compartment, same locale)
    for (int sl = 0; sl < DIM(n_slot); sl++) LA(mvD, sl) = 0.0f;
  }
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    bool live = on && op_live(P, c.act, variant, o);
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    bool mine = live && P.mv_g[mo * G + c.u] && P.mv_l1[mo * L + c.j];
    double part = 0;
    if (mine)
      for (int q = 0; q < nc1; q++) part += LA(X, c1s[q]);
    double depart_sum = team_sum(P, c, part, on);
    if (live && depart_sum > 0) {
      double value = c.expr[P.op_expr[o]];
      if (mine) {
        for (int q1 = 0; q1 < nc1; q1++) {
          int c1 = c1s[q1];
          double depart = LA(X, c1) / depart_sum * value;
          double arrive = 0;
          for (int q = 0; q < nc2; q++) arrive += LA(X, c2s[q]);
          for (int q = 0; q < nc2; q++) {
            double amt = arrive > 0 ? LA(X, c2s[q]) / arrive * depart : 1.0 / nc2 * depart;
            int sl = P.slot_of_pair[c1 * C + c2s[q]];
            float* cellp = &LA(mvD, sl);
            *cellp = (float)((double)*cellp + amt);  // CooMatrix.element_add: f32 cell (coo.py:52-57)
          }
        }
      }
      for (int q1 = 0; q1 < nc1; q1++)
        for (int q = 0; q < nc2; q++) ex_d |= 1 << P.slot_of_pair[c1s[q1] * C + c2s[q]];
    }
  }
  if (on) {
    // 6. cost rates of today (nb_add, executor.py:236-243)
    for (int it = c.cell; it < P.I * L; it += LG) {
      int i = it / L, l = it % L;
      double cr = 0;
      for (int ao = 0; ao < P.n_add_op; ao++) {
        int o = P.add_op[ao];
        if (P.add_i[ao * P.I + i] && P.add_l[ao * L + l] && op_live(P, c.act, variant, o))
          cr += c.expr[P.op_expr[o]] / (double)P.add_parts[ao];
      }
      c.crD[it] = cr;
    }
  }
  __syncwarp();
  if (on && !dyn) fill_tables<real>(P, c, 0.0f);  // constant over the day: filled once (the final barrier of rhs orders it)
  __syncwarp();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: scipy RK45 (rk.py / common.py restated) for the envs of this warp
// ------------------------------------------------------------------------------------------
template <class real>
__device__ void rk45_day(const DevPlan& P, const Cx<real>& c, bool on, bool dyn, int ex_bits, real* accY_g, int* trace,
                         double* last_err_out, double* dbg) {
  const double rtol = 1e-3, atol = 1e-6;
  const int C = DIM(C), LG = DIM(LG), nC = DIM(I) * DIM(L), nS = DIM(has_src64) ? DIM(n_slot) : 0, NSRC = DIM(NSRC),
            n_acc = DIM(n_acc);
  const double inv_n = 1.0 / (double)P.n_flat;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0, n_att = 0;
  double last_err = 0, t = 0.0, h_abs = 0;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };
  real* K0 = c.K;
  real* K1 = c.K + C * EPI_W;

  if (on) {
    for (int k = 0; k < C; k++) LA(ys, k) = (real)LA(X, k);
    for (int s = 0; s < nS; s++) LA(ysS, s) = LA(X, SLOT_C1(s));
  }
  rhs<real>(P, c, on, dyn, 0.0, ex_bits, K0, 0, 0, 0, 0, true);
  nfev++;
  // ---- select_initial_step (common.py:68-134)
  {
    double d0 = 0, d1 = 0;
    if (on) {
      for (int k = 0; k < C; k++) {
        double yv = LA(X, k), sc = atol + fabs(yv) * rtol;
        double a = yv / sc, b = (double)K0[k * EPI_W + c.lane] / sc;
        d0 += a * a; d1 += b * b;
      }
      ACC_DECL(g0, c.fl);
#pragma unroll
      for (int a_ = 0; a_ < n_acc; a_++) {
        double yv = (double)accY_g[a_ * LG], sc = atol + fabs(yv) * rtol;
        double a = yv / sc, b = (double)ACC_AT(g0, c.fl, a_) / sc;
        d0 += a * a; d1 += b * b;
      }
      for (int i = c.cell; i < nC; i += LG) {
        double yv = c.cost[i], sc = atol + fabs(yv) * rtol;
        double a = yv / sc, b = c.crY[i] / sc;  // q(0) = 0
        d0 += a * a; d1 += b * b;
      }
      for (int i = 0; i < NSRC; i++) LA(FE, i) = LA(fl, i);  // stage-0 flows, needed once more below
    }
    d0 = sqrt(team_sum(P, c, d0, on) * inv_n);
    d1 = sqrt(team_sum(P, c, d1, on) * inv_n);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = fmin(h0, 1.0);
    if (on) {
      for (int k = 0; k < C; k++) LA(ys, k) = (real)(LA(X, k) + h0 * (double)K0[k * EPI_W + c.lane]);
      for (int s = 0; s < nS; s++) LA(ysS, s) = LA(X, SLOT_C1(s)) + h0 * LA(KS, s);
    }
    rhs<real>(P, c, on, dyn, h0, ex_bits, K1, 1, 0, 0, 0, true);
    nfev++;
    double d2 = 0, q0 = qd(h0);
    if (on) {
      for (int k = 0; k < C; k++) {
        double sc = atol + fabs(LA(X, k)) * rtol, a = ((double)K1[k * EPI_W + c.lane] - (double)K0[k * EPI_W + c.lane]) / sc;
        d2 += a * a;
      }
      ACC_DECL(g1, c.fl);
      ACC_DECL(g0, c.FE);
#pragma unroll
      for (int a_ = 0; a_ < n_acc; a_++) {
        double sc = atol + fabs((double)accY_g[a_ * LG]) * rtol;
        double a = ((double)ACC_AT(g1, c.fl, a_) - (double)ACC_AT(g0, c.FE, a_)) / sc;
        d2 += a * a;
      }
      for (int i = c.cell; i < nC; i += LG) {
        double sc = atol + fabs(c.cost[i]) * rtol, a = ((c.crD[i] - c.crY[i]) * q0) / sc;
        d2 += a * a;
      }
      // stage-0 contributions of the first attempt
      for (int i = 0; i < NSRC; i++) {
        real f = LA(FE, i);
        LA(FB, i) = (real)RK_B[0] * f;
        LA(FE, i) = (real)RK_E[0] * f;
      }
    }
    d2 = sqrt(team_sum(P, c, d2, on) * inv_n) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 0.2);
    h_abs = fmin(fmin(100 * h0, h1), 1.0);
  }
  // ---- step loop (rk.py:111-176), one ATTEMPT per iteration, warp-uniform
  bool run = on, rejected = false;
  int guard = 0;
  while (__any_sync(EPI_FULL, run)) {
    double h = 0, t_new = t;
    if (run) {
      double min_step = 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t);  // nextafter(t, inf), t >= 0
      if (h_abs < min_step) {
        if (rejected) { status = -1; run = false; }  // step size too small after a rejection
        else h_abs = min_step;
      }
      if (++guard > 10000 || !(h_abs == h_abs)) { status = -1; run = false; }
    }
    if (run) {
      h = h_abs;
      t_new = t + h;
      if (t_new - 1.0 > 0) t_new = 1.0;
      h = t_new - t;
      h_abs = fabs(h);
    }
    // stages 1..5 (rk_step, rk.py:14-71)
#pragma unroll 1
    for (int s = 1; s < 6; s++) {
      if (run) {
        for (int k = 0; k < C; k++) {
          real dy = 0;
          for (int j = 0; j < s; j++) dy += c.K[(j * C + k) * EPI_W + c.lane] * (real)RK_A[s][j];
          LA(ys, k) = (real)(LA(X, k) + (double)dy * h);
        }
        for (int q = 0; q < nS; q++) {
          double dy = 0;
          for (int j = 0; j < s; j++) dy += LA(KS, j * nS + q) * RK_A[s][j];
          LA(ysS, q) = LA(X, SLOT_C1(q)) + dy * h;
        }
      }
      rhs<real>(P, c, run, dyn, t + RK_C[s] * h, ex_bits, c.K + s * C * EPI_W, s, (real)RK_B[s], (real)RK_E[s], 1, false);
    }
    if (run) {
      for (int k = 0; k < C; k++) {
        real dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) dy += c.K[(j * C + k) * EPI_W + c.lane] * (real)RK_B[j];
        LA(ys, k) = (real)(LA(X, k) + h * (double)dy);  // y_new
      }
      for (int q = 0; q < nS; q++) {
        double dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) dy += LA(KS, j * nS + q) * RK_B[j];
        LA(ysS, q) = LA(X, SLOT_C1(q)) + h * dy;
      }
    }
    rhs<real>(P, c, run, dyn, t + h, ex_bits, c.K + 6 * C * EPI_W, 6, 0, (real)RK_E[6], 1, true);
    // error norm over ALL n_flat components (common.py:63-65, rk.py:146-148)
    double e2 = 0;
    double qs[7];
    if (run) {
      nfev += 6;
      for (int k = 0; k < C; k++) {
        double sc = atol + fmax(fabs(LA(X, k)), fabs((double)LA(ys, k))) * rtol;
        real e = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) e += c.K[(j * C + k) * EPI_W + c.lane] * (real)RK_E[j];
        double a = (double)e * h / sc;
        e2 += a * a;
      }
      ACC_DECL(gB, c.FB);
      ACC_DECL(gE, c.FE);
#pragma unroll
      for (int a_ = 0; a_ < n_acc; a_++) {
        double yv = (double)accY_g[a_ * LG], yn = yv + h * (double)ACC_AT(gB, c.FB, a_);
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = (double)ACC_AT(gE, c.FE, a_) * h / sc;
        e2 += a * a;
      }
#pragma unroll
      for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
      for (int i = c.cell; i < nC; i += LG) {
        double cy = c.crY[i], cg = c.crD[i] - cy, sb = 0, se = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          double k = cy + cg * qs[j];
          sb += k * RK_B[j];
          se += k * RK_E[j];
        }
        double yv = c.cost[i], yn = yv + h * sb;
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = se * h / sc;
        e2 += a * a;
      }
    }
    double err = sqrt(team_sum(P, c, e2, run) * inv_n);
    bool accepted = false;
    if (run) {
      last_err = err;
      if (dbg && c.cell == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
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
    }
    if (run && accepted) {
      // commit: y = y_new, f = f_new. The state itself is advanced in double so that float rounding of X
      // does not pile up over hundreds of days; in the fp64 mode this is the same arithmetic as y_new above
      const real hr = (real)h;
      for (int k = 0; k < C; k++) {
        real dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) dy += c.K[(j * C + k) * EPI_W + c.lane] * (real)RK_B[j];
        LA(X, k) = LA(X, k) + h * (double)dy;
        K0[k * EPI_W + c.lane] = c.K[(6 * C + k) * EPI_W + c.lane];
      }
      for (int q = 0; q < nS; q++) {
        LA(X, SLOT_C1(q)) = LA(ysS, q);  // = X + h * sum_j B_j KS_j in double
        LA(KS, q) = LA(KS, 6 * nS + q);
      }
      {
        ACC_DECL(gB, c.FB);
#pragma unroll
        for (int a_ = 0; a_ < n_acc; a_++) accY_g[a_ * LG] += hr * ACC_AT(gB, c.FB, a_);
      }
      for (int i = 0; i < NSRC; i++) {  // the last stage's flows open the next attempt (FSAL)
        real f = LA(fl, i);
        LA(FB, i) = (real)RK_B[0] * f;
        LA(FE, i) = (real)RK_E[0] * f;
      }
      for (int i = c.cell; i < nC; i += LG) {
        double cy = c.crY[i], cg = c.crD[i] - cy, sb = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
        c.cost[i] += h * sb;
      }
      nsteps++;
      t = t_new;
      rejected = false;
      if (!(t < 1.0)) run = false;
    }
    // a rejected attempt: restore the stage-0 flow sums by re-evaluating f(t, y) (identical arithmetic to
    // the evaluation that produced K0; not counted in nfev)
    bool redo = run && !accepted;
    if (__any_sync(EPI_FULL, redo)) {
      if (redo) {
        for (int k = 0; k < C; k++) LA(ys, k) = (real)LA(X, k);
        for (int q = 0; q < nS; q++) LA(ysS, q) = LA(X, SLOT_C1(q));
      }
      rhs<real>(P, c, redo, dyn, t, ex_bits, K0, 0, (real)RK_B[0], (real)RK_E[0], 2, false);
    }
  }
  if (on && c.cell == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
}

// ------------------------------------------------------------------------------------------
// the launch: EpiEnv.step for a batch (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
template <class real>
__device__ void step_warp(const DevPlan& P, const DevState& S, const StepArgs& a, char* warp_base, int lane, int group) {
  Cx<real> c;
  bind<real>(P, warp_base, lane, c);
  const int C = DIM(C), CLG = DIM(CLG), LG = DIM(LG), nA = DIM(n_acc) * LG, nC = DIM(I) * DIM(L), nM = DIM(n_slot) * LG;
  const int env = group * P.cl.epw + c.slot;
  const bool on0 = c.valid && env < a.N;
  const size_t e = on0 ? (size_t)env : 0;

  // ---- load the env's persistent state
  double* gX = S.X + e * CLG;
  real* gAcc = (real*)S.acc + e * nA + c.cell;  // accumulators stay in HBM/L2: touched once per RK45 attempt
  double* gXprev = S.Xprev + e * P.n_chg * LG;
  int* meta = S.meta + e * 8;
  if (on0) {
    for (int k = 0; k < C; k++) LA(X, k) = gX[k * LG + c.cell];
    for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = S.mvY[e * nM + s * LG + c.cell];
    for (int i = c.cell; i < nC; i += LG) {
      c.cost[i] = S.cost[e * nC + i];
      c.crY[i] = S.crY[e * nC + i];
    }
    for (int i = c.cell; i < P.n_cls; i += LG) c.mult_y[i] = S.multY[e * P.n_cls + i];
    // control parameters (epipolicy_environment.py:40-54) and active flags
    for (int k = c.cell; k < P.NCP; k += LG) {
      float v;
      if (a.cpv) v = a.cpv[e * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[e * P.A + s] : P.cp_fixed[k]; }
      c.cpv[k] = v;
    }
    for (int i = c.cell; i < P.I; i += LG)
      c.act[i] = a.active ? a.active[e * P.I + i] : (a.init_only ? P.init_active[i] : 1);
  }
  int t_day = on0 ? meta[0] : 0, ex_y = on0 ? meta[1] : 0;
  __syncwarp();

  double reward = 0;
  bool on = on0;
  const int n_days = a.init_only ? 1 : a.n_days;
  for (int day = 0; day < n_days; day++) {
    if (!__any_sync(EPI_FULL, on)) break;
    bool dyn = true;
    int ex_bits = prologue<real>(P, c, on, a.variant, gXprev, ex_y, dyn);
    if (!a.init_only) {
      double c0 = 0, c1 = 0;
      if (on)
        for (int i = c.cell; i < nC; i += LG) c0 += c.cost[i];
      rk45_day<real>(P, c, on, dyn, ex_bits, gAcc, meta, S.last_err + e,
                     a.dbg_attempts ? a.dbg_attempts + e * 128 : nullptr);
      if (on)
        for (int i = c.cell; i < nC; i += LG) c1 += c.cost[i];
      // reward = -sum(cumulative_cost_next - cumulative_cost_current)  (epidemic.py:115)
      double dr = team_sum(P, c, c1 - c0, on);
      if (on) { reward -= dr; t_day++; }
    }
    // today's parameters become yesterday's (State(t+1, next_obs, dest_parameter), epidemic.py:89)
    if (on) {
      for (int i = c.cell; i < P.n_cls; i += LG) c.mult_y[i] = c.mult_d[i];
      for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = LA(mvD, s);
      for (int i = c.cell; i < nC; i += LG) c.crY[i] = c.crD[i];
      ex_y = (ex_bits >> 16) & 0xffff;
    }
    __syncwarp();
    if (on && t_day >= P.horizon) on = false;  // the reference's EpiEnv.step stops its day loop at done
  }
  // ---- store
  if (on0) {
    for (int k = 0; k < C; k++) {
      double x = LA(X, k);
      gX[k * LG + c.cell] = x;
      if (a.obs_out) ((real*)a.obs_out)[e * CLG + k * LG + c.cell] = (real)x;
    }
    for (int s = 0; s < DIM(n_slot); s++) S.mvY[e * nM + s * LG + c.cell] = LA(mvY, s);
    for (int i = c.cell; i < nC; i += LG) {
      S.cost[e * nC + i] = c.cost[i];
      S.crY[e * nC + i] = c.crY[i];
    }
    for (int i = c.cell; i < P.n_cls; i += LG) S.multY[e * P.n_cls + i] = c.mult_y[i];
    if (c.cell == 0) {
      meta[0] = t_day; meta[1] = ex_y;
      if (a.reward_out) a.reward_out[env] = reward;
      if (a.done_out) a.done_out[env] = (uint8_t)(t_day >= P.horizon);
    }
  }
}

#ifndef EPI_HOST_EMU
template <class real>
__global__ void __launch_bounds__(32) cell_kernel(const DevPlan P, const DevState S, const StepArgs a) {
  extern __shared__ __align__(16) char smem[];
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  step_warp<real>(P, S, a, smem + (size_t)wib * P.cl.warp_bytes, threadIdx.x & 31, blockIdx.x * wpb + wib);
}
#endif

#undef LA
#undef LO
}  // namespace cell
}  // namespace epi
