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
// Envs in one warp run their own step-size sequences; the day is a warp-uniform state machine with ONE
// right-hand-side call site (initial step selection, 6 stages per attempt, the rare re-evaluation after
// a rejection) and lanes of finished envs idle through the barriers.
//
// The cumulative_move/gain/lost components the scipy error norm needs are linear in
// the per-stage flows: d acc_a = sum_q coef_q * flow[src_q]. Instead of 5 accumulator arrays the lane
// keeps FB = sum_s B_s flow_s and FE = sum_s E_s flow_s and gathers them once per attempt. A rejected
// attempt (0-2 per 365-day episode) re-evaluates the RHS at (t, y) to restore the stage-0 flows.
//
// Two builds of this file:
//   * generic (inside libepi_b200.so): dimensions, shared-memory layout and the per-cell program come
//     from the DevPlan at run time (interpreted tables);
//   * scenario-specialised (EPI_SPEC, epi_spec_main.cu + a generated header, see gen_spec in
//     epi_engine.cu): dimensions and layout are constexpr, the per-cell program is generated
//     straight-line code on registers, all shared-memory addresses are immediates.
//
// Reference call stack and precision map: see epi_kernels.cuh (same arithmetic, same rounding points).
#pragma once
#include "epi_kernels.cuh"
#include "epi_rk.cuh"

namespace epi {
namespace cell {
using namespace rk;

#define EPI_FULL 0xffffffffu

#ifdef EPI_SPEC
// dimensions and layout are compile-time constants; every access indexes the dynamic shared array itself,
// so the compiler knows the address space and folds the offsets into the instruction
#define DIM(x) (spec::x)
#define EPI_LW (spec::LW)
#define SLOT_C1(i) (spec::slot_c1[i])
#define SLOT_C2(i) (spec::slot_c2[i])
#ifdef EPI_HOST_EMU
static thread_local char* cell_smem = nullptr;  // the emulated warp's block (tests: a specialised build on the host)
#else
extern __shared__ __align__(16) char cell_smem[];
#endif
#define COLP(arr) ((T_##arr*)(::epi::cell::cell_smem + spec::l_##arr))
#define ENVP(arr) ((T_##arr*)(::epi::cell::cell_smem + (spec::e_base + spec::e_##arr) + c.ebase))
#else
#define DIM(x) (P.x)
#define EPI_LW (P.cl.lw)
#define SLOT_C1(i) (P.slot_c1[i])
#define SLOT_C2(i) (P.slot_c2[i])
#define COLP(arr) ((T_##arr*)(c.wbase + P.cl.l_##arr))
#define ENVP(arr) ((T_##arr*)(c.wbase + (P.cl.e_base + P.cl.e_##arr) + c.ebase))
#endif
// element i of my lane / of lane ln in a per-lane column; entry i of a table of my env
// (columns are LW = EPW * L*G lanes wide: the idle lanes of a warp own no shared memory)
#define LA(arr, i) (COLP(arr)[(i) * EPI_LW + c.lane])
#define LO(arr, i, ln) (COLP(arr)[(i) * EPI_LW + (ln)])
#define EV(arr, i) (ENVP(arr)[i])
// the stage derivatives: element (stage j, compartment k) of my lane. The 8 stage slots of one compartment are
// two 4-vectors per lane ([k][j/4][lane][j%4]), so the fp32 build reads them with two 128-bit loads
#define KA(j, k) (COLP(K)[((((k) * 2 + ((j) >> 2)) * EPI_LW + c.lane) << 2) + ((j) & 3)])
#define WT(arr, i) (COLP(arr)[i])  // warp-wide table

struct Cx {
  int lane, slot, cell, j, u, lc, l0;  // l0: first lane of my env
  int ebase;                           // byte offset of my env's table block: slot * e_bytes
  bool valid;                          // lane carries a cell of a slot < EPW
  char* wbase;                         // the warp's shared-memory block (generic build)
  unsigned cover;                      // bit sl: move slot sl covers my cell
  float mB[4], mD[4];                  // dense mobility column / row of my cell (specialised builds with L <= 4)
};

__device__ __forceinline__ void bind(const DevPlan& P, char* warp_base, int lane, Cx& c) {
  const int LG = DIM(LG);
  c.wbase = warp_base;
  c.lane = lane;
  c.slot = lane / LG;
  c.cell = lane - c.slot * LG;
  c.valid = c.slot < P.cl.epw;
  if (!c.valid) { c.slot = 0; c.cell = 0; }  // harmless aliases; every store of an invalid lane is predicated off
  c.j = c.cell / DIM(G);
  c.u = c.cell - c.j * DIM(G);
  c.lc = __ldg(P.lclass + c.j);
  c.l0 = c.slot * LG;
#ifdef EPI_SPEC
  c.ebase = c.slot * (int)spec::e_bytes;
  if (spec::dense_mob) {
#pragma unroll
    for (int i = 0; i < (spec::dense_mob ? spec::L : 1); i++) {
      c.mB[i] = __ldg(P.mob_dense_csc + c.cell * spec::L + i);
      c.mD[i] = __ldg(P.mob_dense_csr + c.cell * spec::L + i);
    }
  }
#else
  c.ebase = c.slot * (int)P.cl.e_bytes;
#endif
  c.cover = 0;
  for (int sl = 0; sl < DIM(n_slot); sl++) c.cover |= (__ldg(P.slot_cover + sl * LG + c.cell) ? 1u : 0u) << sl;
}

// sum over the lanes of my env, in lane order; every lane of the env gets the bit-identical result
#define TEAM_SUM(out, v, on_)                                           \
  {                                                                     \
    if (c.valid) LA(red, 0) = (on_) ? (double)(v) : 0.0;                \
    __syncwarp();                                                       \
    double s_ = 0;                                                      \
    if (on_)                                                            \
      for (int i_ = 0; i_ < DIM(LG); i_++) s_ += LO(red, 0, c.l0 + i_); \
    __syncwarp();                                                       \
    out = s_;                                                           \
  }

// all accumulator derivatives of a flow column: generated straight-line code under EPI_SPEC
// column views for the gathers: (pointer to my lane's element 0, stride between elements in `real`s)
#define COL_fl (&LA(fl, 0)), EPI_LW
#define COL_FB (&LA(FBE, 0).x), 2 * EPI_LW
#define COL_FE (&LA(FBE, 0).y), 2 * EPI_LW
#ifdef EPI_SPEC
#define ACC_DECL(name, ...) real name[spec::n_acc > 0 ? spec::n_acc : 1]; acc_all<real>(__VA_ARGS__, name)
#define ACC_AT(name, a, ...) (name[a])
template <class real>
__device__ __forceinline__ void acc_all(const real* col, int stride, real* out) {
  if (stride == spec::LW) spec::acc<spec::LW>(col, out); else spec::acc<2 * spec::LW>(col, out);
}
#else
#define ACC_DECL(name, ...)
#define ACC_AT(name, a, ...) acc_gather<real>(P, __VA_ARGS__, a)
// d acc_a from a flow column (cumulative_move / gain / lost derivative, obj/edge.py:76-111 inverted)
template <class real>
__device__ __forceinline__ real acc_gather(const DevPlan& P, const real* col, int stride, int a) {
  real d = 0;
  for (int q = __ldg(P.ag_off + a); q < __ldg(P.ag_off + a + 1); q++)
    d += col[__ldg(P.ag_src + q) * stride] * (real)__ldg(P.ag_coef + q);
  return d;
}
#endif

// first-same-as-last across the day boundary (see rk45_day: skip0); fp32 throughput mode only
#ifndef EPI_FSAL_DAYS
#define EPI_FSAL_DAYS 1
#endif
// the stage input y of my cell: registers in the specialised build, the ys column otherwise
#ifdef EPI_SPEC
#define Y_DECL real y[spec::C]
#define YV(k) (y[k])
// fp32 specialised build: the (otherwise unused) ys column holds a float copy of the state for the stage inputs,
// y = fma(dy, h, float(X)); the state itself stays double
#define EPI_XF (sizeof(real) == 4)
#define XF(k) LA(ys, k)
#else
#define Y_DECL
#define YV(k) LA(ys, k)
#define EPI_XF 0
#define XF(k) LA(ys, k)
#endif

// ------------------------------------------------------------------------------------------
// my env's time-t tables (obj/parameter.py:80-92), shared by its lanes. When today's multipliers equal
// yesterday's (days 2..7 of a gym step) every gap is zero and the tables are filled once per day.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void fill_tables(const DevPlan& P, const Cx& c, float tf) {
  const int G = DIM(G), F = DIM(F), LG = DIM(LG), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P);
  for (int i = c.cell; i < n_lc * PP * G; i += LG) EV(mpt0, i) = lerp_rn(EV(mpy, i), EV(mpg, i), tf);
  for (int i = c.cell; i < n_lc * F * G * G; i += LG) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G);  // lv = lc*F + v
    float ft = lerp_rn(EV(ft_y, lv * G + u), EV(ft_gap, lv * G + u), tf);
    int a = (lv * G + u0) * G + u, b = (lv * G + u) * G + u0;
    float fa = lerp_rn(EV(fi_y, a), EV(fi_gap, a), tf), fb = lerp_rn(EV(fi_y, b), EV(fi_gap, b), tf);
    EV(Tnum, i) = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (DIM(has_ft_ops)) EV(Tden, i) = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + b);
  }
  const int FG = F * G, nFG = n_lc * FG;
  for (int i = c.cell; i < NG * nFG; i += LG) {
    int k = i / nFG, r = i - k * nFG;  // r = (lc*F + v)*G + u0
    real cf = (real)lerp_rn(EV(ft_y, r), EV(ft_gap, r), tf);
    int pi = __ldg(P.ig_p0i + k);
    cf *= (real)lerp_rn(EV(mpFy, pi * nFG + r), EV(mpFg, pi * nFG + r), tf);
    for (int q = __ldg(P.ig_np_off + k); q < __ldg(P.ig_np_off + k + 1); q++) {
      pi = __ldg(P.ig_npi + q);
      cf *= (real)lerp_rn(EV(mpFy, pi * nFG + r), EV(mpFg, pi * nFG + r), tf);
    }
    for (int q = __ldg(P.ig_dp_off + k); q < __ldg(P.ig_dp_off + k + 1); q++) {
      pi = __ldg(P.ig_dpi + q);
      float m = lerp_rn(EV(mpFy, pi * nFG + r), EV(mpFg, pi * nFG + r), tf);
      if (fabsf(m) > 0) cf /= (real)m;
    }
    EV(coefS, i) = cf;
  }
  if (DIM(mob_dyn)) {
    const int nnz = DIM(nnz);
    for (int i = c.cell; i < G * nnz; i += LG) {
      const int g = i / nnz, k = i - g * nnz;
      const float v = mob_comb(P, &EV(mob_y, 0), &EV(mob_g, 0), g, k, tf);
      EV(mxr, i) = v;
      EV(mxc, g * nnz + __ldg(P.csc_of_csr + k)) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// right-hand side for my cell (core_optimize.py:165-222); ONE call site (rk45_day).
// kslot: which of the 7 stage-derivative slots receives dX. acc_mode: 0 none, 1 FB += bw*flow, FE += ew*flow,
// 2 FB = bw*flow, FE = ew*flow. `dyn`: my env's tables depend on t today. `keep_fl`: the caller reads the
// flow column afterwards (the generic path always writes it).
// Under EPI_SPEC the per-cell program (alive, weighted sums, edges, scatter, flow sums) is straight-line code
// generated from the plan (epi_engine.cu: gen_spec) working on registers; otherwise it is interpreted.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void rhs(const DevPlan& P, const Cx& c, bool on, bool dyn, double t, int ex_bits, int kslot,
                                    real bw, real ew, int acc_mode, bool keep_fl
#ifdef EPI_SPEC
                                    , const real (&y)[spec::C]
#endif
) {
  const int G = DIM(G), F = DIM(F), LG = DIM(LG), nnz = DIM(nnz), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P),
            C = DIM(C), nE = DIM(E);
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83
  real alive = 0;
  if (on) {
    if (dyn) fill_tables<real>(P, c, tf);
    // ---- phase A: alive (N excludes death and hospitalized, :28-41,:168) and weighted infectious sums
#ifdef EPI_SPEC
    real xw[spec::NG > 0 ? spec::NG : 1];
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
#ifdef EPI_SPEC
    if (spec::dense_mob) {
      // dense over the L source locales in ascending order (zeros where the pattern has no entry: exact)
#pragma unroll
      for (int i = 0; i < (spec::dense_mob ? spec::L : 1); i++) {
        int src = c.l0 + i * G + c.u;
        real val = (real)c.mB[i];
        a += LO(cmA, 0, src) * val;
#pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) xs[k] += LO(cmA, 1 + k, src) * val;
      }
    } else
#endif
    for (int q = __ldg(P.csc_indptr + c.j); q < __ldg(P.csc_indptr + c.j + 1); q++) {
      int src = c.l0 + __ldg(P.csc_indices + q) * G + c.u;
      real val = (real)(DIM(mob_dyn) ? EV(mxc, c.u * nnz + q) : __ldg(P.mx_csc + c.u * nnz + q));
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
        for (int u = 0; u < G; u++) den += LO(cmB, 0, row + u) * EV(Tden, tb + u);
      } else {
        for (int u = 0; u < G; u++) den += LO(cmB, 0, row + u) * WT(tdc, tb + u);
      }
      if (den <= (real)1e-15) den = 1;
      const real inv = sizeof(real) == 8 ? (real)0 : frcp(den);  // fp32 mode: one reciprocal for the NG quotients
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) {
          real num = 0;
          for (int u = 0; u < G; u++) num += LO(cmB, 1 + k, row + u) * (real)EV(Tnum, tb + u);
          real foi = sizeof(real) == 8 ? num / den : num * inv;
          sk[k] += EV(coefS, ((k * n_lc + c.lc) * F + v) * G + c.u) * foi;
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
#ifdef EPI_SPEC
      if (spec::dense_mob) {
#pragma unroll
        for (int i = 0; i < (spec::dense_mob ? spec::L : 1); i++) {
          int dst = c.l0 + i * G + c.u;
          real val = (real)c.mD[i];
#pragma unroll
          for (int k = 0; k < EPI_MAX_NG; k++)
            if (k < NG) rs[k] += LO(cmA, 1 + k, dst) * val;
        }
      } else
#endif
      for (int q = __ldg(P.csr_indptr + c.j); q < __ldg(P.csr_indptr + c.j + 1); q++) {
        int dst = c.l0 + __ldg(P.csr_indices + q) * G + c.u;
        real val = (real)(DIM(mob_dyn) ? EV(mxr, c.u * nnz + q) : __ldg(P.mx_csr + c.u * nnz + q));
#pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) rs[k] += LO(cmA, 1 + k, dst) * val;
      }
    }
    const float* mp = &EV(mpt0, c.lc * PP * G + c.u);  // parameters from facility 0 only (:48)
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
      if (((c.cover >> sl) & 1) && ((ex_bits >> sl) & 0x10001)) {
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
        if (spec::has_src64) LA(KS, kslot * spec::n_slot + sl) = k1;
      } else if (spec::has_src64) {
        LA(KS, kslot * spec::n_slot + sl) = (double)d[c1];
      }
      f[nE + NG + sl] = nv;
    }
#pragma unroll
    for (int k = 0; k < C; k++) KA(kslot, k) = d[k];
    if (acc_mode == 1) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) {
        real2 p = LA(FBE, i);
        p.x += bw * f[i];
        p.y += ew * f[i];
        LA(FBE, i) = p;
      }
    } else if (acc_mode == 2) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) {
        real2 p;
        p.x = bw * f[i];
        p.y = ew * f[i];
        LA(FBE, i) = p;
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
      LA(fl, e) = denom > 0 ? fdiv(numer, denom) : numer;
    }
    // ---- phase F: gather flows into dX
    for (int k = 0; k < C; k++) {
      real d = 0;
      for (int q = __ldg(P.xg_off + k); q < __ldg(P.xg_off + k + 1); q++)
        d += LA(fl, __ldg(P.xg_src + q)) * (real)__ldg(P.xg_coef + q);
      KA(kslot, k) = d;
    }
    // ---- phase G: clamped moves, slots in op order. next = X + dX, clamp, dX = next - X in double in BOTH
    // modes: a clamped move must give dX[c1] = -X[c1] to full relative precision (core_optimize.py:196-208)
    for (int sl = 0; sl < P.n_slot; sl++) {
      real nv = 0;
      const int c1 = P.slot_c1[sl], c2 = P.slot_c2[sl];
      if (((c.cover >> sl) & 1) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? LA(mvY, sl) : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? LA(mvD, sl) : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double y1 = P.has_src64 ? LA(ysS, sl) : (double)LA(ys, c1), y2 = (double)LA(ys, c2);
        double n1 = y1 + (double)KA(kslot, c1), n2 = y2 + (double)KA(kslot, c2);
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - y1;
        KA(kslot, c1) = (real)k1;
        KA(kslot, c2) = (real)((n2 + nvd) - y2);
        nv = (real)nvd;
        if (P.has_src64) LA(KS, kslot * P.n_slot + sl) = k1;
      } else if (P.has_src64) {
        LA(KS, kslot * P.n_slot + sl) = (double)KA(kslot, c1);
      }
      LA(fl, nE + NG + sl) = nv;
    }
    // ---- running flow sums for the cumulative components
    if (acc_mode == 1) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = LA(fl, i);
        real2 p = LA(FBE, i);
        p.x += bw * f;
        p.y += ew * f;
        LA(FBE, i) = p;
      }
    } else if (acc_mode == 2) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = LA(fl, i);
        real2 p;
        p.x = bw * f;
        p.y = ew * f;
        LA(FBE, i) = p;
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
__device__ __forceinline__ int prologue(const DevPlan& P, const Cx& c, bool on, int variant, double* Xprev_g,
                                        int ex_y_bits, bool& dyn, bool& tabs_current) {
  const int LG = DIM(LG), G = DIM(G), L = DIM(L), C = DIM(C), F = DIM(F), PP = DIM(P);
  // 1. select sums (executor.py:346-359, ['Value'].sum())
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
  // generated per-cell partial sums of every distinct select, ONE exchange through shared memory (the flow-sum
  // column FBE is free between days and fully rewritten before it is read), each total reduced in lane order by
  // one lane of the env
  {
    double* scr = (double*)COLP(FBE);
    if (on) {
      double part[spec::n_sel > 0 ? spec::n_sel : 1];
      spec::sel_partials(c.j, c.u, &LA(X, 0), P.X0 + c.cell, Xprev_g + c.cell, part);
#pragma unroll
      for (int s = 0; s < spec::n_sel; s++)
        if (spec::sel_rep[s] == s) scr[s * EPI_LW + c.lane] = part[s];
    }
    __syncwarp();
    if (on)
      for (int s = c.cell; s < spec::n_sel; s += LG) {
        const int r = spec::sel_rep[s];
        double tot = 0;
        for (int i = 0; i < LG; i++) tot += scr[r * EPI_LW + c.l0 + i];
        EV(sel, s) = tot;
      }
  }
#else
  for (int s = 0; s < P.n_sel; s++) {
    double part = 0;
    if (on && P.sel_l[s * L + c.j] && P.sel_g[s * G + c.u]) {
      const int mode = __ldg(P.sel_mode + s);
      for (int q = __ldg(P.selc_off + s); q < __ldg(P.selc_off + s + 1); q++) {  // the select's compartments
        const int k = __ldg(P.selc + q);
        if (mode == 0) part += LA(X, k);
        else if (mode == 1) part += P.X0[k * LG + c.cell];
        else part += LA(X, k) - Xprev_g[P.chg_slot_of_comp[k] * LG + c.cell];
      }
    }
    double tot;
    TEAM_SUM(tot, part, on);
    if (on && c.cell == 0) EV(sel, s) = tot;
  }
#endif
  __syncwarp();
  if (on) {
    // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
#pragma unroll
    for (int k = 0; k < C; k++)
      if (spec::chg_slot[k] >= 0) Xprev_g[spec::chg_slot[k] * LG + c.cell] = LA(X, k);
#else
    for (int k = 0; k < C; k++) {
      int sl = P.chg_slot_of_comp[k];
      if (sl >= 0) Xprev_g[sl * LG + c.cell] = LA(X, k);
    }
#endif
    // 2. expressions
    for (int e = c.cell; e < P.n_expr; e += LG) {
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
      EV(expr, e) = EV(elive, e) ? spec::expr_eval(e, &EV(cpv, 0), &EV(sel, 0)) : 0.0;
#else
      EV(expr, e) = EV(elive, e) ? eval_expr(P, e, &EV(cpv, 0), &EV(sel, 0)) : 0.0;
#endif
    }
  }
  __syncwarp();
  if (on) {
    // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
    for (int k = c.cell; k < P.n_cls; k += LG) {
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
      EV(mult_d, k) = spec::cls_mult(k, &EV(live, 0), &EV(expr, 0));
#else
      float m = 1.0f;
      for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
        int o = P.cls_op[q];
        if (EV(live, o)) m = __fmul_rn(m, (float)EV(expr, P.op_expr[o]));
      }
      EV(mult_d, k) = m;
#endif
    }
  }
  __syncwarp();
  {
    // do today's multipliers differ from yesterday's? (if not, every parameter gap is zero)
    double diff = 0, tot;
    if (on)
      for (int k = c.cell; k < P.n_cls; k += LG) diff += EV(mult_d, k) != EV(mult_y, k) ? 1.0 : 0.0;
    TEAM_SUM(tot, diff, on);
    dyn = tot > 0;
  }
  // the (y, gap) parameter tables and the time-t tables depend on the multipliers only: when yesterday's and
  // today's are equal for the second day in a row they are already right (y = today's value, gap = 0)
  const bool rebuild = dyn || !tabs_current;
  int ex_d = 0;
  if (on && rebuild) {
    // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
    for (int row = c.cell; row < P.n_lc * F * G; row += LG) {
      float a[32], b[32];  // G <= 32 (checked on the host)
      facility_row_fi(P, &EV(mult_y, 0), row, a);
      facility_row_fi(P, &EV(mult_d, 0), row, b);
      for (int g2 = 0; g2 < G; g2++) {
        EV(fi_y, row * G + g2) = a[g2];
        EV(fi_gap, row * G + g2) = __fsub_rn(b[g2], a[g2]);
      }
    }
    for (int it = c.cell; it < P.n_lc * G; it += LG) {
      int lc = it / G, g = it % G;
      float a[32], b[32];  // F <= 32 (checked on the host)
      // facility_timespent normalised over F (get_normalized_facility_timespent, dense.py:6-19)
      float s = 0.0f, s2 = 0.0f;
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        a[f] = __fmul_rn(P.ft0[i], EV(mult_y, P.ft_cls[i]));
        b[f] = __fmul_rn(P.ft0[i], EV(mult_d, P.ft_cls[i]));
        s = __fadd_rn(s, a[f]);
        s2 = __fadd_rn(s2, b[f]);
      }
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        float ya = s > 1e-15f ? __fdiv_rn(a[f], s) : (f == 0 ? 1.0f : 0.0f);
        float yb = s2 > 1e-15f ? __fdiv_rn(b[f], s2) : (f == 0 ? 1.0f : 0.0f);
        EV(ft_y, i) = ya;
        EV(ft_gap, i) = __fsub_rn(yb, ya);
      }
    }
    // model parameters: default * class multiplier of yesterday / today (parameter.py:62,:72)
    for (int i = c.cell; i < P.n_lc * PP * G; i += LG) {
      int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
      int idx = ((lc * PP + p) * F + 0) * G + g;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, EV(mult_y, cls)), d = __fmul_rn(m0, EV(mult_d, cls));
      EV(mpy, i) = y;
      EV(mpg, i) = __fsub_rn(d, y);
    }
    const int nFG = P.n_lc * F * G;
    for (int i = c.cell; i < P.n_igp * nFG; i += LG) {
      int pi = i / nFG, r = i - pi * nFG, lc = r / (F * G), fg = r - lc * F * G;
      int idx = (lc * PP + __ldg(P.igp + pi)) * F * G + fg;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, EV(mult_y, cls)), d = __fmul_rn(m0, EV(mult_d, cls));
      EV(mpFy, i) = y;
      EV(mpFg, i) = __fsub_rn(d, y);
    }
    if (DIM(mob_dyn))  // the mode matrices of yesterday and today -> (y, gap)  (matrix/coo.py:173-188)
      for (int it = c.cell; it < P.M * G * L; it += LG)
        mob_row(P, &EV(mult_y, 0), &EV(mult_d, 0), it / (G * L), (it / L) % G, it % L, &EV(mob_y, 0), &EV(mob_g, 0));
  }
  if (on) {
    // 5. move table of today (nb_move, executor.py:166-234; different compartment, same locale)
    for (int sl = 0; sl < DIM(n_slot); sl++) LA(mvD, sl) = 0.0f;
  }
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
  // the same loop nest over compile-time tables: fully unrolled, no table loads
#pragma unroll
  for (int mo = 0; mo < spec::n_mv_op; mo++) {
    const bool live = on && EV(live, spec::mv_op[mo]);
    const bool mine = live && ((spec::mv_gmask[mo] >> c.u) & 1) && ((spec::mv_lmask[mo] >> c.j) & 1);
    double part = 0;
    if (mine) {
#pragma unroll
      for (int q = 0; q < spec::mv_max_c1; q++)
        if (q < spec::mv_nc1[mo]) part += LA(X, spec::mv_c1[spec::mv_c1_off[mo] + q]);
    }
    double depart_sum;
    TEAM_SUM(depart_sum, part, on);
    if (live && depart_sum > 0) {
      const double value = EV(expr, spec::mv_expr[mo]);
      if (mine) {
#pragma unroll
        for (int q1 = 0; q1 < spec::mv_max_c1; q1++)
          if (q1 < spec::mv_nc1[mo]) {
            const int c1 = spec::mv_c1[spec::mv_c1_off[mo] + q1];
            const double depart = LA(X, c1) / depart_sum * value;
            double arrive = 0;
#pragma unroll
            for (int q = 0; q < spec::mv_max_c2; q++)
              if (q < spec::mv_nc2[mo]) arrive += LA(X, spec::mv_c2[spec::mv_c2_off[mo] + q]);
#pragma unroll
            for (int q = 0; q < spec::mv_max_c2; q++)
              if (q < spec::mv_nc2[mo]) {
                const int c2 = spec::mv_c2[spec::mv_c2_off[mo] + q];
                const double amt = arrive > 0 ? LA(X, c2) / arrive * depart : 1.0 / spec::mv_nc2[mo] * depart;
                float* cellp = &LA(mvD, spec::slot_of_pair[c1 * spec::C + c2]);
                *cellp = (float)((double)*cellp + amt);  // CooMatrix.element_add: f32 cell (coo.py:52-57)
              }
          }
      }
#pragma unroll
      for (int q1 = 0; q1 < spec::mv_max_c1; q1++)
#pragma unroll
        for (int q = 0; q < spec::mv_max_c2; q++)
          if (q1 < spec::mv_nc1[mo] && q < spec::mv_nc2[mo])
            ex_d |= 1 << spec::slot_of_pair[spec::mv_c1[spec::mv_c1_off[mo] + q1] * spec::C + spec::mv_c2[spec::mv_c2_off[mo] + q]];
    }
  }
#else
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    bool live = on && EV(live, o);
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    bool mine = live && P.mv_g[mo * G + c.u] && P.mv_l1[mo * L + c.j];
    double part = 0;
    if (mine)
      for (int q = 0; q < nc1; q++) part += LA(X, c1s[q]);
    double depart_sum;
    TEAM_SUM(depart_sum, part, on);
    if (live && depart_sum > 0) {
      double value = EV(expr, P.op_expr[o]);
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
#endif
  if (on) {
    // 6. cost rates of today (nb_add, executor.py:236-243)
#if defined(EPI_SPEC) && EPI_SPEC_PROLOGUE
    for (int it = c.cell; it < DIM(I) * L; it += LG) EV(crD, it) = spec::cost_rate(it, &EV(live, 0), &EV(expr, 0));
#else
    for (int it = c.cell; it < P.I * L; it += LG) {
      double cr = 0;
      for (int q = __ldg(P.cr_off + it); q < __ldg(P.cr_off + it + 1); q++) {  // the ADD ops covering (i, l), in op order
        const int ao = __ldg(P.cr_ops + q), o = __ldg(P.add_op + ao);
        if (EV(live, o)) cr += EV(expr, __ldg(P.op_expr + o)) / (double)__ldg(P.add_parts + ao);
      }
      EV(crD, it) = cr;
    }
#endif
  }
  __syncwarp();
  if (on && !dyn && rebuild) fill_tables<real>(P, c, 0.0f);  // constant over the day: filled once
  tabs_current = !dyn;
  __syncwarp();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: scipy RK45 (rk.py / common.py restated) for the envs of this warp, as a warp-uniform
// state machine around a single right-hand-side call.
//   phase 0     f(0, y0)                        (nfev 1)        select_initial_step, common.py:68-134
//   phase 1     f(h0, y0 + h0 f0)               (nfev 2)
//   phase 2..6  stages 1..5 of an attempt                       rk_step, rk.py:14-71
//   phase 7     f(t + h, y_new), error norm, accept / reject    rk.py:111-176
//   phase 8     re-evaluation of f(t, y) after a rejection (restores the stage-0 flow sums; not in nfev)
// The error norm is taken over ALL n_flat components (common.py:63-65, rk.py:146-148); the compartment and
// accumulator terms are formed in `real`: the controller only sees err^(-1/5), far from the accept
// threshold.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ bool rk45_day(const DevPlan& P, const Cx& c, bool on, bool dyn, int ex_bits, int* trace,
                                         double* last_err_out, double* dbg, bool skip0) {
  const double rtol = 1e-3, atol = 1e-6;
  const real rtolr = (real)1e-3, atolr = (real)1e-6;
  const int C = DIM(C), LG = DIM(LG), nC = DIM(I) * DIM(L), nS = DIM(has_src64) ? DIM(n_slot) : 0, NSRC = DIM(NSRC),
            n_acc = DIM(n_acc);
  const double inv_n = 1.0 / (double)P.n_flat;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0, n_att = 0, guard = 0;
  double last_err = 0, t = 0.0, h_abs = 0, h = 0, t_new = 0, h0 = 0, d0 = 0, d1 = 0;
  bool run = on, rejected = false, redo = false;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };
  Y_DECL;

  int phase = 0;
  while (phase >= 0) {
    // ---- who evaluates, where, and what is done with the flows
    const int stage = PH_STAGE[phase];  // K slot written
    if (phase == 2 && run) {
      // a new attempt (rk.py:111-128)
      double min_step = 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t);  // nextafter(t, inf), t >= 0
      if (h_abs < min_step) {
        if (rejected) { status = -1; run = false; }  // step size too small after a rejection
        else h_abs = min_step;
      }
      if (++guard > 10000 || !(h_abs == h_abs)) { status = -1; run = false; }
      if (run) {
        h = h_abs;
        t_new = t + h;
        if (t_new - 1.0 > 0) t_new = 1.0;
        h = t_new - t;
        h_abs = fabs(h);
      }
    }
    // skip0 (fp32 mode, days 2.. of a fused launch whose parameters do not move today): f(0, y0) is yesterday's
    // last stage f(1, y_new) - same state, same parameters up to float rounding - and its derivatives (K slot 0),
    // move-source derivative and flows are still in place (FSAL across the day boundary); nfev counts it
    const bool ev = phase == 0 ? (on && !skip0) : (phase == 1 ? on : (phase == 8 ? redo : run));
    // per-phase scalars from constant tables (epi_rk.cuh): no branch chains on the single warp's critical path
    const int nj = PH_NJ[phase];
    const double hh = phase <= 1 ? h0 : h;
    const double t_eval = t + PH_TC[phase] * hh;
    const real bw = phBW(real(0), phase), ew = phEW(real(0), phase);
    // ---- stage input: y = X + hh * sum_{j<nj} coef_j K_j   (coef = 1 | A[stage][:] | B[:])
    if (ev) {
      real co[6];
#pragma unroll
      for (int j = 0; j < 6; j++) co[j] = rkCO(real(0), phase, j);  // zero beyond the stages that exist
#pragma unroll
      for (int k = 0; k < C; k++) {
        real dy = 0;
#ifndef EPI_HOST_EMU
        if (sizeof(real) == 4) {
          // co[j] is zero beyond the stages that exist and K is finite (zeroed at launch): no predicates
          const float4 lo = *(const float4*)&KA(0, k), hi = *(const float4*)&KA(4, k);
          dy += (real)lo.x * co[0]; dy += (real)lo.y * co[1]; dy += (real)lo.z * co[2]; dy += (real)lo.w * co[3];
          dy += (real)hi.x * co[4]; dy += (real)hi.y * co[5];
        } else
#endif
        {
#pragma unroll
          for (int j = 0; j < 6; j++)
            if (j < nj) dy += KA(j, k) * co[j];
        }
        if (EPI_XF) YV(k) = dy * (real)hh + XF(k);
        else YV(k) = (real)(LA(X, k) + (double)dy * hh);
      }
      for (int q = 0; q < nS; q++) {
        // (fp32 mode only) the double-carried move sources: coefficients are zero beyond the stages that exist
        // and the slots are finite (zeroed at launch), so no predicates; two partial sums halve the DFMA chain
        double dy0 = 0, dy1 = 0;
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
          dy0 += LA(KS, j * nS + q) * RK_CO64[phase][j];
          dy1 += LA(KS, (j + 1) * nS + q) * RK_CO64[phase][j + 1];
        }
        LA(ysS, q) = LA(X, SLOT_C1(q)) + (dy0 + dy1) * hh;
      }
    }
    rhs<real>(P, c, ev, dyn, t_eval, ex_bits, stage, bw, ew, PH_ACC[phase], PH_KEEP[phase] != 0
#ifdef EPI_SPEC
              , y
#endif
    );
    // ---- what follows the evaluation
    if (phase == 0) {
      // select_initial_step, first half: d0 = |y|, d1 = |f0| in the scaled RMS norm
      real a0 = 0, a1 = 0;
      double s0 = 0, s1 = 0;
      if (on) {
#pragma unroll
        for (int k = 0; k < C; k++) {
          real yv = (real)LA(X, k), sc = atolr + fabs(yv) * rtolr;
          real a = fdiv(yv, sc), b = fdiv(KA(0, k), sc);
          a0 += a * a; a1 += b * b;
        }
        {
          ACC_DECL(g0, COL_fl);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real yv = LA(accY, a_), sc = atolr + fabs(yv) * rtolr;
            real a = fdiv(yv, sc), b = fdiv(ACC_AT(g0, a_, COL_fl), sc);
            a0 += a * a; a1 += b * b;
          }
        }
        s0 = (double)a0; s1 = (double)a1;
        for (int i = c.cell; i < nC; i += LG) {
          double yv = EV(cost, i), sc = atol + fabs(yv) * rtol;
          double a = norm_term(real(0), yv, sc), b = norm_term(real(0), EV(crY, i), sc);  // q(0) = 0
          s0 += a * a; s1 += b * b;
        }
        for (int i = 0; i < NSRC; i++) LA(FBE, i).y = LA(fl, i);  // stage-0 flows, needed once more in phase 1
      }
      TEAM_SUM(d0, s0, on);
      TEAM_SUM(d1, s1, on);
      d0 = ctrl_sqrt(real(0), d0 * inv_n);
      d1 = ctrl_sqrt(real(0), d1 * inv_n);
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, 1.0);
      nfev++;
      phase = 1;
    } else if (phase == 1) {
      real a2 = 0;
      double s2 = 0, d2, q0 = qd(h0);
      if (on) {
#pragma unroll
        for (int k = 0; k < C; k++) {
          real sc = atolr + fabs((real)LA(X, k)) * rtolr, a = fdiv(KA(1, k) - KA(0, k), sc);
          a2 += a * a;
        }
        {
          ACC_DECL(g1, COL_fl);
          ACC_DECL(g0, COL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real sc = atolr + fabs(LA(accY, a_)) * rtolr;
            real a = fdiv(ACC_AT(g1, a_, COL_fl) - ACC_AT(g0, a_, COL_FE), sc);
            a2 += a * a;
          }
        }
        s2 = (double)a2;
        for (int i = c.cell; i < nC; i += LG) {
          double sc = atol + fabs(EV(cost, i)) * rtol, a = norm_term(real(0), (EV(crD, i) - EV(crY, i)) * q0, sc);
          s2 += a * a;
        }
        // stage-0 contributions of the first attempt
        for (int i = 0; i < NSRC; i++) {
          real f = LA(FBE, i).y;
          real2 p;
          p.x = rkB(real(0), 0) * f;
          p.y = rkE(real(0), 0) * f;
          LA(FBE, i) = p;
        }
      }
      TEAM_SUM(d2, s2, on);
      d2 = ctrl_sqrt(real(0), d2 * inv_n) / h0;
      double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : ctrl_pow(real(0), 0.01 / fmax(d1, d2), 0.2);
      h_abs = fmin(fmin(100 * h0, h1), 1.0);
      nfev++;
      phase = __any_sync(EPI_FULL, run) ? 2 : -1;
    } else if (phase < 7) {
      phase++;
    } else if (phase == 7) {
      // error norm, accept / reject
      real ae = 0;
      double e2 = 0, qs[7];
      if (run) {
        nfev += 6;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real sc = atolr + fmax(fabs((real)LA(X, k)), fabs(YV(k))) * rtolr;
          real e = 0;
#ifndef EPI_HOST_EMU
          if (sizeof(real) == 4) {
            const float4 lo = *(const float4*)&KA(0, k), hi = *(const float4*)&KA(4, k);
            e += (real)lo.x * rkE(real(0), 0); e += (real)lo.y * rkE(real(0), 1); e += (real)lo.z * rkE(real(0), 2);
            e += (real)lo.w * rkE(real(0), 3); e += (real)hi.x * rkE(real(0), 4); e += (real)hi.y * rkE(real(0), 5);
            e += (real)hi.z * rkE(real(0), 6);
          } else
#endif
          {
#pragma unroll
            for (int j = 0; j < 7; j++) e += KA(j, k) * rkE(real(0), j);
          }
          real a = fdiv(e * (real)h, sc);
          ae += a * a;
        }
        {
          ACC_DECL(gB, COL_FB);
          ACC_DECL(gE, COL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real yv = LA(accY, a_), yn = yv + (real)h * ACC_AT(gB, a_, COL_FB);
            real sc = atolr + fmax(fabs(yv), fabs(yn)) * rtolr, a = fdiv(ACC_AT(gE, a_, COL_FE) * (real)h, sc);
            ae += a * a;
          }
        }
        e2 = (double)ae;
#pragma unroll
        for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
        for (int i = c.cell; i < nC; i += LG) {
          double cy = EV(crY, i), cg = EV(crD, i) - cy, sb = 0, se = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) {
            double k = cy + cg * qs[j];
            sb += k * RK_B[j];
            se += k * RK_E[j];
          }
          double yv = EV(cost, i), yn = yv + h * sb;
          double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = norm_term(real(0), se * h, sc);
          e2 += a * a;
        }
      }
      double err;
      TEAM_SUM(err, e2, run);
      err = ctrl_sqrt(real(0), err * inv_n);
      bool accepted = false;
      if (run) {
        last_err = err;
        if (dbg && c.cell == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
        n_att++;
        if (err < 1) {
          double factor = err == 0 ? 10.0 : fmin(10.0, 0.9 * ctrl_pow(real(0), err, -0.2));
          if (rejected) factor = fmin(1.0, factor);
          h_abs *= factor;
          accepted = true;
        } else {
          h_abs *= fmax(0.2, 0.9 * ctrl_pow(real(0), err, -0.2));
          rejected = true;
          nrej++;
        }
      }
      redo = run && !accepted;
      if (run && accepted) {
        // commit: y = y_new, f = f_new. The state itself is advanced in double so that float rounding of X
        // does not pile up over hundreds of days; in the fp64 mode this is the same arithmetic as y_new above
        const real hr = (real)h;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real dy = 0;
#ifndef EPI_HOST_EMU
          if (sizeof(real) == 4) {
            const float4 lo = *(const float4*)&KA(0, k), hi = *(const float4*)&KA(4, k);
            dy += (real)lo.x * rkB(real(0), 0); dy += (real)lo.y * rkB(real(0), 1); dy += (real)lo.z * rkB(real(0), 2);
            dy += (real)lo.w * rkB(real(0), 3); dy += (real)hi.x * rkB(real(0), 4); dy += (real)hi.y * rkB(real(0), 5);
            KA(0, k) = (real)hi.z;
          } else
#endif
          {
#pragma unroll
            for (int j = 0; j < 6; j++) dy += KA(j, k) * rkB(real(0), j);
            KA(0, k) = KA(6, k);
          }
          {
            const double xn = LA(X, k) + h * (double)dy;
            LA(X, k) = xn;
            if (EPI_XF) XF(k) = (real)xn;
          }
        }
        for (int q = 0; q < nS; q++) {
          LA(X, SLOT_C1(q)) = LA(ysS, q);  // = X + h * sum_j B_j KS_j in double
          if (EPI_XF) XF(SLOT_C1(q)) = (real)LA(ysS, q);
          LA(KS, q) = LA(KS, 6 * nS + q);
        }
        {
          ACC_DECL(gB, COL_FB);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) LA(accY, a_) += hr * ACC_AT(gB, a_, COL_FB);
        }
        for (int i = 0; i < NSRC; i++) {  // the last stage's flows open the next attempt (FSAL)
          real f = LA(fl, i);
          real2 p;
          p.x = rkB(real(0), 0) * f;
          p.y = rkE(real(0), 0) * f;
          LA(FBE, i) = p;
        }
        for (int i = c.cell; i < nC; i += LG) {
          double cy = EV(crY, i), cg = EV(crD, i) - cy, sb = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
          EV(cost, i) += h * sb;
        }
        nsteps++;
        t = t_new;
        rejected = false;
        if (!(t < 1.0)) run = false;
      }
      phase = __any_sync(EPI_FULL, redo) ? 8 : (__any_sync(EPI_FULL, run) ? 2 : -1);
    } else {
      redo = false;
      phase = __any_sync(EPI_FULL, run) ? 2 : -1;
    }
  }
  if (on && c.cell == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
  return on && status == 0;
}

// ------------------------------------------------------------------------------------------
// the launch: EpiEnv.step for a batch (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void step_warp(const DevPlan& P, const DevState& S, const StepArgs& a, char* warp_base,
                                          int lane, int group) {
#if defined(EPI_SPEC) && defined(EPI_HOST_EMU)
  cell_smem = warp_base;
#endif
  Cx c;
  bind(P, warp_base, lane, c);
  const int C = DIM(C), CLG = DIM(CLG), LG = DIM(LG), nA = DIM(n_acc) * LG, nC = DIM(I) * DIM(L), nM = DIM(n_slot) * LG;
  const int env = group * P.cl.epw + c.slot;
  const bool on0 = c.valid && env < a.N;
  const size_t e = on0 ? (size_t)env : 0;

  // ---- the constant FOI-denominator kernel (no FT_MUL ops) as a warp-wide table
  if (!DIM(has_ft_ops) && c.lane < EPI_LW)
    for (int i = c.lane; i < DIM(n_lc) * DIM(F) * DIM(G) * DIM(G); i += EPI_LW)
      WT(tdc, i) = sizeof(real) == 8 ? (real)__ldg(P.Tden_const64 + i) : (real)__ldg(P.Tden_const + i);
  // the stage-derivative column starts finite (its unused slots are multiplied by zero coefficients)
  if (c.lane < EPI_LW) {
    for (int i = c.lane; i < 8 * C * EPI_LW; i += EPI_LW) COLP(K)[i] = 0;
    if (DIM(has_src64))
      for (int i = c.lane; i < 7 * DIM(n_slot) * EPI_LW; i += EPI_LW) COLP(KS)[i] = 0;
  }
  // ---- load the env's persistent state
  double* gX = S.X + e * CLG;
  real* gAcc = (real*)S.acc + e * nA + c.cell;
  double* gXprev = S.Xprev + e * P.n_chg * LG;
  int* meta = S.meta + e * 8;
  if (on0) {
    for (int k = 0; k < C; k++) {
      const double x = gX[k * LG + c.cell];
      LA(X, k) = x;
      if (EPI_XF) XF(k) = (real)x;
    }
    for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = S.mvY[e * nM + s * LG + c.cell];
    for (int a_ = 0; a_ < DIM(n_acc); a_++) LA(accY, a_) = gAcc[a_ * LG];
    for (int i = c.cell; i < nC; i += LG) {
      EV(cost, i) = S.cost[e * nC + i];
      EV(crY, i) = S.crY[e * nC + i];
    }
    for (int i = c.cell; i < P.n_cls; i += LG) EV(mult_y, i) = S.multY[e * P.n_cls + i];
    // control parameters (epipolicy_environment.py:40-54) and active flags
    for (int k = c.cell; k < P.NCP; k += LG) {
      float v;
      if (a.cpv) v = a.cpv[e * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[e * P.A + s] : P.cp_fixed[k]; }
      EV(cpv, k) = v;
    }
    for (int i = c.cell; i < P.I; i += LG)
      EV(act, i) = a.active ? a.active[e * P.I + i] : (a.init_only ? P.init_active[i] : 1);
  }
  int t_day = on0 ? meta[0] : 0, ex_y = on0 ? meta[1] : 0;
  __syncwarp();
  if (on0) {
    // which ops / expressions are live for this launch: the intervention is active and the op belongs to the
    // program variant in use (fixed over the launch's days)
    for (int o = c.cell; o < P.n_op; o += LG) EV(live, o) = (uint8_t)op_live(P, &EV(act, 0), a.variant, o);
    for (int x = c.cell; x < P.n_expr; x += LG) {
      int itv = P.expr_itv[x];
      EV(elive, x) = (uint8_t)(EV(act, itv) && P.expr_prog[x] == (a.variant ? P.prog_init[itv] : P.prog_step[itv]));
    }
  }
  __syncwarp();

  double reward = 0;
  bool on = on0, tabs_current = false, fsal_ok = false;
  const int n_days = a.init_only ? 1 : a.n_days;
  for (int day = 0; day < n_days; day++) {
    if (!__any_sync(EPI_FULL, on)) break;
    if (day > 0 && a.cpv_days) {  // schedule mode: today's row of control parameters / Act flags, and what is live with them
      if (on0) {
        const size_t row = (size_t)day * a.N + e;
        for (int k = c.cell; k < P.NCP; k += LG) EV(cpv, k) = a.cpv_days[row * P.NCP + k];
        for (int i = c.cell; i < P.I; i += LG) EV(act, i) = a.active_days ? a.active_days[row * P.I + i] : 1;
      }
      __syncwarp();
      if (on0) {
        for (int o = c.cell; o < P.n_op; o += LG) EV(live, o) = (uint8_t)op_live(P, &EV(act, 0), a.variant, o);
        for (int x = c.cell; x < P.n_expr; x += LG) {
          int itv = P.expr_itv[x];
          EV(elive, x) = (uint8_t)(EV(act, itv) && P.expr_prog[x] == (a.variant ? P.prog_init[itv] : P.prog_step[itv]));
        }
      }
      __syncwarp();
    }
    bool dyn = true;
    int ex_bits = prologue<real>(P, c, on, a.variant, gXprev, ex_y, dyn, tabs_current);
    if (!a.init_only) {
      double c0 = 0, c1 = 0;
      if (on)
        for (int i = c.cell; i < nC; i += LG) c0 += EV(cost, i);
      const bool skip0 = EPI_FSAL_DAYS && sizeof(real) == 4 && fsal_ok && !dyn;
      fsal_ok = rk45_day<real>(P, c, on, dyn, ex_bits, meta, S.last_err + e,
                               a.dbg_attempts ? a.dbg_attempts + e * 128 : nullptr, skip0);
      if (on)
        for (int i = c.cell; i < nC; i += LG) c1 += EV(cost, i);
      // reward = -sum(cumulative_cost_next - cumulative_cost_current)  (epidemic.py:115)
      double dr;
      TEAM_SUM(dr, c1 - c0, on);
      if (on) {
        reward -= dr; t_day++;
        if (a.reward_days && c.cell == 0) a.reward_days[(size_t)day * a.N + env] = -dr;
        const size_t row = (size_t)day * a.N + e;
        if (a.obs_days)
          for (int k = 0; k < C; k++) ((real*)a.obs_days)[row * CLG + k * LG + c.cell] = (real)LA(X, k);
        if (a.x_days)
          for (int k = 0; k < C; k++) a.x_days[row * CLG + k * LG + c.cell] = LA(X, k);
        if (a.acc_days)
          for (int a_ = 0; a_ < DIM(n_acc); a_++) ((real*)a.acc_days)[row * nA + a_ * LG + c.cell] = LA(accY, a_);
        if (a.mult_days)
          for (int i = c.cell; i < P.n_cls; i += LG) a.mult_days[row * P.n_cls + i] = EV(mult_d, i);
      }
    }
    // today's parameters become yesterday's (State(t+1, next_obs, dest_parameter), epidemic.py:89)
    if (on) {
      for (int i = c.cell; i < P.n_cls; i += LG) EV(mult_y, i) = EV(mult_d, i);
      for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = LA(mvD, s);
      for (int i = c.cell; i < nC; i += LG) EV(crY, i) = EV(crD, i);
      ex_y = (ex_bits >> 16) & 0xffff;
    }
    __syncwarp();
    if (on && t_day >= P.horizon) on = false;  // the reference's EpiEnv.step stops its day loop at done
  }
  // ---- store
  if (a.obs_out) {
    // the observations of the warp's envs are one contiguous [EPW][C*L*G] block: lanes write consecutive words
    // (full lines - this matters when obs_out is mapped host memory, epi_step_host) from the [k][lane] state column
#ifdef EPI_HOST_EMU
    const int nthr = EPI_LW;
#else
    const int nthr = 32;
#endif
    const int e0 = group * P.cl.epw, nv = a.N - e0 < P.cl.epw ? a.N - e0 : P.cl.epw;
    real* out = (real*)a.obs_out + (size_t)e0 * CLG;
    real* out2 = a.obs_out2 ? (real*)a.obs_out2 + (size_t)e0 * CLG : nullptr;
    for (int idx = lane; idx < nv * CLG; idx += nthr) {
      const int sl = idx / CLG, rem = idx - sl * CLG, k = rem / LG, cl = rem - k * LG;
      const real v = (real)LO(X, k, sl * LG + cl);
      out[idx] = v;
      if (out2) out2[idx] = v;
    }
  }
  if (on0) {
    for (int k = 0; k < C; k++) gX[k * LG + c.cell] = LA(X, k);
    for (int s = 0; s < DIM(n_slot); s++) S.mvY[e * nM + s * LG + c.cell] = LA(mvY, s);
    for (int a_ = 0; a_ < DIM(n_acc); a_++) gAcc[a_ * LG] = LA(accY, a_);
    for (int i = c.cell; i < nC; i += LG) {
      S.cost[e * nC + i] = EV(cost, i);
      S.crY[e * nC + i] = EV(crY, i);
    }
    for (int i = c.cell; i < P.n_cls; i += LG) S.multY[e * P.n_cls + i] = EV(mult_y, i);
    if (c.cell == 0) {
      meta[0] = t_day; meta[1] = ex_y;
      if (a.reward_out) a.reward_out[env] = reward;
      if (a.done_out) a.done_out[env] = (uint8_t)(t_day >= P.horizon);
      if (a.reward_out2) a.reward_out2[env] = reward;
      if (a.done_out2) a.done_out2[env] = (uint8_t)(t_day >= P.horizon);
    }
  }
}

#ifndef EPI_HOST_EMU
// one warp per CTA (the scenario-specialised build relies on it: its shared-memory offsets are absolute)
template <class real>
__global__ void __launch_bounds__(32) cell_kernel(const DevPlan P, const DevState S, const StepArgs a) {
#ifdef EPI_SPEC
  step_warp<real>(P, S, a, cell_smem, threadIdx.x, blockIdx.x);
#else
  extern __shared__ __align__(16) char smem[];
  step_warp<real>(P, S, a, smem, threadIdx.x, blockIdx.x);
#endif
}
#endif

}  // namespace cell
}  // namespace epi
