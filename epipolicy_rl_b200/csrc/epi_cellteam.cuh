// epi_cellteam.cuh — the cell kernel with R cooperating warps per env group ("roles") (sm_100a).
//
// The cell kernel (epi_cell.cuh) is bound by shared memory: ~12.5 KB of RK45 working set per COVID-sized env
// leaves 5 warps per SM, and with one warp per group of envs the kernel is a handful of latency-bound
// instruction streams per scheduler. Here a CTA of R warps shares ONE group's working set: lane l of every
// warp is the same (env, locale, group) cell, and the per-cell program is split by ROLE (= warp index):
//   * regular edges e, flow-sum entries i, accumulators a:           role = index mod R
//   * compartments k (stage algebra, scatter into dX, error terms):  role = comp_role[k] (k mod R, except that
//     the two compartments of a move slot sit together on role 0 so the clamp needs no exchange)
//   * the force-of-infection quantities (alive, weighted infectious sums): role = quantity mod R
// Roles exchange through the shared-memory columns that already exist (ys, fl, K, cmA/cmB) with a CTA barrier
// between phases (6 per right-hand-side evaluation); role-dependent code is selected by warp-uniform
// branches, so there is no divergence. Same shared memory per CTA, R times the resident warps, 1/R of the
// serial work per warp. R = 1 reproduces the arithmetic of the cell kernel.
//
// Arithmetic, rounding points and the RK45 state machine are those of epi_cell.cuh (see there for the
// reference citations); per-role partial sums of the error norm are added in (role, cell) order.
#pragma once
#include "epi_cell.cuh"

namespace epi {
namespace cteam {
using namespace rk;
using cell::Cx;

#ifdef EPI_HOST_EMU
#define EPI_TWARP (P.cl.lw)  // the emulation runs LW host threads per warp
#else
#define EPI_TWARP 32
#endif
#ifdef EPI_SPEC
#define NROLE (spec::R)
#define COMP_ROLE(k) (spec::comp_role[k])
#else
#define NROLE (P.cl.roles)
#define COMP_ROLE(k) (P.comp_role[k])
#endif
#define TSYNC() __syncthreads()

struct Tx : Cx {
  int role;    // warp index in the CTA
  int tl, tn;  // my index among / the number of the env's (role, cell) lanes: strides for per-env table work
};

// sum over the (role, cell) lanes of my env in a fixed order; every lane gets the bit-identical result
#define TSUM(out, v, on_)                                                                    \
  {                                                                                          \
    if (c.valid) LO(red, 0, c.role * EPI_LW + c.lane) = (on_) ? (double)(v) : 0.0;           \
    TSYNC();                                                                                 \
    double s_ = 0;                                                                           \
    if (on_)                                                                                 \
      for (int r_ = 0; r_ < NROLE; r_++)                                                     \
        for (int i_ = 0; i_ < DIM(LG); i_++) s_ += LO(red, 0, r_ * EPI_LW + c.l0 + i_);      \
    TSYNC();                                                                                 \
    out = s_;                                                                                \
  }

#ifndef EPI_SPEC
template <class real>
__device__ __forceinline__ real acc_gather(const DevPlan& P, const real* col, int stride, int a) {
  real d = 0;
  for (int q = __ldg(P.ag_off + a); q < __ldg(P.ag_off + a + 1); q++)
    d += col[__ldg(P.ag_src + q) * stride] * (real)__ldg(P.ag_coef + q);
  return d;
}
#endif
// accumulator derivatives of my role (a = role, role + R, ...) from a flow column
#ifdef EPI_SPEC
#define TACC_DECL(name, ...) real name[(spec::n_acc + spec::R - 1) / spec::R + 1]; tacc_all<real>(c.role, __VA_ARGS__, name)
#define TACC_AT(name, a, ...) (name[(a) / spec::R])
template <class real>
__device__ __forceinline__ void tacc_all(int role, const real* col, int stride, real* out) {
  if (stride == spec::LW) spec::acc_role<spec::LW>(role, col, out); else spec::acc_role<2 * spec::LW>(role, col, out);
}
#else
#define TACC_DECL(name, ...)
#define TACC_AT(name, a, ...) acc_gather<real>(P, __VA_ARGS__, a)
#endif

template <class real>
__device__ __forceinline__ void fill_tables(const DevPlan& P, const Tx& c, float tf) {
  const int G = DIM(G), F = DIM(F), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P);
  for (int i = c.tl; i < n_lc * PP * G; i += c.tn) EV(mpt0, i) = lerp_rn(EV(mpy, i), EV(mpg, i), tf);
  for (int i = c.tl; i < n_lc * F * G * G; i += c.tn) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G);  // lv = lc*F + v
    float ft = lerp_rn(EV(ft_y, lv * G + u), EV(ft_gap, lv * G + u), tf);
    int a = (lv * G + u0) * G + u, b = (lv * G + u) * G + u0;
    float fa = lerp_rn(EV(fi_y, a), EV(fi_gap, a), tf), fb = lerp_rn(EV(fi_y, b), EV(fi_gap, b), tf);
    EV(Tnum, i) = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (DIM(has_ft_ops)) EV(Tden, i) = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + b);
  }
  const int FG = F * G, nFG = n_lc * FG;
  for (int i = c.tl; i < NG * nFG; i += c.tn) {
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
}

// ------------------------------------------------------------------------------------------
// right-hand side, split by role (see the header). Stage input: y = X + hh * sum_{j<nj} co[j] K_j.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void rhs(const DevPlan& P, const Tx& c, bool on, bool dyn, double t, int ex_bits, int kslot,
                                    real bw, real ew, int acc_mode, double hh, int nj, const real (&co)[6],
                                    const double (&coS)[6]) {
  const int G = DIM(G), F = DIM(F), LG = DIM(LG), nnz = DIM(nnz), NG = DIM(NG), n_lc = DIM(n_lc), PP = DIM(P),
            C = DIM(C), nE = DIM(E), nS = DIM(has_src64) ? DIM(n_slot) : 0, R = NROLE, NSRC = DIM(NSRC);
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83
  const int k0 = kslot * C;
  // ---- phase 0: stage input of my compartments, my share of the time-t tables
  if (on) {
#pragma unroll
    for (int k = 0; k < C; k++)
      if (COMP_ROLE(k) == c.role) {
        real dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++)
          if (j < nj) dy += LA(K, j * C + k) * co[j];
        LA(ys, k) = (real)(LA(X, k) + (double)dy * hh);
      }
    if (c.role == 0)
      for (int q = 0; q < nS; q++) {
        double dy = 0;
#pragma unroll
        for (int j = 0; j < 6; j++)
          if (j < nj) dy += LA(KS, j * nS + q) * coS[j];
        LA(ysS, q) = LA(X, SLOT_C1(q)) + dy * hh;
      }
    if (dyn) fill_tables<real>(P, c, tf);
  }
  TSYNC();
  // ---- phase A: alive (quantity 0) and the weighted infectious sums (quantities 1..NG), quantity q on role q mod R
  if (on) {
    for (int q = c.role; q < 1 + NG; q += R) {
#ifdef EPI_SPEC
      LA(cmA, q) = spec::phaseA_q(q, &LA(ys, 0));
#else
      real x = 0;
      if (q == 0) {
        for (int k = 0; k < C; k++) x += LA(ys, k) * (real)__ldg(P.alive_coef + k);
      } else {
        for (int w = __ldg(P.ig_w_off + q - 1); w < __ldg(P.ig_w_off + q); w++)
          x += (real)__ldg(P.ig_w + w) * LA(ys, __ldg(P.ig_w_c2 + w));
      }
      LA(cmA, q) = x;
#endif
    }
  }
  TSYNC();
  // ---- phase B: CSC gather over source locales (compute_foi_entry inner loop, :105-111)
  if (on && NG > 0) {
    for (int q = c.role; q < 1 + NG; q += R) {
      real a = 0;
#ifdef EPI_SPEC
      if (spec::dense_mob) {
#pragma unroll
        for (int i = 0; i < (spec::dense_mob ? spec::L : 1); i++) a += LO(cmA, q, c.l0 + i * G + c.u) * (real)c.mB[i];
      } else
#endif
      for (int w = __ldg(P.csc_indptr + c.j); w < __ldg(P.csc_indptr + c.j + 1); w++)
        a += LO(cmA, q, c.l0 + __ldg(P.csc_indices + w) * G + c.u) * (real)__ldg(P.mx_csc + c.u * nnz + w);
      LA(cmB, q) = a;
    }
  }
  TSYNC();
  // ---- phase C: G x G facility mixing -> force of infection -> s[k] for my (j, u0), group k on role k mod R
  if (on) {
    for (int k = c.role; k < NG; k += R) {
      real sk = 0;
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
        real num = 0;
        for (int u = 0; u < G; u++) num += LO(cmB, 1 + k, row + u) * (real)EV(Tnum, tb + u);
        real foi = fdiv(num, den);
        sk += EV(coefS, ((k * n_lc + c.lc) * F + v) * G + c.u) * foi;
      }
      LA(cmA, 1 + k) = sk;  // cmA[1..] is free again: phase B has been read out
    }
  }
  TSYNC();
  // ---- phase D + E: CSR gather over destination locales -> the group flows; my regular edges -> fl
  if (on) {
    for (int k = c.role; k < NG; k += R) {
      real rs = 0;
#ifdef EPI_SPEC
      if (spec::dense_mob) {
#pragma unroll
        for (int i = 0; i < (spec::dense_mob ? spec::L : 1); i++) rs += LO(cmA, 1 + k, c.l0 + i * G + c.u) * (real)c.mD[i];
      } else
#endif
      for (int w = __ldg(P.csr_indptr + c.j); w < __ldg(P.csr_indptr + c.j + 1); w++)
        rs += LO(cmA, 1 + k, c.l0 + __ldg(P.csr_indices + w) * G + c.u) * (real)__ldg(P.mx_csr + c.u * nnz + w);
#ifdef EPI_SPEC
      LA(fl, nE + k) = LA(ys, spec::ig_c0[k < spec::NG ? k : 0]) * rs;
#else
      LA(fl, nE + k) = LA(ys, __ldg(P.ig_c0 + k)) * rs;
#endif
    }
    const real alive = LA(cmA, 0);
    const float* mp = &EV(mpt0, c.lc * PP * G + c.u);  // parameters from facility 0 only (:48)
#ifdef EPI_SPEC
    spec::flows_role(c.role, &LA(ys, 0), alive, mp, &LA(fl, 0));
#else
    for (int e = c.role; e < nE; e += R) {
      const int* ep = P.ep + __ldg(P.ep_off + e);
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
#endif
  }
  TSYNC();
  // ---- phase F + G: scatter into dX for my compartments; clamped moves on role 0 (slot compartments live there)
  if (on) {
#ifdef EPI_SPEC
    spec::dx_role(c.role, &LA(fl, 0), &LA(K, k0));
#else
    for (int k = 0; k < C; k++)
      if (COMP_ROLE(k) == c.role) {
        real d = 0;
        for (int q = __ldg(P.xg_off + k); q < __ldg(P.xg_off + k + 1); q++)
          d += LA(fl, __ldg(P.xg_src + q)) * (real)__ldg(P.xg_coef + q);
        LA(K, k0 + k) = d;
      }
#endif
    if (c.role == 0) {
      // next = X + dX, clamp, dX = next - X in double in BOTH modes (core_optimize.py:196-208)
      for (int sl = 0; sl < DIM(n_slot); sl++) {
        real nv = 0;
        const int c1 = SLOT_C1(sl), c2 = SLOT_C2(sl);
        if (((c.cover >> sl) & 1) && ((ex_bits >> sl) & 0x10001)) {
          float my = ((ex_bits >> sl) & 1) ? LA(mvY, sl) : 0.0f;
          float md = ((ex_bits >> (16 + sl)) & 1) ? LA(mvD, sl) : 0.0f;
          float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
          double y1 = DIM(has_src64) ? LA(ysS, sl) : (double)LA(ys, c1), y2 = (double)LA(ys, c2);
          double n1 = y1 + (double)LA(K, k0 + c1), n2 = y2 + (double)LA(K, k0 + c2);
          double nvd = (double)v < n1 ? (double)v : n1;
          double k1 = (n1 - nvd) - y1;
          LA(K, k0 + c1) = (real)k1;
          LA(K, k0 + c2) = (real)((n2 + nvd) - y2);
          nv = (real)nvd;
          if (DIM(has_src64)) LA(KS, kslot * DIM(n_slot) + sl) = k1;
        } else if (DIM(has_src64)) {
          LA(KS, kslot * DIM(n_slot) + sl) = (double)LA(K, k0 + c1);
        }
        LA(fl, nE + NG + sl) = nv;
      }
    }
    // ---- running flow sums of my flows (the slot flows belong to role 0, which has just written them)
    if (acc_mode != 0) {
#pragma unroll
      for (int i = 0; i < NSRC; i++) {
        const int owner = i >= nE + NG ? 0 : i % R;
        if (owner != c.role) continue;
        real f = LA(fl, i);
        real2 p;
        if (acc_mode == 1) {
          p = LA(FBE, i);
          p.x += bw * f;
          p.y += ew * f;
        } else {
          p.x = bw * f;
          p.y = ew * f;
        }
        LA(FBE, i) = p;
      }
    }
  }
  TSYNC();  // K, tables and cmA are read / rewritten by the next evaluation
}

// ------------------------------------------------------------------------------------------
// day prologue: the cell kernel's, with per-env table work strided over the env's (role, cell) lanes and the
// per-cell compartment loops split by role
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ int prologue(const DevPlan& P, const Tx& c, bool on, int variant, double* Xprev_g,
                                        int ex_y_bits, bool& dyn, bool& tabs_current) {
  const int LG = DIM(LG), G = DIM(G), L = DIM(L), C = DIM(C), F = DIM(F), PP = DIM(P), R = NROLE;
  // 1. select sums (executor.py:346-359, ['Value'].sum())
  for (int s = 0; s < P.n_sel; s++) {
    double part = 0;
    if (on && P.sel_l[s * L + c.j] && P.sel_g[s * G + c.u]) {
      int mode = P.sel_mode[s];
      for (int k = c.role; k < C; k += R)
        if (P.sel_c[s * C + k]) {
          if (mode == 0) part += LA(X, k);
          else if (mode == 1) part += P.X0[k * LG + c.cell];
          else part += LA(X, k) - Xprev_g[P.chg_slot_of_comp[k] * LG + c.cell];
        }
    }
    double tot;
    TSUM(tot, part, on);
    if (on && c.tl == 0) EV(sel, s) = tot;
  }
  TSYNC();
  if (on) {
    // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
    for (int k = c.role; k < C; k += R) {
      int sl = P.chg_slot_of_comp[k];
      if (sl >= 0) Xprev_g[sl * LG + c.cell] = LA(X, k);
    }
    // 2. expressions
    for (int e = c.tl; e < P.n_expr; e += c.tn) {
      int itv = P.expr_itv[e];
      bool live = EV(act, itv) && P.expr_prog[e] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
      EV(expr, e) = live ? eval_expr(P, e, &EV(cpv, 0), &EV(sel, 0)) : 0.0;
    }
  }
  TSYNC();
  if (on) {
    // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
    for (int k = c.tl; k < P.n_cls; k += c.tn) {
      float m = 1.0f;
      for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
        int o = P.cls_op[q];
        if (op_live(P, &EV(act, 0), variant, o)) m = __fmul_rn(m, (float)EV(expr, P.op_expr[o]));
      }
      EV(mult_d, k) = m;
    }
  }
  TSYNC();
  {
    double diff = 0, tot;
    if (on)
      for (int k = c.tl; k < P.n_cls; k += c.tn) diff += EV(mult_d, k) != EV(mult_y, k) ? 1.0 : 0.0;
    TSUM(tot, diff, on);
    dyn = tot > 0;
  }
  const bool rebuild = dyn || !tabs_current;
  int ex_d = 0;
  if (on && rebuild) {
    // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
    for (int row = c.tl; row < P.n_lc * F * G; row += c.tn) {
      float a[32], b[32];  // G <= 32 (checked on the host)
      facility_row_fi(P, &EV(mult_y, 0), row, a);
      facility_row_fi(P, &EV(mult_d, 0), row, b);
      for (int g2 = 0; g2 < G; g2++) {
        EV(fi_y, row * G + g2) = a[g2];
        EV(fi_gap, row * G + g2) = __fsub_rn(b[g2], a[g2]);
      }
    }
    for (int it = c.tl; it < P.n_lc * G; it += c.tn) {
      int lc = it / G, g = it % G;
      float a[32], b[32];  // F <= 32 (checked on the host)
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
    for (int i = c.tl; i < P.n_lc * PP * G; i += c.tn) {
      int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
      int idx = ((lc * PP + p) * F + 0) * G + g;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, EV(mult_y, cls)), d = __fmul_rn(m0, EV(mult_d, cls));
      EV(mpy, i) = y;
      EV(mpg, i) = __fsub_rn(d, y);
    }
    const int nFG = P.n_lc * F * G;
    for (int i = c.tl; i < P.n_igp * nFG; i += c.tn) {
      int pi = i / nFG, r = i - pi * nFG, lc = r / (F * G), fg = r - lc * F * G;
      int idx = (lc * PP + __ldg(P.igp + pi)) * F * G + fg;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, EV(mult_y, cls)), d = __fmul_rn(m0, EV(mult_d, cls));
      EV(mpFy, i) = y;
      EV(mpFg, i) = __fsub_rn(d, y);
    }
  }
  // 5. move table of today (nb_move, executor.py:166-234; different compartment, same locale): role 0
  if (on && c.role == 0)
    for (int sl = 0; sl < DIM(n_slot); sl++) LA(mvD, sl) = 0.0f;
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    bool live = on && op_live(P, &EV(act, 0), variant, o);
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    bool mine = live && c.role == 0 && P.mv_g[mo * G + c.u] && P.mv_l1[mo * L + c.j];
    double part = 0;
    if (mine)
      for (int q = 0; q < nc1; q++) part += LA(X, c1s[q]);
    double depart_sum;
    TSUM(depart_sum, part, on);
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
  if (on) {
    // 6. cost rates of today (nb_add, executor.py:236-243)
    for (int it = c.tl; it < P.I * L; it += c.tn) {
      int i = it / L, l = it % L;
      double cr = 0;
      for (int ao = 0; ao < P.n_add_op; ao++) {
        int o = P.add_op[ao];
        if (P.add_i[ao * P.I + i] && P.add_l[ao * L + l] && op_live(P, &EV(act, 0), variant, o))
          cr += EV(expr, P.op_expr[o]) / (double)P.add_parts[ao];
      }
      EV(crD, it) = cr;
    }
  }
  TSYNC();
  if (on && !dyn && rebuild) fill_tables<real>(P, c, 0.0f);  // constant over the day: filled once
  tabs_current = !dyn;
  TSYNC();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: the RK45 state machine of epi_cell.cuh with per-role shares of every per-cell loop
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void rk45_day(const DevPlan& P, const Tx& c, bool on, bool dyn, int ex_bits, int* trace,
                                         double* last_err_out, double* dbg) {
  const double rtol = 1e-3, atol = 1e-6;
  const real rtolr = (real)1e-3, atolr = (real)1e-6;
  const int C = DIM(C), nC = DIM(I) * DIM(L), nS = DIM(has_src64) ? DIM(n_slot) : 0, NSRC = DIM(NSRC),
            n_acc = DIM(n_acc), R = NROLE, nE = DIM(E), NG = DIM(NG);
  const double inv_n = 1.0 / (double)P.n_flat;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0, n_att = 0, guard = 0;
  double last_err = 0, t = 0.0, h_abs = 0, h = 0, t_new = 0, h0 = 0, d0 = 0, d1 = 0;
  bool run = on, rejected = false, redo = false;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };
  auto mine_flow = [&](int i) { return (i >= nE + NG ? 0 : i % R) == c.role; };

  int phase = 0;
  while (phase >= 0) {
    const int stage = phase <= 1 ? phase : (phase == 8 ? 0 : phase - 1);  // K slot written
    if (phase == 2 && run) {
      double min_step = 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t);  // nextafter(t, inf), t >= 0
      if (h_abs < min_step) {
        if (rejected) { status = -1; run = false; }
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
    const bool ev = phase <= 1 ? on : (phase == 8 ? redo : run);
    double t_eval = 0, hh = 0;
    int nj = 0;
    real bw = 0, ew = 0;
    if (phase == 1) { t_eval = h0; hh = h0; nj = 1; }
    else if (phase >= 2 && phase <= 6) { t_eval = t + RK_C[stage] * h; hh = h; nj = stage; bw = rkB(real(0), stage); ew = rkE(real(0), stage); }
    else if (phase == 7) { t_eval = t + h; hh = h; nj = 6; ew = rkE(real(0), 6); }
    else if (phase == 8) { t_eval = t; bw = rkB(real(0), 0); ew = rkE(real(0), 0); }
    real co[6];
    double coS[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      co[j] = phase == 7 ? rkB(real(0), j) : (phase == 1 ? (real)1 : rkA(real(0), stage < 6 ? stage : 0, j < 5 ? j : 4));
      coS[j] = phase == 7 ? RK_B[j] : (phase == 1 ? 1.0 : RK_A[stage < 6 ? stage : 0][j < 5 ? j : 4]);
    }
    rhs<real>(P, c, ev, dyn, t_eval, ex_bits, stage, bw, ew, phase <= 1 ? 0 : (phase == 8 ? 2 : 1), hh, nj, co, coS);
    if (phase == 0) {
      // select_initial_step (common.py:68-134), first half
      real a0 = 0, a1 = 0;
      double s0 = 0, s1 = 0;
      if (on) {
#pragma unroll
        for (int k = 0; k < C; k++)
          if (COMP_ROLE(k) == c.role) {
            real yv = (real)LA(X, k), sc = atolr + fabs(yv) * rtolr;
            real a = fdiv(yv, sc), b = fdiv(LA(K, k), sc);
            a0 += a * a; a1 += b * b;
          }
        {
          TACC_DECL(g0, COL_fl);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++)
            if (a_ % R == c.role) {
              real yv = LA(accY, a_), sc = atolr + fabs(yv) * rtolr;
              real a = fdiv(yv, sc), b = fdiv(TACC_AT(g0, a_, COL_fl), sc);
              a0 += a * a; a1 += b * b;
            }
        }
        s0 = (double)a0; s1 = (double)a1;
        for (int i = c.tl; i < nC; i += c.tn) {
          double yv = EV(cost, i), sc = atol + fabs(yv) * rtol;
          double a = yv / sc, b = EV(crY, i) / sc;  // q(0) = 0
          s0 += a * a; s1 += b * b;
        }
      }
      TSUM(d0, s0, on);
      TSUM(d1, s1, on);  // (its barriers also order the gathers above before the copy below)
      if (on)
        for (int i = 0; i < NSRC; i++)
          if (mine_flow(i)) LA(FBE, i).y = LA(fl, i);  // stage-0 flows, needed once more in phase 1
      d0 = sqrt(d0 * inv_n);
      d1 = sqrt(d1 * inv_n);
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, 1.0);
      nfev++;
      phase = 1;
    } else if (phase == 1) {
      real a2 = 0;
      double s2 = 0, d2, q0 = qd(h0);
      if (on) {
#pragma unroll
        for (int k = 0; k < C; k++)
          if (COMP_ROLE(k) == c.role) {
            real sc = atolr + fabs((real)LA(X, k)) * rtolr, a = fdiv(LA(K, C + k) - LA(K, k), sc);
            a2 += a * a;
          }
        {
          TACC_DECL(g1, COL_fl);
          TACC_DECL(g0, COL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++)
            if (a_ % R == c.role) {
              real sc = atolr + fabs(LA(accY, a_)) * rtolr;
              real a = fdiv(TACC_AT(g1, a_, COL_fl) - TACC_AT(g0, a_, COL_FE), sc);
              a2 += a * a;
            }
        }
        s2 = (double)a2;
        for (int i = c.tl; i < nC; i += c.tn) {
          double sc = atol + fabs(EV(cost, i)) * rtol, a = ((EV(crD, i) - EV(crY, i)) * q0) / sc;
          s2 += a * a;
        }
      }
      TSUM(d2, s2, on);  // (barriers: every role has gathered from FE before it is overwritten)
      if (on)
        for (int i = 0; i < NSRC; i++)
          if (mine_flow(i)) {  // stage-0 contributions of the first attempt
            real f = LA(FBE, i).y;
            real2 p;
            p.x = rkB(real(0), 0) * f;
            p.y = rkE(real(0), 0) * f;
            LA(FBE, i) = p;
          }
      d2 = sqrt(d2 * inv_n) / h0;
      double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : ctrl_pow(real(0), 0.01 / fmax(d1, d2), 0.2);
      h_abs = fmin(fmin(100 * h0, h1), 1.0);
      nfev++;
      phase = __syncthreads_or(run) ? 2 : -1;
    } else if (phase < 7) {
      phase++;
    } else if (phase == 7) {
      // error norm over ALL n_flat components (common.py:63-65, rk.py:146-148), accept / reject
      real ae = 0;
      double e2 = 0, qs[7];
      if (run) {
        nfev += 6;
#pragma unroll
        for (int k = 0; k < C; k++)
          if (COMP_ROLE(k) == c.role) {
            real sc = atolr + fmax(fabs((real)LA(X, k)), fabs(LA(ys, k))) * rtolr;
            real e = 0;
#pragma unroll
            for (int j = 0; j < 7; j++) e += LA(K, j * C + k) * rkE(real(0), j);
            real a = fdiv(e * (real)h, sc);
            ae += a * a;
          }
        {
          TACC_DECL(gB, COL_FB);
          TACC_DECL(gE, COL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++)
            if (a_ % R == c.role) {
              real yv = LA(accY, a_), yn = yv + (real)h * TACC_AT(gB, a_, COL_FB);
              real sc = atolr + fmax(fabs(yv), fabs(yn)) * rtolr, a = fdiv(TACC_AT(gE, a_, COL_FE) * (real)h, sc);
              ae += a * a;
            }
        }
        e2 = (double)ae;
#pragma unroll
        for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
        for (int i = c.tl; i < nC; i += c.tn) {
          double cy = EV(crY, i), cg = EV(crD, i) - cy, sb = 0, se = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) {
            double k = cy + cg * qs[j];
            sb += k * RK_B[j];
            se += k * RK_E[j];
          }
          double yv = EV(cost, i), yn = yv + h * sb;
          double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = se * h / sc;
          e2 += a * a;
        }
      }
      double err;
      TSUM(err, e2, run);
      err = sqrt(err * inv_n);
      bool accepted = false;
      if (run) {
        last_err = err;
        if (dbg && c.tl == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
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
        // commit: y = y_new, f = f_new (state advanced in double in both modes); every role its own share.
        // accY first: its gather reads other roles' FB entries, which are re-initialised after the barrier below.
        const real hr = (real)h;
        {
          TACC_DECL(gB, COL_FB);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++)
            if (a_ % R == c.role) LA(accY, a_) += hr * TACC_AT(gB, a_, COL_FB);
        }
#pragma unroll
        for (int k = 0; k < C; k++)
          if (COMP_ROLE(k) == c.role) {
            real dy = 0;
#pragma unroll
            for (int j = 0; j < 6; j++) dy += LA(K, j * C + k) * rkB(real(0), j);
            LA(X, k) = LA(X, k) + h * (double)dy;
            LA(K, k) = LA(K, 6 * C + k);
          }
        if (c.role == 0)
          for (int q = 0; q < nS; q++) {
            LA(X, SLOT_C1(q)) = LA(ysS, q);  // = X + h * sum_j B_j KS_j in double
            LA(KS, q) = LA(KS, 6 * nS + q);
          }
        for (int i = c.tl; i < nC; i += c.tn) {
          double cy = EV(crY, i), cg = EV(crD, i) - cy, sb = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
          EV(cost, i) += h * sb;
        }
        nsteps++;
        t = t_new;
        rejected = false;
      }
      TSYNC();
      if (run && accepted) {
        for (int i = 0; i < NSRC; i++)
          if (mine_flow(i)) {  // the last stage's flows open the next attempt (FSAL)
            real f = LA(fl, i);
            real2 p;
            p.x = rkB(real(0), 0) * f;
            p.y = rkE(real(0), 0) * f;
            LA(FBE, i) = p;
          }
        if (!(t < 1.0)) run = false;
      }
      phase = __syncthreads_or(redo) ? 8 : (__syncthreads_or(run) ? 2 : -1);
    } else {
      redo = false;
      phase = __syncthreads_or(run) ? 2 : -1;
    }
  }
  if (on && c.tl == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
}

// ------------------------------------------------------------------------------------------
// EpiEnv.step for the group of envs of this CTA (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void step_cta(const DevPlan& P, const DevState& S, const StepArgs& a, char* smem, int tid,
                                         int group) {
  Tx c;
  const int R = NROLE;
  c.role = tid / EPI_TWARP;
  cell::bind(P, smem, tid - c.role * EPI_TWARP, c);
  c.tl = c.role * DIM(LG) + c.cell;
  c.tn = R * DIM(LG);
  const int C = DIM(C), CLG = DIM(CLG), LG = DIM(LG), nA = DIM(n_acc) * LG, nC = DIM(I) * DIM(L), nM = DIM(n_slot) * LG;
  const int env = group * P.cl.epw + c.slot;
  const bool on0 = c.valid && env < a.N;
  const size_t e = on0 ? (size_t)env : 0;

  if (!DIM(has_ft_ops) && c.valid)
    for (int i = c.role * EPI_LW + c.lane; i < DIM(n_lc) * DIM(F) * DIM(G) * DIM(G); i += R * EPI_LW)
      WT(tdc, i) = sizeof(real) == 8 ? (real)__ldg(P.Tden_const64 + i) : (real)__ldg(P.Tden_const + i);
  double* gX = S.X + e * CLG;
  real* gAcc = (real*)S.acc + e * nA + c.cell;
  double* gXprev = S.Xprev + e * P.n_chg * LG;
  int* meta = S.meta + e * 8;
  if (on0) {
    for (int k = c.role; k < C; k += R) LA(X, k) = gX[k * LG + c.cell];
    if (c.role == 0)
      for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = S.mvY[e * nM + s * LG + c.cell];
    for (int a_ = c.role; a_ < DIM(n_acc); a_ += R) LA(accY, a_) = gAcc[a_ * LG];
    for (int i = c.tl; i < nC; i += c.tn) {
      EV(cost, i) = S.cost[e * nC + i];
      EV(crY, i) = S.crY[e * nC + i];
    }
    for (int i = c.tl; i < P.n_cls; i += c.tn) EV(mult_y, i) = S.multY[e * P.n_cls + i];
    for (int k = c.tl; k < P.NCP; k += c.tn) {
      float v;
      if (a.cpv) v = a.cpv[e * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[e * P.A + s] : P.cp_fixed[k]; }
      EV(cpv, k) = v;
    }
    for (int i = c.tl; i < P.I; i += c.tn)
      EV(act, i) = a.active ? a.active[e * P.I + i] : (a.init_only ? P.init_active[i] : 1);
  }
  int t_day = on0 ? meta[0] : 0, ex_y = on0 ? meta[1] : 0;
  TSYNC();

  double reward = 0;
  bool on = on0, tabs_current = false;
  const int n_days = a.init_only ? 1 : a.n_days;
  for (int day = 0; day < n_days; day++) {
    if (!__syncthreads_or(on)) break;
    bool dyn = true;
    int ex_bits = prologue<real>(P, c, on, a.variant, gXprev, ex_y, dyn, tabs_current);
    if (!a.init_only) {
      double c0 = 0, c1 = 0;
      if (on)
        for (int i = c.tl; i < nC; i += c.tn) c0 += EV(cost, i);
      rk45_day<real>(P, c, on, dyn, ex_bits, meta, S.last_err + e, a.dbg_attempts ? a.dbg_attempts + e * 128 : nullptr);
      if (on)
        for (int i = c.tl; i < nC; i += c.tn) c1 += EV(cost, i);
      double dr;
      TSUM(dr, c1 - c0, on);  // reward = -sum(cumulative_cost_next - cumulative_cost_current), epidemic.py:115
      if (on) { reward -= dr; t_day++; }
    }
    if (on) {
      for (int i = c.tl; i < P.n_cls; i += c.tn) EV(mult_y, i) = EV(mult_d, i);
      if (c.role == 0)
        for (int s = 0; s < DIM(n_slot); s++) LA(mvY, s) = LA(mvD, s);
      for (int i = c.tl; i < nC; i += c.tn) EV(crY, i) = EV(crD, i);
      ex_y = (ex_bits >> 16) & 0xffff;
    }
    TSYNC();
    if (on && t_day >= P.horizon) on = false;  // the reference's EpiEnv.step stops its day loop at done
  }
  if (on0) {
    for (int k = c.role; k < C; k += R) {
      double x = LA(X, k);
      gX[k * LG + c.cell] = x;
      if (a.obs_out) ((real*)a.obs_out)[e * CLG + k * LG + c.cell] = (real)x;
    }
    if (c.role == 0)
      for (int s = 0; s < DIM(n_slot); s++) S.mvY[e * nM + s * LG + c.cell] = LA(mvY, s);
    for (int a_ = c.role; a_ < DIM(n_acc); a_ += R) gAcc[a_ * LG] = LA(accY, a_);
    for (int i = c.tl; i < nC; i += c.tn) {
      S.cost[e * nC + i] = EV(cost, i);
      S.crY[e * nC + i] = EV(crY, i);
    }
    for (int i = c.tl; i < P.n_cls; i += c.tn) S.multY[e * P.n_cls + i] = EV(mult_y, i);
    if (c.tl == 0) {
      meta[0] = t_day; meta[1] = ex_y;
      if (a.reward_out) a.reward_out[env] = reward;
      if (a.done_out) a.done_out[env] = (uint8_t)(t_day >= P.horizon);
    }
  }
}

#ifndef EPI_HOST_EMU
template <class real>
__global__ void __launch_bounds__(128) cell_team_kernel(const DevPlan P, const DevState S, const StepArgs a) {
#ifdef EPI_SPEC
  step_cta<real>(P, S, a, cell::cell_smem, threadIdx.x, blockIdx.x);
#else
  extern __shared__ __align__(16) char smem[];
  step_cta<real>(P, S, a, smem, threadIdx.x, blockIdx.x);
#endif
}
#endif

}  // namespace cteam
}  // namespace epi
