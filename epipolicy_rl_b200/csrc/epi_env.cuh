// epi_env.cuh — the "env" kernel: one CTA per environment, threads strided over its (locale, group) cells
// (sm_100a). The lane-per-cell mapping of epi_cell.cuh for scenarios with L*G > 32, up to the synthetic
// 1000-locale x 16-group configuration (BASELINE.json configs[3]).
//
//   * every thread owns the cells tid, tid + T, ...; per cell it runs the same serial per-cell program as the
//     cell kernel (generated straight-line code under EPI_SPEC, interpreted tables otherwise);
//   * the per-cell RK45 working set (state, 7 stage derivatives, flow sums, accumulators) is a set of
//     [word][cell] columns — cells fastest, so a warp's accesses are coalesced — placed in shared memory when
//     the env fits and in a per-CTA global workspace otherwise (then the kernel streams it once per
//     right-hand-side evaluation: this is the HBM-bound regime);
//   * the force of infection crosses cells: alive / weighted-infectious columns -> CSC gather over source
//     locales -> G x G facility mixing over the locale's groups -> CSR gather over destination locales,
//     with block barriers between the four per-cell passes of one evaluation;
//   * one env per CTA: the scipy controller state (t, h, accept/reject) is uniform over the CTA, the error
//     norm is a block reduction; a persistent grid walks the env axis.
// Arithmetic, rounding points and the RK45 state machine are those of epi_cell.cuh / epi_kernels.cuh.
#pragma once
#include "epi_kernels.cuh"
#include "epi_rk.cuh"

namespace epi {
namespace env {
using namespace rk;

#ifndef EPI_FSAL_DAYS
#define EPI_FSAL_DAYS 1  // fp32 mode: first-same-as-last across the day boundary of a fused launch
#endif

#ifdef EPI_SPEC
#define EDIM(x) (spec::x)
#define ESLOT_C1(i) (spec::slot_c1[i])
#define ESLOT_C2(i) (spec::slot_c2[i])
#else
#define EDIM(x) (P.x)
#define ESLOT_C1(i) (P.slot_c1[i])
#define ESLOT_C2(i) (P.slot_c2[i])
#endif

struct Bx {
  int tid, nt;
  char* big;    // per-cell columns X, ys, K, fl, FBE, accY, mv*, ysS, KS
  char* comm;   // per-cell columns cmA, cmB (the gathers' random-access targets)
  char* tab;    // per-env tables (shared memory)
  double* red;  // [nt] reduction scratch (shared memory)
  double *cost, *crY, *crD;  // [I*L] cumulative cost, yesterday's / today's cost rate of the current env (global)
  // flow-slot mode: the accumulators y and the candidate y_new of the current attempt are two columns that swap
  // roles when the attempt is accepted (no copy); byte offsets into `big`
  int64_t o_acc, o_accn;
};
#define ACCY(i) (((real*)(b.big + b.o_acc))[(size_t)(i) * EDIM(LG) + cell])
#define ACCYN(i) (((real*)(b.big + b.o_accn))[(size_t)(i) * EDIM(LG) + cell])

// element i of cell `cell` in a column; entry i of an env table
#define BA(arr, i) (((T_##arr*)(b.big + P.el.l_##arr))[(size_t)(i) * EDIM(LG) + cell])
#define CM(arr, i, cl) (((T_##arr*)(b.comm + P.el.l_##arr))[(size_t)(i) * EDIM(LG) + (cl)])
#define ET(arr, i) (((T_##arr*)(b.tab + P.el.e_##arr))[i])
#define EXF (sizeof(real) == 4)  // fp32 mode keeps a float copy of the state for the stage inputs and the norms
#define FOR_CELLS for (int cell = b.tid; cell < EDIM(LG); cell += b.nt)

// block sum; every thread receives the bit-identical result
__device__ __forceinline__ double block_sum(const Bx& b, double v) {
  b.red[b.tid] = v;
  __syncthreads();
  int n = b.nt;
  while (n > 1) {
    int half = (n + 1) >> 1;
    if (b.tid < n - half) b.red[b.tid] += b.red[b.tid + half];
    __syncthreads();
    n = half;
  }
  double r = b.red[0];
  __syncthreads();
  return r;
}

#ifdef EPI_SPEC
#define EACC_DECL(name, ...) real name[spec::n_acc > 0 ? spec::n_acc : 1]; spec::acc_rt(__VA_ARGS__, name)
#define EACC_AT(name, a, ...) (name[a])
#define EY_DECL real y[spec::C]
#define EYV(k) (y[k])
#else
#define EACC_DECL(name, ...)
#define EACC_AT(name, a, ...) acc_gather<real>(P, __VA_ARGS__, a)
#define EY_DECL
#define EYV(k) BA(ys, k)
template <class real>
__device__ __forceinline__ real acc_gather(const DevPlan& P, const real* col, size_t stride, int a) {
  real d = 0;
  for (int q = __ldg(P.ag_off + a); q < __ldg(P.ag_off + a + 1); q++)
    d += col[(size_t)__ldg(P.ag_src + q) * stride] * (real)__ldg(P.ag_coef + q);
  return d;
}
#endif
// flow-slot mode (global-workspace regime): the 7 stages' flows are stored (one 35-word write per evaluation)
// and combined once per attempt, instead of read-modify-writing the packed flow sums in every evaluation
#ifdef EPI_SPEC
#define EFS (spec::flow_slots)
#else
#define EFS (P.el.flow_slots)
#endif
#define FLW(i) (*(EFS ? &BA(fls, (size_t)fslot * EDIM(NSRC) + (i)) : &BA(fl, i)))
#define ECOL_slot(ps) (&BA(fls, (size_t)(ps) * EDIM(NSRC))), (size_t)EDIM(LG)
#define ECOL_fl (&BA(fl, 0)), (size_t)EDIM(LG)
#define ECOL_FB (&BA(FBE, 0).x), (size_t)2 * EDIM(LG)
#define ECOL_FE (&BA(FBE, 0).y), (size_t)2 * EDIM(LG)

template <class real>
__device__ __forceinline__ void fill_tables(const DevPlan& P, const Bx& b, float tf) {
  const int G = EDIM(G), F = EDIM(F), NG = EDIM(NG), n_lc = EDIM(n_lc), PP = EDIM(P);
  for (int i = b.tid; i < n_lc * PP * G; i += b.nt) ET(mpt0, i) = lerp_rn(ET(mpy, i), ET(mpg, i), tf);
  for (int i = b.tid; i < n_lc * F * G * G; i += b.nt) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G);  // lv = lc*F + v
    float ft = lerp_rn(ET(ft_y, lv * G + u), ET(ft_gap, lv * G + u), tf);
    int a = (lv * G + u0) * G + u, bb = (lv * G + u) * G + u0;
    float fa = lerp_rn(ET(fi_y, a), ET(fi_gap, a), tf), fb = lerp_rn(ET(fi_y, bb), ET(fi_gap, bb), tf);
    // stored TRANSPOSED, [lv][u][u0]: the mixing pass reads T[.][u][u0] with u0 = the lane's group, so a
    // locale's lanes hit consecutive words (no bank conflicts, coalesced)
    const int it = (lv * G + u) * G + u0;
    ET(Tnum, it) = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (EDIM(has_ft_ops)) ET(Tden, it) = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + bb);
  }
  const int FG = F * G, nFG = n_lc * FG;
  for (int i = b.tid; i < NG * nFG; i += b.nt) {
    int k = i / nFG, r = i - k * nFG;  // r = (lc*F + v)*G + u0
    real cf = (real)lerp_rn(ET(ft_y, r), ET(ft_gap, r), tf);
    int pi = __ldg(P.ig_p0i + k);
    cf *= (real)lerp_rn(ET(mpFy, pi * nFG + r), ET(mpFg, pi * nFG + r), tf);
    for (int q = __ldg(P.ig_np_off + k); q < __ldg(P.ig_np_off + k + 1); q++) {
      pi = __ldg(P.ig_npi + q);
      cf *= (real)lerp_rn(ET(mpFy, pi * nFG + r), ET(mpFg, pi * nFG + r), tf);
    }
    for (int q = __ldg(P.ig_dp_off + k); q < __ldg(P.ig_dp_off + k + 1); q++) {
      pi = __ldg(P.ig_dpi + q);
      float m = lerp_rn(ET(mpFy, pi * nFG + r), ET(mpFg, pi * nFG + r), tf);
      if (fabsf(m) > 0) cf /= (real)m;
    }
    ET(coefS, i) = cf;
  }
}

// ------------------------------------------------------------------------------------------
// one right-hand-side evaluation of the env (core_optimize.py:165-222) in four per-cell passes.
// Stage input: y = X + hh * sum_{j<nj} co[j] K_j. kslot / acc_mode / keep_fl as in epi_cell.cuh.
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void rhs(const DevPlan& P, const Bx& b, bool dyn, double t, int ex_bits, int kslot, real bw,
                                    real ew, int acc_mode, bool keep_fl, double hh, int nj, const real (&co)[6],
                                    const double (&coS)[6], int fslot) {
  const int G = EDIM(G), F = EDIM(F), LG = EDIM(LG), nnz = EDIM(nnz), NG = EDIM(NG), n_lc = EDIM(n_lc), PP = EDIM(P),
            C = EDIM(C), nE = EDIM(E), nS = EDIM(has_src64) ? EDIM(n_slot) : 0;
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83
  if (dyn) fill_tables<real>(P, b, tf);
  // ---- pass 1: stage input, alive (N excludes death and hospitalized, :28-41,:168), weighted infectious sums
  FOR_CELLS {
    EY_DECL;
#pragma unroll
    for (int k = 0; k < C; k++) {
      real dy = 0;
#pragma unroll
      for (int j = 0; j < 6; j++)
        if (j < nj) dy += BA(K, j * C + k) * co[j];
      if (EXF) EYV(k) = dy * (real)hh + BA(XF, k);  // fp32 mode: float copy of the state (half the bytes of X)
      else EYV(k) = (real)(BA(X, k) + (double)dy * hh);
    }
    for (int q = 0; q < nS; q++) {
      double dy = 0;
#pragma unroll
      for (int j = 0; j < 6; j++)
        if (j < nj) dy += BA(KS, j * nS + q) * coS[j];
      BA(ysS, q) = BA(X, ESLOT_C1(q)) + dy * hh;
    }
    real alive = 0;
#ifdef EPI_SPEC
    real xw[spec::NG > 0 ? spec::NG : 1];
    spec::phaseA(y, alive, xw);
#pragma unroll
    for (int k = 0; k < C; k++) BA(ys, k) = y[k];
    CM(cmA, 0, cell) = alive;
#pragma unroll
    for (int k = 0; k < NG; k++) CM(cmA, 1 + k, cell) = xw[k];
#else
    for (int k = 0; k < C; k++) alive += BA(ys, k) * (real)__ldg(P.alive_coef + k);
    CM(cmA, 0, cell) = alive;
    for (int k = 0; k < NG; k++) {
      real x = 0;
      for (int q = __ldg(P.ig_w_off + k); q < __ldg(P.ig_w_off + k + 1); q++)
        x += (real)__ldg(P.ig_w + q) * BA(ys, __ldg(P.ig_w_c2 + q));
      CM(cmA, 1 + k, cell) = x;
    }
#endif
  }
  __syncthreads();
  if (NG > 0) {
    // ---- pass 2: CSC gather over source locales (compute_foi_entry inner loop, :105-111)
    FOR_CELLS {
      const int j = cell / G, u = cell - j * G;
      real a = 0, xs[EPI_MAX_NG];
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++) xs[k] = 0;
      for (int q = __ldg(P.csc_indptr + j); q < __ldg(P.csc_indptr + j + 1); q++) {
        int src = __ldg(P.csc_indices + q) * G + u;
        real val = (real)__ldg(P.mx_csc + (size_t)u * nnz + q);
        a += CM(cmA, 0, src) * val;
#pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) xs[k] += CM(cmA, 1 + k, src) * val;
      }
      CM(cmB, 0, cell) = a;
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) CM(cmB, 1 + k, cell) = xs[k];
    }
    __syncthreads();
    // ---- pass 3: G x G facility mixing -> force of infection -> s[k] for (j, u0)  (:112-124, :146-157)
    FOR_CELLS {
      const int j = cell / G, u0 = cell - j * G, lc = __ldg(P.lclass + j);
      real sk[EPI_MAX_NG];
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++) sk[k] = 0;
      const int row = j * G;
#ifdef EPI_SPEC
      // the locale's gathered columns once, in registers (G is a compile-time constant)
      real cb[1 + spec::NG][spec::G];
#pragma unroll
      for (int k = 0; k < 1 + spec::NG; k++)
#pragma unroll
        for (int u = 0; u < spec::G; u++) cb[k][u] = CM(cmB, k, row + u);
#define CB(k, u) cb[k][u]
#else
#define CB(k, u) CM(cmB, k, row + (u))
#endif
      for (int v = 0; v < F; v++) {
        const int tb = (lc * F + v) * G * G + u0;  // transposed tables: element (u0, u) at tb + u*G
        real den = 0;
        if (EDIM(has_ft_ops)) {
#pragma unroll
          for (int u = 0; u < G; u++) den += CB(0, u) * ET(Tden, tb + u * G);
        } else if (sizeof(real) == 8) {
#pragma unroll
          for (int u = 0; u < G; u++) den += CB(0, u) * (real)__ldg(P.Tden_const64T + tb + u * G);
        } else {
#pragma unroll
          for (int u = 0; u < G; u++) den += CB(0, u) * (real)__ldg(P.Tden_constT + tb + u * G);
        }
        if (den <= (real)1e-15) den = 1;
  #pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) {
            real num = 0;
#pragma unroll
            for (int u = 0; u < G; u++) num += CB(1 + k, u) * (real)ET(Tnum, tb + u * G);
            real foi = fdiv(num, den);
            sk[k] += ET(coefS, ((k * n_lc + lc) * F + v) * G + u0) * foi;
          }
      }
#undef CB
#pragma unroll
      for (int k = 0; k < EPI_MAX_NG; k++)
        if (k < NG) CM(cmA, 1 + k, cell) = sk[k];  // cmA[1..] is free again: pass 2 has been read out
    }
    __syncthreads();
  }
  // ---- pass 4: CSR gather over destination locales (:144-158), edges, scatter into dX, clamped moves, flow sums
  FOR_CELLS {
    const int j = cell / G, u = cell - j * G, lc = __ldg(P.lclass + j);
    real rs[EPI_MAX_NG];
#pragma unroll
    for (int k = 0; k < EPI_MAX_NG; k++) rs[k] = 0;
    if (NG > 0) {
      for (int q = __ldg(P.csr_indptr + j); q < __ldg(P.csr_indptr + j + 1); q++) {
        int dst = __ldg(P.csr_indices + q) * G + u;
        real val = (real)__ldg(P.mx_csr + (size_t)u * nnz + q);
#pragma unroll
        for (int k = 0; k < EPI_MAX_NG; k++)
          if (k < NG) rs[k] += CM(cmA, 1 + k, dst) * val;
      }
    }
    const real alive = CM(cmA, 0, cell);
    const float* mp = &ET(mpt0, lc * PP * G + u);  // parameters from facility 0 only (:48)
    const int k0 = kslot * C;
#ifdef EPI_SPEC
    if (EFS) {
      // streaming form: flows go straight to their slot and into dX (see gen_spec: flows_scatter)
      real y[spec::C], d[spec::C];
#pragma unroll
      for (int k = 0; k < C; k++) y[k] = BA(ys, k);
      real* out = &BA(fls, (size_t)fslot * spec::NSRC);
      const bool st = keep_fl || kslot != 1;  // stage 1 of an attempt has B = E = 0: its flows are never combined
      spec::flows_scatter(y, alive, mp, rs, out, (size_t)LG, d, st);
#pragma unroll
      for (int sl = 0; sl < spec::n_slot; sl++) {
        real nv = 0;
        const int c1 = spec::slot_c1[sl], c2 = spec::slot_c2[sl];
        if (__ldg(P.slot_cover + (size_t)sl * LG + cell) && ((ex_bits >> sl) & 0x10001)) {
          float my = ((ex_bits >> sl) & 1) ? BA(mvY, sl) : 0.0f;
          float md = ((ex_bits >> (16 + sl)) & 1) ? BA(mvD, sl) : 0.0f;
          float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
          double y1 = spec::has_src64 ? BA(ysS, sl) : (double)y[c1], y2 = (double)y[c2];
          double n1 = y1 + (double)d[c1], n2 = y2 + (double)d[c2];
          double nvd = (double)v < n1 ? (double)v : n1;
          double k1 = (n1 - nvd) - y1;
          d[c1] = (real)k1;
          d[c2] = (real)((n2 + nvd) - y2);
          nv = (real)nvd;
          if (spec::has_src64) BA(KS, kslot * spec::n_slot + sl) = k1;
        } else if (spec::has_src64) {
          BA(KS, kslot * spec::n_slot + sl) = (double)d[c1];
        }
        if (st) out[(size_t)(nE + NG + sl) * LG] = nv;
      }
#pragma unroll
      for (int k = 0; k < C; k++) BA(K, k0 + k) = d[k];
      continue;
    }
    real y[spec::C], f[spec::NSRC], d[spec::C];
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = BA(ys, k);
    spec::flows(y, alive, mp, f);
#pragma unroll
    for (int k = 0; k < NG; k++) f[nE + k] = y[spec::ig_c0[k]] * rs[k];
    spec::dx(f, d);
#pragma unroll
    for (int sl = 0; sl < spec::n_slot; sl++) {
      real nv = 0;
      const int c1 = spec::slot_c1[sl], c2 = spec::slot_c2[sl];
      if (__ldg(P.slot_cover + (size_t)sl * LG + cell) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? BA(mvY, sl) : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? BA(mvD, sl) : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double y1 = spec::has_src64 ? BA(ysS, sl) : (double)y[c1], y2 = (double)y[c2];
        double n1 = y1 + (double)d[c1], n2 = y2 + (double)d[c2];
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - y1;
        d[c1] = (real)k1;
        d[c2] = (real)((n2 + nvd) - y2);
        nv = (real)nvd;
        if (spec::has_src64) BA(KS, kslot * spec::n_slot + sl) = k1;
      } else if (spec::has_src64) {
        BA(KS, kslot * spec::n_slot + sl) = (double)d[c1];
      }
      f[nE + NG + sl] = nv;
    }
#pragma unroll
    for (int k = 0; k < C; k++) BA(K, k0 + k) = d[k];
    if (EFS) {
#pragma unroll
      for (int i = 0; i < spec::NSRC; i++) BA(fls, (size_t)fslot * spec::NSRC + i) = f[i];
    } else {
      if (acc_mode == 1) {
#pragma unroll
        for (int i = 0; i < spec::NSRC; i++) {
          real2 p = BA(FBE, i);
          p.x += bw * f[i];
          p.y += ew * f[i];
          BA(FBE, i) = p;
        }
      } else if (acc_mode == 2) {
#pragma unroll
        for (int i = 0; i < spec::NSRC; i++) {
          real2 p;
          p.x = bw * f[i];
          p.y = ew * f[i];
          BA(FBE, i) = p;
        }
      }
      if (keep_fl) {
#pragma unroll
        for (int i = 0; i < spec::NSRC; i++) BA(fl, i) = f[i];
      }
    }
#else
    for (int k = 0; k < NG; k++) FLW(nE + k) = BA(ys, __ldg(P.ig_c0 + k)) * rs[k];
    // regular edges (compute_fraction, :43-65)
    const int* ep = P.ep;
    for (int e = 0; e < nE; e++) {
      auto val = [&](int fac) -> real {
        int kind = fac >> 24, idx = fac & 0xffffff;
        if (kind == 0) return (real)mp[idx * G];
        if (kind == 1) return alive;
        return BA(ys, idx);
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
      FLW(e) = denom > 0 ? fdiv(numer, denom) : numer;
    }
    for (int k = 0; k < C; k++) {
      real d = 0;
      for (int q = __ldg(P.xg_off + k); q < __ldg(P.xg_off + k + 1); q++)
        d += FLW(__ldg(P.xg_src + q)) * (real)__ldg(P.xg_coef + q);
      BA(K, k0 + k) = d;
    }
    // clamped moves, slots in op order (core_optimize.py:196-208), double round trip as in epi_cell.cuh
    for (int sl = 0; sl < P.n_slot; sl++) {
      real nv = 0;
      const int c1 = P.slot_c1[sl], c2 = P.slot_c2[sl];
      if (__ldg(P.slot_cover + (size_t)sl * LG + cell) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? BA(mvY, sl) : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? BA(mvD, sl) : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double y1 = P.has_src64 ? BA(ysS, sl) : (double)BA(ys, c1), y2 = (double)BA(ys, c2);
        double n1 = y1 + (double)BA(K, k0 + c1), n2 = y2 + (double)BA(K, k0 + c2);
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - y1;
        BA(K, k0 + c1) = (real)k1;
        BA(K, k0 + c2) = (real)((n2 + nvd) - y2);
        nv = (real)nvd;
        if (P.has_src64) BA(KS, kslot * P.n_slot + sl) = k1;
      } else if (P.has_src64) {
        BA(KS, kslot * P.n_slot + sl) = (double)BA(K, k0 + c1);
      }
      FLW(nE + NG + sl) = nv;
    }
    if (EFS) {
    } else if (acc_mode == 1) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = BA(fl, i);
        real2 p = BA(FBE, i);
        p.x += bw * f;
        p.y += ew * f;
        BA(FBE, i) = p;
      }
    } else if (acc_mode == 2) {
      for (int i = 0; i < P.NSRC; i++) {
        real f = BA(fl, i);
        real2 p;
        p.x = bw * f;
        p.y = ew * f;
        BA(FBE, i) = p;
      }
    }
#endif
  }
  __syncthreads();  // tables and cmA are rewritten by the next evaluation
}

// ------------------------------------------------------------------------------------------
// day prologue (Executor.execute + get_parameter_from_delta + get_gap_parameter) for the CTA's env
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ int prologue(const DevPlan& P, const Bx& b, int variant, double* Xprev_g, int ex_y_bits,
                                        bool& dyn, bool& tabs_current) {
  const int LG = EDIM(LG), G = EDIM(G), L = EDIM(L), C = EDIM(C), F = EDIM(F), PP = EDIM(P);
  // 1. select sums (executor.py:346-359, ['Value'].sum())
  for (int s = 0; s < P.n_sel; s++) {
    double part = 0;
    const int mode = P.sel_mode[s];
    FOR_CELLS {
      const int j = cell / G, u = cell - j * G;
      if (P.sel_l[s * L + j] && P.sel_g[s * G + u])
        for (int k = 0; k < C; k++)
          if (P.sel_c[s * C + k]) {
            if (mode == 0) part += BA(X, k);
            else if (mode == 1) part += P.X0[(size_t)k * LG + cell];
            else part += BA(X, k) - Xprev_g[(size_t)P.chg_slot_of_comp[k] * LG + cell];
          }
    }
    double tot = block_sum(b, part);
    if (b.tid == 0) ET(sel, s) = tot;
  }
  __syncthreads();
  // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
  FOR_CELLS
    for (int k = 0; k < C; k++) {
      int sl = P.chg_slot_of_comp[k];
      if (sl >= 0) Xprev_g[(size_t)sl * LG + cell] = BA(X, k);
    }
  // 2. expressions
  for (int e = b.tid; e < P.n_expr; e += b.nt) {
    int itv = P.expr_itv[e];
    bool live = ET(act, itv) && P.expr_prog[e] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
    ET(expr, e) = live ? eval_expr(P, e, &ET(cpv, 0), &ET(sel, 0)) : 0.0;
  }
  __syncthreads();
  // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
  for (int k = b.tid; k < P.n_cls; k += b.nt) {
    float m = 1.0f;
    for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
      int o = P.cls_op[q];
      if (op_live(P, &ET(act, 0), variant, o)) m = __fmul_rn(m, (float)ET(expr, P.op_expr[o]));
    }
    ET(mult_d, k) = m;
  }
  __syncthreads();
  {
    double diff = 0;
    for (int k = b.tid; k < P.n_cls; k += b.nt) diff += ET(mult_d, k) != ET(mult_y, k) ? 1.0 : 0.0;
    dyn = block_sum(b, diff) > 0;
  }
  const bool rebuild = dyn || !tabs_current;
  if (rebuild) {
    // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
    for (int row = b.tid; row < P.n_lc * F * G; row += b.nt) {
      float a[32], bb[32];  // G <= 32 (checked on the host)
      facility_row_fi(P, &ET(mult_y, 0), row, a);
      facility_row_fi(P, &ET(mult_d, 0), row, bb);
      for (int g2 = 0; g2 < G; g2++) {
        ET(fi_y, row * G + g2) = a[g2];
        ET(fi_gap, row * G + g2) = __fsub_rn(bb[g2], a[g2]);
      }
    }
    for (int it = b.tid; it < P.n_lc * G; it += b.nt) {
      int lc = it / G, g = it % G;
      float a[32], bb[32];  // F <= 32 (checked on the host)
      float s = 0.0f, s2 = 0.0f;
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        a[f] = __fmul_rn(P.ft0[i], ET(mult_y, P.ft_cls[i]));
        bb[f] = __fmul_rn(P.ft0[i], ET(mult_d, P.ft_cls[i]));
        s = __fadd_rn(s, a[f]);
        s2 = __fadd_rn(s2, bb[f]);
      }
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        float ya = s > 1e-15f ? __fdiv_rn(a[f], s) : (f == 0 ? 1.0f : 0.0f);
        float yb = s2 > 1e-15f ? __fdiv_rn(bb[f], s2) : (f == 0 ? 1.0f : 0.0f);
        ET(ft_y, i) = ya;
        ET(ft_gap, i) = __fsub_rn(yb, ya);
      }
    }
    for (int i = b.tid; i < P.n_lc * PP * G; i += b.nt) {
      int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
      int idx = ((lc * PP + p) * F + 0) * G + g;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, ET(mult_y, cls)), d = __fmul_rn(m0, ET(mult_d, cls));
      ET(mpy, i) = y;
      ET(mpg, i) = __fsub_rn(d, y);
    }
    const int nFG = P.n_lc * F * G;
    for (int i = b.tid; i < P.n_igp * nFG; i += b.nt) {
      int pi = i / nFG, r = i - pi * nFG, lc = r / (F * G), fg = r - lc * F * G;
      int idx = (lc * PP + __ldg(P.igp + pi)) * F * G + fg;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float y = __fmul_rn(m0, ET(mult_y, cls)), d = __fmul_rn(m0, ET(mult_d, cls));
      ET(mpFy, i) = y;
      ET(mpFg, i) = __fsub_rn(d, y);
    }
  }
  // 5. move table of today (nb_move, executor.py:166-234; different compartment, same locale)
  FOR_CELLS
    for (int sl = 0; sl < EDIM(n_slot); sl++) BA(mvD, sl) = 0.0f;
  int ex_d = 0;
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    bool live = op_live(P, &ET(act, 0), variant, o);
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    double part = 0;
    if (live)
      FOR_CELLS {
        const int j = cell / G, u = cell - j * G;
        if (P.mv_g[mo * G + u] && P.mv_l1[mo * L + j])
          for (int q = 0; q < nc1; q++) part += BA(X, c1s[q]);
      }
    double depart_sum = block_sum(b, part);
    if (live && depart_sum > 0) {
      double value = ET(expr, P.op_expr[o]);
      FOR_CELLS {
        const int j = cell / G, u = cell - j * G;
        if (!(P.mv_g[mo * G + u] && P.mv_l1[mo * L + j])) continue;
        for (int q1 = 0; q1 < nc1; q1++) {
          int c1 = c1s[q1];
          double depart = BA(X, c1) / depart_sum * value;
          double arrive = 0;
          for (int q = 0; q < nc2; q++) arrive += BA(X, c2s[q]);
          for (int q = 0; q < nc2; q++) {
            double amt = arrive > 0 ? BA(X, c2s[q]) / arrive * depart : 1.0 / nc2 * depart;
            int sl = P.slot_of_pair[c1 * C + c2s[q]];
            float* cellp = &BA(mvD, sl);
            *cellp = (float)((double)*cellp + amt);  // CooMatrix.element_add: f32 cell (coo.py:52-57)
          }
        }
      }
      for (int q1 = 0; q1 < nc1; q1++)
        for (int q = 0; q < nc2; q++) ex_d |= 1 << P.slot_of_pair[c1s[q1] * C + c2s[q]];
    }
  }
  // 6. cost rates of today (nb_add, executor.py:236-243)
  for (int it = b.tid; it < P.I * L; it += b.nt) {
    int i = it / L, l = it % L;
    double cr = 0;
    for (int ao = 0; ao < P.n_add_op; ao++) {
      int o = P.add_op[ao];
      if (P.add_i[ao * P.I + i] && P.add_l[(size_t)ao * L + l] && op_live(P, &ET(act, 0), variant, o))
        cr += ET(expr, P.op_expr[o]) / (double)P.add_parts[ao];
    }
    b.crD[it] = cr;
  }
  __syncthreads();
  if (!dyn && rebuild) fill_tables<real>(P, b, 0.0f);  // constant over the day: filled once
  tabs_current = !dyn;
  __syncthreads();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: scipy RK45 state machine (see epi_cell.cuh: rk45_day), uniform over the CTA
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ bool rk45_day(const DevPlan& P, Bx& b, bool dyn, int ex_bits, int* trace,
                                         double* last_err_out, double* dbg, bool skip0, int& rot) {
  const double rtol = 1e-3, atol = 1e-6;
  const real rtolr = (real)1e-3, atolr = (real)1e-6;
  const int C = EDIM(C), nC = EDIM(I) * EDIM(L), nS = EDIM(has_src64) ? EDIM(n_slot) : 0, NSRC = EDIM(NSRC),
            n_acc = EDIM(n_acc);
  const double inv_n = 1.0 / (double)P.n_flat;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0, n_att = 0, guard = 0;
  double last_err = 0, t = 0.0, h_abs = 0, h = 0, t_new = 0, h0 = 0, d0 = 0, d1 = 0;
  bool run = true, rejected = false;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };
  // flow-slot mode: stage s of the current attempt lives in slot (s + rot) % 7; rot persists over the days of a launch
  // so that yesterday's last stage is found in today's slot 0 (skip0, see epi_cell.cuh: rk45_day)
  auto ps = [&](int st) { int v = st + rot; return v >= 7 ? v - 7 : v; };

  int phase = 0;
  while (phase >= 0) {
    const int stage = PH_STAGE[phase];  // K slot written
    if (phase == 2) {
      // a new attempt (rk.py:111-128)
      double min_step = 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t);  // nextafter(t, inf), t >= 0
      if (h_abs < min_step) {
        if (rejected) { status = -1; run = false; }
        else h_abs = min_step;
      }
      if (++guard > 10000 || !(h_abs == h_abs)) { status = -1; run = false; }
      if (!run) break;
      h = h_abs;
      t_new = t + h;
      if (t_new - 1.0 > 0) t_new = 1.0;
      h = t_new - t;
      h_abs = fabs(h);
    }
    // per-phase scalars and stage-input coefficients from constant tables (epi_rk.cuh)
    const int nj = PH_NJ[phase];
    const double hh = phase <= 1 ? h0 : h;
    const double t_eval = t + PH_TC[phase] * hh;
    const real bw = phBW(real(0), phase), ew = phEW(real(0), phase);
    real co[6];
    double coS[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      co[j] = rkCO(real(0), phase, j);
      coS[j] = RK_CO64[phase][j];
    }
    if (!(phase == 0 && skip0))
      rhs<real>(P, b, dyn, t_eval, ex_bits, stage, bw, ew, PH_ACC[phase], PH_KEEP[phase] != 0,
                hh, nj, co, coS, ps(stage));
    if (phase == 0) {
      // select_initial_step (common.py:68-134), first half
      double s0 = 0, s1 = 0;
      FOR_CELLS {
        real a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real yv = EXF ? BA(XF, k) : (real)BA(X, k), sc = atolr + fabs(yv) * rtolr;
          real a = fdiv(yv, sc), bq = fdiv(BA(K, k), sc);
          a0 += a * a; a1 += bq * bq;
        }
        if (EFS) {
          EACC_DECL(g0, ECOL_slot(ps(0)));
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real yv = ACCY(a_), sc = atolr + fabs(yv) * rtolr;
            real a = fdiv(yv, sc), bq = fdiv(EACC_AT(g0, a_, ECOL_slot(ps(0))), sc);
            a0 += a * a; a1 += bq * bq;
          }
        } else {
          EACC_DECL(g0, ECOL_fl);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real yv = ACCY(a_), sc = atolr + fabs(yv) * rtolr;
            real a = fdiv(yv, sc), bq = fdiv(EACC_AT(g0, a_, ECOL_fl), sc);
            a0 += a * a; a1 += bq * bq;
          }
        }
        s0 += (double)a0; s1 += (double)a1;
        if (!EFS)
          for (int i = 0; i < NSRC; i++) BA(FBE, i).y = BA(fl, i);  // stage-0 flows, needed once more in phase 1
      }
      for (int i = b.tid; i < nC; i += b.nt) {
        double yv = b.cost[i], sc = atol + fabs(yv) * rtol;
        double a = norm_term(real(0), yv, sc), bq = norm_term(real(0), b.crY[i], sc);  // q(0) = 0
        s0 += a * a; s1 += bq * bq;
      }
      d0 = ctrl_sqrt(real(0), block_sum(b, s0) * inv_n);
      d1 = ctrl_sqrt(real(0), block_sum(b, s1) * inv_n);
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, 1.0);
      nfev++;
      phase = 1;
    } else if (phase == 1) {
      double s2 = 0, q0 = qd(h0);
      FOR_CELLS {
        real a2 = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real sc = atolr + fabs(EXF ? BA(XF, k) : (real)BA(X, k)) * rtolr;
          real a = fdiv(BA(K, C + k) - BA(K, k), sc);
          a2 += a * a;
        }
        if (EFS) {
          EACC_DECL(g1, ECOL_slot(ps(1)));
          EACC_DECL(g0, ECOL_slot(ps(0)));
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real sc = atolr + fabs(ACCY(a_)) * rtolr;
            real a = fdiv(EACC_AT(g1, a_, ECOL_slot(ps(1))) - EACC_AT(g0, a_, ECOL_slot(ps(0))), sc);
            a2 += a * a;
          }
        } else {
          EACC_DECL(g1, ECOL_fl);
          EACC_DECL(g0, ECOL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real sc = atolr + fabs(ACCY(a_)) * rtolr;
            real a = fdiv(EACC_AT(g1, a_, ECOL_fl) - EACC_AT(g0, a_, ECOL_FE), sc);
            a2 += a * a;
          }
        }
        s2 += (double)a2;
        if (!EFS)
        for (int i = 0; i < NSRC; i++) {  // stage-0 contributions of the first attempt
          real f = BA(FBE, i).y;
          real2 p;
          p.x = rkB(real(0), 0) * f;
          p.y = rkE(real(0), 0) * f;
          BA(FBE, i) = p;
        }
      }
      for (int i = b.tid; i < nC; i += b.nt) {
        double sc = atol + fabs(b.cost[i]) * rtol, a = norm_term(real(0), (b.crD[i] - b.crY[i]) * q0, sc);
        s2 += a * a;
      }
      double d2 = ctrl_sqrt(real(0), block_sum(b, s2) * inv_n) / h0;
      double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : ctrl_pow(real(0), 0.01 / fmax(d1, d2), 0.2);
      h_abs = fmin(fmin(100 * h0, h1), 1.0);
      nfev++;
      phase = 2;
    } else if (phase < 7) {
      phase++;
    } else if (phase == 7) {
      // error norm over ALL n_flat components (common.py:63-65, rk.py:146-148), accept / reject
      double e2 = 0, qs[7];
      nfev += 6;
      FOR_CELLS {
        real ae = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real sc = atolr + fmax(fabs(EXF ? BA(XF, k) : (real)BA(X, k)), fabs(BA(ys, k))) * rtolr;
          // one pass over the stage derivatives gives the error estimate and the 5th-order increment; stage 1 has
          // B = E = 0 (not read), and its slot - dead until the next attempt rewrites it - keeps the increment for
          // the commit
          real e = 0, dy = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) {
            if (j == 1) continue;
            const real kj = BA(K, j * C + k);
            e += kj * rkE(real(0), j);
            if (j < 6) dy += kj * rkB(real(0), j);
          }
          BA(K, C + k) = dy;
          real a = fdiv(e * (real)h, sc);
          ae += a * a;
        }
        if (EFS) {
          // combine the 7 stored stage flows once: FB = sum_s B_s f_s, FE = sum_s E_s f_s, then the accumulator
          // gathers; the candidate y_new of the accumulators is kept for the commit
          size_t so[7];
#pragma unroll
          for (int st = 0; st < 7; st++) so[st] = (size_t)ps(st) * NSRC;
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real vB = 0, vE = 0;
#ifdef EPI_SPEC
#pragma unroll
            for (int q = spec::ag_off[a_]; q < spec::ag_off[a_ + 1]; q++) {
              const int src = spec::ag_src[q];
              real sb = 0, se = 0;
#pragma unroll
              for (int st = 0; st < 7; st++) {
                if (st == 1) continue;  // B[1] = E[1] = 0 (and the slot is not written)
                real f = BA(fls, so[st] + src);
                sb += rkB(real(0), st) * f;
                se += rkE(real(0), st) * f;
              }
              vB += sb * (real)spec::ag_coef[q];
              vE += se * (real)spec::ag_coef[q];
            }
#else
            for (int q = __ldg(P.ag_off + a_); q < __ldg(P.ag_off + a_ + 1); q++) {
              const int src = __ldg(P.ag_src + q);
              real sb = 0, se = 0;
              for (int st = 0; st < 7; st++) {
                if (st == 1) continue;  // B[1] = E[1] = 0
                real f = BA(fls, so[st] + src);
                sb += rkB(real(0), st) * f;
                se += rkE(real(0), st) * f;
              }
              vB += sb * (real)__ldg(P.ag_coef + q);
              vE += se * (real)__ldg(P.ag_coef + q);
            }
#endif
            real yv = ACCY(a_), yn = yv + (real)h * vB;
            real sc = atolr + fmax(fabs(yv), fabs(yn)) * rtolr, a = fdiv(vE * (real)h, sc);
            ae += a * a;
            ACCYN(a_) = yn;
          }
        } else {
          EACC_DECL(gB, ECOL_FB);
          EACC_DECL(gE, ECOL_FE);
#pragma unroll
          for (int a_ = 0; a_ < n_acc; a_++) {
            real yv = ACCY(a_), yn = yv + (real)h * EACC_AT(gB, a_, ECOL_FB);
            real sc = atolr + fmax(fabs(yv), fabs(yn)) * rtolr, a = fdiv(EACC_AT(gE, a_, ECOL_FE) * (real)h, sc);
            ae += a * a;
          }
        }
        e2 += (double)ae;
      }
#pragma unroll
      for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
      for (int i = b.tid; i < nC; i += b.nt) {
        double cy = b.crY[i], cg = b.crD[i] - cy, sb = 0, se = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          double k = cy + cg * qs[j];
          sb += k * RK_B[j];
          se += k * RK_E[j];
        }
        double yv = b.cost[i], yn = yv + h * sb;
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = norm_term(real(0), se * h, sc);
        e2 += a * a;
      }
      double err = ctrl_sqrt(real(0), block_sum(b, e2) * inv_n);
      last_err = err;
      if (dbg && b.tid == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
      n_att++;
      bool accepted = false;
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
      if (accepted) {
        // commit: y = y_new, f = f_new (the state is advanced in double in both modes)
        const real hr = (real)h;
        FOR_CELLS {
#pragma unroll
          for (int k = 0; k < C; k++) {
            const real dy = BA(K, C + k);  // sum_j B_j K_j, left there by the error-norm pass
            const double xn = BA(X, k) + h * (double)dy;
            BA(X, k) = xn;
            if (EXF) BA(XF, k) = (real)xn;
            BA(K, k) = BA(K, 6 * C + k);
          }
          for (int q = 0; q < nS; q++) {
            BA(X, ESLOT_C1(q)) = BA(ysS, q);  // = X + h * sum_j B_j KS_j in double
            if (EXF) BA(XF, ESLOT_C1(q)) = (real)BA(ysS, q);
            BA(KS, q) = BA(KS, 6 * nS + q);
          }
          if (EFS) {
            // (the candidate column becomes the accumulator column after this pass: see below)
          } else {
            EACC_DECL(gB, ECOL_FB);
#pragma unroll
            for (int a_ = 0; a_ < n_acc; a_++) ACCY(a_) += hr * EACC_AT(gB, a_, ECOL_FB);
          }
          if (!EFS)
          for (int i = 0; i < NSRC; i++) {  // the last stage's flows open the next attempt (FSAL)
            real f = BA(fl, i);
            real2 p;
            p.x = rkB(real(0), 0) * f;
            p.y = rkE(real(0), 0) * f;
            BA(FBE, i) = p;
          }
        }
        for (int i = b.tid; i < nC; i += b.nt) {
          double cy = b.crY[i], cg = b.crD[i] - cy, sb = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
          b.cost[i] += h * sb;
        }
        nsteps++;
        t = t_new;
        rejected = false;
        if (!(t < 1.0)) run = false;
        if (EFS) {
          rot = ps(6);  // FSAL: the slot of stage 6 becomes the slot of stage 0
          const int64_t t_ = b.o_acc;  // y = y_new for the accumulators: the columns swap roles
          b.o_acc = b.o_accn;
          b.o_accn = t_;
        }
        __syncthreads();
        phase = run ? 2 : -1;
      } else {
        // flow sums: restore the stage-0 terms by re-evaluating f(t, y) (not counted in nfev);
        // flow slots: slot 0 still holds them
        phase = EFS ? 2 : 8;
      }
    } else {
      phase = 2;
    }
  }
  if (b.tid == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
  return status == 0;
}

// ------------------------------------------------------------------------------------------
// EpiEnv.step for the envs this CTA walks (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
template <class real>
__device__ __forceinline__ void step_cta(const DevPlan& P, const DevState& S, const StepArgs& a, const Bx& b0, int cta,
                                         int n_ctas) {
  const int C = EDIM(C), CLG = EDIM(CLG), LG = EDIM(LG), nC = EDIM(I) * EDIM(L);
  const size_t nA = (size_t)EDIM(n_acc) * LG, nM = (size_t)EDIM(n_slot) * LG;
  Bx b = b0;
  b.crD = a.cr_scratch + (size_t)cta * nC;
  for (int env = cta; env < a.N; env += n_ctas) {
    const size_t e = (size_t)env;
    b.cost = S.cost + e * nC;  // the cost arrays (I*L doubles) are touched once per RK45 attempt: they stay in HBM/L2
    b.crY = S.crY + e * nC;
    double* gX = S.X + e * CLG;
    real* gAcc = (real*)S.acc + e * nA;
    double* gXprev = S.Xprev + e * P.n_chg * LG;
    int* meta = S.meta + e * 8;
    b.o_acc = P.el.l_accY;
    b.o_accn = P.el.l_accYn;
    FOR_CELLS {
      for (int k = 0; k < C; k++) {
        const double x = gX[(size_t)k * LG + cell];
        BA(X, k) = x;
        if (EXF) BA(XF, k) = (real)x;
      }
      for (int s = 0; s < EDIM(n_slot); s++) BA(mvY, s) = S.mvY[e * nM + (size_t)s * LG + cell];
      for (int a_ = 0; a_ < EDIM(n_acc); a_++) ACCY(a_) = gAcc[(size_t)a_ * LG + cell];
    }
    for (int i = b.tid; i < P.n_cls; i += b.nt) ET(mult_y, i) = S.multY[e * P.n_cls + i];
    for (int k = b.tid; k < P.NCP; k += b.nt) {
      float v;
      if (a.cpv) v = a.cpv[e * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[e * P.A + s] : P.cp_fixed[k]; }
      ET(cpv, k) = v;
    }
    for (int i = b.tid; i < P.I; i += b.nt)
      ET(act, i) = a.active ? a.active[e * P.I + i] : (a.init_only ? P.init_active[i] : 1);
    int t_day = meta[0], ex_y = meta[1];
    __syncthreads();

    double reward = 0;
    bool tabs_current = false, fsal_ok = false;
    int rot = 0;
    const int n_days = a.init_only ? 1 : a.n_days;
    for (int day = 0; day < n_days; day++) {
      if (day > 0 && a.cpv_days) {  // schedule mode: today's row
        const size_t row = (size_t)day * a.N + e;
        for (int k = b.tid; k < P.NCP; k += b.nt) ET(cpv, k) = a.cpv_days[row * P.NCP + k];
        for (int i = b.tid; i < P.I; i += b.nt) ET(act, i) = a.active_days ? a.active_days[row * P.I + i] : 1;
        __syncthreads();
      }
      bool dyn = true;
      int ex_bits = prologue<real>(P, b, a.variant, gXprev, ex_y, dyn, tabs_current);
      if (!a.init_only) {
        double c0 = 0, c1 = 0;
        for (int i = b.tid; i < nC; i += b.nt) c0 += b.cost[i];
        const bool skip0 = EPI_FSAL_DAYS && sizeof(real) == 4 && fsal_ok && !dyn;
        fsal_ok = rk45_day<real>(P, b, dyn, ex_bits, meta, S.last_err + e,
                                 a.dbg_attempts ? a.dbg_attempts + e * 128 : nullptr, skip0, rot);
        for (int i = b.tid; i < nC; i += b.nt) c1 += b.cost[i];
        const double dr = block_sum(b, c1 - c0);
        reward -= dr;  // -sum(cumulative_cost_next - cumulative_cost_current), epidemic.py:115
        t_day++;
        if (a.reward_days && b.tid == 0) a.reward_days[(size_t)day * a.N + env] = -dr;
        const size_t row = (size_t)day * a.N + e;
        if (a.obs_days)
          FOR_CELLS
            for (int k = 0; k < C; k++) ((real*)a.obs_days)[row * CLG + (size_t)k * LG + cell] = (real)BA(X, k);
        if (a.x_days)
          FOR_CELLS
            for (int k = 0; k < C; k++) a.x_days[row * CLG + (size_t)k * LG + cell] = BA(X, k);
        if (a.acc_days)
          FOR_CELLS
            for (int a_ = 0; a_ < EDIM(n_acc); a_++) ((real*)a.acc_days)[row * nA + (size_t)a_ * LG + cell] = ACCY(a_);
        if (a.mult_days)
          for (int i = b.tid; i < P.n_cls; i += b.nt) a.mult_days[row * P.n_cls + i] = ET(mult_d, i);
      }
      // today's parameters become yesterday's (State(t+1, next_obs, dest_parameter), epidemic.py:89)
      for (int i = b.tid; i < P.n_cls; i += b.nt) ET(mult_y, i) = ET(mult_d, i);
      FOR_CELLS
        for (int s = 0; s < EDIM(n_slot); s++) BA(mvY, s) = BA(mvD, s);
      for (int i = b.tid; i < nC; i += b.nt) b.crY[i] = b.crD[i];
      ex_y = (ex_bits >> 16) & 0xffff;
      __syncthreads();
      if (t_day >= P.horizon) break;  // the reference's EpiEnv.step stops its day loop at done
    }
    FOR_CELLS {
      for (int k = 0; k < C; k++) {
        double x = BA(X, k);
        gX[(size_t)k * LG + cell] = x;
        if (a.obs_out) ((real*)a.obs_out)[e * CLG + (size_t)k * LG + cell] = (real)x;
        if (a.obs_out2) ((real*)a.obs_out2)[e * CLG + (size_t)k * LG + cell] = (real)x;
      }
      for (int s = 0; s < EDIM(n_slot); s++) S.mvY[e * nM + (size_t)s * LG + cell] = BA(mvY, s);
      for (int a_ = 0; a_ < EDIM(n_acc); a_++) gAcc[(size_t)a_ * LG + cell] = ACCY(a_);
    }
    for (int i = b.tid; i < P.n_cls; i += b.nt) S.multY[e * P.n_cls + i] = ET(mult_y, i);
    if (b.tid == 0) {
      meta[0] = t_day; meta[1] = ex_y;
      if (a.reward_out) a.reward_out[env] = reward;
      if (a.done_out) a.done_out[env] = (uint8_t)(t_day >= P.horizon);
      if (a.reward_out2) a.reward_out2[env] = reward;
      if (a.done_out2) a.done_out2[env] = (uint8_t)(t_day >= P.horizon);
    }
    __syncthreads();
  }
}

// shared memory of a CTA: [env tables | reduction scratch | comm columns if they fit | big columns if they fit];
// whatever does not fit lives in the CTA's slice of the global workspace (StepArgs.small_scratch / big_scratch)
__device__ __forceinline__ Bx bind(const DevPlan& P, const StepArgs& a, char* smem, int tid, int nt, int cta) {
  Bx b;
  b.tid = tid; b.nt = nt;
  b.tab = smem;
  b.red = (double*)(smem + P.el.tab_bytes);
  char* p = smem + P.el.tab_bytes + P.el.red_bytes;
  b.comm = P.el.comm_in_smem ? p : a.small_scratch + (size_t)cta * P.el.comm_bytes;
  if (P.el.comm_in_smem) p += P.el.comm_bytes;
  b.big = P.el.big_in_smem ? p : a.big_scratch + (size_t)cta * P.el.big_bytes;
  b.cost = b.crY = b.crD = nullptr;
  return b;
}

#ifndef EPI_HOST_EMU
#ifndef EPI_ENV_MINB
#define EPI_ENV_MINB 2  // CTAs per SM the register allocation aims at (256 threads -> <= 128 registers)
#endif
template <class real>
__global__ void __launch_bounds__(256, EPI_ENV_MINB) env_kernel(const DevPlan P, const DevState S, const StepArgs a) {
  extern __shared__ __align__(16) char env_smem[];
  Bx b = bind(P, a, env_smem, threadIdx.x, blockDim.x, blockIdx.x);
  step_cta<real>(P, S, a, b, blockIdx.x, gridDim.x);
}
#endif

}  // namespace env
}  // namespace epi
