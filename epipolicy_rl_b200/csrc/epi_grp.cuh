// epi_grp.cuh — the "env-group" kernel (sm_100a): ONE environment is stepped by a GROUP of S co-resident CTAs with
// one (locale, group) cell per thread, so that the RK45 working set of a 1000-locale scenario never leaves the chip.
// Scenario-specialised builds only (EPI_SPEC, fp32 throughput mode); the env kernel (epi_env.cuh) remains the
// interpreted / fp64 path and the A/B baseline. BASELINE.json configs[3]: L=1000 x G=16 x F=8 -> 4 groups of 37 CTAs.
//
// Where the working set lives (per cell; the env kernel streamed all of it through HBM, 1.9 GB per env-step):
//   * 6 stage-derivative slots K (stage 6 reuses the slot of stage 1, whose B = E = 0; accepted steps swap the slots
//     of stages 0 and 1 instead of copying K_6 -> K_0) and the double-precision move-source derivatives: SHARED MEMORY,
//     [word][thread] columns (conflict-free);
//   * stage input y, the packed flow sums FB = sum_s B_s f_s and FE = sum_s E_s f_s (the scipy error norm's cumulative
//     components are linear in the stage flows), the cell's alive count: REGISTERS, live across the whole attempt;
//   * what crosses cells — alive / weighted-infectious columns for the CSC gather, the mixed force-of-infection terms
//     for the CSR gather — goes through an L2-resident group scratch (12 + 8 B per cell and evaluation); the mixing
//     itself only needs the locale's own groups: shared memory;
//   * touched once per RK45 attempt: the float copy of the state, the double state X, the accumulators (global, L2).
// One right-hand-side evaluation = pass 1 (stage input, alive, weighted sums) | GROUP BARRIER | pass 2 (CSC gather) |
// CTA barrier | pass 3 (G x G facility mixing, 128-bit shared-memory table reads) | GROUP BARRIER | pass 4 (CSR gather,
// edges, dX, flow sums), and pass 1 of the next stage follows without a barrier. The group barrier is an arrive counter
// in L2 (release fence + atomic, acquire spin by one thread per CTA); the kernel is launched cooperatively so that all
// CTAs of a group are resident. The controller state (t, h, accept / reject) is uniform over the group: every CTA
// reduces the per-CTA partials of a norm in the same fixed order and so takes the same decision, without a broadcast.
// Arithmetic, rounding points and the RK45 state machine are those of epi_env.cuh (flow-sum form).
#pragma once
#include "epi_kernels.cuh"
#include "epi_rk.cuh"
#ifdef EPI_GRP_TIMING
#include <stdio.h>
#endif

#ifdef EPI_SPEC
namespace epi {
namespace grp {
using namespace rk;
typedef spec::real real;

#ifndef EPI_FSAL_DAYS
#define EPI_FSAL_DAYS 1
#endif
// entries of a cell's CSC / CSR row gathered together (their loads are in flight at the same time): 8 where the
// register budget allows it (<= 12 resident warps per SM), else 4
// two locales per half-warp in the mixing pass (half the table traffic, twice the accumulators): needs the 168-register
// budget of <= 12 resident warps per SM
#ifndef EPI_GRP_MIX2
#define EPI_GRP_MIX2 ((EPI_GRP_R * (EPI_GRP_THREADS / 32)) <= 12)
#endif
#ifndef EPI_GRP_CHUNK
#define EPI_GRP_CHUNK ((EPI_GRP_R * (EPI_GRP_THREADS / 32)) <= 12 ? 8 : 4)
#endif

// -DEPI_GRP_TIMING: thread 0 of every CTA accumulates the cycles it spends in each section of the step (critical-path
// breakdown; printed by CTA 0 of group 0 when the kernel ends). Debug builds only.
#ifdef EPI_GRP_TIMING
#define TSEC(i) do { if (b.tid == 0) { long long now_ = clock64(); b.tacc[b.tcur] += now_ - b.tlast; b.tlast = now_; b.tcur = (i); } } while (0)
#else
#define TSEC(i) do { } while (0)
#endif

struct Gx {
#ifdef EPI_GRP_TIMING
  long long* tacc; long long tlast; int tcur;
#endif
  int tid, T;        // thread, threads per CTA (= width of a shared-memory column)
  int cta, S;        // CTA inside the group, CTAs per group
  int l0, nl;        // first locale / number of locales of this CTA
  int j, u, lc, cell;  // this thread's locale, group, locale class, cell index j*G+u
  bool valid;        // the thread carries a cell
  char *smem, *tab;  // shared memory; per-env tables inside it
  char *gs, *cs;     // the group's / this CTA's global scratch
  unsigned* bar;     // the group's arrive counter
  unsigned epoch;    // arrivals expected at the next barrier
  int parity;        // reduction-partials buffer in use
  int s0, s1;        // K slots of stage 0 and of stages 1 / 6
  unsigned selmask, mvmask;  // selects / MOVE ops whose locale and group masks cover this thread's cell (static)
  int nq_w;          // chunks' worth of entries of the longest CSC (low half) / CSR (high half) row in this warp
  // current env
  double *gX, *cost, *crY, *crD;
  real *acc, *accn;  // accumulator column and candidate column (swap roles on an accepted step)
  float *mvY, *mvD;
};

#define GL (P.gl)
// threads per CTA and CTAs per group are compile-time constants of the specialised build (immediate shared-memory offsets)
#define GTH (spec::grp_threads)
#define GSS (spec::grp_S)
#define GT(arr, i) (((T_##arr*)(b.tab + GL.e_##arr))[i])
#define CT(arr, i) (((T_##arr*)(b.cs + GL.c_##arr))[i])  // per-CTA tables in global memory (read on dyn days only)
#define SK(slot, k) (((real*)(b.smem + GL.s_K))[((slot) * spec::C + (k)) * GTH + b.tid])
#define SKS(slot, q) (((double*)(b.smem + GL.s_KS))[((slot) * spec::n_slot + (q)) * GTH + b.tid])
#define SCMB(k, t) (((real*)(b.smem + GL.s_cmB))[(k) * GTH + (t)])
#define GCOL(off, T_, i) (((T_*)(b.gs + (off)))[(size_t)(i) * spec::LG + b.cell])
// the float copy of the state: a shared-memory column when the layout has room for it
#define GXF(k) (*(spec::grp_xf_smem ? &((real*)(b.smem + GL.s_XF))[(k) * GTH + b.tid] : &GCOL(GL.g_XF, real, k)))

#ifdef EPI_HOST_EMU
__device__ __forceinline__ real ldcg(const real* p) { return *p; }
__device__ __forceinline__ double ldcg(const double* p) { return *p; }
__device__ __forceinline__ real ldcg_ord(const real* p) { return *p; }
#else
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
// the gathers: volatile, so that the loads of a chunk are issued in program order, all of them before the first
// multiply-add that consumes one (otherwise the register allocator serialises load -> use -> load)
__device__ __forceinline__ float ldcg_ord(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
#endif

// ---- group barrier: all S CTAs of the group; global writes before it are visible (in L2) to every CTA after it.
// Arrive = release-add on the group's counter (the release orders the CTA's writes, which thread 0 has observed through
// the CTA barrier, before the count); wait = relaxed polls by thread 0, then the CTA barrier. No acquire fence and so
// no L1 invalidation (an acquire costs CCTL.IVALL: every spill slot, table and index list would be re-fetched from L2
// after every barrier): everything another CTA writes is read with ld.global.cg, i.e. from L2, the point of coherence.
__device__ __forceinline__ void grp_arrive(Gx& b) {
  __syncthreads();
  b.epoch += (unsigned)b.S;
#ifdef EPI_HOST_EMU
  emu::group_arrive_wait();
#else
  if (b.tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(b.bar) : "memory");
#endif
}
__device__ __forceinline__ void grp_wait(Gx& b) {
#ifndef EPI_HOST_EMU
  if (b.tid == 0) {
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(b.bar) : "memory");
    } while ((int)(v - b.epoch) < 0);
  }
#endif
  __syncthreads();
}
__device__ __forceinline__ void grp_sync(Gx& b) {
  grp_arrive(b);
  grp_wait(b);
}

// butterfly over the warp: every lane ends with the same bits (each level adds the same two values, a + b == b + a)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block sum: warp butterflies, then every warp adds the warp partials the same way; every thread receives the
// bit-identical result
__device__ __forceinline__ double block_sum(const DevPlan& P, const Gx& b, double v) {
  v = warp_sum(v);
  double* red = (double*)(b.smem + GL.s_red);
  constexpr int nw = GTH >> 5;
  if ((b.tid & 31) == 0) red[b.tid >> 5] = v;
  __syncthreads();
  const int lane = b.tid & 31;
  double r = lane < nw ? red[lane] : 0.0;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// group sum of n <= 2 values: block sums -> partials in the group scratch -> barrier -> every warp of every CTA adds
// the S partials the same way (so every thread of the group holds the same bits and takes the same decision)
template <int n>
__device__ __forceinline__ void grp_sum(const DevPlan& P, Gx& b, double (&v)[n]) {
  double* part = (double*)(b.gs + GL.g_part) + (size_t)b.parity * GSS * GL.np;
#pragma unroll
  for (int i = 0; i < n; i++) v[i] = block_sum(P, b, v[i]);
  if (b.tid == 0) {
#pragma unroll
    for (int i = 0; i < n; i++) part[(size_t)b.cta * GL.np + i] = v[i];
  }
  grp_sync(b);
  const int lane = b.tid & 31;
#pragma unroll
  for (int i = 0; i < n; i++) {
    double r = 0;
#pragma unroll
    for (int c0 = 0; c0 < GSS; c0 += 32)
      if (c0 + lane < GSS) r += ldcg(part + (size_t)(c0 + lane) * GL.np + i);
    v[i] = warp_sum(r);
  }
  b.parity ^= 1;
}

// time-t tables of the env (every CTA keeps its own copy in shared memory); mixing tables in the layout
// T[lc][u0][v][u] with padded rows: a cell's thread reads its row with 128-bit loads, the two locales of a warp
// read the same addresses (broadcast) and the 16 groups of a locale hit distinct banks
__device__ __forceinline__ void fill_tables(const DevPlan& P, const Gx& b, float tf) {
  constexpr int G = spec::G, F = spec::F, NG = spec::NG, n_lc = spec::n_lc, PP = spec::P;
  const float* fi_y = (const float*)(b.cs + GL.c_fi_y);
  const float* fi_gap = (const float*)(b.cs + GL.c_fi_gap);
  for (int i = b.tid; i < n_lc * PP * G; i += b.T) GT(mpt0, i) = lerp_rn(CT(mpy, i), CT(mpg, i), tf);
  for (int i = b.tid; i < n_lc * F * G * G; i += b.T) {
    int u = i % G, u0 = (i / G) % G, lv = i / (G * G), lc = lv / F, v = lv - lc * F;
    float ft = lerp_rn(GT(ft_y, lv * G + u), GT(ft_gap, lv * G + u), tf);
    int a = (lv * G + u0) * G + u, bb = (lv * G + u) * G + u0;
    float fa = lerp_rn(fi_y[a], fi_gap[a], tf), fb = lerp_rn(fi_y[bb], fi_gap[bb], tf);
    const int it = (lc * G + u0) * GL.trow + v * G + u;
    GT(Tnum, it) = __fmul_rn(__fmul_rn(ft, fa), fb);  // f32 product (core_optimize.py:114)
    if (spec::has_ft_ops) CT(Tden, it) = (real)ft * (real)__ldg(P.fi0 + a) * (real)__ldg(P.fi0 + bb);
  }
  constexpr int FG = F * G, nFG = n_lc * FG;
  for (int i = b.tid; i < NG * nFG; i += b.T) {
    int k = i / nFG, r = i - k * nFG;  // r = (lc*F + v)*G + u0
    real cf = (real)lerp_rn(GT(ft_y, r), GT(ft_gap, r), tf);
    int pi = __ldg(P.ig_p0i + k);
    cf *= (real)lerp_rn(CT(mpFy, pi * nFG + r), CT(mpFg, pi * nFG + r), tf);
    for (int q = __ldg(P.ig_np_off + k); q < __ldg(P.ig_np_off + k + 1); q++) {
      pi = __ldg(P.ig_npi + q);
      cf *= (real)lerp_rn(CT(mpFy, pi * nFG + r), CT(mpFg, pi * nFG + r), tf);
    }
    for (int q = __ldg(P.ig_dp_off + k); q < __ldg(P.ig_dp_off + k + 1); q++) {
      pi = __ldg(P.ig_dpi + q);
      float m = lerp_rn(CT(mpFy, pi * nFG + r), CT(mpFg, pi * nFG + r), tf);
      if (fabsf(m) > 0) cf /= (real)m;
    }
    GT(coefS, i) = cf;
  }
}

// ------------------------------------------------------------------------------------------
// one right-hand-side evaluation (core_optimize.py:165-222). Stage input y = XF + hh * sum_{j<nj} co[j] K_j;
// FB += bw * f, FE += ew * f for every flow f of the evaluation; `st6`: the flows are also stored (last stage of an
// attempt: they open the next attempt, first-same-as-last).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rhs(const DevPlan& P, Gx& b, bool dyn, double t, int ex_bits, int phase, int stage, real bw,
                                    real ew, bool st6, double hh, int nj, float2 (&FBE)[spec::NSRC]) {
  constexpr int G = spec::G, F = spec::F, LG = spec::LG, NG = spec::NG, n_lc = spec::n_lc, PP = spec::P, C = spec::C,
                nE = spec::E, nS = spec::has_src64 ? spec::n_slot : 0, CH = EPI_GRP_CHUNK;
  const float tf = (float)t;
  const float qf = (float)((-3.0 * t + 4.0) * t);  // utility/utils.py:34-37, rounded like parameter.py:83
  if (dyn) {
    TSEC(13);
    __syncthreads();  // the previous evaluation's readers of the time-t tables are done
    fill_tables(P, b, tf);
    __syncthreads();
    TSEC(0);
  }
  const int kslot = stage == 0 ? b.s0 : (stage == 1 || stage == 6) ? b.s1 : stage;
  const float2 bwew = float2{bw, ew};
  real* out = (real*)(b.gs + GL.g_fl6) + b.cell;
  real d[C];
  real yc[NG > 0 ? NG : 1];                          // the susceptible compartments of the infectious groups
  real y1[spec::n_slot > 0 ? spec::n_slot : 1], y2[spec::n_slot > 0 ? spec::n_slot : 1];  // move slots: source, target
  double ysS[nS > 0 ? nS : 1];
  TSEC(1);
  // ---- pass 1: stage input, alive (N excludes death and hospitalized, :28-41,:168), weighted infectious sums
  real y[C], alive = 0;
  if (b.valid) {
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = 0;
    double dS[nS > 0 ? nS : 1];
#pragma unroll
    for (int q = 0; q < nS; q++) dS[q] = 0;
#pragma unroll 1
    for (int j = 0; j < nj; j++) {
      const int sl = j == 0 ? b.s0 : j == 1 ? b.s1 : j;
      const real cj = rkCO(real(0), phase, j);
      const real* kj = &SK(sl, 0);
#pragma unroll
      for (int k = 0; k < C; k++) y[k] += kj[k * GTH] * cj;
#pragma unroll
      for (int q = 0; q < nS; q++) dS[q] += SKS(sl, q) * RK_CO64[phase][j];
    }
#pragma unroll
    for (int k = 0; k < C; k++) y[k] = y[k] * (real)hh + GXF(k);
#pragma unroll
    for (int q = 0; q < nS; q++) ysS[q] = b.gX[(size_t)spec::slot_c1[q] * LG + b.cell] + dS[q] * hh;
    real xw[NG > 0 ? NG : 1];
    spec::phaseA(y, alive, xw);
    GCOL(GL.g_cmA, real, 0) = alive;
#pragma unroll
    for (int k = 0; k < NG; k++) GCOL(GL.g_cmA, real, 1 + k) = xw[k];
  }
  TSEC(2);
  if (NG > 0) grp_arrive(b);
  TSEC(3);
  if (b.valid) {
    // ---- the regular edges need the cell's own state only: they run while the other CTAs arrive. Their share of dX
    // waits in the slot this evaluation will fill (its old content is dead), so that the gathers and the mixing run
    // with the flow sums as the only long-lived registers
    const float* mp = &GT(mpt0, b.lc * PP * G + b.u);  // parameters from facility 0 only (:48)
    spec::flows_reg(y, alive, mp, d, FBE, bwew, out, (size_t)LG, st6);
#pragma unroll
    for (int k = 0; k < NG; k++) yc[k] = y[spec::ig_c0[k]];
#pragma unroll
    for (int sl = 0; sl < spec::n_slot; sl++) { y1[sl] = y[spec::slot_c1[sl]]; y2[sl] = y[spec::slot_c2[sl]]; }
    if (NG > 0) {
#pragma unroll
      for (int k = 0; k < C; k++) SK(kslot, k) = d[k];
    }
  }
  if (NG > 0) {
    real rs[NG];
#pragma unroll
    for (int k = 0; k < NG; k++) rs[k] = 0;
    TSEC(4);
    grp_wait(b);
    TSEC(5);
    // ---- pass 2: CSC gather over source locales (compute_foi_entry inner loop, :105-111) -> shared memory.
    // The CTA's rows are padded to a common length (ELL; padding gathers the own locale with weight 0): the source
    // locale of (entry, locale) comes from shared memory, the weight of (entry, thread) from the CTA's global table,
    // CH entries at a time with all their loads in flight together
    const int jl = b.valid ? b.j - b.l0 : 0;
    {
      real a = 0, xs[NG];
#pragma unroll
      for (int k = 0; k < NG; k++) xs[k] = 0;
      const real* cmA = (const real*)(b.gs + GL.g_cmA) + b.u;
      const int* ei = (const int*)(b.smem + GL.s_eci) + jl;
      const float* ev = (const float*)(b.cs + GL.c_evc) + b.tid;
      const int nq = b.nq_w & 0xffff;
#pragma unroll 1
      for (int q0 = 0; q0 < nq; q0 += CH) {
        real val[CH], g[CH][1 + NG];
#pragma unroll
        for (int i = 0; i < CH; i++) {
          const int src = ei[(q0 + i) * spec::grp_lpc] * G;
          val[i] = ev[(q0 + i) * GTH];
#pragma unroll
          for (int k = 0; k < 1 + NG; k++) g[i][k] = ldcg_ord(cmA + (size_t)k * LG + src);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CH; i++) {
          a += g[i][0] * val[i];
#pragma unroll
          for (int k = 0; k < NG; k++) xs[k] += g[i][1 + k] * val[i];
        }
      }
      SCMB(0, b.tid) = a;
#pragma unroll
      for (int k = 0; k < NG; k++) SCMB(1 + k, b.tid) = xs[k];
    }
    TSEC(6);
    __syncthreads();
    // ---- pass 3: G x G facility mixing -> force of infection -> s[k] for (j, u0)  (:112-124, :146-157)
    // Every cell needs the whole table row of its group (F x G values of Tnum and of Tden): read per cell, the tables
    // cross the shared-memory port once per locale and the pass is bound by it. With G == 16 a warp holds two locales
    // (lanes 0-15 / 16-31): each half-warp instead takes HALF of the facilities for BOTH locales, so every table value
    // read feeds two cells; the two halves exchange their partial sums with one shuffle per infectious group.
    if (G == 16 && F % 2 == 0 && GTH % 32 == 0 && EPI_GRP_MIX2) {
      constexpr int FH = F / 2;
      const int lane = b.tid & 31, half = lane >> 4, u0 = lane & 15;
      const int rowA = (b.tid & ~31), rowB = rowA + 16;           // first column of the warp's two locales
      // (tables are per locale class; both locales of a warp almost always share it, the general case reads both rows)
      const int lca = __shfl_sync(0xffffffffu, b.lc, 0), lcb = __shfl_sync(0xffffffffu, b.lc, 16);
      const float* TnA = &GT(Tnum, (lca * G + u0) * GL.trow) + half * FH * G;
      const float* TnB = &GT(Tnum, (lcb * G + u0) * GL.trow) + half * FH * G;
      const real* TdBase = (spec::has_ft_ops ? (const real*)(b.cs + GL.c_Tden)
                            : spec::grp_tden_smem ? (const real*)(b.tab + GL.e_Tden) : (const real*)P.Tden_constP);
      const real* TdA = TdBase + (lca * G + u0) * GL.trow + half * FH * G;
      const real* TdB = TdBase + (lcb * G + u0) * GL.trow + half * FH * G;
      const bool same = lca == lcb;
      float2 denA[FH], denB[FH], numA[NG][FH], numB[NG][FH];
#pragma unroll
      for (int v = 0; v < FH; v++) {
        denA[v] = denB[v] = float2{0.f, 0.f};
#pragma unroll
        for (int k = 0; k < NG; k++) numA[k][v] = numB[k][v] = float2{0.f, 0.f};
      }
#pragma unroll 1
      for (int u4 = 0; u4 < G / 4; u4++) {
        float4 cA[1 + NG], cB[1 + NG];
#pragma unroll
        for (int k = 0; k < 1 + NG; k++) {
          cA[k] = *(const float4*)&SCMB(k, rowA + 4 * u4);
          cB[k] = *(const float4*)&SCMB(k, rowB + 4 * u4);
        }
#pragma unroll
        for (int v = 0; v < FH; v++) {
          const float4 tn = *(const float4*)(TnA + v * G + 4 * u4);
          const float4 td = *(const float4*)(TdA + v * G + 4 * u4);
          const float4 tnb = same ? tn : *(const float4*)(TnB + v * G + 4 * u4);
          const float4 tdb = same ? td : *(const float4*)(TdB + v * G + 4 * u4);
          denA[v] = ffma2(float2{cA[0].x, cA[0].y}, float2{td.x, td.y}, denA[v]);
          denA[v] = ffma2(float2{cA[0].z, cA[0].w}, float2{td.z, td.w}, denA[v]);
          denB[v] = ffma2(float2{cB[0].x, cB[0].y}, float2{tdb.x, tdb.y}, denB[v]);
          denB[v] = ffma2(float2{cB[0].z, cB[0].w}, float2{tdb.z, tdb.w}, denB[v]);
#pragma unroll
          for (int k = 0; k < NG; k++) {
            numA[k][v] = ffma2(float2{cA[1 + k].x, cA[1 + k].y}, float2{tn.x, tn.y}, numA[k][v]);
            numA[k][v] = ffma2(float2{cA[1 + k].z, cA[1 + k].w}, float2{tn.z, tn.w}, numA[k][v]);
            numB[k][v] = ffma2(float2{cB[1 + k].x, cB[1 + k].y}, float2{tnb.x, tnb.y}, numB[k][v]);
            numB[k][v] = ffma2(float2{cB[1 + k].z, cB[1 + k].w}, float2{tnb.z, tnb.w}, numB[k][v]);
          }
        }
      }
      real skA[NG], skB[NG];
#pragma unroll
      for (int k = 0; k < NG; k++) skA[k] = skB[k] = 0;
#pragma unroll
      for (int v = 0; v < FH; v++) {
        real dA = denA[v].x + denA[v].y, dB = denB[v].x + denB[v].y;
        if (dA <= (real)1e-15) dA = 1;
        if (dB <= (real)1e-15) dB = 1;
        const real rA = frcp(dA), rB = frcp(dB);
#pragma unroll
        for (int k = 0; k < NG; k++) {
          skA[k] += GT(coefS, ((k * n_lc + lca) * F + half * FH + v) * G + u0) * ((numA[k][v].x + numA[k][v].y) * rA);
          skB[k] += GT(coefS, ((k * n_lc + lcb) * F + half * FH + v) * G + u0) * ((numB[k][v].x + numB[k][v].y) * rB);
        }
      }
      // lanes 0-15 own locale A's cells, lanes 16-31 locale B's: each sends the other half's partial and keeps its own
#pragma unroll
      for (int k = 0; k < NG; k++) {
        const real mine = half ? skB[k] : skA[k], theirs = half ? skA[k] : skB[k];
        const real got = __shfl_xor_sync(0xffffffffu, theirs, 16);
        // facilities in a fixed order: the lower half first
        const real sk = half ? got + mine : mine + got;
        if (b.valid) GCOL(GL.g_cmS, real, k) = sk;
      }
    } else if (b.valid) {
      real sk[NG];
#pragma unroll
      for (int k = 0; k < NG; k++) sk[k] = 0;
      const int row = b.tid - b.u;  // the locale's first column
      const float* Tn = &GT(Tnum, (b.lc * G + b.u) * GL.trow);
      // the denominator's table: this CTA's time-t table with FT ops (global scratch); else env-independent: the
      // shared-memory copy when the layout has one, or the padded copy in global memory (L1 / L2 hits)
      const real* Td = (spec::has_ft_ops ? (const real*)(b.cs + GL.c_Tden)
                        : spec::grp_tden_smem ? (const real*)(b.tab + GL.e_Tden) : (const real*)P.Tden_constP) +
                       (b.lc * G + b.u) * GL.trow;
      constexpr int VH = (F % 4 == 0 && G % 4 == 0) ? 4 : 1;  // facilities per sweep over the locale's groups
#pragma unroll 1
      for (int v0 = 0; v0 < F; v0 += VH) {
        real den[VH], num[NG][VH];
        if (G % 4 == 0) {
          // four groups of the locale at a time: their gathered columns in registers (128-bit broadcast reads),
          // against VH rows of the cell's tables (128-bit reads, conflict-free over the locale's groups); packed
          // multiply-adds (FFMA2): even and odd groups accumulate in the two halves of a register pair
          float2 den2[VH], num2[NG][VH];
#pragma unroll
          for (int v = 0; v < VH; v++) {
            den2[v] = float2{0.f, 0.f};
#pragma unroll
            for (int k = 0; k < NG; k++) num2[k][v] = float2{0.f, 0.f};
          }
#pragma unroll 1
          for (int u4 = 0; u4 < G / 4; u4++) {
            float4 c4[1 + NG];
#pragma unroll
            for (int k = 0; k < 1 + NG; k++) c4[k] = *(const float4*)&SCMB(k, row + 4 * u4);
#pragma unroll
            for (int v = 0; v < VH; v++) {
              const float4 tn = *(const float4*)(Tn + (v0 + v) * G + 4 * u4);
              const float4 td = *(const float4*)(Td + (v0 + v) * G + 4 * u4);
              den2[v] = ffma2(float2{c4[0].x, c4[0].y}, float2{td.x, td.y}, den2[v]);
              den2[v] = ffma2(float2{c4[0].z, c4[0].w}, float2{td.z, td.w}, den2[v]);
#pragma unroll
              for (int k = 0; k < NG; k++) {
                num2[k][v] = ffma2(float2{c4[1 + k].x, c4[1 + k].y}, float2{tn.x, tn.y}, num2[k][v]);
                num2[k][v] = ffma2(float2{c4[1 + k].z, c4[1 + k].w}, float2{tn.z, tn.w}, num2[k][v]);
              }
            }
          }
#pragma unroll
          for (int v = 0; v < VH; v++) {
            den[v] = den2[v].x + den2[v].y;
#pragma unroll
            for (int k = 0; k < NG; k++) num[k][v] = num2[k][v].x + num2[k][v].y;
          }
        } else {
#pragma unroll
          for (int v = 0; v < VH; v++) {
            den[v] = 0;
#pragma unroll
            for (int k = 0; k < NG; k++) num[k][v] = 0;
          }
#pragma unroll 1
          for (int u = 0; u < G; u++) {
            real c1[1 + NG];
#pragma unroll
            for (int k = 0; k < 1 + NG; k++) c1[k] = SCMB(k, row + u);
#pragma unroll
            for (int v = 0; v < VH; v++) {
              den[v] += c1[0] * Td[(v0 + v) * G + u];
#pragma unroll
              for (int k = 0; k < NG; k++) num[k][v] += c1[1 + k] * (real)Tn[(v0 + v) * G + u];
            }
          }
        }
#pragma unroll
        for (int v = 0; v < VH; v++) {
          real dn = den[v];
          if (dn <= (real)1e-15) dn = 1;
          const real rden = frcp(dn);
#pragma unroll
          for (int k = 0; k < NG; k++)
            sk[k] += GT(coefS, ((k * n_lc + b.lc) * F + v0 + v) * G + b.u) * (num[k][v] * rden);
        }
      }
#pragma unroll
      for (int k = 0; k < NG; k++) GCOL(GL.g_cmS, real, k) = sk[k];
    }
    TSEC(7);
    grp_sync(b);
    TSEC(8);
    // ---- pass 4a: CSR gather over destination locales (:144-158), then the infectious flows
    {
      const real* cmS = (const real*)(b.gs + GL.g_cmS) + b.u;
      const int* ei = (const int*)(b.smem + GL.s_eri) + jl;
      const float* ev = (const float*)(b.cs + GL.c_evr) + b.tid;
      const int nq = b.nq_w >> 16;
#pragma unroll 1
      for (int q0 = 0; q0 < nq; q0 += CH) {
        real val[CH], g[CH][NG];
#pragma unroll
        for (int i = 0; i < CH; i++) {
          const int dst = ei[(q0 + i) * spec::grp_lpc] * G;
          val[i] = ev[(q0 + i) * GTH];
#pragma unroll
          for (int k = 0; k < NG; k++) g[i][k] = ldcg_ord(cmS + (size_t)k * LG + dst);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CH; i++)
#pragma unroll
          for (int k = 0; k < NG; k++) rs[k] += g[i][k] * val[i];
      }
    }
    if (b.valid) {
#pragma unroll
      for (int k = 0; k < C; k++) d[k] = SK(kslot, k);
      spec::flows_inf(yc, rs, d, FBE, bwew, out, (size_t)LG, st6);
    }
  }
  TSEC(9);
  // ---- pass 4b: clamped moves, the stage derivative
  if (b.valid) {
#pragma unroll
    for (int sl = 0; sl < spec::n_slot; sl++) {
      real nv = 0;
      const int c1 = spec::slot_c1[sl], c2 = spec::slot_c2[sl];
      if (__ldg(P.slot_cover + (size_t)sl * LG + b.cell) && ((ex_bits >> sl) & 0x10001)) {
        float my = ((ex_bits >> sl) & 1) ? b.mvY[(size_t)sl * LG + b.cell] : 0.0f;
        float md = ((ex_bits >> (16 + sl)) & 1) ? b.mvD[(size_t)sl * LG + b.cell] : 0.0f;
        float v = __fadd_rn(__fmul_rn(__fsub_rn(md, my), qf), my);  // coo.py scale/add in f32
        double ya = spec::has_src64 ? ysS[sl] : (double)y1[sl], yb = (double)y2[sl];
        double n1 = ya + (double)d[c1], n2 = yb + (double)d[c2];
        double nvd = (double)v < n1 ? (double)v : n1;
        double k1 = (n1 - nvd) - ya;
        d[c1] = (real)k1;
        d[c2] = (real)((n2 + nvd) - yb);
        nv = (real)nvd;
        if (spec::has_src64) SKS(kslot, sl) = k1;
      } else if (spec::has_src64) {
        SKS(kslot, sl) = (double)d[c1];
      }
      FBE[nE + NG + sl] = ffma2(float2{nv, nv}, bwew, FBE[nE + NG + sl]);
      if (st6) out[(size_t)(nE + NG + sl) * LG] = nv;
    }
#pragma unroll
    for (int k = 0; k < C; k++) SK(kslot, k) = d[k];
  }
  TSEC(0);
}

// ------------------------------------------------------------------------------------------
// day prologue (Executor.execute + get_parameter_from_delta + get_gap_parameter). Select sums and the departing
// populations of the MOVE ops need the whole env: one group reduction for all of them; the rest is per cell or a
// per-CTA copy of the env's small tables.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int prologue(const DevPlan& P, Gx& b, int variant, double* Xprev_g, int ex_y_bits, bool& dyn,
                                        bool& tabs_current) {
  constexpr int LG = spec::LG, G = spec::G, L = spec::L, C = spec::C, F = spec::F, PP = spec::P;
  double x[C];
#pragma unroll
  for (int k = 0; k < C; k++) x[k] = b.valid ? b.gX[(size_t)k * LG + b.cell] : 0.0;
  // 1. select sums (executor.py:346-359, ['Value'].sum()) and departing populations (executor.py:181-183)
  double* part = (double*)(b.gs + GL.g_part) + (size_t)b.parity * GSS * GL.np;
  const int2* selinfo = (const int2*)(b.tab + GL.e_selinfo);  // (compartment mask, mode) per select, cached per kernel
  for (int s = 0; s < P.n_sel; s++) {
    double v = 0;
    if ((b.selmask >> s) & 1u) {
      const int2 si = selinfo[s];
#pragma unroll
      for (int k = 0; k < C; k++)
        if ((si.x >> k) & 1) {
          if (si.y == 0) v += x[k];
          else if (si.y == 1) v += P.X0[(size_t)k * LG + b.cell];
          else v += x[k] - Xprev_g[(size_t)P.chg_slot_of_comp[k] * LG + b.cell];
        }
    }
    v = block_sum(P, b, v);
    if (b.tid == 0) part[(size_t)b.cta * GL.np + s] = v;
  }
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    double v = 0;
    if ((b.mvmask >> mo) & 1u)
      for (int q = P.mv_c1_off[mo]; q < P.mv_c1_off[mo + 1]; q++) {
        const int c1 = P.mv_c1[q];
#pragma unroll
        for (int k = 0; k < C; k++)
          if (k == c1) v += x[k];
      }
    v = block_sum(P, b, v);
    if (b.tid == 0) part[(size_t)b.cta * GL.np + P.n_sel + mo] = v;
  }
  grp_sync(b);
  double* dep = (double*)(b.smem + GL.s_red) + 32;  // departing populations, [n_mv_op]
  // one warp per total: the lanes fetch the S partials together (one L2 round trip instead of S), the butterfly adds
  // them the same way in every CTA
  for (int i0 = 0; i0 < P.n_sel + P.n_mv_op; i0 += GTH >> 5) {
    const int i = i0 + (b.tid >> 5), lane = b.tid & 31;
    const bool on = i < P.n_sel + P.n_mv_op;
    double r = 0;
#pragma unroll
    for (int c0 = 0; c0 < GSS; c0 += 32)
      if (on && c0 + lane < GSS) r += ldcg(part + (size_t)(c0 + lane) * GL.np + i);
    r = warp_sum(r);
    if (on && lane == 0) {
      if (i < P.n_sel) GT(sel, i) = r;
      else dep[i - P.n_sel] = r;
    }
  }
  b.parity ^= 1;
  __syncthreads();
  // the previous state for tomorrow's 'change' selects is today's start state (epidemic.py:96-99)
  if (b.valid) {
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int sl = P.chg_slot_of_comp[k];
      if (sl >= 0) Xprev_g[(size_t)sl * LG + b.cell] = x[k];
    }
  }
  // 2. expressions
  for (int e = b.tid; e < P.n_expr; e += b.T) {
    int itv = P.expr_itv[e];
    bool live = GT(act, itv) && P.expr_prog[e] == (variant ? P.prog_init[itv] : P.prog_step[itv]);
    GT(expr, e) = live ? eval_expr(P, e, &GT(cpv, 0), &GT(sel, 0)) : 0.0;
  }
  __syncthreads();
  // 3. class multipliers: sequential f32 products in op order (nb_apply, executor.py:116-164)
  for (int k = b.tid; k < P.n_cls; k += b.T) {
    float m = 1.0f;
    for (int q = P.cls_off[k]; q < P.cls_off[k + 1]; q++) {
      int o = P.cls_op[q];
      if (op_live(P, &GT(act, 0), variant, o)) m = __fmul_rn(m, (float)GT(expr, P.op_expr[o]));
    }
    GT(mult_d, k) = m;
  }
  __syncthreads();
  {
    int diff = 0;
    for (int k = b.tid; k < P.n_cls; k += b.T) diff |= GT(mult_d, k) != GT(mult_y, k);
    dyn = __syncthreads_or(diff) != 0;
  }
  const bool rebuild = dyn || !tabs_current;
  if (rebuild) {
    // 4. facility tables of yesterday and today -> (y, gap)   (parameter.py:64-65,:73-74)
    float* fi_y = (float*)(b.cs + GL.c_fi_y);
    float* fi_gap = (float*)(b.cs + GL.c_fi_gap);
    for (int row = b.tid; row < P.n_lc * F * G; row += b.T) {
      float a[32], bb[32];  // G <= 32 (checked on the host)
      facility_row_fi(P, &GT(mult_y, 0), row, a);
      facility_row_fi(P, &GT(mult_d, 0), row, bb);
      for (int g2 = 0; g2 < G; g2++) {
        fi_y[row * G + g2] = a[g2];
        fi_gap[row * G + g2] = __fsub_rn(bb[g2], a[g2]);
      }
    }
    for (int it = b.tid; it < P.n_lc * G; it += b.T) {
      int lc = it / G, g = it % G;
      float a[32], bb[32];  // F <= 32 (checked on the host)
      float s = 0.0f, s2 = 0.0f;
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        a[f] = __fmul_rn(P.ft0[i], GT(mult_y, P.ft_cls[i]));
        bb[f] = __fmul_rn(P.ft0[i], GT(mult_d, P.ft_cls[i]));
        s = __fadd_rn(s, a[f]);
        s2 = __fadd_rn(s2, bb[f]);
      }
      for (int f = 0; f < F; f++) {
        int i = (lc * F + f) * G + g;
        float ya = s > 1e-15f ? __fdiv_rn(a[f], s) : (f == 0 ? 1.0f : 0.0f);
        float yb = s2 > 1e-15f ? __fdiv_rn(bb[f], s2) : (f == 0 ? 1.0f : 0.0f);
        GT(ft_y, i) = ya;
        GT(ft_gap, i) = __fsub_rn(yb, ya);
      }
    }
    for (int i = b.tid; i < P.n_lc * PP * G; i += b.T) {
      int g = i % G, p = (i / G) % PP, lc = i / (G * PP);
      int idx = ((lc * PP + p) * F + 0) * G + g;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float yv = __fmul_rn(m0, GT(mult_y, cls)), dv = __fmul_rn(m0, GT(mult_d, cls));
      CT(mpy, i) = yv;
      CT(mpg, i) = __fsub_rn(dv, yv);
    }
    const int nFG = P.n_lc * F * G;
    for (int i = b.tid; i < P.n_igp * nFG; i += b.T) {
      int pi = i / nFG, r = i - pi * nFG, lc = r / (F * G), fg = r - lc * F * G;
      int idx = (lc * PP + __ldg(P.igp + pi)) * F * G + fg;
      float m0 = __ldg(P.mp0 + idx);
      int cls = __ldg(P.mp_cls + idx);
      float yv = __fmul_rn(m0, GT(mult_y, cls)), dv = __fmul_rn(m0, GT(mult_d, cls));
      CT(mpFy, i) = yv;
      CT(mpFg, i) = __fsub_rn(dv, yv);
    }
  }
  // 5. move table of today (nb_move, executor.py:166-234; different compartment, same locale)
  if (b.valid)
    for (int sl = 0; sl < spec::n_slot; sl++) b.mvD[(size_t)sl * LG + b.cell] = 0.0f;
  int ex_d = 0;
  for (int mo = 0; mo < P.n_mv_op; mo++) {
    int o = P.mv_op[mo];
    bool live = op_live(P, &GT(act, 0), variant, o);
    int nc1 = P.mv_c1_off[mo + 1] - P.mv_c1_off[mo], nc2 = P.mv_c2_off[mo + 1] - P.mv_c2_off[mo];
    const int* c1s = P.mv_c1 + P.mv_c1_off[mo];
    const int* c2s = P.mv_c2 + P.mv_c2_off[mo];
    const double depart_sum = dep[mo];
    if (live && depart_sum > 0) {
      double value = GT(expr, P.op_expr[o]);
      if ((b.mvmask >> mo) & 1u) {
        auto xv = [&](int c) {
          double r = 0;
#pragma unroll
          for (int k = 0; k < C; k++)
            if (k == c) r = x[k];
          return r;
        };
        for (int q1 = 0; q1 < nc1; q1++) {
          int c1 = c1s[q1];
          double depart = xv(c1) / depart_sum * value;
          double arrive = 0;
          for (int q = 0; q < nc2; q++) arrive += xv(c2s[q]);
          for (int q = 0; q < nc2; q++) {
            double amt = arrive > 0 ? xv(c2s[q]) / arrive * depart : 1.0 / nc2 * depart;
            int sl = P.slot_of_pair[c1 * C + c2s[q]];
            float* cellp = &b.mvD[(size_t)sl * LG + b.cell];
            *cellp = (float)((double)*cellp + amt);  // CooMatrix.element_add: f32 cell (coo.py:52-57)
          }
        }
      }
      for (int q1 = 0; q1 < nc1; q1++)
        for (int q = 0; q < nc2; q++) ex_d |= 1 << P.slot_of_pair[c1s[q1] * C + c2s[q]];
    }
  }
  // 6. cost rates of today (nb_add, executor.py:236-243) for the (intervention, locale) entries of this CTA
  for (int q = b.tid; q < P.I * b.nl; q += b.T) {
    int i = q / b.nl, l = b.l0 + q % b.nl;
    double cr = 0;
    for (int ao = 0; ao < P.n_add_op; ao++) {
      int o = P.add_op[ao];
      if (P.add_i[ao * P.I + i] && P.add_l[(size_t)ao * L + l] && op_live(P, &GT(act, 0), variant, o))
        cr += GT(expr, P.op_expr[o]) / (double)P.add_parts[ao];
    }
    b.crD[(size_t)i * L + l] = cr;
  }
  __syncthreads();
  if (!dyn && rebuild) fill_tables(P, b, 0.0f);  // constant over the day: filled once
  tabs_current = !dyn;
  __syncthreads();
  return (ex_y_bits & 0xffff) | (ex_d << 16);
}

// ------------------------------------------------------------------------------------------
// one simulated day: scipy RK45 state machine (rk.py:86-176, common.py:63-134), uniform over the group
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool rk45_day(const DevPlan& P, Gx& b, bool dyn, int ex_bits, int* trace, double* last_err_out,
                                         double* dbg, bool skip0) {
  const double rtol = 1e-3, atol = 1e-6;
  const real rtolr = (real)1e-3, atolr = (real)1e-6;
  constexpr int C = spec::C, L = spec::L, LG = spec::LG, nS = spec::has_src64 ? spec::n_slot : 0, NSRC = spec::NSRC,
                n_acc = spec::n_acc;
  const int nCl = spec::I * b.nl;  // cost entries of this CTA
  const double inv_n = 1.0 / (double)P.n_flat;
  int nfev = 0, nsteps = 0, nrej = 0, status = 0, n_att = 0, guard = 0;
  double last_err = 0, t = 0.0, h_abs = 0, h = 0, t_new = 0, h0 = 0, d0 = 0, d1 = 0;
  bool run = true, rejected = false;
  auto qd = [](double tt) { return (double)(float)((-3.0 * tt + 4.0) * tt); };
  auto centry = [&](int q) { return (size_t)(q / b.nl) * L + b.l0 + q % b.nl; };
  // the packed flow sums of the current attempt live in registers for the whole day
  float2 FBE[NSRC];  // (.x, .y) = (FB, FE): one packed multiply-add (FFMA2) updates both
#pragma unroll
  for (int i = 0; i < NSRC; i++) FBE[i] = float2{0.f, 0.f};
  const real* fl6 = (const real*)(b.gs + GL.g_fl6) + b.cell;

  int phase = 0;
  while (phase >= 0) {
    const int stage = PH_STAGE[phase];
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
    const int nj = PH_NJ[phase];
    const double hh = phase <= 1 ? h0 : h;
    const double t_eval = t + PH_TC[phase] * hh;
    real bw = phBW(real(0), phase), ew = phEW(real(0), phase);
    // the flow sums are accumulated (FB += bw f, FE += ew f); the phases that SET them clear first:
    // phase 0: FE = f(0, y0); phase 1: FB = f(h0, y0 + h0 f0); phase 8: FB = B_0 f, FE = E_0 f
    if (phase == 0) {
      bw = 0; ew = 1;
#pragma unroll
      for (int i = 0; i < NSRC; i++) FBE[i].y = 0;
    } else if (phase == 1) {
      bw = 1; ew = 0;
#pragma unroll
      for (int i = 0; i < NSRC; i++) FBE[i].x = 0;
    } else if (phase == 8) {
#pragma unroll
      for (int i = 0; i < NSRC; i++) FBE[i] = float2{0.f, 0.f};
    }
    if (phase == 0 && skip0) {
      // f(0, y0) of this day is the previous day's last stage (same state, same parameters): its derivatives sit in
      // slot s0 already, its flows in the last-stage column
      if (b.valid) {
#pragma unroll
        for (int i = 0; i < NSRC; i++) FBE[i].y = fl6[(size_t)i * LG];
      }
    } else {
      rhs(P, b, dyn, t_eval, ex_bits, phase, stage, bw, ew, phase == 7, hh, nj, FBE);
    }
    if (phase == 0) {
      TSEC(11);
      // select_initial_step (common.py:68-134), first half
      double s01[2] = {0, 0};
      if (b.valid) {
        real a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real yv = GXF(k), sc = atolr + fabs(yv) * rtolr;
          real a = fdiv(yv, sc), bq = fdiv(SK(b.s0, k), sc);
          a0 += a * a; a1 += bq * bq;
        }
#pragma unroll
        for (int a_ = 0; a_ < n_acc; a_++) {
          real g0 = 0;
#pragma unroll
          for (int q = spec::ag_off[a_]; q < spec::ag_off[a_ + 1]; q++) g0 += FBE[spec::ag_src[q]].y * (real)spec::ag_coef[q];
          real yv = b.acc[(size_t)a_ * LG + b.cell], sc = atolr + fabs(yv) * rtolr;
          real a = fdiv(yv, sc), bq = fdiv(g0, sc);
          a0 += a * a; a1 += bq * bq;
        }
        s01[0] = (double)a0; s01[1] = (double)a1;
      }
      for (int q = b.tid; q < nCl; q += b.T) {
        const size_t i = centry(q);
        double yv = b.cost[i], sc = atol + fabs(yv) * rtol;
        double a = norm_term(real(0), yv, sc), bq = norm_term(real(0), b.crY[i], sc);  // q(0) = 0
        s01[0] += a * a; s01[1] += bq * bq;
      }
      TSEC(10);
      grp_sum<2>(P, b, s01);
      TSEC(0);
      d0 = ctrl_sqrt(real(0), s01[0] * inv_n);
      d1 = ctrl_sqrt(real(0), s01[1] * inv_n);
      h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, 1.0);
      nfev++;
      phase = 1;
    } else if (phase == 1) {
      TSEC(11);
      double s2[1] = {0};
      const double q0 = qd(h0);
      if (b.valid) {
        real a2 = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real sc = atolr + fabs(GXF(k)) * rtolr;
          real a = fdiv(SK(b.s1, k) - SK(b.s0, k), sc);
          a2 += a * a;
        }
#pragma unroll
        for (int a_ = 0; a_ < n_acc; a_++) {
          real g = 0;
#pragma unroll
          for (int q = spec::ag_off[a_]; q < spec::ag_off[a_ + 1]; q++)
            g += (FBE[spec::ag_src[q]].x - FBE[spec::ag_src[q]].y) * (real)spec::ag_coef[q];
          real sc = atolr + fabs(b.acc[(size_t)a_ * LG + b.cell]) * rtolr;
          real a = fdiv(g, sc);
          a2 += a * a;
        }
        s2[0] = (double)a2;
      }
      // stage-0 contributions of the first attempt
#pragma unroll
      for (int i = 0; i < NSRC; i++) {
        const real f = FBE[i].y;
        FBE[i] = float2{rkB(real(0), 0) * f, rkE(real(0), 0) * f};
      }
      for (int q = b.tid; q < nCl; q += b.T) {
        const size_t i = centry(q);
        double sc = atol + fabs(b.cost[i]) * rtol, a = norm_term(real(0), (b.crD[i] - b.crY[i]) * q0, sc);
        s2[0] += a * a;
      }
      TSEC(10);
      grp_sum<1>(P, b, s2);
      TSEC(0);
      double d2 = ctrl_sqrt(real(0), s2[0] * inv_n) / h0;
      double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : ctrl_pow(real(0), 0.01 / fmax(d1, d2), 0.2);
      h_abs = fmin(fmin(100 * h0, h1), 1.0);
      nfev++;
      phase = 2;
    } else if (phase < 7) {
      phase++;
    } else if (phase == 7) {
      TSEC(11);
      // error norm over ALL n_flat components (common.py:63-65, rk.py:146-148), accept / reject
      double e2[1] = {0}, qs[7];
      nfev += 6;
      real dy[C];
      if (b.valid) {
        real ae = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          real e = 0, inc = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) {
            if (j == 1) continue;  // B_1 = E_1 = 0 (and slot s1 holds stage 6 by now)
            const real kj = SK(j == 0 ? b.s0 : j == 6 ? b.s1 : j, k);
            e += kj * rkE(real(0), j);
            if (j < 6) inc += kj * rkB(real(0), j);
          }
          dy[k] = inc;
          // the 5th-order candidate y_new (= the input stage 6 was evaluated at)
          const real xf = GXF(k), yn = inc * (real)h + xf;
          real sc = atolr + fmax(fabs(xf), fabs(yn)) * rtolr;
          real a = fdiv(e * (real)h, sc);
          ae += a * a;
        }
#pragma unroll
        for (int a_ = 0; a_ < n_acc; a_++) {
          real vB = 0, vE = 0;
#pragma unroll
          for (int q = spec::ag_off[a_]; q < spec::ag_off[a_ + 1]; q++) {
            vB += FBE[spec::ag_src[q]].x * (real)spec::ag_coef[q];
            vE += FBE[spec::ag_src[q]].y * (real)spec::ag_coef[q];
          }
          real yv = b.acc[(size_t)a_ * LG + b.cell], yn = yv + (real)h * vB;
          real sc = atolr + fmax(fabs(yv), fabs(yn)) * rtolr, a = fdiv(vE * (real)h, sc);
          ae += a * a;
          b.accn[(size_t)a_ * LG + b.cell] = yn;
        }
        e2[0] = (double)ae;
      }
#pragma unroll
      for (int j = 0; j < 7; j++) qs[j] = qd(t + RK_C[j] * h);  // cost rate is a closed form in t
      for (int q = b.tid; q < nCl; q += b.T) {
        const size_t i = centry(q);
        double cy = b.crY[i], cg = b.crD[i] - cy, sb = 0, se = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
          double k = cy + cg * qs[j];
          sb += k * RK_B[j];
          se += k * RK_E[j];
        }
        double yv = b.cost[i], yn = yv + h * sb;
        double sc = atol + fmax(fabs(yv), fabs(yn)) * rtol, a = norm_term(real(0), se * h, sc);
        e2[0] += a * a;
      }
      TSEC(10);
      grp_sum<1>(P, b, e2);
      TSEC(0);
      double err = ctrl_sqrt(real(0), e2[0] * inv_n);
      last_err = err;
      if (dbg && b.tid == 0 && b.cta == 0 && n_att < 64) { dbg[2 * n_att] = h; dbg[2 * n_att + 1] = err; }
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
        TSEC(12);
        // commit: y = y_new, f = f_new (the state is advanced in double in both modes)
        if (b.valid) {
          double xs[nS > 0 ? nS : 1];
#pragma unroll
          for (int q = 0; q < nS; q++) {
            // the move slots' source compartments in double: X + h * sum_j B_j KS_j
            double inc = 0;
#pragma unroll
            for (int j = 0; j < 6; j++) {
              if (j == 1) continue;
              inc += SKS(j == 0 ? b.s0 : j, q) * RK_B[j];
            }
            xs[q] = b.gX[(size_t)spec::slot_c1[q] * LG + b.cell] + inc * h;
          }
#pragma unroll
          for (int k = 0; k < C; k++) {
            double xn = b.gX[(size_t)k * LG + b.cell] + h * (double)dy[k];
#pragma unroll
            for (int q = 0; q < nS; q++)
              if (k == spec::slot_c1[q]) xn = xs[q];
            b.gX[(size_t)k * LG + b.cell] = xn;
            GXF(k) = (real)xn;
          }
          // the last stage's flows open the next attempt (first same as last)
#pragma unroll
          for (int i = 0; i < NSRC; i++) {
            const real f = fl6[(size_t)i * LG];
            FBE[i] = float2{rkB(real(0), 0) * f, rkE(real(0), 0) * f};
          }
        }
        for (int q = b.tid; q < nCl; q += b.T) {
          const size_t i = centry(q);
          double cy = b.crY[i], cg = b.crD[i] - cy, sb = 0;
#pragma unroll
          for (int j = 0; j < 7; j++) sb += (cy + cg * qs[j]) * RK_B[j];
          b.cost[i] += h * sb;
        }
        nsteps++;
        t = t_new;
        rejected = false;
        if (!(t < 1.0)) run = false;
        {  // stage 6 becomes stage 0 of the next attempt; the accumulators' candidate column becomes the column
          const int s_ = b.s0; b.s0 = b.s1; b.s1 = s_;
          real* t_ = b.acc; b.acc = b.accn; b.accn = t_;
        }
        TSEC(0);
        phase = run ? 2 : -1;
      } else {
        phase = 8;  // restore the stage-0 terms of the flow sums by re-evaluating f(t, y) (not counted in nfev)
      }
    } else {
      phase = 2;
    }
  }
  if (b.tid == 0 && b.cta == 0) {
    trace[3] = nfev; trace[4] = nsteps; trace[5] = nrej; trace[2] = status;
    *last_err_out = last_err;
  }
  return status == 0;
}

// ------------------------------------------------------------------------------------------
// EpiEnv.step for the envs this group walks (epipolicy_environment.py:39-62), n_days fused
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void step_group(const DevPlan& P, const DevState& S, const StepArgs& a, Gx& b, int group,
                                           int n_groups) {
  constexpr int C = spec::C, CLG = spec::CLG, LG = spec::LG, L = spec::L, G = spec::G, F = spec::F, nC = spec::I * spec::L;
  const size_t nA = (size_t)spec::n_acc * LG, nM = (size_t)spec::n_slot * LG;
  {
    // the CSC / CSR rows of this CTA's locales in padded (ELL) form, built once per kernel: source / destination
    // locale of (entry q, locale) in shared memory, mobility value of (entry q, thread) in the CTA's global scratch;
    // padding entries point at the own locale with weight 0
    int* eci = (int*)(b.smem + GL.s_eci); int* eri = (int*)(b.smem + GL.s_eri);
    float* evc = (float*)(b.cs + GL.c_evc); float* evr = (float*)(b.cs + GL.c_evr);
    for (int i = b.tid; i < spec::grp_nqc * spec::grp_lpc; i += b.T) {
      const int q = i / spec::grp_lpc, jl = i - q * spec::grp_lpc, j = jl < b.nl ? b.l0 + jl : b.l0;
      const int beg = P.csc_indptr[j], len = jl < b.nl ? P.csc_indptr[j + 1] - beg : 0;
      eci[i] = q < len ? P.csc_indices[beg + q] : j;
    }
    for (int i = b.tid; i < spec::grp_nqr * spec::grp_lpc; i += b.T) {
      const int q = i / spec::grp_lpc, jl = i - q * spec::grp_lpc, j = jl < b.nl ? b.l0 + jl : b.l0;
      const int beg = P.csr_indptr[j], len = jl < b.nl ? P.csr_indptr[j + 1] - beg : 0;
      eri[i] = q < len ? P.csr_indices[beg + q] : j;
    }
    {
      const int cb = P.csc_indptr[b.j], cl = b.valid ? P.csc_indptr[b.j + 1] - cb : 0;
      for (int q = 0; q < spec::grp_nqc; q++) evc[q * GTH + b.tid] = q < cl ? P.mx_cscT[(size_t)(cb + q) * spec::G + b.u] : 0.0f;
      const int rb = P.csr_indptr[b.j], rl = b.valid ? P.csr_indptr[b.j + 1] - rb : 0;
      for (int q = 0; q < spec::grp_nqr; q++) evr[q * GTH + b.tid] = q < rl ? P.mx_csrT[(size_t)(rb + q) * spec::G + b.u] : 0.0f;
    }
    {
      // entries (a multiple of the chunk) that cover the longest CSC / CSR row among this WARP's cells
      int mc = b.valid ? P.csc_indptr[b.j + 1] - P.csc_indptr[b.j] : 0, mr = b.valid ? P.csr_indptr[b.j + 1] - P.csr_indptr[b.j] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, mc, o), orr = __shfl_xor_sync(0xffffffffu, mr, o);
        mc = oc > mc ? oc : mc; mr = orr > mr ? orr : mr;
      }
      mc = (mc + EPI_GRP_CHUNK - 1) / EPI_GRP_CHUNK * EPI_GRP_CHUNK;
      mr = (mr + EPI_GRP_CHUNK - 1) / EPI_GRP_CHUNK * EPI_GRP_CHUNK;
#ifdef EPI_HOST_EMU
      // the host emulation's __syncwarp is a CTA-wide barrier: every warp has to walk the same number of chunks (the
      // CTA's padded row length; the padding gathers the own locale with weight 0, so the sums do not change)
      mc = spec::grp_nqc; mr = spec::grp_nqr;
#endif
      b.nq_w = mc | (mr << 16);
    }
    // selects / MOVE ops covering this thread's cell, and each select's compartment mask and mode
    b.selmask = b.mvmask = 0;
    if (b.valid) {
      for (int s_ = 0; s_ < P.n_sel; s_++)
        if (P.sel_l[s_ * spec::L + b.j] && P.sel_g[s_ * spec::G + b.u]) b.selmask |= 1u << s_;
      for (int mo = 0; mo < P.n_mv_op; mo++)
        if (P.mv_g[mo * spec::G + b.u] && P.mv_l1[mo * spec::L + b.j]) b.mvmask |= 1u << mo;
    }
    for (int s_ = b.tid; s_ < P.n_sel; s_ += b.T) {
      int cm = 0;
      for (int k = 0; k < spec::C; k++) cm |= P.sel_c[s_ * spec::C + k] ? 1 << k : 0;
      ((int2*)(b.tab + GL.e_selinfo))[s_] = int2{cm, P.sel_mode[s_]};
    }
    if (!spec::has_ft_ops && spec::grp_tden_smem)
      for (int i = b.tid; i < spec::n_lc * spec::G * GL.trow; i += b.T) ((real*)(b.tab + GL.e_Tden))[i] = (real)__ldg(P.Tden_constP + i);
    __syncthreads();
  }
  b.crD = (double*)(b.gs + GL.g_crD);
  b.mvD = (float*)(b.gs + GL.g_mvD);
  for (int env = group; env < a.N; env += n_groups) {
    const size_t e = (size_t)env;
    b.cost = S.cost + e * nC;
    b.crY = S.crY + e * nC;
    b.gX = S.X + e * CLG;
    real* gAcc = (real*)S.acc + e * nA;
    b.acc = gAcc;
    b.accn = (real*)(b.gs + GL.g_accN);
    b.mvY = S.mvY + e * nM;
    double* gXprev = S.Xprev + e * P.n_chg * LG;
    int* meta = S.meta + e * 8;
    b.s0 = 0; b.s1 = 1;
    if (b.valid) {
#pragma unroll
      for (int k = 0; k < C; k++) GXF(k) = (real)b.gX[(size_t)k * LG + b.cell];
    }
    for (int i = b.tid; i < P.n_cls; i += b.T) GT(mult_y, i) = S.multY[e * P.n_cls + i];
    for (int k = b.tid; k < P.NCP; k += b.T) {
      float v;
      if (a.cpv) v = a.cpv[e * P.NCP + k];
      else if (a.init_only) v = P.init_cpv[k];
      else { int s = P.cp_src[k]; v = s >= 0 ? a.actions[e * P.A + s] : P.cp_fixed[k]; }
      GT(cpv, k) = v;
    }
    for (int i = b.tid; i < P.I; i += b.T)
      GT(act, i) = a.active ? a.active[e * P.I + i] : (a.init_only ? P.init_active[i] : 1);
    int t_day = meta[0], ex_y = meta[1];
    __syncthreads();

    double reward = 0;  // this thread's share: -(cost after - cost before) of its cost entries, summed over the days
    bool tabs_current = false, fsal_ok = false;
    const int n_days = a.init_only ? 1 : a.n_days;
    const int nCl = spec::I * b.nl;
    for (int day = 0; day < n_days; day++) {
      if (day > 0 && a.cpv_days) {  // schedule mode: today's row (every CTA keeps its own copy of the env's tables)
        const size_t row = (size_t)day * a.N + e;
        for (int k = b.tid; k < P.NCP; k += b.T) GT(cpv, k) = a.cpv_days[row * P.NCP + k];
        for (int i = b.tid; i < P.I; i += b.T) GT(act, i) = a.active_days ? a.active_days[row * P.I + i] : 1;
        __syncthreads();
      }
      bool dyn = true;
      TSEC(14);
      int ex_bits = prologue(P, b, a.variant, gXprev, ex_y, dyn, tabs_current);
      TSEC(0);
      if (!a.init_only) {
        double c0 = 0, c1 = 0;
        for (int q = b.tid; q < nCl; q += b.T) c0 += b.cost[(size_t)(q / b.nl) * L + b.l0 + q % b.nl];
        const bool skip0 = EPI_FSAL_DAYS && fsal_ok && !dyn;
        fsal_ok = rk45_day(P, b, dyn, ex_bits, meta, S.last_err + e, a.dbg_attempts ? a.dbg_attempts + e * 128 : nullptr,
                           skip0);
        for (int q = b.tid; q < nCl; q += b.T) c1 += b.cost[(size_t)(q / b.nl) * L + b.l0 + q % b.nl];
        reward -= c1 - c0;  // -sum(cumulative_cost_next - cumulative_cost_current), epidemic.py:115
        t_day++;
        if (a.reward_days) {  // schedule mode: the day's reward needs its own group sum
          double dr[1] = {-(c1 - c0)};
          grp_sum<1>(P, b, dr);
          if (b.cta == 0 && b.tid == 0) a.reward_days[(size_t)day * a.N + env] = dr[0];
        }
        const size_t row = (size_t)day * a.N + e;
        if (b.valid) {
          if (a.obs_days)
            for (int k = 0; k < C; k++) ((real*)a.obs_days)[row * CLG + (size_t)k * LG + b.cell] = (real)b.gX[(size_t)k * LG + b.cell];
          if (a.x_days)
            for (int k = 0; k < C; k++) a.x_days[row * CLG + (size_t)k * LG + b.cell] = b.gX[(size_t)k * LG + b.cell];
          if (a.acc_days)
            for (int a_ = 0; a_ < spec::n_acc; a_++) ((real*)a.acc_days)[row * nA + (size_t)a_ * LG + b.cell] = b.acc[(size_t)a_ * LG + b.cell];
        }
        if (a.mult_days && b.cta == 0)
          for (int i = b.tid; i < P.n_cls; i += b.T) a.mult_days[row * P.n_cls + i] = GT(mult_d, i);
      }
      // today's parameters become yesterday's (State(t+1, next_obs, dest_parameter), epidemic.py:89)
      __syncthreads();
      for (int i = b.tid; i < P.n_cls; i += b.T) GT(mult_y, i) = GT(mult_d, i);
      if (b.valid)
        for (int s = 0; s < spec::n_slot; s++) b.mvY[(size_t)s * LG + b.cell] = b.mvD[(size_t)s * LG + b.cell];
      for (int q = b.tid; q < nCl; q += b.T) {
        const size_t i = (size_t)(q / b.nl) * L + b.l0 + q % b.nl;
        b.crY[i] = b.crD[i];
      }
      ex_y = (ex_bits >> 16) & 0xffff;
      __syncthreads();
      if (t_day >= P.horizon) break;  // the reference's EpiEnv.step stops its day loop at done
    }
    double rw[1] = {reward};
    TSEC(10);
    grp_sum<1>(P, b, rw);
    TSEC(0);
    if (b.valid) {
#pragma unroll
      for (int k = 0; k < C; k++) {
        const double x = b.gX[(size_t)k * LG + b.cell];
        if (a.obs_out) ((real*)a.obs_out)[e * CLG + (size_t)k * LG + b.cell] = (real)x;
        if (a.obs_out2) ((real*)a.obs_out2)[e * CLG + (size_t)k * LG + b.cell] = (real)x;
      }
      if (b.acc != gAcc)
        for (int a_ = 0; a_ < spec::n_acc; a_++) gAcc[(size_t)a_ * LG + b.cell] = b.acc[(size_t)a_ * LG + b.cell];
    }
    if (b.cta == 0) {
      for (int i = b.tid; i < P.n_cls; i += b.T) S.multY[e * P.n_cls + i] = GT(mult_y, i);
      if (b.tid == 0) {
        meta[0] = t_day; meta[1] = ex_y;
        if (a.reward_out) a.reward_out[env] = rw[0];
        if (a.done_out) a.done_out[env] = (uint8_t)(t_day >= P.horizon);
        if (a.reward_out2) a.reward_out2[env] = rw[0];
        if (a.done_out2) a.done_out2[env] = (uint8_t)(t_day >= P.horizon);
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ Gx bind(const DevPlan& P, const StepArgs& a, char* smem, int tid, int nt, int block) {
  Gx b;
  b.tid = tid; b.T = nt;   // nt == GTH
  b.S = GSS;
  const int group = block / GSS;
  b.cta = block - group * GSS;
  b.l0 = (b.cta * spec::L) / GSS;
  b.nl = ((b.cta + 1) * spec::L) / GSS - b.l0;
  const int jl = tid / spec::G;
  b.u = tid - jl * spec::G;
  b.valid = jl < b.nl;
  b.j = b.valid ? b.l0 + jl : b.l0;
  b.cell = b.j * spec::G + b.u;
  b.lc = P.lclass[b.j];
  b.smem = smem;
  b.tab = smem + GL.s_tab;
  b.gs = a.grp_scratch + (size_t)group * GL.g_bytes;
  b.cs = a.grp_cta_scratch + (size_t)block * GL.c_bytes;
  b.bar = a.grp_bar + (size_t)group * 32;  // one 128-byte line per group
  b.epoch = 0;
  b.parity = 0;
  b.s0 = 0; b.s1 = 1;
  b.gX = b.cost = b.crY = b.crD = nullptr;
  b.acc = b.accn = nullptr;
  b.mvY = b.mvD = nullptr;
  return b;
}

#ifndef EPI_HOST_EMU
// EPI_GRP_R CTAs per SM. The register file is four 16 K-register partitions, one per scheduler, and a CTA's warps are dealt
// round-robin to them: registers per thread <= 16384 / (32 * ceil(warps / 4))  (448 threads -> 128; a kernel compiled
// for 144 cannot be resident at all, which a cooperative launch reports as "too many blocks")
#define EPI_GRP_MAXREG_RAW (16384 / (32 * ((EPI_GRP_R * (EPI_GRP_THREADS / 32) + 3) / 4)) / 8 * 8)
#define EPI_GRP_MAXREG (EPI_GRP_MAXREG_RAW > 255 ? 255 : EPI_GRP_MAXREG_RAW)
__global__ void __maxnreg__(EPI_GRP_MAXREG) grp_kernel(const DevPlan P, const DevState S, const StepArgs a) {
  extern __shared__ __align__(16) char grp_smem[];
  int vb = blockIdx.x;
#if EPI_GRP_R > 1
  {
    // R CTAs share an SM: they must belong to DIFFERENT groups to fill each other's barrier and gather stalls, and the
    // hardware's block placement is not ours to choose. The k-th CTA to arrive on an SM joins class k; the CTAs of a
    // class are numbered in arrival order and cut into groups of S (all n_groups * S CTAs are resident: cooperative
    // launch, exactly R per SM). Counters: a.grp_bar + 32 * n_groups ... (zeroed per launch).
    __shared__ int s_vb;
    if (threadIdx.x == 0) s_vb = blockIdx.x;
    // (a launch with fewer groups than the device holds leaves SMs half empty: any numbering will do)
    if (threadIdx.x == 0 && GL.dyn_assign && (int)gridDim.x == GL.n_groups * GSS) {
      const int n_groups = gridDim.x / GSS, per_class = gridDim.x / EPI_GRP_R;
      unsigned* ctr = a.grp_bar + (size_t)32 * n_groups;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const unsigned cls = atomicAdd(ctr + 8 + smid, 1u) % EPI_GRP_R;
      const unsigned r = atomicAdd(ctr + cls, 1u);
      s_vb = (int)(cls * per_class + r);
    }
    __syncthreads();
    vb = s_vb;
  }
#endif
  Gx b = bind(P, a, grp_smem, threadIdx.x, blockDim.x, vb);
#ifdef EPI_GRP_TIMING
  __shared__ long long s_tacc[16];
  b.tacc = s_tacc; b.tlast = clock64(); b.tcur = 0;
  if (threadIdx.x == 0) for (int i = 0; i < 16; i++) s_tacc[i] = 0;
  __syncthreads();
#endif
  step_group(P, S, a, b, vb / GSS, gridDim.x / GSS);
#ifdef EPI_GRP_TIMING
  TSEC(0);
  if (threadIdx.x == 0 && (vb == 0 || vb == 5)) {
    printf("grp timing cta %d:", vb);
    for (int i = 0; i < 15; i++) printf(" s%d=%.3fms", i, (double)b.tacc[i] / 1.9e6);
    printf("\n");
  }
#endif
}
#endif

}  // namespace grp
}  // namespace epi
#endif  // EPI_SPEC
