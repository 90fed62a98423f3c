/* epi_oracle.c — CPU ORACLE for the Epipolicy_RL hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file. The product (epipolicy_rl_b200/) never does.
 *
 * It is a plain-C restatement, for ONE environment at a time, of what the reference does per
 * simulated day, in the reference's own precision map (A.5):
 *   execute()            <- epipolicy/core/executor.py:272-284, nb_apply :116-164, nb_move :166-234,
 *                           nb_add :236-243, select :292-370 (compartment shape, ['Value'].sum())
 *   dest_parameter()     <- epipolicy/obj/parameter.py:53-66, matrix/coo.py:173-188 (mobility rows),
 *                           matrix/dense.py:6-32 (facility normalisers)
 *   gap / interpolation  <- obj/parameter.py:68-92, utility/utils.py:34-37, matrix/coo.py:193-227
 *   rhs()                <- core/core_optimize.py:28-222 (literal loop nest, incl. its quirks)
 *   rk45_day()           <- scipy.integrate.solve_ivp(method="RK45") as called at core/epidemic.py:87
 *                           with default rtol=1e-3, atol=1e-6; scipy is a third-party dependency
 *                           (requirements pin scipy==1.6.2, installed 1.18.1): restated from
 *                           scipy/integrate/_ivp/rk.py:14-71,86-176,538-552 and common.py:63-134.
 *   day_step()/env_step()<- core/epidemic.py:82-116, epipolicy_environment.py:39-62
 *
 * PARITY PINNING: the reference ships no tests or golden vectors. This oracle is
 * pinned against outputs of the reference itself run in the build container: tests/golden/ (npz)
 * (made by tools/gen_golden.py), incl. the one recorded number in the reference repo
 * (user.ipynb:82, reward -4566.504778395923). See tests/test_oracle_golden.py.
 *
 * Known deviation: the reference iterates the move table in nested-dict order (g, c1, c2, l1, l2
 * by first insertion at each level, core_optimize.py:197-207); this file iterates in flat
 * first-insertion order. The two coincide unless two MOVE ops interleave their keys.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/epi_desc.h"

#define FLOAT_EPSILON 1e-15 /* utility/singleton.py:1 */

typedef struct { int g, c1, c2, l1, l2; float v; } Move;
typedef struct { Move* e; int n, cap; } MoveTab;

typedef struct {
  float *mp, *ft, *fi, *mob; /* [P,L,F,G] [L,F,G] [L,F,G,G] [G*MODE_NNZ_TOTAL] */
  double* cost;              /* [I,L] */
  MoveTab mv;
} Param;

typedef struct Oracle {
  EpiDesc d;
  void* owned[80];
  int n_owned;
  int CLG, n; /* n = flat ODE length (obj/observable.py:24-25) */
  int nmp, nft, nfi, nmob, ncost;
  double* y;      /* current flat observable */
  double* Xprev;  /* current_comp of the previous state (epidemic.py:96-99) */
  double* X0;
  Param yest;     /* state.unobs */
  Param init;     /* schedule.initial_state.unobs */
  int t;          /* len(history.states) */
  /* scratch */
  Param dest, gap, cur;
  float *dmp, *dft, *dfi, *dmob;
  double *K[7], *ytmp, *ynew, *scale, *f1;
  double *alive, *foi, *edge, *next;
  float *comb_csr, *comb_csc;
  int* csc_to_csr_dummy;
  /* trace of the last day */
  int nfev, nsteps, nrej, status;
  double last_err;
  int last_nfev_total;
  double attempts[128]; /* (h, err) of each attempted RK step of the last day */
  int n_attempts;
} Oracle;

/* ------------------------------------------------------------------ utilities */
static void* xcalloc(size_t n, size_t s) {
  void* p = calloc(n ? n : 1, s);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void mv_clear(MoveTab* t) { t->n = 0; }
static void mv_free(MoveTab* t) { free(t->e); t->e = NULL; t->n = t->cap = 0; }
static int mv_find(const MoveTab* t, int g, int c1, int c2, int l1, int l2) {
  for (int i = 0; i < t->n; i++) {
    const Move* m = &t->e[i];
    if (m->g == g && m->c1 == c1 && m->c2 == c2 && m->l1 == l1 && m->l2 == l2) return i;
  }
  return -1;
}
static int mv_slot(MoveTab* t, int g, int c1, int c2, int l1, int l2) {
  int i = mv_find(t, g, c1, c2, l1, l2);
  if (i >= 0) return i;
  if (t->n == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 16;
    t->e = (Move*)realloc(t->e, sizeof(Move) * t->cap);
  }
  Move m = {g, c1, c2, l1, l2, 0.0f};
  t->e[t->n] = m;
  return t->n++;
}
/* CooMatrix.element_add (matrix/coo.py:52-57): f32 cell += value, stored back as f32 */
static void mv_add(MoveTab* t, int g, int c1, int c2, int l1, int l2, double amount) {
  int i = mv_slot(t, g, c1, c2, l1, l2);
  t->e[i].v = (float)((double)t->e[i].v + amount);
}
static void mv_copy(MoveTab* dst, const MoveTab* src) {
  if (dst->cap < src->n) {
    dst->cap = src->n;
    dst->e = (Move*)realloc(dst->e, sizeof(Move) * (dst->cap ? dst->cap : 1));
  }
  dst->n = src->n;
  if (src->n) memcpy(dst->e, src->e, sizeof(Move) * src->n);
}

static void param_alloc(Oracle* o, Param* p) {
  p->mp = (float*)xcalloc(o->nmp, 4);
  p->ft = (float*)xcalloc(o->nft, 4);
  p->fi = (float*)xcalloc(o->nfi, 4);
  p->mob = (float*)xcalloc(o->nmob, 4);
  p->cost = (double*)xcalloc(o->ncost, 8);
  p->mv.e = NULL; p->mv.n = p->mv.cap = 0;
}
static void param_free(Param* p) {
  free(p->mp); free(p->ft); free(p->fi); free(p->mob); free(p->cost); mv_free(&p->mv);
}
static void param_copy(Oracle* o, Param* dst, const Param* src) {
  memcpy(dst->mp, src->mp, 4 * (size_t)o->nmp);
  memcpy(dst->ft, src->ft, 4 * (size_t)o->nft);
  memcpy(dst->fi, src->fi, 4 * (size_t)o->nfi);
  memcpy(dst->mob, src->mob, 4 * (size_t)o->nmob);
  memcpy(dst->cost, src->cost, 8 * (size_t)o->ncost);
  mv_copy(&dst->mv, &src->mv);
}

#define LST(o, id, ptr, len) \
  const int32_t* ptr = (o)->d.list_data + (o)->d.list_off[id]; \
  int len = (o)->d.list_off[(id) + 1] - (o)->d.list_off[id]

static int in_list(const int32_t* p, int n, int v) {
  for (int i = 0; i < n; i++) if (p[i] == v) return 1;
  return 0;
}

/* ------------------------------------------------------------------ executor */
/* select(...)['Value'].sum() over current_comp (executor.py:346-359) */
static double eval_select(const Oracle* o, int sel, const double* X) {
  const EpiDesc* d = &o->d;
  const int32_t* a = d->sel_arg + 4 * sel;
  LST(o, a[0], cs, nc); LST(o, a[1], ls, nl); LST(o, a[2], gs, ng);
  int LG = d->L * d->G;
  double s = 0.0;
  for (int ic = 0; ic < nc; ic++)
    for (int il = 0; il < nl; il++)
      for (int ig = 0; ig < ng; ig++) {
        int k = cs[ic] * LG + ls[il] * d->G + gs[ig];
        double v = X[k];
        if (a[3] == EPI_SEL_DEFAULT) v = o->X0[k];
        else if (a[3] == EPI_SEL_CHANGE) v = X[k] - o->Xprev[k];
        s += v;
      }
  return s;
}

/* expression VM: numpy scalar semantics recorded by export.py (float32 nodes are rounded) */
static double eval_expr(const Oracle* o, int e, const float* cpv, const double* X) {
  const EpiDesc* d = &o->d;
  double st[32];
  int sp = 0;
  for (int k = d->expr_off[e]; k < d->expr_off[e + 1]; k++) {
    int op = d->tok_op[k];
    double r;
    switch (op) {
      case EPI_TOK_CONST: r = d->tok_val[k]; break;
      case EPI_TOK_CP: r = (double)cpv[d->tok_arg[k]]; break;
      case EPI_TOK_SEL: r = eval_select(o, d->tok_arg[k], X); break;
      case EPI_TOK_NEG: r = -st[--sp]; break;
      default: {
        double b = st[--sp], a = st[--sp];
        r = op == EPI_TOK_ADD ? a + b : op == EPI_TOK_SUB ? a - b : op == EPI_TOK_MUL ? a * b : a / b;
      }
    }
    if (d->tok_f32[k]) r = (double)(float)r;
    st[sp++] = r;
  }
  return st[0];
}

/* nb_move (executor.py:166-234) */
static void exec_move(Oracle* o, const int32_t* arg, double value, const double* X, MoveTab* mv) {
  const EpiDesc* d = &o->d;
  int G = d->G, LG = d->L * d->G;
  LST(o, arg[0], c1s, nc1); LST(o, arg[1], l1s, nl1); LST(o, arg[2], g1s, ng1);
  LST(o, arg[3], c2s, nc2); LST(o, arg[4], l2s, nl2); LST(o, arg[5], g2s, ng2);
  double depart_sum = 0;
  for (int ig = 0; ig < ng1; ig++) {
    int g = g1s[ig];
    if (!in_list(g2s, ng2, g)) continue;
    for (int ic = 0; ic < nc1; ic++)
      for (int il = 0; il < nl1; il++) depart_sum += X[c1s[ic] * LG + l1s[il] * G + g];
  }
  if (!(depart_sum > 0)) return;
  for (int ig = 0; ig < ng1; ig++) {
    int g = g1s[ig];
    if (!in_list(g2s, ng2, g)) continue;
    for (int ic = 0; ic < nc1; ic++) {
      int c1 = c1s[ic];
      for (int il = 0; il < nl1; il++) {
        int l1 = l1s[il];
        double depart = X[c1 * LG + l1 * G + g] / depart_sum * value;
        if (in_list(c2s, nc2, c1)) {
          if (!in_list(l2s, nl2, l1)) {
            double arrive = 0;
            for (int j = 0; j < nl2; j++) arrive += X[c1 * LG + l2s[j] * G + g];
            for (int j = 0; j < nl2; j++) {
              double amt = arrive > 0 ? X[c1 * LG + l2s[j] * G + g] / arrive * depart : 1.0 / nl2 * depart;
              mv_add(mv, g, c1, c1, l1, l2s[j], amt);
            }
          }
        } else if (in_list(l2s, nl2, l1)) {
          double arrive = 0;
          for (int j = 0; j < nc2; j++) arrive += X[c2s[j] * LG + l1 * G + g];
          for (int j = 0; j < nc2; j++) {
            double amt = arrive > 0 ? X[c2s[j] * LG + l1 * G + g] / arrive * depart : 1.0 / nc2 * depart;
            mv_add(mv, g, c1, c2s[j], l1, l1, amt);
          }
        } else {
          double arrive = 0;
          for (int j = 0; j < nc2; j++)
            for (int k = 0; k < nl2; k++) arrive += X[c2s[j] * LG + l2s[k] * G + g];
          for (int j = 0; j < nc2; j++)
            for (int k = 0; k < nl2; k++) {
              double amt = arrive > 0 ? X[c2s[j] * LG + l2s[k] * G + g] / arrive * depart
                                      : 1.0 / (nc2 * nl2) * depart;
              mv_add(mv, g, c1, c2s[j], l1, l2s[k], amt);
            }
        }
      }
    }
  }
}

/* Executor.execute (executor.py:272-284) on the compiled programs; delta multipliers start at 1,
 * cost at 0 (obj/delta_parameter.py:58-67). */
static void execute(Oracle* o, const double* X, const float* cpv, const int32_t* active,
                    const int32_t* prog_of, double* cost, MoveTab* mv) {
  const EpiDesc* d = &o->d;
  int L = d->L, G = d->G, F = d->F;
  for (int i = 0; i < o->nmp; i++) o->dmp[i] = 1.0f;
  for (int i = 0; i < o->nft; i++) o->dft[i] = 1.0f;
  for (int i = 0; i < o->nfi; i++) o->dfi[i] = 1.0f;
  for (int i = 0; i < o->nmob; i++) o->dmob[i] = 1.0f;
  memset(cost, 0, 8 * (size_t)o->ncost);
  mv_clear(mv);
  for (int itv = 0; itv < d->I; itv++) {
    if (!active[itv] || prog_of[itv] < 0) continue;
    int pr = prog_of[itv];
    for (int op = d->prog_off[pr]; op < d->prog_off[pr + 1]; op++) {
      const int32_t* arg = d->op_arg + 8 * op;
      double val = eval_expr(o, d->op_expr[op], cpv, X);
      switch (d->op_kind[op]) {
        case EPI_OP_PARAM_MUL: {
          LST(o, arg[0], ps, np_); LST(o, arg[1], ls, nl); LST(o, arg[2], fs, nf); LST(o, arg[3], gs, ng);
          float m = (float)val;
          for (int a = 0; a < np_; a++) for (int b = 0; b < nl; b++) for (int c = 0; c < nf; c++) for (int e = 0; e < ng; e++)
            o->dmp[((ps[a] * L + ls[b]) * F + fs[c]) * G + gs[e]] *= m;
        } break;
        case EPI_OP_FI_MUL: {
          LST(o, arg[0], ls, nl); LST(o, arg[1], fs, nf); LST(o, arg[2], g1, n1); LST(o, arg[3], g2, n2);
          float m = (float)val;
          for (int b = 0; b < nl; b++) for (int c = 0; c < nf; c++) for (int e = 0; e < n1; e++) for (int h = 0; h < n2; h++)
            o->dfi[((ls[b] * F + fs[c]) * G + g1[e]) * G + g2[h]] *= m;
        } break;
        case EPI_OP_FT_MUL: {
          LST(o, arg[0], ls, nl); LST(o, arg[1], fs, nf); LST(o, arg[2], gs, ng);
          float m = (float)val;
          for (int b = 0; b < nl; b++) for (int c = 0; c < nf; c++) for (int e = 0; e < ng; e++)
            o->dft[(ls[b] * F + fs[c]) * G + gs[e]] *= m;
        } break;
        case EPI_OP_MOB_MUL: {
          /* CooMatrix.pairwise_multiply (coo.py:112-117): existing (row in from, col in to) cells */
          int m_ = arg[0];
          LST(o, arg[1], gs, ng); LST(o, arg[2], fr, nfr); LST(o, arg[3], to, nto);
          float m = (float)val;
          const int32_t* ip = d->mode_indptr + m_ * (L + 1);
          const int32_t* ix = d->mode_indices + d->mode_off[m_];
          int nnz_m = d->mode_off[m_ + 1] - d->mode_off[m_];
          float* blk = o->dmob + (size_t)G * d->mode_off[m_];
          for (int e = 0; e < ng; e++) for (int b = 0; b < nfr; b++) {
            int r = fr[b];
            for (int k = ip[r]; k < ip[r + 1]; k++)
              if (in_list(to, nto, ix[k])) blk[gs[e] * nnz_m + k] *= m;
          }
        } break;
        case EPI_OP_MOVE: exec_move(o, arg, val, X, mv); break;
        case EPI_OP_ADD: {
          LST(o, arg[0], is, ni); LST(o, arg[1], ls, nl);
          int parts = ni * nl;
          if (parts > 0)
            for (int a = 0; a < ni; a++) for (int b = 0; b < nl; b++) cost[is[a] * L + ls[b]] += val / parts;
        } break;
        default: break;
      }
    }
  }
}

/* get_parameter_from_delta (obj/parameter.py:53-66) */
static void dest_parameter(Oracle* o, Param* dst) {
  const EpiDesc* d = &o->d;
  int L = d->L, G = d->G, F = d->F;
  /* mobility: origin (*) delta, rows rescaled to the origin row sum (matrix/coo.py:173-188) */
  for (int m = 0; m < d->M; m++) {
    const int32_t* ip = d->mode_indptr + m * (L + 1);
    const int32_t* ix = d->mode_indices + d->mode_off[m];
    int nnz_m = d->mode_off[m + 1] - d->mode_off[m];
    for (int g = 0; g < G; g++) {
      const float* org = d->mode_origin + (size_t)G * d->mode_off[m] + (size_t)g * nnz_m;
      const float* del = o->dmob + (size_t)G * d->mode_off[m] + (size_t)g * nnz_m;
      float* out = dst->mob + (size_t)G * d->mode_off[m] + (size_t)g * nnz_m;
      for (int r = 0; r < L; r++) {
        double osum = 0, rsum = 0;
        for (int k = ip[r]; k < ip[r + 1]; k++) {
          osum += org[k];
          rsum += (double)(float)(org[k] * del[k]);
        }
        for (int k = ip[r]; k < ip[r + 1]; k++) {
          if (rsum > FLOAT_EPSILON) out[k] = (float)((double)(float)(org[k] * del[k]) / rsum * osum);
          else out[k] = ix[k] == r ? (float)osum : 0.0f;
        }
      }
    }
  }
  for (int i = 0; i < o->nmp; i++) dst->mp[i] = d->mp0[i] * o->dmp[i];
  /* facility_timespent normalised over F (matrix/dense.py:6-19); numba np.sum(axis=1) accumulates in f32 */
  for (int l = 0; l < L; l++) for (int g = 0; g < G; g++) {
    float s = 0.0f;
    for (int f = 0; f < F; f++) s += d->ft0[(l * F + f) * G + g] * o->dft[(l * F + f) * G + g];
    for (int f = 0; f < F; f++) {
      float v = d->ft0[(l * F + f) * G + g] * o->dft[(l * F + f) * G + g];
      dst->ft[(l * F + f) * G + g] = s > FLOAT_EPSILON ? v / s : (f == 0 ? 1.0f : 0.0f);
    }
  }
  /* facility_interaction rows renormalised only if the row sum is > 1 or < 0 (dense.py:21-32) */
  for (int l = 0; l < L; l++) for (int f = 0; f < F; f++) for (int g1 = 0; g1 < G; g1++) {
    size_t b = ((size_t)(l * F + f) * G + g1) * G;
    float s = 0.0f;
    for (int g2 = 0; g2 < G; g2++) s += d->fi0[b + g2] * o->dfi[b + g2];
    for (int g2 = 0; g2 < G; g2++) {
      float v = d->fi0[b + g2] * o->dfi[b + g2];
      dst->fi[b + g2] = (s > 1.0f || s < 0.0f) ? v / s : v;
    }
  }
}

/* get_gap_parameter (obj/parameter.py:68-78): gap = current - previous */
static void gap_parameter(Oracle* o, const Param* prev, const Param* cur, Param* gap) {
  for (int i = 0; i < o->nmp; i++) gap->mp[i] = cur->mp[i] - prev->mp[i];
  for (int i = 0; i < o->nft; i++) gap->ft[i] = cur->ft[i] - prev->ft[i];
  for (int i = 0; i < o->nfi; i++) gap->fi[i] = cur->fi[i] - prev->fi[i];
  for (int i = 0; i < o->nmob; i++) gap->mob[i] = cur->mob[i] - prev->mob[i];
  for (int i = 0; i < o->ncost; i++) gap->cost[i] = cur->cost[i] - prev->cost[i];
  /* add_d3(scale_d3(copy_d3_coo(prev), -1), cur)  (coo.py:193-227) */
  mv_copy(&gap->mv, &prev->mv);
  for (int i = 0; i < gap->mv.n; i++) gap->mv.e[i].v = (float)((double)gap->mv.e[i].v * -1.0);
  for (int i = 0; i < cur->mv.n; i++) {
    const Move* m = &cur->mv.e[i];
    int k = mv_slot(&gap->mv, m->g, m->c1, m->c2, m->l1, m->l2);
    gap->mv.e[k].v = gap->mv.e[k].v + m->v;
  }
}

/* get_interpolated_parameter (obj/parameter.py:80-92) */
static void interpolate(Oracle* o, const Param* src, const Param* gap, double step, Param* out) {
  float fs = (float)step;
  float qs = (float)((-3.0 * step + 4.0) * step); /* utility/utils.py:34-37 */
  for (int i = 0; i < o->nmob; i++) out->mob[i] = src->mob[i] + gap->mob[i] * fs;
  for (int i = 0; i < o->nmp; i++) out->mp[i] = src->mp[i] + gap->mp[i] * fs;
  for (int i = 0; i < o->nft; i++) out->ft[i] = src->ft[i] + gap->ft[i] * fs;
  for (int i = 0; i < o->nfi; i++) out->fi[i] = src->fi[i] + gap->fi[i] * fs;
  mv_copy(&out->mv, &gap->mv);
  for (int i = 0; i < out->mv.n; i++) out->mv.e[i].v = out->mv.e[i].v * qs;
  for (int i = 0; i < src->mv.n; i++) {
    const Move* m = &src->mv.e[i];
    int k = mv_slot(&out->mv, m->g, m->c1, m->c2, m->l1, m->l2);
    out->mv.e[k].v = out->mv.e[k].v + m->v;
  }
  for (int i = 0; i < o->ncost; i++) out->cost[i] = src->cost[i] + gap->cost[i] * (double)qs;
}

/* ------------------------------------------------------------------ right-hand side */
static double factor_val(const Oracle* o, int32_t f, const double* X, const float* mp, int l, int g) {
  const EpiDesc* d = &o->d;
  int kind = f >> 24, idx = f & 0xffffff;
  if (kind == EPI_FAC_PARAM) return (double)mp[((idx * d->L + l) * d->F + 0) * d->G + g]; /* facility 0! :48 */
  if (kind == EPI_FAC_N) return o->alive[l * d->G + g];
  return X[(idx * d->L + l) * d->G + g];
}

static void scatter_edge(const Oracle* o, int e, const double* flow, double* dy) {
  const EpiDesc* d = &o->d;
  int C = d->C, LG = d->L * d->G, CLG = o->CLG;
  for (int k = d->sc_off[e]; k < d->sc_off[e + 1]; k++) {
    double coef = (double)d->sc_coef[k];
    double* dst;
    switch (d->sc_kind[k]) {
      case EPI_SC_COMP: dst = dy + d->sc_a[k] * LG; break;
      case EPI_SC_GAIN: dst = dy + CLG + d->sc_a[k] * LG; break;
      case EPI_SC_LOST: dst = dy + 2 * CLG + d->sc_a[k] * LG; break;
      default: dst = dy + 3 * CLG + (d->sc_a[k] * C + d->sc_b[k]) * LG; break;
    }
    for (int i = 0; i < LG; i++) dst[i] += flow[i] * coef;
  }
}

/* scipy_fun / forward / compute_delta_observable (core_optimize.py:165-222) */
static void rhs(Oracle* o, double t, const double* y, double* dy) {
  const EpiDesc* d = &o->d;
  int C = d->C, L = d->L, G = d->G, F = d->F, LG = L * G, CLG = o->CLG, nnz = d->nnz;
  o->nfev++;
  interpolate(o, &o->yest, &o->gap, t, &o->cur);
  const Param* p = &o->cur;
  const double* X = y;
  memset(dy, 0, 8 * (size_t)o->n);
  /* alive matrix: N excludes death AND hospitalized (core_optimize.py:28-41,168) */
  for (int i = 0; i < LG; i++) o->alive[i] = 0;
  for (int c = 0; c < C; c++) for (int i = 0; i < LG; i++) o->alive[i] += X[c * LG + i] * d->alive_coef[c];
  /* compute_combined_mobility (:69-92): f32 sum over modes of bias*reduced on the union pattern */
  {
    int* cptr = (int*)alloca(sizeof(int) * (L + 1));
    int* mi = (int*)alloca(sizeof(int) * d->M);
    for (int i = 0; i <= L; i++) cptr[i] = d->csc_indptr[i];
    for (int r = 0; r < L; r++) {
      for (int m = 0; m < d->M; m++) mi[m] = d->mode_indptr[m * (L + 1) + r];
      for (int k = d->csr_indptr[r]; k < d->csr_indptr[r + 1]; k++) {
        int c = d->csr_indices[k];
        int dest = cptr[c]++;
        for (int g = 0; g < G; g++) o->comb_csr[g * nnz + k] = 0.0f;
        for (int m = 0; m < d->M; m++) {
          const int32_t* ip = d->mode_indptr + m * (L + 1);
          const int32_t* ix = d->mode_indices + d->mode_off[m];
          int nnz_m = d->mode_off[m + 1] - d->mode_off[m];
          if (mi[m] < ip[r + 1] && c == ix[mi[m]]) {
            const float* blk = p->mob + (size_t)G * d->mode_off[m];
            for (int g = 0; g < G; g++) o->comb_csr[g * nnz + k] += d->mode_bias[m] * blk[g * nnz_m + mi[m]];
            mi[m]++;
          }
        }
        for (int g = 0; g < G; g++) o->comb_csc[g * nnz + dest] = o->comb_csr[g * nnz + k];
      }
    }
  }
  /* regular edges (:171-181, compute_fraction :55-65) */
  for (int e = 0; e < d->E; e++) {
    for (int l = 0; l < L; l++) for (int g = 0; g < G; g++) {
      double numer = 1.0;
      for (int k = d->edge_numer_off[e]; k < d->edge_numer_off[e + 1]; k++) numer *= factor_val(o, d->fac[k], X, p->mp, l, g);
      double denom = 0.0;
      for (int s = d->edge_den_off[e]; s < d->edge_den_off[e + 1]; s++) {
        double sv = 1.0;
        for (int k = d->sum_fac_off[s]; k < d->sum_fac_off[s + 1]; k++) sv *= factor_val(o, d->fac[k], X, p->mp, l, g);
        denom += (double)d->sum_coef[s] * sv;
      }
      if (denom > 0) numer /= denom;
      o->edge[l * G + g] = numer;
    }
    scatter_edge(o, e, o->edge, dy);
  }
  /* infectious edges (:182-195) */
  for (int e = 0; e < d->IE; e++) {
    int c2 = d->ie_c2[e], c0 = d->ie_c0[e], p0 = d->ie_p0[e];
    /* compute_foi_matrix / compute_foi_entry (:98-134) */
    for (int u0 = 0; u0 < G; u0++) for (int j = 0; j < L; j++) for (int v = 0; v < F; v++) {
      double numer = 0, denom = 0;
      for (int u = 0; u < G; u++) {
        double ns = 0, ds = 0;
        for (int r = d->csc_indptr[j]; r < d->csc_indptr[j + 1]; r++) {
          int i = d->csc_indices[r];
          double val = (double)o->comb_csc[u * nnz + r];
          ns += X[(c2 * L + i) * G + u] * val;
          ds += o->alive[i * G + u] * val;
        }
        float ft = p->ft[(j * F + v) * G + u];
        float factor = ft * p->fi[((j * F + v) * G + u0) * G + u] * p->fi[((j * F + v) * G + u) * G + u0];
        numer += ns * (double)factor;
        denom += ds * (double)ft * (double)d->fi0[((j * F + v) * G + u0) * G + u] * (double)d->fi0[((j * F + v) * G + u) * G + u0];
      }
      if (denom <= FLOAT_EPSILON) denom = 1.0;
      o->foi[(u0 * L + j) * F + v] = numer / denom;
    }
    /* compute_infectious_edge (:138-163) */
    for (int i0 = 0; i0 < L; i0++) for (int u0 = 0; u0 < G; u0++) {
      double entry = 0;
      for (int k = d->csr_indptr[i0]; k < d->csr_indptr[i0 + 1]; k++) {
        int j = d->csr_indices[k];
        double val = (double)o->comb_csr[u0 * nnz + k];
        double tmp = 0;
        for (int v = 0; v < F; v++) {
          double coef = 1.0;
          for (int q = d->ie_np_off[e]; q < d->ie_np_off[e + 1]; q++) coef *= (double)p->mp[((d->ie_np[q] * L + j) * F + v) * G + u0];
          for (int q = d->ie_dp_off[e]; q < d->ie_dp_off[e + 1]; q++) {
            float mv = p->mp[((d->ie_dp[q] * L + j) * F + v) * G + u0];
            if (fabsf(mv) > 0) coef /= (double)mv;
          }
          tmp += (double)p->ft[(j * F + v) * G + u0] * o->foi[(u0 * L + j) * F + v] * (double)p->mp[((p0 * L + j) * F + v) * G + u0] * coef;
        }
        entry += tmp * val;
      }
      o->edge[i0 * G + u0] = entry * X[(c0 * L + i0) * G + u0];
    }
    scatter_edge(o, d->E + e, o->edge, dy);
  }
  /* clamped moves (:196-208) */
  for (int i = 0; i < CLG; i++) o->next[i] = X[i] + dy[i];
  for (int i = 0; i < p->mv.n; i++) {
    const Move* m = &p->mv.e[i];
    double v = (double)m->v;
    double have = o->next[(m->c1 * L + m->l1) * G + m->g];
    double nv = v < have ? v : have;
    o->next[(m->c1 * L + m->l1) * G + m->g] -= nv;
    o->next[(m->c2 * L + m->l2) * G + m->g] += nv;
    if (m->l1 == m->l2) dy[3 * CLG + ((m->c1 * C + m->c2) * L + m->l1) * G + m->g] += nv;
  }
  for (int i = 0; i < CLG; i++) dy[i] = o->next[i] - X[i];
  /* cost rate (:209) */
  for (int i = 0; i < o->ncost; i++) dy[(3 + C) * CLG + i] += p->cost[i];
}

/* ------------------------------------------------------------------ scipy RK45 */
static const double RK_C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

static double rms(const double* x, int n) { /* common.py:63-65 */
  double s = 0;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  return sqrt(s) / sqrt((double)n);
}

/* one day: solve_ivp(scipy_fun, (0,1), y0, method="RK45") -> y(1) written back into o->y */
static void rk45_day(Oracle* o) {
  const double rtol = 1e-3, atol = 1e-6, t_bound = 1.0;
  int n = o->n;
  double* y = o->y;
  double* f = o->K[0];
  double t = 0.0;
  o->nfev = 0; o->nsteps = 0; o->nrej = 0; o->status = 0; o->last_err = 0; o->n_attempts = 0;
  rhs(o, t, y, f);
  /* select_initial_step (common.py:68-134), order = 4 */
  double h_abs;
  {
    for (int i = 0; i < n; i++) o->scale[i] = atol + fabs(y[i]) * rtol;
    for (int i = 0; i < n; i++) o->ytmp[i] = y[i] / o->scale[i];
    double d0 = rms(o->ytmp, n);
    for (int i = 0; i < n; i++) o->ytmp[i] = f[i] / o->scale[i];
    double d1 = rms(o->ytmp, n);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > 1.0) h0 = 1.0;
    for (int i = 0; i < n; i++) o->ytmp[i] = y[i] + h0 * f[i];
    rhs(o, t + h0, o->ytmp, o->f1);
    for (int i = 0; i < n; i++) o->ytmp[i] = (o->f1[i] - f[i]) / o->scale[i];
    double d2 = rms(o->ytmp, n) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5);
    h_abs = fmin(fmin(100 * h0, h1), 1.0);
  }
  int guard = 0;
  while (t < t_bound) {
    if (++guard > 100000 || !(h_abs == h_abs)) { o->status = -2; return; }
    double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    int accepted = 0, rejected = 0;
    double t_new = t, h = 0;
    while (!accepted) {
      if (h_abs < min_step || ++guard > 100000) { o->status = -1; return; }
      h = h_abs;
      t_new = t + h;
      if (t_new - t_bound > 0) t_new = t_bound;
      h = t_new - t;
      h_abs = fabs(h);
      /* rk_step (rk.py:14-71) */
      for (int s = 1; s < 6; s++) {
        for (int i = 0; i < n; i++) {
          double dy = 0;
          for (int j = 0; j < s; j++) dy += o->K[j][i] * RK_A[s][j];
          o->ytmp[i] = y[i] + dy * h;
        }
        rhs(o, t + RK_C[s] * h, o->ytmp, o->K[s]);
      }
      for (int i = 0; i < n; i++) {
        double dy = 0;
        for (int j = 0; j < 6; j++) dy += o->K[j][i] * RK_B[j];
        o->ynew[i] = y[i] + h * dy;
      }
      rhs(o, t + h, o->ynew, o->K[6]);
      for (int i = 0; i < n; i++) {
        double sc = atol + fmax(fabs(y[i]), fabs(o->ynew[i])) * rtol;
        double e = 0;
        for (int j = 0; j < 7; j++) e += o->K[j][i] * RK_E[j];
        o->ytmp[i] = e * h / sc;
      }
      double err = rms(o->ytmp, n);
      o->last_err = err;
      if (o->n_attempts < 64) { o->attempts[2 * o->n_attempts] = h; o->attempts[2 * o->n_attempts + 1] = err; o->n_attempts++; }
      if (err < 1) {
        double factor = err == 0 ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
        if (rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        accepted = 1;
      } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
        rejected = 1;
        o->nrej++;
      }
    }
    o->nsteps++;
    t = t_new;
    memcpy(y, o->ynew, 8 * (size_t)n);
    memcpy(o->K[0], o->K[6], 8 * (size_t)n); /* f = f_new */
  }
}

/* ------------------------------------------------------------------ public API */
static void* own(Oracle* o, const void* src, size_t n, size_t s) {
  void* p = xcalloc(n, s);
  if (n) memcpy(p, src, n * s);
  o->owned[o->n_owned++] = p;
  return p;
}

static void set_initial(Oracle* o) {
  /* make_setting (epidemic.py:77-80): execute the schedule's day-0 action on the default state */
  const EpiDesc* d = &o->d;
  memcpy(o->Xprev, o->X0, 8 * (size_t)o->CLG);
  execute(o, o->X0, d->init_cpv, d->init_active, d->prog_init, o->init.cost, &o->init.mv);
  dest_parameter(o, &o->init);
}

Oracle* oracle_create(const EpiDesc* src) {
  Oracle* o = (Oracle*)xcalloc(1, sizeof(Oracle));
  o->d = *src;
  EpiDesc* d = &o->d;
  int C = d->C, L = d->L, G = d->G, F = d->F, P = d->P, I = d->I, M = d->M;
#define OWN(field, count) d->field = own(o, src->field, (size_t)(count), sizeof(*src->field))
  OWN(alive_coef, C); OWN(X0, C * L * G); OWN(pop, L * G); OWN(mp0, P * L * F * G); OWN(ft0, L * F * G);
  OWN(fi0, L * F * G * G); OWN(mode_bias, M); OWN(mode_indptr, M * (L + 1)); OWN(mode_off, M + 1);
  OWN(mode_indices, d->MODE_NNZ_TOTAL); OWN(mode_origin, G * d->MODE_NNZ_TOTAL);
  OWN(csr_indptr, L + 1); OWN(csr_indices, d->nnz); OWN(csc_indptr, L + 1); OWN(csc_indices, d->nnz);
  OWN(edge_numer_off, d->E + 1); OWN(edge_den_off, d->E + 1); OWN(sum_coef, d->NS); OWN(sum_fac_off, d->NS + 1);
  OWN(fac, d->NFAC); OWN(ie_p0, d->IE); OWN(ie_c0, d->IE); OWN(ie_c2, d->IE); OWN(ie_np_off, d->IE + 1);
  OWN(ie_np, d->NNP); OWN(ie_dp_off, d->IE + 1); OWN(ie_dp, d->NDP); OWN(sc_off, d->E + d->IE + 1);
  OWN(sc_kind, d->NSC); OWN(sc_a, d->NSC); OWN(sc_b, d->NSC); OWN(sc_coef, d->NSC); OWN(itv_is_cost, I);
  OWN(itv_cp_off, I + 1); OWN(cp_default, d->NCP); OWN(cp_min, d->NCP); OWN(cp_max, d->NCP);
  OWN(op_kind, d->NOP); OWN(op_itv, d->NOP); OWN(op_expr, d->NOP); OWN(op_arg, 8 * d->NOP);
  OWN(sel_arg, 4 * d->NSEL); OWN(expr_off, d->NEXPR + 1); OWN(tok_op, d->NTOK); OWN(tok_arg, d->NTOK);
  OWN(tok_f32, d->NTOK); OWN(tok_val, d->NTOK); OWN(prog_off, d->NPROG + 1); OWN(prog_step, I);
  OWN(prog_init, I); OWN(init_active, I); OWN(init_cpv, d->NCP); OWN(list_off, d->NLIST + 1);
  OWN(list_data, d->NLISTDATA);
#undef OWN
  o->CLG = C * L * G;
  o->n = (3 + C) * o->CLG + I * L;
  o->nmp = P * L * F * G; o->nft = L * F * G; o->nfi = L * F * G * G; o->nmob = G * d->MODE_NNZ_TOTAL; o->ncost = I * L;
  o->y = (double*)xcalloc(o->n, 8);
  o->Xprev = (double*)xcalloc(o->CLG, 8);
  o->X0 = (double*)d->X0;
  param_alloc(o, &o->yest); param_alloc(o, &o->init); param_alloc(o, &o->dest); param_alloc(o, &o->gap); param_alloc(o, &o->cur);
  o->dmp = (float*)xcalloc(o->nmp, 4); o->dft = (float*)xcalloc(o->nft, 4);
  o->dfi = (float*)xcalloc(o->nfi, 4); o->dmob = (float*)xcalloc(o->nmob, 4);
  for (int i = 0; i < 7; i++) o->K[i] = (double*)xcalloc(o->n, 8);
  o->ytmp = (double*)xcalloc(o->n, 8); o->ynew = (double*)xcalloc(o->n, 8);
  o->scale = (double*)xcalloc(o->n, 8); o->f1 = (double*)xcalloc(o->n, 8);
  o->alive = (double*)xcalloc(L * G, 8); o->foi = (double*)xcalloc((size_t)G * L * F, 8);
  o->edge = (double*)xcalloc(L * G, 8); o->next = (double*)xcalloc(o->CLG, 8);
  o->comb_csr = (float*)xcalloc((size_t)G * d->nnz, 4); o->comb_csc = (float*)xcalloc((size_t)G * d->nnz, 4);
  set_initial(o);
  return o;
}

void oracle_reset(Oracle* o) { /* Epidemic.reset (epidemic.py:104-106) */
  memset(o->y, 0, 8 * (size_t)o->n);
  memcpy(o->y, o->X0, 8 * (size_t)o->CLG);
  memcpy(o->Xprev, o->X0, 8 * (size_t)o->CLG);
  param_copy(o, &o->yest, &o->init);
  o->t = 0;
}

void oracle_destroy(Oracle* o) {
  if (!o) return;
  for (int i = 0; i < o->n_owned; i++) free(o->owned[i]);
  free(o->y); free(o->Xprev);
  param_free(&o->yest); param_free(&o->init); param_free(&o->dest); param_free(&o->gap); param_free(&o->cur);
  free(o->dmp); free(o->dft); free(o->dfi); free(o->dmob);
  for (int i = 0; i < 7; i++) free(o->K[i]);
  free(o->ytmp); free(o->ynew); free(o->scale); free(o->f1);
  free(o->alive); free(o->foi); free(o->edge); free(o->next); free(o->comb_csr); free(o->comb_csc);
  free(o);
}

/* Epidemic.step (epidemic.py:108-116) with full control-parameter vector + active flags.
 * Returns the day's reward. */
double oracle_day_step(Oracle* o, const float* cpv, const int32_t* active, const int32_t* prog_of) {
  const EpiDesc* d = &o->d;
  if (!prog_of) prog_of = d->prog_step;
  double* cur_cost = (double*)alloca(8 * (size_t)o->ncost);
  memcpy(cur_cost, o->y + (3 + d->C) * o->CLG, 8 * (size_t)o->ncost);
  double* Xcur = (double*)malloc(8 * (size_t)o->CLG);
  memcpy(Xcur, o->y, 8 * (size_t)o->CLG);
  execute(o, o->y, cpv, active, prog_of, o->dest.cost, &o->dest.mv);
  dest_parameter(o, &o->dest);
  gap_parameter(o, &o->yest, &o->dest, &o->gap);
  rk45_day(o);
  param_copy(o, &o->yest, &o->dest);
  /* previous state for value-mode 'change' = the state this day started from */
  memcpy(o->Xprev, Xcur, 8 * (size_t)o->CLG);
  free(Xcur);
  o->t++;
  double r = 0;
  for (int i = 0; i < o->ncost; i++) r += o->y[(3 + d->C) * o->CLG + i] - cur_cost[i];
  return -r;
}

/* EpiEnv.step (epipolicy_environment.py:39-62): expand action, n_days daily steps, summed reward */
double oracle_env_step(Oracle* o, const float* action, int n_days, int* done_out) {
  const EpiDesc* d = &o->d;
  float* cpv = (float*)alloca(4 * (size_t)(d->NCP ? d->NCP : 1));
  int32_t* active = (int32_t*)alloca(4 * (size_t)d->I);
  int a = 0;
  for (int i = 0; i < d->I; i++) {
    active[i] = 1; /* cost interventions are appended by get_normalized_action (obj/act.py:46-57) */
    for (int k = d->itv_cp_off[i]; k < d->itv_cp_off[i + 1]; k++) {
      if (d->itv_is_cost[i]) cpv[k] = d->cp_default[k];
      else if (d->cp_min[k] == d->cp_max[k]) cpv[k] = d->cp_min[k];
      else cpv[k] = action[a++];
    }
  }
  double total = 0;
  int done = 0;
  for (int k = 0; k < n_days; k++) {
    total += oracle_day_step(o, cpv, active, NULL);
    done = o->t >= d->horizon;
    if (done) break;
  }
  if (done_out) *done_out = done;
  return total;
}

void oracle_get_X(const Oracle* o, double* out) { memcpy(out, o->y, 8 * (size_t)o->CLG); }
void oracle_get_flat(const Oracle* o, double* out) { memcpy(out, o->y, 8 * (size_t)o->n); }
void oracle_set_flat(Oracle* o, const double* in) { memcpy(o->y, in, 8 * (size_t)o->n); }
void oracle_get_cost(const Oracle* o, double* out) {
  memcpy(out, o->y + (3 + o->d.C) * o->CLG, 8 * (size_t)o->ncost);
}
int oracle_n_flat(const Oracle* o) { return o->n; }
int oracle_day(const Oracle* o) { return o->t; }
void oracle_get_trace(const Oracle* o, int32_t* out4, double* last_err) {
  out4[0] = o->nfev; out4[1] = o->nsteps; out4[2] = o->nrej; out4[3] = o->status;
  if (last_err) *last_err = o->last_err;
}
int oracle_get_attempts(const Oracle* o, double* out, int cap) {
  for (int i = 0; i < o->n_attempts && i < cap; i++) { out[2 * i] = o->attempts[2 * i]; out[2 * i + 1] = o->attempts[2 * i + 1]; }
  return o->n_attempts;
}
/* which: 0 = yesterday (state.unobs), 1 = initial */
void oracle_get_param(const Oracle* o, int which, float* mp, float* ft, float* fi, float* mob, double* cost) {
  const Param* p = which ? &o->init : &o->yest;
  if (mp) memcpy(mp, p->mp, 4 * (size_t)o->nmp);
  if (ft) memcpy(ft, p->ft, 4 * (size_t)o->nft);
  if (fi) memcpy(fi, p->fi, 4 * (size_t)o->nfi);
  if (mob) memcpy(mob, p->mob, 4 * (size_t)o->nmob);
  if (cost) memcpy(cost, p->cost, 8 * (size_t)o->ncost);
}
int oracle_get_moves(const Oracle* o, int which, double* out6, int cap) {
  const Param* p = which ? &o->init : &o->yest;
  for (int i = 0; i < p->mv.n && i < cap; i++) {
    const Move* m = &p->mv.e[i];
    out6[6 * i + 0] = m->g; out6[6 * i + 1] = m->c1; out6[6 * i + 2] = m->c2;
    out6[6 * i + 3] = m->l1; out6[6 * i + 4] = m->l2; out6[6 * i + 5] = (double)m->v;
  }
  return p->mv.n;
}

/* Batch helper for the CPU baseline: run `n_envs` independent envs for `n_steps` gym steps each
 * (actions [n_envs, n_steps, A]) on the calling thread; returns the sum of rewards (a checksum). */
double oracle_rollout(const EpiDesc* desc, const float* actions, int n_envs, int n_steps, int n_days,
                      double* final_X /* [n_envs, CLG] or NULL */) {
  Oracle* o = oracle_create(desc);
  double acc = 0;
  for (int e = 0; e < n_envs; e++) {
    oracle_reset(o);
    for (int s = 0; s < n_steps; s++) {
      int done;
      acc += oracle_env_step(o, actions + ((size_t)e * n_steps + s) * desc->A, n_days, &done);
    }
    if (final_X) memcpy(final_X + (size_t)e * o->CLG, o->y, 8 * (size_t)o->CLG);
  }
  oracle_destroy(o);
  return acc;
}
