// epi_plan.h — the device-side "compiled scenario" shared by the host builder and the kernels.
//
// A DevPlan is built once per engine from an EpiDesc (include/epi_desc.h). It holds
//   * the reference's static tables, de-duplicated by LOCALE CLASS (locales whose default
//     parameters, facility matrices and op coverage are identical share one table row; the
//     shipped scenarios and the synthetic 1000-locale scenario all collapse to ONE class),
//   * multiplier classes: for every parameter / facility cell the ordered list of PARAM_MUL /
//     FI_MUL / FT_MUL ops that touch it (core/executor.py:116-164), so "apply the action" is a
//     handful of per-env scalars instead of a dense [P,L,F,G] delta tensor,
//   * infectious edges grouped by (susceptible compartment, transmission parameter, extra
//     parameters): the 16 COVID edges become 2 weighted force-of-infection problems
//     (core_optimize.py:98-163 recomputes the whole FOI matrix per edge),
//   * scatter maps turned into gather lists (no atomics),
//   * the per-env workspace layout (offsets) for the RK45 stages.
#pragma once
#include <stdint.h>

#define EPI_MAX_SLOTS 8
#define EPI_MAX_SEL 32

// shared-memory layout of the cell kernel (epi_cell.cuh): byte offsets inside one warp's block.
// Per-lane columns hold element i of lane ln at off + (i * W + ln) * sizeof(T); the per-env (slot)
// tables start at e_base + slot * e_bytes.
struct CellLayout {
  int enabled, epw;  // envs per warp = 32 / (L*G)
  int lw;            // lanes that carry a cell = epw * L*G = width of a per-lane column
  int roles;         // cooperating warps per CTA (epi_cellteam.cuh); 1 = the plain cell kernel
  int64_t l_X, l_ys, l_K, l_fl, l_FBE, l_accY, l_tdc, l_cmA, l_cmB, l_mvY, l_mvD, l_ysS, l_KS, l_red;
  int64_t e_base, e_bytes;
  int64_t e_mult_y, e_mult_d, e_cpv, e_act, e_live, e_elive, e_expr, e_sel, e_mpy, e_mpg, e_mpt0, e_mpFy, e_mpFg, e_coefS, e_Tnum,
      e_Tden, e_fi_y, e_fi_gap, e_ft_y, e_ft_gap, e_cost, e_crY, e_crD;
  int64_t e_mob_y, e_mob_g, e_mxr, e_mxc;  // mobility multipliers present: the env's mode matrices (y, gap) and combined mobility at time t
  int64_t warp_bytes;
};

// layout of the env kernel (epi_env.cuh): byte offsets of the [word][cell] columns inside the "big" and "comm"
// regions and of the per-env tables inside the table block; where each region lives.
struct EnvLayout {
  int enabled, threads, comm_in_smem, big_in_smem;
  int flow_slots;  // 1: the 7 stages' flows are stored (columns fls, accYn) instead of the flow sums (fl, FBE)
  int64_t l_fls, l_accYn;
  int64_t l_X, l_XF, l_ys, l_K, l_fl, l_FBE, l_accY, l_mvY, l_mvD, l_ysS, l_KS;  // big region
  int64_t l_cmA, l_cmB;                                                      // comm region
  int64_t e_mult_y, e_mult_d, e_cpv, e_act, e_expr, e_sel, e_mpy, e_mpg, e_mpt0, e_mpFy, e_mpFg, e_coefS, e_Tnum,
      e_Tden, e_fi_y, e_fi_gap, e_ft_y, e_ft_gap, e_cost, e_crY, e_crD;
  int64_t tab_bytes, red_bytes, comm_bytes, big_bytes;
};

// layout of the env-group kernel (epi_grp.cuh): one env is stepped by a GROUP of S co-resident CTAs, one cell per
// thread; byte offsets of the [word][thread] columns in shared memory, of the per-env tables, and of the group's
// global scratch (comm columns, float state copy, last-stage flows, candidate accumulators, reduction partials).
struct GrpLayout {
  int enabled, S, n_groups, threads;
  int R;    // CTAs per SM (2: CTAs of two different groups share an SM and fill each other's barrier / gather stalls)
  int lpc;  // locales per CTA (upper bound; CTA c owns locales [c*L/S, (c+1)*L/S))
  int np;   // doubles per CTA and parity in the reduction-partials array
  int trow; // row stride (floats) of the mixing tables T[lc][u0][v][u] (F*G padded against bank conflicts)
  int64_t s_K, s_KS, s_cmB, s_red, s_eci, s_eri, s_XF, s_tab, smem_bytes;
  int dyn_assign;  // R > 1 and the grid fills every SM exactly R times: CTAs pick their group by arrival order per SM
  int tden_smem;   // the denominator's mixing table has a copy in shared memory (else it is read from global memory)
  int xf_smem;  // the float copy of the state lives in shared memory (else in the group scratch)
  int nqc, nqr;  // longest CSC / CSR row, rounded up to the gather chunk: rows are padded to it (ELL layout per CTA)
  int64_t e_mult_y, e_mult_d, e_cpv, e_act, e_expr, e_sel, e_mpt0, e_coefS, e_Tnum, e_Tden, e_ft_y, e_ft_gap, e_selinfo, tab_bytes;
  int64_t g_XF, g_cmA, g_cmS, g_fl6, g_accN, g_mvD, g_crD, g_part, g_bytes;  // per group
  int64_t c_fi_y, c_fi_gap, c_mpy, c_mpg, c_mpFy, c_mpFg, c_Tden, c_evc, c_evr, c_bytes;  // per CTA (tables only the day prologue / a dyn day reads)
};

struct DevPlan {
  // ---- dimensions
  int C, L, G, F, P, I, A, NCP, LG, CLG, nnz, n_lc;
  int E;      // regular edges
  int NG;     // infectious groups
  int NSRC;   // E + NG + n_slot : rows of the flows buffer
  int n_acc;  // cumulative accumulators with a non-zero derivative (pairs / gain / lost)
  int n_slot; // move slots (distinct (c1,c2) of MOVE ops)
  int n_chg;  // compartments referenced by value-mode 'change' selects
  int n_sel, n_expr, n_cls, n_op;
  int n_flat; // length of the reference's flat ODE state (error-norm denominator)
  int horizon;
  int has_ft_ops; // FT_MUL ops present -> T_den depends on the env
  // ---- static tables (device pointers)
  const double* alive_coef;  // [C]
  const double* X0;          // [CLG]
  const int* lclass;         // [L]
  const float* mp0;          // [n_lc,P,F,G]
  const float* ft0;          // [n_lc,F,G] raw default
  const float* fi0;          // [n_lc,F,G,G] raw default
  const int* mp_cls;         // [n_lc,P,F,G] multiplier class (0 = identity)
  const int* ft_cls;         // [n_lc,F,G]
  const int* fi_cls;         // [n_lc,F,G,G]
  const float* Tden_const;   // [n_lc,F,G,G] ft*fi0*fi0^T with the default (normalised) ft, when !has_ft_ops
  const double* Tden_const64;
  const float* Tden_constT;  // the same, transposed [n_lc,F,u,u0] (env kernel)
  const float* Tden_constP;  // the same as rows [n_lc,u0][F*G + 4]: element (v,u) of cell group u0 at v*G+u (env-group kernel)
  const double* Tden_const64T;
  const float* mx_csr;       // [G,nnz] combined mobility on the union pattern (core_optimize.py:69-92)
  const float* mx_csc;       // [G,nnz]
  const float *mob_dense_csc, *mob_dense_csr; // [LG,L] dense column / row of every cell when L <= 4 (else null)
  const float *mx_csrT, *mx_cscT;             // [nnz,G] the same values, groups fastest (env-group kernel: a locale's lanes read one line)
  const int *csr_indptr, *csr_indices, *csc_indptr, *csc_indices;
  // ---- mobility multipliers (MOB_MUL ops, core/executor.py:121-127): the combined mobility then depends on the env
  // and the day; the mode matrices are rebuilt in the day prologue (matrix/coo.py:173-188) and combined per evaluation
  int mob_dyn, M, nmob;      // MOB_MUL ops present; modes; G * (sum of the modes' nnz)
  const float* mob_org;      // [nmob] static.mode_group_coo values, block m at G*mode_off[m], shape [G,nnz_m]
  const int* mob_cls;        // [nmob] multiplier class of every cell (0 = identity)
  const int *mode_indptr, *mode_indices, *mode_off;  // [M,L+1] [sum nnz_m] [M+1] the modes' CSR patterns
  const float* mode_bias;    // [M]
  const int* comb_src;       // [M,nnz] position of a union-pattern entry (CSR order) in mode m's pattern, or -1
  const int* csc_of_csr;     // [nnz] CSC position of a CSR entry
  // ---- multiplier classes
  const int* cls_off;        // [n_cls+1]
  const int* cls_op;         // op ids in execution order
  // ---- ops / programs / expressions (EpiDesc arrays, uploaded verbatim)
  const int *op_kind, *op_itv, *op_expr, *op_prog; // op_prog: program the op belongs to
  const int *prog_step, *prog_init;                // [I]
  const int *expr_off, *tok_op, *tok_arg, *tok_f32;
  const double* tok_val;
  const int* expr_itv;       // [n_expr] intervention owning the expression
  const int* expr_prog;      // [n_expr]
  // selects: byte masks
  const uint8_t *sel_c, *sel_l, *sel_g; // [n_sel,C] [n_sel,L] [n_sel,G]
  const int* sel_mode;                  // [n_sel]
  const int *selc_off, *selc;           // [n_sel+1], compartments of each select (the set bits of sel_c)
  const int *cr_off, *cr_ops;           // [I*L+1], ADD ops (indices into add_op) covering cost cell (i, l), in op order
  const int* chg_slot_of_comp;          // [C] index into Xprev rows or -1
  // MOVE ops (case "different compartment, same locale", executor.py:201-212)
  int n_mv_op;
  const int* mv_op;                     // [n_mv_op] op id
  const uint8_t *mv_g, *mv_l1;          // [n_mv_op,G] (groups in both selections) [n_mv_op,L]
  const int *mv_c1_off, *mv_c1, *mv_c2_off, *mv_c2; // per MOVE op compartment lists
  const int* slot_of_pair;              // [C*C] slot id or -1
  int slot_c1[EPI_MAX_SLOTS], slot_c2[EPI_MAX_SLOTS];
  const uint8_t* slot_cover;            // [n_slot,LG]
  // general move entries (a MOVE op with targets in other locales, executor.py:185-200,:213-233; team kernels only)
  int xmove, n_ent, n_src, n_tgt;
  int mv_len;                           // floats of a move table per env: n_slot*LG, or n_ent with xmove
  const int *ent_g, *ent_c1, *ent_c2, *ent_l1, *ent_l2, *ent_mo, *ent_sl;  // [n_ent] in the reference's insertion order
  const int *src_off, *src_mo, *src_idx, *src_sl;  // source cells: entries [src_off[s], src_off[s+1]), op, c1*LG+cell, slot with that c1
  const int* mo_src_off;                // [n_mv_op+1] the source cells of an op are contiguous
  const int *tgt_off, *tgt_ent, *tgt_idx, *tgt_sl;  // target cells: incoming entries in list order, c2*LG+cell, slot whose c1 is c2 (or -1)
  // ADD ops
  int n_add_op;
  const int* add_op;                    // [n_add_op]
  const uint8_t *add_i, *add_l;         // [n_add_op,I] [n_add_op,L]
  const int* add_parts;                 // [n_add_op]
  // control-parameter expansion (epipolicy_environment.py:40-47)
  const int* cp_src;                    // [NCP] action index or -1
  const float* cp_fixed;                // [NCP] value when cp_src < 0
  const float* init_cpv;                // [NCP]
  const uint8_t* init_active;           // [I]
  // ---- regular edges: interpreted programs
  const int* ep_off;                    // [E+1]
  const int* ep;                        // nNumer, facs..., nSummant, {coef_idx, nFac, facs...}
  const float* ep_coef;
  // ---- infectious groups
  const int *ig_c0, *ig_p0, *ig_np_off, *ig_np, *ig_dp_off, *ig_dp, *ig_w_off, *ig_w_c2;
  const double* ig_w;
  int n_igp;                            // distinct parameters used by infectious groups
  const int* comp_role;                 // [C] role owning a compartment in the cell-team kernel
  const int *igp, *ig_p0i, *ig_npi, *ig_dpi; // [n_igp] parameter ids; ig_p0 / ig_np / ig_dp as indices into igp
  // ---- gather lists
  const int *xg_off, *xg_src; const float* xg_coef;    // [C+1]
  const int *ag_off, *ag_src; const float* ag_coef;    // [n_acc+1]
  const int* acc_flat;                                  // [n_acc] offset/LG in the reference flat layout
  // ---- workspace layout (element offsets; see build_layout)
  int small_in_smem, big_in_smem;
  int64_t small_bytes, big_bytes;
  int64_t o_mult_y, o_mult_d, o_cpv, o_act, o_expr, o_sel, o_mpt0, o_coefS, o_Tnum, o_Tden, o_fi_y, o_fi_gap,
      o_ft_y, o_ft_gap, o_mob_y, o_mob_g, o_mxr, o_mxc;  // small region (bytes)
  int64_t o_X, o_ys, o_K, o_accY, o_accB, o_accE, o_accF, o_accF2, o_flows, o_alive, o_Xw, o_Ag, o_Wg, o_s, o_r,
      o_mvY, o_mvD, o_cost, o_crY, o_crD, o_ysS, o_KS, o_xnv;   // big region (bytes)
  CellLayout cl;
  EnvLayout el;
  GrpLayout gl;
  int has_src64;  // fp32 mode with move slots: the slots' source compartments are carried in double
};

// persistent per-env state in HBM (all [N, ...] arrays)
struct DevState {
  double* X;        // [N,CLG]            current_comp (double in both precision modes)
  void* acc;        // real [N,n_acc*LG]  live cumulative_move / gain / lost components
  double* cost;     // [N,I*L]            cumulative_cost
  double* crY;      // [N,I*L]            yesterday's cost rate
  float* mvY;       // [N,n_slot*LG]      yesterday's move table
  double* Xprev;    // [N,n_chg*LG]       previous state's compartments used by 'change' selects
  float* multY;     // [N,n_cls]          yesterday's class multipliers
  int* meta;        // [N,8]: t, ex_y bits, status, nfev, nsteps, nrej, (2 spare)
  double* last_err; // [N]
};
