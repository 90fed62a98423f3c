// epi_engine.cu — host side of libepi_b200.so: EpiDesc -> DevPlan compiler, state management, C-ABI.
//
// Replaces: the per-day work of Epidemic.get_next_state / step / reset
// (core/epidemic.py:82-116) and the sparse/dense helpers of epipolicy/matrix/* that only exist to
// feed it. Static construction (make_static, the parsers) stays in the reference; its output
// arrives here as an EpiDesc.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <dlfcn.h>

#include <algorithm>
#include <functional>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/epi_b200.h"
#include "epi_kernels.cuh"
#include "epi_cell.cuh"
#include "epi_env.cuh"

namespace {

thread_local std::string g_create_error;

struct Unsupported { std::string msg; };
struct Invalid { std::string msg; };
struct CudaFail { std::string msg; };

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) throw CudaFail{std::string(#call) + ": " + cudaGetErrorString(e_)};     \
  } while (0)

template <class T>
std::vector<T> vec(const T* p, size_t n) { return std::vector<T>(p, p + n); }

}  // namespace

// host copies of the plan tables the scenario-specialised code generator (gen_spec) walks
struct HostTabs {
  std::vector<double> alive_coef, ig_w;
  std::vector<int> ig_c0, ig_w_off, ig_w_c2, ep_off, ep, xg_off, xg_src, ag_off, ag_src;
  std::vector<float> ep_coef, xg_coef, ag_coef;
  std::vector<int> comp_role;
  std::vector<int> csr_indptr, csc_indptr;
  // per-locale-class default parameters and multiplier classes (epi_get_plan: the reporter rebuilds a day's parameters)
  std::vector<int> lclass, mp_cls, ft_cls;
  std::vector<float> mp0, ft0;
  // prologue tables for the generator
  std::vector<uint8_t> sel_c, sel_l, sel_g;
  std::vector<int> sel_mode, chg_slot, selc_off, selc, cls_off, cls_op, op_expr, expr_off, tok_op, tok_arg, tok_f32;
  std::vector<double> tok_val;
  std::vector<int> mv_op, mv_c1_off, mv_c1, mv_c2_off, mv_c2, slot_of_pair, cr_off, cr_ops, add_op, add_parts;
  std::vector<uint8_t> mv_g, mv_l1;
};

struct epi_engine {
  int device = 0, precision = 0, N = 0;
  HostTabs tabs;
  // scenario-specialised build of the cell kernel (epi_spec_attach)
  void* spec_lib = nullptr;
  int (*spec_launch)(const DevPlan*, const DevState*, const epi::StepArgs*, int, int, size_t, void*) = nullptr;
  // env-group kernel of the specialised build (epi_grp.cuh): cooperative launch of n_groups x S CTAs
  int (*spec_launch_grp)(const DevPlan*, const DevState*, const epi::StepArgs*, int, int, size_t, void*) = nullptr;
  char *grp_scratch = nullptr, *grp_cta_scratch = nullptr;
  unsigned* grp_bar = nullptr;
  int sms = 148;
  EpiDesc d{};                      // scalars only are valid after create
  DevPlan plan{};
  DevState S{}, init{};
  std::vector<void*> allocs;
  char *big_scratch = nullptr, *small_scratch = nullptr;
  double* cr_scratch = nullptr;
  int cta_mode = 0, threads = 128, wpb = 4, grid = 1;  // cta_mode: 0 warp team, 1 CTA team, 2 cell kernel
  int cellW = 32;                   // lanes per warp of the cell kernel (32 on the device; the host emulation shrinks it)
  size_t smem_bytes = 0, scratch_bytes = 0;
  int64_t launches = 0;
  // host-buffer path
  float* h_act = nullptr; float* d_act = nullptr;
  void* h_obs = nullptr; void* d_obs = nullptr;
  double* h_rew = nullptr; double* d_rew = nullptr;
  uint8_t* h_done = nullptr; uint8_t* d_done = nullptr;
  // device-side addresses of the pinned (mapped) host buffers: the kernel reads actions from / writes results to
  // host memory directly over PCIe (EPI_HOST_ZEROCOPY: bit 0 results, bit 1 actions)
  float* m_act = nullptr; void* m_obs = nullptr; double* m_rew = nullptr; uint8_t* m_done = nullptr;
  int zerocopy = 1;
  // epi_bind_mirror: second destination of every device-buffer step's results (peer memory of the learner rank)
  void* mir_obs = nullptr; double* mir_rew = nullptr; uint8_t* mir_done = nullptr;
  cudaStream_t stream = nullptr;
  mutable std::string err;
  double* dbg_attempts = nullptr;   // [N,64,2], allocated on first use of state selector 9
  std::vector<int> acc_flat_host;   // [n_acc]
  size_t real_size() const { return precision == EPI_FP64 ? 8 : 4; }

  bool host_only = false;           // plan compiled without a device (epi_spec_source_desc): tables stay on the host
  std::vector<void*> host_allocs;
  template <class T>
  T* up(const std::vector<T>& v) {
    T* p = nullptr;
    size_t n = v.size() ? v.size() : 1;
    if (host_only) {
      p = (T*)calloc(n, sizeof(T));
      if (!p) throw std::bad_alloc();
      host_allocs.push_back(p);
      if (v.size()) memcpy(p, v.data(), v.size() * sizeof(T));
      return p;
    }
    CK(cudaMalloc(&p, n * sizeof(T)));
    allocs.push_back(p);
    if (v.size()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
  template <class T>
  T* dalloc(size_t n, bool zero = true) {
    T* p = nullptr;
    CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    allocs.push_back(p);
    if (zero) CK(cudaMemset(p, 0, (n ? n : 1) * sizeof(T)));
    return p;
  }
};

namespace {

int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }

// ------------------------------------------------------------------------------------------
// EpiDesc -> DevPlan
// ------------------------------------------------------------------------------------------
void build_plan(epi_engine* h, const EpiDesc& d) {
  DevPlan& P = h->plan;
  const int C = d.C, L = d.L, G = d.G, F = d.F, PP = d.P, I = d.I, LG = L * G;
  if (C < 1 || L < 1 || G < 1 || F < 1 || I < 0) throw Invalid{"bad dimensions"};
  if (G > 32 || F > 32) throw Unsupported{"more than 32 groups or facilities"};
  P.C = C; P.L = L; P.G = G; P.F = F; P.P = PP; P.I = I; P.A = d.A; P.NCP = d.NCP; P.LG = LG; P.CLG = C * LG;
  P.nnz = d.nnz; P.E = d.E; P.horizon = d.horizon; P.n_op = d.NOP; P.n_expr = d.NEXPR; P.n_sel = d.NSEL;
  P.n_flat = (3 + C) * C * LG + I * L;

  auto list = [&](int id) { return vec(d.list_data + d.list_off[id], (size_t)(d.list_off[id + 1] - d.list_off[id])); };
  auto mask = [&](int id, int n) {
    std::vector<uint8_t> m(n, 0);
    for (int v : list(id)) { if (v < 0 || v >= n) throw Invalid{"list entry out of range"}; m[v] = 1; }
    return m;
  };

  // ---- op -> program, expression -> owner
  std::vector<int> op_prog(d.NOP, -1), expr_itv(d.NEXPR, 0), expr_prog(d.NEXPR, -1);
  for (int p = 0; p < d.NPROG; p++)
    for (int o = d.prog_off[p]; o < d.prog_off[p + 1]; o++) op_prog[o] = p;
  for (int o = 0; o < d.NOP; o++) {
    int e = d.op_expr[o];
    if (e < 0 || e >= d.NEXPR) throw Invalid{"op_expr out of range"};
    expr_itv[e] = d.op_itv[o];
    expr_prog[e] = op_prog[o];
  }
  for (int e = 0; e < d.NEXPR; e++) {
    int depth = 0, mx = 0;
    for (int k = d.expr_off[e]; k < d.expr_off[e + 1]; k++) {
      int op = d.tok_op[k];
      depth += op <= 2 ? 1 : (op == 7 ? 0 : -1);
      mx = depth > mx ? depth : mx;
    }
    if (mx > 16 || depth != 1) throw Unsupported{"expression too deep or malformed"};
  }

  // ---- locale classes: identical default tables and identical op coverage
  std::vector<int> lclass(L), rep;
  {
    std::map<std::string, int> seen;
    for (int l = 0; l < L; l++) {
      std::string key;
      for (int p = 0; p < PP; p++) key.append((const char*)(d.mp0 + ((size_t)p * L + l) * F * G), sizeof(float) * F * G);
      key.append((const char*)(d.ft0 + (size_t)l * F * G), sizeof(float) * F * G);
      key.append((const char*)(d.fi0 + (size_t)l * F * G * G), sizeof(float) * F * G * G);
      for (int o = 0; o < d.NOP; o++) {
        int k = d.op_kind[o];
        if (k == EPI_OP_PARAM_MUL) key.push_back((char)mask(d.op_arg[8 * o + 1], L)[l]);
        else if (k == EPI_OP_FI_MUL || k == EPI_OP_FT_MUL) key.push_back((char)mask(d.op_arg[8 * o + 0], L)[l]);
      }
      auto it = seen.find(key);
      if (it == seen.end()) { it = seen.emplace(key, (int)rep.size()).first; rep.push_back(l); }
      lclass[l] = it->second;
    }
  }
  const int n_lc = (int)rep.size();
  P.n_lc = n_lc;

  // ---- multiplier classes
  std::vector<int> mp_cls((size_t)n_lc * PP * F * G, 0), ft_cls((size_t)n_lc * F * G, 0), fi_cls((size_t)n_lc * F * G * G, 0);
  std::vector<float> mp0((size_t)n_lc * PP * F * G), ft0((size_t)n_lc * F * G), fi0((size_t)n_lc * F * G * G);
  std::map<std::vector<int>, int> cls_id;
  std::vector<std::vector<int>> cls_list;
  cls_id[{}] = 0; cls_list.push_back({});
  auto get_cls = [&](const std::vector<int>& ops) {
    auto it = cls_id.find(ops);
    if (it == cls_id.end()) { it = cls_id.emplace(ops, (int)cls_list.size()).first; cls_list.push_back(ops); }
    return it->second;
  };
  int has_ft_ops = 0, has_mob_ops = 0;
  std::vector<int> mob_cls;
  {
    // per op masks
    struct OpM { int kind; std::vector<uint8_t> a, l, f, g, g2; };
    std::vector<OpM> om(d.NOP);
    for (int o = 0; o < d.NOP; o++) {
      const int32_t* a = d.op_arg + 8 * o;
      om[o].kind = d.op_kind[o];
      if (om[o].kind == EPI_OP_PARAM_MUL) { om[o].a = mask(a[0], PP); om[o].l = mask(a[1], L); om[o].f = mask(a[2], F); om[o].g = mask(a[3], G); }
      else if (om[o].kind == EPI_OP_FI_MUL) { om[o].l = mask(a[0], L); om[o].f = mask(a[1], F); om[o].g = mask(a[2], G); om[o].g2 = mask(a[3], G); }
      else if (om[o].kind == EPI_OP_FT_MUL) { om[o].l = mask(a[0], L); om[o].f = mask(a[1], F); om[o].g = mask(a[2], G); has_ft_ops = 1; }
      else if (om[o].kind == EPI_OP_MOB_MUL) {
        // sim.apply({'locale-from', 'locale-to'}, m): CooMatrix.pairwise_multiply on the mode's pattern
        // (core/executor.py:121-127, matrix/coo.py:112-117): a = mode, g = groups, l = rows (from), f = columns (to)
        if (a[0] < 0 || a[0] >= d.M) throw Unsupported{"mobility multiplier on an unknown mode"};
        om[o].a = std::vector<uint8_t>(1, (uint8_t)a[0]); om[o].g = mask(a[1], G); om[o].l = mask(a[2], L); om[o].f = mask(a[3], L);
        has_mob_ops = 1;
      }
    }
    if (has_mob_ops) {
      // multiplier class of every mobility cell (mode, group, pattern entry): the MOB_MUL ops covering it, in op order
      mob_cls.assign((size_t)G * d.MODE_NNZ_TOTAL, 0);
      for (int m = 0; m < d.M; m++) {
        const int32_t* ip = d.mode_indptr + (size_t)m * (L + 1);
        const int32_t* ix = d.mode_indices + d.mode_off[m];
        const int nnz_m = d.mode_off[m + 1] - d.mode_off[m];
        for (int g = 0; g < G; g++) for (int r = 0; r < L; r++) for (int k = ip[r]; k < ip[r + 1]; k++) {
          std::vector<int> ops;
          for (int o = 0; o < d.NOP; o++)
            if (om[o].kind == EPI_OP_MOB_MUL && om[o].a[0] == m && om[o].g[g] && om[o].l[r] && om[o].f[ix[k]]) ops.push_back(o);
          mob_cls[(size_t)G * d.mode_off[m] + (size_t)g * nnz_m + k] = get_cls(ops);
        }
      }
    }
    for (int lc = 0; lc < n_lc; lc++) {
      int l = rep[lc];
      for (int p = 0; p < PP; p++) for (int f = 0; f < F; f++) for (int g = 0; g < G; g++) {
        std::vector<int> ops;
        for (int o = 0; o < d.NOP; o++)
          if (om[o].kind == EPI_OP_PARAM_MUL && om[o].a[p] && om[o].l[l] && om[o].f[f] && om[o].g[g]) ops.push_back(o);
        size_t idx = (((size_t)lc * PP + p) * F + f) * G + g;
        mp_cls[idx] = get_cls(ops);
        mp0[idx] = d.mp0[(((size_t)p * L + l) * F + f) * G + g];
      }
      for (int f = 0; f < F; f++) for (int g = 0; g < G; g++) {
        std::vector<int> ops;
        for (int o = 0; o < d.NOP; o++)
          if (om[o].kind == EPI_OP_FT_MUL && om[o].l[l] && om[o].f[f] && om[o].g[g]) ops.push_back(o);
        size_t idx = ((size_t)lc * F + f) * G + g;
        ft_cls[idx] = get_cls(ops);
        ft0[idx] = d.ft0[((size_t)l * F + f) * G + g];
        for (int g2 = 0; g2 < G; g2++) {
          std::vector<int> ops2;
          for (int o = 0; o < d.NOP; o++)
            if (om[o].kind == EPI_OP_FI_MUL && om[o].l[l] && om[o].f[f] && om[o].g[g] && om[o].g2[g2]) ops2.push_back(o);
          fi_cls[idx * G + g2] = get_cls(ops2);
          fi0[idx * G + g2] = d.fi0[(((size_t)l * F + f) * G + g) * G + g2];
        }
      }
    }
  }
  P.has_ft_ops = has_ft_ops;
  P.mob_dyn = has_mob_ops;
  P.M = d.M; P.nmob = G * d.MODE_NNZ_TOTAL;
  P.n_cls = (int)cls_list.size();
  std::vector<int> cls_off{0}, cls_op;
  for (auto& v : cls_list) { cls_op.insert(cls_op.end(), v.begin(), v.end()); cls_off.push_back((int)cls_op.size()); }

  // ---- default (normalised) facility_timespent and the constant FOI-denominator kernel
  std::vector<float> tden((size_t)n_lc * F * G * G);
  std::vector<double> tden64((size_t)n_lc * F * G * G);
  for (int lc = 0; lc < n_lc; lc++) {
    std::vector<float> ftn((size_t)F * G);
    for (int g = 0; g < G; g++) {
      float s = 0.0f;
      for (int f = 0; f < F; f++) s += ft0[((size_t)lc * F + f) * G + g];
      for (int f = 0; f < F; f++) ftn[(size_t)f * G + g] = s > 1e-15f ? ft0[((size_t)lc * F + f) * G + g] / s : (f == 0 ? 1.0f : 0.0f);
    }
    for (int v = 0; v < F; v++) for (int u0 = 0; u0 < G; u0++) for (int u = 0; u < G; u++) {
      size_t base = ((size_t)lc * F + v) * G;
      double t64 = (double)ftn[(size_t)v * G + u] * (double)fi0[(base + u0) * G + u] * (double)fi0[(base + u) * G + u0];
      tden64[(base + u0) * G + u] = t64;
      tden[(base + u0) * G + u] = (float)t64;
    }
  }

  // ---- constant combined mobility (no mobility ops): get_normalized_matrix (matrix/coo.py:173-188)
  // with all multipliers 1, then compute_combined_mobility (core_optimize.py:69-92) in f32
  std::vector<float> mx_csr((size_t)G * d.nnz, 0.0f), mx_csc((size_t)G * d.nnz, 0.0f);
  std::vector<int> comb_src, csc_of_csr;
  {
    std::vector<std::vector<float>> red(d.M);
    for (int m = 0; m < d.M; m++) {
      const int32_t* ip = d.mode_indptr + (size_t)m * (L + 1);
      int nnz_m = d.mode_off[m + 1] - d.mode_off[m];
      red[m].assign((size_t)G * nnz_m, 0.0f);
      for (int g = 0; g < G; g++) {
        const float* org = d.mode_origin + (size_t)G * d.mode_off[m] + (size_t)g * nnz_m;
        const int32_t* ix = d.mode_indices + d.mode_off[m];
        for (int r = 0; r < L; r++) {
          double osum = 0;
          for (int k = ip[r]; k < ip[r + 1]; k++) osum += org[k];
          for (int k = ip[r]; k < ip[r + 1]; k++)
            red[m][(size_t)g * nnz_m + k] = osum > 1e-15 ? (float)((double)org[k] / osum * osum) : (ix[k] == r ? (float)osum : 0.0f);
        }
      }
    }
    std::vector<int> cptr(d.csc_indptr, d.csc_indptr + L + 1), mi(d.M);
    comb_src.assign((size_t)d.M * d.nnz, -1); csc_of_csr.assign(d.nnz, 0);
    for (int r = 0; r < L; r++) {
      for (int m = 0; m < d.M; m++) mi[m] = d.mode_indptr[(size_t)m * (L + 1) + r];
      for (int k = d.csr_indptr[r]; k < d.csr_indptr[r + 1]; k++) {
        int c = d.csr_indices[k];
        int dest = cptr[c]++;
        csc_of_csr[k] = dest;
        for (int m = 0; m < d.M; m++) {
          const int32_t* ip = d.mode_indptr + (size_t)m * (L + 1);
          const int32_t* ix = d.mode_indices + d.mode_off[m];
          int nnz_m = d.mode_off[m + 1] - d.mode_off[m];
          if (mi[m] < ip[r + 1] && c == ix[mi[m]]) {
            for (int g = 0; g < G; g++) mx_csr[(size_t)g * d.nnz + k] += d.mode_bias[m] * red[m][(size_t)g * nnz_m + mi[m]];
            comb_src[(size_t)m * d.nnz + k] = mi[m];  // the union entry's position in mode m's pattern
            mi[m]++;
          }
        }
        for (int g = 0; g < G; g++) mx_csc[(size_t)g * d.nnz + dest] = mx_csr[(size_t)g * d.nnz + k];
      }
    }
  }

  // dense column / row of every cell for tiny locale counts (specialised cell kernel); only when the pattern's
  // indices are ascending, so that the dense loop adds the non-zeros in the reference's order
  std::vector<float> dense_csc, dense_csr;
  if (L <= 4 && !has_mob_ops) {  // (with mobility ops the values are per env and per day: the env's own tables)
    bool asc = true;
    for (int r = 0; r < L && asc; r++) {
      for (int k = d.csr_indptr[r] + 1; k < d.csr_indptr[r + 1]; k++) asc = asc && d.csr_indices[k - 1] < d.csr_indices[k];
      for (int k = d.csc_indptr[r] + 1; k < d.csc_indptr[r + 1]; k++) asc = asc && d.csc_indices[k - 1] < d.csc_indices[k];
    }
    if (asc) {
      dense_csc.assign((size_t)LG * L, 0.0f); dense_csr.assign((size_t)LG * L, 0.0f);
      for (int j = 0; j < L; j++) for (int g = 0; g < G; g++) {
        for (int k = d.csc_indptr[j]; k < d.csc_indptr[j + 1]; k++) dense_csc[((size_t)j * G + g) * L + d.csc_indices[k]] = mx_csc[(size_t)g * d.nnz + k];
        for (int k = d.csr_indptr[j]; k < d.csr_indptr[j + 1]; k++) dense_csr[((size_t)j * G + g) * L + d.csr_indices[k]] = mx_csr[(size_t)g * d.nnz + k];
      }
    }
  }

  // ---- selects
  std::vector<uint8_t> sel_c((size_t)d.NSEL * C), sel_l((size_t)d.NSEL * L), sel_g((size_t)d.NSEL * G);
  std::vector<int> sel_mode(d.NSEL), chg_slot(C, -1);
  int n_chg = 0;
  for (int s = 0; s < d.NSEL; s++) {
    auto mc = mask(d.sel_arg[4 * s + 0], C), ml = mask(d.sel_arg[4 * s + 1], L), mg = mask(d.sel_arg[4 * s + 2], G);
    std::copy(mc.begin(), mc.end(), sel_c.begin() + (size_t)s * C);
    std::copy(ml.begin(), ml.end(), sel_l.begin() + (size_t)s * L);
    std::copy(mg.begin(), mg.end(), sel_g.begin() + (size_t)s * G);
    sel_mode[s] = d.sel_arg[4 * s + 3];
    if (sel_mode[s] == EPI_SEL_CHANGE)
      for (int c = 0; c < C; c++) if (mc[c] && chg_slot[c] < 0) chg_slot[c] = n_chg++;
  }
  P.n_chg = n_chg;

  // ---- MOVE ops -> slots
  // Regular case (every shipped intervention): the move stays inside a locale ("different compartment, same locale",
  // executor.py:201-212) and is a per-cell table. A move whose targets lie in other locales (executor.py:185-200,
  // :213-233) switches the engine to GENERAL MOVE ENTRIES (xmove): a static list of (group, c1, c2, l1, l2) entries in
  // the reference's insertion order, per-env amounts per entry, applied source cell by source cell with the
  // reference's sequential clamp (core_optimize.py:196-208). Team kernels only.
  std::vector<int> mv_op, mv_c1_off{0}, mv_c1, mv_c2_off{0}, mv_c2, slot_of_pair((size_t)C * C, -1);
  std::vector<uint8_t> mv_g, mv_l1;
  std::vector<std::pair<int, int>> slots;
  bool xmove = false;
  for (int o = 0; o < d.NOP; o++) {
    if (d.op_kind[o] != EPI_OP_MOVE) continue;
    const int32_t* a = d.op_arg + 8 * o;
    auto c1 = list(a[0]), c2 = list(a[3]);
    auto l1 = mask(a[1], L), g1 = mask(a[2], G), l2 = mask(a[4], L), g2 = mask(a[5], G);
    for (int x : c1) for (int y : c2) if (x == y) xmove = true;         // targets in the same compartment: other locales
    for (int l = 0; l < L; l++) if (l1[l] && !l2[l]) xmove = true;     // a source locale that is not a target locale
    mv_op.push_back(o);
    for (int g = 0; g < G; g++) mv_g.push_back(g1[g] && g2[g]);
    mv_l1.insert(mv_l1.end(), l1.begin(), l1.end());
    mv_c1.insert(mv_c1.end(), c1.begin(), c1.end()); mv_c1_off.push_back((int)mv_c1.size());
    mv_c2.insert(mv_c2.end(), c2.begin(), c2.end()); mv_c2_off.push_back((int)mv_c2.size());
    for (int x : c1) for (int y : c2)
      if (slot_of_pair[(size_t)x * C + y] < 0) { slot_of_pair[(size_t)x * C + y] = (int)slots.size(); slots.push_back({x, y}); }
  }
  const int n_slot = (int)slots.size();
  if (n_slot > EPI_MAX_SLOTS) throw Unsupported{"more than 8 distinct move (c1,c2) pairs"};
  for (int a = 0; a < n_slot; a++) for (int b = 0; b < n_slot; b++) {
    if (a != b && slots[a].first == slots[b].first) throw Unsupported{"two moves out of one compartment (order-dependent clamp)"};
    if ((a != b || !xmove) && slots[a].second == slots[b].first) throw Unsupported{"a move target that is another move's source (order-dependent clamp)"};
  }
  // general move entries
  std::vector<int> ent_g, ent_c1, ent_c2, ent_l1, ent_l2, ent_mo, ent_sl, src_off{0}, src_mo, src_idx, src_sl, mo_src_off{0},
      tgt_off{0}, tgt_ent, tgt_idx, tgt_sl;
  if (xmove) {
    if (mv_op.size() > 16) throw Unsupported{"more than 16 move ops with cross-locale moves"};
    for (size_t mo = 0; mo < mv_op.size(); mo++) {
      const int32_t* a = d.op_arg + 8 * mv_op[mo];
      auto c1s = list(a[0]), l1s = list(a[1]), g1s = list(a[2]), c2s = list(a[3]), l2s = list(a[4]);
      auto c2m = mask(a[3], C), l2m = mask(a[4], L), g2m = mask(a[5], G);
      for (int g : g1s) {
        if (!g2m[g]) continue;
        for (int c1 : c1s) for (int l1 : l1s) {
          // one source cell (executor.py:178-183: it counts in the departing population whether or not it has targets)
          src_mo.push_back((int)mo); src_idx.push_back(c1 * LG + l1 * G + g);
          auto add = [&](int c2, int l2) {
            ent_g.push_back(g); ent_c1.push_back(c1); ent_c2.push_back(c2); ent_l1.push_back(l1); ent_l2.push_back(l2);
            ent_mo.push_back((int)mo); ent_sl.push_back(slot_of_pair[(size_t)c1 * C + c2]);
          };
          if (c2m[c1]) { if (!l2m[l1]) for (int l2 : l2s) add(c1, l2); }           // :185-200
          else if (l2m[l1]) { for (int c2 : c2s) add(c2, l1); }                     // :201-212
          else { for (int c2 : c2s) for (int l2 : l2s) add(c2, l2); }              // :213-233
          src_off.push_back((int)ent_g.size());
        }
      }
      mo_src_off.push_back((int)src_mo.size());
    }
    if (ent_g.size() > (size_t)1 << 22) throw Unsupported{"more than 4 M cross-locale move entries"};
    // order independence between ops (the reference applies the entries in dict-insertion order, which depends on the
    // history of live ops): among ops that can be live together (same program variant), every entry comes from one op,
    // a source cell with entries belongs to one op, and no target cell is a source cell with entries
    for (int variant = 0; variant < 2; variant++) {
      auto in_set = [&](int mo) { int o = mv_op[mo], itv = d.op_itv[o]; return op_prog[o] == (variant ? d.prog_init[itv] : d.prog_step[itv]); };
      std::map<std::vector<int>, int> owner;
      std::map<int, int> src_owner;  // (compartment, cell) with outgoing entries -> op
      for (size_t s_ = 0; s_ < src_mo.size(); s_++) {
        if (!in_set(src_mo[s_]) || src_off[s_ + 1] == src_off[s_]) continue;
        auto it = src_owner.find(src_idx[s_]);
        if (it != src_owner.end() && it->second != src_mo[s_]) throw Unsupported{"two move ops out of one cell (order-dependent clamp)"};
        src_owner[src_idx[s_]] = src_mo[s_];
      }
      for (size_t e = 0; e < ent_g.size(); e++) {
        if (!in_set(ent_mo[e])) continue;
        std::vector<int> key{ent_g[e], ent_c1[e], ent_c2[e], ent_l1[e], ent_l2[e]};
        if (owner.count(key)) throw Unsupported{"two move ops feeding one (compartment, locale) pair (order-dependent sum)"};
        owner[key] = ent_mo[e];
        if (src_owner.count(ent_c2[e] * LG + ent_l2[e] * G + ent_g[e])) throw Unsupported{"a move target that is another move's source (order-dependent clamp)"};
      }
    }
    for (size_t s_ = 0; s_ < src_mo.size(); s_++) {
      int c1 = src_idx[s_] / LG, sl = -1;
      for (int q = 0; q < n_slot; q++) if (slots[q].first == c1) sl = q;
      src_sl.push_back(sl);
    }
    // target cells: their incoming entries in list order
    std::map<int, std::vector<int>> by_tgt;
    for (size_t e = 0; e < ent_g.size(); e++) by_tgt[ent_c2[e] * LG + ent_l2[e] * G + ent_g[e]].push_back((int)e);
    for (auto& kv : by_tgt) {
      tgt_idx.push_back(kv.first);
      int c2 = kv.first / LG, sl = -1;
      for (int q = 0; q < n_slot; q++) if (slots[q].first == c2) sl = q;  // the target's compartment is carried in double too
      tgt_sl.push_back(sl);
      tgt_ent.insert(tgt_ent.end(), kv.second.begin(), kv.second.end());
      tgt_off.push_back((int)tgt_ent.size());
    }
  }
  P.xmove = xmove ? 1 : 0; P.n_ent = (int)ent_g.size(); P.n_src = (int)src_mo.size(); P.n_tgt = (int)tgt_idx.size();
  P.mv_len = xmove ? (P.n_ent > 0 ? P.n_ent : 1) : n_slot * LG;
  P.n_slot = n_slot; P.n_mv_op = (int)mv_op.size();
  std::vector<uint8_t> slot_cover((size_t)n_slot * LG, 0);
  for (size_t mo = 0; mo < mv_op.size(); mo++)
    for (int q1 = mv_c1_off[mo]; q1 < mv_c1_off[mo + 1]; q1++) for (int q2 = mv_c2_off[mo]; q2 < mv_c2_off[mo + 1]; q2++) {
      int sl = slot_of_pair[(size_t)mv_c1[q1] * C + mv_c2[q2]];
      for (int l = 0; l < L; l++) for (int g = 0; g < G; g++)
        if (mv_l1[mo * L + l] && mv_g[mo * G + g]) slot_cover[(size_t)sl * LG + l * G + g] = 1;
    }
  for (int s = 0; s < n_slot; s++) { P.slot_c1[s] = slots[s].first; P.slot_c2[s] = slots[s].second; }

  // ---- ADD ops
  std::vector<int> add_op, add_parts;
  std::vector<uint8_t> add_i, add_l;
  for (int o = 0; o < d.NOP; o++) {
    if (d.op_kind[o] != EPI_OP_ADD) continue;
    const int32_t* a = d.op_arg + 8 * o;
    auto mi = mask(a[0], I), ml = mask(a[1], L);
    add_op.push_back(o);
    add_parts.push_back((int)(list(a[0]).size() * list(a[1]).size()));
    add_i.insert(add_i.end(), mi.begin(), mi.end());
    add_l.insert(add_l.end(), ml.begin(), ml.end());
  }
  P.n_add_op = (int)add_op.size();
  // per cost cell (i, l): the ADD ops that cover it; per select: its compartments
  std::vector<int> cr_off{0}, cr_ops, selc_off{0}, selc;
  for (int i = 0; i < I; i++) for (int l = 0; l < L; l++) {
    for (size_t ao = 0; ao < add_op.size(); ao++)
      if (add_i[ao * I + i] && add_l[ao * L + l]) cr_ops.push_back((int)ao);
    cr_off.push_back((int)cr_ops.size());
  }
  for (int s = 0; s < d.NSEL; s++) {
    for (int c = 0; c < C; c++) if (sel_c[(size_t)s * C + c]) selc.push_back(c);
    selc_off.push_back((int)selc.size());
  }

  // ---- control-parameter expansion (epipolicy_environment.py:24-31,:40-47)
  std::vector<int> cp_src(d.NCP, -1);
  std::vector<float> cp_fixed(d.NCP, 0.0f);
  std::vector<uint8_t> init_active(I);
  int n_act = 0;
  for (int i = 0; i < I; i++) {
    init_active[i] = (uint8_t)(d.init_active[i] != 0);
    for (int k = d.itv_cp_off[i]; k < d.itv_cp_off[i + 1]; k++) {
      if (d.itv_is_cost[i]) cp_fixed[k] = d.cp_default[k];
      else if (d.cp_min[k] == d.cp_max[k]) cp_fixed[k] = d.cp_min[k];
      else cp_src[k] = n_act++;
    }
  }
  if (n_act != d.A) throw Invalid{"EpiDesc.A does not match the number of free control parameters"};

  // ---- regular edge programs
  std::vector<int> ep_off{0}, ep;
  for (int e = 0; e < d.E; e++) {
    ep.push_back(d.edge_numer_off[e + 1] - d.edge_numer_off[e]);
    for (int k = d.edge_numer_off[e]; k < d.edge_numer_off[e + 1]; k++) ep.push_back(d.fac[k]);
    ep.push_back(d.edge_den_off[e + 1] - d.edge_den_off[e]);
    for (int s = d.edge_den_off[e]; s < d.edge_den_off[e + 1]; s++) {
      ep.push_back(s);
      ep.push_back(d.sum_fac_off[s + 1] - d.sum_fac_off[s]);
      for (int k = d.sum_fac_off[s]; k < d.sum_fac_off[s + 1]; k++) ep.push_back(d.fac[k]);
    }
    ep_off.push_back((int)ep.size());
  }

  // ---- infectious groups: edges with the same susceptible side and proportional scatter are one
  // weighted force-of-infection problem (everything downstream of the FOI numerator is linear in X[c2])
  struct Grp { int c0, p0; std::vector<int> np, dp; int ref_edge; std::vector<int> c2; std::vector<double> w; };
  std::vector<Grp> grps;
  for (int e = 0; e < d.IE; e++) {
    std::vector<int> np = vec(d.ie_np + d.ie_np_off[e], (size_t)(d.ie_np_off[e + 1] - d.ie_np_off[e]));
    std::vector<int> dp = vec(d.ie_dp + d.ie_dp_off[e], (size_t)(d.ie_dp_off[e + 1] - d.ie_dp_off[e]));
    int s0 = d.sc_off[d.E + e], s1 = d.sc_off[d.E + e + 1];
    bool placed = false;
    for (auto& g : grps) {
      if (g.c0 != d.ie_c0[e] || g.p0 != d.ie_p0[e] || g.np != np || g.dp != dp) continue;
      int r0 = d.sc_off[d.E + g.ref_edge], r1 = d.sc_off[d.E + g.ref_edge + 1];
      if (r1 - r0 != s1 - s0 || s1 == s0 || d.sc_coef[r0] == 0.0f) continue;
      double ratio = (double)d.sc_coef[s0] / (double)d.sc_coef[r0];
      bool ok = true;
      for (int k = 0; k < s1 - s0 && ok; k++) {
        ok = d.sc_kind[s0 + k] == d.sc_kind[r0 + k] && d.sc_a[s0 + k] == d.sc_a[r0 + k] && d.sc_b[s0 + k] == d.sc_b[r0 + k] &&
             fabs((double)d.sc_coef[s0 + k] - ratio * (double)d.sc_coef[r0 + k]) <= 1e-13 * fabs((double)d.sc_coef[s0 + k]);
      }
      if (!ok) continue;
      g.c2.push_back(d.ie_c2[e]); g.w.push_back(ratio); placed = true;
      break;
    }
    if (!placed) grps.push_back(Grp{d.ie_c0[e], d.ie_p0[e], np, dp, e, {d.ie_c2[e]}, {1.0}});
  }
  const int NG = (int)grps.size();
  if (NG > EPI_MAX_NG) throw Unsupported{"more than 8 infectious-edge groups"};
  P.NG = NG; P.NSRC = d.E + NG + n_slot;
  std::vector<int> ig_c0, ig_p0, ig_np_off{0}, ig_np, ig_dp_off{0}, ig_dp, ig_w_off{0}, ig_w_c2;
  std::vector<double> ig_w;
  for (auto& g : grps) {
    ig_c0.push_back(g.c0); ig_p0.push_back(g.p0);
    ig_np.insert(ig_np.end(), g.np.begin(), g.np.end()); ig_np_off.push_back((int)ig_np.size());
    ig_dp.insert(ig_dp.end(), g.dp.begin(), g.dp.end()); ig_dp_off.push_back((int)ig_dp.size());
    ig_w_c2.insert(ig_w_c2.end(), g.c2.begin(), g.c2.end()); ig_w.insert(ig_w.end(), g.w.begin(), g.w.end());
    ig_w_off.push_back((int)ig_w.size());
  }

  // distinct parameters the infectious groups read (the cell kernel keeps their per-env (y, gap) tables)
  std::vector<int> igp, ig_p0i, ig_npi, ig_dpi;
  {
    auto idx_of = [&](int p) {
      for (size_t i = 0; i < igp.size(); i++) if (igp[i] == p) return (int)i;
      igp.push_back(p);
      return (int)igp.size() - 1;
    };
    for (int p : ig_p0) ig_p0i.push_back(idx_of(p));
    for (int p : ig_np) ig_npi.push_back(idx_of(p));
    for (int p : ig_dp) ig_dpi.push_back(idx_of(p));
  }
  P.n_igp = (int)igp.size();

  // ---- gather lists
  std::vector<std::vector<std::pair<int, float>>> xg(C);
  std::map<int, int> acc_of_flat;  // flat row (in LG units) -> acc id
  std::vector<std::vector<std::pair<int, float>>> ag;
  std::vector<int> acc_flat;
  auto acc_id = [&](int flat_row) {
    auto it = acc_of_flat.find(flat_row);
    if (it == acc_of_flat.end()) { it = acc_of_flat.emplace(flat_row, (int)ag.size()).first; ag.emplace_back(); acc_flat.push_back(flat_row); }
    return it->second;
  };
  auto add_scatter = [&](int src, int s0, int s1) {
    for (int k = s0; k < s1; k++) {
      float coef = d.sc_coef[k];
      int a = d.sc_a[k], b = d.sc_b[k];
      if (a < 0 || a >= C) throw Invalid{"scatter compartment out of range"};
      switch (d.sc_kind[k]) {
        case EPI_SC_COMP: xg[a].push_back({src, coef}); break;
        case EPI_SC_GAIN: ag[acc_id(C + a)].push_back({src, coef}); break;
        case EPI_SC_LOST: ag[acc_id(2 * C + a)].push_back({src, coef}); break;
        default: ag[acc_id(3 * C + a * C + b)].push_back({src, coef}); break;
      }
    }
  };
  for (int e = 0; e < d.E; e++) add_scatter(e, d.sc_off[e], d.sc_off[e + 1]);
  for (int k = 0; k < NG; k++) add_scatter(d.E + k, d.sc_off[d.E + grps[k].ref_edge], d.sc_off[d.E + grps[k].ref_edge + 1]);
  for (int s = 0; s < n_slot; s++) ag[acc_id(3 * C + slots[s].first * C + slots[s].second)].push_back({d.E + NG + s, 1.0f});
  P.n_acc = (int)ag.size();
  h->acc_flat_host = acc_flat;
  std::vector<int> xg_off{0}, xg_src, ag_off{0}, ag_src;
  std::vector<float> xg_coef, ag_coef;
  for (auto& v : xg) { for (auto& pr : v) { xg_src.push_back(pr.first); xg_coef.push_back(pr.second); } xg_off.push_back((int)xg_src.size()); }
  for (auto& v : ag) { for (auto& pr : v) { ag_src.push_back(pr.first); ag_coef.push_back(pr.second); } ag_off.push_back((int)ag_src.size()); }

  // (a role-split "cell-team" variant - R warps per env group, epi_cellteam.cuh in round 1 - measured 2.1-2.3 ms per
  // COVID_C gym step against the cell kernel's 1.24 ms at the time and was removed in round 2)
  {
    P.cl.roles = 1;
    std::vector<int> comp_role(C, 0);
    h->tabs.comp_role = comp_role;
    P.comp_role = h->up(comp_role);
  }
  {
    HostTabs& T = h->tabs;
    T.alive_coef = vec(d.alive_coef, C); T.ig_w = ig_w; T.ig_c0 = ig_c0; T.ig_w_off = ig_w_off; T.ig_w_c2 = ig_w_c2;
    T.ep_off = ep_off; T.ep = ep; T.ep_coef = vec(d.sum_coef, d.NS);
    T.xg_off = xg_off; T.xg_src = xg_src; T.xg_coef = xg_coef; T.ag_off = ag_off; T.ag_src = ag_src; T.ag_coef = ag_coef;
    T.sel_c = sel_c; T.sel_l = sel_l; T.sel_g = sel_g; T.sel_mode = sel_mode; T.chg_slot = chg_slot;
    T.selc_off = selc_off; T.selc = selc; T.cls_off = cls_off; T.cls_op = cls_op;
    T.op_expr = vec(d.op_expr, d.NOP); T.expr_off = vec(d.expr_off, d.NEXPR + 1);
    T.tok_op = vec(d.tok_op, d.NTOK); T.tok_arg = vec(d.tok_arg, d.NTOK); T.tok_f32 = vec(d.tok_f32, d.NTOK);
    T.tok_val = vec(d.tok_val, d.NTOK);
    T.mv_op = mv_op; T.mv_c1_off = mv_c1_off; T.mv_c1 = mv_c1; T.mv_c2_off = mv_c2_off; T.mv_c2 = mv_c2;
    T.slot_of_pair = slot_of_pair; T.mv_g = mv_g; T.mv_l1 = mv_l1;
    T.cr_off = cr_off; T.cr_ops = cr_ops; T.add_op = add_op; T.add_parts = add_parts;
  }

  // ---- workspace layout
  const int64_t rs = (int64_t)h->real_size();
  {
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align16(o + bytes); return r; };
    P.o_mult_y = take(4 * (int64_t)P.n_cls); P.o_mult_d = take(4 * (int64_t)P.n_cls);
    P.o_cpv = take(4 * (int64_t)d.NCP); P.o_act = take(I);
    P.o_expr = take(8 * (int64_t)d.NEXPR); P.o_sel = take(8 * (int64_t)d.NSEL);
    P.o_mpt0 = take(4 * (int64_t)n_lc * PP * G); P.o_coefS = take(rs * (int64_t)NG * n_lc * F * G);
    P.o_Tnum = take(4 * (int64_t)n_lc * F * G * G); P.o_Tden = take(has_ft_ops ? rs * (int64_t)n_lc * F * G * G : 0);
    P.o_fi_y = take(4 * (int64_t)n_lc * F * G * G); P.o_fi_gap = take(4 * (int64_t)n_lc * F * G * G);
    P.o_ft_y = take(4 * (int64_t)n_lc * F * G); P.o_ft_gap = take(4 * (int64_t)n_lc * F * G);
    P.o_mob_y = take(has_mob_ops ? 4 * (int64_t)P.nmob : 0); P.o_mob_g = take(has_mob_ops ? 4 * (int64_t)P.nmob : 0);
    P.o_mxr = take(has_mob_ops ? 4 * (int64_t)G * d.nnz : 0); P.o_mxc = take(has_mob_ops ? 4 * (int64_t)G * d.nnz : 0);
    P.small_bytes = o;
    o = 0;
    const int64_t CLG = P.CLG, nA = (int64_t)P.n_acc * LG;
    P.o_X = take(8 * CLG); P.o_ys = take(rs * CLG); P.o_K = take(rs * 7 * CLG);
    P.o_accY = take(rs * nA); P.o_accB = take(rs * nA); P.o_accE = take(rs * nA); P.o_accF = take(rs * nA); P.o_accF2 = take(rs * nA);
    P.o_flows = take(rs * (int64_t)P.NSRC * LG);
    P.o_alive = take(rs * LG); P.o_Xw = take(rs * (int64_t)NG * LG); P.o_Ag = take(rs * LG); P.o_Wg = take(rs * (int64_t)NG * LG);
    P.o_s = take(rs * (int64_t)NG * LG); P.o_r = take(rs * (int64_t)NG * LG);
    P.o_mvY = take(4 * (int64_t)P.mv_len); P.o_mvD = take(4 * (int64_t)P.mv_len); P.o_xnv = take(xmove ? 8 * (int64_t)P.mv_len : 0);
    P.o_cost = take(8 * (int64_t)I * L); P.o_crY = take(8 * (int64_t)I * L); P.o_crD = take(8 * (int64_t)I * L);
    P.has_src64 = (rs == 4 && n_slot > 0) ? 1 : 0;
    P.o_ysS = take(P.has_src64 ? 8 * (int64_t)n_slot * LG : 0); P.o_KS = take(P.has_src64 ? 8 * 7 * (int64_t)n_slot * LG : 0);
    P.big_bytes = o;
  }

  // ---- upload
  std::vector<uint8_t> dummy8(1, 0);
  P.alive_coef = h->up(vec(d.alive_coef, C)); P.X0 = h->up(vec(d.X0, (size_t)C * LG)); P.lclass = h->up(lclass);
  P.mp0 = h->up(mp0); P.ft0 = h->up(ft0); P.fi0 = h->up(fi0); P.mp_cls = h->up(mp_cls); P.ft_cls = h->up(ft_cls); P.fi_cls = h->up(fi_cls);
  h->tabs.lclass = lclass; h->tabs.mp_cls = mp_cls; h->tabs.ft_cls = ft_cls; h->tabs.mp0 = mp0; h->tabs.ft0 = ft0;
  {
    std::vector<float> tdT(tden.size());
    std::vector<double> td64T(tden64.size());
    for (int lv = 0; lv < n_lc * F; lv++) for (int u0 = 0; u0 < G; u0++) for (int u = 0; u < G; u++) {
      tdT[((size_t)lv * G + u) * G + u0] = tden[((size_t)lv * G + u0) * G + u];
      td64T[((size_t)lv * G + u) * G + u0] = tden64[((size_t)lv * G + u0) * G + u];
    }
    P.Tden_constT = h->up(tdT); P.Tden_const64T = h->up(td64T);
    // padded rows for the env-group kernel: row (lc, u0) holds element (v, u) at v*G + u, F*G + 4 floats per row
    const size_t trow = (size_t)F * G + 4;
    std::vector<float> tdP((size_t)n_lc * G * trow, 0.0f);
    for (int lc = 0; lc < n_lc; lc++) for (int v = 0; v < F; v++) for (int u0 = 0; u0 < G; u0++) for (int u = 0; u < G; u++)
      tdP[((size_t)lc * G + u0) * trow + (size_t)v * G + u] = tden[((((size_t)lc * F + v) * G) + u0) * G + u];
    P.Tden_constP = h->up(tdP);
  }
  P.Tden_const = h->up(tden); P.Tden_const64 = h->up(tden64); P.mx_csr = h->up(mx_csr); P.mx_csc = h->up(mx_csc);
  {
    std::vector<float> tr((size_t)G * d.nnz), tc((size_t)G * d.nnz);
    for (int g = 0; g < G; g++)
      for (int k = 0; k < d.nnz; k++) { tr[(size_t)k * G + g] = mx_csr[(size_t)g * d.nnz + k]; tc[(size_t)k * G + g] = mx_csc[(size_t)g * d.nnz + k]; }
    P.mx_csrT = h->up(tr); P.mx_cscT = h->up(tc);
    h->tabs.csr_indptr = vec(d.csr_indptr, (size_t)L + 1); h->tabs.csc_indptr = vec(d.csc_indptr, (size_t)L + 1);
  }
  P.mob_dense_csc = dense_csc.empty() ? nullptr : h->up(dense_csc);
  P.mob_dense_csr = dense_csr.empty() ? nullptr : h->up(dense_csr);
  P.mob_org = h->up(vec(d.mode_origin, (size_t)G * d.MODE_NNZ_TOTAL)); P.mob_cls = h->up(has_mob_ops ? mob_cls : std::vector<int>(1, 0));
  P.mode_indptr = h->up(vec(d.mode_indptr, (size_t)d.M * (L + 1))); P.mode_indices = h->up(vec(d.mode_indices, d.MODE_NNZ_TOTAL));
  P.mode_off = h->up(vec(d.mode_off, d.M + 1)); P.mode_bias = h->up(vec(d.mode_bias, d.M));
  P.comb_src = h->up(comb_src); P.csc_of_csr = h->up(csc_of_csr);
  P.csr_indptr = h->up(vec(d.csr_indptr, L + 1)); P.csr_indices = h->up(vec(d.csr_indices, d.nnz));
  P.csc_indptr = h->up(vec(d.csc_indptr, L + 1)); P.csc_indices = h->up(vec(d.csc_indices, d.nnz));
  P.cls_off = h->up(cls_off); P.cls_op = h->up(cls_op);
  P.op_kind = h->up(vec(d.op_kind, d.NOP)); P.op_itv = h->up(vec(d.op_itv, d.NOP)); P.op_expr = h->up(vec(d.op_expr, d.NOP));
  P.op_prog = h->up(op_prog); P.prog_step = h->up(vec(d.prog_step, I)); P.prog_init = h->up(vec(d.prog_init, I));
  P.expr_off = h->up(vec(d.expr_off, d.NEXPR + 1)); P.tok_op = h->up(vec(d.tok_op, d.NTOK)); P.tok_arg = h->up(vec(d.tok_arg, d.NTOK));
  P.tok_f32 = h->up(vec(d.tok_f32, d.NTOK)); P.tok_val = h->up(vec(d.tok_val, d.NTOK));
  P.expr_itv = h->up(expr_itv); P.expr_prog = h->up(expr_prog);
  P.sel_c = h->up(sel_c); P.sel_l = h->up(sel_l); P.sel_g = h->up(sel_g); P.sel_mode = h->up(sel_mode); P.chg_slot_of_comp = h->up(chg_slot);
  P.mv_op = h->up(mv_op); P.mv_g = h->up(mv_g); P.mv_l1 = h->up(mv_l1); P.mv_c1_off = h->up(mv_c1_off); P.mv_c1 = h->up(mv_c1);
  P.mv_c2_off = h->up(mv_c2_off); P.mv_c2 = h->up(mv_c2); P.slot_of_pair = h->up(slot_of_pair); P.slot_cover = h->up(slot_cover);
  P.ent_g = h->up(ent_g); P.ent_c1 = h->up(ent_c1); P.ent_c2 = h->up(ent_c2); P.ent_l1 = h->up(ent_l1); P.ent_l2 = h->up(ent_l2);
  P.ent_mo = h->up(ent_mo); P.ent_sl = h->up(ent_sl); P.src_off = h->up(src_off); P.src_mo = h->up(src_mo); P.src_idx = h->up(src_idx);
  P.src_sl = h->up(src_sl); P.mo_src_off = h->up(mo_src_off); P.tgt_off = h->up(tgt_off); P.tgt_ent = h->up(tgt_ent);
  P.tgt_idx = h->up(tgt_idx); P.tgt_sl = h->up(tgt_sl);
  P.cr_off = h->up(cr_off); P.cr_ops = h->up(cr_ops); P.selc_off = h->up(selc_off); P.selc = h->up(selc);
  P.add_op = h->up(add_op); P.add_i = h->up(add_i); P.add_l = h->up(add_l); P.add_parts = h->up(add_parts);
  P.cp_src = h->up(cp_src); P.cp_fixed = h->up(cp_fixed); P.init_cpv = h->up(vec(d.init_cpv, d.NCP)); P.init_active = h->up(init_active);
  P.ep_off = h->up(ep_off); P.ep = h->up(ep); P.ep_coef = h->up(vec(d.sum_coef, d.NS));
  P.ig_c0 = h->up(ig_c0); P.ig_p0 = h->up(ig_p0); P.ig_np_off = h->up(ig_np_off); P.ig_np = h->up(ig_np);
  P.ig_dp_off = h->up(ig_dp_off); P.ig_dp = h->up(ig_dp); P.ig_w_off = h->up(ig_w_off); P.ig_w_c2 = h->up(ig_w_c2); P.ig_w = h->up(ig_w);
  P.igp = h->up(igp); P.ig_p0i = h->up(ig_p0i); P.ig_npi = h->up(ig_npi); P.ig_dpi = h->up(ig_dpi);
  P.xg_off = h->up(xg_off); P.xg_src = h->up(xg_src); P.xg_coef = h->up(xg_coef);
  P.ag_off = h->up(ag_off); P.ag_src = h->up(ag_src); P.ag_coef = h->up(ag_coef); P.acc_flat = h->up(acc_flat);
}

// shared-memory layout of the cell kernel for warps of W lanes (epi_cell.cuh)
void build_cell_layout(epi_engine* h, int W) {
  DevPlan& P = h->plan;
  CellLayout& Y = P.cl;
  const int64_t rs = (int64_t)h->real_size();
  const int64_t C = P.C, G = P.G, F = P.F, n_lc = P.n_lc, NG = P.NG, nC = (int64_t)P.I * P.L;
  const int roles = Y.roles > 0 ? Y.roles : 1;
  Y = CellLayout{};
  Y.roles = roles;
  Y.epw = W / P.LG;
  Y.lw = W = Y.epw * P.LG;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = align16(o + bytes); return r; };
  const int64_t nS = P.has_src64 ? P.n_slot : 0;
  Y.l_X = take(8 * C * W); Y.l_ys = take(rs * C * W); Y.l_K = take(rs * 8 * C * W);  // 8 stage slots: two 4-vectors
  Y.l_fl = take(rs * P.NSRC * W); Y.l_FBE = take(2 * rs * P.NSRC * W); Y.l_accY = take(rs * (int64_t)P.n_acc * W);
  Y.l_tdc = take(P.has_ft_ops ? 0 : rs * n_lc * F * G * G);
  Y.l_cmA = take(rs * (1 + NG) * W); Y.l_cmB = take(rs * (1 + NG) * W);
  Y.l_mvY = take(4 * (int64_t)P.n_slot * W); Y.l_mvD = take(4 * (int64_t)P.n_slot * W);
  Y.l_ysS = take(8 * nS * W); Y.l_KS = take(8 * 7 * nS * W); Y.l_red = take(8 * (int64_t)W * roles);
  Y.e_base = o;
  o = 0;
  Y.e_mult_y = take(4 * (int64_t)P.n_cls); Y.e_mult_d = take(4 * (int64_t)P.n_cls);
  Y.e_cpv = take(4 * (int64_t)P.NCP); Y.e_act = take(P.I);
  Y.e_live = take(P.n_op); Y.e_elive = take(P.n_expr);
  Y.e_expr = take(8 * (int64_t)P.n_expr); Y.e_sel = take(8 * (int64_t)P.n_sel);
  Y.e_mpy = take(4 * n_lc * P.P * G); Y.e_mpg = take(4 * n_lc * P.P * G); Y.e_mpt0 = take(4 * n_lc * P.P * G);
  Y.e_mpFy = take(4 * (int64_t)P.n_igp * n_lc * F * G); Y.e_mpFg = take(4 * (int64_t)P.n_igp * n_lc * F * G);
  Y.e_coefS = take(rs * NG * n_lc * F * G); Y.e_Tnum = take(4 * n_lc * F * G * G);
  Y.e_Tden = take(P.has_ft_ops ? rs * n_lc * F * G * G : 0);
  Y.e_fi_y = take(4 * n_lc * F * G * G); Y.e_fi_gap = take(4 * n_lc * F * G * G);
  Y.e_ft_y = take(4 * n_lc * F * G); Y.e_ft_gap = take(4 * n_lc * F * G);
  Y.e_cost = take(8 * nC); Y.e_crY = take(8 * nC); Y.e_crD = take(8 * nC);
  Y.e_mob_y = take(P.mob_dyn ? 4 * (int64_t)P.nmob : 0); Y.e_mob_g = take(P.mob_dyn ? 4 * (int64_t)P.nmob : 0);
  Y.e_mxr = take(P.mob_dyn ? 4 * G * P.nnz : 0); Y.e_mxc = take(P.mob_dyn ? 4 * G * P.nnz : 0);
  Y.e_bytes = o;
  Y.warp_bytes = Y.e_base + Y.e_bytes * Y.epw;
}


int env_threads(int LG) {
  if (const char* ov = getenv("EPI_ENV_THREADS")) { int v = atoi(ov); if (v >= 32 && v <= 256 && v % 32 == 0) return v; }  // tuning aid
  return LG >= 256 ? 256 : ((LG + 31) / 32) * 32;
}

// layout of the env kernel (epi_env.cuh) for CTAs of `threads` threads
void build_env_layout(epi_engine* h, int threads, int64_t smem_max) {
  DevPlan& P = h->plan;
  EnvLayout& Y = P.el;
  const int64_t rs = (int64_t)h->real_size();
  const int64_t C = P.C, G = P.G, F = P.F, n_lc = P.n_lc, NG = P.NG, nC = (int64_t)P.I * P.L, LG = P.LG;
  Y = EnvLayout{};
  Y.threads = threads;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = align16(o + bytes); return r; };
  const int64_t nS = P.has_src64 ? P.n_slot : 0;
  // first with the flow sums (fl + packed FB/FE): does the env fit in shared memory? if not, the global-workspace
  // regime stores the 7 stages' flows instead (less traffic per evaluation, more workspace)
  for (int pass = 0; pass < 2; pass++) {
    o = 0;
    Y.l_X = take(8 * C * LG); Y.l_XF = take(rs == 4 ? 4 * C * LG : 0); Y.l_ys = take(rs * C * LG); Y.l_K = take(rs * 7 * C * LG);
    if (Y.flow_slots) {
      Y.l_fls = take(rs * 7 * P.NSRC * LG); Y.l_accYn = take(rs * (int64_t)P.n_acc * LG);
      Y.l_fl = Y.l_FBE = Y.l_fls;
    } else {
      Y.l_fl = take(rs * P.NSRC * LG); Y.l_FBE = take(2 * rs * P.NSRC * LG);
      Y.l_fls = Y.l_accYn = Y.l_fl;
    }
    Y.l_accY = take(rs * (int64_t)P.n_acc * LG);
    Y.l_mvY = take(4 * (int64_t)P.n_slot * LG); Y.l_mvD = take(4 * (int64_t)P.n_slot * LG);
    Y.l_ysS = take(8 * nS * LG); Y.l_KS = take(8 * 7 * nS * LG);
    Y.big_bytes = o;
    if (pass == 0) {
      // (tables + reduction scratch + comm columns are laid out below; a conservative bound decides the regime)
      int64_t small = 48 * 1024;
      if (getenv("EPI_ENV_FLOWSLOTS")) Y.flow_slots = atoi(getenv("EPI_ENV_FLOWSLOTS")) != 0;
      else Y.flow_slots = Y.big_bytes + small > smem_max;
      if (!Y.flow_slots) break;
    }
  }
  o = 0;
  Y.l_cmA = take(rs * (1 + NG) * LG); Y.l_cmB = take(rs * (1 + NG) * LG);
  Y.comm_bytes = o;
  o = 0;
  Y.e_mult_y = take(4 * (int64_t)P.n_cls); Y.e_mult_d = take(4 * (int64_t)P.n_cls);
  Y.e_cpv = take(4 * (int64_t)P.NCP); Y.e_act = take(P.I);
  Y.e_expr = take(8 * (int64_t)P.n_expr); Y.e_sel = take(8 * (int64_t)P.n_sel);
  Y.e_mpy = take(4 * n_lc * P.P * G); Y.e_mpg = take(4 * n_lc * P.P * G); Y.e_mpt0 = take(4 * n_lc * P.P * G);
  Y.e_mpFy = take(4 * (int64_t)P.n_igp * n_lc * F * G); Y.e_mpFg = take(4 * (int64_t)P.n_igp * n_lc * F * G);
  Y.e_coefS = take(rs * NG * n_lc * F * G); Y.e_Tnum = take(4 * n_lc * F * G * G);
  Y.e_Tden = take(P.has_ft_ops ? rs * n_lc * F * G * G : 0);
  Y.e_fi_y = take(4 * n_lc * F * G * G); Y.e_fi_gap = take(4 * n_lc * F * G * G);
  Y.e_ft_y = take(4 * n_lc * F * G); Y.e_ft_gap = take(4 * n_lc * F * G);
  Y.e_cost = Y.e_crY = Y.e_crD = o;  // the cost arrays of the env kernel live in global memory
  Y.tab_bytes = o;
  Y.red_bytes = align16(8 * (int64_t)threads);
  int64_t used = Y.tab_bytes + Y.red_bytes;
  Y.comm_in_smem = used + Y.comm_bytes <= smem_max;
  if (Y.comm_in_smem) used += Y.comm_bytes;
  Y.big_in_smem = Y.comm_in_smem && used + Y.big_bytes <= smem_max;
}

// layout of the env-group kernel (epi_grp.cuh): the fewest CTAs per env whose cells fit one per thread with the
// stage derivatives in shared memory; then as many groups as the device has room for (all CTAs must be co-resident)
void build_grp_layout(epi_engine* h, int sms, int64_t smem_max) {
  DevPlan& P = h->plan;
  GrpLayout& Y = P.gl;
  Y = GrpLayout{};
  const char* ev = getenv("EPI_GRP");  // "0": never; "2": also when the env fits one CTA's shared memory (tests)
  const int mode = ev ? atoi(ev) : 1;
  if (mode == 0 || h->precision != EPI_FP32 || P.LG <= 32 || P.G > 32 || P.F > 32 || P.C > 32 || P.n_sel > 32 || P.n_mv_op > 32) return;
  if (P.el.big_in_smem && mode < 2) return;
  const int64_t rs = 4, C = P.C, G = P.G, F = P.F, n_lc = P.n_lc, NG = P.NG;
  const int64_t nS = P.has_src64 ? P.n_slot : 0;
  const int forceS = getenv("EPI_GRP_S") ? atoi(getenv("EPI_GRP_S")) : 0;
  const int forceR = getenv("EPI_GRP_R") ? atoi(getenv("EPI_GRP_R")) : 0;
  const int maxw = getenv("EPI_GRP_WARPS") ? atoi(getenv("EPI_GRP_WARPS")) : 16;  // resident warps per SM
  // shared-memory layout of a CTA for groups of S CTAs, R CTAs per SM; false if it does not fit
  auto layout = [&](int S, int R) {
    const int lpc = (P.L + S - 1) / S;
    const int threads = (int)(((int64_t)lpc * G + 31) / 32 * 32);
    // registers: the SM's warps are dealt to four 16 K-register partitions; 16 warps per SM leave 128 registers per
    // thread, 12 warps 168 (the kernel keeps 70 flow sums per cell in registers)
    if (R * (threads / 32) > maxw) return false;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align16(o + bytes); return r; };
    Y.trow = (int)(F * G + 4);
    Y.np = P.n_sel + P.n_mv_op > 2 ? P.n_sel + P.n_mv_op : 2;
    Y.s_K = take(rs * 6 * C * threads); Y.s_KS = take(8 * 6 * nS * threads); Y.s_cmB = take(rs * (1 + NG) * threads);
    Y.s_red = take(8 * (32 + (int64_t)(P.n_mv_op > 2 ? P.n_mv_op : 2)) + 16);  // + (chunks of this CTA's longest CSC / CSR row)
    // the CSC / CSR rows of the CTA's locales, padded to the longest row (ELL): source / destination locale per
    // (entry, locale) in shared memory, the mobility value per (entry, thread) in the CTA's global scratch
    {
      int mc = 0, mr = 0;
      for (int l = 0; l < P.L; l++) {
        mc = std::max(mc, h->tabs.csc_indptr[l + 1] - h->tabs.csc_indptr[l]);
        mr = std::max(mr, h->tabs.csr_indptr[l + 1] - h->tabs.csr_indptr[l]);
      }
      const int ch = R * (threads / 32) <= 12 ? 8 : 4;  // EPI_GRP_CHUNK (epi_grp.cuh)
      Y.nqc = (mc + ch - 1) / ch * ch; Y.nqr = (mr + ch - 1) / ch * ch;
    }
    Y.s_eci = take(4 * (int64_t)lpc * Y.nqc); Y.s_eri = take(4 * (int64_t)lpc * Y.nqr);
    const int64_t o_cols = o;
    o = 0;
    Y.e_mult_y = take(4 * (int64_t)P.n_cls); Y.e_mult_d = take(4 * (int64_t)P.n_cls);
    Y.e_cpv = take(4 * (int64_t)P.NCP); Y.e_act = take(P.I);
    Y.e_expr = take(8 * (int64_t)P.n_expr); Y.e_sel = take(8 * (int64_t)P.n_sel);
    Y.e_mpt0 = take(4 * n_lc * P.P * G);
    Y.e_coefS = take(rs * NG * n_lc * F * G);
    Y.e_Tnum = take(4 * n_lc * G * Y.trow);
    Y.e_ft_y = take(4 * n_lc * F * G); Y.e_ft_gap = take(4 * n_lc * F * G);
    Y.e_selinfo = take(8 * (int64_t)(P.n_sel > 0 ? P.n_sel : 1));  // (compartment mask, mode) per select
    Y.tab_bytes = o;
    // when there is room: first a shared-memory copy of the denominator's mixing table (read 2 F G / 4 times per cell
    // and evaluation), then the float copy of the state (read C times)
    const int64_t td_bytes = align16(rs * n_lc * G * Y.trow);
    Y.tden_smem = R * (o_cols + Y.tab_bytes + td_bytes + 1024) <= smem_max + 1024 && !getenv("EPI_GRP_TDEN_GLOBAL");
    Y.e_Tden = Y.tab_bytes;
    if (Y.tden_smem) Y.tab_bytes += td_bytes;
    Y.s_XF = o_cols;
    const int64_t xf_bytes = align16(rs * C * threads);
    Y.xf_smem = R * (o_cols + xf_bytes + Y.tab_bytes + 1024) <= smem_max + 1024 && !getenv("EPI_GRP_XF_GLOBAL");
    Y.s_tab = o_cols + (Y.xf_smem ? xf_bytes : 0);
    Y.smem_bytes = Y.s_tab + Y.tab_bytes;
    Y.lpc = lpc; Y.threads = threads; Y.S = S; Y.R = R;
    // R CTAs share the SM's shared memory (1 KB of it is reserved per resident CTA)
    return R * (Y.smem_bytes + 1024) <= smem_max + 1024;
  };
  int S = 0, Rsel = 0;
  // (two CTAs of different groups per SM were measured and do not pay: the step is a latency chain per env, and the
  // number of envs in flight is set by the shared memory either way; EPI_GRP_R=2 keeps the variant reachable)
  for (int R = forceR > 0 ? forceR : 1; R >= 1 && S == 0; R--) {
    if (forceS > 0) {
      if (forceS <= P.L && layout(forceS, R)) { S = forceS; Rsel = R; }
    } else {
      for (int s_ = 1; s_ <= sms * R && s_ <= P.L; s_++)
        if (layout(s_, R)) { S = s_; Rsel = R; break; }
      if (S > 0) {
        // as many groups as fit at this size, then all the CTA slots' worth of CTAs per group (fewer cells per CTA)
        const int ng = sms * R / S;
        int S2 = sms * R / ng;
        if (S2 > P.L) S2 = P.L;
        if (!layout(S2, R)) layout(S, R);
      }
    }
    if (forceR > 0) break;
  }
  if (S == 0) { Y = GrpLayout{}; return; }
  (void)Rsel;
  Y.n_groups = sms * Y.R / Y.S;
  Y.dyn_assign = Y.R > 1 && Y.n_groups * Y.S == sms * Y.R;
  const int64_t LG = P.LG;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = (o + bytes + 255) & ~int64_t(255); return r; };
  Y.g_XF = take(rs * C * LG); Y.g_cmA = take(rs * (1 + NG) * LG); Y.g_cmS = take(rs * (NG > 0 ? NG : 1) * LG);
  Y.g_fl6 = take(rs * P.NSRC * LG); Y.g_accN = take(rs * (int64_t)P.n_acc * LG); Y.g_mvD = take(4 * (int64_t)P.n_slot * LG);
  Y.g_crD = take(8 * (int64_t)P.I * P.L); Y.g_part = take(8 * 2 * (int64_t)Y.S * Y.np);
  Y.g_bytes = o;
  o = 0;
  Y.c_fi_y = take(4 * n_lc * F * G * G); Y.c_fi_gap = take(4 * n_lc * F * G * G);
  Y.c_mpy = take(4 * n_lc * P.P * G); Y.c_mpg = take(4 * n_lc * P.P * G);
  Y.c_mpFy = take(4 * (int64_t)P.n_igp * n_lc * F * G); Y.c_mpFg = take(4 * (int64_t)P.n_igp * n_lc * F * G);
  Y.c_Tden = take(P.has_ft_ops ? rs * n_lc * G * Y.trow : 0);
  Y.c_evc = take(4 * (int64_t)Y.nqc * Y.threads); Y.c_evr = take(4 * (int64_t)Y.nqr * Y.threads);
  Y.c_bytes = o;
  Y.enabled = 1;
}

// ------------------------------------------------------------------------------------------
// scenario-specialised code generator: the per-cell program of the cell kernel as straight-line CUDA
// (same operations in the same order as the interpreted path of epi_cell.cuh)
// ------------------------------------------------------------------------------------------
std::string hexf(float v) { char b[64]; snprintf(b, sizeof b, "(real)%af", (double)v); return b; }
std::string hexd(double v) { char b[64]; snprintf(b, sizeof b, "(real)%a", v); return b; }

std::string gen_spec(const epi_engine* h) {
  const DevPlan& P = h->plan;
  const HostTabs& T = h->tabs;
  std::ostringstream o;
  o << "#include \"epi_fdiv.cuh\"\nnamespace epi { namespace spec {\n";
  o << "typedef " << (h->precision == EPI_FP64 ? "double" : "float") << " real;\n";
  o << "constexpr int C = " << P.C << ", L = " << P.L << ", G = " << P.G << ", F = " << P.F << ", P = " << P.P << ", I = " << P.I
    << ", LG = " << P.LG << ", CLG = " << P.CLG << ", nnz = " << P.nnz << ", n_lc = " << P.n_lc << ", E = " << P.E
    << ", NG = " << P.NG << ", NSRC = " << P.NSRC << ", n_acc = " << P.n_acc << ", n_slot = " << P.n_slot
    << ", has_ft_ops = " << P.has_ft_ops << ", has_src64 = " << P.has_src64
    << ", dense_mob = " << (P.mob_dense_csc ? 1 : 0) << ", mob_dyn = " << P.mob_dyn << ", flow_slots = " << (P.LG > 32 ? P.el.flow_slots : 0) << ", R = " << (P.cl.roles > 0 ? P.cl.roles : 1) << ";\n";
  o << "constexpr int big = " << (P.LG > 32 ? 1 : 0) << ";  // 1: env kernel (epi_env.cuh), 0: cell kernel (epi_cell.cuh)\n";
  if (P.LG <= 32) {
    const CellLayout& Y = P.cl;  // built for 32-lane warps
    o << "constexpr int l_X = " << Y.l_X << ", l_ys = " << Y.l_ys << ", l_K = " << Y.l_K << ", l_fl = " << Y.l_fl << ", l_FBE = " << Y.l_FBE
      << ", l_accY = " << Y.l_accY << ", l_tdc = " << Y.l_tdc << ", LW = " << Y.lw << ", l_cmA = " << Y.l_cmA << ", l_cmB = " << Y.l_cmB << ", l_mvY = " << Y.l_mvY << ", l_mvD = " << Y.l_mvD
      << ", l_ysS = " << Y.l_ysS << ", l_KS = " << Y.l_KS << ", l_red = " << Y.l_red << ";\n";
    o << "constexpr int e_base = " << Y.e_base << ", e_bytes = " << Y.e_bytes << ", e_mult_y = " << Y.e_mult_y << ", e_mult_d = " << Y.e_mult_d
      << ", e_cpv = " << Y.e_cpv << ", e_act = " << Y.e_act << ", e_live = " << Y.e_live << ", e_elive = " << Y.e_elive << ", e_expr = " << Y.e_expr << ", e_sel = " << Y.e_sel << ", e_mpy = " << Y.e_mpy
      << ", e_mpg = " << Y.e_mpg << ", e_mpt0 = " << Y.e_mpt0 << ", e_mpFy = " << Y.e_mpFy << ", e_mpFg = " << Y.e_mpFg
      << ", e_coefS = " << Y.e_coefS << ", e_Tnum = " << Y.e_Tnum << ", e_Tden = " << Y.e_Tden << ", e_fi_y = " << Y.e_fi_y
      << ", e_fi_gap = " << Y.e_fi_gap << ", e_ft_y = " << Y.e_ft_y << ", e_ft_gap = " << Y.e_ft_gap << ", e_cost = " << Y.e_cost
      << ", e_crY = " << Y.e_crY << ", e_crD = " << Y.e_crD << ", e_mob_y = " << Y.e_mob_y << ", e_mob_g = " << Y.e_mob_g
      << ", e_mxr = " << Y.e_mxr << ", e_mxc = " << Y.e_mxc << ", warp_bytes = " << Y.warp_bytes << ";\n";
  }
  auto arr = [&](const char* name, const std::vector<int>& v) {
    o << "__device__ constexpr int " << name << "[" << (v.empty() ? 1 : v.size()) << "] = {";
    if (v.empty()) o << "0";
    for (size_t i = 0; i < v.size(); i++) o << (i ? ", " : "") << v[i];
    o << "};\n";
  };
  std::vector<int> s1(P.slot_c1, P.slot_c1 + P.n_slot), s2(P.slot_c2, P.slot_c2 + P.n_slot);
  arr("slot_c1", s1); arr("slot_c2", s2); arr("ig_c0", T.ig_c0); arr("comp_role", T.comp_role);
  // phase A
  o << "__device__ __forceinline__ void phaseA(const real* y, real& alive, real* xw) {\n  alive = 0;\n";
  for (int c = 0; c < P.C; c++)
    if (T.alive_coef[c] != 0.0) o << "  alive += y[" << c << "] * " << hexd(T.alive_coef[c]) << ";\n";
  for (int k = 0; k < P.NG; k++) {
    o << "  xw[" << k << "] = 0;\n";
    for (int q = T.ig_w_off[k]; q < T.ig_w_off[k + 1]; q++)
      o << "  xw[" << k << "] += " << hexd(T.ig_w[q]) << " * y[" << T.ig_w_c2[q] << "];\n";
  }
  o << "}\n";

  // denominator of a regular edge: `dn > 0 ? n / dn : n` (compute_fraction, core_optimize.py:55-65). A denominator
  // made of constants only is folded here (the usual case is the single summant "1").
  const std::map<std::string, std::string>* rcp_of = nullptr;
  auto emit_denominator = [&](std::ostringstream& out, const int*& ep, const std::function<std::string(int)>& valfn,
                              const std::string& target) {
    int ns = *ep++;
    if (ns == 0) { out << " " << target << " = n; }\n"; return; }
    bool all_const = true;
    double csum = 0;
    {
      const int* q = ep;
      for (int sidx = 0; sidx < ns; sidx++) { float cf = T.ep_coef[*q++]; int nf = *q++; if (nf) all_const = false; q += nf; csum += (double)cf; }
    }
    if (all_const) {
      for (int sidx = 0; sidx < ns; sidx++) { ep++; ep++; }
      float c = 0.0f;  // summed in f32 like `denom += sv` on real=float; exact for the single-summant case
      { const int* q = ep - 2 * ns; for (int sidx = 0; sidx < ns; sidx++) { c += T.ep_coef[*q]; q += 2; } }
      (void)csum;
      if (c == 1.0f) out << " " << target << " = n; }\n";
      else if (c > 0) out << " " << target << " = fdiv(n, " << hexf(c) << "); }\n";
      else out << " " << target << " = n; }\n";
      return;
    }
    std::ostringstream dn;
    for (int sidx = 0; sidx < ns; sidx++) {
      dn << " dn += " << hexf(T.ep_coef[*ep++]);
      int nf = *ep++;
      for (int q = 0; q < nf; q++) dn << " * " << valfn(*ep++);
      dn << ";";
    }
    out << " real dn = 0;" << dn.str();
    if (h->precision == EPI_FP32 && rcp_of) {
      // fp32 mode: edges sharing a denominator share its reciprocal (declared by the caller from rcp_of)
      auto it = rcp_of->find(dn.str());
      if (it != rcp_of->end()) { out << " " << target << " = dn > 0 ? n * " << it->second << " : n; }\n"; return; }
    }
    out << " " << target << " = dn > 0 ? fdiv(n, dn) : n; }\n";
  };
  // denominators (as generated text) that occur in more than one regular edge -> a shared reciprocal variable
  auto shared_denominators = [&](const std::function<std::string(int)>& valfn, std::vector<std::pair<std::string, std::string>>& decl) {
    std::map<std::string, int> count;
    std::vector<std::string> order;
    for (int e = 0; e < P.E; e++) {
      const int* ep = T.ep.data() + T.ep_off[e];
      int n = *ep++; ep += n;
      int ns = *ep++;
      bool all_const = true;
      std::ostringstream dn;
      for (int sidx = 0; sidx < ns; sidx++) {
        dn << " dn += " << hexf(T.ep_coef[*ep++]);
        int nf = *ep++;
        if (nf) all_const = false;
        for (int q = 0; q < nf; q++) dn << " * " << valfn(*ep++);
        dn << ";";
      }
      if (ns == 0 || all_const) continue;
      if (!count[dn.str()]++) order.push_back(dn.str());
    }
    std::map<std::string, std::string> m;
    for (auto& k : order)
      if (count[k] > 1) { std::string v = "rcp" + std::to_string(m.size()); m[k] = v; decl.push_back({v, k}); }
    return m;
  };

  // regular edges
  o << "__device__ __forceinline__ void flows(const real* y, real alive, const float* mp, real* f) {\n";
  auto val = [&](int fac) {
    int kind = fac >> 24, idx = fac & 0xffffff;
    std::ostringstream v;
    if (kind == 0) v << "(real)mp[" << idx * P.G << "]";
    else if (kind == 1) v << "alive";
    else v << "y[" << idx << "]";
    return v.str();
  };
  std::vector<std::pair<std::string, std::string>> rdecl;
  std::map<std::string, std::string> rmap;
  if (h->precision == EPI_FP32) {
    rmap = shared_denominators(val, rdecl);
    for (auto& dc : rdecl) o << "  real " << dc.first << "; { real dn = 0;" << dc.second << " " << dc.first << " = frcp(dn); }\n";
    rcp_of = &rmap;
  }
  for (int e = 0; e < P.E; e++) {
    const int* ep = T.ep.data() + T.ep_off[e];
    int n = *ep++;
    o << "  { real n = (real)1";
    for (int q = 0; q < n; q++) o << " * " << val(*ep++);
    o << ";";
    emit_denominator(o, ep, val, "f[" + std::to_string(e) + "]");
  }
  rcp_of = nullptr;
  o << "}\n";
  // scatter into dX
  o << "__device__ __forceinline__ void dx(const real* f, real* d) {\n";
  for (int c = 0; c < P.C; c++) {
    o << "  d[" << c << "] = 0;";
    for (int q = T.xg_off[c]; q < T.xg_off[c + 1]; q++) o << " d[" << c << "] += f[" << T.xg_src[q] << "] * " << hexf(T.xg_coef[q]) << ";";
    o << "\n";
  }
  o << "}\n";
  // accumulator derivatives of a flow column (col already offset by the lane; stride 32)
  o << "template <int STRIDE> __device__ __forceinline__ void acc(const real* col, real* out) {\n";
  for (int a = 0; a < P.n_acc; a++) {
    o << "  out[" << a << "] = 0;";
    for (int q = T.ag_off[a]; q < T.ag_off[a + 1]; q++) o << " out[" << a << "] += col[" << T.ag_src[q] << " * STRIDE] * " << hexf(T.ag_coef[q]) << ";";
    o << "\n";
  }
  o << "}\n";

  // ---- streaming form for the env kernel's flow-slot regime: every flow is stored as soon as it is known and
  // scattered into dX on the fly (few live registers: y, d), instead of materialising all flows first
  if (P.LG > 32) {
    std::vector<std::vector<std::pair<int, float>>> sc_of(P.NSRC);  // source -> (compartment, coef), in gather order
    for (int c = 0; c < P.C; c++)
      for (int q = T.xg_off[c]; q < T.xg_off[c + 1]; q++) sc_of[T.xg_src[q]].push_back({c, T.xg_coef[q]});
    auto scatter = [&](int src) {
      for (auto& pr : sc_of[src]) o << " d[" << pr.first << "] += f * " << hexf(pr.second) << ";";
    };
    o << "__device__ __forceinline__ void flows_scatter(const real* y, real alive, const float* mp, const real* rs, real* out, size_t stride, real* d, bool st) {\n";
    for (int c = 0; c < P.C; c++) o << "  d[" << c << "] = 0;\n";
    if (h->precision == EPI_FP32) {
      for (auto& dc : rdecl) o << "  real " << dc.first << "; { real dn = 0;" << dc.second << " " << dc.first << " = frcp(dn); }\n";
      rcp_of = &rmap;
    }
    for (int k = 0; k < P.NG; k++) {
      o << "  { real f = y[" << T.ig_c0[k] << "] * rs[" << k << "]; if (st) out[" << P.E + k << " * stride] = f;";
      scatter(P.E + k);
      o << " }\n";
    }
    for (int e = 0; e < P.E; e++) {
      const int* ep = T.ep.data() + T.ep_off[e];
      int n = *ep++;
      o << "  { real f; { real n = (real)1";
      for (int q = 0; q < n; q++) o << " * " << val(*ep++);
      o << ";";
      emit_denominator(o, ep, val, "f");
      o << "    if (st) out[" << e << " * stride] = f;";
      scatter(e);
      o << " }\n";
    }
    rcp_of = nullptr;
    o << "}\n";
    if (P.gl.enabled) {
    // ---- the same program for the env-group kernel (epi_grp.cuh): the flow sums FB += bw f, FE += ew f live in
    // registers; `st`: the flows are also stored (last stage of an attempt)
    o << "constexpr int grp_threads = " << P.gl.threads << ", grp_S = " << P.gl.S << ", grp_groups = " << P.gl.n_groups << ", grp_R = " << P.gl.R << ", grp_xf_smem = " << P.gl.xf_smem << ", grp_tden_smem = " << P.gl.tden_smem << ", grp_lpc = " << P.gl.lpc << ", grp_nqc = " << P.gl.nqc << ", grp_nqr = " << P.gl.nqr << ";\n";
    // regular edges (need the cell's own state only: they run while the group barrier of the gathers is pending)
    o << "__device__ __forceinline__ void flows_reg(const real* y, real alive, const float* mp, real* d, float2* FBE, float2 bwew, real* out, size_t stride, bool st) {\n";
    for (int c = 0; c < P.C; c++) o << "  d[" << c << "] = 0;\n";
    if (h->precision == EPI_FP32) {
      for (auto& dc : rdecl) o << "  real " << dc.first << "; { real dn = 0;" << dc.second << " " << dc.first << " = frcp(dn); }\n";
      rcp_of = &rmap;
    }
    auto sums = [&](int src) {
      o << " FBE[" << src << "] = ffma2(float2{f, f}, bwew, FBE[" << src << "]); if (st) out[" << src << " * stride] = f;";
    };
    for (int e = 0; e < P.E; e++) {
      const int* ep = T.ep.data() + T.ep_off[e];
      int n = *ep++;
      o << "  { real f; { real n = (real)1";
      for (int q = 0; q < n; q++) o << " * " << val(*ep++);
      o << ";";
      emit_denominator(o, ep, val, "f");
      o << "   ";
      sums(e);
      scatter(e);
      o << " }\n";
    }
    rcp_of = nullptr;
    o << "}\n";
    // infectious groups: yc[k] = the susceptible compartment of group k, rs[k] = its gathered force of infection
    o << "__device__ __forceinline__ void flows_inf(const real* yc, const real* rs, real* d, float2* FBE, float2 bwew, real* out, size_t stride, bool st) {\n";
    for (int k = 0; k < P.NG; k++) {
      o << "  { real f = yc[" << k << "] * rs[" << k << "];";
      sums(P.E + k);
      scatter(P.E + k);
      o << " }\n";
    }
    o << "}\n";
    }
    arr("ag_off", T.ag_off); arr("ag_src", T.ag_src);
    o << "__device__ constexpr float ag_coef[" << (T.ag_coef.empty() ? 1 : T.ag_coef.size()) << "] = {";
    if (T.ag_coef.empty()) o << "0";
    for (size_t i = 0; i < T.ag_coef.size(); i++) { char b[64]; snprintf(b, sizeof b, "%af", (double)T.ag_coef[i]); o << (i ? ", " : "") << b; }
    o << "};\n";
  }

  // ---- day prologue pieces of the cell kernel: per-cell partial sums of the distinct selects, the expression
  // programs and the class-multiplier products as straight-line code (executor.py:116-164, :346-359)
  bool spec_prologue = false;
  if (P.LG <= 32 && (int64_t)P.n_sel * 8 <= (int64_t)2 * P.NSRC * (int64_t)h->real_size() && P.L <= 32 && P.G <= 32) {
    spec_prologue = true;
    const int NS = P.n_sel;
    std::vector<int> rep(NS);
    for (int s = 0; s < NS; s++) {
      rep[s] = s;
      for (int r = 0; r < s; r++) {
        bool same = T.sel_mode[r] == T.sel_mode[s];
        for (int c = 0; c < P.C && same; c++) same = T.sel_c[(size_t)r * P.C + c] == T.sel_c[(size_t)s * P.C + c];
        for (int l = 0; l < P.L && same; l++) same = T.sel_l[(size_t)r * P.L + l] == T.sel_l[(size_t)s * P.L + l];
        for (int g = 0; g < P.G && same; g++) same = T.sel_g[(size_t)r * P.G + g] == T.sel_g[(size_t)s * P.G + g];
        if (same) { rep[s] = rep[r]; break; }
      }
    }
    o << "constexpr int n_sel = " << NS << ";\n";
    arr("sel_rep", rep);
    o << "__device__ __forceinline__ void sel_partials(int j, int u, const double* X, const double* X0c, const double* Xpc, double* part) {\n";
    for (int s = 0; s < NS; s++) {
      if (rep[s] != s) continue;
      unsigned lm = 0, gm = 0;
      for (int l = 0; l < P.L; l++) if (T.sel_l[(size_t)s * P.L + l]) lm |= 1u << l;
      for (int g = 0; g < P.G; g++) if (T.sel_g[(size_t)s * P.G + g]) gm |= 1u << g;
      o << "  part[" << s << "] = 0; if (((" << lm << "u >> j) & 1u) && ((" << gm << "u >> u) & 1u)) {";
      for (int q = T.selc_off[s]; q < T.selc_off[s + 1]; q++) {
        int k = T.selc[q];
        if (T.sel_mode[s] == 0) o << " part[" << s << "] += X[" << k << " * LW];";
        else if (T.sel_mode[s] == 1) o << " part[" << s << "] += X0c[" << k << " * LG];";
        else o << " part[" << s << "] += X[" << k << " * LW] - Xpc[" << T.chg_slot[k] << " * LG];";
      }
      o << " }\n";
    }
    o << "}\n";
    o << "__device__ __forceinline__ double expr_eval(int e, const float* cpv, const double* sel) {\n  switch (e) {\n";
    for (int e = 0; e < P.n_expr; e++) {
      std::vector<std::string> st;
      for (int k = T.expr_off[e]; k < T.expr_off[e + 1]; k++) {
        int op = T.tok_op[k];
        std::string r;
        char b[64];
        if (op == 0) { snprintf(b, sizeof b, "%a", T.tok_val[k]); r = b; }
        else if (op == 1) r = "(double)cpv[" + std::to_string(T.tok_arg[k]) + "]";
        else if (op == 2) r = "sel[" + std::to_string(T.tok_arg[k]) + "]";
        else if (op == 7) { std::string a = st.back(); st.pop_back(); r = "(-" + a + ")"; }
        else {
          std::string bb = st.back(); st.pop_back();
          std::string a = st.back(); st.pop_back();
          r = op == 3 ? "__dadd_rn(" + a + ", " + bb + ")" : op == 4 ? "__dsub_rn(" + a + ", " + bb + ")"
            : op == 5 ? "__dmul_rn(" + a + ", " + bb + ")" : "(" + a + " / " + bb + ")";
        }
        if (T.tok_f32[k]) r = "(double)(float)(" + r + ")";
        st.push_back(r);
      }
      o << "  case " << e << ": return " << (st.empty() ? std::string("0.0") : st[0]) << ";\n";
    }
    o << "  }\n  return 0.0;\n}\n";
    o << "__device__ __forceinline__ float cls_mult(int k, const unsigned char* live, const double* expr) {\n  float m = 1.0f;\n  switch (k) {\n";
    for (int k = 0; k < P.n_cls; k++) {
      o << "  case " << k << ":";
      for (int q = T.cls_off[k]; q < T.cls_off[k + 1]; q++) {
        int op = T.cls_op[q];
        o << " if (live[" << op << "]) m = __fmul_rn(m, (float)expr[" << T.op_expr[op] << "]);";
      }
      o << " break;\n";
    }
    o << "  }\n  return m;\n}\n";
    // compartments whose start-of-day value feeds tomorrow's 'change' selects (epidemic.py:96-99)
    arr("chg_slot", T.chg_slot);
    // today's cost rate of (intervention, locale) entry `it`: the ADD ops covering it, in op order (nb_add,
    // executor.py:236-243)
    o << "__device__ __forceinline__ double cost_rate(int it, const unsigned char* live, const double* expr) {\n  double cr = 0;\n  switch (it) {\n";
    for (int it = 0; it < P.I * P.L; it++) {
      o << "  case " << it << ":";
      for (int q = T.cr_off[it]; q < T.cr_off[it + 1]; q++) {
        int ao = T.cr_ops[q], op = T.add_op[ao];
        o << " if (live[" << op << "]) cr += expr[" << T.op_expr[op] << "] / " << T.add_parts[ao] << ".0;";
      }
      o << " break;\n";
    }
    o << "  }\n  return cr;\n}\n";
    // move ops (nb_move, executor.py:166-234): op, expression, group / source-locale masks, compartment lists
    {
      const int NM = (int)T.mv_op.size();
      o << "constexpr int n_mv_op = " << NM << ";\n";
      std::vector<int> mexpr(NM), gm(NM), lm(NM), nc1(NM), nc2(NM);
      int m1 = 1, m2 = 1;
      for (int mo = 0; mo < NM; mo++) {
        mexpr[mo] = T.op_expr[T.mv_op[mo]];
        for (int g = 0; g < P.G; g++) if (T.mv_g[(size_t)mo * P.G + g]) gm[mo] |= 1 << g;
        for (int l = 0; l < P.L; l++) if (T.mv_l1[(size_t)mo * P.L + l]) lm[mo] |= 1 << l;
        nc1[mo] = T.mv_c1_off[mo + 1] - T.mv_c1_off[mo]; nc2[mo] = T.mv_c2_off[mo + 1] - T.mv_c2_off[mo];
        m1 = std::max(m1, nc1[mo]); m2 = std::max(m2, nc2[mo]);
      }
      arr("mv_op", T.mv_op); arr("mv_expr", mexpr); arr("mv_gmask", gm); arr("mv_lmask", lm);
      arr("mv_nc1", nc1); arr("mv_nc2", nc2); arr("mv_c1_off", T.mv_c1_off); arr("mv_c2_off", T.mv_c2_off);
      arr("mv_c1", T.mv_c1); arr("mv_c2", T.mv_c2); arr("slot_of_pair", T.slot_of_pair);
      o << "constexpr int mv_max_c1 = " << m1 << ", mv_max_c2 = " << m2 << ";\n";
    }
  }
  o << "__device__ __forceinline__ void acc_rt(const real* col, size_t stride, real* out) {\n";
  for (int a = 0; a < P.n_acc; a++) {
    o << "  out[" << a << "] = 0;";
    for (int q = T.ag_off[a]; q < T.ag_off[a + 1]; q++) o << " out[" << a << "] += col[" << T.ag_src[q] << " * stride] * " << hexf(T.ag_coef[q]) << ";";
    o << "\n";
  }
  o << "}\n}}\n";
  std::string body = o.str();
  // key: FNV-1a of the body + the ABI of the structs passed by value
  uint64_t k = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { for (size_t i = 0; i < n; i++) { k ^= ((const unsigned char*)p)[i]; k *= 1099511628211ull; } };
  mix(body.data(), body.size());
  size_t abi[4] = {sizeof(DevPlan), sizeof(DevState), sizeof(epi::StepArgs), 31 /* generator revision */};
  mix(abi, sizeof abi);
  char kb[32];
  snprintf(kb, sizeof kb, "%016llx", (unsigned long long)k);
  return std::string("// GENERATED by libepi_b200 (epi_engine.cu: gen_spec) - scenario-specialised per-cell program\n#define EPI_SPEC 1\n#define EPI_SPEC_KEY \"") +
         kb + "\"\n#define EPI_SPEC_BIG " + (P.LG > 32 ? "1" : "0") + "\n" +
         "#define EPI_SPEC_PROLOGUE " + (spec_prologue ? "1" : "0") + "\n" +
         "#define EPI_SPEC_ROLES " + std::to_string(P.cl.roles > 0 ? P.cl.roles : 1) + "\n" +
         (P.LG > 32 ? std::string("#define EPI_ENV_MINB ") + (P.el.big_in_smem ? "2" : "4") + "\n" : std::string()) +
         (P.LG > 32 ? std::string("#define EPI_GRP_THREADS ") + std::to_string(P.gl.enabled ? P.gl.threads : 0) + "\n#define EPI_GRP_R " + std::to_string(P.gl.enabled ? P.gl.R : 1) + "\n" : std::string()) + body;
}

void alloc_state(epi_engine* h, DevState& S, int N) {
  const DevPlan& P = h->plan;
  size_t rs = h->real_size(), LG = P.LG;
  S.X = h->dalloc<double>((size_t)N * P.CLG);
  S.acc = h->dalloc<char>((size_t)N * P.n_acc * LG * rs);
  S.cost = h->dalloc<double>((size_t)N * P.I * P.L);
  S.crY = h->dalloc<double>((size_t)N * P.I * P.L);
  S.mvY = h->dalloc<float>((size_t)N * P.mv_len);
  S.Xprev = h->dalloc<double>((size_t)N * P.n_chg * LG);
  S.multY = h->dalloc<float>((size_t)N * P.n_cls);
  S.meta = h->dalloc<int>((size_t)N * 8);
  S.last_err = h->dalloc<double>((size_t)N);
}

void choose_launch(epi_engine* h) {
  DevPlan& P = h->plan;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  const int64_t smem_max = (int64_t)prop.sharedMemPerBlockOptin - 1024;  // leave room for static smem
  const int sms = prop.multiProcessorCount;
  int64_t per_env = P.small_bytes + P.big_bytes;
  const char* force = getenv("EPI_TEAM");  // debugging aid: "warp" / "cta" force the team kernels
  P.cl.enabled = 0;
  // (cross-locale moves, xmove: general move entries exist in the team kernels only)
  if (P.LG <= 32 && !P.xmove && !(force && (!strcmp(force, "warp") || !strcmp(force, "cta") || !strcmp(force, "env")))) {
#ifdef EPI_HOST_EMU
    int W = P.LG * (h->N < 32 / P.LG ? h->N : 32 / P.LG);  // as many lanes as the test's envs need
#else
    int W = 32;
#endif
    build_cell_layout(h, W);
    if (P.cl.warp_bytes <= smem_max) {
      // CELL kernel: one lane per (env, locale, group), EPW envs per warp, one warp per CTA
      P.cl.enabled = 1;
      h->cta_mode = 2; h->cellW = W;
      P.small_in_smem = P.big_in_smem = 1;
      h->wpb = P.cl.roles; h->threads = W * P.cl.roles;  // (the emulation runs LW host threads per warp)
      h->smem_bytes = (size_t)P.cl.warp_bytes;
      h->grid = (h->N + P.cl.epw - 1) / P.cl.epw;
      h->scratch_bytes = 0;
      return;
    }
  }
  P.el.enabled = 0;
  // (mobility multipliers, MOB_MUL: per-env mobility tables exist in the cell kernel and the team kernels only)
  if (P.mob_dyn && force && !strcmp(force, "env")) throw Unsupported{"mobility multipliers in the env kernel"};
  if (P.xmove && force && !strcmp(force, "env")) throw Unsupported{"cross-locale moves in the env kernel"};
  if (!P.mob_dyn && !P.xmove && (P.LG > 32 || (force && !strcmp(force, "env"))) && !(force && (!strcmp(force, "warp") || !strcmp(force, "cta")))) {
    // ENV kernel: one CTA per env, threads strided over the cells
#ifdef EPI_HOST_EMU
    int threads = P.LG < 24 ? P.LG : 24;  // host threads per emulated CTA
#else
    int threads = env_threads(P.LG);
#endif
    build_env_layout(h, threads, smem_max);
    P.el.enabled = 1;
    h->cta_mode = 3; h->threads = threads; h->wpb = 1;
    P.small_in_smem = 1; P.big_in_smem = P.el.big_in_smem;
    h->smem_bytes = (size_t)(P.el.tab_bytes + P.el.red_bytes + (P.el.comm_in_smem ? P.el.comm_bytes : 0) +
                             (P.el.big_in_smem ? P.el.big_bytes : 0));
    int ctas_per_sm = (int)((prop.sharedMemPerMultiprocessor - 1024) / (h->smem_bytes + 1024));
    // __launch_bounds__(256, EPI_ENV_MINB): 2 CTAs/SM (<= 128 registers) when the working set is in shared
    // memory, 4 (<= 64 registers, more loads in flight) when it streams from the global workspace
    int by_regs = 65536 / (threads * (P.el.big_in_smem ? 128 : 64));
    if (by_regs < 1) by_regs = 1;
    if (const char* ov = getenv("EPI_ENV_CTAS")) by_regs = atoi(ov) > 0 ? atoi(ov) : by_regs;  // tuning aid
    if (ctas_per_sm > by_regs) ctas_per_sm = by_regs;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int cap = sms * ctas_per_sm;
    h->grid = h->N < cap ? h->N : cap;
    h->scratch_bytes = (size_t)h->grid * ((P.el.big_in_smem ? 0 : P.el.big_bytes) + (P.el.comm_in_smem ? 0 : P.el.comm_bytes));
    if (!P.el.big_in_smem) h->big_scratch = h->dalloc<char>((size_t)P.el.big_bytes * h->grid, false);
    if (!P.el.comm_in_smem) h->small_scratch = h->dalloc<char>((size_t)P.el.comm_bytes * h->grid, false);
    h->cr_scratch = h->dalloc<double>((size_t)P.I * P.L * h->grid);
    // the env-group kernel of the specialised build (used once such a build is attached, epi_spec_attach)
    h->sms = sms;
    build_grp_layout(h, sms, smem_max);
#ifndef EPI_HOST_EMU
    if (P.gl.enabled && !prop.cooperativeLaunch) P.gl.enabled = 0;
#endif
    if (P.gl.enabled) {
      const int ng = P.gl.n_groups < h->N ? P.gl.n_groups : h->N;
      h->grp_scratch = h->dalloc<char>((size_t)P.gl.g_bytes * ng, false);
      h->grp_cta_scratch = h->dalloc<char>((size_t)P.gl.c_bytes * ng * P.gl.S, false);
      h->grp_bar = h->dalloc<unsigned>((size_t)32 * ng + 8 + 1024);  // arrive counters | class counters | per-SM counters
    }
    return;
  }
  if (force && !strcmp(force, "cta")) per_env = INT64_MAX / 16;
  if (per_env * 8 <= (int64_t)prop.sharedMemPerMultiprocessor - 8 * 1024 || per_env <= 28 * 1024) {
    // WARP teams: the whole working set of an env in shared memory
    h->cta_mode = 0;
    P.small_in_smem = P.big_in_smem = 1;
    int wpb = 4;
    while (wpb > 1 && per_env * wpb > smem_max) wpb >>= 1;
    h->wpb = wpb; h->threads = 32 * wpb;
    h->smem_bytes = (size_t)(per_env * wpb);
    int ctas_per_sm = (int)((prop.sharedMemPerMultiprocessor - 1024) / (h->smem_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * wpb > 48) ctas_per_sm = 48 / wpb;
    int want = (h->N + wpb - 1) / wpb;
    int cap = sms * ctas_per_sm;
    h->grid = want < cap ? want : cap;
    h->scratch_bytes = 0;
  } else {
    h->cta_mode = 1;
    P.small_in_smem = P.small_bytes <= 96 * 1024;
    P.big_in_smem = P.small_in_smem && (P.small_bytes + P.big_bytes <= smem_max);
    h->threads = P.LG >= 2048 ? 512 : 256;
    h->wpb = 1;
    h->smem_bytes = (size_t)((P.small_in_smem ? P.small_bytes : 0) + (P.big_in_smem ? P.big_bytes : 0));
    int ctas_per_sm = (int)((prop.sharedMemPerMultiprocessor - 1024) / (h->smem_bytes + 2048));
    int by_threads = 1536 / h->threads;
    if (ctas_per_sm > by_threads) ctas_per_sm = by_threads;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int cap = sms * ctas_per_sm;
    h->grid = h->N < cap ? h->N : cap;
    int64_t per_team = (P.big_in_smem ? 0 : P.big_bytes) + (P.small_in_smem ? 0 : P.small_bytes);
    h->scratch_bytes = (size_t)per_team * h->grid;
    if (!P.big_in_smem) h->big_scratch = h->dalloc<char>((size_t)P.big_bytes * h->grid, false);
    if (!P.small_in_smem) h->small_scratch = h->dalloc<char>((size_t)P.small_bytes * h->grid, false);
  }
}

template <class real, bool CTA>
void launch_t(epi_engine* h, const DevState& S, const epi::StepArgs& a, int grid, cudaStream_t st) {
#ifdef EPI_HOST_EMU
  (void)grid; (void)st;
  epi::step_emulated<real, CTA>(h->plan, S, a);
#else
  auto kern = epi::step_kernel<real, CTA>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  kern<<<grid, h->threads, h->smem_bytes, st>>>(h->plan, S, a);
  CK(cudaGetLastError());
#endif
  h->launches++;
}

template <class real>
void launch_cell(epi_engine* h, const DevState& S, const epi::StepArgs& a, cudaStream_t st) {
  const int epw = h->plan.cl.epw;
  const int grid = (a.N + epw - 1) / epw;
#ifdef EPI_HOST_EMU
  (void)st;
  if (h->spec_launch) {
    if (h->spec_launch(&h->plan, &S, &a, grid, h->cellW, (size_t)h->plan.cl.warp_bytes, nullptr) != 0)
      throw Invalid{"specialised cell kernel (host emulation)"};
  } else
    emu::run_warps(grid, h->cellW, (size_t)h->plan.cl.warp_bytes, [&](char* smem, int lane, int group) {
      epi::cell::step_warp<real>(h->plan, S, a, smem, lane, group);
    });
#else
  if (h->spec_launch) {
    int rc = h->spec_launch(&h->plan, &S, &a, grid, h->threads, h->smem_bytes, (void*)st);
    if (rc != 0) throw CudaFail{std::string("specialised cell kernel launch: ") + cudaGetErrorString((cudaError_t)rc)};
  } else {
    auto kern = epi::cell::cell_kernel<real>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<grid, h->threads, h->smem_bytes, st>>>(h->plan, S, a);
    CK(cudaGetLastError());
  }
#endif
  h->launches++;
}

template <class real>
void launch_env(epi_engine* h, const DevState& S, const epi::StepArgs& a, cudaStream_t st) {
  const int grid = a.N < h->grid ? a.N : h->grid;
#ifdef EPI_HOST_EMU
  (void)st;
  if (h->spec_launch) {
    if (h->spec_launch(&h->plan, &S, &a, grid, h->threads, h->smem_bytes, nullptr) != 0)
      throw Invalid{"specialised env kernel (host emulation)"};
  } else
  emu::run_warps(grid, h->threads, h->smem_bytes, [&](char* smem, int tid, int cta) {
    epi::env::Bx b = epi::env::bind(h->plan, a, smem, tid, h->threads, cta);
    epi::env::step_cta<real>(h->plan, S, a, b, cta, grid);
  });
#else
  if (h->spec_launch) {
    int rc = h->spec_launch(&h->plan, &S, &a, grid, h->threads, h->smem_bytes, (void*)st);
    if (rc != 0) throw CudaFail{std::string("specialised env kernel launch: ") + cudaGetErrorString((cudaError_t)rc)};
  } else {
    auto kern = epi::env::env_kernel<real>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    kern<<<grid, h->threads, h->smem_bytes, st>>>(h->plan, S, a);
    CK(cudaGetLastError());
  }
#endif
  h->launches++;
}

// env-group kernel: n_groups x S co-resident CTAs (cooperative launch), arrive counters cleared in stream order
void launch_grp(epi_engine* h, const DevState& S, epi::StepArgs a, cudaStream_t st) {
  const GrpLayout& Y = h->plan.gl;
  const int ng = Y.n_groups < a.N ? Y.n_groups : a.N;
  a.grp_scratch = h->grp_scratch; a.grp_cta_scratch = h->grp_cta_scratch; a.grp_bar = h->grp_bar;
  CK(cudaMemsetAsync(h->grp_bar, 0, sizeof(unsigned) * (32 * (size_t)ng + 8 + 1024), st));
  int rc = h->spec_launch_grp(&h->plan, &S, &a, ng * Y.S, Y.threads, (size_t)Y.smem_bytes, (void*)st);
  if (rc != 0) throw CudaFail{std::string("env-group kernel launch: ") + cudaGetErrorString((cudaError_t)rc)};
  h->launches++;
}

void launch(epi_engine* h, const DevState& S, epi::StepArgs a, cudaStream_t st) {
  a.dbg_attempts = (a.N == h->N) ? h->dbg_attempts : nullptr;
  if (h->cta_mode == 4) {
    launch_grp(h, S, a, st);
    return;
  }
  if (h->cta_mode == 3) {
    a.big_scratch = h->big_scratch;
    a.small_scratch = h->small_scratch;
    a.cr_scratch = h->cr_scratch;
    if (h->precision == EPI_FP64) launch_env<double>(h, S, a, st); else launch_env<float>(h, S, a, st);
    return;
  }
  if (h->cta_mode == 2) {
    if (h->precision == EPI_FP64) launch_cell<double>(h, S, a, st); else launch_cell<float>(h, S, a, st);
    return;
  }
  a.big_scratch = h->big_scratch;
  a.small_scratch = h->small_scratch;
  int grid = h->grid;
  if (a.N < h->N) grid = 1;
  if (h->precision == EPI_FP64) {
    if (h->cta_mode) launch_t<double, true>(h, S, a, grid, st); else launch_t<double, false>(h, S, a, grid, st);
  } else {
    if (h->cta_mode) launch_t<float, true>(h, S, a, grid, st); else launch_t<float, false>(h, S, a, grid, st);
  }
}

void do_reset(epi_engine* h, const uint8_t* mask, void* obs_out, cudaStream_t st) {
  int grid = h->N < 4096 ? h->N : 4096;
#ifdef EPI_HOST_EMU
  (void)grid; (void)st;
  if (h->precision == EPI_FP64) epi::reset_kernel<double>(h->plan, h->S, h->init, mask, h->N, obs_out);
  else epi::reset_kernel<float>(h->plan, h->S, h->init, mask, h->N, obs_out);
#else
  if (h->precision == EPI_FP64) epi::reset_kernel<double><<<grid, 128, 0, st>>>(h->plan, h->S, h->init, mask, h->N, obs_out);
  else epi::reset_kernel<float><<<grid, 128, 0, st>>>(h->plan, h->S, h->init, mask, h->N, obs_out);
#endif
  CK(cudaGetLastError());
  h->launches++;
}

// pinned host staging + device mirrors of the host-buffer path
void ensure_host_buffers(epi_engine* h) {
  if (h->h_act) return;
  const size_t N = h->N, A = h->plan.A ? h->plan.A : 1, rs = h->real_size(), CLG = h->plan.CLG;
  CK(cudaHostAlloc(&h->h_act, N * A * 4, cudaHostAllocMapped)); h->d_act = h->dalloc<float>(N * A);
  CK(cudaHostAlloc(&h->h_obs, N * CLG * rs, cudaHostAllocMapped)); h->d_obs = h->dalloc<char>(N * CLG * rs);
  CK(cudaHostAlloc(&h->h_rew, N * 8, cudaHostAllocMapped)); h->d_rew = h->dalloc<double>(N);
  CK(cudaHostAlloc(&h->h_done, N, cudaHostAllocMapped)); h->d_done = h->dalloc<uint8_t>(N);
  CK(cudaHostGetDevicePointer((void**)&h->m_act, h->h_act, 0)); CK(cudaHostGetDevicePointer(&h->m_obs, h->h_obs, 0));
  CK(cudaHostGetDevicePointer((void**)&h->m_rew, h->h_rew, 0)); CK(cudaHostGetDevicePointer((void**)&h->m_done, h->h_done, 0));
  if (const char* z = getenv("EPI_HOST_ZEROCOPY")) h->zerocopy = atoi(z);
  // dalloc's cudaMemset runs on the legacy default stream, which h->stream (non-blocking) does not wait for
  CK(cudaDeviceSynchronize());
}

template <class F>
int guarded(const epi_engine* h, F&& f) {
  try {
    f();
    return EPI_OK;
  } catch (const Unsupported& e) {
    (h ? h->err : g_create_error) = "unsupported: " + e.msg;
    return EPI_E_UNSUPPORTED;
  } catch (const Invalid& e) {
    (h ? h->err : g_create_error) = "invalid: " + e.msg;
    return EPI_E_INVALID;
  } catch (const CudaFail& e) {
    (h ? h->err : g_create_error) = "cuda: " + e.msg;
    return EPI_E_CUDA;
  } catch (const std::bad_alloc&) {
    (h ? h->err : g_create_error) = "out of host memory";
    return EPI_E_NOMEM;
  } catch (...) {
    (h ? h->err : g_create_error) = "unknown failure";
    return EPI_E_INVALID;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int epi_create(const EpiDesc* desc, int n_envs, int device, int precision, epi_t** out) {
  if (!desc || !out || n_envs < 1 || (precision != EPI_FP64 && precision != EPI_FP32)) {
    g_create_error = "invalid: null descriptor/out pointer, n_envs < 1 or unknown precision";
    return EPI_E_INVALID;
  }
  epi_engine* h = new epi_engine();
  h->device = device; h->precision = precision; h->N = n_envs;
  int rc = guarded(nullptr, [&] {
    CK(cudaSetDevice(device));
    h->d = *desc;
    build_plan(h, *desc);
    choose_launch(h);
    alloc_state(h, h->S, n_envs);
    alloc_state(h, h->init, 1);
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // make_setting (core/epidemic.py:77-80): execute the schedule's day-0 action on the default state,
    // keep the resulting parameters as "yesterday" of every fresh episode
    {
      DevState tmp = h->init;
      std::vector<double> x0 = vec(desc->X0, (size_t)h->plan.CLG);
      CK(cudaMemcpy(tmp.X, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice));
      std::vector<float> ones(h->plan.n_cls, 1.0f);
      CK(cudaMemcpy(tmp.multY, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
      {
        // the previous state of the default state is itself ('change' selects give 0, epidemic.py:96-99)
        std::vector<int> chg(h->plan.C);
        CK(cudaMemcpy(chg.data(), h->plan.chg_slot_of_comp, sizeof(int) * h->plan.C, cudaMemcpyDeviceToHost));
        size_t LG = h->plan.LG;
        for (int c = 0; c < h->plan.C; c++)
          if (chg[c] >= 0) CK(cudaMemcpy(tmp.Xprev + (size_t)chg[c] * LG, tmp.X + (size_t)c * LG, LG * 8, cudaMemcpyDeviceToDevice));
      }
      epi::StepArgs a{};
      a.variant = 1; a.n_days = 0; a.init_only = 1; a.N = 1;
      launch(h, h->init, a, h->stream);
      CK(cudaStreamSynchronize(h->stream));
    }
    do_reset(h, nullptr, nullptr, h->stream);
    CK(cudaStreamSynchronize(h->stream));
  });
  if (rc != EPI_OK) {
    for (void* p : h->allocs) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    *out = nullptr;
    return rc;
  }
  *out = h;
  return EPI_OK;
}

void epi_destroy(epi_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->h_act) cudaFreeHost(h->h_act);
  if (h->h_obs) cudaFreeHost(h->h_obs);
  if (h->h_rew) cudaFreeHost(h->h_rew);
  if (h->h_done) cudaFreeHost(h->h_done);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->spec_lib) dlclose(h->spec_lib);
  delete h;
}

const char* epi_last_error(const epi_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int epi_info(const epi_t* h, EpiInfo* o) {
  if (!h || !o) return EPI_E_INVALID;
  const DevPlan& P = h->plan;
  memset(o, 0, sizeof(*o));
  o->n_envs = h->N; o->precision = h->precision; o->device = h->device; o->obs_dim = P.CLG; o->action_dim = P.A;
  o->ncp = P.NCP; o->n_itv = P.I; o->horizon = P.horizon; o->n_flat = P.n_flat; o->n_acc = P.n_acc; o->n_slot = P.n_slot;
  o->n_lc = P.n_lc; o->n_groups_inf = P.NG; o->n_cls = P.n_cls; o->team_kind = h->cta_mode;
  o->threads = h->cta_mode == 4 ? P.gl.threads : h->threads;
  o->grid = h->cta_mode == 4 ? (P.gl.n_groups < h->N ? P.gl.n_groups : h->N) * P.gl.S : h->grid; o->envs_per_cta = h->cta_mode == 2 ? h->plan.cl.epw : (h->cta_mode ? 1 : h->wpb);
  o->smem_bytes = h->cta_mode == 4 ? P.gl.smem_bytes : (int64_t)h->smem_bytes;
  o->specialized = h->spec_launch ? 1 : 0; o->reserved = h->cta_mode == 4 ? P.gl.S : 0;  // CTAs per env group
  o->small_bytes = P.small_bytes; o->big_bytes = P.big_bytes; o->scratch_bytes = (int64_t)h->scratch_bytes;
  int64_t rs = (int64_t)h->real_size();
  o->mv_len = P.mv_len;
  o->state_bytes_per_env = 8 * (P.CLG + (int64_t)P.n_chg * P.LG) + rs * (int64_t)P.n_acc * P.LG + 16 * (int64_t)P.I * P.L +
                           4 * ((int64_t)P.mv_len + P.n_cls) + 32 + 8;
  return EPI_OK;
}

int epi_reset(epi_t* h, const uint8_t* mask_dev, void* obs_out_dev, void* cuda_stream) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    CK(cudaSetDevice(h->device));
    do_reset(h, mask_dev, obs_out_dev, (cudaStream_t)cuda_stream);
  });
}

// the team kernels (A/B baselines) do not write the mirror in their epilogue: copy behind the launch instead
static void mirror_fallback(epi_engine* h, const epi::StepArgs& a, cudaStream_t st) {
  if (h->cta_mode >= 2) return;
  const size_t N = h->N;
  if (a.obs_out2 && a.obs_out) CK(cudaMemcpyAsync(a.obs_out2, a.obs_out, N * h->plan.CLG * h->real_size(), cudaMemcpyDefault, st));
  if (a.reward_out2 && a.reward_out) CK(cudaMemcpyAsync(a.reward_out2, a.reward_out, N * 8, cudaMemcpyDefault, st));
  if (a.done_out2 && a.done_out) CK(cudaMemcpyAsync(a.done_out2, a.done_out, N, cudaMemcpyDefault, st));
}

int epi_bind_mirror(epi_t* h, void* obs_dev, double* reward_dev, uint8_t* done_dev) {
  if (!h) return EPI_E_INVALID;
  h->mir_obs = obs_dev; h->mir_rew = reward_dev; h->mir_done = done_dev;
  return EPI_OK;
}

int epi_step(epi_t* h, const float* actions_dev, int n_days, void* obs_out_dev, double* reward_out_dev,
             uint8_t* done_out_dev, void* cuda_stream) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (!actions_dev && h->plan.A > 0) throw Invalid{"actions_dev is NULL"};
    if (n_days < 1) throw Invalid{"n_days < 1"};
    CK(cudaSetDevice(h->device));
    epi::StepArgs a{};
    a.actions = actions_dev; a.variant = 0; a.n_days = n_days; a.N = h->N;
    a.obs_out = obs_out_dev; a.reward_out = reward_out_dev; a.done_out = done_out_dev;
    a.obs_out2 = obs_out_dev ? h->mir_obs : nullptr; a.reward_out2 = h->mir_rew; a.done_out2 = h->mir_done;
    launch(h, h->S, a, (cudaStream_t)cuda_stream);
    mirror_fallback(h, a, (cudaStream_t)cuda_stream);
  });
}

int epi_step_plan(epi_t* h, const float* cpv_dev, const uint8_t* active_dev, int schedule_programs, int n_days,
                  void* obs_out_dev, double* reward_out_dev, uint8_t* done_out_dev, void* cuda_stream) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (!cpv_dev) throw Invalid{"cpv_dev is NULL"};
    if (n_days < 1) throw Invalid{"n_days < 1"};
    CK(cudaSetDevice(h->device));
    epi::StepArgs a{};
    a.cpv = cpv_dev; a.active = active_dev; a.variant = schedule_programs ? 1 : 0; a.n_days = n_days; a.N = h->N;
    a.obs_out = obs_out_dev; a.reward_out = reward_out_dev; a.done_out = done_out_dev;
    a.obs_out2 = obs_out_dev ? h->mir_obs : nullptr; a.reward_out2 = h->mir_rew; a.done_out2 = h->mir_done;
    launch(h, h->S, a, (cudaStream_t)cuda_stream);
    mirror_fallback(h, a, (cudaStream_t)cuda_stream);
  });
}

int epi_run_plan(epi_t* h, const float* cpv_days_dev, const uint8_t* active_days_dev, int schedule_programs, int n_days,
                 const EpiRunOut* days_out, void* obs_out_dev, double* reward_out_dev, uint8_t* done_out_dev,
                 void* cuda_stream) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (!cpv_days_dev) throw Invalid{"cpv_days_dev is NULL"};
    if (n_days < 1) throw Invalid{"n_days < 1"};
    CK(cudaSetDevice(h->device));
    epi::StepArgs a{};
    a.cpv = cpv_days_dev; a.active = active_days_dev;  // row 0; the kernels move on to row d on day d
    a.cpv_days = cpv_days_dev; a.active_days = active_days_dev;
    if (days_out) {
      a.obs_days = days_out->obs_days; a.reward_days = days_out->reward_days; a.x_days = days_out->x_days;
      a.acc_days = days_out->acc_days; a.mult_days = days_out->mult_days;
    }
    a.variant = schedule_programs ? 1 : 0; a.n_days = n_days; a.N = h->N;
    a.obs_out = obs_out_dev; a.reward_out = reward_out_dev; a.done_out = done_out_dev;
    launch(h, h->S, a, (cudaStream_t)cuda_stream);
  });
}

int epi_step_host(epi_t* h, const float* actions_host, int n_days, void* obs_out_host, double* reward_out_host,
                  uint8_t* done_out_host) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (!actions_host && h->plan.A > 0) throw Invalid{"actions_host is NULL"};
    if (n_days < 1) throw Invalid{"n_days < 1"};
    CK(cudaSetDevice(h->device));
    const size_t N = h->N, rs = h->real_size(), CLG = h->plan.CLG;
    ensure_host_buffers(h);
    // The buffers are pinned and mapped: by default the kernel epilogue writes obs / reward / done straight into host
    // memory (posted PCIe writes that overlap the envs still being stepped) and no copy is enqueued behind the launch;
    // the actions (24 B per env) are read by the kernel head the same way when bit 1 is set.
    const bool zc_out = h->zerocopy & 1, zc_act = h->zerocopy & 2;
    if (h->plan.A > 0) {
      if (actions_host != h->h_act) memcpy(h->h_act, actions_host, N * h->plan.A * 4);
      if (!zc_act) CK(cudaMemcpyAsync(h->d_act, h->h_act, N * h->plan.A * 4, cudaMemcpyHostToDevice, h->stream));
    }
    epi::StepArgs a{};
    a.actions = zc_act ? h->m_act : h->d_act; a.variant = 0; a.n_days = n_days; a.N = h->N;
    a.obs_out = obs_out_host ? (zc_out ? h->m_obs : h->d_obs) : nullptr;
    a.reward_out = zc_out ? h->m_rew : h->d_rew; a.done_out = zc_out ? h->m_done : h->d_done;
    launch(h, h->S, a, h->stream);
    if (!zc_out) {
      if (obs_out_host) CK(cudaMemcpyAsync(h->h_obs, h->d_obs, N * CLG * rs, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(h->h_rew, h->d_rew, N * 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(h->h_done, h->d_done, N, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    // callers that use the library's own pinned buffers (epi_host_buffers) read the results in place
    if (obs_out_host && obs_out_host != h->h_obs) memcpy(obs_out_host, h->h_obs, N * CLG * rs);
    if (reward_out_host && reward_out_host != h->h_rew) memcpy(reward_out_host, h->h_rew, N * 8);
    if (done_out_host && done_out_host != h->h_done) memcpy(done_out_host, h->h_done, N);
  });
}

int epi_host_buffers(epi_t* h, float** actions, void** obs, double** reward, uint8_t** done) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    CK(cudaSetDevice(h->device));
    ensure_host_buffers(h);
    if (actions) *actions = h->h_act;
    if (obs) *obs = h->h_obs;
    if (reward) *reward = h->h_rew;
    if (done) *done = h->h_done;
  });
}

int epi_get_state(epi_t* h, int what, void* out, size_t bytes) {
  if (!h || !out) return EPI_E_INVALID;
  return guarded(h, [&] {
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const DevPlan& P = h->plan;
    const size_t N = h->N, rs = h->real_size(), LG = P.LG, nC = (size_t)P.I * P.L;
    auto need = [&](size_t n) { if (bytes != n) throw Invalid{"buffer size mismatch: expected " + std::to_string(n) + " bytes"}; };
    switch (what) {
      case 0: need(N * P.CLG * 8); CK(cudaMemcpy(out, h->S.X, bytes, cudaMemcpyDeviceToHost)); break;
      case 1: need(N * nC * 8); CK(cudaMemcpy(out, h->S.cost, bytes, cudaMemcpyDeviceToHost)); break;
      case 2: {
        need(N * 4);
        std::vector<int> m(N * 8);
        CK(cudaMemcpy(m.data(), h->S.meta, N * 32, cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < N; e++) ((int32_t*)out)[e] = m[e * 8];
      } break;
      case 3: {  // the reference's flat observable (obj/observable.py:24-25), structural zeros included
        need(N * (size_t)P.n_flat * 8);
        std::vector<double> X(N * P.CLG);
        std::vector<char> acc(N * P.n_acc * LG * rs);
        std::vector<double> cost(N * nC);
        CK(cudaMemcpy(X.data(), h->S.X, X.size() * 8, cudaMemcpyDeviceToHost));
        if (!acc.empty()) CK(cudaMemcpy(acc.data(), h->S.acc, acc.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cost.data(), h->S.cost, cost.size() * 8, cudaMemcpyDeviceToHost));
        double* o = (double*)out;
        memset(o, 0, bytes);
        auto rd = [&](const std::vector<char>& v, size_t i) { return rs == 8 ? ((const double*)v.data())[i] : (double)((const float*)v.data())[i]; };
        for (size_t e = 0; e < N; e++) {
          double* f = o + e * (size_t)P.n_flat;
          for (size_t i = 0; i < (size_t)P.CLG; i++) f[i] = X[e * P.CLG + i];
          for (int a = 0; a < P.n_acc; a++)
            for (size_t i = 0; i < LG; i++) f[(size_t)h->acc_flat_host[a] * LG + i] = rd(acc, (e * P.n_acc + a) * LG + i);
          for (size_t i = 0; i < nC; i++) f[(size_t)(3 + P.C) * P.CLG + i] = cost[e * nC + i];
        }
      } break;
      case 4: {
        need(N * 16);
        std::vector<int> m(N * 8);
        CK(cudaMemcpy(m.data(), h->S.meta, N * 32, cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < N; e++) {
          int32_t* t = (int32_t*)out + e * 4;
          t[0] = m[e * 8 + 3]; t[1] = m[e * 8 + 4]; t[2] = m[e * 8 + 5]; t[3] = m[e * 8 + 2];
        }
      } break;
      case 5: need(N * 8); CK(cudaMemcpy(out, h->S.last_err, bytes, cudaMemcpyDeviceToHost)); break;
      case 6: need(N * nC * 8); CK(cudaMemcpy(out, h->S.crY, bytes, cudaMemcpyDeviceToHost)); break;
      case 7: need(N * P.n_cls * 4); CK(cudaMemcpy(out, h->S.multY, bytes, cudaMemcpyDeviceToHost)); break;
      case 8: need(N * (size_t)P.mv_len * 4); if (bytes) CK(cudaMemcpy(out, h->S.mvY, bytes, cudaMemcpyDeviceToHost)); break;
      case 9: {  // (h, error_norm) per attempted RK step of the last day; recording starts at the first query
        need(N * 128 * 8);
        if (!h->dbg_attempts) h->dbg_attempts = h->dalloc<double>(N * 128);
        CK(cudaMemcpy(out, h->dbg_attempts, bytes, cudaMemcpyDeviceToHost));
      } break;
      default: throw Invalid{"unknown state selector"};
    }
  });
}

int epi_set_state(epi_t* h, int what, const void* in, size_t bytes) {
  if (!h || !in) return EPI_E_INVALID;
  return guarded(h, [&] {
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const DevPlan& P = h->plan;
    const size_t N = h->N, rs = h->real_size(), nC = (size_t)P.I * P.L;
    auto need = [&](size_t n) { if (bytes != n) throw Invalid{"buffer size mismatch: expected " + std::to_string(n) + " bytes"}; };
    switch (what) {
      case 0: need(N * P.CLG * 8); CK(cudaMemcpy(h->S.X, in, bytes, cudaMemcpyHostToDevice)); break;
      case 1: need(N * nC * 8); CK(cudaMemcpy(h->S.cost, in, bytes, cudaMemcpyHostToDevice)); break;
      case 2: {
        need(N * 4);
        std::vector<int> m(N * 8);
        CK(cudaMemcpy(m.data(), h->S.meta, N * 32, cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < N; e++) m[e * 8] = ((const int32_t*)in)[e];
        CK(cudaMemcpy(h->S.meta, m.data(), N * 32, cudaMemcpyHostToDevice));
      } break;
      case 6: need(N * nC * 8); CK(cudaMemcpy(h->S.crY, in, bytes, cudaMemcpyHostToDevice)); break;
      case 7: need(N * P.n_cls * 4); CK(cudaMemcpy(h->S.multY, in, bytes, cudaMemcpyHostToDevice)); break;
      case 8: need(N * (size_t)P.mv_len * 4); if (bytes) CK(cudaMemcpy(h->S.mvY, in, bytes, cudaMemcpyHostToDevice)); break;
      default: throw Invalid{"state selector not settable"};
    }
  });
}

int epi_get_state_dev(epi_t* h, int what, void* out_dev, size_t bytes, void* cuda_stream) {
  if (!h || !out_dev) return EPI_E_INVALID;
  return guarded(h, [&] {
    CK(cudaSetDevice(h->device));
    const DevPlan& P = h->plan;
    const size_t N = h->N, rs = h->real_size(), LG = P.LG, nC = (size_t)P.I * P.L;
    auto need = [&](size_t n) { if (bytes != n) throw Invalid{"buffer size mismatch: expected " + std::to_string(n) + " bytes"}; };
    const void* src = nullptr;
    switch (what) {
      case 0: need(N * P.CLG * 8); src = h->S.X; break;
      case 1: need(N * nC * 8); src = h->S.cost; break;
      case 7: need(N * P.n_cls * 4); src = h->S.multY; break;
      case 10: need(N * P.n_acc * LG * rs); src = h->S.acc; break;
      default: throw Invalid{"state selector has no device copy"};
    }
    if (bytes) CK(cudaMemcpyAsync(out_dev, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
  });
}

int epi_get_plan(const epi_t* h, int what, void* out_host, size_t bytes) {
  if (!h || !out_host) return EPI_E_INVALID;
  return guarded(h, [&] {
    const HostTabs& T = h->tabs;
    auto put = [&](const void* p, size_t n) {
      if (bytes != n) throw Invalid{"buffer size mismatch: expected " + std::to_string(n) + " bytes"};
      memcpy(out_host, p, n);
    };
    switch (what) {
      case 0: put(h->acc_flat_host.data(), h->acc_flat_host.size() * 4); break;
      case 1: put(T.lclass.data(), T.lclass.size() * 4); break;
      case 2: put(T.mp0.data(), T.mp0.size() * 4); break;
      case 3: put(T.mp_cls.data(), T.mp_cls.size() * 4); break;
      case 4: put(T.ft0.data(), T.ft0.size() * 4); break;
      case 5: put(T.ft_cls.data(), T.ft_cls.size() * 4); break;
      default: throw Invalid{"unknown plan table"};
    }
  });
}

int64_t epi_launch_count(const epi_t* h) { return h ? h->launches : 0; }

int epi_spec_source(epi_t* h, char* buf, size_t cap, size_t* need) {
  if (!h) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (h->cta_mode < 2) throw Unsupported{"scenario specialisation exists for the cell and env kernels only"};
    std::string src = gen_spec(h);
    if (need) *need = src.size() + 1;
    if (buf && cap) {
      size_t n = src.size() < cap - 1 ? src.size() : cap - 1;
      memcpy(buf, src.data(), n);
      buf[n] = 0;
    }
  });
}

int epi_spec_source_desc(const EpiDesc* desc, int precision, char* buf, size_t cap, size_t* need) {
  if (!desc || (precision != EPI_FP64 && precision != EPI_FP32)) return EPI_E_INVALID;
  epi_engine* h = new epi_engine();
  h->precision = precision; h->N = 1; h->host_only = true;
  int rc = guarded(nullptr, [&] {
    h->d = *desc;
    build_plan(h, *desc);
    // scenarios that run on the team kernels have no specialised build
    if (h->plan.xmove) throw Unsupported{"cross-locale moves run on the team kernels: no specialised build"};
    if (h->plan.mob_dyn && h->plan.LG > 32) throw Unsupported{"mobility multipliers with L*G > 32 run on the team kernels: no specialised build"};
    if (h->plan.LG <= 32) build_cell_layout(h, 32);
    else {
      build_env_layout(h, env_threads(h->plan.LG), 232448 - 1024);  // B200: 227 KB opt-in shared memory per CTA
      build_grp_layout(h, 148, 232448 - 1024);                      // B200: 148 SMs
    }
    std::string src = gen_spec(h);
    if (need) *need = src.size() + 1;
    if (buf && cap) {
      size_t n = src.size() < cap - 1 ? src.size() : cap - 1;
      memcpy(buf, src.data(), n);
      buf[n] = 0;
    }
  });
  for (void* p : h->host_allocs) free(p);
  delete h;
  return rc;
}

int epi_spec_attach(epi_t* h, const char* so_path) {
  if (!h || !so_path) return EPI_E_INVALID;
  return guarded(h, [&] {
    if (h->cta_mode < 2) throw Unsupported{"scenario specialisation exists for the cell and env kernels only"};
    // the generated build is the cell kernel for L*G <= 32 and the env kernel above: a kernel kind forced with EPI_TEAM
    // (debugging aid) on the other side of that line has no specialised build and runs interpreted
    if ((h->cta_mode >= 3) != (h->plan.LG > 32)) throw Unsupported{"no specialised build for the forced kernel kind"};
    void* lib = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw Invalid{std::string("dlopen failed: ") + dlerror()};
    auto keyf = (const char* (*)())dlsym(lib, "epi_spec_key");
    auto lf = (int (*)(const DevPlan*, const DevState*, const epi::StepArgs*, int, int, size_t, void*))dlsym(lib, "epi_spec_launch");
    std::string src = gen_spec(h);
    size_t p0 = src.find("EPI_SPEC_KEY \"") + 14;
    std::string want = src.substr(p0, 16);
    if (!keyf || !lf || want != keyf()) {
      dlclose(lib);
      throw Invalid{"specialised library does not match this scenario / build (key mismatch)"};
    }
    h->spec_lib = lib;
    h->spec_launch = lf;
    // the env-group kernel, when this build carries one and the layout is on for this engine
    auto lg = (int (*)(const DevPlan*, const DevState*, const epi::StepArgs*, int, int, size_t, void*))dlsym(lib, "epi_spec_launch_grp");
    if (lg && h->plan.gl.enabled && h->cta_mode == 3) {
      h->spec_launch_grp = lg;
      h->cta_mode = 4;
    }
  });
}

}  // extern "C"
