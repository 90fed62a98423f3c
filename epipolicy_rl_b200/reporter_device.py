"""The reference Reporter's epidemiological quantities for ALL environments of a batch, computed on the device from
what the engine already carries: average infectious duration, windowed incidence / infectious
counts, instantaneous, basic and case reproduction numbers per reporting location (core/reporter.py:54-80, :132-204),
on top of the incidences of :45-51.

Inputs are torch tensors recorded day by day by `BatchedEpidemic.run(..., report=True)` with stream-ordered device
copies (`epi_get_state_dev`): the state X, the live cumulative_move components (the accumulators the RK45 error norm
needs anyway) and the class multipliers that define a day's parameters. Nothing is copied to the host; every
expression is batched over the env axis. torch CPU tensors work the same (tests on the host emulation).

Restated literally, including two things worth knowing about the reference:
  * `compute_basic_reproductive_number` ("beta_based", the default) multiplies its per-location transmission-rate row
    IN PLACE: `add = self.location_avg_model_parameters[t, idx]` is a view, `add *= ...` and `true_divide(add, ..)`
    (out = a) write through it, so later edges of the same day see the already scaled row (core/reporter.py:181-186).
    The restatement keeps the aliasing.
  * `location_cumulative_moves[t, :, :, loc]` is assigned, not accumulated, in the loop over the location's locales
    (core/reporter.py:112): the basic R's edge weights see the LAST leaf locale only. Kept.
  * the windowed sums are running sums of window differences (core/reporter.py:132-150), kept as written.

`desc.meta["reporter"]` holds what comes from the reference Static / session (export.reporter_meta): incidence
edges, INF / SUS / DIE compartment tags, the reporting locations (session["locales"] -> leaf locale lists), the
infectious edges with their parameter lists, and the generation-interval weights of utility/utils.py:9-19.
"""
from __future__ import annotations

import numpy as np
import torch

R_WINDOW = 7  # utility/singleton.py:11


def _meta(desc):
    m = desc.meta.get("reporter")
    if m is None or "locations" not in m:
        raise ValueError("desc.meta['reporter'] lacks the reporter tables: re-export the scenario with "
                         "tools/gen_reporter_golden.py (needs the reference importable)")
    return m


def true_divide(a, b):
    """utility/utils.py:57-60 with out = a: a / b where b != 0, a unchanged elsewhere"""
    return torch.where(b != 0, a / torch.where(b != 0, b, torch.ones_like(b)), a)


class DeviceReporter:
    """day-by-day recorder + the Reporter's derived quantities. All tensors [T+1, N, ...] on the engine's device."""

    def __init__(self, desc, plan: dict, device, dtype=torch.float64):
        self.desc, self.dev, self.dt = desc, torch.device(device), dtype
        m = _meta(desc)
        C, L, G = desc.C, desc.L, desc.G
        self.inf = torch.tensor(m["inf_comps"], dtype=dtype, device=self.dev)
        self.sus = torch.tensor(m["sus_comps"], dtype=dtype, device=self.dev)
        self.die = torch.tensor(m["die_comps"], dtype=dtype, device=self.dev)
        self.inc_edges = [tuple(e) for e in m["incidence_edges_order"]]      # insertion order of the reference dict
        self.locations = [list(l) for l in m["locations"]]                     # leaf locales per reporting location
        self.inf_edges = m["infectious_edges"]                                 # dicts: sus, comps{c: coef}, beta, numer, denom
        self.Ws = torch.tensor(m["infectivity"], dtype=dtype, device=self.dev)
        nloc = len(self.locations)
        M = torch.zeros((nloc, L), dtype=dtype, device=self.dev)
        for i, ls in enumerate(self.locations):
            for l in ls:
                M[i, l] += 1.0
        self.M = M
        self.M_last = torch.zeros_like(M)
        for i, ls in enumerate(self.locations):
            if ls:
                self.M_last[i, ls[-1]] = 1.0   # leaf locales are inserted in ascending order (core_utils.py:44-53)
        # position of every live accumulator in the dense layouts
        self.acc_flat = np.asarray(plan["acc_flat"], np.int64)                # offset / (L*G) in the flat observable
        self.plan = {k: torch.as_tensor(np.asarray(v), device=self.dev) for k, v in plan.items() if k != "acc_flat"}
        self.X, self.acc, self.mult = [], [], []

    # ---- recording --------------------------------------------------------------------------------
    def record(self, X, acc, mult):
        """state after day t (call once after reset for t = 0): X [N, C*L*G] f64, acc [N, n_acc, L*G] real,
        mult [N, n_cls] f32 — device tensors, cloned here"""
        d = self.desc
        self.X.append(X.to(self.dt).reshape(-1, d.C, d.L, d.G).clone())
        self.acc.append(acc.to(self.dt).reshape(acc.shape[0], -1, d.L, d.G).clone())
        self.mult.append(mult.clone())

    # ---- pieces -----------------------------------------------------------------------------------
    def _cumulative_move(self, acc):
        """dense cumulative_move [.., C, C, L, G] from the live components (structural zeros stay zero)"""
        d = self.desc
        C, L, G = d.C, d.L, d.G
        mv = torch.zeros(acc.shape[:-3] + (C * C, L, G), dtype=self.dt, device=self.dev)
        for a, off in enumerate(self.acc_flat):
            k = int(off) - 3 * C  # flat order: current_comp, gain, lost [C each], then move [C*C]
            if k >= 0:
                mv[..., k, :, :] = acc[..., a, :, :]
        return mv.reshape(acc.shape[:-3] + (C, C, L, G))

    def _parameters(self, mult):
        """a state's model_parameters [N, n_lc, P, F, G] and normalised facility_timespent [N, n_lc, F, G] in float32
        with the reference's rounding (obj/parameter.py:53-66, matrix/dense.py:6-19)"""
        d, pl = self.desc, self.plan
        n_lc = pl["mp0"].numel() // (d.P * d.F * d.G)
        mp = (pl["mp0"].float()[None] * mult[:, pl["mp_cls"].long()]).reshape(-1, n_lc, d.P, d.F, d.G)
        ft = (pl["ft0"].float()[None] * mult[:, pl["ft_cls"].long()]).reshape(-1, n_lc, d.F, d.G)
        s = torch.zeros_like(ft[:, :, 0])
        for f in range(d.F):
            s = s + ft[:, :, f]  # sequential float32 sum over the facilities
        first = torch.zeros_like(ft)
        first[:, :, 0] = 1.0
        ftn = torch.where((s > 1e-15)[:, :, None], ft / torch.where(s > 1e-15, s, torch.ones_like(s))[:, :, None], first)
        return mp, ftn

    # ---- the Reporter chain -------------------------------------------------------------------------
    def compute(self) -> dict:
        d = self.desc
        C, L, G = d.C, d.L, d.G
        X = torch.stack(self.X)        # [T+1, N, C, L, G]
        acc = torch.stack(self.acc)    # [T+1, N, n_acc, L, G]
        T1, N = X.shape[0], X.shape[1]
        mv = self._cumulative_move(acc)
        tag = lambda x, w: torch.einsum("c,...clg->...lg", w, x)  # get_matrix_with_tag  # noqa: E731
        # incidences (core/reporter.py:45-48)
        tot = torch.zeros((T1, N, L, G), dtype=self.dt, device=self.dev)
        for c1, c2 in self.inc_edges:
            tot = tot + mv[:, :, c1, c2] - mv[:, :, c2, c1]
        tot = tot + tag(X[0], self.inf)[None]
        inc = tot.clone()
        inc[1:] = tot[1:] - tot[:-1]
        # average infectious duration, "instantaneous_based" (core/reporter.py:54-62)
        inf_t = tag(X, self.inf).sum((-1, -2))                       # [T+1, N]
        cum_inf, cum_inc = torch.cumsum(inf_t, 0), torch.cumsum(inc.sum((-1, -2)), 0)
        avg_dur = torch.where(cum_inc > 0, cum_inf / torch.where(cum_inc > 0, cum_inc, torch.ones_like(cum_inc)),
                              torch.zeros_like(cum_inf))
        # location statistics (core/reporter.py:82-130; the pieces the R numbers use)
        loc_inc = torch.einsum("il,tnlg->tnig", self.M, inc)          # [T+1, N, nloc, G]
        loc_cnt = torch.einsum("il,tnclg->tncig", self.M, X)          # [T+1, N, C, nloc, G]
        # location_cumulative_moves: the reference ASSIGNS inside its loop over the location's locales
        # (core/reporter.py:112, `=` not `+=`): what remains is the last leaf locale's moves, summed over the groups
        loc_mv = torch.einsum("il,tnabl->tnabi", self.M_last, mv.sum(-1))  # [T+1, N, C, C, nloc]
        alive = X.sum(2) - tag(X, self.die)                           # get_matrix_without_tag(.., DIE)  [T+1, N, L, G]
        lclass = self.plan["lclass"].long()
        n_lc = int(lclass.max().item()) + 1
        onehot = torch.zeros((n_lc, L), dtype=self.dt, device=self.dev)
        onehot[lclass, torch.arange(L, device=self.dev)] = 1.0
        # A[t, n, lc, loc, g] = alive summed over the location's locales of class lc
        A = torch.einsum("kl,il,tnlg->tnkig", onehot, self.M, alive)
        loc_mp = torch.zeros((T1, N, d.P, len(self.locations)), dtype=self.dt, device=self.dev)
        loc_mp_cnt = torch.zeros((T1, N, len(self.locations)), dtype=self.dt, device=self.dev)
        for t in range(T1):
            mp, ftn = self._parameters(self.mult[t])
            # sum_{lc,f,g} mp[lc,p,f,g] * ft[lc,f,g] * A[lc,loc,g]
            loc_mp[t] = torch.einsum("nkpfg,nkfg,nkig->npi", mp.to(self.dt), ftn.to(self.dt), A[t])
            loc_mp_cnt[t] = torch.einsum("nkfg,nkig->ni", ftn.to(self.dt), A[t])
        loc_mp = true_divide(loc_mp, loc_mp_cnt[:, :, None, :].expand_as(loc_mp))
        # window statistics (core/reporter.py:132-150)
        nloc = len(self.locations)
        w_inc = torch.zeros((T1, N, nloc), dtype=self.dt, device=self.dev)
        w_cnt = torch.zeros_like(w_inc)
        w_dur = torch.zeros((T1, N), dtype=self.dt, device=self.dev)
        loc_inf = torch.einsum("c,tncig->tnig", self.inf, loc_cnt)
        for t in range(T1):
            if t >= R_WINDOW:
                w_inc[t] = (loc_inc[t] - loc_inc[t - R_WINDOW]).sum(-1)
                w_cnt[t] = (loc_inf[t] - loc_inf[t - R_WINDOW]).sum(-1)
                w_dur[t] = avg_dur[t] - avg_dur[t - R_WINDOW]
            else:
                w_inc[t] = loc_inc[t].sum(-1)
                w_cnt[t] = loc_inf[t].sum(-1)
                w_dur[t] = avg_dur[t]
            if t > 0:
                w_inc[t] += w_inc[t - 1]; w_cnt[t] += w_cnt[t - 1]; w_dur[t] += w_dur[t - 1]
        # instantaneous R, "infectious_duration_based" (core/reporter.py:152-158)
        inst = torch.zeros_like(w_inc)
        for t in range(T1):
            coef = (w_dur[t] / min(t + 1, R_WINDOW))[:, None]
            q = torch.where(w_cnt[t] != 0, w_inc[t] * coef / torch.where(w_cnt[t] != 0, w_cnt[t], torch.ones_like(w_cnt[t])),
                            torch.zeros_like(w_cnt[t]))
            inst[t] = torch.clamp(q, min=0)
        # basic R, "beta_based" (core/reporter.py:167-191), with the reference's in-place row update
        basic = torch.zeros_like(w_inc)
        for t in range(T1):
            rows = loc_mp[t].clone()                       # [N, P, nloc]: written through, as the reference does
            final_beta = torch.zeros((N, nloc), dtype=self.dt, device=self.dev)
            sum_w = torch.zeros_like(final_beta)
            for c1, c2 in self.inc_edges:
                weight = 1 + loc_mv[t][:, c1, c2]
                beta = torch.zeros_like(final_beta)
                for e in self.inf_edges:
                    for comp, coef in e["comps"]:
                        if e["sus"] == c1 and comp == c2:
                            add = rows[:, e["beta"]]        # a VIEW of the transmission-rate row
                            for p in e["numer"]:
                                add *= rows[:, p]
                            for p in e["denom"]:
                                add.copy_(true_divide(add, rows[:, p]))
                            beta = beta + add * coef
                final_beta = final_beta + beta * weight
                sum_w = sum_w + weight
            final_beta = true_divide(final_beta, sum_w)
            basic[t] = final_beta * w_dur[t][:, None] / min(t + 1, R_WINDOW)
        # case R, "infectivity_based" (core/reporter.py:198-204)
        case = torch.zeros_like(w_inc)
        nW = self.Ws.shape[0]
        for t in range(T1):
            hi = min(T1, t + nW)
            case[t] = (inst[t:hi] * self.Ws[:hi - t, None, None]).sum(0)
        return {"incidences": inc, "average_infectious_duration": avg_dur, "location_incidences": loc_inc,
                "location_counts": loc_cnt, "location_avg_model_parameters": loc_mp,
                "window_infectious_duration": w_dur, "location_window_incidence": w_inc,
                "location_window_infectious_count": w_cnt, "location_instant_R": inst, "location_basic_R": basic,
                "location_case_R": case}
