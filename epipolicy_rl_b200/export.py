"""Build an `EpiDesc` from a live reference `Epidemic` (the reference package must be importable).

Per BASELINE.json's north_star the reference's JSON scenario parser, `make_static`
(core/core_utils.py:362-396) and the intervention objects stay as they are; this module is
the bridge from their output (a numba `Static` jitclass) to the flat tables the B200 engine
consumes. It runs once per scenario at construction time, never per step.

Two parts:
  * `export_static`  — tables, edges, mobility patterns, initial state.
  * `compile_programs` — runs every JSON `effect`/`cost`/`func` source ONCE against a
    recording `sim` proxy (symbolic control parameters) to obtain a static op table, instead
    of `exec`-ing Python and re-parsing regexes every simulated day (core/executor.py:272-284,
    :111-114). Anything outside the op algebra raises `UnsupportedProgram` rather than
    silently diverging.
"""
from __future__ import annotations

import numpy as np

from . import desc as D
from .desc import EpiDesc


class UnsupportedProgram(Exception):
    pass


# ---------------------------------------------------------------------------
# symbolic expressions with numpy (NEP 50) scalar promotion
# ---------------------------------------------------------------------------
class Sym:
    """A recorded scalar expression. dtype: 'w' weak python number, 'f32', 'f64'."""

    __slots__ = ("toks", "dt")

    def __init__(self, toks, dt):
        self.toks, self.dt = toks, dt

    @staticmethod
    def lift(x):
        if isinstance(x, Sym):
            return x
        if isinstance(x, (np.float32,)):
            return Sym([(D.TOK_CONST, 0, 1, float(x))], "f32")
        if isinstance(x, (np.floating,)):
            return Sym([(D.TOK_CONST, 0, 0, float(x))], "f64")
        if isinstance(x, (int, float, np.integer)):
            return Sym([(D.TOK_CONST, 0, 0, float(x))], "w")
        raise UnsupportedProgram(f"operand of type {type(x)} in intervention code")

    def _bin(self, other, op, swap=False):
        a, b = self, Sym.lift(other)
        if swap:
            a, b = b, a
        if a.dt == "f64" or b.dt == "f64":
            dt = "f64"
        elif a.dt == "f32" or b.dt == "f32":
            dt = "f32"
        else:
            dt = "w"
        if dt == "w":  # python-number folding
            va, vb = a.toks[0][3], b.toks[0][3]
            v = {D.TOK_ADD: va + vb, D.TOK_SUB: va - vb, D.TOK_MUL: va * vb, D.TOK_DIV: va / vb}[op]
            return Sym([(D.TOK_CONST, 0, 0, v)], "w")

        def cast(s):
            # a weak python constant adopts the other operand's dtype
            if s.dt == "w" and dt == "f32":
                return [(D.TOK_CONST, 0, 1, float(np.float32(s.toks[0][3])))]
            return s.toks

        return Sym(cast(a) + cast(b) + [(op, 0, 1 if dt == "f32" else 0, 0.0)], dt)

    def __add__(self, o): return self._bin(o, D.TOK_ADD)
    def __radd__(self, o): return self._bin(o, D.TOK_ADD, True)
    def __sub__(self, o): return self._bin(o, D.TOK_SUB)
    def __rsub__(self, o): return self._bin(o, D.TOK_SUB, True)
    def __mul__(self, o): return self._bin(o, D.TOK_MUL)
    def __rmul__(self, o): return self._bin(o, D.TOK_MUL, True)
    def __truediv__(self, o): return self._bin(o, D.TOK_DIV)
    def __rtruediv__(self, o): return self._bin(o, D.TOK_DIV, True)

    def __neg__(self):
        return Sym(self.toks + [(D.TOK_NEG, 0, 1 if self.dt == "f32" else 0, 0.0)], self.dt)

    def _no(self, *a, **k):
        raise UnsupportedProgram("intervention code branches on / transforms a control value "
                                 "(only + - * / of control params, constants and select-sums are compiled)")

    __bool__ = __lt__ = __le__ = __gt__ = __ge__ = __eq__ = __ne__ = _no
    __pow__ = __rpow__ = __abs__ = __float__ = __int__ = __round__ = _no
    __hash__ = None


class _SelProxy:
    """What `sim.select(...)` returns; only `['Value'].sum()` is compiled (Appendix B.3)."""

    def __init__(self, rec, sel_id):
        self.rec, self.sel_id = rec, sel_id

    def __getitem__(self, key):
        if key != "Value":
            raise UnsupportedProgram(f"select(...)[{key!r}]")
        return self

    def sum(self):
        return Sym([(D.TOK_SEL, self.sel_id, 0, 0.0)], "f64")

    def __getattr__(self, name):
        raise UnsupportedProgram(f"select(...).{name}")


class _Lists:
    def __init__(self):
        self.off, self.data, self.memo = [0], [], {}

    def add(self, seq) -> int:
        t = tuple(int(x) for x in seq)
        if t not in self.memo:
            self.memo[t] = len(self.off) - 1
            self.data.extend(t)
            self.off.append(len(self.data))
        return self.memo[t]


class _Recorder:
    """The `sim` object seen by the JSON code while a program is being recorded.
    Selector resolution is the reference's own (`get_regsult`, executor.py:111-114)."""

    def __init__(self, epi, lists, itv_index, cp_base):
        from epipolicy.core.executor import get_regsult
        self._get_regsult = get_regsult
        self.epi, self.static, self.lists = epi, epi.static, lists
        self.itv, self.cp_base = itv_index, cp_base
        self.ops, self.sels, self.exprs = [], [], []

    def _res(self, func, regex):
        sd = self.epi.executor.to_string_dict_regex(regex)
        return self._get_regsult(self.static, self.epi.parser, self.itv, func, sd)

    def _expr(self, v) -> int:
        self.exprs.append(Sym.lift(v))
        return len(self.exprs) - 1

    @staticmethod
    def _loc(res, key):
        return [int(k) for k in res.locale_result[key].keys()]

    @staticmethod
    def _idx(res, key):
        return [int(k) for k in res.index_result[key]]

    # sim.select (executor.py:292-370): only the compartment shape is compiled
    def select(self, regex):
        res = self._res("select", regex)
        mode = "current"
        if "value-mode" in res.string_result:
            mode = str(res.string_result["value-mode"])
        lr, ir = res.locale_result, res.index_result
        if ("locale-from" in lr and "locale-to" in lr) or "parameter" in ir or \
                ("group-from" in ir and "group-to" in ir) or "compartment-from" not in ir:
            raise UnsupportedProgram("select() over parameters/mobility/facilities")
        vm = {"current": D.SEL_CURRENT, "default": D.SEL_DEFAULT, "change": D.SEL_CHANGE}.get(mode)
        if vm is None:
            raise UnsupportedProgram(f"value-mode {mode!r}")
        self.sels.append((self.lists.add(self._idx(res, "compartment-from")),
                          self.lists.add(self._loc(res, "locale-from")),
                          self.lists.add(self._idx(res, "group-from")), vm))
        return _SelProxy(self, len(self.sels) - 1)

    # sim.apply (executor.py:372-374, nb_apply :116-164)
    def apply(self, regex, multiplier):
        res = self._res("apply", regex)
        lr, ir, L = res.locale_result, res.index_result, self.lists
        e = self._expr(multiplier)
        if "locale-from" in lr and "locale-to" in lr:
            mode = self.static.get_property_index("mode", "border")
            if "mode" in res.string_result:
                mode = self.static.get_property_index("mode", str(res.string_result["mode"]))
            frm, to = self._loc(res, "locale-from"), self._loc(res, "locale-to")
            if len(frm) == 0 or len(to) == 0:
                return  # e.g. '~' + '*' parses to the empty set (regex_parser.py:145-170): no-op
            self.ops.append((D.OP_MOB_MUL, e, [int(mode), L.add(self._idx(res, "group-from")), L.add(frm), L.add(to)]))
        elif "parameter" in ir:
            self.ops.append((D.OP_PARAM_MUL, e, [L.add(self._idx(res, "parameter")), L.add(self._loc(res, "locale-from")),
                                                  L.add(self._idx(res, "facility")), L.add(self._idx(res, "group-from"))]))
        elif "group-from" in ir and "group-to" in ir:
            self.ops.append((D.OP_FI_MUL, e, [L.add(self._loc(res, "locale-from")), L.add(self._idx(res, "facility")),
                                               L.add(self._idx(res, "group-from")), L.add(self._idx(res, "group-to"))]))
        elif "facility" in ir:
            self.ops.append((D.OP_FT_MUL, e, [L.add(self._loc(res, "locale-from")), L.add(self._idx(res, "facility")),
                                               L.add(self._idx(res, "group-from"))]))

    # sim.move (executor.py:376-379, nb_move :166-234)
    def move(self, regex1, regex2, value):
        r1, r2, L = self._res("move", regex1), self._res("move", regex2), self.lists
        e = self._expr(value)
        self.ops.append((D.OP_MOVE, e, [L.add(self._idx(r1, "compartment-from")), L.add(self._loc(r1, "locale-from")),
                                         L.add(self._idx(r1, "group-from")), L.add(self._idx(r2, "compartment-from")),
                                         L.add(self._loc(r2, "locale-from")), L.add(self._idx(r2, "group-from"))]))

    # sim.add (executor.py:381-383, nb_add :236-243)
    def add(self, regex, value):
        res = self._res("add", regex)
        if "intervention" not in res.index_result:
            return
        e = self._expr(value)
        self.ops.append((D.OP_ADD, e, [self.lists.add(self._idx(res, "intervention")),
                                        self.lists.add(self._loc(res, "locale-from"))]))


def _record(epi, lists, itv_index, cp_base, sources, locale_regex):
    """Run the effect/cost sources of one intervention against a recorder."""
    static = epi.static
    rec = _Recorder(epi, lists, itv_index, cp_base)
    cp = {}
    for k, c in enumerate(static.interventions[itv_index].cp_list):
        cp[str(c.name)] = Sym([(D.TOK_CP, cp_base + k, 1, 0.0)], "f32")
    for fname, src, takes_locales in sources:
        ns = {"sim": rec, "__builtins__": {"len": len, "range": range, "float": float, "int": int,
                                           "min": min, "max": max, "abs": abs, "sum": sum}}
        try:
            exec(src, ns)
            fn = ns[fname]
            fn(cp, locale_regex) if takes_locales else fn(cp)
        except UnsupportedProgram:
            raise
        except Exception as ex:  # noqa: BLE001
            raise UnsupportedProgram(f"intervention {static.interventions[itv_index].name!r}: {ex!r}") from ex
    return rec


def compile_programs(epi, out_arrays, lists):
    """Record one program per (intervention, locale regex) and flatten into EpiDesc arrays."""
    static, session = epi.static, epi.session
    I = static.intervention_count
    itv_cp_off = [0]
    for itv in static.interventions:
        itv_cp_off.append(itv_cp_off[-1] + len(itv.cp_list))
    src_of = {}
    for itv in session["interventions"]:
        i = static.get_property_index("intervention", itv["name"])
        src_of[i] = [("effect", itv["effect"], True), ("cost", itv["cost"], True)]
    for c in session["costs"]:
        i = static.get_property_index("intervention", c["name"])
        src_of[i] = [("cost", c["func"], False)]

    # schedule day-0 action (used by the detached initial state, epidemic.py:77-80)
    init_active = np.zeros(I, np.int32)
    init_cpv = np.zeros(itv_cp_off[-1], np.float32)
    init_regex = {}
    for k, c in enumerate(c for itv in static.interventions for c in itv.cp_list):
        init_cpv[k] = c.default_value
    for act in static.schedule.action_list[0]:
        i = int(act.index)
        if init_active[i]:
            raise UnsupportedProgram("two schedule entries for one intervention on day 0")
        init_active[i] = 1
        init_regex[i] = str(act.locale_regex)
        init_cpv[itv_cp_off[i]:itv_cp_off[i + 1]] = np.asarray(act.cpv_list, np.float32)

    op_kind, op_itv, op_expr, op_arg = [], [], [], []
    sel_arg, expr_off, toks = [], [0], []
    prog_off, memo = [0], {}

    def emit(i, regex):
        rec = _record(epi, lists, i, itv_cp_off[i], src_of[i], regex)
        key = (i, tuple((k, tuple(a)) for k, _, a in rec.ops), tuple(rec.sels),
               tuple(tuple(e.toks) for e in rec.exprs))
        if key in memo:
            return memo[key]
        sel_base, expr_base = len(sel_arg), len(expr_off) - 1
        sel_arg.extend(rec.sels)
        for e in rec.exprs:
            for (op, arg, f32, val) in e.toks:
                toks.append((op, arg + sel_base if op == D.TOK_SEL else arg, f32, val))
            expr_off.append(len(toks))
        for kind, e, args in rec.ops:
            op_kind.append(kind)
            op_itv.append(i)
            op_expr.append(e + expr_base)
            op_arg.append(list(args) + [-1] * (8 - len(args)))
        prog_off.append(len(op_kind))
        memo[key] = len(prog_off) - 2
        return memo[key]

    prog_step = np.full(I, -1, np.int32)
    prog_init = np.full(I, -1, np.int32)
    for i in range(I):
        prog_step[i] = emit(i, "*")
        prog_init[i] = emit(i, init_regex[i]) if i in init_regex else prog_step[i]

    out_arrays.update(
        op_kind=op_kind, op_itv=op_itv, op_expr=op_expr, op_arg=np.array(op_arg, np.int32).reshape(-1),
        sel_arg=np.array(sel_arg, np.int32).reshape(-1), expr_off=expr_off,
        tok_op=[t[0] for t in toks], tok_arg=[t[1] for t in toks], tok_f32=[t[2] for t in toks],
        tok_val=[t[3] for t in toks], prog_off=prog_off, prog_step=prog_step, prog_init=prog_init,
        init_active=init_active, init_cpv=init_cpv, itv_cp_off=itv_cp_off,
    )


def export_static(epi) -> EpiDesc:
    """Flatten a reference `Epidemic` (core/epidemic.py:34-75) into an EpiDesc."""
    from epipolicy.utility.singleton import DIE_TAG, HOS_TAG, ALIVE_COMPARTMENT

    st = epi.static
    C, L, G, F, P, I, M = (int(st.compartment_count), int(st.locale_count), int(st.group_count),
                           int(st.facility_count), int(st.parameter_count), int(st.intervention_count),
                           int(st.mode_count))
    a = {}
    alive = np.ones(C)
    for tag in (DIE_TAG, HOS_TAG):
        if st.has_property_name("compartment_tag", tag):
            ti = st.get_property_index("compartment_tag", tag)
            alive *= 1 - np.asarray(st.compartment_tags)[:, ti]
    a["alive_coef"] = alive
    unobs = st.default_state.unobs
    a["X0"] = np.array(st.default_state.obs.current_comp)
    a["pop"] = np.array(st.locale_group_pop)
    a["mp0"] = np.array(unobs.model_parameters)
    a["ft0"] = np.array(unobs.facility_timespent)
    a["fi0"] = np.array(unobs.facility_interaction)
    a["mode_bias"] = np.array(st.mode_bias)

    mode_indptr, mode_off, mode_indices, mode_origin = [], [0], [], []
    for m in range(M):
        csr = st.mode_csr[m]
        indptr, indices = np.array(csr.indptr), np.array(csr.indices)
        mode_indptr.append(indptr)
        mode_indices.append(indices)
        mode_off.append(mode_off[-1] + len(indices))
        org = np.zeros((G, len(indices)), np.float32)
        for g in range(G):
            coo = st.mode_group_coo[m][g]
            for r in range(L):
                for k in range(indptr[r], indptr[r + 1]):
                    org[g, k] = coo.get(r, int(indices[k]))
        mode_origin.append(org.reshape(-1))
        # sanity: the reference's default reduced CSR is the same numbers
        assert np.array_equal(org, np.array(unobs.mode_reduced_csr[m]))
    a["mode_indptr"] = np.concatenate(mode_indptr)
    a["mode_off"] = mode_off
    a["mode_indices"] = np.concatenate(mode_indices) if mode_indices else []
    a["mode_origin"] = np.concatenate(mode_origin) if mode_origin else []
    a["csr_indptr"] = np.array(st.sum_mode_csr.indptr)
    a["csr_indices"] = np.array(st.sum_mode_csr.indices)
    a["csc_indptr"] = np.array(st.sum_mode_csc.indptr)
    a["csc_indices"] = np.array(st.sum_mode_csc.indices)

    # ---- edges -----------------------------------------------------------------
    fac, sum_coef, sum_fac_off = [], [], [0]
    edge_numer_off, edge_den_off = [0], [0]
    sc_off, sc_kind, sc_a, sc_b, sc_coef = [0], [], [], [], []

    def factors(muls):
        out = []
        for var in muls:
            var = str(var)
            # same precedence as compute_summant (core_optimize.py:46-52); unknown names are ignored there
            if st.has_property_name("parameter", var):
                out.append((D.FAC_PARAM << 24) | int(st.get_property_index("parameter", var)))
            elif var == ALIVE_COMPARTMENT:
                out.append(D.FAC_N << 24)
            elif st.has_property_name("compartment", var):
                out.append((D.FAC_COMP << 24) | int(st.get_property_index("compartment", var)))
        return out

    def scatter(edge):
        for c, coef in edge.compartment_map.items():
            sc_kind.append(D.SC_COMP); sc_a.append(int(c)); sc_b.append(-1); sc_coef.append(np.float32(coef))
        for c, coef in edge.gain_map.items():
            sc_kind.append(D.SC_GAIN); sc_a.append(int(c)); sc_b.append(-1); sc_coef.append(np.float32(coef))
        for c, coef in edge.lost_map.items():
            sc_kind.append(D.SC_LOST); sc_a.append(int(c)); sc_b.append(-1); sc_coef.append(np.float32(coef))
        for c1 in edge.move_map:
            for c2, coef in edge.move_map[c1].items():
                sc_kind.append(D.SC_MOVE); sc_a.append(int(c1)); sc_b.append(int(c2)); sc_coef.append(np.float32(coef))
        sc_off.append(len(sc_kind))

    # numerator factors of all edges first, then the summants' factors (both ranges index `fac`)
    for edge in st.edges.values():
        fac.extend(factors(edge.fraction.numer.muls))
        edge_numer_off.append(len(fac))
    sum_fac_off = [len(fac)]
    for edge in st.edges.values():
        for s in edge.fraction.denom:
            sum_coef.append(np.float32(s.coef))
            fac.extend(factors(s.muls))
            sum_fac_off.append(len(fac))
        edge_den_off.append(len(sum_coef))
        scatter(edge)
    ie_p0, ie_c0, ie_c2, ie_np_off, ie_np, ie_dp_off, ie_dp = [], [], [], [0], [], [0], []
    for ie in st.infectious_edges.values():
        ie_p0.append(int(ie.transmission_rate_index))
        ie_c0.append(int(ie.susceptible_comp_index))
        ie_c2.append(int(ie.infectious_comp_index))
        ie_np.extend(int(x) for x in ie.numer_param_index_list)
        ie_np_off.append(len(ie_np))
        ie_dp.extend(int(x) for x in ie.denom_param_index_list)
        ie_dp_off.append(len(ie_dp))
        scatter(ie)
    a.update(edge_numer_off=edge_numer_off, edge_den_off=edge_den_off, sum_coef=sum_coef,
             sum_fac_off=sum_fac_off, fac=fac, ie_p0=ie_p0, ie_c0=ie_c0, ie_c2=ie_c2,
             ie_np_off=ie_np_off, ie_np=ie_np, ie_dp_off=ie_dp_off, ie_dp=ie_dp,
             sc_off=sc_off, sc_kind=sc_kind, sc_a=sc_a, sc_b=sc_b, sc_coef=sc_coef)

    # ---- interventions ----------------------------------------------------------
    a["itv_is_cost"] = [1 if itv.is_cost else 0 for itv in st.interventions]
    cps = [c for itv in st.interventions for c in itv.cp_list]
    a["cp_default"] = [c.default_value for c in cps]
    a["cp_min"] = [c.min_value for c in cps]
    a["cp_max"] = [c.max_value for c in cps]
    n_act = 0
    for itv in st.interventions:
        if not itv.is_cost:
            n_act += sum(1 for c in itv.cp_list if c.min_value != c.max_value)

    lists = _Lists()
    compile_programs(epi, a, lists)
    a["list_off"], a["list_data"] = lists.off, lists.data

    d = EpiDesc()
    d.arrays = a
    d.scalars = dict(C=C, L=L, G=G, F=F, P=P, I=I, A=n_act, NCP=len(cps), M=M,
                     horizon=int(st.schedule.horizon), E=len(st.edges), IE=len(st.infectious_edges))
    for n, _ in D.SCALARS:
        d.scalars.setdefault(n, 0)
    names = lambda prop, n: [str(st.get_property_name(prop, i)) for i in range(n)]  # noqa: E731
    d.meta = dict(
        compartments=names("compartment", C), locales=names("locale", L), groups=names("group", G),
        facilities=names("facility", F), parameters=names("parameter", P),
        interventions=[str(itv.name) for itv in st.interventions],
        control_params=[[str(c.name) for c in itv.cp_list] for itv in st.interventions],
        total_population=float(np.sum(a["X0"])),
        n_action_interventions=sum(1 for itv in st.interventions if not itv.is_cost),
        reporter=reporter_meta(st),
    )
    return d.finalize()


def reporter_meta(st, session=None) -> dict:
    """What reporter.py / reporter_device.py need from the Static (and, for the reporting locations, the session):
    the hashed incidence edges (core_utils.py:243-254) as (c1, c2) pairs - sorted, and in the reference dict's own
    iteration order -, the INF / SUS / DIE compartment tag weights (get_matrix_with_tag, core_optimize.py:10-16), the
    infectious edges with their parameter lists (obj/edge.py:51-67), the leaf locales of every reporting location
    (session["locales"] through static.locale_hierarchy, core/reporter.py:102-103) and the generation-interval
    weights of utility/utils.py:9-19."""
    from epipolicy.utility.singleton import DIE_TAG, INF_TAG, SUS_TAG
    from epipolicy.utility.utils import Ws, get_normalized_string
    C = st.compartment_count

    def tagw(tag):
        if st.has_property_name("compartment_tag", tag):
            ti = st.get_property_index("compartment_tag", tag)
            return [float(st.compartment_tags[c, ti]) for c in range(C)]
        return [0.0] * C
    meta = {"incidence_edges": [[int(h) // C, int(h) % C] for h in sorted(int(h) for h in st.hashed_incidence_edges)],
            "incidence_edges_order": [[int(h) // C, int(h) % C] for h in st.hashed_incidence_edges],
            "inf_comps": tagw(INF_TAG), "sus_comps": tagw(SUS_TAG), "die_comps": tagw(DIE_TAG),
            "has_die_tag": bool(st.has_property_name("compartment_tag", DIE_TAG)),
            "infectious_edges": [{"sus": int(e.susceptible_comp_index), "beta": int(e.transmission_rate_index),
                                  "comps": [[int(c), float(v)] for c, v in e.compartment_map.items()],
                                  "numer": [int(p) for p in e.numer_param_index_list],
                                  "denom": [int(p) for p in e.denom_param_index_list]}
                                 for e in st.infectious_edges.values()],
            "infectivity": [float(w) for w in Ws]}
    if session is not None:
        sess = session.get("session", session)
        meta["locations"] = [sorted(int(l) for l in st.locale_hierarchy[get_normalized_string(loc["name"])])
                             for loc in sess["locales"]]
        meta["location_names"] = [loc["name"] for loc in sess["locales"]]
    return meta


def export_initial_parameter(epi) -> dict:
    """The reference's detached initial `Parameter` literal (epidemic.py:77-80), for validating
    the restated executor/parameter build (not part of EpiDesc)."""
    st = epi.static
    un = st.schedule.initial_state.unobs
    moves = []
    for g in un.delta_comp_to_comp:
        for c1 in un.delta_comp_to_comp[g]:
            for c2 in un.delta_comp_to_comp[g][c1]:
                coo = un.delta_comp_to_comp[g][c1][c2]
                for l1 in coo.mat:
                    for l2, v in coo.mat[l1].items():
                        moves.append((int(g), int(c1), int(c2), int(l1), int(l2), float(v)))
    return dict(
        init_mp=np.array(un.model_parameters), init_ft=np.array(un.facility_timespent),
        init_fi=np.array(un.facility_interaction),
        init_mob=np.concatenate([np.array(m).reshape(-1) for m in un.mode_reduced_csr]),
        init_cost=np.array(un.delta_cost), init_moves=np.array(moves, np.float64).reshape(-1, 6),
    )
