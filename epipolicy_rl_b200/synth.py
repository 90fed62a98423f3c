"""Synthetic COVID-like scenarios in the reference's own JSON session schema (SURVEY.md Appendix F,
BASELINE.json configs[3]): L leaf locales x G groups x F facilities on top of COVID_C's model,
interventions, costs, schedules and optimize ranges. The session is consumed by the reference's unmodified
`construct_epidemic` (tools/gen_syn.py exports the resulting Static to epipolicy_rl_b200/scenarios/desc_syn_*.npz), so the
GPU path and the oracle see exactly what the reference's parser / make_static produce for it."""
from __future__ import annotations

import copy
import os

import numpy as np

SIZES = {"syn_s": (20, 4, 8), "syn_m": (100, 8, 8), "syn": (1000, 16, 8)}
_HERE = os.path.dirname(os.path.abspath(__file__))


def make_session(covid_c_session: dict, L: int, G: int, F: int, seed: int = 20260101) -> dict:
    """COVID_C session with locales / groups / facilities / mobility replaced (Appendix F recipe)."""
    rng = np.random.default_rng(seed)
    s = copy.deepcopy(covid_c_session.get("session", covid_c_session))
    pop = np.maximum(rng.lognormal(np.log(1e5), 1.0, L), 1000.0).round()
    s["locales"] = [{"id": "UP", "name": "UnitedProvinces", "parent_id": "", "population": float(pop.sum())}] + [
        {"id": f"UP.L{i:04d}", "name": f"UnitedProvinces.L{i:04d}", "parent_id": "UP", "population": float(pop[i])}
        for i in range(L)]
    gnames = ["Children", "Adults", "Seniors"] + [f"Group{g}" for g in range(3, G)]
    frac = rng.dirichlet(2.0 * np.ones(G))
    s["groups"] = [{"name": gnames[g], "locales": [{"id": f"UP.L{i:04d}", "population": float(frac[g])} for i in range(L)]}
                   for g in range(G)]
    fnames = ["Household", "School", "Workplace", "Community"] + [f"Facility{f}" for f in range(4, F)]
    s["facilities"] = [{"name": fnames[f]} for f in range(F)]
    ts = rng.dirichlet(np.ones(F), G).T  # [F, G], columns sum to 1
    s["facilities_timespent"] = [{"locales": "UnitedProvinces.*", "matrix": ts.tolist()}]
    fi = rng.dirichlet(np.ones(G), (F, G)) * 0.98  # [F, G, G], rows sum to 0.98
    s["facilities_interactions"] = [{"locales": "UnitedProvinces.*", "facilities": fi.tolist()}]
    data = []
    for i in range(L):
        near = [(i + d) % L for d in (-2, -1, 1, 2)]
        far = []
        while len(far) < 4 and L > 9:
            c = int(rng.integers(0, L))
            if c != i and c not in near and c not in far:
                far.append(c)
        dst = list(dict.fromkeys([d for d in near + far if d != i]))
        w = rng.dirichlet(np.ones(len(dst)))
        data += [{"src_locale_id": f"UP.L{i:04d}", "dst_locale_id": f"UP.L{d:04d}", "group": "*", "value": float(v)}
                 for d, v in zip(dst, w)]
    s["border"] = {"data": data, "specifications": [{"impedance": 70}]}
    s["airport"] = {"data": [], "specifications": []}
    s["groups_locales_parameters"] = []
    s["initial_info"] = {"initializers": [
        {"locale_regex": f"UnitedProvinces.L{i:04d}", "group": "*", "compartment": "Ipre", "value": "100"}
        for i in range(0, L, max(1, L // 8))]}
    return s


def workload_desc(name: str):
    """EpiDesc of a synthetic workload: the fixture exported from the reference's Static (tools/gen_syn.py)."""
    from . import scenarios
    return scenarios.load(name)
