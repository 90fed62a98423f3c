"""Reporter quantities of the reference from the engine's flat observables.

The reference's Reporter (core/reporter.py) derives incidence and per-compartment "new" counts from the cumulative
components of every day's Observable (obj/observable.py:6-25). The batched engine carries exactly those components
for its RK45 error norm and returns them in the reference's flat layout (`BatchedEpidemic.get_state("flat")`,
[N, n_flat]), so these quantities are a few array expressions on top of the states of a run — for all N
environments at once, no per-env Python loop. Host-side numpy: the reporter is not on the hot path.

`desc.meta["reporter"]` (written by tools/gen_reporter_golden.py / export.py from the reference Static) holds the
scenario's incidence edges (core_utils.py:243-254: edges out of a SUS-tagged compartment that carry a transmission
rate) and the INF-tagged compartments."""
from __future__ import annotations

import numpy as np


def split_flat(desc, flat: np.ndarray) -> dict:
    """Observable.flatten order (observable.py:24-25): current_comp, cumulative_gain, cumulative_lost
    [C,L,G] each, cumulative_move [C,C,L,G], cumulative_cost [I,L]; leading axes of `flat` are kept."""
    C, L, G, I = desc.C, desc.L, desc.G, desc.I
    lead, n = flat.shape[:-1], C * L * G
    if flat.shape[-1] != 3 * n + C * n + I * L:
        raise ValueError(f"flat has {flat.shape[-1]} components, the scenario has {3 * n + C * n + I * L}")
    return {
        "current_comp": flat[..., :n].reshape(*lead, C, L, G),
        "cumulative_gain": flat[..., n:2 * n].reshape(*lead, C, L, G),
        "cumulative_lost": flat[..., 2 * n:3 * n].reshape(*lead, C, L, G),
        "cumulative_move": flat[..., 3 * n:3 * n + C * n].reshape(*lead, C, C, L, G),
        "cumulative_cost": flat[..., 3 * n + C * n:].reshape(*lead, I, L),
    }


def _meta(desc):
    m = desc.meta.get("reporter")
    if m is None:
        raise ValueError("desc.meta['reporter'] is missing: export the scenario with tools/gen_reporter_golden.py / export.py")
    return m


def total_incidence(desc, flat: np.ndarray) -> np.ndarray:
    """get_total_incidence_matrix (obj/static.py:137-143): sum over the incidence edges (c1, c2) of
    cumulative_move[c1, c2] - cumulative_move[c2, c1] -> [..., L, G]."""
    mv = split_flat(desc, flat)["cumulative_move"]
    res = np.zeros(mv.shape[:-4] + mv.shape[-2:])
    for c1, c2 in _meta(desc)["incidence_edges"]:
        res += mv[..., c1, c2, :, :]
        res -= mv[..., c2, c1, :, :]
    return res


def total_new(desc, flat: np.ndarray) -> np.ndarray:
    """get_total_new_matrix (obj/static.py:145-152): cumulative_gain[c1] + sum_{c2 != c1} max(move[c2, c1] -
    move[c1, c2], 0) -> [..., C, L, G] (c2 ascending, like the reference's loop)."""
    o = split_flat(desc, flat)
    res, mv = o["cumulative_gain"].copy(), o["cumulative_move"]
    for c1 in range(desc.C):
        for c2 in range(desc.C):
            if c1 != c2:
                res[..., c1, :, :] += np.maximum(mv[..., c2, c1, :, :] - mv[..., c1, c2, :, :], 0)
    return res


def total_delta(desc, flat: np.ndarray) -> np.ndarray:
    """get_total_delta_matrix (obj/static.py:154-162): gain - lost + net moves into each compartment."""
    o = split_flat(desc, flat)
    mv = o["cumulative_move"]
    return o["cumulative_gain"] - o["cumulative_lost"] - mv.sum(-3) + mv.sum(-4)


def _consecutive_difference(a: np.ndarray) -> np.ndarray:
    """utility/utils.py:49-55 along axis 0: the first entry is kept, then day-to-day differences."""
    out = a.copy()
    out[1:] = a[1:] - a[:-1]
    return out


def incidences(desc, flats: np.ndarray) -> np.ndarray:
    """Reporter.compute_incidences (core/reporter.py:45-48). flats: [T+1, ..., n_flat], the initial state first.
    Returns [T+1, ..., L, G]: entry 0 is the initially infected (INF-tagged compartments of day 0), entry t the
    incidence of day t."""
    inf = np.asarray(_meta(desc)["inf_comps"], np.float64)
    x0 = split_flat(desc, flats[0])["current_comp"]
    initial = np.tensordot(inf, np.moveaxis(x0, -3, 0), axes=(0, 0))  # get_matrix_with_tag(.., INF_TAG)
    return _consecutive_difference(total_incidence(desc, flats) + initial)


def comp_news(desc, flats: np.ndarray) -> np.ndarray:
    """Reporter.compute_comp_news (core/reporter.py:50-51): [T+1, ..., C, L, G]."""
    return _consecutive_difference(total_new(desc, flats))
