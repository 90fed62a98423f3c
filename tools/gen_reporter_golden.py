"""Golden vectors for the reporter quantities from the UNMODIFIED reference, run in this
container: the scenario's own schedule for T days through Epidemic.step, then the reference Reporter's
compute_history / compute_incidences / compute_comp_news (core/reporter.py:34-51) and its R-number chain
(average infectious duration, location / window statistics, instantaneous, basic and case reproduction numbers,
core/reporter.py:54-204, default configuration).

  PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=tools/gym_shim:/root/reference:. python tools/gen_reporter_golden.py [names...]

Writes tests/golden/reporter_<name>.npz: incidences [T+1, L, G], comp_news [T+1, C, L, G], the scenario's hashed
incidence edges (core_utils.py:243-254) and infectious-compartment mask (compartment tag INF), and stores the
latter two in the desc's meta ("reporter") so that epipolicy_rl_b200/reporter.py needs no reference at run time."""
import json
import os
import sys

import numpy as np

REF = os.environ.get("EPI_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
SCN = os.path.join(os.path.dirname(__file__), "..", "epipolicy_rl_b200", "scenarios")  # product data: the descriptors

from epipolicy.core.epidemic import construct_epidemic  # noqa: E402
from epipolicy_rl_b200.desc import EpiDesc  # noqa: E402
from epipolicy_rl_b200.export import reporter_meta  # noqa: E402


def main(names, T=28):
    for name in names:
        session = json.load(open(os.path.join(REF, "jsons", name + ".json")))
        epi = construct_epidemic(session)
        epi.reset()
        for t in range(T):
            epi.step(epi.static.schedule.action_list[t])
        rep = epi.reporter
        rep.compute_history()
        rep.compute_incidences()
        rep.compute_comp_news()
        # the R-number chain with the reference's default configuration (utility/singleton.py:30-34)
        rep.compute_average_infectious_duration()
        rep.compute_location_statistics()
        rep.compute_window_statistics()
        rep.compute_instantaneous_reproductive_number()
        rep.compute_basic_reproductive_number()
        rep.compute_case_reproductive_number()
        meta = reporter_meta(epi.static, session)
        np.savez_compressed(os.path.join(OUT, f"reporter_{name}.npz"), incidences=np.asarray(rep.incidences),
                            comp_news=np.asarray(rep.comp_news),
                            incidence_edges=np.array(meta["incidence_edges"], np.int32).reshape(-1, 2),
                            inf_comps=np.array(meta["inf_comps"]), T=T,
                            average_infectious_duration=np.asarray(rep.average_infectious_duration),
                            window_infectious_duration=np.asarray(rep.window_infectious_duration),
                            location_window_incidence=np.asarray(rep.location_window_incidence),
                            location_window_infectious_count=np.asarray(rep.location_window_infectious_count),
                            location_instant_R=np.asarray(rep.location_instant_R),
                            location_basic_R=np.asarray(rep.location_basic_R),
                            location_case_R=np.asarray(rep.location_case_R),
                            location_avg_model_parameters=np.asarray(rep.location_avg_model_parameters))
        dpath = os.path.join(SCN, f"desc_{name}.npz")
        d = EpiDesc.load(dpath)
        d.meta["reporter"] = meta
        d.save(dpath)
        print(name, "incidence edges", meta["incidence_edges"], "inf", meta["inf_comps"], flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["SIR_A", "SIRV_B", "COVID_C"])
