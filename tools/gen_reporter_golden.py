"""Golden vectors for the reporter quantities from the UNMODIFIED reference, run in this
container: the scenario's own schedule for T days through Epidemic.step, then the reference Reporter's
compute_history / compute_incidences / compute_comp_news (core/reporter.py:34-51).

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
        meta = reporter_meta(epi.static)
        np.savez_compressed(os.path.join(OUT, f"reporter_{name}.npz"), incidences=np.asarray(rep.incidences),
                            comp_news=np.asarray(rep.comp_news),
                            incidence_edges=np.array(meta["incidence_edges"], np.int32).reshape(-1, 2),
                            inf_comps=np.array(meta["inf_comps"]), T=T)
        dpath = os.path.join(SCN, f"desc_{name}.npz")
        d = EpiDesc.load(dpath)
        d.meta["reporter"] = meta
        d.save(dpath)
        print(name, "incidence edges", meta["incidence_edges"], "inf", meta["inf_comps"], flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["SIR_A", "SIRV_B", "COVID_C"])
