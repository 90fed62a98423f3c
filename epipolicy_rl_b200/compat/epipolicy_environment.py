"""Same-named stand-in for the reference's `epipolicy_environment` module (epipolicy_environment.py:1-71), so that
`runner.py`, `baselines.py` and `plots.py` — which do `from epipolicy_environment import EpiEnv` — pick up the
B200-backed environment without an edit:

    PYTHONPATH=/path/to/epi-b200/epipolicy_rl_b200/compat:/path/to/epi-b200:/path/to/Epipolicy_RL python runner.py ...

`EpiEnv(session)` has the reference's constructor shape, attributes (`epi`, `act_domain`, `action_space`,
`observation_space`) and reset/step/render/close surface; the session is parsed by the reference's own
parser / make_static (kept as is), exported to an EpiDesc and stepped by libepi_b200.so."""
from epipolicy_rl_b200.engine import PERIOD  # noqa: F401  (module-level constant of the reference, line 1)
from epipolicy_rl_b200.env import BatchedEpiEnv, EpiEnv  # noqa: F401
