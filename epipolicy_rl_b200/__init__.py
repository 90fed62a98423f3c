"""epipolicy_rl_b200 — B200-native batched epidemic-simulation step behind the Epipolicy_RL API.

Only the hot path named by BASELINE.json's north_star lives here:
`EpiEnv.step` -> `Epidemic.step` -> `Epidemic.get_next_state` of the reference, rebuilt as
hand-written sm_100a CUDA behind a C-ABI (include/epi_b200.h). The reference's JSON parser,
`make_static` and intervention objects are used as they are (see export.py).
"""
__version__ = "0.1.0"
