"""Synthetic weights / clips for the oracle and the tests: re-export of cutie_b200/utils/synth.py (data generation only;
it lives outside oracle/ so that the measured GPU arm of bench.py imports nothing from this package)."""
from cutie_b200.utils.synth import *  # noqa: F401,F403
from cutie_b200.utils.synth import synthetic_state_dict, synthetic_video  # noqa: F401
