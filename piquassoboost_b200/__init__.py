"""piquassoboost_b200 -- B200 (sm_100a) CUDA engine for piquassoboost's Glynn/BBFG permanent path.

Only the hot path of the reference is here: the permanent calculators, their
batched forms and the sub-permanent evaluation of the Clifford-Clifford B sampler, behind the
reference's own class names.  The compute lives in libpqb200.so (csrc/, C ABI in include/pqb200.h).
"""
from . import _lib
from ._lib import Pqb200Error, last_stats, load
from .Boson_Sampling_Utilities import *  # noqa: F401,F403
from .simulation_strategies import (  # noqa: F401
    GeneralizedCliffordsBLossySimulationStrategy,
    GeneralizedCliffordsBLossySimulationStrategy_wrapper,
    GeneralizedCliffordsBSimulationStrategy,
    GeneralizedCliffordsBSimulationStrategy_wrapper,
    GeneralizedCliffordsBUniformLossesSimulationStrategy,
    GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper,
    compute_pmf,
)

__all__ = ["_lib", "Pqb200Error", "last_stats", "load"]
