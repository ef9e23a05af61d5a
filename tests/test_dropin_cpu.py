"""CPU check of the drop-in build (oracle/_ref/dropin, made by `make -C oracle dropin` where /root/reference
exists): the reference's UNMODIFIED CPython wrapper code imports, constructs its objects over
include/dropin/*.h, and -- there being no GPU here -- reports the engine's "no CUDA device" failure as a Python
exception through the reference's own `catch (std::string)` (no silent CPU fallback behind the reference's API)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")


def _built():
    d = os.path.join(DROPIN, "piquassoboost", "sampling")
    return os.path.isdir(d) and any(f.startswith("Boson_Sampling_Utilities_wrapper") and f.endswith(".so")
                                    for f in os.listdir(d))


def _cuda_present():
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return False
    n = ctypes.c_int(0)
    return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0


pytestmark = pytest.mark.skipif(not _built(), reason="oracle/_ref/dropin not built (needs /root/reference)")


def test_reference_wrapper_types_import_and_fail_loudly_without_a_gpu():
    if _cuda_present():
        pytest.skip("a GPU is present: tests/test_reference_wrappers_gpu.py covers this stack")
    sys.path.insert(0, DROPIN)
    try:
        from piquassoboost.sampling.Boson_Sampling_Utilities import BBFGRepeatedPermanentCalculatorDouble, GlynnPermanent
        import piquassoboost.sampling.Boson_Sampling_Utilities_wrapper as w
        assert os.path.realpath(w.__file__).startswith(os.path.realpath(DROPIN))
        with pytest.raises(Exception, match="no usable CUDA device"):
            GlynnPermanent(np.eye(3, dtype=complex)).calculate()
        with pytest.raises(Exception, match="no usable CUDA device"):
            BBFGRepeatedPermanentCalculatorDouble(np.eye(3, dtype=complex), np.array([1, 1, 0]), np.array([0, 1, 1])).calculate()
        # the trivial branch the reference answers before touching a calculator still works
        from piquassoboost.sampling.simulation_strategies.GeneralizedCliffordsBSimulationStrategy_wrapper import (
            GeneralizedCliffordsBSimulationStrategy_wrapper)
        sim = GeneralizedCliffordsBSimulationStrategy_wrapper(np.eye(3, dtype=complex), 5, 0)
        assert sim.simulate(np.array([1, 1, 0], dtype=np.int64), 0) == ()
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "piquassoboost" or k.startswith("piquassoboost.")]:
            del sys.modules[k]
