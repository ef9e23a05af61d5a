"""The reference's OWN CPython wrapper code over the B200 engine (missing #3).

`make -C oracle dropin` compiles sampling/GlynnPermanentCalculator_Wrapper.hpp and
sampling/ChinHuhPermanentCalculator_Wrapper.hpp UNMODIFIED from /root/reference with include/dropin/ in front
of the reference's calculator headers, and stages them with the reference's own Boson_Sampling_Utilities.py and
unittests/test_permanent.py under oracle/_ref/dropin/ (built in the dev container, shipped to the GPU box as
built files).  Here: (1) the reference's unit tests run unmodified through that stack; (2) the reference
wrapper's list branches are the yardstick for this package's Python mirror."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")


def _have_dropin():
    d = os.path.join(DROPIN, "piquassoboost", "sampling")
    return os.path.isdir(d) and any(f.startswith("Boson_Sampling_Utilities_wrapper") and f.endswith(".so")
                                    for f in os.listdir(d))


needs_dropin = pytest.mark.skipif(not _have_dropin(), reason="oracle/_ref/dropin not built (make -C oracle dropin)")


@needs_dropin
def test_reference_unittests_run_unmodified():
    """unittests/test_permanent.py of the reference, byte for byte, importing
    piquassoboost.sampling.Boson_Sampling_Utilities (the reference's file) over the reference's wrapper
    types over libpqb200.so."""
    env = dict(os.environ, PYTHONPATH=DROPIN + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", DROPIN,
                          os.path.join(DROPIN, "unittests", "test_permanent.py")],
                         capture_output=True, text=True, env=env, cwd=DROPIN, timeout=600)
    tail = res.stdout[-1500:] + res.stderr[-500:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout and "failed" not in res.stdout, tail


@pytest.fixture(scope="module")
def refwrap():
    if not _have_dropin():
        pytest.skip("oracle/_ref/dropin not built")
    code_path = DROPIN
    if code_path not in sys.path:
        sys.path.insert(0, code_path)
    import importlib
    mod = importlib.import_module("piquassoboost.sampling.Boson_Sampling_Utilities")
    return mod


@needs_dropin
def test_reference_classes_on_the_engine(refwrap, ref):
    """Every class of the reference's Boson_Sampling_Utilities.py that maps to a supported lib id."""
    A = np.ascontiguousarray(haar_unitary(12, 3))
    want = ref.bbfg(A, True)
    for name in ("GlynnPermanent", "GlynnPermanentDouble", "BBFGPermanentDouble", "BBFGPermanentLongDouble",
                 "GlynnPermanentSimpleDouble", "GlynnPermanentSimpleLongDouble"):
        assert relerr(getattr(refwrap, name)(A).calculate(), want) < 1e-12, name
    vals = refwrap.BBFGPermanentDouble([A, A[:6, :6].copy()]).calculate()      # list-of-matrices branch
    assert relerr(vals[0], want) < 1e-12 and relerr(vals[1], ref.bbfg(A[:6, :6].copy(), True)) < 1e-12
    U = np.ascontiguousarray(haar_unitary(6, 8))
    a, b = np.array([2, 1, 0, 1, 0, 1], dtype=np.int64), np.array([0, 0, 3, 1, 1, 0], dtype=np.int64)
    want = ref.bbfg_repeated(U, a, b, True)
    for name in ("ChinHuhPermanentCalculator", "GlynnRepeatedPermanentCalculator", "GlynnRepeatedPermanentCalculatorDouble",
                 "BBFGRepeatedPermanentCalculatorDouble", "BBFGRepeatedPermanentCalculatorLongDouble"):
        assert relerr(getattr(refwrap, name)(U, a, b).calculate(), want) < 1e-12, name
    with pytest.raises(Exception, match="Number of input and output states should be equal"):
        refwrap.BBFGRepeatedPermanentCalculatorDouble(U, a, np.array([1, 0, 0, 0, 0, 0], dtype=np.int64)).calculate()


@needs_dropin
def test_python_mirror_matches_the_reference_wrapper_list_branches(refwrap):
    """The reference wrapper's nested-list semantics (ChinHuhPermanentCalculator_Wrapper.hpp:233-326), executed
    by the reference's own C++ wrapper code, against this package's Python mirror of it.

    The reference stores the converted inner item back with `PyList_SetItem(oOut, i, o)` (index i of the OUTER
    loop, :258): for the second and later outer entries that overwrites inner items before they are read and
    raises IndexError on ragged input, so the comparison uses one outer entry, where the reference's branch
    computes what it documents (ret[0][j] = perm(input[0], output[0][j])).  The inner lists are rebuilt per
    call because that statement also rewrites the caller's list."""
    import piquassoboost_b200 as pqb
    U = np.ascontiguousarray(haar_unitary(5, 22))
    one = np.array([1, 1, 1, 0, 0], dtype=np.int64)

    def many():
        return [np.array([0, 1, 1, 1, 0], dtype=np.int64), np.array([3, 0, 0, 0, 0], dtype=np.int64),
                np.array([1, 0, 1, 0, 1], dtype=np.int64)]

    theirs = refwrap.ChinHuhPermanentCalculator_wrapper(7, U, [one], [many()]).calculate()
    ours = pqb.ChinHuhPermanentCalculator_wrapper(7, U, [one], [many()]).calculate()
    assert [len(r) for r in theirs] == [len(r) for r in ours] == [3]
    for x, y in zip(theirs[0], ours[0]):
        assert relerr(y, x) < 1e-13
    theirs = refwrap.ChinHuhPermanentCalculator_wrapper(7, U, [many()], [one]).calculate()    # multiInput
    ours = pqb.ChinHuhPermanentCalculator_wrapper(7, U, [many()], [one]).calculate()
    assert [len(r) for r in theirs] == [len(r) for r in ours] == [3]
    for x, y in zip(theirs[0], ours[0]):
        assert relerr(y, x) < 1e-13
    # both Python classes refuse the nested form the same way (Boson_Sampling_Utilities.py:40-55)
    for bad in ([[one], [many()]], [[many()], [one]]):
        with pytest.raises(Exception, match="should be a numpy array, or a list of numpy arrays"):
            refwrap.BBFGRepeatedPermanentCalculatorDouble(U, *bad)
        with pytest.raises(Exception, match="should be a numpy array, or a list of numpy arrays"):
            pqb.BBFGRepeatedPermanentCalculatorDouble(U, *bad)


@needs_dropin
def test_reference_strategy_wrappers_on_the_engine(ref):
    """SURVEY call stack 3.3: the reference's three B-strategy wrapper MODULES
    (sampling/simulation_strategies/GeneralizedCliffordsB*SimulationStrategy_wrapper.cpp, compiled unmodified)
    over include/dropin/CGeneralizedCliffordsB*SimulationStrategy.h -> libpqb200.so: seeded shots must be
    identical to the unmodified reference's own C++ strategies (oracle/_ref), as tuples of int64 arrays."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import importlib
    pkg = "piquassoboost.sampling.simulation_strategies."
    B = importlib.import_module(pkg + "GeneralizedCliffordsBSimulationStrategy_wrapper").GeneralizedCliffordsBSimulationStrategy_wrapper
    UL = importlib.import_module(pkg + "GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper").GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper
    LO = importlib.import_module(pkg + "GeneralizedCliffordsBLossySimulationStrategy_wrapper").GeneralizedCliffordsBLossySimulationStrategy_wrapper
    U = np.ascontiguousarray(haar_unitary(10, 77))
    state = np.array([1, 1, 1, 0, 1, 0, 2, 0, 0, 0], dtype=np.int64)
    sim = B(interferometer_matrix=U, seed=123, lib=0)
    got = sim.simulate(state, 40)
    assert isinstance(got, tuple) and len(got) == 40 and got[0].dtype == np.int64
    assert np.array_equal(np.array(got), ref.simulate_B(U, 123, state, 40))
    sim.seed(9)                                                     # .seed(K) then simulate again
    assert np.array_equal(np.array(sim.simulate(state, 25)), ref.simulate_B(U, 9, state, 25))
    ul = UL(interferometer_matrix=U, transmissivity=0.8, seed=5, lib=0)
    assert np.array_equal(np.array(ul.simulate(state, 60)), ref.simulate_B_uniform_losses(U, 5, state, 60, 0.8))
    M = np.ascontiguousarray(U * np.linspace(0.7, 1.0, 10)[None, :])  # a lossy (sub-unitary) interferometer
    lo = LO(interferometer_matrix=M.copy(), approximated_modes=3, seed=11, lib=0)
    assert np.array_equal(np.array(lo.simulate(state, 50)), ref.simulate_B_lossy(M.copy(), 11, state, 50, 3))


@needs_dropin
def test_reference_free_function_module_on_the_engine(ref):
    """sampling/permanent_calculators_wrapper.cpp (permanent_CPU_repeated_double / _long_double), a complete
    module of the reference compiled unmodified over include/dropin."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import importlib
    mod = importlib.import_module("piquassoboost.sampling.permanent_calculators")
    U = np.ascontiguousarray(haar_unitary(6, 8))
    a, b = np.array([2, 1, 0, 1, 0, 1], dtype=np.int64), np.array([0, 0, 3, 1, 1, 0], dtype=np.int64)
    want = ref.bbfg_repeated(U, a, b, True)
    assert relerr(mod.permanent_CPU_repeated_double(U, a, b), want) < 1e-12
    assert relerr(mod.permanent_CPU_repeated_long_double(U, a, b), want) < 1e-12
