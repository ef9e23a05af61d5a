"""BASELINE.json configs at their STATED size, in the driver-run GPU suite (missing #1):
C2 n=30 plain, C4 30 photons / 60 modes repeated, C5 20 photons / 40 modes / 1000 shots sampling --
each against the unmodified reference's long double calculators run live on the box's host cores
(oracle/_ref), with the timings written to gpurun_out/configs_gpu_tests.json."""
import json
import os
import time

import numpy as np
import pytest

from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOG = {}


@pytest.fixture(scope="module")
def pqb():
    import piquassoboost_b200 as p
    p.load()
    return p


def _timed(fn, reps=3):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        v = fn()
    return v, (time.perf_counter() - t) / reps


def _save():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(_LOG, open(os.path.join(ROOT, "gpurun_out", "configs_gpu_tests.json"), "w"), indent=1)


def test_c2_n30_unitary_and_subblock(pqb, ref):
    """C2: n=30 complex permanent, single B200, vs the reference's extended precision
    (boostbenchmarks/permanent_benchmark.py:55-62 pattern: unitary_group.rvs)."""
    for name, A in (("haar30", haar_unitary(30, 30)), ("subblock30_of_60", haar_unitary(60, 4)[:30, :30])):
        A = np.ascontiguousarray(A)
        got, sec = _timed(lambda: pqb.BBFGPermanentLongDouble(A).calculate())
        st = pqb.last_stats()
        t = time.perf_counter()
        want = ref.bbfg(A, True)
        ref_s = time.perf_counter() - t
        err = relerr(got, want)
        _LOG["C2_" + name] = {"call_s": sec, "kernel_ms": st["kernel_ms"], "ndev": st["ndev"], "rel_err": err,
                              "reference_long_double_s": ref_s, "reference_cores": ref.hardware_concurrency()}
        assert err < 1e-11, (name, got, want)
    _save()


@pytest.mark.parametrize("name", ["15x2", "10x3", "multinomial", "30_distinct", "repeated_both_sides"])
def test_c4_thirty_photons_in_sixty_modes(pqb, ref, name):
    """C4: repeated-row permanents for 30 photons in 60 modes with multiplicities."""
    U = np.ascontiguousarray(haar_unitary(60, 4))
    rng = np.random.default_rng(4)
    inp = np.array([1] * 30 + [0] * 30)
    if name == "15x2":
        out = np.array([2] * 15 + [0] * 45)
    elif name == "10x3":
        out = np.array([3] * 10 + [0] * 50)
    elif name == "multinomial":
        out = rng.multinomial(30, rng.dirichlet(np.ones(60) * 0.3))
    elif name == "30_distinct":
        out = np.array(([0, 1] * 30))
    else:
        inp = rng.multinomial(30, rng.dirichlet(np.ones(60) * 0.2))
        out = rng.multinomial(30, rng.dirichlet(np.ones(60) * 0.2))
    got, sec = _timed(lambda: pqb.BBFGRepeatedPermanentCalculatorLongDouble(U, inp, out).calculate())
    st = pqb.last_stats()
    from piquassoboost_b200 import _lib
    bound, used_dd = _lib.last_repeated_bound()
    t = time.perf_counter()
    want = ref.bbfg_repeated(U, inp, out, True)
    ref_s = time.perf_counter() - t
    err = relerr(got, want)
    _LOG["C4_" + name] = {"call_s": sec, "kernel_ms": st["kernel_ms"], "gray_steps": st["gray_steps"], "rel_err": err,
                          "bound": bound, "double_double": used_dd, "reference_long_double_s": ref_s,
                          "out": np.asarray(out).tolist(), "inp": np.asarray(inp).tolist()}
    _save()
    assert err < 1e-10, (name, got, want, bound)


def test_c5_twenty_photons_forty_modes(pqb, ref):
    """C5: Clifford-Clifford B sampling, 20 photons / 40 modes.  The first 50 shots must be IDENTICAL to the
    reference's (same std::mt19937 stream, boson_sampling_benchmark.py:84-94 pattern); the 1000-shot run
    is checked for its invariants and timed."""
    U = np.ascontiguousarray(haar_unitary(40, 5))
    state = np.array([1] * 20 + [0] * 20, dtype=np.int64)
    sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=123)
    got = np.array(sim.simulate(state, 50))
    t = time.perf_counter()
    want = ref.simulate_B(U, 123, state, 50)
    ref_s = time.perf_counter() - t
    assert got.shape == (50, 40)
    assert np.array_equal(got, want)
    sim.seed(7)
    sim.simulate(state, 1000)                      # warm (buffers, module load)
    sim.seed(7)
    t = time.perf_counter()
    shots = sim.simulate(state, 1000)
    call_s = time.perf_counter() - t
    st = pqb.last_stats()
    arr = np.array(shots)
    assert isinstance(shots, tuple) and len(shots) == 1000 and shots[0].dtype == np.int64
    assert arr.shape == (1000, 40) and np.all(arr >= 0) and np.all(arr.sum(axis=1) == 20)
    assert np.array_equal(arr[:50], ref.simulate_B(U, 7, state, 50))
    # the same seed gives the same shots whatever the wave / cluster decomposition
    sim.seed(7)
    again = np.array(sim.simulate(state, 300))
    assert np.array_equal(again, arr[:300])
    _LOG["C5_20ph_40modes"] = {"shots": 1000, "call_s": call_s, "kernel_ms": st["kernel_ms"], "ndev": st["ndev"],
                               "launches": st["launches"], "gray_terms": st["gray_steps"],
                               "reference_s_per_shot": ref_s / 50, "reference_cores": ref.hardware_concurrency()}
    _save()
