"""GPU parity tests of the repeated-row permanent (pqb200_permanent_repeated and the Python mirror
of ChinHuhPermanentCalculator_wrapper / RepeatedPermanentCalculator) against the reference's own
KATs, the committed golden vectors, and the live long-double reference calculator."""
import numpy as np
import pytest

import golden_util as G
from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-11


@pytest.fixture(scope="module")
def pqb():
    import piquassoboost_b200 as p
    p.load()
    return p


def test_reference_unittest_kats(pqb):
    """unittests/test_permanent.py:47-251 -- same matrices, same call, same np.isclose assertion."""
    for k in G.load("kat_test_permanent.json"):
        m = G.mat(k["matrix"])
        res = pqb.BBFGRepeatedPermanentCalculatorDouble(
            m, np.asarray(k["cols"], dtype=np.int64), np.asarray(k["rows"], dtype=np.int64)).calculate()
        assert np.isclose(res, G.z(k["expected"])), k["name"]
        assert relerr(res, G.z(k["ref_long_double"])) < TOL, k["name"]


def test_golden_repeated_cases(pqb):
    for c in G.load("repeated_reference.json"):
        m = G.mat(c["matrix"])
        got = pqb.BBFGRepeatedPermanentCalculatorLongDouble(m, c["cols"], c["rows"]).calculate()
        want = G.z(c["bbfg_rep_long_double"])
        if want == 0:
            assert abs(got) < 1e-14
        else:
            assert relerr(got, want) < TOL, (c["rows"], c["cols"], got, want)


@pytest.mark.parametrize("photons,modes,seed", [(4, 4, 0), (6, 5, 1), (8, 6, 2), (10, 8, 3), (12, 7, 4),
                                                (14, 9, 5), (16, 6, 6), (18, 12, 7), (20, 10, 8)])
def test_random_occupations_vs_reference_long_double(pqb, ref, photons, modes, seed):
    rng = np.random.default_rng(seed)
    U = haar_unitary(modes, 300 + seed)
    for trial in range(4):
        rows = rng.multinomial(photons, rng.dirichlet(np.ones(modes)))
        cols = rng.multinomial(photons, rng.dirichlet(np.ones(modes)))
        got = pqb.GlynnRepeatedPermanentCalculator(U, cols, rows).calculate()
        want = ref.bbfg_repeated(U, cols, rows, True)
        assert relerr(got, want) < TOL, (rows.tolist(), cols.tolist(), got, want)


def test_expanded_equals_compact(pqb):
    """ctests/permanents_accuracy.cpp:41-79: Glynn on the physically expanded matrix == repeated
    calculator on the compact matrix with multiplicities."""
    rng = np.random.default_rng(9)
    A = rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5))
    rows = np.array([2, 1, 3, 0, 2])
    cols = np.array([1, 3, 1, 2, 1])
    expanded = np.repeat(np.repeat(A, rows, axis=0), cols, axis=1)
    assert relerr(pqb.GlynnPermanent(expanded).calculate(),
                  pqb.BBFGRepeatedPermanentCalculatorDouble(A, cols, rows).calculate()) < 1e-12


def test_large_space_and_all_single_route(pqb, ref):
    # 24 photons: rows 8 x 3 -> 3 * 4^7 = 49152 steps; columns all single -> plain route when transposed
    U = haar_unitary(24, 77)
    rows = np.array([3] * 8 + [0] * 16)
    cols = np.array([1] * 24)
    got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, cols, rows).calculate()
    assert relerr(got, ref.bbfg_repeated(U, cols, rows, True)) < TOL
    ones = np.ones(14, dtype=np.int64)
    V = haar_unitary(14, 5)
    assert relerr(pqb.BBFGRepeatedPermanentCalculatorDouble(V, ones, ones).calculate(), ref.bbfg(V, True)) < TOL


def test_photon_number_mismatch_raises(pqb):
    """performance_tests/test_permanent_calculators.py:43-74."""
    with pytest.raises(Exception) as e:
        pqb.BBFGRepeatedPermanentCalculatorDouble(np.eye(3), [1, 1, 0], [1, 1, 1]).calculate()
    assert "Number of input and output states should be equal" in str(e.value)


def test_zero_and_trivial_states(pqb):
    U = haar_unitary(4, 1)
    assert pqb.BBFGRepeatedPermanentCalculatorDouble(U, [0, 0, 0, 0], [0, 0, 0, 0]).calculate() == 1
    one = pqb.BBFGRepeatedPermanentCalculatorDouble(U, [0, 1, 0, 0], [0, 0, 1, 0]).calculate()
    assert relerr(one, U[2, 1]) < 1e-15
    pw = pqb.BBFGRepeatedPermanentCalculatorDouble(U, [0, 3, 0, 0], [0, 0, 3, 0]).calculate()
    assert relerr(pw, 6 * U[2, 1] ** 3) < 1e-13     # 3! * a^3


def test_list_of_states_batch(pqb, ref):
    U = haar_unitary(5, 21)
    ins = [np.array([1, 1, 1, 0, 0]), np.array([2, 0, 0, 1, 0])]
    outs = [np.array([0, 1, 1, 1, 0]), np.array([3, 0, 0, 0, 0]), np.array([1, 0, 1, 0, 1])]
    res = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs).calculate()
    assert len(res) == 2 and all(len(r) == 3 for r in res)
    for i, a in enumerate(ins):
        for j, b in enumerate(outs):
            assert relerr(res[i][j], ref.bbfg_repeated(U, a, b, True)) < TOL


def test_free_functions(pqb, ref):
    U = haar_unitary(6, 8)
    a, b = [2, 1, 0, 1, 0, 1], [0, 0, 3, 1, 1, 0]
    want = ref.bbfg_repeated(U, a, b, True)
    assert relerr(pqb.permanent_CPU_repeated_double(U, a, b), want) < TOL
    assert relerr(pqb.permanent_CPU_repeated_long_double(U, a, b), want) < TOL


def test_connector_permanent_hook(pqb, ref):
    """piquassoboost/connector.py:23-31: the callable installed as BoostConnector.permanent."""
    from piquassoboost_b200 import connector
    U = haar_unitary(6, 31)
    a, b = np.array([1, 0, 2, 0, 1, 0]), np.array([0, 1, 1, 1, 0, 1])
    assert relerr(connector.cpp_permanent_function(U, a, b), ref.bbfg_repeated(U, a, b, True)) < TOL
    # list-valued states give a list, which the reference hook maps to 1.0 (connector.py:27-29)
    assert connector.cpp_permanent_function(U, [a], [b]) == 1.0
    res = connector.permanent_function_batch(U, [a, [2, 2, 0, 0, 0, 0]], [b])
    assert relerr(res[0][0], ref.bbfg_repeated(U, a, b, True)) < TOL
    assert relerr(res[1][0], ref.bbfg_repeated(U, [2, 2, 0, 0, 0, 0], b, True)) < TOL


@pytest.mark.parametrize("photons", list(range(3, 45)))
def test_every_repeated_kernel_instantiation(pqb, ref, photons):
    """glynn_repeated_kernel is instantiated per total photon number N (= physically repeated
    columns): run each one on 6 modes with high multiplicities (small mixed-radix Gray space, so that
    the reference's long double answer is cheap).  Many photons in few modes is the ill-conditioned
    corner of the formula (huge binomial-weighted terms cancelling): there the bound is north_star's
    1e-10 and "no worse than 10x the reference's own double path", below 25 photons the usual 1e-11."""
    rng = np.random.default_rng(500 + photons)
    U = haar_unitary(6, 77)
    for trial in range(2):
        inp = rng.multinomial(photons, [1 / 6] * 6).astype(np.int64)
        out = rng.multinomial(photons, [0.3, 0.2, 0.2, 0.1, 0.1, 0.1] if trial else [1 / 6] * 6).astype(np.int64)
        want = ref.bbfg_repeated(U, inp, out, True)
        got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
        err = relerr(got, want)
        if photons <= 24:
            assert err < TOL, (photons, inp, out, got, want)
        else:
            ref_double_err = relerr(ref.bbfg_repeated(U, inp, out, False), want)
            assert err < 1e-10 or err < 10 * ref_double_err, (photons, inp, out, got, want, ref_double_err)


def test_double_double_repeated_mode(pqb, ref):
    """pqb200_permanent_repeated_dd (a straight double-double walk of the reference's n-ary Gray code)
    against the unmodified reference's long double repeated calculator; also serves lib 9 (GlynnRepInf)."""
    from piquassoboost_b200.Boson_Sampling_Utilities import permanent_repeated_double_double
    rng = np.random.default_rng(99)
    for m, photons in [(4, 5), (5, 9), (6, 12), (8, 16), (7, 7), (5, 22)]:
        U = haar_unitary(m, 60 + m)
        inp = rng.multinomial(photons, [1 / m] * m).astype(np.int64)
        out = rng.multinomial(photons, [1 / m] * m).astype(np.int64)
        want = ref.bbfg_repeated(U, inp, out, True)
        got, parts = permanent_repeated_double_double(U, inp, out)
        assert relerr(got, want) < 2e-15, (m, photons, inp, out, got, want)
        # the fast kernel agrees with the double-double truth as well
        assert relerr(pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate(), got) < 1e-10
    U = haar_unitary(6, 5)
    ones, = [np.array([1, 1, 1, 0, 1, 0])]
    assert relerr(permanent_repeated_double_double(U, ones, ones[::-1].copy())[0],
                  ref.bbfg_repeated(U, ones, ones[::-1].copy(), True)) < 2e-15         # all singles
    assert permanent_repeated_double_double(U, np.zeros(6, int), np.zeros(6, int))[0] == 1.0
    with pytest.raises(Exception):
        permanent_repeated_double_double(U, [1, 0, 0, 0, 0, 0], [2, 0, 0, 0, 0, 0])
    a, b = np.array([2, 0, 1, 0, 1, 0]), np.array([0, 1, 1, 1, 0, 1])
    assert relerr(pqb.GlynnRepeatedInfPermanentCalculator(U, a, b).calculate(), ref.bbfg_repeated(U, a, b, True)) < 2e-15
