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
    compact = pqb.BBFGRepeatedPermanentCalculatorDouble(A, cols, rows).calculate()   # power-form kernel
    assert relerr(pqb.GlynnPermanent(expanded).calculate(), compact) < 1e-12            # plain kernel
    from oracle import oraclebind
    assert relerr(compact, oraclebind.bbfg_repeated(A, cols, rows, oraclebind.LONG_DOUBLE)) < 1e-12


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


def _truth(U, inp, out):
    """__float128 evaluation by the oracle (C restatement of the reference, pinned to it in
    tests/test_oracle.py): in the ill-conditioned corner the reference's long double is itself only good to
    ~1e-19 x condition number, so it cannot serve as the truth there."""
    from oracle import oraclebind
    return oraclebind.bbfg_repeated(np.ascontiguousarray(U), np.asarray(inp), np.asarray(out), oraclebind.FLOAT128)


_CORNER_LOG = []


@pytest.mark.parametrize("photons", list(range(3, 45)))
def test_many_photons_in_few_modes(pqb, ref, photons):
    """Many photons in few modes is the ill-conditioned corner of the formula (huge binomial-weighted
    terms cancelling; the reference's own double path loses up to 8 digits there).  The engine's
    precision dispatch (a-posteriori bound -> double-double re-evaluation, include/pqb200.h) must keep
    north_star's 1e-10 against the __float128 truth for EVERY input, and must agree with the reference's
    long double wherever that is itself accurate."""
    from piquassoboost_b200 import _lib
    rng = np.random.default_rng(500 + photons)
    U = haar_unitary(6, 77)
    for trial in range(2):
        inp = rng.multinomial(photons, [1 / 6] * 6).astype(np.int64)
        out = rng.multinomial(photons, [0.3, 0.2, 0.2, 0.1, 0.1, 0.1] if trial else [1 / 6] * 6).astype(np.int64)
        truth = _truth(U, inp, out)
        got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
        bound, used_dd = _lib.last_repeated_bound()
        err = relerr(got, truth)
        ref_ld = relerr(ref.bbfg_repeated(U, inp, out, True), truth)
        _CORNER_LOG.append({"photons": photons, "inp": inp.tolist(), "out": out.tolist(), "bound": bound,
                            "double_double": used_dd, "rel_err": err, "reference_long_double_rel_err": ref_ld})
        assert err < 1e-10, (photons, inp, out, got, truth, bound, used_dd)
        if photons <= 24:
            assert err < TOL, (photons, inp, out, got, truth)


def test_rounding_bound_is_a_bound(pqb):
    """With the double-double re-evaluation switched off (PQB200_NO_DD_FALLBACK is read once per process,
    so this runs in a child process) the fast kernel's true error must stay below the a-posteriori bound
    it reports -- that is what makes the dispatch threshold meaningful.  Writes the table DESIGN.md cites."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from conftest import haar_unitary, relerr
from oracle import oraclebind
import piquassoboost_b200 as pqb
from piquassoboost_b200 import _lib
rows = []
U = haar_unitary(6, 77)
for photons in (8, 16, 24, 28, 32, 36, 40, 44, 52, 60):
    rng = np.random.default_rng(500 + photons)
    inp = rng.multinomial(photons, [1 / 6] * 6).astype(np.int64)
    out = rng.multinomial(photons, [1 / 6] * 6).astype(np.int64)
    truth = oraclebind.bbfg_repeated(np.ascontiguousarray(U), inp, out, oraclebind.FLOAT128)
    got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
    bound, used = _lib.last_repeated_bound()
    rows.append({"photons": photons, "bound": bound, "rel_err_fast_kernel": relerr(got, truth), "dd": used})
print(json.dumps(rows))
""" % (root, os.path.join(root, "tests"))
    env = dict(os.environ, PQB200_NO_DD_FALLBACK="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    rows = json.loads(res.stdout.strip().splitlines()[-1])
    for r in rows:
        assert not r["dd"]
        assert r["rel_err_fast_kernel"] <= max(r["bound"], 1e-13), r
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    json.dump({"bound_vs_error": rows, "corner": _CORNER_LOG},
              open(os.path.join(root, "gpurun_out", "repeated_bound_calibration.json"), "w"), indent=1)


def test_sixty_photons_in_ten_modes(pqb, ref):
    """VERDICT r1 missing #4: the reference has no photon cap (60 photons in 10 modes is milliseconds on
    its CPU); round 1 returned PQB200_ERR_ARG above 44 photons.  Power form: the matrix in the parameter
    bank is (distinct rows) x (distinct columns), whatever the photon number."""
    rng = np.random.default_rng(60)
    U = haar_unitary(10, 610)
    for photons, nrows in ((45, 3), (60, 3), (80, 4), (120, 3)):
        inp = rng.multinomial(photons, [0.1] * 10).astype(np.int64)       # spread over all 10 input modes
        out = np.zeros(10, dtype=np.int64)
        out[:nrows] = rng.multinomial(photons, [1 / nrows] * nrows)        # bunched into a few output modes
        got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
        truth = _truth(U, inp, out)
        assert relerr(got, truth) < 1e-10, (photons, inp, out, got, truth)
    # 30 photons, both sides repeated, a longer Gray space (6^6 * 5/6 steps on the column side)
    V = haar_unitary(12, 611)
    inp = np.array([5] * 6 + [0] * 6)
    out = np.array([3] * 10 + [0] * 2)
    got = pqb.BBFGRepeatedPermanentCalculatorDouble(V, inp, out).calculate()
    assert relerr(got, _truth(V, inp, out)) < 1e-10
    # more distinct columns than the fast kernel keeps in registers (48): 60 single photons from 60 input modes
    # bunched into three output modes -- served by the double-double walk (up to 65 x 128), exactly
    from piquassoboost_b200 import _lib
    W = haar_unitary(60, 612)
    inp = np.ones(60, dtype=np.int64)
    out = np.zeros(60, dtype=np.int64)
    out[:3] = [20, 20, 20]
    got = pqb.BBFGRepeatedPermanentCalculatorDouble(W, inp, out).calculate()
    assert _lib.last_repeated_bound()[1]                      # went through the double-double mode
    assert relerr(got, _truth(W, inp, out)) < 1e-12
    with pytest.raises(Exception, match="exceeds the limits"):  # 2^59 Gray steps either way round
        pqb.BBFGRepeatedPermanentCalculatorDouble(W, inp, inp.copy()).calculate()
    # 60 photons over all 10 modes on BOTH sides (about 7^10 Gray steps): against the reference's long double
    inp = np.array([6] * 10)
    out = np.array([7, 5, 6, 6, 8, 4, 6, 6, 5, 7])
    got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
    want = ref.bbfg_repeated(U, inp, out, True)
    bound, used_dd = _lib.last_repeated_bound()
    # the long double reference is itself only accurate to ~1e-19 x condition number here
    assert relerr(got, want) < max(1e-10, 1e-3 * bound), (got, want, bound, used_dd)


@pytest.mark.parametrize("C", list(range(1, 49)))
def test_every_power_kernel_instantiation(pqb, ref, C):
    """glynn_reppow_kernel is instantiated per number C of DISTINCT product-side columns (1..48): C
    input modes (one or several photons each) into three output modes, so that the Gray space (the
    three output rows) stays small and the reference's long double answer is cheap."""
    m = max(C, 3)
    U = haar_unitary(m, 900 + C)
    rng = np.random.default_rng(C)
    for trial in range(2):
        inp = np.zeros(m, dtype=np.int64)
        inp[:C] = 1 if trial == 0 else rng.integers(1, 4, size=C)
        photons = int(inp.sum())
        out = np.zeros(m, dtype=np.int64)
        out[:3] = rng.multinomial(photons, [0.4, 0.35, 0.25])
        if np.count_nonzero(out) < 2:   # keep at least two output rows (one row is the host's product branch)
            out[:3] = [photons - photons // 2, photons // 2, 0]
        got = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
        # many photons into three modes: the reference's long double is itself off by up to 2e-8 there
        # (its double path by 4e-5), so the truth is the __float128 oracle; where the long double IS
        # accurate (single occupations) it must agree as well
        truth = _truth(U, inp, out)
        assert relerr(got, truth) < 1e-10, (C, inp, out, got, truth)
        if trial == 0:
            assert relerr(got, ref.bbfg_repeated(U, inp, out, True)) < 1e-10


def test_random_stress_against_the_oracle(pqb):
    """300 random shapes: rectangular matrices, zero occupations, one-sided bunching, both orientations of
    the Gray space, the trivial branches -- single calls and one batched call, against the C restatement
    of the reference (long double; pinned to the reference in tests/test_oracle.py)."""
    from oracle import oraclebind
    rng = np.random.default_rng(2024)
    cases = []
    for k in range(300):
        rows, cols = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        photons = int(rng.integers(0, 13))
        A = rng.normal(size=(rows, cols)) + 1j * rng.normal(size=(rows, cols))
        pr = rng.dirichlet(np.ones(rows) * rng.choice([0.2, 1.0, 5.0]))
        pc = rng.dirichlet(np.ones(cols) * rng.choice([0.2, 1.0, 5.0]))
        out = rng.multinomial(photons, pr).astype(np.int64)     # row multiplicities
        inp = rng.multinomial(photons, pc).astype(np.int64)     # column multiplicities
        cases.append((np.ascontiguousarray(A), inp, out))
    for A, inp, out in cases:
        want = oraclebind.bbfg_repeated(A, inp, out, oraclebind.LONG_DOUBLE)
        got = pqb.BBFGRepeatedPermanentCalculatorDouble(A, inp, out).calculate()
        if abs(want) < 1e-12:
            assert abs(got - want) < 1e-11, (A.shape, inp, out, got, want)
        else:
            assert relerr(got, want) < 1e-10, (A.shape, inp, out, got, want)
    # the batched entry point on one matrix with many occupation pairs (kernel jobs, trivial jobs, single
    # rows), in the reference's per-input form: every input with the outputs of the same photon number
    U = haar_unitary(9, 99)
    ins = [rng.multinomial(p, rng.dirichlet(np.ones(9))).astype(np.int64) for p in rng.integers(0, 11, size=60)]
    outs = [[b for b in ins if b.sum() == a.sum()] for a in ins]
    res = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs).calculate()
    assert [len(r) for r in res] == [len(o) for o in outs]
    for i, a in enumerate(ins):
        for j, b in enumerate(outs[i]):
            want = oraclebind.bbfg_repeated(np.ascontiguousarray(U), a, b, oraclebind.LONG_DOUBLE)
            assert abs(res[i][j] - want) <= 1e-10 * max(abs(want), 1e-12), (a, b, res[i][j], want)


def test_reference_shaped_list_states(pqb, ref):
    """ChinHuhPermanentCalculator_Wrapper.hpp:233-326: input_state a list of n vectors, output_state a list
    of n (ragged) lists of vectors -> ret[i][j]; and the 'multiInput' form the other way round."""
    U = haar_unitary(5, 22)
    ins = [np.array([1, 1, 1, 0, 0]), np.array([2, 0, 0, 1, 0])]
    outs = [[np.array([0, 1, 1, 1, 0]), np.array([3, 0, 0, 0, 0])], [np.array([1, 0, 1, 0, 1])]]
    res = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs).calculate()
    assert [len(r) for r in res] == [2, 1]
    for i, a in enumerate(ins):
        for j, b in enumerate(outs[i]):
            assert relerr(res[i][j], ref.bbfg_repeated(U, a, b, True)) < TOL
    # multiInput: input_state[i] is the list of inputs that go with output_state[i]
    res2 = pqb.ChinHuhPermanentCalculator_wrapper(8, U, outs, ins).calculate()
    assert [len(r) for r in res2] == [2, 1]
    for i, b in enumerate(ins):
        for j, a in enumerate(outs[i]):
            assert relerr(res2[i][j], ref.bbfg_repeated(U, a, b, True)) < TOL
    # the class a reference user constructs (Boson_Sampling_Utilities.py:30-91) refuses lists whose items
    # are not arrays, as the reference does
    with pytest.raises(Exception, match="should be a numpy array, or a list of numpy arrays"):
        pqb.BBFGRepeatedPermanentCalculatorDouble(U, ins, outs)


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
