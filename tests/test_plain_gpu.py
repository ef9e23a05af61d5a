"""GPU parity tests of the plain Glynn/BBFG permanent through the C ABI / Python mirror.

Oracle: the reference's own BBFGPermanentCalculator in long double (oracle/_ref), as required by
SURVEY.md section 8(c).  Tolerance: 1e-10 relative (BASELINE.json north_star); the tests assert a
tighter 1e-11 so that regressions in the compensated accumulation are caught early.
"""
import os

import numpy as np
import pytest

from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-11


@pytest.fixture(scope="module")
def pqb():
    import piquassoboost_b200 as p
    p.load()
    return p


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 10, 12, 13, 16, 17, 20, 22])
def test_haar_unitary_vs_reference_long_double(pqb, ref, n):
    A = haar_unitary(n, n)
    got = pqb.GlynnPermanent(A).calculate()
    want = ref.bbfg(A, long_double=True)
    assert relerr(got, want) < TOL, (n, got, want)


@pytest.mark.parametrize("n", [24, 26])
def test_haar_submatrix_vs_reference_long_double(pqb, ref, n):
    A = np.ascontiguousarray(haar_unitary(2 * n, 2 * n)[:n, :n])
    got = pqb.BBFGPermanentLongDouble(A).calculate()
    want = ref.bbfg(A, long_double=True)
    assert relerr(got, want) < TOL, (n, got, want)


def test_all_plain_classes_agree(pqb, ref):
    A = haar_unitary(9, 99)
    want = ref.bbfg(A, True)
    for cls in (pqb.GlynnPermanent, pqb.GlynnPermanentDouble, pqb.BBFGPermanentDouble,
                pqb.BBFGPermanentLongDouble, pqb.GlynnPermanentSimpleDouble,
                pqb.GlynnPermanentSimpleLongDouble):
        assert relerr(cls(A).calculate(), want) < TOL


def test_edge_cases_match_reference(pqb, ref):
    # rows==0 or cols==0 -> 1 ; rows >= cols+2 -> 0 (BBFGPermanentCalculator.hpp:90-95)
    assert pqb.GlynnPermanent(np.zeros((0, 0), dtype=complex)).calculate() == 1
    assert pqb.GlynnPermanent(np.ones((4, 2), dtype=complex)).calculate() == 0
    one = np.array([[4.2 - 1j]])
    assert relerr(pqb.GlynnPermanent(one).calculate(), ref.bbfg(one)) < 1e-15
    # non-square inputs the reference accepts (rows <= cols+1): product over columns, sum over rows
    rng = np.random.default_rng(3)
    for shape in [(3, 5), (5, 4), (1, 3), (6, 6)]:
        A = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        assert relerr(pqb.GlynnPermanent(A).calculate(), ref.bbfg(A)) < TOL, shape


def test_non_contiguous_and_strided_input(pqb, ref):
    big = haar_unitary(12, 5)
    view = big[2:9, 3:10]              # non-contiguous view: wrapper must normalise it
    assert relerr(pqb.GlynnPermanent(view).calculate(), ref.bbfg(np.ascontiguousarray(view))) < TOL
    # explicit stride through the C ABI
    from piquassoboost_b200 import _lib
    out = np.zeros(2)
    sub = big[:7]                      # 7 x 12 contiguous; use the first 7 columns with stride 12
    _lib.check(_lib.load().pqb200_permanent(_lib.dptr(sub), 7, 7, 12, 1, _lib.dptr(out)))
    assert relerr(complex(out[0], out[1]), ref.bbfg(np.ascontiguousarray(sub[:, :7]))) < TOL


def test_structured_known_answers(pqb):
    import math
    # all-ones: n! ; rank one u v^T: n! prod u_i v_i ; identity: 1
    for n in (5, 11, 18):
        ones = np.ones((n, n), dtype=complex)
        assert relerr(pqb.GlynnPermanent(ones).calculate(), math.factorial(n)) < 1e-12
        rng = np.random.default_rng(n)
        u = rng.normal(size=n) + 1j * rng.normal(size=n)
        v = rng.normal(size=n) + 1j * rng.normal(size=n)
        want = math.factorial(n) * np.prod(u) * np.prod(v)
        assert relerr(pqb.GlynnPermanent(np.outer(u, v)).calculate(), want) < 1e-11
        assert abs(pqb.GlynnPermanent(np.eye(n)).calculate() - 1) < 1e-12


def test_list_of_matrices(pqb, ref):
    mats = [haar_unitary(n, 100 + n) for n in (3, 6, 4, 9)]
    got = pqb.GlynnPermanentCalculator_wrapper(mats, 7).calculate()
    assert isinstance(got, list) and len(got) == 4
    for g, m in zip(got, mats):
        assert relerr(g, ref.bbfg(m)) < TOL


def test_shard_invariance_and_combine(pqb):
    from piquassoboost_b200 import _lib
    lib = _lib.load()
    A = haar_unitary(21, 21)
    whole = pqb.GlynnPermanent(A).calculate()
    for shards in (1, 2, 3, 8):
        parts = np.zeros(4 * shards)
        for k in range(shards):
            out = np.zeros(4)
            _lib.check(lib.pqb200_permanent_shard(_lib.dptr(A), 21, 21, 21, 0, k, shards,
                                                  _lib.dptr(out), None, None))
            parts[4 * k:4 * k + 4] = out
        res = np.zeros(2)
        _lib.check(lib.pqb200_permanent_combine(_lib.dptr(parts), shards, 21, _lib.dptr(res)))
        assert relerr(complex(res[0], res[1]), whole) < 1e-13


def test_errors_are_raised_not_swallowed(pqb):
    with pytest.raises(Exception):
        pqb.GlynnPermanent(np.ones((46, 46), dtype=complex)).calculate()   # above the size limit
    with pytest.raises(Exception):
        pqb.GlynnPermanentCalculator_wrapper(np.eye(3), 3).calculate()    # DFE id


def test_inf_request_is_served_by_the_double_double_mode(pqb, ref):
    """lib 1 (GlynnInf: MPFR in the reference, BBFGPermanentCalculator.cpp:100-107) -> pqb200_permanent_dd."""
    A = haar_unitary(14, 141)
    assert relerr(pqb.GlynnPermanentInf(A).calculate(), ref.bbfg(A, True)) < 1e-15
    vals = pqb.GlynnPermanentCalculator_wrapper([A, A[:9, :9]], 1).calculate()
    assert relerr(vals[1], ref.bbfg(np.ascontiguousarray(A[:9, :9]), True)) < 1e-15


@pytest.mark.parametrize("L", [2, 3, 4, 5, 7, 9, 10])
def test_every_chunk_length_gives_the_same_value(pqb, ref, L):
    """The chunk length only changes the order of summation (and exercises the short-chunk paths of
    the kernel: chunks shorter than the double-double fold period, the last-pair dummy toggle)."""
    import os
    A = haar_unitary(11, 1234)
    want = ref.bbfg(A, True)
    os.environ["PQB200_FORCE_L"] = str(L)
    try:
        got = pqb.GlynnPermanent(A).calculate()
    finally:
        del os.environ["PQB200_FORCE_L"]
    assert relerr(got, want) < TOL, (L, got, want)


@pytest.mark.parametrize("N", list(range(3, 46)))
def test_every_kernel_instantiation(pqb, ref, N):
    """The kernel is instantiated per column count N with a per-N tuned loop form (static unroll /
    rolled, chain count): run each one, with and without the anchor re-seeding path (forced chunk
    length 2^12 > sub-chunk length 2^8), against the reference on a matrix with few rows."""
    import os
    rng = np.random.default_rng(1000 + N)
    if N <= 14:
        A = haar_unitary(N, N)
    else:  # 14 x (N-1): rows < cols runs in the N = cols+1 kernel (constant padding column)
        A = (rng.normal(size=(14, N - 1)) + 1j * rng.normal(size=(14, N - 1))) / np.sqrt(N - 1)
    want = ref.bbfg(np.ascontiguousarray(A), True)
    assert relerr(pqb.GlynnPermanent(A).calculate(), want) < TOL, N
    if A.shape[0] >= 14:
        os.environ["PQB200_FORCE_L"] = "12"
        try:
            got = pqb.GlynnPermanent(A).calculate()
        finally:
            del os.environ["PQB200_FORCE_L"]
        assert relerr(got, want) < TOL, N


def _golden_dd():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plain_dd_reference.json")
    return json.load(open(path))["cases"]


def test_double_double_evaluation_matches_reference_long_double(pqb, ref):
    """pqb200_permanent_dd (every operation in double-double) against the unmodified reference's long
    double calculator, live for n=18 and through the committed values the reference produced on CPU
    for n=32..36 (tests/golden/plain_dd_reference.json, tools/ref_big.py)."""
    from piquassoboost_b200.Boson_Sampling_Utilities import permanent_double_double
    A = haar_unitary(18, 18)
    v, parts = permanent_double_double(A)
    assert relerr(v, ref.bbfg(A, True)) < 1e-15
    gold = _golden_dd()
    A = haar_unitary(32, 32)
    v, parts = permanent_double_double(A, ndev=1)
    assert relerr(v, complex(*gold["32"]["ref_long_double"])) < 5e-15  # the long double's own error
    # deterministic, bit for bit, for a given device count (the golden value was made on one GPU; on more
    # devices the low words of the partial sums are grouped differently: the high words stay the same)
    assert [float(x).hex() for x in parts] == gold["32"]["dd_parts_hex"]
    v2, parts2 = permanent_double_double(A)
    assert [float(x).hex() for x in parts2[0::2]] == gold["32"]["dd_parts_hex"][0::2]


@pytest.mark.parametrize("n,tol", [(30, 1e-11), (34, 1e-11), (36, 1e-11), (38, 5e-11), (40, 5e-11)])
def test_large_haar_unitary_vs_committed_double_double(pqb, n, tol):
    """BASELINE sizes beyond the CPU reference's reach (n=40: ~2 h on 16 cores): the fast kernel against
    the committed double-double evaluation, itself pinned to the reference's long double up to n=36.
    north_star tolerance: 1e-10."""
    gold = _golden_dd()[str(n)]
    want = complex(*[float.fromhex(x) for x in gold["dd_parts_hex"]][0::2])
    got = pqb.GlynnPermanent(haar_unitary(n, n)).calculate()
    assert relerr(got, want) < tol, (n, got, want)
    if "ref_long_double" in gold:
        assert relerr(got, complex(*gold["ref_long_double"])) < tol


def test_concurrent_callers(pqb, ref):
    """The reference's calculators are called concurrently from TBB workers
    (samplingHelperFunctions.cpp:137, BatchedPermanentCalculator.cpp:343): the C ABI must be re-entrant
    (one mutex inside the library serialises the launches; results must not mix)."""
    import threading
    mats = [haar_unitary(n, 300 + n) for n in (5, 9, 12, 14, 15, 10, 7, 13)]
    want = [ref.bbfg(m, True) for m in mats]
    got = [None] * len(mats)
    errs = []

    def work(k):
        try:
            for _ in range(20):
                got[k] = pqb.GlynnPermanent(mats[k]).calculate()
                if relerr(got[k], want[k]) > TOL:
                    errs.append((k, got[k], want[k]))
        except Exception as e:  # noqa: BLE001
            errs.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(mats))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs[:3]


def test_block_diagonal_n42_full_size(pqb, ref):
    """A full-size run beyond n=40 with a known answer: perm(A (+) B) = perm(A) perm(B) for two 21x21
    Haar blocks (identity (ii)); 2^41 Gray steps (~30 s on one B200).  The zero
    blocks also put exact zeros into the row / column that the normalisation makes real."""
    A, B = haar_unitary(21, 2101), haar_unitary(21, 2102)
    M = np.zeros((42, 42), dtype=complex)
    M[:21, :21] = A
    M[21:, 21:] = B
    want = ref.bbfg(A, True) * ref.bbfg(B, True)
    got = pqb.GlynnPermanent(M).calculate()
    assert relerr(got, want) < 1e-10, (got, want)


def test_n40_haar_submatrix_vs_committed_double_double(pqb):
    """BASELINE config C3, second input class: the top-left 40x40 block of a Haar
    unitary of size 80, against the committed double-double evaluation (tools/accuracy_subblock.py)."""
    from scipy.stats import unitary_group
    gold = _golden_dd()["40_sub80"]
    want = complex(*[float.fromhex(x) for x in gold["dd_parts_hex"]][0::2])
    A = np.ascontiguousarray(unitary_group.rvs(80, random_state=80)[:40, :40])
    assert relerr(pqb.GlynnPermanent(A).calculate(), want) < 1e-11


def test_random_shapes_single_and_batched_against_the_oracle(pqb):
    """200 random shapes incl. the reference's trivial branches (empty -> 1, one row -> product,
    rows >= cols + 2 -> 0) and rectangular input, each through the single-matrix call and all of them
    through ONE list-of-matrices call (GlynnPermanentCalculator_Wrapper.hpp:204-293), against the C
    restatement of the reference (long double)."""
    from oracle import oraclebind
    rng = np.random.default_rng(77)
    mats = []
    for k in range(200):
        rows, cols = int(rng.integers(1, 13)), int(rng.integers(1, 13))
        mats.append(np.ascontiguousarray(rng.normal(size=(rows, cols)) + 1j * rng.normal(size=(rows, cols))))
    want = [oraclebind.bbfg(m, oraclebind.LONG_DOUBLE) for m in mats]
    batched = pqb.BBFGPermanentDouble(mats).calculate()
    assert len(batched) == len(mats)
    for m, w, b in zip(mats, want, batched):
        single = pqb.BBFGPermanentLongDouble(m).calculate()
        scale = max(abs(w), 1e-12)
        assert abs(single - w) <= 1e-11 * scale, (m.shape, single, w)
        assert abs(b - w) <= 1e-11 * scale, (m.shape, b, w)
    # a non-contiguous view goes through the same conversion as in the reference wrapper
    big = haar_unitary(14, 3)
    view = big[::2, ::2]
    assert relerr(pqb.GlynnPermanent(view).calculate(), oraclebind.bbfg(np.ascontiguousarray(view), oraclebind.LONG_DOUBLE)) < 1e-12
