"""CPU tests of the HOST logic in front of the kernels (no GPU): the unit-modulus normalisation of
plain_host.h / rep_normalise and the repeated-row planning of repeated_host.h (step table, base vectors,
binomials, chunk decomposition).  The headers are compiled into a small test library
(tests/cpp/host_logic.cpp) and the kernels' sweeps are emulated in numpy / Python integers; the oracle
(C restatement of the reference, long double) is the checker."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import haar_unitary, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lp = ctypes.POINTER(ctypes.c_int64)


@pytest.fixture(scope="module")
def hl(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hl") / "libhostlogic.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", so,
                    os.path.join(ROOT, "tests", "cpp", "host_logic.cpp")], check=True, capture_output=True)
    return ctypes.CDLL(so)


def _oracle(A):
    from oracle import oraclebind
    return oraclebind.bbfg(np.ascontiguousarray(A), oraclebind.LONG_DOUBLE)


def _normalise(hl, A):
    A = np.ascontiguousarray(A, dtype=np.complex128)
    rows, cols = A.shape
    N = hl.hl_plain_cols(rows, cols)
    a2 = np.zeros((N, N), dtype=np.complex128)
    phase = np.zeros(4)
    hl.hl_plain_normalise(A.ctypes.data_as(_dp), rows, cols, ctypes.c_size_t(cols), N, a2.ctypes.data_as(_dp),
                          phase.ctypes.data_as(_dp))
    return N, a2, complex(phase[0] + phase[1], phase[2] + phase[3])


@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (7, 7), (12, 12), (5, 4), (4, 6), (9, 9)])
def test_plain_normalisation_keeps_the_permanent(hl, shape):
    """The kernel's contract (real Gray-bit-0 row, real last column, kernel row order) and the value:
    the Glynn sum of the normalised matrix times the correction factor is the reference's value."""
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    A = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    if shape == (9, 9):
        A[8, 3] = 0.0      # zeros in the row / column that is made real must be harmless
        A[2, 8] = 0.0
    rows, cols = shape
    N, a2, corr = _normalise(hl, A)
    assert abs(abs(corr) - 1.0) < 1e-15
    # the real row (index 0) is stored compactly: N real parts in the first N doubles of the row
    packed = a2[0].view(np.float64).copy()
    assert np.all(packed[N:] == 0.0)
    a2[0] = packed[:N]
    assert np.all(a2[:, N - 1].imag == 0.0)
    assert np.allclose(np.abs(a2[:rows, :cols]), 2 * np.abs(np.vstack([A[1:][::-1], A[:1]])), rtol=1e-15)
    # emulate the kernel's sum: row rows-1 has delta = +1, row k is toggled by Gray bit k
    total = 0.0
    for g in range(1 << (rows - 1)):
        gray = g ^ (g >> 1)
        c = 0.5 * a2[rows - 1].copy()
        for k in range(rows - 1):
            c += (-0.5 if (gray >> k) & 1 else 0.5) * a2[k]
        total += (-1) ** bin(gray).count("1") * np.prod(c)
    got = total * corr / 2.0 ** (rows - 1)
    assert relerr(got, _oracle(A)) < 1e-12


def _rep_plan(hl, U, inp, out, max_len, normalise=1):
    U = np.ascontiguousarray(U, dtype=np.complex128)
    rows, cols = U.shape
    rm = np.ascontiguousarray(out, dtype=np.int64)
    cm = np.ascontiguousarray(inp, dtype=np.int64)
    P = int(rm.sum())
    mat = np.zeros((P + 1, max(P, 1)), dtype=np.complex128)
    r = np.zeros(64, dtype=np.int32)
    cmv = np.zeros(64, dtype=np.int32)
    tw = np.zeros(max_len + 1)
    trow = np.zeros(max_len + 1, dtype=np.int32)
    tfl = np.zeros(max_len + 1, dtype=np.int32)
    base = np.zeros((2, max(P, 1)), dtype=np.complex128)
    radix = np.zeros(64, dtype=np.int32)
    weight = np.zeros(64, dtype=np.uint64)
    binom = np.zeros(64 * 64)
    phase = np.zeros(4)
    misc = np.zeros(8, dtype=np.int64)
    triv = np.zeros(2)
    rc = hl.hl_rep_plan(U.ctypes.data_as(_dp), ctypes.c_size_t(rows), ctypes.c_size_t(cols), ctypes.c_size_t(cols),
                        rm.ctypes.data_as(_lp), cm.ctypes.data_as(_lp), max_len, normalise, triv.ctypes.data_as(_dp),
                        mat.ctypes.data_as(_dp), r.ctypes.data_as(_ip), cmv.ctypes.data_as(_ip), tw.ctypes.data_as(_dp), trow.ctypes.data_as(_ip),
                        tfl.ctypes.data_as(_ip), base.ctypes.data_as(_dp), radix.ctypes.data_as(_ip),
                        weight.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), binom.ctypes.data_as(_dp),
                        phase.ctypes.data_as(_dp), misc.ctypes.data_as(_lp))
    return rc, dict(mat=mat, r=r, cm=cmv, tw=tw, trow=trow, tfl=tfl, base=base, radix=radix, weight=weight,
                    binom=binom.reshape(64, 64), phase=complex(phase[0] + phase[1], phase[2] + phase[3]),
                    R=int(misc[0]), N=int(misc[1]), q=int(misc[2]), len=int(misc[3]), nchunks=int(misc[4]),
                    lowpar=(int(misc[5]), int(misc[6])), photons=int(misc[7]), triv=complex(triv[0], triv[1]))


def _emulate_repeated_kernel(pl):
    """glynn_reppow_kernel in Python: one 'thread' per chunk, term t (product of the column sums raised
    to their multiplicities) then the toggle of table entry t+1."""
    R, N, q, length = pl["R"], pl["N"], pl["q"], pl["len"]
    D = R - 1
    cm = pl["cm"][:N].astype(np.int64)
    mat = pl["mat"].reshape(-1)[:R * N].reshape(R, N)
    A2 = 2.0 * mat
    total = 0.0
    for cc in range(pl["nchunks"]):
        par, ksum, whigh = 0, 0, 1.0
        ks = {}
        for d in range(D - 1, q - 1, -1):
            radix = int(pl["radix"][d])
            cnt = (cc // int(pl["weight"][d])) % radix
            k = radix - 1 - cnt if par else cnt
            par ^= k & 1
            ksum += k
            whigh *= pl["binom"][d, k]
            ks[d] = k
        P = par
        c = pl["base"].reshape(-1)[:2 * N].reshape(2, N)[P].copy()
        for d in range(D - 1, q - 1, -1):
            c += 0.5 * (int(pl["radix"][d]) - 1 - 2 * ks[d]) * A2[d + 1]
        acc = 0.0
        for t in range(length):
            acc += pl["tw"][t] * np.prod(c ** cm)
            if t + 1 < length:
                row, fl = int(pl["trow"][t + 1]), int(pl["tfl"][t + 1])
                c += (-1.0 if (fl >> P) & 1 else 1.0) * A2[row]
        f = -whigh if (ksum + pl["lowpar"][P]) & 1 else whigh
        total += f * acc
    return total * pl["phase"] / 2.0 ** (pl["photons"] - 1)


@pytest.mark.parametrize("case", [
    ([2, 1, 0, 1], [1, 1, 2, 0], 64),
    ([3, 0, 2, 0, 1], [1, 2, 1, 1, 1], 12),      # short chunks: several high digits
    ([0, 4, 0, 3], [2, 2, 2, 1], 5),
    ([2, 2, 2, 2, 0, 0], [1, 1, 1, 1, 2, 2], 9),
    ([5, 0, 0, 0], [1, 1, 1, 2], 100),             # a single repeated column after transposition
    ([6, 5, 0, 0], [4, 4, 3, 0], 50),              # power form: multiplicities up to 6 on both sides
    ([3, 3, 3, 3], [12, 0, 0, 0], 20),             # one column twelve times
    ([2, 2, 3, 1, 1, 0], [1, 1, 1, 2, 2, 2], 30),  # mixed multiplicities: sorted descending, largest last
])
@pytest.mark.parametrize("normalise", [0, 1])
def test_repeated_plan_tables_reproduce_the_reference(hl, case, normalise):
    from oracle import oraclebind
    inp, out, max_len = case
    m = len(inp)
    U = haar_unitary(m, 40 + m)
    rc, pl = _rep_plan(hl, U, inp, out, max_len, normalise)
    assert rc == 0, rc
    assert pl["len"] <= max_len
    cm = pl["cm"][:pl["N"]]
    assert cm.sum() == pl["photons"] and cm[-1] == cm.max() and np.all(np.diff(cm[:-1]) <= 0)
    if normalise:
        mat = pl["mat"].reshape(-1)[:pl["R"] * pl["N"]].reshape(pl["R"], pl["N"])
        assert np.all(mat[1].imag == 0.0) and np.all(mat[:, -1].imag == 0.0)
    want = oraclebind.bbfg_repeated(np.ascontiguousarray(U), np.array(inp), np.array(out), oraclebind.LONG_DOUBLE)
    assert relerr(_emulate_repeated_kernel(pl), want) < 1e-12


def test_repeated_prepare_edge_cases(hl):
    U = haar_unitary(4, 3)
    rc, pl = _rep_plan(hl, U, [0, 0, 0, 0], [0, 0, 0, 0], 16)
    assert rc == 1 and pl["triv"] == 1.0                      # empty permanent is 1
    rc, _ = _rep_plan(hl, U, [1, 1, 0, 0], [1, 0, 0, 0], 16)
    assert rc == -3                                           # photon-number mismatch
    rc, pl = _rep_plan(hl, U, [1, 1, 1, 0], [0, 1, 1, 1], 16)
    assert rc == 2                                            # all singles: plain route
    rc, pl = _rep_plan(hl, U, [0, 0, 1, 0], [1, 0, 0, 0], 16)
    assert rc == 1                                            # single row: product of entries
    assert relerr(pl["triv"], U[0, 2]) < 1e-15
    rc, pl = _rep_plan(hl, U, [0, 2, 0, 1], [3, 0, 0, 0], 16)  # one row three times: 3! * prod
    assert rc == 0
    assert relerr(_emulate_repeated_kernel(pl), 6 * U[0, 1] ** 2 * U[0, 3]) < 1e-13


def test_rng_replay_header_matches_std_mt19937(tmp_path):
    """csrc/rng_replay.h (block generator, remainder by multiplication, PDEP / portable photon pick) against
    std::mt19937 -- the reference's generator (samplingHelperFunctions.cpp:58-74) -- and a plain restatement of
    the per-step draws: the stream, random_index, random_double_0_1, the photon loops and the stream position
    after them (tests/cpp/rng_replay.cpp)."""
    exe = str(tmp_path / "rng_replay")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "rng_replay.cpp")],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_sampler_launch_policy(hl, monkeypatch):
    """laplace_host.h: clusters only while all their CTAs are resident at once (3 CTAs per SM), a second,
    cost-ordered launch only when the shots fill the device, never both (the second launch runs one CTA per
    shot), and the tuning knobs override the choice."""
    for k in ("PQB200_LAP_CLUSTER", "PQB200_LAP_TAIL"):
        monkeypatch.delenv(k, raising=False)
    hl.hl_lap_cluster_for.argtypes = [ctypes.c_size_t, ctypes.c_int]
    hl.hl_lap_tail_steps.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    sms, slots = 148, 148 * 3
    for shots in list(range(1, 700)) + [1000, 2664, 2665, 5000, 20000, 10**6]:
        cs = hl.hl_lap_cluster_for(shots, sms)
        tail = hl.hl_lap_tail_steps(shots, 20, sms)
        assert 1 <= cs <= 8
        assert cs == 1 or cs * shots <= slots            # every CTA of every cluster resident in one wave
        assert cs > 1 or shots * 2 > slots                # SMs to spare are used
        assert tail in (0, 4) and not (tail and cs > 1)
        assert (tail == 4) == (slots <= shots <= 6 * slots)
    assert hl.hl_lap_cluster_for(125, sms) == 3 and hl.hl_lap_cluster_for(25, sms) == 8
    assert hl.hl_lap_cluster_for(60, sms) == 4          # 8 CTAs per shot only with SMs to spare
    assert hl.hl_lap_tail_steps(1000, 8, sms) == 0      # few photons: one launch
    monkeypatch.setenv("PQB200_LAP_TAIL", "2")
    monkeypatch.setenv("PQB200_LAP_CLUSTER", "2")
    assert hl.hl_lap_tail_steps(100, 20, sms) == 2 and hl.hl_lap_cluster_for(1000, sms) == 2


def _laplace_pmf(U, sample, cur):
    """laplace_sampler_kernel's photon step in numpy (csrc/laplace_sampler.cuh): ONE mixed-radix Gray sweep over
    the row multiplicities; every term is the vector of column sums with one factor left out (prefix / suffix
    products), which yields all k sub-permanents of compute_pmf (samplingHelperFunctions.cpp:137-172) at once."""
    import itertools
    from math import comb
    m = len(sample)
    cols = [j for j in range(m) for _ in range(int(cur[j]))]          # physically repeated columns
    first = [i == 0 or cols[i] != cols[i - 1] for i in range(len(cols))]
    N, K = len(cols), int(np.sum(sample))
    if K == 0:
        addend = np.ones(N, dtype=complex)
    else:
        occ = [i for i in range(m) if sample[i] > 0]
        amin = min(occ, key=lambda i: (sample[i], i))                   # anchor: first minimum-multiplicity row
        rows = [(i, int(sample[i]) - (1 if i == amin else 0)) for i in occ]
        rows = [(i, r) for i, r in rows if r > 0]
        A0 = U[amin, cols]
        addend = np.zeros(N, dtype=complex)
        for ks in itertools.product(*[range(r + 1) for _, r in rows]):
            s = A0.copy()
            w = 1.0
            for (i, r), k in zip(rows, ks):
                s = s + (r - 2 * k) * U[i, cols]
                w *= comb(r, k) * (-1) ** k
            pre = np.concatenate(([1.0 + 0j], np.cumprod(s)[:-1]))
            suf = np.concatenate((np.cumprod(s[::-1])[::-1][1:], [1.0 + 0j]))
            addend += w * pre * suf                                     # leave-one-out products of the term
        addend *= 2.0 ** -(K - 1)
    amp = np.zeros(m, dtype=complex)
    for j in range(N):
        if first[j]:
            amp += cur[cols[j]] * addend[j] * U[:, cols[j]]
    p = np.abs(amp) ** 2
    return p / p.sum()


@pytest.mark.parametrize("seed", range(6))
def test_laplace_sweep_restatement_matches_oracle_pmf(seed):
    """The sampler kernel's algorithm (one sweep, leave-one-out products: Clifford & Clifford's Lemma 1 as in
    piquasso-module/src/permanent_laplace.cpp:43-233) against the oracle's compute_pmf, which evaluates the k
    repeated sub-permanents one by one as the reference does -- including collisions on both sides."""
    from oracle import oraclebind
    rng = np.random.default_rng(seed)
    m = 5 + seed % 4
    U = haar_unitary(m, 40 + seed)
    for _ in range(6):
        k = int(rng.integers(1, 7))                                    # photons inserted so far
        cur = np.bincount(rng.integers(0, m, size=k), minlength=m)
        sample = np.bincount(rng.integers(0, max(2, m // 2), size=k - 1), minlength=m)  # k-1 placed, with collisions
        want = oraclebind.compute_pmf(U, sample, cur)
        got = _laplace_pmf(U, sample, cur)
        assert np.allclose(got, want, rtol=1e-10, atol=1e-14), (sample, cur)


def test_double_double_primitives_against_float128(tmp_path):
    """csrc/dd.cuh compiled for the host (tests/cpp/dd_arith.cpp): TwoSum is error-free, dd add / multiply and
    the complex dd product are accurate to 2^-100 against __float128, a 2e6-term cancelling sum keeps 2^-95 of
    sum|x| (plain double: 1e-17) -- the accumulation every kernel uses and the arithmetic of pqb200_permanent_dd,
    which serves the reference's long double / MPFR requests."""
    exe = str(tmp_path / "dd_arith")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "dd_arith.cpp")],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def _batch_job(hl, rm, cm):
    rm = np.ascontiguousarray(rm, dtype=np.int64)
    cm = np.ascontiguousarray(cm, dtype=np.int64)
    head = np.zeros(8, dtype=np.int64)
    gidx, radix, pidx = (np.zeros(64, dtype=np.int32) for _ in range(3))
    rc = hl.hl_batch_job(ctypes.c_size_t(len(rm)), ctypes.c_size_t(len(cm)), rm.ctypes.data_as(_lp), cm.ctypes.data_as(_lp),
                         head.ctypes.data_as(_lp), gidx.ctypes.data_as(_ip), radix.ctypes.data_as(_ip), pidx.ctypes.data_as(_ip))
    kind, R, N, q, length, nchunks, tr = (int(x) for x in head[:7])
    return rc, dict(kind=kind, R=R, N=N, q=q, len=length, nchunks=nchunks, transpose=tr,
                    gidx=gidx[:R].tolist(), radix=radix[:max(R - 1, 0)].tolist(), pidx=pidx[:N].tolist())


def test_batch_job_construction_matches_oracle(hl):
    """batch_host.h builds a job of the many-small-permanents path from the two states alone (index lists: the
    gather of adaptInterferometer, BatchedPermanentCalculator.cpp:412-451, happens in the kernel prologue).
    glynn_batch_kernel's sum is emulated from the job in numpy and compared with the oracle's repeated
    permanent; the Gray side is the one with the smaller space, the digit split covers it exactly."""
    import itertools
    from math import comb
    from oracle import oraclebind
    rng = np.random.default_rng(11)
    m = 7
    U = haar_unitary(m, 77)
    kinds = set()
    for case in range(60):
        photons = int(rng.integers(0, 7))
        rm = np.bincount(rng.integers(0, m, size=photons), minlength=m)
        cm = np.bincount(rng.integers(0, 3 if case % 2 else m, size=photons), minlength=m)
        rc, job = _batch_job(hl, rm, cm)
        assert rc == 0
        kinds.add(job["kind"])
        want = oraclebind.bbfg_repeated(U, cm, rm)      # (matrix, input = columns, output = rows)
        if job["kind"] == 1:                             # empty permanent
            assert photons == 0 and want == 1
            continue
        assert job["kind"] in (0, 2) and job["N"] == photons
        el = (lambda g, p: U[p, g]) if job["transpose"] else (lambda g, p: U[g, p])
        A = np.array([[el(g, p) for p in job["pidx"]] for g in job["gidx"]])
        if job["kind"] == 2:                             # a single row of multiplicity 1: the product of its entries
            assert job["R"] == 1 and photons == 1
            got = np.prod(A[0])
        else:
            total = int(np.prod(job["radix"]))
            assert job["len"] * job["nchunks"] == total and job["len"] == int(np.prod(job["radix"][:job["q"]]))
            got = 0.0
            for ks in itertools.product(*[range(r) for r in job["radix"]]):
                s = A[0].copy()
                w = 1.0
                for d, k in enumerate(ks):
                    r = job["radix"][d] - 1
                    s = s + (r - 2 * k) * A[d + 1]
                    w *= comb(r, k) * (-1) ** k
                got += w * np.prod(s)
            got *= 2.0 ** -(photons - 1)
            # the Gray side is the smaller space
            space = lambda mult: np.prod([v + 1 for v in mult if v > 0]) * min(v for v in mult if v > 0) / (min(v for v in mult if v > 0) + 1)
            assert (space(cm) < space(rm)) == bool(job["transpose"])
        assert relerr(got, want) < 1e-12, (rm, cm, job)
    assert kinds >= {0, 1}
    # error codes of the reference: unequal photon numbers, negative occupation
    assert _batch_job(hl, [1, 0, 2], [1, 1, 0])[0] == 3
    assert _batch_job(hl, [1, -1, 0], [0, 0, 0])[0] == 2
