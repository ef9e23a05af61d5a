"""GPU parity tests of the batched sub-permanent path: pmf layer and the Clifford-Clifford B sampler
against the reference's seeded golden shots (unittests/test_seeding.py) and compute_pmf rows."""
import numpy as np
import pytest

import golden_util as G
from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pqb():
    import piquassoboost_b200 as p
    p.load()
    return p


def test_pmf_rows_match_reference_golden(pqb):
    g = G.load("sampling_seed123.json")
    U = G.mat(g["U"])
    for row in g["pmf_rows"]:
        pmf = pqb.compute_pmf(U, row["sample"], row["current_input"])
        assert np.allclose(pmf, row["pmf"], rtol=1e-11, atol=1e-15), row
        assert abs(pmf.sum() - 1) < 1e-13


def test_seeded_sampler_reproduces_reference_shots(pqb):
    """seed 123, input [2,1,1,0,1], 20 shots: exact equality with the C++ generation order of the
    unmodified reference, and multiset equality with unittests/test_seeding.py:62-83."""
    g = G.load("sampling_seed123.json")
    U = G.mat(g["U"])
    strat = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=g["seed"])
    shots = strat.simulate(np.array(g["input"], dtype=np.int64), g["shots"])
    assert isinstance(shots, tuple) and len(shots) == g["shots"]
    got = [s.tolist() for s in shots]
    assert got == g["cpp_generation_order"]
    assert sorted(map(tuple, got)) == sorted(map(tuple, g["expected_python_order"]))


def test_pmf_layer_many_shots_vs_oracle(pqb, ref):
    rng = np.random.default_rng(3)
    m, k = 9, 6
    U = haar_unitary(m, 17)
    S = 40
    samples = np.zeros((S, m), dtype=np.int64)
    inputs = np.zeros((S, m), dtype=np.int64)
    for s in range(S):
        samples[s] = rng.multinomial(k - 1, [1 / m] * m)
        inputs[s] = rng.multinomial(k, [1 / m] * m)
    pmf = pqb.compute_pmf(U, samples, inputs)
    for s in range(S):
        want = ref.compute_pmf(U, samples[s], inputs[s])
        assert np.allclose(pmf[s], want, rtol=1e-10, atol=1e-14), s


def test_sampler_larger_instance_matches_oracle_shots(pqb, ref):
    # 8 photons in 12 modes, 25 shots: identical shots to the unmodified reference sampler
    U = haar_unitary(12, 4)
    inp = np.array([1, 1, 0, 2, 0, 1, 0, 1, 1, 0, 1, 0], dtype=np.int64)
    want = ref.simulate_B(U, 77, inp, 25)
    got = np.array(pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=77).simulate(inp, 25))
    assert (got == want).all()
    assert (got.sum(axis=1) == inp.sum()).all()


def test_batch_of_small_plain_matrices(pqb, ref):
    mats = [haar_unitary(n, 500 + n) for n in (2, 3, 5, 8, 11, 14, 17, 20, 7, 7)]
    got = pqb.GlynnPermanentCalculator_wrapper(mats, 7).calculate()
    for g_, m_ in zip(got, mats):
        assert relerr(g_, ref.bbfg(m_, True)) < 1e-11


def test_repeated_batch_many_states(pqb, ref):
    rng = np.random.default_rng(8)
    m = 10
    U = haar_unitary(m, 9)
    ins = [rng.multinomial(7, [1 / m] * m) for _ in range(12)]
    outs = [rng.multinomial(7, [1 / m] * m) for _ in range(9)]
    res = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs).calculate()
    for i, a in enumerate(ins):
        for j, b in enumerate(outs):
            want = ref.bbfg_repeated(U, a, b, True)
            assert abs(res[i][j] - want) <= 1e-11 * max(abs(want), 1e-12), (i, j)


def test_many_shots_exercise_the_threaded_host_path(pqb, ref):
    """>= 128 shots switch the host-side job construction to several threads; the shots must still
    be exactly the reference's."""
    U = haar_unitary(8, 31)
    inp = np.array([1, 1, 1, 0, 1, 0, 1, 0], dtype=np.int64)
    want = ref.simulate_B(U, 5, inp, 300)
    got = np.array(pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=5).simulate(inp, 300))
    assert (got == want).all()


def test_waves_and_cost_ordered_second_launch_give_the_same_shots(pqb, ref):
    """The launch policy (several waves per device, the last photon steps as a second, cost-ordered launch
    that overlaps the replay of the random stream, clusters of CTAs per shot) must not change a single shot:
    2600 shots of 12 photons in 16 modes under the default policy, as one plain launch, and with forced
    clusters / waves, against each other and -- the first 40 -- against the unmodified reference."""
    import os
    U = np.ascontiguousarray(haar_unitary(16, 160))
    state = np.array([1] * 12 + [0] * 4, dtype=np.int64)
    knobs = ("PQB200_LAP_TAIL", "PQB200_LAP_WAVES", "PQB200_LAP_CLUSTER")
    saved = {k: os.environ.pop(k, None) for k in knobs}

    def run(shots=2600, **env):
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in env.items()})
        sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=77)
        return np.array(sim.simulate(state, shots)), pqb.last_stats()["launches"]

    try:
        default, launches = run()
        assert launches >= 4                                   # two waves x two launches
        single, launches1 = run(PQB200_LAP_TAIL=0, PQB200_LAP_WAVES=1)
        assert launches1 == 1
        assert np.array_equal(default, single)
        assert np.array_equal(default, run(PQB200_LAP_TAIL=4, PQB200_LAP_WAVES=3)[0])
        small = run(shots=120)[0]
        assert np.array_equal(small, default[:120])
        for cl in (1, 3, 8):
            assert np.array_equal(run(shots=120, PQB200_LAP_CLUSTER=cl)[0], small)
        assert np.array_equal(default[:40], ref.simulate_B(U, 77, state, 40))
        assert np.all(default.sum(axis=1) == 12)
    finally:
        for k in knobs:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
