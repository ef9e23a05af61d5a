"""The in-process multi-device route (`ndev` argument) and the re-entrancy of the C ABI: results must
not depend on how a call is sharded, small automatic calls must stay on ONE GPU, and concurrent callers
must not disturb each other (missing #1, weak #4, #5; ADVICE r1: stream race)."""
import ctypes
import threading

import numpy as np
import pytest

from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pqb():
    import piquassoboost_b200 as p
    p.load()
    return p


def _ndev():
    import piquassoboost_b200 as p
    return p.load().pqb200_device_count()


needs2 = pytest.mark.skipif("_ndev() < 2", reason="needs two GPUs")


def test_small_automatic_calls_use_one_device(pqb):
    """ndev=0 is automatic: a 14x14 permanent launches ONE kernel on ONE device whatever the box has."""
    A = haar_unitary(14, 14)
    pqb.GlynnPermanent(A).calculate()
    st = pqb.last_stats()
    assert st["ndev"] == 1 and st["launches"] == 1
    U = haar_unitary(8, 1)
    pqb.BBFGRepeatedPermanentCalculatorDouble(U, [2, 1, 1, 0, 0, 2, 0, 0], [1, 1, 1, 1, 1, 1, 0, 0]).calculate()
    assert pqb.last_stats()["ndev"] == 1
    sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=3)
    sim.simulate([1, 1, 1, 1, 0, 0, 0, 0], 10)
    assert pqb.last_stats()["ndev"] == 1


def test_explicit_single_device_is_bit_identical_to_automatic(pqb):
    A = haar_unitary(26, 26)
    a = pqb.GlynnPermanentCalculator_wrapper(A, 7, ndev=1).calculate()
    b = pqb.GlynnPermanentCalculator_wrapper(A, 7, ndev=0).calculate()
    assert a == b


@needs2
def test_two_devices_give_the_same_bits(pqb):
    """Every entry point with an `ndev` argument: ndev=2 against ndev=1."""
    from piquassoboost_b200 import _lib
    lib = pqb.load()
    A = np.ascontiguousarray(haar_unitary(24, 24))
    one = pqb.GlynnPermanentCalculator_wrapper(A, 7, ndev=1).calculate()
    assert pqb.last_stats()["ndev"] == 1
    two = pqb.GlynnPermanentCalculator_wrapper(A, 7, ndev=2).calculate()
    assert pqb.last_stats()["ndev"] == 2 and pqb.last_stats()["launches"] == 2
    assert relerr(two, one) < 1e-13     # the chunk length follows the shard count: same sum, other rounding
    big = np.ascontiguousarray(haar_unitary(32, 32))    # at the maximal chunk length the decomposition is the
    assert pqb.GlynnPermanentCalculator_wrapper(big, 7, ndev=2).calculate() == \
        pqb.GlynnPermanentCalculator_wrapper(big, 7, ndev=1).calculate()   # same and so are the bits
    # double-double mode
    B = np.ascontiguousarray(haar_unitary(16, 16))
    assert pqb.permanent_double_double(B, ndev=2)[1][0::2].tolist() == pqb.permanent_double_double(B, ndev=1)[1][0::2].tolist()
    # repeated: one big permanent, and the batch
    U = np.ascontiguousarray(haar_unitary(12, 5))
    inp, out = [3, 2, 2, 1, 1, 1, 1, 1, 0, 0, 0, 0], [2, 2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0]
    r1 = pqb.ChinHuhPermanentCalculator_wrapper(7, U, np.array(inp), np.array(out), ndev=1).calculate()
    r2 = pqb.ChinHuhPermanentCalculator_wrapper(7, U, np.array(inp), np.array(out), ndev=2).calculate()
    assert pqb.last_stats()["ndev"] == 2
    assert relerr(r2, r1) < 1e-13      # the chunk length follows the shard count: same sum, other rounding
    rng = np.random.default_rng(0)
    ins = [rng.multinomial(6, [1 / 12] * 12) for _ in range(40)]
    outs = [rng.multinomial(6, [1 / 12] * 12) for _ in range(30)]
    b1 = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs, ndev=1).calculate()
    b2 = pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, outs, ndev=2).calculate()
    assert b1 == b2
    # pmf layer and the sampler (shots shard over the devices, no collective)
    samples = np.array([rng.multinomial(5, [1 / 12] * 12) for _ in range(200)], dtype=np.int64)
    inputs = np.array([rng.multinomial(6, [1 / 12] * 12) for _ in range(200)], dtype=np.int64)
    p1 = pqb.compute_pmf(U, samples, inputs, ndev=1)
    p2 = pqb.compute_pmf(U, samples, inputs, ndev=2)
    assert np.allclose(p1, p2, rtol=1e-11, atol=1e-300)   # cluster size follows the shots per device
    state = np.array([1] * 8 + [0] * 4, dtype=np.int64)
    s1 = np.array(pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=11, ndev=1).simulate(state, 500))
    s2 = np.array(pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=11, ndev=2).simulate(state, 500))
    assert pqb.last_stats()["ndev"] == 2
    assert np.array_equal(s1, s2)


def test_concurrent_callers(pqb, ref):
    """Several host threads call the library at once (the reference's TBB workers do,
    samplingHelperFunctions.cpp:137): every result must equal the single-threaded one."""
    mats = [np.ascontiguousarray(haar_unitary(n, 700 + n)) for n in (12, 16, 18, 20, 22, 14, 17, 19)]
    want = [pqb.GlynnPermanent(m).calculate() for m in mats]
    U = np.ascontiguousarray(haar_unitary(8, 2))
    st_in, st_out = np.array([2, 1, 1, 0, 0, 2, 0, 0]), np.array([1, 1, 1, 1, 1, 1, 0, 0])
    want_rep = pqb.BBFGRepeatedPermanentCalculatorDouble(U, st_in, st_out).calculate()
    errors = []

    def worker(k):
        try:
            for rep in range(20):
                i = (k + rep) % len(mats)
                if pqb.GlynnPermanent(mats[i]).calculate() != want[i]:
                    errors.append(("plain", k, rep))
                if pqb.BBFGRepeatedPermanentCalculatorDouble(U, st_in, st_out).calculate() != want_rep:
                    errors.append(("repeated", k, rep))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_async_shard_on_the_default_stream(pqb):
    """pqb200_permanent_shard with d_out and a NULL stream must run on the caller's
    legacy default stream, so that work the caller enqueues there afterwards (a collective, a copy)
    sees the finished partial.  world_size 1 through torch.distributed + NCCL on the default stream."""
    import os
    import torch
    import torch.distributed as dist
    from piquassoboost_b200.distributed import permanent_distributed
    A = haar_unitary(27, 27)
    want = pqb.GlynnPermanentCalculator_wrapper(A, 7, ndev=1).calculate()
    import socket
    with socket.socket() as sk:                       # a free port: nothing else on the box may hold it
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        for _ in range(5):
            assert permanent_distributed(A) == want          # torch's default stream (handle 0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            assert permanent_distributed(A) == want          # a side stream
        # raw C ABI: NULL stream + d_out, then a copy on the legacy default stream
        from piquassoboost_b200 import _lib
        from piquassoboost_b200.distributed import combine_partials, gpu_shard_fn
        part = torch.zeros(4, dtype=torch.float64, device="cuda:0")
        for _ in range(5):
            part.zero_()
            gpu_shard_fn(A, 0, 1, 0, d_out=part.data_ptr(), stream=None)
            assert combine_partials(part.cpu().numpy(), 27) == want
    finally:
        dist.destroy_process_group()


def test_dd_edge_cases_follow_the_reference(pqb, ref):
    """the double-double entry point keeps the reference's trivial branches."""
    dd = pqb.permanent_double_double
    assert dd(np.zeros((0, 0), dtype=complex))[0] == 1.0
    row = haar_unitary(5, 1)[:1]
    assert relerr(dd(row)[0], np.prod(row)) < 1e-15
    tall = haar_unitary(6, 2)[:, :3]                        # rows >= cols + 2 -> 0 (BBFGPermanentCalculator.hpp:90-95)
    assert dd(tall)[0] == 0.0 and pqb.GlynnPermanent(tall).calculate() == 0.0
    big_tall = np.ones((60, 3), dtype=complex)
    assert pqb.GlynnPermanent(big_tall).calculate() == 0.0   # also beyond the size limit of the kernels
    wide = haar_unitary(7, 3)[:4, :]                        # rows < cols: the reference's rectangular sum
    assert relerr(dd(wide)[0], ref.bbfg(np.ascontiguousarray(wide), True)) < 1e-14
    assert relerr(pqb.GlynnPermanent(wide).calculate(), ref.bbfg(np.ascontiguousarray(wide), True)) < 1e-12
    one_more = haar_unitary(7, 4)[:5, :4]                   # rows == cols + 1
    assert relerr(dd(one_more)[0], ref.bbfg(np.ascontiguousarray(one_more), True)) < 1e-14
    # GlynnPermanentCalculatorSimple: non-square input gives 0 (GlynnPermanentCalculatorSimple.hpp:67-70)
    assert pqb.GlynnPermanentSimpleDouble(wide).calculate() == 0
    sq = haar_unitary(6, 9)
    assert relerr(pqb.GlynnPermanentSimpleLongDouble(sq).calculate(), ref.bbfg(sq, True)) < 1e-12
