"""CPU tests around the lossy B strategies: the committed golden shots are what the unmodified
reference produces (oracle/_ref), the reference's own LossyInterferometer KATs are among them, and
the host-side loss extraction of pqb200_simulate_B_lossy (no device involved when shots == 0)
agrees with numpy's SVD."""
import ctypes

import numpy as np
import pytest

import golden_util as G
from conftest import haar_unitary


def test_golden_contains_reference_kats():
    g = G.load("sampling_lossy.json")
    u, d = g["uniform"], g["dilation"]
    assert sorted(map(tuple, u["cpp_generation_order"])) == sorted(map(tuple, u["expected_python_order"]))
    trimmed = [row[:5] for row in d["cpp_generation_order"]]
    assert sorted(map(tuple, trimmed)) == sorted(map(tuple, d["expected_python_order_trimmed"]))
    assert len(u["expected_python_order"]) == 20 and len(d["expected_python_order_trimmed"]) == 20


def test_reference_reproduces_lossy_golden(ref):
    g = G.load("sampling_lossy.json")
    u = g["uniform"]
    got = ref.simulate_B_uniform_losses(G.mat(u["matrix"]), g["seed"], g["input"], g["shots"], u["transmissivity"])
    assert got.tolist() == u["cpp_generation_order"]
    for r in g["runs"]:
        assert ref.simulate_B_uniform_losses(G.mat(r["unitary"]), r["seed"], r["input"], r["shots"],
                                             r["uniform_losses"]["transmissivity"]).tolist() == r["uniform_losses"]["samples"]
        assert ref.simulate_B_lossy(G.mat(r["lossy_matrix"]), r["seed"], r["input"], r["shots"], 0).tolist() == r["lossy_0"]
        assert ref.simulate_B_lossy(G.mat(r["lossy_matrix"]), r["seed"], r["input"], r["shots"], 2).tolist() == r["lossy_2"]
        assert ref.simulate_B_two_calls(G.mat(r["unitary"]), r["seed"], r["input"], 7, r["shots"] - 7).tolist() == r["two_calls"]


@pytest.mark.parametrize("m", [1, 2, 5, 12, 40])
def test_loss_extraction_matches_numpy_svd(m):
    """extract_losses_from_interferometer (CGeneralizedCliffordsBLossySimulationStrategy.cpp:386-452):
    uniform_loss = 1 - sigma_max^2 and U <- U / sigma_max."""
    from piquassoboost_b200 import _lib
    rng = np.random.default_rng(m)
    M = haar_unitary(m, 70 + m) * rng.uniform(0.3, 0.97, size=m)[:, None]
    M = np.ascontiguousarray(M @ haar_unitary(m, 170 + m))
    work = M.copy()
    state = np.zeros(m, dtype=np.int64)
    loss = ctypes.c_double(-1.0)
    _lib.check(_lib.load().pqb200_simulate_B_lossy(_lib.dptr(work), m, m, _lib.iptr(state), 0, 0, 1,
                                                   ctypes.byref(loss), None))
    s = np.linalg.svd(M, compute_uv=False)
    assert abs(loss.value - (1.0 - s[0] ** 2)) < 1e-13
    assert np.allclose(work, M / s[0], rtol=1e-13, atol=1e-16)
    assert abs(np.linalg.svd(work, compute_uv=False)[0] - 1.0) < 1e-13


def test_sampler_argument_errors_need_no_device():
    from piquassoboost_b200 import _lib, Pqb200Error
    lib = _lib.load()
    U = haar_unitary(4, 1)
    out = np.zeros((1, 4), dtype=np.int64)
    neg = np.array([1, -1, 0, 0], dtype=np.int64)
    with pytest.raises(Pqb200Error, match="negative"):
        _lib.check(lib.pqb200_simulate_B_next(_lib.dptr(U), 4, 4, _lib.iptr(neg), 1, 1, _lib.iptr(out)))
    big = np.array([41, 0, 0, 0], dtype=np.int64)
    with pytest.raises(Pqb200Error, match="40 photons"):
        _lib.check(lib.pqb200_simulate_B_uniform_losses(_lib.dptr(U), 4, 4, _lib.iptr(big), 1, 0.9, 1, _lib.iptr(out)))
    with pytest.raises(Pqb200Error, match="stride"):
        _lib.check(lib.pqb200_simulate_B_next(_lib.dptr(U), 4, 3, _lib.iptr(big), 1, 1, _lib.iptr(out)))
    assert lib.pqb200_seed(5) == 0
