"""Python mirror of the reference's Clifford-Clifford "B" sampling strategies (lossless, uniform
losses, lossy), backed by the CUDA engine (pqb200_simulate_B* / pqb200_pmf_layer).

Same class names, constructor arguments and return types as
piquassoboost/sampling/simulation_strategies/GeneralizedCliffordsSimulationStrategy.py and the
CPython type in .../GeneralizedCliffordsBSimulationStrategy_wrapper.cpp:124-260:
`(interferometer_matrix, seed, lib)`, `.seed(K)`, `.simulate(input_state, n)` -> tuple of int64
arrays in the C++ generation order.
"""
import time

import numpy as np

from . import _lib
from ._lib import Pqb200Error


def compute_pmf(interferometer_matrix, sample, current_input, ndev=0, batched=False):
    """pic::compute_pmf (samplingHelperFunctions.cpp:78-342) for one shot, or for S shots at once
    when `sample` / `current_input` are S x m arrays.  batched=True evaluates one kernel job per
    sub-permanent (pqb200_pmf_layer_batched) instead of the single Laplace sweep."""
    u = _lib.as_matrix(interferometer_matrix)
    m = u.shape[0]
    s = np.ascontiguousarray(sample, dtype=np.int64).reshape(-1, m)
    c = np.ascontiguousarray(current_input, dtype=np.int64).reshape(-1, m)
    if s.shape != c.shape:
        raise Pqb200Error("sample and current_input must have the same shape")
    pmf = np.zeros(s.shape, dtype=np.float64)
    fn = _lib.load().pqb200_pmf_layer_batched if batched else _lib.load().pqb200_pmf_layer
    _lib.check(fn(_lib.dptr(u), m, u.strides[0] // 16, _lib.iptr(s), _lib.iptr(c), s.shape[0], ndev,
                  _lib.dptr(pmf)))
    return pmf[0] if np.ndim(sample) == 1 else pmf


class _BStrategyBase:
    """Shared behaviour of the three B strategy wrappers: `(interferometer_matrix, <param>, seed,
    lib)`, `.seed(K)`, `.simulate(input_state, n)` -> tuple of int64 arrays.

    The reference keeps ONE process-global std::mt19937 (samplingHelperFunctions.cpp:58-62): every
    constructor seeds it (with time(NULL) unless a seed is given) and every simulate() continues it.
    The library owns the same global generator (pqb200_seed)."""

    def __init__(self, interferometer_matrix, seed=None, lib=0, ndev=0):
        # own copy: the lossy strategy updates its matrix, and the reference also stores a copy of
        # the matrix header (Update_interferometer_matrix)
        self.interferometer_matrix = np.array(_lib.as_matrix(interferometer_matrix))
        if self.interferometer_matrix.shape[0] != self.interferometer_matrix.shape[1]:
            raise Pqb200Error("interferometer matrix must be square")
        self.lib = int(lib)
        self.ndev = int(ndev)
        self.seed(int(time.time()) if seed is None else int(seed))

    def seed(self, value=0):
        """C...SimulationStrategy::seed -> seed_random_generator."""
        _lib.check(_lib.load().pqb200_seed(int(value) & 0xFFFFFFFFFFFFFFFF))

    def _state(self, input_state):
        state = _lib.as_state(input_state)
        if len(state) != self.interferometer_matrix.shape[0]:
            raise Pqb200Error("input state length does not match the interferometer")
        return state

    def _run(self, u, m, state, n, out):
        raise NotImplementedError

    def simulate(self, input_state, samples_number=1):
        u = self.interferometer_matrix
        m = u.shape[0]
        state = self._state(input_state)
        n = max(int(samples_number), 0)
        out = np.zeros((n, m), dtype=np.int64)
        self._run(u, m, state, n, out)
        # one (shots, m) array; the tuple holds its rows as views (20 000 per-shot copies cost more
        # than the kernel)
        return tuple(out)


class GeneralizedCliffordsBSimulationStrategy_wrapper(_BStrategyBase):
    """GeneralizedCliffordsBSimulationStrategy_wrapper.cpp:124-260."""

    def _run(self, u, m, state, n, out):
        _lib.check(_lib.load().pqb200_simulate_B_next(_lib.dptr(u), m, u.strides[0] // 16,
                                                      _lib.iptr(state), n, self.ndev, _lib.iptr(out)))


class GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper(_BStrategyBase):
    """GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper.cpp:132-180:
    `(interferometer_matrix, transmissivity, seed, lib)`; the value is the amplitude transmittance
    (the survival probability of a photon is its square, ...UniformLosses...cpp:309)."""

    def __init__(self, interferometer_matrix, transmissivity=1.0, seed=None, lib=0, ndev=0):
        super().__init__(interferometer_matrix, seed=seed, lib=lib, ndev=ndev)
        self.transmissivity = float(transmissivity)

    def _run(self, u, m, state, n, out):
        _lib.check(_lib.load().pqb200_simulate_B_uniform_losses(
            _lib.dptr(u), m, u.strides[0] // 16, _lib.iptr(state), n, self.transmissivity, self.ndev,
            _lib.iptr(out)))


class GeneralizedCliffordsBLossySimulationStrategy_wrapper(_BStrategyBase):
    """GeneralizedCliffordsBLossySimulationStrategy_wrapper.cpp:132-180:
    `(interferometer_matrix, approximated_modes, seed, lib)`.  As in the reference, simulate()
    extracts the uniform losses from the stored matrix and KEEPS the rescaled matrix
    (CGeneralizedCliffordsBLossySimulationStrategy.cpp:386-452)."""

    def __init__(self, interferometer_matrix, approximated_modes=0, seed=None, lib=0, ndev=0):
        super().__init__(interferometer_matrix, seed=seed, lib=lib, ndev=ndev)
        self.approximated_modes = int(approximated_modes)
        self.uniform_loss = None

    def set_approximated_modes_number(self, value):
        self.approximated_modes = int(value)

    def _run(self, u, m, state, n, out):
        import ctypes
        loss = ctypes.c_double(0.0)
        _lib.check(_lib.load().pqb200_simulate_B_lossy(
            _lib.dptr(u), m, u.strides[0] // 16, _lib.iptr(state), n, self.approximated_modes,
            self.ndev, ctypes.byref(loss), _lib.iptr(out)))
        self.uniform_loss = loss.value


class GeneralizedCliffordsBSimulationStrategy(GeneralizedCliffordsBSimulationStrategy_wrapper):
    """GeneralizedCliffordsSimulationStrategy.py: `(interferometer_matrix, seed)` with lib=0."""

    def __init__(self, interferometer_matrix, seed=None, ndev=0):
        super().__init__(interferometer_matrix, seed=seed, lib=0, ndev=ndev)


class GeneralizedCliffordsBUniformLossesSimulationStrategy(
        GeneralizedCliffordsBUniformLossesSimulationStrategy_wrapper):
    """GeneralizedCliffordsSimulationStrategy.py: `(interferometer_matrix, transmissivity, seed)`."""

    def __init__(self, interferometer_matrix, transmissivity, seed=None, ndev=0):
        super().__init__(interferometer_matrix, transmissivity=transmissivity, seed=seed, lib=0, ndev=ndev)


class GeneralizedCliffordsBLossySimulationStrategy(GeneralizedCliffordsBLossySimulationStrategy_wrapper):
    """GeneralizedCliffordsSimulationStrategy.py: `(interferometer_matrix, approximated_modes, seed)`."""

    def __init__(self, interferometer_matrix, approximated_modes, seed=None, ndev=0):
        super().__init__(interferometer_matrix, approximated_modes=approximated_modes, seed=seed, lib=0,
                         ndev=ndev)
