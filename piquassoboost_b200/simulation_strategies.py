"""Python mirror of the reference's Clifford-Clifford "B" sampling strategy, backed by the CUDA
engine (pqb200_simulate_B / pqb200_pmf_layer).

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


def compute_pmf(interferometer_matrix, sample, current_input, ndev=0):
    """pic::compute_pmf (samplingHelperFunctions.cpp:78-342) for one shot, or for S shots at once
    when `sample` / `current_input` are S x m arrays."""
    u = _lib.as_matrix(interferometer_matrix)
    m = u.shape[0]
    s = np.ascontiguousarray(sample, dtype=np.int64).reshape(-1, m)
    c = np.ascontiguousarray(current_input, dtype=np.int64).reshape(-1, m)
    if s.shape != c.shape:
        raise Pqb200Error("sample and current_input must have the same shape")
    pmf = np.zeros(s.shape, dtype=np.float64)
    _lib.check(_lib.load().pqb200_pmf_layer(_lib.dptr(u), m, u.strides[0] // 16, _lib.iptr(s),
                                            _lib.iptr(c), s.shape[0], ndev, _lib.dptr(pmf)))
    return pmf[0] if np.ndim(sample) == 1 else pmf


class GeneralizedCliffordsBSimulationStrategy_wrapper:
    def __init__(self, interferometer_matrix, seed=None, lib=0, ndev=0):
        self.interferometer_matrix = _lib.as_matrix(interferometer_matrix)
        if self.interferometer_matrix.shape[0] != self.interferometer_matrix.shape[1]:
            raise Pqb200Error("interferometer matrix must be square")
        self.lib = int(lib)
        self.ndev = int(ndev)
        # the reference constructor seeds with time(NULL) (CGeneralizedCliffordsBSimulationStrategy.cpp:81-86)
        self._seed = int(time.time()) if seed is None else int(seed)
        self._calls = 0

    def seed(self, value):
        """CGeneralizedCliffordsBSimulationStrategy::seed (cpp:113-116)."""
        self._seed = int(value)
        self._calls = 0

    def simulate(self, input_state, samples_number=1):
        u = self.interferometer_matrix
        m = u.shape[0]
        state = _lib.as_state(input_state)
        if len(state) != m:
            raise Pqb200Error("input state length does not match the interferometer")
        n = int(samples_number)
        out = np.zeros((max(n, 0), m), dtype=np.int64)
        if self._calls:
            # the reference keeps ONE global generator running across simulate() calls; here every
            # call restarts from the seed, so a second call on the same object is made explicit
            raise Pqb200Error("call seed() before simulating again: the generator state is per call")
        _lib.check(_lib.load().pqb200_simulate_B(_lib.dptr(u), m, u.strides[0] // 16, _lib.iptr(state),
                                                 n, self._seed & 0xFFFFFFFFFFFFFFFF, self.ndev,
                                                 _lib.iptr(out)))
        self._calls += 1
        return tuple(out[k].copy() for k in range(n))


class GeneralizedCliffordsBSimulationStrategy(GeneralizedCliffordsBSimulationStrategy_wrapper):
    """GeneralizedCliffordsSimulationStrategy.py: `(interferometer_matrix, seed)` with lib=0."""

    def __init__(self, interferometer_matrix, seed=None, ndev=0):
        super().__init__(interferometer_matrix, seed=seed, lib=0, ndev=ndev)
