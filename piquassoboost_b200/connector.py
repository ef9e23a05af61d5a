"""piquasso connector hook backed by the CUDA engine (reference: piquassoboost/connector.py:23-38).

`cpp_permanent_function(matrix, input, output)` is the callable the reference installs as
`BoostConnector.permanent`; piquasso's Fock-space / sampling calculations call it once per
(input, output) occupation pair (piquasso-module/piquasso/_simulators/sampling/utils.py:136-143).
`permanent_function_batch` is the batched form (one launch for many occupation pairs), which is
the regime where a GPU pays off for tiny permanents.

piquasso itself is not a dependency of this package: `BoostConnector` is defined only when
`piquasso` can be imported, exactly mirroring the reference class.
"""
import numpy as np

from .Boson_Sampling_Utilities import BBFGRepeatedPermanentCalculatorDouble


def cpp_permanent_function(matrix, input, output):  # noqa: A002  (reference argument names)
    result = BBFGRepeatedPermanentCalculatorDouble(matrix, input, output).calculate()
    if isinstance(result, list):
        # the reference maps the list returned for list-valued states to 1.0 (connector.py:27-29)
        result = 1.0
    return result


def permanent_function_batch(matrix, inputs, outputs):
    """Nested list [input][output] of permanents for lists of occupation vectors, evaluated by one
    batched call (ChinHuhPermanentCalculator_Wrapper.hpp:233-326 list branch)."""
    ins = [np.asarray(i, dtype=np.int64) for i in inputs]
    outs = [np.asarray(o, dtype=np.int64) for o in outputs]
    return BBFGRepeatedPermanentCalculatorDouble(matrix, ins, outs).calculate()


try:  # pragma: no cover - piquasso is not part of this image
    from piquasso._simulators.connectors import NumpyConnector
except Exception:  # noqa: BLE001
    NumpyConnector = None

if NumpyConnector is not None:  # pragma: no cover
    class BoostConnector(NumpyConnector):
        def __init__(self) -> None:
            super().__init__()
            self.permanent = cpp_permanent_function
            self.number_of_approximated_modes = None
