"""Python mirror of the reference's permanent-calculator classes, backed by the CUDA engine.

Same class names, constructor arguments, `calculate()` behaviour, `lib` integers and attribute
names (`matrix`, `input_state`, `output_state`) as
piquassoboost/sampling/Boson_Sampling_Utilities.py:30-335 and the CPython types it derives from
(sampling/GlynnPermanentCalculator_Wrapper.hpp:34-44,132-378;
sampling/ChinHuhPermanentCalculator_Wrapper.hpp:36-47,143-408), so that

    from piquassoboost_b200.Boson_Sampling_Utilities import GlynnPermanent
    GlynnPermanent(A).calculate()

is a drop-in for the reference import.  Every supported `lib` id is routed to the one numeric mode
of the engine (FP64 products, double-double accumulation): the ids keep
their meaning only as far as "which reference calculator would have been used".
"""
import ctypes
import sys

import numpy as np

from . import _lib
from ._lib import Pqb200Error

# lib ids of GlynnPermanentCalculator_wrapper (GlynnPermanentCalculator_Wrapper.hpp:34-44)
GlynnRep = 0  # GlynnPermanentCalculatorLongDouble
GlynnInf = 1
GlynnSingleDFE = 2
GlynnDualDFE = 3
GlynnSingleDFEF = 4
GlynnDualDFEF = 5
GlynnDouble = 6
BBFGDouble = 7
BBFGLongDouble = 8
GlynnSimpleDouble = 10
GlynnSimpleLongDouble = 11
GlynnB200 = 12  # new id: explicit request for the CUDA engine

# lib ids of ChinHuhPermanentCalculator_wrapper (ChinHuhPermanentCalculator_Wrapper.hpp:36-47)
ChinHuh = 0
GlynnRepLongDouble = 1
GlynnRepSingleDFE = 2
GlynnRepDualDFE = 3
GlynnRepMultiSingleDFE = 4
GlynnRepMultiDualDFE = 5
GlynnRepDouble = 6
BBFGRepDouble = 7
BBFGRepLongDouble = 8
GlynnRepInf = 9
GlynnRepB200 = 12

_PLAIN_LIBS = {GlynnRep, GlynnDouble, BBFGDouble, BBFGLongDouble, GlynnSimpleDouble,
               GlynnSimpleLongDouble, GlynnB200}
_REPEATED_LIBS = {ChinHuh, GlynnRepLongDouble, GlynnRepDouble, BBFGRepDouble, BBFGRepLongDouble,
                  GlynnRepB200}

#: number of GPUs one calculate() call is sharded over (0 = automatic: as many devices as the call's
#: work is worth -- small permanents run on one GPU, see include/pqb200.h)
DEFAULT_NDEV = 0


def _plain_permanent(matrix, ndev):
    m = _lib.as_matrix(matrix)
    out = np.zeros(2, dtype=np.float64)
    _lib.check(_lib.load().pqb200_permanent(_lib.dptr(m), m.shape[0], m.shape[1],
                                            m.strides[0] // 16 if m.shape[1] else 0, ndev,
                                            _lib.dptr(out)))
    return complex(out[0], out[1])


def _repeated_permanent(matrix, input_state, output_state, ndev):
    m = _lib.as_matrix(matrix)
    cols = _lib.as_state(input_state)   # input state  <-> column multiplicities
    rows = _lib.as_state(output_state)  # output state <-> row multiplicities
    if len(cols) != m.shape[1] or len(rows) != m.shape[0]:
        raise Pqb200Error("state lengths do not match the matrix shape")
    out = np.zeros(2, dtype=np.float64)
    _lib.check(_lib.load().pqb200_permanent_repeated(
        _lib.dptr(m), m.shape[0], m.shape[1], m.strides[0] // 16 if m.shape[1] else 0,
        _lib.iptr(rows), _lib.iptr(cols), ndev, _lib.dptr(out)))
    return complex(out[0], out[1])


class GlynnPermanentCalculator_wrapper:
    """Mirror of the CPython type of the same name: `(matrix, lib)`; `matrix` may be a list of
    arrays, then `calculate()` returns a list (GlynnPermanentCalculator_Wrapper.hpp:204-293)."""

    def __init__(self, matrix=None, lib=BBFGDouble, ndev=None):
        self.lib = int(lib)
        self.ndev = DEFAULT_NDEV if ndev is None else int(ndev)
        self.matrix = matrix

    @property
    def matrix(self):
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        if isinstance(value, (list, tuple)):
            self._matrix = [_lib.as_matrix(v) for v in value]
        elif value is None:
            self._matrix = None
        else:
            self._matrix = _lib.as_matrix(value)

    def calculate(self):
        if self.lib not in _PLAIN_LIBS:
            if self.lib in (GlynnSingleDFE, GlynnDualDFE, GlynnSingleDFEF, GlynnDualDFEF):
                raise Pqb200Error("DFE (FPGA) permanent calculators are not part of this build")
            if self.lib == GlynnInf:
                # the reference's "Inf" calculators need an MPFR build (BBFGPermanentCalculator.cpp:100-107);
                # here the request is served by the extended mode: every operation in double-double
                # (~106 bits, pqb200_permanent_dd), for square matrices
                if self._matrix is None:
                    raise Pqb200Error("matrix is not set")
                mats = self._matrix if isinstance(self._matrix, list) else [self._matrix]
                vals = [permanent_double_double(m, self.ndev)[0] for m in mats]
                return vals if isinstance(self._matrix, list) else vals[0]
            raise Pqb200Error("Wrong value set for permanent library.")
        if self._matrix is None:
            raise Pqb200Error("matrix is not set")
        if self.lib in (GlynnSimpleDouble, GlynnSimpleLongDouble):
            # GlynnPermanentCalculatorSimple only takes square matrices: it prints the message below and
            # returns 0 otherwise (GlynnPermanentCalculatorSimple.hpp:67-70)
            mats = self._matrix if isinstance(self._matrix, list) else [self._matrix]
            if any(m.shape[0] != m.shape[1] for m in mats):
                vals = []
                for m in mats:
                    if m.shape[0] != m.shape[1]:
                        print("Error: GlynnPermanentCalculatorSimple only works for square matrices",
                              file=sys.stderr)
                        vals.append(0j)
                    else:
                        vals.append(_plain_permanent(m, self.ndev))
                return vals if isinstance(self._matrix, list) else vals[0]
        if isinstance(self._matrix, list):
            return self._calculate_list(self._matrix)
        return _plain_permanent(self._matrix, self.ndev)

    def _calculate_list(self, mats):
        lib = _lib.load()
        n = len(mats)
        if n == 0:
            return []
        ptrs = (_lib.c_double_p * n)(*[_lib.dptr(m) for m in mats])
        rows = (ctypes.c_size_t * n)(*[m.shape[0] for m in mats])
        cols = (ctypes.c_size_t * n)(*[m.shape[1] for m in mats])
        strides = (ctypes.c_size_t * n)(*[m.strides[0] // 16 if m.shape[1] else 0 for m in mats])
        out = np.zeros(2 * n, dtype=np.float64)
        _lib.check(lib.pqb200_permanent_batch(ptrs, rows, cols, strides, n, self.ndev,
                                              _lib.dptr(out)))
        return out.view(np.complex128).tolist()


class ChinHuhPermanentCalculator_wrapper:
    """Mirror of the CPython type of the same name: `(lib, matrix, input_state, output_state)`.
    List-valued states follow the reference (ChinHuhPermanentCalculator_Wrapper.hpp:233-326): a list
    of n input vectors with a list of n (ragged) lists of output vectors -- or the other way round --
    gives ret[i][j]; two flat lists give the full cross product (extension, see `_pairs`).
    lib 0 (ChinHuh: the reference's Ryser-type CChinHuhPermanentCalculator) is served by the same
    Glynn/BBFG engine -- the value is the same permanent, only the reference's algorithm differs."""

    def __init__(self, lib=BBFGRepDouble, matrix=None, input_state=None, output_state=None,
                 ndev=None):
        self.lib = int(lib)
        self.ndev = DEFAULT_NDEV if ndev is None else int(ndev)
        self.matrix = None if matrix is None else _lib.as_matrix(matrix)
        self.input_state = self._norm_state(input_state)
        self.output_state = self._norm_state(output_state)

    @staticmethod
    def _norm_state(state):
        """None, one occupation vector, a flat list of vectors, or a list of lists of vectors (the
        reference's per-input / per-output form, ChinHuhPermanentCalculator_Wrapper.hpp:233-263)."""
        if state is None:
            return None
        if isinstance(state, (list, tuple)):
            if len(state) == 0:
                return []
            if all(isinstance(s, np.ndarray) for s in state):
                return [_lib.as_state(s) for s in state]                        # flat list of vectors
            if all(isinstance(inner, (list, tuple)) and all(isinstance(s, np.ndarray) for s in inner)
                   for inner in state):
                return [[_lib.as_state(s) for s in inner] for inner in state]   # list of (ragged) lists
        return _lib.as_state(state)

    @staticmethod
    def _is_nested(state):
        return isinstance(state, list) and len(state) > 0 and isinstance(state[0], list)

    def _pairs(self):
        """(input, output) occupation pairs of a list-valued call and the shape of the result.

        Reference forms (ChinHuhPermanentCalculator_Wrapper.hpp:233-326): `input_state` a list of n
        vectors and `output_state` a list of n lists of vectors -> ret[i][j] = perm(input[i],
        output[i][j]); or the other way round ("multiInput": `input_state` a list of lists, one per
        output vector).  Extension of this package: two FLAT lists -> the full cross product
        ret[i][o] (what piquasso's batched connector call needs)."""
        ins, outs = self.input_state, self.output_state
        if self._is_nested(ins) and not self._is_nested(outs):
            outs = outs if isinstance(outs, list) else [outs]
            if len(ins) != len(outs):
                raise Pqb200Error("input_state (list of lists) and output_state must have the same length")
            return [(i, outs[k]) for k, inner in enumerate(ins) for i in inner], [len(inner) for inner in ins]
        if self._is_nested(outs):
            ins = ins if isinstance(ins, list) else [ins]
            if len(ins) != len(outs):
                raise Pqb200Error("input_state and output_state (list of lists) must have the same length")
            return [(ins[k], o) for k, inner in enumerate(outs) for o in inner], [len(inner) for inner in outs]
        ins = ins if isinstance(ins, list) else [ins]
        outs = outs if isinstance(outs, list) else [outs]
        return None, (ins, outs)

    def calculate(self):
        if self.lib not in _REPEATED_LIBS:
            if self.lib in (GlynnRepSingleDFE, GlynnRepDualDFE, GlynnRepMultiSingleDFE,
                            GlynnRepMultiDualDFE, 10, 11):
                raise Pqb200Error("DFE (FPGA) permanent calculators are not part of this build")
            if self.lib != GlynnRepInf:
                raise Pqb200Error("Wrong value set for permanent library.")
        ins, outs = self.input_state, self.output_state
        if isinstance(ins, list) or isinstance(outs, list):
            pairs, shape = self._pairs()
            if pairs is None:
                cross_ins, cross_outs = shape
                if self.lib == GlynnRepInf:
                    return [[permanent_repeated_double_double(self.matrix, i, o, self.ndev)[0] for o in cross_outs]
                            for i in cross_ins]
                return self._calculate_lists(cross_ins, cross_outs)
            if self.lib == GlynnRepInf:
                flat = [permanent_repeated_double_double(self.matrix, i, o, self.ndev)[0] for i, o in pairs]
            else:
                flat = self._calculate_pairs(pairs)
            ret, pos = [], 0
            for n in shape:
                ret.append(flat[pos:pos + n])
                pos += n
            return ret
        if self.lib == GlynnRepInf:
            # the reference needs an MPFR build for this id (BBFGPermanentCalculatorRepeated.cpp:117-124);
            # here it is served by the extended mode (every operation in double-double)
            return permanent_repeated_double_double(self.matrix, ins, outs, self.ndev)[0]
        return _repeated_permanent(self.matrix, ins, outs, self.ndev)

    def _calculate_pairs(self, pairs):
        m = self.matrix
        count = len(pairs)
        if count == 0:
            return []
        cols = np.ascontiguousarray(np.array([p[0] for p in pairs], dtype=np.int64).reshape(count, -1))
        rows = np.ascontiguousarray(np.array([p[1] for p in pairs], dtype=np.int64).reshape(count, -1))
        if cols.shape[1] != m.shape[1] or rows.shape[1] != m.shape[0]:
            raise Pqb200Error("state lengths do not match the matrix shape")
        out = np.zeros(2 * count, dtype=np.float64)
        _lib.check(_lib.load().pqb200_permanent_repeated_batch(
            _lib.dptr(m), m.shape[0], m.shape[1], m.strides[0] // 16 if m.shape[1] else 0, _lib.iptr(rows),
            _lib.iptr(cols), count, self.ndev, _lib.dptr(out)))
        return out.view(np.complex128).tolist()

    def _calculate_lists(self, ins, outs):
        m = self.matrix
        count = len(ins) * len(outs)
        if count == 0:
            return [[] for _ in ins]
        # (input i, output o) pairs in row-major order; bulk numpy conversions -- for tens of thousands
        # of tiny permanents the Python list handling, not the GPU, is the cost of this call
        o = np.array(outs, dtype=np.int64).reshape(len(outs), -1)
        i = np.array(ins, dtype=np.int64).reshape(len(ins), -1)
        rows = np.ascontiguousarray(np.tile(o, (len(ins), 1)))
        cols = np.ascontiguousarray(np.repeat(i, len(outs), axis=0))
        out = np.zeros(2 * count, dtype=np.float64)
        _lib.check(_lib.load().pqb200_permanent_repeated_batch(
            _lib.dptr(m), m.shape[0], m.shape[1], m.strides[0] // 16, _lib.iptr(rows),
            _lib.iptr(cols), count, self.ndev, _lib.dptr(out)))
        return out.view(np.complex128).reshape(len(ins), len(outs)).tolist()


class RepeatedPermanentCalculator(ChinHuhPermanentCalculator_wrapper):
    """Boson_Sampling_Utilities.py:30-91."""

    def __init__(self, lib, matrix, input_state, output_state):
        if not isinstance(input_state, np.ndarray):
            if not isinstance(input_state, list):
                input_state = np.array(input_state, dtype=np.int64)
            elif any(not isinstance(item, np.ndarray) for item in input_state):
                if all(np.isscalar(item) for item in input_state):
                    input_state = np.array(input_state, dtype=np.int64)
                else:
                    raise Exception('The input state should be a numpy array, or a list of numpy arrays')
        if not isinstance(output_state, np.ndarray):
            if not isinstance(output_state, list):
                output_state = np.array(output_state, dtype=np.int64)
            elif any(not isinstance(item, np.ndarray) for item in output_state):
                if all(np.isscalar(item) for item in output_state):
                    output_state = np.array(output_state, dtype=np.int64)
                else:
                    raise Exception('The output state should be a numpy array, or a list of numpy arrays')
        super().__init__(lib=lib, matrix=matrix, input_state=input_state, output_state=output_state)


def _repeated_class(name, lib, doc):
    def __init__(self, matrix, input_state, output_state):
        RepeatedPermanentCalculator.__init__(self, lib, matrix, input_state, output_state)
    return type(name, (RepeatedPermanentCalculator,), {"__init__": __init__, "__doc__": doc})


ChinHuhPermanentCalculator = _repeated_class(
    "ChinHuhPermanentCalculator", ChinHuh, "Boson_Sampling_Utilities.py:96-99 (lib 0)")
GlynnRepeatedPermanentCalculator = _repeated_class(
    "GlynnRepeatedPermanentCalculator", GlynnRepLongDouble, "Boson_Sampling_Utilities.py:105-116 (lib 1)")
GlynnRepeatedPermanentCalculatorDouble = _repeated_class(
    "GlynnRepeatedPermanentCalculatorDouble", GlynnRepDouble, "Boson_Sampling_Utilities.py:167-177 (lib 6)")
BBFGRepeatedPermanentCalculatorDouble = _repeated_class(
    "BBFGRepeatedPermanentCalculatorDouble", BBFGRepDouble, "Boson_Sampling_Utilities.py:180-189 (lib 7)")
BBFGRepeatedPermanentCalculatorLongDouble = _repeated_class(
    "BBFGRepeatedPermanentCalculatorLongDouble", BBFGRepLongDouble, "Boson_Sampling_Utilities.py:193-203 (lib 8)")
GlynnRepeatedInfPermanentCalculator = _repeated_class(
    "GlynnRepeatedInfPermanentCalculator", GlynnRepInf, "Boson_Sampling_Utilities.py:120-123 (lib 9; served by the double-double mode instead of MPFR)")


def _plain_class(name, lib, doc):
    def __init__(self, matrix):
        GlynnPermanentCalculator_wrapper.__init__(self, matrix=matrix, lib=lib)
    return type(name, (GlynnPermanentCalculator_wrapper,), {"__init__": __init__, "__doc__": doc})


GlynnPermanent = _plain_class("GlynnPermanent", GlynnRep, "Boson_Sampling_Utilities.py:206-216 (lib 0)")
GlynnPermanentInf = _plain_class("GlynnPermanentInf", GlynnInf, "Boson_Sampling_Utilities.py:218-228 (lib 1; served by the double-double mode instead of MPFR)")
GlynnPermanentDouble = _plain_class("GlynnPermanentDouble", GlynnDouble, "Boson_Sampling_Utilities.py:265-275 (lib 6)")
BBFGPermanentDouble = _plain_class("BBFGPermanentDouble", BBFGDouble, "Boson_Sampling_Utilities.py:282-292 (lib 7)")
BBFGPermanentLongDouble = _plain_class("BBFGPermanentLongDouble", BBFGLongDouble, "Boson_Sampling_Utilities.py:296-307 (lib 8)")
GlynnPermanentSimpleDouble = _plain_class("GlynnPermanentSimpleDouble", GlynnSimpleDouble, "Boson_Sampling_Utilities.py:310-320 (lib 10)")
GlynnPermanentSimpleLongDouble = _plain_class("GlynnPermanentSimpleLongDouble", GlynnSimpleLongDouble, "Boson_Sampling_Utilities.py:324-334 (lib 11)")


def permanent_CPU_repeated_double(matrix, input_state, output_state):
    """Name kept from sampling/permanent_calculators_wrapper.cpp:84-87 (runs on the GPU here)."""
    return _repeated_permanent(matrix, input_state, output_state, DEFAULT_NDEV)


def permanent_CPU_repeated_long_double(matrix, input_state, output_state):
    """sampling/permanent_calculators_wrapper.cpp:157-160."""
    return _repeated_permanent(matrix, input_state, output_state, DEFAULT_NDEV)


def permanent_double_double(matrix, ndev=None):
    """perm(matrix) with every operation in double-double arithmetic (`pqb200_permanent_dd`): the
    engine's extended-precision evaluation (cf. BBFGPermanentCalculator.cpp:109-116, long double in
    the reference).  Returns `(value, parts)`: the value rounded to complex128 and the four doubles
    re.hi, re.lo, im.hi, im.lo."""
    m = _lib.as_matrix(matrix)
    parts = np.zeros(4, dtype=np.float64)
    _lib.check(_lib.load().pqb200_permanent_dd(
        _lib.dptr(m), m.shape[0], m.shape[1], m.strides[0] // 16 if m.shape[1] else 0,
        DEFAULT_NDEV if ndev is None else int(ndev), _lib.dptr(parts)))
    return complex(parts[0] + parts[1], parts[2] + parts[3]), parts


def permanent_repeated_double_double(matrix, input_state, output_state, ndev=None):
    """The repeated-row permanent with every operation in double-double arithmetic
    (`pqb200_permanent_repeated_dd`).  Returns `(value, parts)` like `permanent_double_double`."""
    m = _lib.as_matrix(matrix)
    cols = _lib.as_state(input_state)
    rows = _lib.as_state(output_state)
    if len(cols) != m.shape[1] or len(rows) != m.shape[0]:
        raise Pqb200Error("state lengths do not match the matrix shape")
    parts = np.zeros(4, dtype=np.float64)
    _lib.check(_lib.load().pqb200_permanent_repeated_dd(
        _lib.dptr(m), m.shape[0], m.shape[1], m.strides[0] // 16 if m.shape[1] else 0,
        _lib.iptr(rows), _lib.iptr(cols), DEFAULT_NDEV if ndev is None else int(ndev), _lib.dptr(parts)))
    return complex(parts[0] + parts[1], parts[2] + parts[3]), parts
