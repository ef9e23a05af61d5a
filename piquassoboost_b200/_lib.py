"""ctypes binding of libpqb200.so (the C ABI in include/pqb200.h).

This is the reference-side binding a maintainer would otherwise write in C++;
here it plays the role of the CPython extension layer of the reference
(piquassoboost/sampling/GlynnPermanentCalculator_Wrapper.hpp, ChinHuhPermanentCalculator_Wrapper.hpp).
There is no fallback: if the shared library is missing or CUDA is unusable, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PQB200_LIB: alternative build of the library (tuning experiments); default: the in-tree build
_LIB_PATH = os.environ.get("PQB200_LIB") or os.path.join(_HERE, "libpqb200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_size_p = ctypes.POINTER(ctypes.c_size_t)


class Pqb200Stats(ctypes.Structure):
    _fields_ = [
        ("kernel_ms", ctypes.c_double),
        ("gray_steps", ctypes.c_double),
        ("launches", ctypes.c_uint64),
        ("ndev", ctypes.c_int),
    ]


class Pqb200Error(Exception):
    """Raised for every non-zero return code of the C ABI (the reference raises plain Exception
    from `catch (std::string)`; this subclasses it so `except Exception` call sites keep working)."""


_SIGNATURES = {
    "pqb200_init": (ctypes.c_int, [ctypes.c_int]),
    "pqb200_shutdown": (None, []),
    "pqb200_device_count": (ctypes.c_int, []),
    "pqb200_last_error": (ctypes.c_char_p, []),
    "pqb200_version": (ctypes.c_char_p, []),
    "pqb200_permanent": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, c_double_p],
    ),
    "pqb200_permanent_dd": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, c_double_p],
    ),
    "pqb200_nccl_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "pqb200_nccl_init": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "pqb200_nccl_shutdown": (ctypes.c_int, []),
    "pqb200_permanent_nccl": (
        ctypes.c_int, [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, c_double_p]),
    "pqb200_permanent_nccl_async": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "pqb200_permanent_shard": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
         ctypes.c_uint64, ctypes.c_uint64, c_double_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "pqb200_permanent_combine": (
        ctypes.c_int, [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_double_p]),
    "pqb200_permanent_batch": (
        ctypes.c_int,
        [ctypes.POINTER(c_double_p), c_size_p, c_size_p, c_size_p, ctypes.c_size_t, ctypes.c_int,
         c_double_p],
    ),
    "pqb200_permanent_repeated": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, c_int64_p,
         ctypes.c_int, c_double_p],
    ),
    "pqb200_last_repeated_bound": (ctypes.c_int, [c_double_p, ctypes.POINTER(ctypes.c_int)]),
    "pqb200_permanent_repeated_dd": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, c_int64_p,
         ctypes.c_int, c_double_p],
    ),
    "pqb200_permanent_repeated_batch": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, c_int64_p,
         ctypes.c_size_t, ctypes.c_int, c_double_p],
    ),
    "pqb200_pmf_layer": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, c_int64_p, ctypes.c_size_t,
         ctypes.c_int, c_double_p],
    ),
    "pqb200_simulate_B": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, ctypes.c_size_t, ctypes.c_uint64,
         ctypes.c_int, c_int64_p],
    ),
    "pqb200_pmf_layer_batched": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, c_int64_p, ctypes.c_size_t,
         ctypes.c_int, c_double_p],
    ),
    "pqb200_seed": (ctypes.c_int, [ctypes.c_uint64]),
    "pqb200_fill_r_samples": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
         ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8), c_double_p, ctypes.c_int,
         c_int64_p],
    ),
    "pqb200_simulate_B_next": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, ctypes.c_size_t, ctypes.c_int,
         c_int64_p],
    ),
    "pqb200_simulate_B_uniform_losses": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, ctypes.c_size_t, ctypes.c_double,
         ctypes.c_int, c_int64_p],
    ),
    "pqb200_simulate_B_lossy": (
        ctypes.c_int,
        [c_double_p, ctypes.c_size_t, ctypes.c_size_t, c_int64_p, ctypes.c_size_t, ctypes.c_int,
         ctypes.c_int, c_double_p, c_int64_p],
    ),
    "pqb200_last_stats": (ctypes.c_int, [ctypes.POINTER(Pqb200Stats)]),
    "pqb200_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_double, c_double_p]),
}

_lib = None


def lib_path():
    return _LIB_PATH


def exported_symbols():
    """Names of every entry point include/pqb200.h declares."""
    return sorted(_SIGNATURES)


def load():
    """Load libpqb200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise Pqb200Error(
                f"{_LIB_PATH} is missing: build it with `make -C piquassoboost_b200/csrc` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().pqb200_last_error()
        raise Pqb200Error(msg.decode() if msg else f"pqb200 error code {rc}")


def as_matrix(matrix):
    """complex128, C-contiguous 2-D -- what PyArray_FROM_OTF(..., NPY_COMPLEX128, NPY_ARRAY_IN_ARRAY)
    does in the reference wrappers (GlynnPermanentCalculator_Wrapper.hpp:151-157)."""
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    if m.ndim != 2:
        raise Pqb200Error("matrix must be 2-dimensional")
    return m


def as_state(state):
    return np.ascontiguousarray(state, dtype=np.int64).reshape(-1)


def dptr(arr):
    return arr.ctypes.data_as(c_double_p)


def iptr(arr):
    return arr.ctypes.data_as(c_int64_p)


def last_repeated_bound():
    """(a-posteriori rounding bound, re-evaluated in double-double?) of this thread's most recent
    repeated permanent (pqb200_last_repeated_bound)."""
    b = ctypes.c_double(0.0)
    d = ctypes.c_int(0)
    check(load().pqb200_last_repeated_bound(ctypes.byref(b), ctypes.byref(d)))
    return b.value, bool(d.value)


def last_stats():
    st = Pqb200Stats()
    check(load().pqb200_last_stats(ctypes.byref(st)))
    return {"kernel_ms": st.kernel_ms, "gray_steps": st.gray_steps, "launches": int(st.launches),
            "ndev": st.ndev}
