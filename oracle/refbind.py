"""oracle/refbind.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

ctypes binding of oracle/_ref/libpqbref.so: the UNMODIFIED reference C++ calculators compiled from
/root/reference by oracle/Makefile (`make ref`), driven through oracle/ref_driver/ref_capi.cpp.
Used by tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libpqbref.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)
_lib = None


def available():
    return os.path.exists(REF_PATH)


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(REF_PATH)
        lib.ref_last_error.restype = ctypes.c_char_p
        lib.ref_hardware_concurrency.restype = ctypes.c_int
        for name in ("ref_bbfg", "ref_glynn", "ref_glynn_simple"):
            f = getattr(lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [_dp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, _dp]
        for name in ("ref_bbfg_repeated", "ref_glynn_repeated"):
            f = getattr(lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [_dp, ctypes.c_size_t, ctypes.c_size_t, _ip, _ip, ctypes.c_int, _dp]
        lib.ref_chinhuh.restype = ctypes.c_int
        lib.ref_chinhuh.argtypes = [_dp, ctypes.c_size_t, ctypes.c_size_t, _ip, _ip, _dp]
        lib.ref_compute_pmf.restype = ctypes.c_int
        lib.ref_compute_pmf.argtypes = [_dp, ctypes.c_size_t, _ip, _ip, _dp]
        lib.ref_simulate_B.restype = ctypes.c_int
        lib.ref_simulate_B.argtypes = [_dp, ctypes.c_size_t, ctypes.c_uint64, _ip, ctypes.c_int, _ip]
        lib.ref_simulate_B_two_calls.restype = ctypes.c_int
        lib.ref_simulate_B_two_calls.argtypes = [_dp, ctypes.c_size_t, ctypes.c_uint64, _ip, ctypes.c_int,
                                                 ctypes.c_int, _ip]
        lib.ref_simulate_B_uniform_losses.restype = ctypes.c_int
        lib.ref_simulate_B_uniform_losses.argtypes = [_dp, ctypes.c_size_t, ctypes.c_uint64, _ip, ctypes.c_int,
                                                      ctypes.c_double, _ip]
        lib.ref_simulate_B_lossy.restype = ctypes.c_int
        lib.ref_simulate_B_lossy.argtypes = [_dp, ctypes.c_size_t, ctypes.c_uint64, _ip, ctypes.c_int,
                                             ctypes.c_int, _ip]
        lib.ref_adapt_interferometer.restype = ctypes.c_int
        lib.ref_adapt_interferometer.argtypes = [_dp, ctypes.c_size_t, _ip, _ip, _dp,
                                                 ctypes.POINTER(ctypes.c_size_t),
                                                 ctypes.POINTER(ctypes.c_size_t)]
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(load().ref_last_error().decode())


def _m(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(_dp)


def _s(s):
    s = np.ascontiguousarray(s, dtype=np.int64)
    return s, s.ctypes.data_as(_ip)


def hardware_concurrency():
    return load().ref_hardware_concurrency()


def _plain(fn, a, long_double):
    a, pa = _m(a)
    out = np.zeros(2)
    _check(fn(pa, a.shape[0], a.shape[1], 1 if long_double else 0, out.ctypes.data_as(_dp)))
    return complex(out[0], out[1])


def bbfg(a, long_double=True):
    """BBFGPermanentCalculator::calculate(A, use_extended)."""
    return _plain(load().ref_bbfg, a, long_double)


def glynn(a, long_double=True):
    """GlynnPermanentCalculator{LongDouble,Double}::calculate(A)."""
    return _plain(load().ref_glynn, a, long_double)


def glynn_simple(a, long_double=True):
    return _plain(load().ref_glynn_simple, a, long_double)


def _rep(fn, a, input_state, output_state, long_double):
    a, pa = _m(a)
    i, pi = _s(input_state)
    o, po = _s(output_state)
    out = np.zeros(2)
    _check(fn(pa, a.shape[0], a.shape[1], pi, po, 1 if long_double else 0, out.ctypes.data_as(_dp)))
    return complex(out[0], out[1])


def bbfg_repeated(a, input_state, output_state, long_double=True):
    """BBFGPermanentCalculatorRepeated::calculate(A, col_mult=input, row_mult=output, use_extended)."""
    return _rep(load().ref_bbfg_repeated, a, input_state, output_state, long_double)


def glynn_repeated(a, input_state, output_state, long_double=True):
    return _rep(load().ref_glynn_repeated, a, input_state, output_state, long_double)


def chinhuh(a, input_state, output_state):
    a, pa = _m(a)
    i, pi = _s(input_state)
    o, po = _s(output_state)
    out = np.zeros(2)
    _check(load().ref_chinhuh(pa, a.shape[0], a.shape[1], pi, po, out.ctypes.data_as(_dp)))
    return complex(out[0], out[1])


def compute_pmf(u, sample, current_input):
    u, pu = _m(u)
    s, ps = _s(sample)
    c, pc = _s(current_input)
    pmf = np.zeros(u.shape[0])
    _check(load().ref_compute_pmf(pu, u.shape[0], ps, pc, pmf.ctypes.data_as(_dp)))
    return pmf


def simulate_B(u, seed, input_state, shots):
    """CGeneralizedCliffordsBSimulationStrategy(U,0); seed(seed); simulate(input, shots) -> shots x m."""
    u, pu = _m(u)
    i, pi = _s(input_state)
    out = np.zeros((shots, u.shape[0]), dtype=np.int64)
    _check(load().ref_simulate_B(pu, u.shape[0], seed, pi, shots, out.ctypes.data_as(_ip)))
    return out


def simulate_B_two_calls(u, seed, input_state, shots1, shots2):
    """One strategy object, seed(seed), then simulate(input, shots1) and simulate(input, shots2):
    the global generator keeps running between the calls."""
    u, pu = _m(u)
    i, pi = _s(input_state)
    out = np.zeros((shots1 + shots2, u.shape[0]), dtype=np.int64)
    _check(load().ref_simulate_B_two_calls(pu, u.shape[0], seed, pi, shots1, shots2, out.ctypes.data_as(_ip)))
    return out


def simulate_B_uniform_losses(u, seed, input_state, shots, transmittance):
    """CGeneralizedCliffordsBUniformLossesSimulationStrategy(U, transmittance, 0); seed; simulate."""
    u, pu = _m(u)
    i, pi = _s(input_state)
    out = np.zeros((shots, u.shape[0]), dtype=np.int64)
    _check(load().ref_simulate_B_uniform_losses(pu, u.shape[0], seed, pi, shots, transmittance,
                                                out.ctypes.data_as(_ip)))
    return out


def simulate_B_lossy(u, seed, input_state, shots, approximated_modes):
    """CGeneralizedCliffordsBLossySimulationStrategy(U, approximated_modes, 0); seed; simulate."""
    u, pu = _m(u)
    i, pi = _s(input_state)
    out = np.zeros((shots, u.shape[0]), dtype=np.int64)
    _check(load().ref_simulate_B_lossy(pu, u.shape[0], seed, pi, shots, approximated_modes,
                                       out.ctypes.data_as(_ip)))
    return out
