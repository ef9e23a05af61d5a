"""oracle/oraclebind.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

ctypes binding of oracle/libpqb_oracle.so, this repository's plain-C restatement of the reference
hot path (oracle/pqb_oracle.c).  Unlike oracle/_ref it can always be rebuilt (gcc only), so it is the
checker used on the GPU box if the prebuilt reference library is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpqb_oracle.so")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)
_lib = None

DOUBLE, LONG_DOUBLE, FLOAT128 = 0, 1, 2


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.pqbo_bbfg.restype = ctypes.c_int
        lib.pqbo_bbfg.argtypes = [_dp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, _dp]
        lib.pqbo_bbfg_repeated.restype = ctypes.c_int
        lib.pqbo_bbfg_repeated.argtypes = [_dp, ctypes.c_size_t, ctypes.c_size_t, _ip, _ip,
                                           ctypes.c_int, _dp]
        lib.pqbo_compute_pmf.restype = ctypes.c_int
        lib.pqbo_compute_pmf.argtypes = [_dp, ctypes.c_size_t, _ip, _ip, ctypes.c_int, _dp]
        lib.pqbo_simulate_B.restype = ctypes.c_int
        lib.pqbo_simulate_B.argtypes = [_dp, ctypes.c_size_t, ctypes.c_uint64, _ip, ctypes.c_int, _ip]
        lib.pqbo_mt19937.restype = None
        lib.pqbo_mt19937.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.c_size_t]
        lib.pqbo_threads.restype = ctypes.c_int
        lib.pqbo_set_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def threads():
    return load().pqbo_threads()


def set_threads(n):
    load().pqbo_set_threads(int(n))


def bbfg(a, prec=LONG_DOUBLE):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    out = np.zeros(2)
    load().pqbo_bbfg(a.ctypes.data_as(_dp), a.shape[0], a.shape[1], prec, out.ctypes.data_as(_dp))
    return complex(out[0], out[1])


def bbfg_repeated(a, input_state, output_state, prec=LONG_DOUBLE):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    cm = np.ascontiguousarray(input_state, dtype=np.int64)
    rm = np.ascontiguousarray(output_state, dtype=np.int64)
    out = np.zeros(2)
    rc = load().pqbo_bbfg_repeated(a.ctypes.data_as(_dp), a.shape[0], a.shape[1],
                                   cm.ctypes.data_as(_ip), rm.ctypes.data_as(_ip), prec,
                                   out.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("Number of input and output states should be equal")
    return complex(out[0], out[1])


def compute_pmf(u, sample, current_input, prec=DOUBLE):
    u = np.ascontiguousarray(u, dtype=np.complex128)
    s = np.ascontiguousarray(sample, dtype=np.int64)
    c = np.ascontiguousarray(current_input, dtype=np.int64)
    pmf = np.zeros(u.shape[0])
    rc = load().pqbo_compute_pmf(u.ctypes.data_as(_dp), u.shape[0], s.ctypes.data_as(_ip),
                                 c.ctypes.data_as(_ip), prec, pmf.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("compute_pmf failed")
    return pmf


def simulate_B(u, seed, input_state, shots):
    u = np.ascontiguousarray(u, dtype=np.complex128)
    i = np.ascontiguousarray(input_state, dtype=np.int64)
    out = np.zeros((shots, u.shape[0]), dtype=np.int64)
    rc = load().pqbo_simulate_B(u.ctypes.data_as(_dp), u.shape[0], seed, i.ctypes.data_as(_ip),
                                shots, out.ctypes.data_as(_ip))
    if rc:
        raise RuntimeError("simulate_B failed")
    return out


def mt19937(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    load().pqbo_mt19937(seed, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n)
    return out
