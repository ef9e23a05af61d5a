"""CPU tests of the drop-in boundary: libpqb200.so loads and exports every symbol that
include/pqb200.h declares; with no CUDA device every compute call fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pqb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pqb200_[A-Za-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from piquassoboost_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations found in include/pqb200.h"
    assert sorted(_lib.exported_symbols()) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_strings():
    from piquassoboost_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.pqb200_version()
    assert isinstance(lib.pqb200_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    """In the dev container there is no GPU: the product must raise, never compute on the CPU."""
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        has_driver = True
    except OSError:
        has_driver = False
    import piquassoboost_b200 as pqb
    if has_driver and pqb.load().pqb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pqb.Pqb200Error):
        pqb.GlynnPermanent(np.eye(3)).calculate()
    with pytest.raises(pqb.Pqb200Error):
        pqb.BBFGRepeatedPermanentCalculatorDouble(np.eye(3), [1, 1, 1], [1, 1, 1]).calculate()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "piquassoboost_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


def test_argument_validation_happens_before_any_launch():
    from piquassoboost_b200 import _lib
    lib = _lib.load()
    out = np.zeros(2)
    a = np.ones((3, 3), dtype=np.complex128)
    assert lib.pqb200_permanent(None, 3, 3, 3, 1, _lib.dptr(out)) == 2          # PQB200_ERR_ARG
    assert lib.pqb200_permanent(_lib.dptr(a), 3, 3, 2, 1, _lib.dptr(out)) == 2  # stride < cols
    big = np.ones((46, 46), dtype=np.complex128)
    assert lib.pqb200_permanent(_lib.dptr(big), 46, 46, 46, 1, _lib.dptr(out)) == 2
    parts = np.array([1.0, 0.0, 2.0, 0.0, 3.0, 0.0, -1.0, 0.0])
    assert lib.pqb200_permanent_combine(_lib.dptr(parts), 2, 3, _lib.dptr(out)) == 0
    assert out.tolist() == [1.0, 0.25]
