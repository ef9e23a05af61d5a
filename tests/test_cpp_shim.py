"""The C++ calculator shim (include/pqb200_calculators.hpp) compiles against stand-in containers,
links libpqb200.so, and behaves: with a GPU it returns the reference's 2x2 known answer, without one
it throws std::string (exit code 10) -- it never computes on the CPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "shim_smoke")
    libdir = os.path.join(ROOT, "piquassoboost_b200")
    cmd = ["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp"), "-o", exe, "-L" + libdir,
           "-l:libpqb200.so", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_shim_compiles_and_refuses_without_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode in (0, 10), res.stdout + res.stderr
    if res.returncode == 10:
        assert "no usable CUDA device" in res.stdout


@pytest.mark.gpu
def test_shim_known_answer_on_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "repeated=(-1," in res.stdout


REF_TYPES_EXE = os.path.join(ROOT, "oracle", "_ref", "shim_reference_types")


def test_shim_compiles_against_the_reference_container_types():
    """In the dev container (reference tree present) `make -C oracle ref` instantiates the shim with
    pic::matrix / pic::PicState_int64 from the reference's own headers; without a GPU it must throw the
    reference's error type (std::string -> exit code 10)."""
    if not os.path.isdir("/root/reference/piquassoboost"):
        pytest.skip("reference tree not present")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    res = subprocess.run([REF_TYPES_EXE], capture_output=True, text=True)
    assert res.returncode in (0, 10), res.stdout + res.stderr


@pytest.mark.gpu
def test_shim_with_reference_types_known_answer_on_gpu():
    if not os.path.exists(REF_TYPES_EXE):
        pytest.skip("oracle/_ref/shim_reference_types was not built (needs the reference tree)")
    res = subprocess.run([REF_TYPES_EXE], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "repeated=(-1," in res.stdout
