"""Runs the C++ shim tests (tests/cpp/test_shim.cpp: the reference's own
test_sparse_operator / test_documentation written against
include/aboria_b200/Aboria.h) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_shim")


def test_shim_binary_is_built():
    if not os.path.exists(BIN):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_shim_on_gpu():
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all shim tests passed" in r.stdout


GLUE = os.path.join(ROOT, "tests", "cpp", "test_eigen_glue")


def test_eigen_glue_type_checks_against_the_mock():
    # compiling include/aboria_b200/EigenGlue.h against tests/cpp/mock_eigen IS the test on a CPU-only box
    if not os.path.exists(GLUE):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(GLUE)


@pytest.mark.gpu
def test_eigen_glue_on_gpu():
    r = subprocess.run([GLUE], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "eigen glue tests passed" in r.stdout
