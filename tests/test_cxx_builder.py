"""The C++ host mirror (rust_lbfgs_b200/cxx/lbfgsb200.hpp) compiled with plain g++ against liblbfgsb200.so:
the reference's own tests/simple.rs + doc-tests restated in C++ (tests/cxx/test_builder.cpp)."""
import os
import subprocess

import pytest

from rust_lbfgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_builder")


def compile_it():
    _lib.lib()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(_lib.SO_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-ffp-contract=off", os.path.join(ROOT, "tests", "cxx", "test_builder.cpp"),
           "-o", EXE, "-L" + libdir, "-llbfgsb200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cxx_builder_compiles_and_checks_params():
    compile_it()
    r = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok (no-gpu)" in r.stdout


@pytest.mark.gpu
def test_cxx_builder_reference_tests_on_gpu():
    compile_it()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.strip().endswith("ok")
