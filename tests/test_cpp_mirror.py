"""Builds and runs the C++ fluent mirror (include/enterpolation.hpp) against libenterp_b200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "test_fluent")
    libdir = os.path.join(ROOT, "enterpolation_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_fluent.cpp"), "-o", exe,
                    "-L" + libdir, "-l:libenterp_b200.so", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    return exe


def test_cpp_mirror_builders(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_device(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK (device path)" in out.stdout
