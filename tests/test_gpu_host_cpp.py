"""Compiles and runs the C++ host mirror's test (algo_dsp_b200/host/conv_host_test.cpp), which
re-states reference tests against conv.hpp -> C ABI -> CUDA."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = str(tmp_path / "conv_host_test")
    pkg = os.path.join(ROOT, "algo_dsp_b200")
    subprocess.run(["g++", "-std=c++17", os.path.join(pkg, "host", "conv_host_test.cpp"), "-o", exe, "-L" + pkg, "-lalgodsp_cuda",
                    "-Wl,-rpath," + pkg], check=True)
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    _compile(tmp_path)


@pytest.mark.gpu
def test_cpp_host_mirror_runs(tmp_path):
    r = subprocess.run([_compile(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "conv_host_test: ok" in r.stdout, r.stdout + r.stderr
