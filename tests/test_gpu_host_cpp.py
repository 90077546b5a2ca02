"""Compiles and runs the C++ host mirror's test (algo_dsp_b200/host/conv_host_test.cpp), which
re-states reference tests against conv.hpp -> C ABI -> CUDA."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path, name="conv_host_test"):
    exe = str(tmp_path / name)
    pkg = os.path.join(ROOT, "algo_dsp_b200")
    subprocess.run(["g++", "-std=c++17", "-Wall", os.path.join(pkg, "host", name + ".cpp"), "-o", exe, "-L" + pkg, "-lalgodsp_cuda",
                    "-Wl,-rpath," + pkg], check=True)
    return exe


@pytest.mark.parametrize("name", ["conv_host_test", "post_host_test"])
def test_cpp_host_mirror_compiles_and_links(tmp_path, name):
    _compile(tmp_path, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["conv_host_test", "post_host_test"])
def test_cpp_host_mirror_runs(tmp_path, name):
    """conv.hpp (dsp/conv) and post.hpp (measure/ir, measure/sweep, dsp/filter/fir, dsp/resample, dsp/signal)."""
    r = subprocess.run([_compile(tmp_path, name)], capture_output=True, text=True)
    assert r.returncode == 0 and name + ": ok" in r.stdout, r.stdout + r.stderr
