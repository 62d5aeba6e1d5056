"""The C++ mirror of the reference's interface (include/audio_matcher.hpp) running the reference's own unit tests
(tests/cpp_reference_tests.cpp).  On a CPU box the host-side tests run and the GPU algorithm must refuse to
construct; on a GPU box (`-m gpu`) the correlation and find_peaks KATs run through the same binary."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from audio_matcher_b200 import _native
    _native.build_native()
    exe = tmp_path / "cpp_reference_tests"
    libdir = os.path.join(ROOT, "audio_matcher_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp_reference_tests.cpp"), "-L", libdir, "-laudio_matcher_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    return exe


def test_cpp_mirror_host_side(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    out = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp reference tests ok (host only)" in out.stdout, out.stderr


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu(tmp_path):
    out = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "cpp reference tests ok", out.stdout + out.stderr
