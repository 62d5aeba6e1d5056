"""Runs the CUDA kernels' register-FFT index arithmetic on the CPU (tests/host_emul.cu): every length
2^4..2^14, row and column-tile batching, forward and inverse, and the forward->inverse register hand-over."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_fft_emulation(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc) and not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    exe = tmp_path / "host_emul"
    subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-diag-suppress", "128,20011", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_emul.cu")], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-2000:]
