"""CPU pins of the FFT-domain path (no GPU needed).

* ``scripts/fft/emul.cu`` runs the per-thread phases of ``csrc/fftcore.cuh`` (the very functions the kernels call)
  thread by thread on the host and compares with a float64 DFT: index algebra, paddings, twiddle conventions and the
  transposed (permuted -> natural) factorisation for L = 1024 / 2048 / 4096.
* ``scripts/fft/model.py`` states in numpy what ``csrc/fftcorr.cuh`` computes (block partition of the LS lag sums,
  overlap-save FIR, fused clean-and-correlate CAF segments) and compares with direct float64 sums of the reference's
  loops (clutter_removal.py:34-51, range_doppler_processing.py:81-86).
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


@pytest.mark.skipif(_nvcc() is None, reason="nvcc not available")
def test_fftcore_thread_emulation_matches_float64_dft(tmp_path):
    exe = str(tmp_path / "fft_emul")
    src = os.path.join(ROOT, "scripts", "fft", "emul.cu")
    res = subprocess.run([_nvcc(), "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I",
                          os.path.join(ROOT, "passiveradar_b200", "csrc"), "-o", exe, src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.strip().endswith("OK")
    for line in run.stdout.splitlines()[:3]:
        assert "perm bijective: yes" in line


def test_numpy_model_of_fft_kernels_matches_direct_sums():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fft", "model.py")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "model OK" in res.stdout
