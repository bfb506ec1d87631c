"""Argument validation of the drop-in wrappers happens BEFORE the C call ("Wrapper validates
before crossing into C"), so it can be checked without a GPU; and without a GPU every operator must raise
-- there is no CPU fallback."""
import numpy as np
import pytest

import passiveradar_b200 as prb
from passiveradar_b200 import _lib, build as prb_build


def _sig(n, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


@pytest.mark.parametrize("call", [
    lambda a, b: prb.fast_xambg(a, b, 10, 8),
    lambda a, b: prb.direct_xambg(a, b, 10, 8, 1000.0),
    lambda a, b: prb.LS_Filter(a, b, 10),
    lambda a, b: prb.LS_Filter_Toeplitz(a, b, 10),
    lambda a, b: prb.LS_Filter_Multiple(a, b, 10, 1000.0),
])
def test_length_mismatch_raises_the_references_value_error(call, capsys):
    a, b = _sig(256), _sig(255)
    with pytest.raises(ValueError, match="same length"):
        call(a, b)


def test_front_end_and_cfar_argument_errors():
    x = _sig(100)
    with pytest.raises(ValueError):
        prb.frequency_shift(x.reshape(10, 10), 1.0, 10.0)
    with pytest.raises(ValueError):
        prb.frequency_shift(x, 1.0, 10.0, np.zeros(3))
    with pytest.raises(ValueError):
        prb.resample(x, 0, 3)
    with pytest.raises(ValueError):
        prb.resample(x, 2.5, 3)
    with pytest.raises(ValueError):
        prb.deinterleave_IQ(np.zeros((4, 4), np.int8))
    with pytest.raises(ValueError):
        prb.frontend(np.zeros(0, np.int8), 1.0, 2.0, 0, 13, 119)
    with pytest.raises(ValueError):
        prb.CFAR_2D(np.zeros(8, np.float32), 18, 4)
    with pytest.raises(TypeError):
        prb.CFAR_2D(np.zeros((8, 8), np.complex64), 3, 1)
    with pytest.raises(ValueError):
        prb.block_NLMS(x, x, 10, 0.05, blockLen=0)
    # up == down: resample_poly returns a copy without touching the device
    y = prb.resample(x, 5, 5)
    assert y is not x and np.array_equal(y, x)
    assert prb.LS_Filter_Multiple(x, x, 10, 1000.0, dopplerBins=[]) is x      # the reference's loop never runs


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load()
    count = C_int_device_count(lib)
    if count > 0:
        pytest.skip("a CUDA device is present")
    x = _sig(512)
    for call in (lambda: prb.LS_Filter(x, x, 8), lambda: prb.fast_xambg(x, x, 8, 8), lambda: prb.NLMS_filter(x, x, 8, 0.05),
                 lambda: prb.deinterleave_IQ(np.zeros(64, np.int8)), lambda: prb.CFAR_2D(np.ones((8, 8), np.float32), 3, 1)):
        with pytest.raises(_lib.PrcoreError):
            call()


def C_int_device_count(lib):
    import ctypes as C
    n = C.c_int(0)
    st = lib.prc_device_count(C.byref(n))
    return n.value if st == 0 else 0


def test_build_tracks_every_header():
    """A stale libprcore.so after editing a header is the worst kind of bug: the up-to-date check must see all
    of csrc/ (it once listed only two of the eight headers)."""
    import os
    listed = {os.path.basename(h) for h in prb_build.HEADERS}
    on_disk = {f for f in os.listdir(prb_build.CSRC) if f.endswith((".cuh", ".h"))}
    assert on_disk <= listed and "prcore.h" in listed
