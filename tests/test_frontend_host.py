"""Host-side logic of the front-end wrappers (no GPU): NumPy-2 promotion rules that decide the phase
arithmetic, raw-IQ classification, and the resample_poly filter design handed to the device."""
import numpy as np
import pytest
import scipy.signal as signal

from passiveradar_b200 import _lib, signal_utils as su
from oracle.clutter_oracle import frequency_shift_oracle


@pytest.mark.parametrize("fc,fs,po,mode", [
    (300e3, 2.4e6, 0, _lib.MIX_C64),                       # python scalars: complex64 throughout
    (300e3, 2.4e6, 0.5, _lib.MIX_C64),
    (np.float64(300e3), 2.4e6, np.float64(0.5), _lib.MIX_C64),   # np.float64 subclasses float: Python multiplies first
    (300e3, 2.4e6, np.float32(0.5), _lib.MIX_C64),
    (300e3, 2.4e6, np.array([0.5]), _lib.MIX_C128),        # what main.py:127-130 passes
    (300e3, np.float64(2.4e6), 0, _lib.MIX_FS64),          # NumPy sees the divisor
    (np.int64(300000), 2.4e6, 0, _lib.MIX_F64),
])
def test_mix_mode_matches_numpy_result_type(fc, fs, po, mode):
    got_mode, got_po = su._mix_mode(fc, fs, po)
    assert got_mode == mode
    assert got_po == pytest.approx(float(np.asarray(po).reshape(())))
    # the mode's output type is what NumPy really produces for a complex64 input
    want = frequency_shift_oracle(np.ones(4, np.complex64), fc, fs, po).dtype
    assert (np.complex64 if mode == _lib.MIX_C64 else np.complex128) == want


def test_mix_mode_rejects_vector_and_complex_offsets():
    with pytest.raises(ValueError):
        su._mix_mode(1.0, 2.0, np.zeros(3))
    with pytest.raises(TypeError):
        su._mix_mode(1.0, 2.0, 1j)


def test_raw_iq_kinds_and_odd_lengths():
    for dt, kind in [(np.int8, _lib.IQ_I8), (np.int16, _lib.IQ_I16), (np.float32, _lib.IQ_C64), (np.float64, _lib.IQ_C64),
                     (np.uint8, _lib.IQ_C64)]:
        buf, k, n = su._raw_iq(np.arange(11).astype(dt))
        assert k == kind and n == 5 and buf.shape[0] == 10 and buf.flags.c_contiguous
        if k == _lib.IQ_C64:
            assert buf.dtype == np.float32
    with pytest.raises(ValueError):
        su._raw_iq(np.zeros((2, 2)))


@pytest.mark.parametrize("up,dn", [(13, 119), (3, 2), (1, 4)])
def test_resample_taps_are_scipys(up, dn):
    """resample_poly designs firwin(2*10*max+1, 1/max, kaiser 5.0) * up; passing our taps back as `window`
    must reproduce its output exactly."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4000) + 1j * rng.standard_normal(4000)
    h = su._resample_taps(up, dn, False)
    assert h.shape[0] == 20 * max(up, dn) + 1
    a = signal.resample_poly(x, up, dn, padtype='line')
    b = signal.resample_poly(x, up, dn, window=h / up, padtype='line')
    np.testing.assert_array_equal(a, b)
    h32 = su._resample_taps(up, dn, True)
    x32 = x.astype(np.complex64)
    a32 = signal.resample_poly(x32, up, dn, padtype='line')
    hw = signal.firwin(20 * max(up, dn) + 1, 1.0 / max(up, dn), window=('kaiser', 5.0)).astype(np.float32)
    assert np.array_equal(h32, (hw * np.float32(up)).astype(np.float64))
    assert a32.dtype == np.complex64


def test_reduced_ratio_and_errors():
    assert su._reduced(26, 238) == (13, 119)
    for bad in [(0, 3), (2, 0), (1.5, 2), (2, 2.5)]:
        with pytest.raises(ValueError):
            su._reduced(*bad)
