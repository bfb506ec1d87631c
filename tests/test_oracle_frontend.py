"""oracle/frontend_oracle.py pinned against goldens produced by the reference's own functions
(deinterleave_IQ / frequency_shift / resample, CFAR_2D, direct_xambg) -- CPU only."""
import numpy as np
import pytest

import _golden as G
from oracle import frontend_oracle as FO
from oracle.clutter_oracle import frequency_shift_oracle


@pytest.mark.parametrize("name", G.CFAR_ALL)
def test_cfar_oracle_is_the_reference(name):
    g = G.load(name)
    thresh = float(g["thresh"]) if bool(g["has_thresh"]) else None
    out = FO.cfar_2d_oracle(g["x"], int(g["fw"]), int(g["gw"]), thresh)
    assert out.dtype == g["out"].dtype
    np.testing.assert_array_equal(out, g["out"])


@pytest.mark.parametrize("name", G.DIRECT_ALL)
def test_direct_xambg_oracle_is_the_reference(name):
    g = G.load(name)
    ref, srv = G.inputs(g)
    out = FO.direct_xambg_oracle(ref, srv, int(g["R"]), int(g["F"]), float(g["fs"]))
    assert out.dtype == np.complex64 and out.shape == g["out"].shape
    np.testing.assert_array_equal(out, g["out"])


def test_direct_xambg_reference_self_noise_is_small():
    g = G.load("direct_small")
    ref, srv = G.inputs(g)
    truth = FO.direct_xambg_truth(ref, srv, int(g["R"]), int(g["F"]), float(g["fs"]))
    assert G.rel_inf(g["out"], truth) < 1e-5


@pytest.mark.parametrize("name", G.FRONT_SMALL)
def test_frontend_oracle_is_the_reference(name):
    g = G.load(name)
    iq, fc, fs, po, up, dn = G.front_inputs(g)
    x = FO.deinterleave_iq_oracle(iq)
    xs = frequency_shift_oracle(x, fc, fs, po)
    assert str(xs.dtype) == str(g["shift_dtype"])
    np.testing.assert_array_equal(x[g["x_idx"]], g["deint_sub"])
    np.testing.assert_array_equal(xs[g["x_idx"]], g["shift_sub"])
    y = FO.frontend_oracle(iq, fc, fs, po, up, dn)
    assert str(y.dtype) == str(g["out_dtype"]) and y.shape[0] == int(g["out_len"])
    np.testing.assert_array_equal(y, g["out"])


@pytest.mark.parametrize("name", G.RESAMPLE_ALL)
def test_resample_oracle_is_the_reference(name):
    g = G.load(name)
    ref, _ = G.synth.make_frame(int(g["n"]), str(g["profile"]), 3)
    y = FO.resample_oracle(ref.astype(str(g["dtype"])), int(g["up"]), int(g["dn"]))
    np.testing.assert_array_equal(y, g["out"])


def test_polyphase_statement_matches_resample_poly():
    """The index arithmetic the device kernel uses (frontend.cuh) written out in float64 equals SciPy's
    upfirdn-based resample_poly(padtype='line'), including the line extension at both ends."""
    rng = np.random.default_rng(3)
    for n, up, dn in [(700, 13, 119), (257, 3, 2), (400, 1, 4), (300, 26, 238)]:
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        a = FO.resample_truth(x, up, dn)
        b = FO.resample_oracle(x, up, dn)
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())
