"""GPU parity of the rows either side of the hot path: front end (deinterleave_IQ /
frequency_shift / resample / fused frontend), CFAR_2D and direct_xambg, through Python -> ctypes ->
C ABI, against goldens produced by the reference's own functions and against the live oracle."""
import numpy as np
import pytest

import _golden as G
import passiveradar_b200 as prb
from passiveradar_b200 import synth
from oracle import frontend_oracle as FO
from oracle.clutter_oracle import frequency_shift_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5          # relative infinity norm, SURVEY.md 8d


# ------------------------------------------------------------------ CFAR_2D
@pytest.mark.parametrize("name", G.CFAR_ALL)
def test_cfar_matches_reference_golden(name):
    g = G.load(name)
    fw, gw = int(g["fw"]), int(g["gw"])
    if bool(g["has_thresh"]):
        th = float(g["thresh"])
        det = prb.CFAR_2D(g["x"], fw, gw, th)
        assert det.dtype == np.bool_ and det.shape == g["out"].shape
        cr = FO.cfar_2d_oracle(g["x"], fw, gw)
        flips = det != g["out"]
        assert np.all(np.abs(cr[flips] - th) <= 1e-5 * th), "only cells within 1e-5 of the threshold may differ"
    else:
        cr = prb.CFAR_2D(g["x"], fw, gw)
        assert cr.dtype == np.float64 and cr.shape == g["out"].shape
        assert G.rel_inf(cr, g["out"]) <= TOL
        assert np.max(np.abs(cr - g["out"]) / np.abs(g["out"])) <= 1e-4      # per cell as well


def test_cfar_on_a_gpu_map_finds_the_target():
    ref, srv = synth.make_frame(2 ** 18, "P1", 2)
    clean = prb.LS_Filter(ref, srv, 100, 1.0, 10)
    amb = np.abs(prb.fast_xambg(ref, clean, 100, 64, 2 ** 18, ("kaiser", 5.0))[:, :, 0])
    cr = prb.CFAR_2D(amb, 18, 4)
    assert G.rel_inf(cr, FO.cfar_2d_oracle(amb, 18, 4)) <= TOL
    peak = np.unravel_index(np.argmax(cr), cr.shape)
    assert peak[1] == 100 - 40              # P1 target at delay 40 (column R - d)


def test_cfar_argument_errors():
    with pytest.raises(ValueError):
        prb.CFAR_2D(np.zeros(5, np.float32), 18, 4)
    with pytest.raises(TypeError):
        prb.CFAR_2D(np.zeros((4, 4), np.complex64), 3, 1)
    with pytest.raises(prb._lib.PrcoreError):
        prb.CFAR_2D(np.ones((4, 4), np.float32), 100, 4)


# ------------------------------------------------------------------ direct_xambg
@pytest.mark.parametrize("name", G.DIRECT_ALL)
def test_direct_xambg_matches_reference_golden(name):
    g = G.load(name)
    ref, srv = G.inputs(g)
    out = prb.direct_xambg(ref, srv, int(g["R"]), int(g["F"]), float(g["fs"]))
    assert out.dtype == np.complex64 and out.shape == g["out"].shape
    assert G.rel_inf(out, g["out"]) <= TOL


def test_direct_xambg_peak_agrees_with_fast_xambg():
    n, R, F = 2 ** 16, 60, 64
    ref, srv = synth.make_frame(n, "P1", 4)
    clean = prb.LS_Filter(ref, srv, R, 1.0, 10)
    fast = np.abs(prb.fast_xambg(ref, clean, R, F, n, None)[:, :, 0])
    direct = np.abs(prb.direct_xambg(ref, clean, R, F, float(n))[:, :, 0])
    assert np.unravel_index(np.argmax(fast), fast.shape)[1] == np.unravel_index(np.argmax(direct), direct.shape)[1]
    with pytest.raises(ValueError):
        prb.direct_xambg(ref, srv[:-1], R, F, 1.0)


# ------------------------------------------------------------------ front end
@pytest.mark.parametrize("kind", ["int8", "int16", "float32"])
@pytest.mark.parametrize("n", [1000, 1001])
def test_deinterleave_is_exact(kind, n):
    iq = synth.raw_iq(n, kind, 5)[: 2 * n - (n & 1)]          # odd total length for odd n
    out = prb.deinterleave_IQ(iq)
    ref = FO.deinterleave_iq_oracle(iq)
    assert out.dtype == np.complex64
    np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("name", G.FRONT_SMALL + G.FRONT_BIG)
def test_frontend_stages_match_reference_golden(name):
    g = G.load(name)
    iq, fc, fs, po, up, dn = G.front_inputs(g)
    x = prb.deinterleave_IQ(iq)
    np.testing.assert_array_equal(x[g["x_idx"]], g["deint_sub"])
    xs = prb.frequency_shift(x, fc, fs, po)
    assert str(xs.dtype) == str(g["shift_dtype"])
    den = np.abs(g["shift_sub"]).max()
    assert np.abs(xs[g["x_idx"]] - g["shift_sub"]).max() / den <= TOL
    y = prb.resample(xs, up, dn)
    assert str(y.dtype) == str(g["out_dtype"]) and y.shape[0] == int(g["out_len"])
    yden = float(g["out_absmax"])
    assert np.abs(y[g["out_idx"]] - g["out_sub"]).max() / yden <= TOL
    # fused: raw IQ in, resampled block out, one pass on the device
    yf = prb.frontend(iq, fc, fs, po, up, dn)
    assert yf.dtype == np.complex64 and yf.shape[0] == int(g["out_len"])
    assert np.abs(yf[g["out_idx"]] - g["out_sub"]).max() / yden <= TOL
    if "out" in g:
        assert G.rel_inf(yf, g["out"]) <= TOL
        assert G.rel_inf(y, g["out"]) <= TOL
    assert abs(yf.astype(np.complex128).sum() - complex(g["out_sum"])) <= 1e-4 * yden * np.sqrt(yf.shape[0])


@pytest.mark.parametrize("name", G.RESAMPLE_ALL)
def test_resample_matches_reference_golden(name):
    g = G.load(name)
    ref, _ = synth.make_frame(int(g["n"]), str(g["profile"]), 3)
    x = ref.astype(str(g["dtype"]))
    y = prb.resample(x, int(g["up"]), int(g["dn"]))
    assert y.dtype == g["out"].dtype and y.shape == g["out"].shape
    assert G.rel_inf(y, g["out"]) <= TOL


@pytest.mark.parametrize("fc,fs,po", [(300e3, 2.4e6, 0), (-777.25, 250000.0, 0.75), (12.5, np.float64(48000.0), 0),
                                       (np.float64(1234.5), 1e6, np.float64(0.1)), (np.int64(1200), 1e6, np.array([0.3])), (0.0, 1.0, 0)])
def test_frequency_shift_follows_numpy_promotion(fc, fs, po):
    ref, _ = synth.make_frame(300_000, "P0", 6)
    want = frequency_shift_oracle(ref, fc, fs, po)
    got = prb.frequency_shift(ref, fc, fs, po)
    assert got.dtype == want.dtype
    assert G.rel_inf(got, want) <= TOL


def test_resample_edges_and_identity():
    ref, _ = synth.make_frame(5000, "P1", 7)
    np.testing.assert_array_equal(prb.resample(ref, 7, 7), ref)            # up == down: a copy
    y = prb.resample(ref, 26, 238)                                         # reduces to 13/119
    assert G.rel_inf(y, FO.resample_oracle(ref, 13, 119)) <= TOL
    short = ref[:40]                                                       # shorter than the filter: line extension dominates
    assert G.rel_inf(prb.resample(short, 13, 119), FO.resample_oracle(short, 13, 119)) <= TOL
    assert G.rel_inf(prb.resample(short, 3, 2), FO.resample_oracle(short, 3, 2)) <= TOL
    with pytest.raises(ValueError):
        prb.resample(ref, 0, 3)
    with pytest.raises(ValueError):
        prb.resample(ref, 1.5, 3)
