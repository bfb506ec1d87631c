"""Development check of the FFT-domain path on a GPU box (not collected by pytest: no test_ prefix): errors against the
float64 truths / the CPU oracle on both paths and a first timing.  Prints everything, asserts nothing.  It lives under
tests/ because it uses the oracle, which only tests/, smoke() and bench.py's CPU arm may import.
    python tests/dev_gpu_check.py"""
import os
import sys
import time

import numpy as np
import scipy.signal as signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import passiveradar_b200 as prb
from passiveradar_b200 import _lib, synth
from oracle import clutter_oracle as co
from oracle import xambg_oracle as xo


def rel(a, b, den=None):
    d = np.abs(b).max() if den is None else den
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / d)


def ls_case(n, R, peek=10, reg=1.0, profile="P1"):
    ref, srv = synth.make_frame(n, profile, 1)
    t_out, t_taps = co.ls_filter_truth(ref, srv, R, reg, peek)
    res = {}
    for mode in (1, 0):
        _lib.set_option("fft", mode)
        out, taps = prb.LS_Filter(ref, srv, R, reg, peek, True)
        res[mode] = (rel(taps, t_taps), rel(out, t_out, np.abs(srv).max()))
    print(f"LS n={n} M={R + peek} {profile}: fft taps {res[1][0]:.2e} out {res[1][1]:.2e} | direct taps {res[0][0]:.2e} out {res[0][1]:.2e}", flush=True)


def x_case(n, F, R, win=True, profile="P1"):
    ref, srv = synth.make_frame(n, profile, 2)
    w = signal.get_window(("kaiser", 5.0), n) if win else None
    want = xo.fast_xambg_oracle(ref, srv, R, F, n, w)
    res = {}
    for mode in (1, 0):
        _lib.set_option("fft", mode)
        res[mode] = rel(prb.fast_xambg(ref, srv, R, F, n, w), want)
    print(f"CAF n={n} F={F} R={R} win={win}: fft {res[1]:.2e} | direct {res[0]:.2e}", flush=True)


def frame_case(n, F, R, profile="P0"):
    from passiveradar_b200.frames import FramePipeline
    import torch
    refs, srvs = zip(*[synth.make_frame(n, profile, 10 + i) for i in range(3)])
    ref_h = np.stack(refs)
    srv_h = np.stack(srvs)
    w = signal.get_window(("kaiser", 5.0), n)
    res = {}
    for mode in (1, 0):
        _lib.set_option("fft", mode)
        pipe = FramePipeline(n, R, F, batch=2, nslots=2)
        res[mode] = pipe.run_host(ref_h, srv_h)
    # truth: float64 LS then oracle CAF
    errs = []
    for i in range(3):
        t_out, _ = co.ls_filter_truth(ref_h[i], srv_h[i], R, 1.0, 10)
        want = xo.fast_xambg_oracle(ref_h[i], t_out.astype(np.complex64), R, F, n, w)
        errs.append((rel(res[1][i], want), rel(res[0][i], want)))
    print(f"FRAME n={n} F={F} R={R} {profile}: (fft, direct) vs truth-chained oracle {['%.2e/%.2e' % e for e in errs]}; fft vs direct {rel(res[1], res[0]):.2e}", flush=True)


def timing(n=2 ** 20, F=256, R=300, B=32):
    import torch
    from passiveradar_b200.frames import FramePipeline
    dev = torch.device("cuda", 0)
    ref_h = np.stack([synth.make_frame(n, "P1", i)[0] for i in range(4)])
    srv_h = np.stack([synth.make_frame(n, "P1", i)[1] for i in range(4)])
    ref_d = torch.from_numpy(np.tile(ref_h, (B // 4, 1))).to(dev)
    srv_d = torch.from_numpy(np.tile(srv_h, (B // 4, 1))).to(dev)
    maps = torch.empty((B, F, R + 1), dtype=torch.complex64, device=dev)
    for mode, batch, slots in ((1, 1, 8), (1, 4, 4), (1, 8, 3), (1, 16, 2), (1, 32, 1), (0, 1, 8)):
        _lib.set_option("fft", mode)
        pipe = FramePipeline(n, R, F, batch=batch, nslots=slots)
        for _ in range(3):
            pipe.run_device(ref_d, srv_d, maps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            pipe.run_device(ref_d, srv_d, maps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"TIMING fft={mode} batch={batch} slots={slots}: {B * reps / ms * 1e3:.0f} frames/s ({ms / (B * reps) * 1e3:.1f} us/frame)", flush=True)
        # per-kernel
        _lib.profile_reset()
        _lib.profile(True)
        single = FramePipeline(n, R, F, batch=batch, nslots=1)
        single.run_device(ref_d, srv_d, maps)
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile(False)
        print("   per launch (one stream): " + ", ".join(f"{k} {1e3 * v[0] / v[1]:.1f}us x{v[1]}" for k, v in prof.items() if v[1]), flush=True)


if __name__ == "__main__":
    _lib.set_option("fft_min_n", 0)
    t0 = time.time()
    ls_case(16384, 40)
    ls_case(2 ** 18, 100)
    ls_case(200000, 100)
    ls_case(2 ** 20, 300)
    ls_case(2 ** 20, 300, profile="P0")
    ls_case(2 ** 17, 1200, peek=10)            # L = 4096
    x_case(16384, 32, 40)
    x_case(200000, 64, 100)
    x_case(2 ** 20, 256, 300)
    x_case(2 ** 19, 1024, 175)
    x_case(10 ** 6, 256, 300, win=False, profile="P0")
    frame_case(2 ** 18, 128, 100)
    frame_case(2 ** 20, 256, 300)
    frame_case(200000, 64, 100, "P1")
    print(f"checks took {time.time() - t0:.1f}s", flush=True)
    timing()
