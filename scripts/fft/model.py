"""numpy model of the FFT-domain kernels (fftcorr.cuh): pins the algebra (block partition, aliasing margins, shifts,
conjugation conventions) against direct float64 sums before the CUDA transliteration is trusted.  Run: python scripts/fft/model.py"""
import numpy as np

rng = np.random.default_rng(7)


def direct_lags(x, s, lags, n, linear=False):
    """C[l] = sum_i x[i] conj(s[i + l]) (circular mod n, or zero outside [0, n) when linear)"""
    out = np.zeros(len(lags), complex)
    for q, l in enumerate(lags):
        if linear:
            sh = np.zeros(n, complex)
            if l >= 0:
                sh[:n - l] = s[l:]
            else:
                sh[-l:] = s[:n + l]
        else:
            sh = np.roll(s, -l)
        out[q] = np.sum(x * np.conj(sh))
    return out


def model_lscorr(ref, srv, M, peek, L, linear=False):
    """segments of Bs <= L - M + 1 samples; X = zero-padded reference segment, Yr / Ys = the L reference / surveillance
    samples that start with it (surveillance shifted by -peek); sum over segments of X conj(Y) is the spectrum of the lag sums"""
    n = len(ref)
    bmax = L - M + 1
    nseg = -(-n // bmax)
    Bs = -(-n // nseg)
    nseg = -(-n // Bs)

    def take(sig, base, ln):
        idx = base + np.arange(L)
        if linear:
            v = np.where((idx >= 0) & (idx < n), sig[np.clip(idx, 0, n - 1)], 0)
        else:
            v = sig[idx % n]
        v = v.astype(complex)
        v[ln:] = 0
        return np.fft.fft(v)

    accC = np.zeros(L, complex)
    accX = np.zeros(L, complex)
    for q in range(nseg):
        i0 = q * Bs
        ln = min(Bs, n - i0)
        X = take(ref, i0, ln)
        accC += X * np.conj(take(ref, i0, L))
        accX += X * np.conj(take(srv, i0 - peek, L))
    return np.fft.fft(accC)[:M] / L, np.fft.fft(accX)[:M] / L


def taps_spectrum(w, L):
    M = len(w)
    h = np.zeros(L, complex)
    h[(np.arange(M) - (M - 1)) % L] = w
    return np.fft.fft(h)


def model_fir(ref, srv, w, peek, L, linear=False):
    n, M = len(ref), len(w)
    Wp = taps_spectrum(w, L)
    Bf = L - M + 1
    out = np.zeros(n, complex)
    for p0 in range(0, n, Bf):
        idx = p0 + peek - (M - 1) + np.arange(L)
        r = np.where((idx >= 0) & (idx < n), ref[np.clip(idx, 0, n - 1)], 0) if linear else ref[idx % n]
        Y = np.fft.fft(r) * Wp
        yc = np.fft.fft(np.conj(Y))          # = L conj(y)
        y = np.conj(yc) / L
        m = min(Bf, n - p0)
        out[p0:p0 + m] = srv[p0:p0 + m] - y[:m]
    return out


def model_caf(ref, srv, win, R, F, L, w=None, peek=0):
    """block sums P[j, d], d = 0..R, optionally with the clutter FIR fused in the frequency domain"""
    n = len(ref)
    D = n // F
    ntaps = D + 1 if D > 1 else 1
    half = (ntaps - 1) // 2
    pre_pad = D - half % D
    pre_rem = (half + pre_pad) // D
    c0 = pre_rem * D - pre_pad if D > 1 else 0
    M = len(w) if w is not None else 1
    Bmax = L - R - (M - 1 if w is not None else 0)
    Wp = taps_spectrum(w, L) if w is not None else None
    xw = ref * (win if win is not None else 1.0)
    P = np.zeros((F, R + 1), complex)
    for j in range(F):
        lo = j * D + c0 - (ntaps - 1)
        hi = j * D + c0 + 1            # exclusive
        lo, hi = max(lo, 0), min(hi, n)
        if hi <= lo:
            continue
        nseg = -(-(hi - lo) // Bmax)
        Bs = -(-(hi - lo) // nseg)
        acc = np.zeros(L, complex)
        for q in range(nseg):
            i0 = lo + q * Bs
            ln = min(Bs, hi - i0)
            x = np.zeros(L, complex)
            x[:ln] = xw[i0:i0 + ln]
            S = np.fft.fft(srv[(i0 + np.arange(L)) % n])
            if w is not None:
                Rr = np.fft.fft(ref[(i0 + peek - (M - 1) + np.arange(L)) % n])
                S = S - Rr * Wp
            acc += np.fft.fft(x) * np.conj(S)
        P[j] = np.fft.fft(acc)[:R + 1] / L
    return P


def check(name, a, b, den=None):
    e = np.abs(a - b).max() / (den if den else np.abs(b).max())
    print(f"{name:56s} {e:.2e}")
    assert e < 1e-9, name


for n, M, peek, L in [(8192, 110, 10, 2048), (20000, 310, 10, 2048), (5000, 64, 0, 1024), (12345, 300, 7, 2048)]:
    ref = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    srv = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for linear in (False, True):
        C, X = model_lscorr(ref, srv, M, peek, L, linear)
        check(f"lscorr n={n} M={M} L={L} linear={linear} autocorr", C, direct_lags(ref, ref, range(M), n, linear))
        check(f"lscorr n={n} M={M} L={L} linear={linear} xcorr", X, direct_lags(ref, srv, range(-peek, M - peek), n, linear))
    w = (rng.standard_normal(M) + 1j * rng.standard_normal(M)) / M
    for linear in (False, True):
        got = model_fir(ref, srv, w, peek, L, linear)
        want = srv.copy()
        for k in range(M):
            sh = np.roll(ref, k - peek)
            if linear:
                d = k - peek
                sh = np.zeros(n, complex)
                if d >= 0:
                    sh[d:] = ref[:n - d]
                else:
                    sh[:n + d] = ref[-d:]
            want = want - w[k] * sh
        check(f"fir n={n} M={M} linear={linear}", got, want)

for n, F, R, M, peek, L in [(16384, 8, 40, 50, 10, 2048), (20000, 16, 100, 110, 10, 2048), (8192, 32, 20, 30, 10, 1024), (10000, 7, 33, 0, 0, 1024)]:
    ref = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    srv = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    win = np.kaiser(n, 5.0)
    D = n // F
    ntaps = D + 1
    half = (ntaps - 1) // 2
    pre_pad = D - half % D
    c0 = ((half + pre_pad) // D) * D - pre_pad
    w = (rng.standard_normal(M) + 1j * rng.standard_normal(M)) / M if M else None
    s_eff = srv if w is None else srv - sum(w[k] * np.roll(ref, k - peek) for k in range(M))
    want = np.zeros((F, R + 1), complex)
    xw = ref * win
    for j in range(F):
        for m in range(ntaps):
            i = j * D + c0 - m
            if 0 <= i < n:
                want[j] += xw[i] * np.conj(s_eff[(i + np.arange(R + 1)) % n])
    got = model_caf(ref, srv, win, R, F, L, w, peek)
    check(f"caf n={n} F={F} R={R} M={M} L={L}", got, want)
print("model OK")
