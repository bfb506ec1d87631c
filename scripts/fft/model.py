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


# ---------------------------------------------------------------------------------------------------------------------
# "Shared spectra" frame (fftcorr.cuh: lsspec_fft_kernel + cafspec_fft_kernel): the LS lag sums and the CAF use ONE
# segmentation of the channel and share the spectra of the reference / surveillance windows of every segment.
def segment_table(n, F, R, M, peek, L):
    """Python twin of build_seg_table() in prcore.cu.  Doppler block j sums samples [lo_j, hi_j) (boxcar of D + 1 taps,
    blocks overlap by one sample); each block is cut into nseg_j segments of <= bmax samples.  Per segment:
    (i0, ln, mlo, mhi): CAF uses ref[i0 : i0 + ln) (ln = 0: no CAF), the LS sums own ref[i0 + mlo : i0 + mhi) -- every
    sample of [0, n) is owned exactly once (the shared boundary samples once, the tail beyond the last block by extra
    segments).  blk[j] .. blk[j + 1] are the segments of block j."""
    D = n // F
    ntaps = D + 1
    half = (ntaps - 1) // 2
    pre_pad = D - half % D
    c0 = ((half + pre_pad) // D) * D - pre_pad
    off = M - 1 - peek                              # position of sample i0 inside the reference window
    bmax = min(L - R - (M - 1), L - off - (M - 1))  # CAF lags and LS lags stay inside the windows
    assert bmax >= 1
    segs, blk = [], [0]
    owned = 0                                       # samples [0, owned) already belong to an earlier segment
    for j in range(F):
        lo = max(j * D + c0 - (ntaps - 1), 0)
        hi = min(j * D + c0 + 1, n)
        if hi > lo:
            nseg = -(-(hi - lo) // bmax)
            Bs = -(-(hi - lo) // nseg)
            for q in range(nseg):
                i0 = lo + q * Bs
                ln = min(Bs, hi - i0)
                mlo = max(owned - i0, 0)
                mhi = max(ln, mlo)
                segs.append((i0, ln, mlo, mhi))
                owned = max(owned, i0 + ln)
        blk.append(len(segs))
    while owned < n:                                # what the CAF ignores still counts for the LS sums
        ln = min(bmax, n - owned)
        segs.append((owned, 0, 0, ln))
        owned += ln
    return segs, blk, off, c0, D


def model_shared_frame(ref, srv, win, R, F, M, peek, reg, L):
    n = len(ref)
    segs, blk, off, c0, D = segment_table(n, F, R, M, peek, L)
    cover = np.zeros(n, int)
    accC = np.zeros(L, complex)
    accX = np.zeros(L, complex)
    spec = []
    for (i0, ln, mlo, mhi) in segs:
        a = i0 - off                                # reference window start; surveillance window starts at a - peek
        Rw = np.fft.fft(ref[(a + np.arange(L)) % n])
        Sw = np.fft.fft(srv[(a - peek + np.arange(L)) % n])
        xu = np.zeros(L, complex)
        xu[mlo:mhi] = ref[i0 + mlo:i0 + mhi]
        cover[i0 + mlo:i0 + mhi] += 1
        Xu = np.fft.fft(xu)
        accC += Xu * np.conj(Rw)
        accX += Xu * np.conj(Sw)
        spec.append((Rw, Sw))
    assert (cover == 1).all()
    c = np.fft.fft(accC)[off:off + M] / L          # lag m sits at output index off + m
    x = np.fft.fft(accX)[off:off + M] / L
    # normal equations as levinson_kernel sees them (it conjugates): T[a, b] = conj(c)[a - b], rhs = conj(x)
    t = np.conj(c)
    T = np.array([[t[a - b] if a >= b else np.conj(t[b - a]) for b in range(M)] for a in range(M)]) + reg * np.eye(M)
    w = np.linalg.solve(T, np.conj(x))
    h = np.zeros(L, complex)
    h[:M] = w                                       # plain causal taps: valid for window positions >= M - 1
    W = np.fft.fft(h)
    xw = ref * (win if win is not None else 1.0)
    P = np.zeros((F, R + 1), complex)
    for j in range(F):
        acc = np.zeros(L, complex)
        for s in range(blk[j], blk[j + 1]):
            i0, ln, _, _ = segs[s]
            Rw, Sw = spec[s]
            x_ = np.zeros(L, complex)
            x_[:ln] = xw[i0:i0 + ln]
            acc += np.fft.fft(x_) * np.conj(Sw - Rw * W)
        P[j] = np.fft.fft(acc)[M - 1:M - 1 + R + 1] / L     # lag d sits at output index (M - 1) + d
    return w, P


for n, F, R, M, peek, L in [(16384, 8, 40, 50, 10, 2048), (20000, 16, 100, 110, 10, 2048), (12345, 7, 33, 30, 0, 1024), (8192, 32, 20, 30, 10, 1024)]:
    ref = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    srv = rng.standard_normal(n) + 1j * rng.standard_normal(n) + 0.7 * np.roll(ref, 3)
    win = np.kaiser(n, 5.0)
    w, P = model_shared_frame(ref, srv, win, R, F, M, peek, 1.0, L)
    # direct: LS_Filter then the block sums of fast_xambg
    A = np.stack([np.roll(ref, k - peek) for k in range(M)], axis=1)
    w_ref = np.linalg.solve(A.conj().T @ A + np.eye(M), A.conj().T @ srv)
    check(f"shared n={n} F={F} R={R} M={M} L={L} taps", w, w_ref)
    clean = srv - A @ w_ref
    D = n // F
    _, _, _, c0, _ = segment_table(n, F, R, M, peek, L)
    want = np.zeros((F, R + 1), complex)
    xw = ref * win
    for j in range(F):
        for m in range(D + 1):
            i = j * D + c0 - m
            if 0 <= i < n:
                want[j] += xw[i] * np.conj(clean[(i + np.arange(R + 1)) % n])
    check(f"shared n={n} F={F} R={R} M={M} L={L} block sums", P, want)
print("shared model OK")
