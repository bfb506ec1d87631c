// kernels.cuh -- sm_100a device code of libprcore (passive-radar hot path).
//
// Four kernels carry the frame  LS_Filter -> fast_xambg  (DESIGN.md section 3):
//
//   lagcorr_kernel   partial lag-correlation sums  C[d] = sum_i x[i] conj(s[i+d])
//                    used three times: Gram column + right-hand side of the LS
//                    normal equations (clutter_removal.py:39,45) and the per-Doppler-
//                    block lag products of the CAF (range_doppler_processing.py:81-86)
//   levinson_kernel  (T + reg I) w = rhs for the Hermitian Toeplitz Gram matrix, fp64
//                    (replaces np.linalg.solve, clutter_removal.py:42-45)
//   fir_apply_kernel out = srv - circular FIR(ref, w)              (clutter_removal.py:51)
//   doppler_fft_*    chunk sum + FFT along Doppler + fftshift  (range_doppler_processing.py:89)
//
// lagcorr and fir_apply share one inner loop, slide_mac<>: a register-tiled sliding
// window complex multiply-accumulate.  A thread owns TD accumulators and a ring of
// TI+TD window samples; every step it pulls TI fresh window samples (128-bit LDS,
// lane stride TD*8 B = an odd multiple of 16 B, hence bank-conflict free) and TI
// broadcast operands, and issues TI*TD complex MACs (4 FFMA each).  That is
// 4*TI*TD FFMA for TI 128-bit shared loads: the loop is FP32-pipe bound by design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prc {

template <int A, int B> struct Gcd { static constexpr int v = Gcd<B, A % B>::v; };
template <int A> struct Gcd<A, 0> { static constexpr int v = A; };

// acc += x * conj(w)   (CONJ)      or      acc += x * w
template <bool CONJ>
__device__ __forceinline__ void cmac(float2& acc, const float2 x, const float2 w) {
    if (CONJ) {
        acc.x = fmaf(x.x, w.x, acc.x);
        acc.x = fmaf(x.y, w.y, acc.x);
        acc.y = fmaf(x.y, w.x, acc.y);
        acc.y = fmaf(-x.x, w.y, acc.y);
    } else {
        acc.x = fmaf(x.x, w.x, acc.x);
        acc.x = fmaf(-x.y, w.y, acc.x);
        acc.y = fmaf(x.x, w.y, acc.y);
        acc.y = fmaf(x.y, w.x, acc.y);
    }
}

// acc[v] += sum_{t<nsteps, u<TI}  xs[t*TI+u] (*) ws[t*TI+u+v]       v = 0..TD-1
// xs: operand shared by the lanes of a group (broadcast loads); ws: this thread's window.
// Both must be 16-byte aligned; ws must be readable up to index nsteps*TI + TD - 1.
template <int TI, int TD, bool CONJ>
__device__ __forceinline__ void slide_mac(float2 (&acc)[TD], const float2* __restrict__ xs,
                                          const float2* __restrict__ ws, int nsteps) {
    static_assert(TI % 2 == 0 && TD % 2 == 0, "128-bit shared loads need even tile sizes");
    constexpr int RS = TI + TD;                       // ring size
    constexpr int PERIOD = RS / Gcd<RS, TI>::v;       // steps after which ring slots repeat
    static_assert(PERIOD <= 4, "unroll factor too large for the instruction cache");
    float2 W[RS];
#pragma unroll
    for (int q = 0; q < TD; q += 2) {
        const float4 v = *reinterpret_cast<const float4*>(ws + q);
        W[q] = make_float2(v.x, v.y);
        W[q + 1] = make_float2(v.z, v.w);
    }
    for (int t0 = 0; t0 < nsteps; t0 += PERIOD) {
#pragma unroll
        for (int p = 0; p < PERIOD; ++p) {
            if (t0 + p < nsteps) {
                const float2* wp = ws + (t0 + p) * TI + TD;
                const float2* xp = xs + (t0 + p) * TI;
#pragma unroll
                for (int q = 0; q < TI; q += 2) {
                    const float4 v = *reinterpret_cast<const float4*>(wp + q);
                    W[(p * TI + TD + q) % RS] = make_float2(v.x, v.y);
                    W[(p * TI + TD + q + 1) % RS] = make_float2(v.z, v.w);
                }
#pragma unroll
                for (int u = 0; u < TI; u += 2) {
                    const float4 xv = *reinterpret_cast<const float4*>(xp + u);
                    const float2 x0 = make_float2(xv.x, xv.y);
                    const float2 x1 = make_float2(xv.z, xv.w);
#pragma unroll
                    for (int v = 0; v < TD; ++v) cmac<CONJ>(acc[v], x0, W[(p * TI + u + v) % RS]);
#pragma unroll
                    for (int v = 0; v < TD; ++v) cmac<CONJ>(acc[v], x1, W[(p * TI + u + 1 + v) % RS]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// lagcorr: partial[prob][blk][chunk][l] = sum_{i in chunk of block blk} xw[i] * conj(s[(i + dmin + l) mod n])
//   xw[i] = x[i] * win[i] * taps[blk_hi - i]   (either factor optional), 0 outside [0, n)
// One CTA per (block, chunk, problem).  Threads = H lag groups (TD lags each) x G sample groups.
// ------------------------------------------------------------------------------------------
struct LagCorrParams {
    const float2* x;
    const float2* s[2];
    int dmin[2];
    const float* win;
    const float* taps;
    int n;
    long long blk_first_lo;
    long long blk_stride;
    int blk_len;
    int nblk;
    int nchunk;
    int chunk_len;
    int H;        // lag groups; padded lag count = H*TD
    int G;        // sample groups
    int steps;    // TI-sample steps per sample group
    float2* partial;
};

template <int TI, int TD>
__global__ void __launch_bounds__(512) lagcorr_kernel(const __grid_constant__ LagCorrParams p) {
    extern __shared__ __align__(16) float2 smem[];
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int blk = blockIdx.x / p.nchunk;
    const int chunk = blockIdx.x - blk * p.nchunk;
    const int prob = blockIdx.y;
    const int Sg = p.steps * TI;
    const int Lpad = p.G * Sg;
    const int HT = p.H * TD;
    float2* xs = smem;
    float2* ss = smem + Lpad;

    const long long blk_lo = p.blk_first_lo + (long long)blk * p.blk_stride;
    const long long blk_hi = blk_lo + p.blk_len - 1;
    const long long chunk_lo = blk_lo + (long long)chunk * p.chunk_len;
    int len = p.blk_len - chunk * p.chunk_len;
    if (len > p.chunk_len) len = p.chunk_len;

    // ---- stage the weighted x chunk (zero outside the block / the signal)
    const float2* __restrict__ x = p.x;
    for (int q = tid; q < Lpad; q += nthr) {
        float2 v = make_float2(0.f, 0.f);
        const long long i = chunk_lo + q;
        if (q < len && i >= 0 && i < p.n) {
            v = x[i];
            if (p.win) { const float w = p.win[i]; v.x *= w; v.y *= w; }
            if (p.taps) { const float w = p.taps[blk_hi - i]; v.x *= w; v.y *= w; }
        }
        xs[q] = v;
    }
    // ---- stage the circular window of s that the chunk's lags touch
    const float2* __restrict__ s = p.s[prob];
    {
        long long s0 = (chunk_lo + p.dmin[prob]) % p.n;
        if (s0 < 0) s0 += p.n;
        const unsigned start = (unsigned)s0;
        const unsigned n = (unsigned)p.n;
        for (int q = tid; q < Lpad + HT; q += nthr) {
            unsigned idx = start + (unsigned)q;
            if (idx >= n) { idx -= n; if (idx >= n) idx %= n; }
            ss[q] = s[idx];
        }
    }
    __syncthreads();

    float2 acc[TD];
#pragma unroll
    for (int v = 0; v < TD; ++v) acc[v] = make_float2(0.f, 0.f);
    const int h = tid % p.H;
    const int g = tid / p.H;
    const bool active = g < p.G;
    if (active) slide_mac<TI, TD, true>(acc, xs + g * Sg, ss + g * Sg + h * TD, p.steps);
    __syncthreads();

    // ---- reduce the G sample groups in a fixed order (deterministic), write the partial row
    float2* red = smem;
    if (active) {
#pragma unroll
        for (int v = 0; v < TD; ++v) red[g * HT + h * TD + v] = acc[v];
    }
    __syncthreads();
    float2* out = p.partial + (((size_t)prob * p.nblk + blk) * p.nchunk + chunk) * (size_t)HT;
    for (int l = tid; l < HT; l += nthr) {
        float2 sum = red[l];
        for (int gg = 1; gg < p.G; ++gg) {
            const float2 t = red[gg * HT + l];
            sum.x += t.x;
            sum.y += t.y;
        }
        out[l] = sum;
    }
}

// ------------------------------------------------------------------------------------------
// fir_apply: out[i] = srv[i] - sum_{k<M} taps[k] * ref[(i + peek - k) mod n]
// One CTA per LO = blockDim*TO consecutive outputs; taps reversed and zero-padded to Mpad.
// ------------------------------------------------------------------------------------------
struct FirParams {
    const float2* ref;
    const float2* srv;
    const float2* taps;
    float2* out;
    int n;
    int M;
    int Mpad;     // multiple of TK
    int peek;
};

template <int TK, int TO>
__global__ void __launch_bounds__(512) fir_apply_kernel(const __grid_constant__ FirParams p) {
    extern __shared__ __align__(16) float2 smem[];
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int LO = nthr * TO;
    float2* tr = smem;              // Mpad reversed taps
    float2* rs = smem + p.Mpad;     // LO + Mpad window of ref
    const long long I0 = (long long)blockIdx.x * LO;

    for (int q = tid; q < p.Mpad; q += nthr) {
        const int k = p.Mpad - 1 - q;
        tr[q] = (k < p.M) ? p.taps[k] : make_float2(0.f, 0.f);
    }
    {
        long long r0 = (I0 + p.peek - (p.Mpad - 1)) % p.n;
        if (r0 < 0) r0 += p.n;
        const unsigned start = (unsigned)r0;
        const unsigned n = (unsigned)p.n;
        const float2* __restrict__ ref = p.ref;
        for (int q = tid; q < LO + p.Mpad; q += nthr) {
            unsigned idx = start + (unsigned)q;
            if (idx >= n) { idx -= n; if (idx >= n) idx %= n; }
            rs[q] = ref[idx];
        }
    }
    __syncthreads();

    float2 acc[TO];
#pragma unroll
    for (int v = 0; v < TO; ++v) acc[v] = make_float2(0.f, 0.f);
    slide_mac<TK, TO, false>(acc, tr, rs + tid * TO, p.Mpad / TK);

    const long long i0 = I0 + (long long)tid * TO;
#pragma unroll
    for (int v = 0; v < TO; ++v) {
        const long long i = i0 + v;
        if (i < p.n) {
            const float2 d = p.srv[i];
            p.out[i] = make_float2(d.x - acc[v].x, d.y - acc[v].y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// levinson: reduce the chunk partials of the two LS correlations in fp64, then solve the
// Hermitian Toeplitz system (T + reg I) w = rhs by Levinson recursion (fp64, one warp).
//   t[m]   = conj( sum_c partial[0][c][m] )          m = 0..M-1   (first column of A^H A)
//   rhs[a] = conj( sum_c partial[1][c][a] )          a = 0..M-1   (A^H srv)
// status: 0 ok, 1 = not positive definite / non-finite pivot.
// ------------------------------------------------------------------------------------------
struct LevinsonParams {
    const float2* partial;   // [2][nchunk][HT]
    int nchunk;
    int HT;
    int M;
    double reg;
    float2* taps;            // M
    int* status;
};

__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 zmulc(double2 a, double2 b) {   // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) levinson_kernel(const __grid_constant__ LevinsonParams p) {
    extern __shared__ __align__(16) double2 zs[];
    const int M = p.M;
    double2* t = zs;            // Toeplitz first column (t[0] real, + reg)
    double2* r = zs + M;        // right-hand side
    double2* f0 = zs + 2 * M;   // forward vector (double buffered)
    double2* f1 = zs + 3 * M;
    double2* w = zs + 4 * M;    // solution
    const int tid = threadIdx.x;

    // phase A: fp64 sum over chunks (fixed order), conjugate
    for (int m = tid; m < 2 * M; m += blockDim.x) {
        const int prob = m / M;
        const int l = m - prob * M;
        const float2* src = p.partial + (size_t)prob * p.nchunk * p.HT + l;
        double sx = 0.0, sy = 0.0;
        for (int c = 0; c < p.nchunk; ++c) {
            const float2 v = src[(size_t)c * p.HT];
            sx += (double)v.x;
            sy += (double)v.y;
        }
        if (prob == 0) t[l] = make_double2(sx, -sy);
        else r[l] = make_double2(sx, -sy);
    }
    __syncthreads();
    if (tid >= 32) return;
    const int lane = tid;

    const double t0 = t[0].x + p.reg;
    int bad = !(t0 > 0.0) || !isfinite(t0);
    if (lane == 0) {
        f0[0] = make_double2(1.0 / t0, 0.0);
        w[0] = make_double2(r[0].x / t0, r[0].y / t0);
    }
    __syncwarp();
    double2* f = f0;
    double2* fn = f1;
    for (int n = 1; n < M; ++n) {
        // ef = sum_i t[n-i] f[i],  ex = sum_i t[n-i] w[i],  i = 0..n-1
        double efx = 0, efy = 0, exx = 0, exy = 0;
        for (int i = lane; i < n; i += 32) {
            const double2 tv = t[n - i];
            const double2 a = zmul(tv, f[i]);
            const double2 b = zmul(tv, w[i]);
            efx += a.x; efy += a.y; exx += b.x; exy += b.y;
        }
        efx = warp_sum(efx); efy = warp_sum(efy); exx = warp_sum(exx); exy = warp_sum(exy);
        const double den = 1.0 - (efx * efx + efy * efy);
        if (!(den > 0.0) || !isfinite(den)) bad = 1;
        const double inv = 1.0 / den;
        const double2 ef = make_double2(efx, efy);
        // fn[i] = ( f[i]*[i<n] - ef * conj(f[n-i])*[i>=1] ) / den ,  i = 0..n
        for (int i = lane; i <= n; i += 32) {
            double2 v = make_double2(0.0, 0.0);
            if (i < n) v = f[i];
            if (i >= 1) {
                const double2 c = zmulc(ef, f[n - i]);
                v.x -= c.x; v.y -= c.y;
            }
            fn[i] = make_double2(v.x * inv, v.y * inv);
        }
        __syncwarp();
        // w[i] += (r[n] - ex) * conj(fn[n-i]),  i = 0..n   (w[n] starts at 0)
        const double2 dlt = make_double2(r[n].x - exx, r[n].y - exy);
        for (int i = lane; i <= n; i += 32) {
            const double2 c = zmulc(dlt, fn[n - i]);
            double2 v = (i < n) ? w[i] : make_double2(0.0, 0.0);
            v.x += c.x; v.y += c.y;
            w[i] = v;
        }
        __syncwarp();
        double2* tmp = f; f = fn; fn = tmp;
    }
    for (int i = lane; i < M; i += 32) p.taps[i] = make_float2((float)w[i].x, (float)w[i].y);
    if (lane == 0) *p.status = bad;
}

// ------------------------------------------------------------------------------------------
// Doppler stage: P[j][lag] = sum_chunks partial[j][c][lag];  X = FFT_j(P);  out[(f+F/2)%F][R-lag]
// ------------------------------------------------------------------------------------------
__global__ void twiddle_kernel(float2* tw, int F) {      // tw[m] = exp(-2 pi i m / F), m < F
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < F) {
        double s, c;
        sincospi(-2.0 * (double)m / (double)F, &s, &c);
        tw[m] = make_float2((float)c, (float)s);
    }
}

struct DopplerParams {
    const float2* partial;   // [F][nchunk][HT]
    const float2* tw;        // [F]
    float2* out;             // [F][R+1]
    int F;
    int logF;
    int R;
    int nchunk;
    int HT;
};

// power-of-two F: radix-2 Stockham autosort in shared memory, CT range columns per CTA
template <int CT>
__global__ void __launch_bounds__(256) doppler_fft_pow2_kernel(const __grid_constant__ DopplerParams p) {
    extern __shared__ __align__(16) float2 smem[];
    const int F = p.F;
    float2* a = smem;
    float2* b = smem + (size_t)F * CT;
    const int k0 = blockIdx.x * CT;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < F * CT; idx += blockDim.x) {
        const int j = idx / CT, c = idx - j * CT;
        const int k = k0 + c;
        float2 sum = make_float2(0.f, 0.f);
        if (k <= p.R) {
            const float2* src = p.partial + (size_t)j * p.nchunk * p.HT + (p.R - k);
            for (int ch = 0; ch < p.nchunk; ++ch) {
                const float2 v = src[(size_t)ch * p.HT];
                sum.x += v.x;
                sum.y += v.y;
            }
        }
        a[idx] = sum;
    }
    __syncthreads();
    const int half = F >> 1;
    for (int st = 0; st < p.logF; ++st) {
        const int Ns = 1 << st;
        for (int idx = tid; idx < half * CT; idx += blockDim.x) {
            const int bf = idx / CT, c = idx - bf * CT;
            const int kk = bf & (Ns - 1);
            const float2 w = p.tw[kk * (half >> st)];          // exp(-2 pi i kk / (2 Ns))
            const float2 u = a[bf * CT + c];
            const float2 v0 = a[(bf + half) * CT + c];
            const float2 v = make_float2(v0.x * w.x - v0.y * w.y, v0.x * w.y + v0.y * w.x);
            const int j0 = ((bf - kk) << 1) + kk;
            b[j0 * CT + c] = make_float2(u.x + v.x, u.y + v.y);
            b[(j0 + Ns) * CT + c] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
    }
    for (int idx = tid; idx < F * CT; idx += blockDim.x) {
        const int fp = idx / CT, c = idx - fp * CT;
        const int k = k0 + c;
        if (k <= p.R) {
            const int f = (fp + half) & (F - 1);
            p.out[(size_t)f * (p.R + 1) + k] = a[idx];
        }
    }
}

// any F: chunk-sum into a compact [F][R+1] buffer, then a direct DFT with an exact twiddle table
__global__ void chunk_sum_kernel(const float2* __restrict__ partial, float2* __restrict__ P,
                                 int F, int R, int nchunk, int HT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F * (R + 1)) return;
    const int j = idx / (R + 1), k = idx - j * (R + 1);
    const float2* src = partial + (size_t)j * nchunk * HT + (R - k);
    float2 sum = make_float2(0.f, 0.f);
    for (int ch = 0; ch < nchunk; ++ch) {
        const float2 v = src[(size_t)ch * HT];
        sum.x += v.x;
        sum.y += v.y;
    }
    P[idx] = sum;
}

__global__ void doppler_dft_kernel(const float2* __restrict__ P, const float2* __restrict__ tw,
                                   float2* __restrict__ out, int F, int R) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int fp = blockIdx.y;
    if (k > R) return;
    float ax = 0.f, ay = 0.f;
    int m = 0;
    for (int j = 0; j < F; ++j) {
        const float2 v = P[(size_t)j * (R + 1) + k];
        const float2 w = tw[m];
        ax = fmaf(v.x, w.x, ax); ax = fmaf(-v.y, w.y, ax);
        ay = fmaf(v.x, w.y, ay); ay = fmaf(v.y, w.x, ay);
        m += fp;
        if (m >= F) m -= F;
    }
    const int f = (fp + F / 2) % F;
    out[(size_t)f * (R + 1) + k] = make_float2(ax, ay);
}

// ------------------------------------------------------------------------------------------
// small element-wise helpers
// ------------------------------------------------------------------------------------------
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

}  // namespace prc
