// fftcorr.cuh -- the frame's three heavy steps as block correlations / overlap-save filtering in the frequency
// domain (FP32, fftcore.cuh), reading the caller's complex64 channels directly: no intermediate planes in HBM.
//
//   lscorr_fft_kernel   c[m] = sum_i ref[i] conj(ref[i+m]),  x[m] = sum_i ref[i] conj(srv[i+m-peek]),  m < M
//                       (first column of A^H A and A^H srv of LS_Filter, reference clutter_removal.py:34-45, up to the
//                       conjugation levinson_kernel applies).  The channel is cut into nb zero-padded blocks of
//                       Bu <= L/2 samples; with X_b, S_b their length-L spectra the sum over blocks of
//                       X_b conj(X_b + e_b X_{b+1}) (e_b = shift by the block length) is the spectrum of the lag
//                       sums for 0 <= m < M <= L/2 -- two transforms per block, spectra accumulated in registers,
//                       one permuted->natural transform per correlation and CTA at the end.
//   taps_spectrum_kernel  W'(f): spectrum of the solved taps placed at (k - (M-1)) mod L, in permuted order.
//   fir_fft_kernel      out = srv - circular FIR(ref, w)  (clutter_removal.py:51) by overlap-save: one forward
//                       transform of L reference samples, times W', one transform back, L - M + 1 outputs.
//   caf_fft_kernel      CAF block sums P[j][d] = sum_{i in Doppler block j} ref[i] win[i] conj(s[(i+d) mod n]),
//                       d = 0..R (range_doppler_processing.py:81-86), s = srv or -- FUSED -- the cleaned channel
//                       formed in the frequency domain as S - R_seg W' (valid because the circular-convolution
//                       wrap lands in the part of the segment no lag 0..R touches).  Segments of <= L - R - (M-1)
//                       samples; spectra X conj(S_clean) accumulated per Doppler block; one transform back.
//
// All kernels take a batch of frames in gridDim.y (frame_stride samples apart).  Algebra pinned on the CPU by
// scripts/fft/model.py (numpy) and scripts/fft/emul.cu (thread-by-thread run of fftcore.cuh).
#pragma once
#include <cuda_runtime.h>

#include "fftcore.cuh"

namespace prc {
namespace fftc {

using fft::cmul;
using fft::cmulc;

// idx may be up to one period outside [0, n)
__device__ __forceinline__ float2 load_sig(const float2* __restrict__ sig, long long idx, int n, int linear) {
    if (idx < 0) {
        if (linear) return make_float2(0.f, 0.f);
        idx += n;
    } else if (idx >= n) {
        if (linear) return make_float2(0.f, 0.f);
        idx -= n;
    }
    return __ldg(sig + idx);
}

// exp(-2 pi i f len / L) for register r of thread t (exact argument reduction in integers)
template <int R3>
__device__ __forceinline__ float2 shift_tw(int t, int r, int len) {
    constexpr int L = fft::Geo<R3>::L;
    const int f = fft::perm_freq<R3>(t, r);
    const int q = (int)(((long long)f * len) & (L - 1));
    float s, c;
    sincospif(-2.0f * (float)q / (float)L, &s, &c);
    return make_float2(c, s);
}

// ------------------------------------------------------------------------------------------------ LS correlations
struct LsCorrParams {
    const float2* ref;
    const float2* srv;
    long long frame_stride;
    int n, M, peek, linear;
    int nb, Bu, last;          // nb blocks of Bu samples, the last one `last` samples
    int bpc;                   // blocks per CTA
    float2* partial;           // [frame][2][gridDim.x][HT]
    int HT;
    const float2* tw;
};

template <int R3>
__global__ void __launch_bounds__(16 * R3) lscorr_fft_kernel(const __grid_constant__ LsCorrParams p) {
    using G = fft::Geo<R3>;
    extern __shared__ __align__(16) float2 sm[];
    const int t = threadIdx.x;
    const fft::Smem<R3> S(sm);
    fft::stage_twiddles<R3>(sm, p.tw, t);
    __syncthreads();
    const float2* ref = p.ref + (size_t)blockIdx.y * p.frame_stride;
    const float2* srv = p.srv + (size_t)blockIdx.y * p.frame_stride;
    const int b0 = blockIdx.x * p.bpc;
    const int b1 = min(b0 + p.bpc, p.nb);
    // a previous block of exactly L/2 samples: e_b = (-1)^f = (-1)^k1, one sign per thread
    const float sgn = ((t / R3) & 1) ? -1.f : 1.f;

    float2 accC[16], accX[16], Xp[16], Xn[16], v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) accC[r] = accX[r] = Xp[r] = make_float2(0.f, 0.f);
    int prev_len = 0;
    for (int b = b0; b <= b1; ++b) {                       // b == b1: look-ahead only (cross terms of block b1 - 1)
        const bool owned = b < b1;
        long long start;
        int ln;
        if (b < p.nb) { start = (long long)b * p.Bu; ln = (b == p.nb - 1) ? p.last : p.Bu; }
        else if (p.linear) { start = p.n; ln = p.Bu; }     // what follows the channel: zeros (and the peek samples of srv)
        else { start = 0; ln = p.Bu; }                     // circular: block 0 follows the last block
        // ---- reference block
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int i = n1 * G::T + t;
            Xn[n1] = (i < ln) ? load_sig(ref, start + i, p.n, p.linear) : make_float2(0.f, 0.f);
        }
        // surveillance block (shifted by -peek): issue the loads before the first transform
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int i = n1 * G::T + t;
            v[n1] = (i < ln) ? load_sig(srv, start + i - p.peek, p.n, p.linear) : make_float2(0.f, 0.f);
        }
        fft::fft_n2p<R3>(Xn, t, S);
        fft::fft_n2p<R3>(v, t, S);
        const bool has_prev = b > b0;
        const bool half = (prev_len * 2 == G::L);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            // Q = Xp conj(e) [prev block's cross term] + Xn [own term];  accC += Q conj(Xn), accX += Q conj(S)
            float2 q = make_float2(0.f, 0.f);
            if (has_prev) {
                if (half) q = make_float2(sgn * Xp[r].x, sgn * Xp[r].y);
                else q = cmulc(Xp[r], shift_tw<R3>(t, r, prev_len));
            }
            if (owned) { q.x += Xn[r].x; q.y += Xn[r].y; }
            const float2 zc = cmulc(q, Xn[r]);
            const float2 zx = cmulc(q, v[r]);
            accC[r].x += zc.x; accC[r].y += zc.y;
            accX[r].x += zx.x; accX[r].y += zx.y;
            Xp[r] = Xn[r];
        }
        prev_len = ln;
    }
    const float inv = 1.0f / (float)G::L;
    float2* row = p.partial + ((size_t)(blockIdx.y * 2 + 0) * gridDim.x + blockIdx.x) * p.HT;
    fft::fft_p2n<R3>(accC, t, S);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int m = n1 * G::T + t;
        if (m < p.M) row[m] = make_float2(accC[n1].x * inv, accC[n1].y * inv);
    }
    row += (size_t)gridDim.x * p.HT;
    fft::fft_p2n<R3>(accX, t, S);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int m = n1 * G::T + t;
        if (m < p.M) row[m] = make_float2(accX[n1].x * inv, accX[n1].y * inv);
    }
}

// ------------------------------------------------------------------------------------------------ taps spectrum
struct TapSpecParams {
    const float2* taps;        // [frame][M]
    int M;
    float2* wp;                // [frame][L], permuted order: wp[r * T + t]
    const float2* tw;
};

template <int R3>
__global__ void __launch_bounds__(16 * R3) taps_spectrum_kernel(const __grid_constant__ TapSpecParams p) {
    using G = fft::Geo<R3>;
    extern __shared__ __align__(16) float2 sm[];
    const int t = threadIdx.x;
    const fft::Smem<R3> S(sm);
    fft::stage_twiddles<R3>(sm, p.tw, t);
    __syncthreads();
    const float2* taps = p.taps + (size_t)blockIdx.x * p.M;
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int k = (n1 * G::T + t + p.M - 1) & (G::L - 1);      // h[(k - (M-1)) mod L] = w[k]
        v[n1] = (k < p.M) ? taps[k] : make_float2(0.f, 0.f);
    }
    fft::fft_n2p<R3>(v, t, S);
    float2* wp = p.wp + (size_t)blockIdx.x * G::L;
#pragma unroll
    for (int r = 0; r < 16; ++r) wp[r * G::T + t] = v[r];
}

// ------------------------------------------------------------------------------------------------ overlap-save FIR
struct FirFftParams {
    const float2* ref;
    const float2* srv;
    float2* out;
    long long frame_stride;
    const float2* wp;          // [frame][L]
    int n, M, peek, linear;
    int nseg;                  // segments of L - M + 1 outputs; CTA walks segments blockIdx.x, + gridDim.x, ...
    const float2* tw;
};

template <int R3>
__global__ void __launch_bounds__(16 * R3) fir_fft_kernel(const __grid_constant__ FirFftParams p) {
    using G = fft::Geo<R3>;
    extern __shared__ __align__(16) float2 sm[];
    const int t = threadIdx.x;
    const fft::Smem<R3> S(sm);
    fft::stage_twiddles<R3>(sm, p.tw, t);
    __syncthreads();
    const float2* ref = p.ref + (size_t)blockIdx.y * p.frame_stride;
    const float2* srv = p.srv + (size_t)blockIdx.y * p.frame_stride;
    float2* out = p.out + (size_t)blockIdx.y * p.frame_stride;
    const float2* wp = p.wp + (size_t)blockIdx.y * G::L;
    const int Bf = G::L - p.M + 1;
    const float inv = 1.0f / (float)G::L;
    for (int sgm = blockIdx.x; sgm < p.nseg; sgm += gridDim.x) {
        const long long p0 = (long long)sgm * Bf;
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1)
            v[n1] = load_sig(ref, p0 + p.peek - (p.M - 1) + n1 * G::T + t, p.n, p.linear);
        fft::fft_n2p<R3>(v, t, S);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float2 y = cmul(v[r], __ldg(wp + r * G::T + t));
            v[r] = make_float2(y.x, -y.y);                   // transform of conj(Y) = L conj(y)
        }
        fft::fft_p2n<R3>(v, t, S);
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int i = n1 * G::T + t;
            const long long o = p0 + i;
            if (i < Bf && o < p.n) {
                const float2 s = __ldg(srv + o);
                out[o] = make_float2(s.x - v[n1].x * inv, s.y + v[n1].y * inv);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ CAF block sums
struct CafFftParams {
    const float2* ref;
    const float2* srv;
    long long frame_stride;
    const float* win;          // n floats shared by all frames, or NULL
    const float2* wp;          // [frame][L] taps spectrum (FUSED) or NULL
    int n, R, F, M, peek;      // M, peek: the fused clutter filter
    long long D;               // Doppler block j sums samples i = j D + c0 - m, m = 0..ntaps-1, 0 <= i < n
    int c0, ntaps;
    int Bmax;                  // segment length limit: L - R - (M - 1)
    float2* P;                 // [frame][F][HT]; P[j][d], d = 0..R
    int HT;
    const float2* tw;
};

template <int R3, bool FUSED>
__global__ void __launch_bounds__(16 * R3) caf_fft_kernel(const __grid_constant__ CafFftParams p) {
    using G = fft::Geo<R3>;
    extern __shared__ __align__(16) float2 sm[];
    const int t = threadIdx.x;
    const fft::Smem<R3> S(sm);
    fft::stage_twiddles<R3>(sm, p.tw, t);
    __syncthreads();
    const float2* ref = p.ref + (size_t)blockIdx.y * p.frame_stride;
    const float2* srv = p.srv + (size_t)blockIdx.y * p.frame_stride;
    const float2* wp = FUSED ? p.wp + (size_t)blockIdx.y * G::L : nullptr;
    const float inv = 1.0f / (float)G::L;
    for (int j = blockIdx.x; j < p.F; j += gridDim.x) {
        long long lo = (long long)j * p.D + p.c0 - (p.ntaps - 1);
        long long hi = (long long)j * p.D + p.c0 + 1;
        if (lo < 0) lo = 0;
        if (hi > p.n) hi = p.n;
        float2 acc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
        if (hi > lo) {
            const int len = (int)(hi - lo);
            const int nseg = (len + p.Bmax - 1) / p.Bmax;
            const int Bs = (len + nseg - 1) / nseg;
            for (int q = 0; q < nseg; ++q) {
                const long long i0 = lo + (long long)q * Bs;
                const int ln = (int)min((long long)Bs, hi - i0);
                float2 sg[16], v[16];
                // surveillance segment: L samples from i0 (circular); positions >= ln + R are never used by lags 0..R
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    long long idx = i0 + n1 * G::T + t;
                    if (idx >= p.n) idx -= p.n;
                    if (idx >= p.n) idx %= p.n;              // n < L only for tiny inputs
                    sg[n1] = __ldg(srv + idx);
                }
                if (FUSED) {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        long long idx = i0 + p.peek - (p.M - 1) + n1 * G::T + t;
                        if (idx < 0) idx += p.n;
                        if (idx >= p.n) idx -= p.n;
                        if (idx < 0 || idx >= p.n) idx = ((idx % p.n) + p.n) % p.n;
                        v[n1] = __ldg(ref + idx);
                    }
                }
                fft::fft_n2p<R3>(sg, t, S);
                if (FUSED) {
                    fft::fft_n2p<R3>(v, t, S);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float2 y = cmul(v[r], __ldg(wp + r * G::T + t));
                        sg[r].x -= y.x;
                        sg[r].y -= y.y;
                    }
                }
                // x segment: ref * window, zero padded
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int i = n1 * G::T + t;
                    float2 x = make_float2(0.f, 0.f);
                    if (i < ln) {
                        x = __ldg(ref + i0 + i);
                        if (p.win) {
                            const float w = __ldg(p.win + i0 + i);
                            x.x *= w;
                            x.y *= w;
                        }
                    }
                    v[n1] = x;
                }
                fft::fft_n2p<R3>(v, t, S);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 z = cmulc(v[r], sg[r]);
                    acc[r].x += z.x;
                    acc[r].y += z.y;
                }
            }
            fft::fft_p2n<R3>(acc, t, S);
        }
        float2* row = p.P + ((size_t)blockIdx.y * p.F + j) * p.HT;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const int d = n1 * G::T + t;
            if (d <= p.R) row[d] = make_float2(acc[n1].x * inv, acc[n1].y * inv);
        }
    }
}

}  // namespace fftc
}  // namespace prc
