// frontend.cuh -- the per-chunk front end that feeds the hot path:
//   deinterleave_IQ  (reference passiveRadar/signal_utils.py:19-22)
//   frequency_shift  (signal_utils.py:24-27, float32 phase ramp reproduced)
//   resample         (signal_utils.py:15-17: scipy resample_poly(x, up, dn, padtype='line'))
// and the CFAR detector behind it (detect.cuh).  All of it is bandwidth-bound element-wise / short-FIR
// work: one pass over the raw samples, coalesced 128-bit loads, shared-memory staging of the FIR window.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prc {

enum IqKind { IQ_C64 = 0, IQ_I8 = 1, IQ_I16 = 2 };
enum MixMode { MIX_NONE = 0, MIX_C64 = 1, MIX_C128 = 2, MIX_F64 = 3, MIX_FS64 = 4 };

struct MixParams {
    const void* in;
    int kind;            // IqKind
    long long n;         // complex samples
    int mode;            // MixMode
    float B;             // float32(2 pi fc): imaginary part of complex64(1j*2*pi*fc)
    float rFs;           // fl32(1 / float32(Fs)): numpy divides a complex array by a real scalar as a
                         // multiplication by the reciprocal (loops.c.src: scl = 1.0/(in2r + in2i*rat))
    double po;           // phase offset (MIX_C64: rounded to float32 first, like numpy's weak scalar)
    double B64, rFs64;   // MIX_F64 / MIX_FS64: the same in float64 (fc or Fs was a NumPy float64 scalar)
};

__device__ __forceinline__ float2 iq_load(const void* in, int kind, long long i) {
    if (kind == IQ_I8) {
        const char2 v = reinterpret_cast<const char2*>(in)[i];
        return make_float2((float)v.x, (float)v.y);
    }
    if (kind == IQ_I16) {
        const short2 v = reinterpret_cast<const short2*>(in)[i];
        return make_float2((float)v.x, (float)v.y);
    }
    return reinterpret_cast<const float2*>(in)[i];
}

// x[i] * exp(1j * theta_i) with numpy's evaluation order:
//   theta32 = fl32(fl32(B * fl32(i)) * fl32(1/Fs))                                  (complex64 arange, weak python scalars)
//   MIX_C64 : theta = fl32(theta32 + fl32(po)), exp and product in float32     (phase_offset a python scalar)
//   MIX_C128: theta = double(theta32) + po, exp in float64                     (phase_offset a numpy float64, main.py:127-130)
//   MIX_F64 : theta = (2 pi fc) * double(fl32(i)) / Fs + po in float64          (fc a numpy float64 scalar)
//   MIX_FS64: theta = double(fl32(B * fl32(i))) / Fs + po in float64            (only Fs a numpy float64 scalar)
// The float64 angle (up to ~1e7 rad) is reduced to [-pi, pi] in double (two-constant Cody-Waite) and the
// sine/cosine taken in float32: absolute error ~1e-7, far inside the 1e-5 parity budget.
__device__ __forceinline__ float2 mix_sample(const MixParams& p, float2 v, long long i) {
    if (p.mode == MIX_NONE) return v;
    const float th32 = __fmul_rn(__fmul_rn(p.B, (float)i), p.rFs);
    double a;
    if (p.mode == MIX_C64) a = (double)__fadd_rn(th32, (float)p.po);
    else if (p.mode == MIX_C128) a = (double)th32 + p.po;
    else if (p.mode == MIX_F64) a = (p.B64 * (double)(float)i) * p.rFs64 + p.po;   // the ramp itself is still arange(dtype=complex64)
    else a = (double)__fmul_rn(p.B, (float)i) * p.rFs64 + p.po;
    const double k = rint(a * 0.15915494309189535);
    double r = fma(-k, 6.283185307179586, a);
    r = fma(-k, 2.4492935982947064e-16, r);
    float sn, cs;
    __sincosf((float)r, &sn, &cs);        // |r| <= pi: absolute error < 5e-7 (CUDA C Programming Guide, intrinsic functions)
    return make_float2(__fmaf_rn(v.x, cs, -__fmul_rn(v.y, sn)), __fmaf_rn(v.x, sn, __fmul_rn(v.y, cs)));
}

__global__ void iq_mix_kernel(const __grid_constant__ MixParams p, float2* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    out[i] = mix_sample(p, iq_load(p.in, p.kind, i), i);
}

// ---------------------------------------------------------------------------------------------------
// resample_poly(x, up, down, padtype='line') as one polyphase pass.
//   y[m] = sum_k hpad[k] * xup[(m + n_pre_remove) * down - k],   xup[i * up] = xext[i]
// hpad = n_pre_pad zeros + h (already scaled by up), xext = x extended by a line through its end points
// (scipy _upfirdn_apply.pyx MODE_LINE: x[0] + idx * slope on the left, x[n-1] + (idx - n + 1) * slope on
// the right, slope = (x[n-1] - x[0]) / (n - 1)).  One thread per output sample; a CTA stages the
// polyphase taps [phase][j] and its window of (optionally mixed, optionally deinterleaved) input samples
// in shared memory, so the raw samples are read from HBM exactly once.
// ---------------------------------------------------------------------------------------------------
struct ResampleParams {
    MixParams mix;        // input description (mode MIX_NONE for plain resample)
    int up, down;
    int nh;               // taps in h (2 * half_len + 1)
    int n_pre_pad, n_pre_remove;
    int tpp;              // taps per phase = ceil((n_pre_pad + nh) / up), made odd (table row stride)
    long long n_out;
    const float* hp;      // polyphase table [up][tpp]: hp[ph][j] = hpad[ph + j * up]
    float2 x0, x1;        // mixed end points x[0], x[n-1] (line extension); filled by resample_ends_kernel
    const float2* ends;   // device copy of {x0, x1}
    int span;             // input samples staged per CTA
};

constexpr int RS_THREADS = 256;

__global__ void resample_ends_kernel(const __grid_constant__ MixParams p, float2* __restrict__ ends) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ends[0] = mix_sample(p, iq_load(p.in, p.kind, 0), 0);
        ends[1] = mix_sample(p, iq_load(p.in, p.kind, p.n - 1), p.n - 1);
    }
}

__global__ void __launch_bounds__(RS_THREADS) resample_kernel(const __grid_constant__ ResampleParams p, float2* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    float* hp = reinterpret_cast<float*>(rs_smem);                               // [up][tpp]
    float2* xs = reinterpret_cast<float2*>(rs_smem + (((size_t)p.up * p.tpp * 4 + 15) & ~(size_t)15));
    const long long m0 = (long long)blockIdx.x * RS_THREADS;
    const long long n = p.mix.n;
    // input index range of this CTA: output m uses i = floor(T / up) - j, j = 0 .. tpp-1, T = (m + npr) * down
    const long long Tlo = (m0 + p.n_pre_remove) * (long long)p.down;
    const long long i_hi0 = Tlo / p.up;                      // largest index of the first output
    const long long i_lo = i_hi0 - (p.tpp - 1);              // smallest index any output of the CTA needs
    for (int q = threadIdx.x; q < p.up * p.tpp; q += RS_THREADS) hp[q] = p.hp[q];
    const float2 x0 = p.ends[0], x1 = p.ends[1];
    float2 slope = make_float2(0.f, 0.f);
    if (n > 1) slope = make_float2((x1.x - x0.x) / (float)(n - 1), (x1.y - x0.y) / (float)(n - 1));
    for (int q = threadIdx.x; q < p.span; q += RS_THREADS) {
        const long long i = i_lo + q;
        float2 v;
        if (i < 0) v = make_float2(x0.x + (float)i * slope.x, x0.y + (float)i * slope.y);
        else if (i >= n) v = make_float2(x1.x + (float)(i - n + 1) * slope.x, x1.y + (float)(i - n + 1) * slope.y);
        else v = mix_sample(p.mix, iq_load(p.mix.in, p.mix.kind, i), i);
        xs[q] = v;
    }
    __syncthreads();
    const long long m = m0 + threadIdx.x;
    if (m >= p.n_out) return;
    const long long T = (m + p.n_pre_remove) * (long long)p.down;
    const long long ih = T / p.up;
    const int ph = (int)(T - ih * p.up);                     // hpad index of the newest sample: k = ph + j * up
    const float* h = hp + (size_t)ph * p.tpp;
    const float2* x = xs + (ih - i_lo);                      // x[-j] = xext[ih - j]
    float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;            // two chains: shorter dependency, pairwise-like sum
    int j = 0;
    for (; j + 1 < p.tpp; j += 2) {
        const float h0 = h[j], h1 = h[j + 1];
        const float2 v0 = x[-j], v1 = x[-j - 1];
        ar = fmaf(h0, v0.x, ar); ai = fmaf(h0, v0.y, ai);
        br = fmaf(h1, v1.x, br); bi = fmaf(h1, v1.y, bi);
    }
    if (j < p.tpp) {
        const float h0 = h[j];
        const float2 v0 = x[-j];
        ar = fmaf(h0, v0.x, ar); ai = fmaf(h0, v0.y, ai);
    }
    out[m] = make_float2(ar + br, ai + bi);
}

}  // namespace prc
