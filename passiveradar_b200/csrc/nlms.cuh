// nlms.cuh -- NLMS / block-NLMS clutter canceller (reference clutter_removal.py:189-249).
//
// The recurrence carries the tap vector from one sample to the next, so the sample loop
// is serial by construction; version 1 keeps the whole filter inside ONE persistent CTA:
// thread j owns tap j (taps in registers), the reference window lives in a shared-memory
// tile, and each sample costs one warp-shuffle + one cross-warp reduction.
//   u_k[j] = ref[M + k - j]                      (clutter_removal.py:228,237)
//   e_k    = srv[k + filterLen] - w^H u_k        (:213,:241)
//   w     += mu * u_k * conj(e_k) / (u_k^H u_k)  (:214)      [applied every block_len samples]
#pragma once
#include <cuda_runtime.h>

namespace prc {

struct NlmsParams {
    const float2* ref;
    const float2* srv;
    const float2* init;      // M initial taps or nullptr
    float2* out;             // n
    float2* taps_out;        // M or nullptr
    int n;
    int filter_len;
    int peek;
    float mu;
    int block_len;
    // a batch of independent frames, one CTA each (gridDim.x): frame f reads ref/srv + f * frame_stride, writes
    // out + f * frame_stride; init / taps_out advance by filter_len + peek taps per frame (init_shared: every frame
    // starts from the same init taps)
    long long frame_stride;
    int init_shared;
};

__device__ __forceinline__ NlmsParams nlms_frame_params(const NlmsParams& g) {
    NlmsParams p = g;
    const size_t f = blockIdx.x;
    const int M = g.filter_len + g.peek;
    p.ref += f * g.frame_stride;
    p.srv += f * g.frame_stride;
    p.out += f * g.frame_stride;
    if (p.init && !g.init_shared) p.init += f * M;
    if (p.taps_out) p.taps_out += f * M;
    return p;
}

constexpr int NLMS_TILE = 1024;      // samples staged per shared-memory tile
constexpr int NLMS_MAXT = 4;         // max taps per thread (M <= 4096)

template <int KT>
__global__ void __launch_bounds__(1024) nlms_kernel(const __grid_constant__ NlmsParams pg) {
    extern __shared__ __align__(16) float2 nsm[];
    const NlmsParams p = nlms_frame_params(pg);
    const int M = p.filter_len + p.peek;
    const int nsteps = p.n - M;
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    float2* tile = nsm;                               // NLMS_TILE + M window of ref
    float2* dtile = nsm + NLMS_TILE + ((M + 1) & ~1);  // NLMS_TILE desired samples (16 B aligned)
    float4* red = reinterpret_cast<float4*>(dtile + NLMS_TILE);   // [2][32] (re, im, |u|^2, -)

    // output head / tail stay zero (clutter_removal.py:231)
    for (int i = tid; i < p.n; i += nthr)
        if (i < p.filter_len || i >= p.filter_len + (nsteps > 0 ? nsteps : 0)) p.out[i] = make_float2(0.f, 0.f);

    float2 w[KT], grad[KT];
#pragma unroll
    for (int r = 0; r < KT; ++r) {
        const int j = tid + r * nthr;
        w[r] = (p.init && j < M) ? p.init[j] : make_float2(0.f, 0.f);
        grad[r] = make_float2(0.f, 0.f);
    }
    int in_block = 0;
    int buf = 0;
    for (int ts = 0; ts < nsteps; ts += NLMS_TILE) {
        const int tl = min(NLMS_TILE, nsteps - ts);
        __syncthreads();
        // tile[q] = ref[ts + 1 + q]; u_k[j] = tile[M - 1 - j + (k - ts)]
        for (int q = tid; q < tl + M - 1; q += nthr) tile[q] = p.ref[ts + 1 + q];
        for (int q = tid; q < tl; q += nthr) dtile[q] = p.srv[p.filter_len + ts + q];
        __syncthreads();
        for (int kk = 0; kk < tl; ++kk) {
            float2 u[KT];
            float dr = 0.f, di = 0.f, nn = 0.f;
#pragma unroll
            for (int r = 0; r < KT; ++r) {
                const int j = tid + r * nthr;
                u[r] = (j < M) ? tile[M - 1 - j + kk] : make_float2(0.f, 0.f);
                // conj(w) * u
                dr = fmaf(w[r].x, u[r].x, dr); dr = fmaf(w[r].y, u[r].y, dr);
                di = fmaf(w[r].x, u[r].y, di); di = fmaf(-w[r].y, u[r].x, di);
                nn = fmaf(u[r].x, u[r].x, nn); nn = fmaf(u[r].y, u[r].y, nn);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dr += __shfl_xor_sync(0xffffffffu, dr, o);
                di += __shfl_xor_sync(0xffffffffu, di, o);
                nn += __shfl_xor_sync(0xffffffffu, nn, o);
            }
            float4* rb = red + buf * 32;
            if (lane == 0) rb[warp] = make_float4(dr, di, nn, 0.f);
            __syncthreads();
            float sr = 0.f, si = 0.f, sn = 0.f;
            for (int q = 0; q < nwarp; ++q) {
                const float4 v = rb[q];
                sr += v.x; si += v.y; sn += v.z;
            }
            buf ^= 1;
            const float2 d = dtile[kk];
            const float er = d.x - sr, ei = d.y - si;
            if (tid == 0) p.out[p.filter_len + ts + kk] = make_float2(er, ei);
            // grad += mu * u * conj(e) / |u|^2
            const float sc = p.mu / sn;
            const float cr = er * sc, ci = -ei * sc;
#pragma unroll
            for (int r = 0; r < KT; ++r) {
                grad[r].x = fmaf(u[r].x, cr, grad[r].x); grad[r].x = fmaf(-u[r].y, ci, grad[r].x);
                grad[r].y = fmaf(u[r].x, ci, grad[r].y); grad[r].y = fmaf(u[r].y, cr, grad[r].y);
            }
            if (++in_block == p.block_len || ts + kk == nsteps - 1) {
#pragma unroll
                for (int r = 0; r < KT; ++r) {
                    w[r].x += grad[r].x; w[r].y += grad[r].y;
                    grad[r] = make_float2(0.f, 0.f);
                }
                in_block = 0;
            }
        }
    }
    if (p.taps_out) {
#pragma unroll
        for (int r = 0; r < KT; ++r) {
            const int j = tid + r * nthr;
            if (j < M) p.taps_out[j] = w[r];
        }
    }
}

}  // namespace prc
