// Microbenchmark: peak issue rate of FFMA vs FFMA2 on one SM (run on the B200 box).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-9f + i, i * 0.5f);
    float2 x = make_float2(a, a), y = make_float2(b, 1.0f - b);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) {            // scalar FFMA, 2 per accumulator pair
                    acc[i].x = fmaf(x.x, y.x, acc[i].x);
                    acc[i].y = fmaf(x.y, y.y, acc[i].y);
                } else {                    // packed FFMA2
                    acc[i] = __ffma2_rn(x, y, acc[i]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
    float* d;
    cudaMalloc(&d, 1 << 20);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 32; warps *= 2) {
            float h;
            if (mode == 0) k<0><<<1, warps * 32>>>(d, iters, 1.0001f, 0.5f);
            else k<1><<<1, warps * 32>>>(d, iters, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
            double fma_lanes = (double)iters * 4 * 16 * 2 * warps * 32;   // scalar FMAs executed
            printf("%s warps/SM=%2d cycles=%.0f  FMA/clk/SM=%.1f\n", mode ? "FFMA2" : "FFMA ", warps, h, fma_lanes / h);
        }
    return 0;
}
