// Microbenchmark 2: FFMA2 with the operand pattern of slide_mac2 (x reused, window operand and
// accumulator change every instruction), optionally interleaved with 128-bit shared loads.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, const float4* gsrc) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float2 acc[10], acc2[10], W[20];
#pragma unroll
    for (int i = 0; i < 10; ++i) { acc[i] = make_float2(threadIdx.x * 1e-9f + i, i * 0.5f); acc2[i] = acc[i]; }
#pragma unroll
    for (int i = 0; i < 20; ++i) W[i] = make_float2(1.0f + i * 1e-3f, 0.5f - i * 1e-3f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 10; ++u) {
            float2 xr = make_float2(a + u, a + u), xi = make_float2(a - u, a - u);
            if (MODE == 1 && (u & 1) == 0) {       // one LDS.128 per two samples, like the real loop
                float4 v = sm[(threadIdx.x * 5 + it + u) & 1023];
                W[(u + 10) % 20] = make_float2(v.x, v.y);
                W[(u + 11) % 20] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int v = 0; v < 10; ++v) acc[v] = __ffma2_rn(xr, W[(u + v) % 20], acc[v]);
#pragma unroll
            for (int v = 0; v < 10; ++v) acc2[v] = __ffma2_rn(xi, W[(u + v) % 20], acc2[v]);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) s += acc[i].x + acc[i].y + acc2[i].x + acc2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
    float* d;
    cudaMalloc(&d, 1 << 20);
    const int iters = 2048;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            float h;
            if (mode == 0) k<0><<<1, warps * 32>>>(d, iters, 1.0001f, nullptr);
            else k<1><<<1, warps * 32>>>(d, iters, 1.0001f, nullptr);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
            double fma_lanes = (double)iters * 10 * 20 * 2 * warps * 32;
            printf("%s warps/SM=%2d cycles=%.0f  FMA/clk/SM=%.1f\n", mode ? "FFMA2+LDS" : "FFMA2    ", warps, h, fma_lanes / h);
        }
    return 0;
}
