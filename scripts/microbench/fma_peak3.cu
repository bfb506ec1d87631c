// Microbenchmark 3: cost of shared-memory loads and register moves interleaved with the FFMA2
// stream of slide_mac2 (200 FFMA2 per iteration).  Prints achieved FMA/clk/SM (peak 128).
#include <cstdio>
#include <cuda_runtime.h>

// NOWN: own-lane LDS.128 per iteration (lane stride 80 B), NBC: broadcast LDS.128, NMOV: dup moves
template <int NOWN, int NBC, int NMOV, int W64>
__global__ void k(float* out, int iters, float a) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float2 acc[10], acc2[10], W[20];
#pragma unroll
    for (int i = 0; i < 10; ++i) { acc[i] = make_float2(threadIdx.x * 1e-9f + i, i * 0.5f); acc2[i] = acc[i]; }
#pragma unroll
    for (int i = 0; i < 20; ++i) W[i] = make_float2(1.0f + i * 1e-3f, 0.5f - i * 1e-3f);
    float4 xb = make_float4(a, a, a, a);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 10; ++u) {
            if (u < NOWN) {
                if (W64) {
                    const float2* s2 = reinterpret_cast<const float2*>(sm);
                    W[(u + 10) % 20] = s2[(threadIdx.x * 10 + it + 2 * u) & 4095];
                    W[(u + 11) % 20] = s2[(threadIdx.x * 10 + it + 2 * u + 1) & 4095];
                } else {
                    float4 v = sm[(threadIdx.x * 5 + it + u) & 2047];
                    W[(u + 10) % 20] = make_float2(v.x, v.y);
                    W[(u + 11) % 20] = make_float2(v.z, v.w);
                }
            }
            if (u < NBC) xb = sm[(it + u) & 2047];
            float2 xr, xi;
            if (u < NMOV) { xr = make_float2(xb.x + 0.f, xb.x); xi = make_float2(xb.y, xb.y); }
            else { xr = make_float2(xb.x, xb.y); xi = make_float2(xb.z, xb.w); }
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

template <int NOWN, int NBC, int NMOV, int W64>
void run(float* d, const char* name) {
    const int iters = 2048;
    for (int warps = 8; warps <= 16; warps *= 2) {
        float h;
        k<NOWN, NBC, NMOV, W64><<<1, warps * 32>>>(d, iters, 1.0001f);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        double fma = (double)iters * 10 * 20 * 2 * warps * 32;
        printf("%-34s warps=%2d FMA/clk/SM=%6.1f  cycles/iter/warp-slot=%.1f\n", name, warps, fma / h, h / iters / (warps / 4));
    }
}

int main() {
    float* d;
    cudaMalloc(&d, 1 << 20);
    run<0, 0, 0, 0>(d, "pure FFMA2");
    run<5, 0, 0, 0>(d, "5 own LDS.128");
    run<10, 0, 0, 0>(d, "10 own LDS.128");
    run<0, 5, 0, 0>(d, "5 bcast LDS.128");
    run<0, 10, 0, 0>(d, "10 bcast LDS.128");
    run<5, 5, 0, 0>(d, "5 own + 5 bcast LDS.128");
    run<5, 5, 10, 0>(d, "5 own + 5 bcast + 10 dup");
    run<5, 0, 0, 1>(d, "10 own LDS.64 (= 5 x 128 bytes)");
    return 0;
}
