// Probe: one tcgen05.mma.kind::tf32 on thread-filled shared memory, dump the accumulator.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "toepcorr.cuh"
using namespace prc::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

// mode 0: all ones; mode 1: A[m][k] = m + 100k (small ints), B[n][k] = (k == kb) pattern -> checks layout
__global__ void probe(float* out, uint32_t* info, int mode, uint32_t idesc_override, uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    float* A = reinterpret_cast<float*>(sm);             // canonical MN-major: chunk c (4 m) at c*128 B, k row r at r*16 B
    float* B = reinterpret_cast<float*>(sm + 8192);
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        const int m = i % 128, k = i / 128;
        const float v = mode == 0 ? 1.f : (float)(m + 1);       // A[m][k] = m+1
        A[(m / 4) * 32 + k * 4 + (m % 4)] = v;
    }
    for (int i = tid; i < 256 * 8; i += blockDim.x) {
        const int n = i % 256, k = i / 256;
        const float v = mode == 0 ? 1.f : ((k == (n % 8)) ? 1.f : 0.f);   // B[n][k] = delta(k, n%8)
        B[(n / 4) * 32 + k * 4 + (n % 4)] = v;
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) tmem_alloc(&tbase, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (tid == 0) {
        info[0] = tmem;
        const uint32_t idesc = idesc_override ? idesc_override : make_idesc(128, 256);
        info[1] = idesc;
        const uint64_t da = make_desc(smem_u32(A), lbo, sbo), db = make_desc(smem_u32(B), lbo, sbo);
        info[2] = (uint32_t)da; info[3] = (uint32_t)(da >> 32);
        umma_tf32(tmem, da, db, idesc, 0);
        umma_commit(&bar);
    }
    if (tid < 128) {
        mbar_wait(&bar, 0);
        tc_fence_after();
        const int warp = tid >> 5, lane = tid & 31;
        for (int j0 = 0; j0 < 256; j0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + j0, v);
            for (int jj = 0; jj < 32; ++jj) out[(warp * 32 + lane) * 256 + j0 + jj] = __uint_as_float(v[jj]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (tid < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
    float* d; uint32_t* info;
    CK(cudaMalloc(&d, 128 * 256 * 4)); CK(cudaMalloc(&info, 64));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    std::vector<float> h(128 * 256);
    uint32_t hi[4];
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(d, 0xff, 128 * 256 * 4));
        probe<<<1, 160, 32768>>>(d, info, mode, 0, 4096, 128);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hi, info, 16, cudaMemcpyDeviceToHost));
        printf("mode %d: tmem=0x%08x idesc=0x%08x desc=0x%08x%08x\n", mode, hi[0], hi[1], hi[3], hi[2]);
        printf("  D[0][0..7]  ="); for (int j = 0; j < 8; ++j) printf(" %g", h[j]); printf("\n");
        printf("  D[1][0..7]  ="); for (int j = 0; j < 8; ++j) printf(" %g", h[256 + j]); printf("\n");
        printf("  D[5][8..15] ="); for (int j = 8; j < 16; ++j) printf(" %g", h[5 * 256 + j]); printf("\n");
        printf("  D[127][248..255] ="); for (int j = 248; j < 256; ++j) printf(" %g", h[127 * 256 + j]); printf("\n");
        if (mode == 1) {   // expected D[m][n] = sum_k A[m][k] B[n][k] = (m+1)
            int bad = 0; for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) if (fabsf(h[m * 256 + n] - (m + 1)) > 1e-3f) ++bad;
            printf("  mode 1 mismatches: %d of %d\n", bad, 128 * 256);
        }
    }
    return 0;
}
