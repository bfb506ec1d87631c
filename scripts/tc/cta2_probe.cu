// cta2_probe.cu -- minimal tcgen05.mma.cta_group::2 bring-up on B200: D[256 x 128] = A[256 x 64] * B[128 x 64]^T,
// BF16 K-major SWIZZLE_128B operands, a cluster of two CTAs: each CTA holds 128 rows of A and 64 of the 128
// rows (N) of B, the leader (cluster rank 0) issues the MMAs, the commit is multicast to both CTAs, each CTA
// reads its 128 lanes of the accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../passiveradar_b200/csrc -o cta2_probe cta2_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "firtc.cuh"
using namespace prc::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

constexpr int K = 64;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = cluster_rank();
    uint8_t* sa = sm;                 // 128 rows x 128 B
    uint8_t* sb = sm + 16384;         // 64 rows x 128 B
    const int tid = threadIdx.x;
    // A rows [128 rank, +128), B rows (N) [64 rank, +64): 16-byte chunks, chunk ^= row & 7
    for (int q = tid; q < 128 * 8; q += blockDim.x) {
        const int row = q >> 3, ch = q & 7;
        *reinterpret_cast<uint4*>(sa + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(A + (size_t)(128 * rank + row) * K + 8 * ch);
    }
    for (int q = tid; q < 64 * 8; q += blockDim.x) {
        const int row = q >> 3, ch = q & 7;
        *reinterpret_cast<uint4*>(sb + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(B + (size_t)(64 * rank + row) * K + 8 * ch);
    }
    if (tid == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_proxy_async();
    __syncthreads();
    if (tid < 32) tmem_alloc2(&tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();               // both CTAs: operands in shared memory, barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = make_idesc_bf16_kmajor(256, 128);
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t da = make_desc_k_sw128(smem_u32(sa) + kk * 32, 1024);
            const uint64_t db = make_desc_k_sw128(smem_u32(sb) + kk * 32, 1024);
            umma2_f16(tmem, da, db, idesc, kk > 0);
        }
        umma2_commit_mc(&done_bar, 0x3);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j0 = 0; j0 < 128; j0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + j0, v);
        for (int jj = 0; jj < 32; ++jj) D[(size_t)(128 * rank + warp * 32 + lane) * 128 + j0 + jj] = __uint_as_float(v[jj]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (tid < 32) { tc_fence_after(); tmem_dealloc2(tmem, 128); }
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<uint16_t> A(256 * K), B(128 * K);
    srand(1);
    for (auto& v : A) v = f2bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = f2bf((rand() % 2001 - 1000) / 1000.f);
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * 128 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 256 * 128 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<2, 128, 32768>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(256 * 128);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxval = 0;
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < 128; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)bf2f(A[m * K + k]) * bf2f(B[n * K + k]);
            maxerr = fmax(maxerr, fabs(s - D[m * 128 + n]));
            maxval = fmax(maxval, fabs(s));
        }
    printf("cta_group::2 M=256 N=128 K=64: max|err| = %.3e, max|val| = %.3e -> %s\n", maxerr, maxval, maxerr < 1e-3 * maxval ? "OK" : "WRONG");
    printf("D[0][0..3] = %f %f %f %f; D[128][0] = %f; D[255][127] = %f\n", D[0], D[1], D[2], D[3], D[128 * 128], D[255 * 128 + 127]);
    return 0;
}
