// tma_probe.cu -- does cuTensorMapEncodeTiled accept OVERLAPPING rows (row stride 256 B, inner extent
// 2048 B), and does a SWIZZLE_128B box land in shared memory as chunk ^ (row & 7)?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2048) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(c1), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(sm)[i];
}

int main() {
    const int N = 1 << 16;
    std::vector<uint16_t> h(N);
    for (int i = 0; i < N; ++i) h[i] = (uint16_t)i;
    uint16_t *d, *o;
    cudaMalloc(&d, N * 2);
    cudaMalloc(&o, 2048);
    cudaMemcpy(d, h.data(), N * 2, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, (void*)enc);
    CUtensorMap tm;
    cuuint64_t dims[2] = {1024, 400};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (overlapping rows): CUresult=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int c0 = 320, c1 = 5;
    probe<<<1, 128, 4096>>>(tm, c0, c1, o);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint16_t> out(1024);
    cudaMemcpy(out.data(), o, 2048, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 16; ++row)
        for (int ch = 0; ch < 8; ++ch)
            for (int el = 0; el < 8; ++el) {
                const uint16_t want = (uint16_t)(128 * (c1 + row) + c0 + 8 * ch + el);
                const uint16_t got = out[row * 64 + ((ch ^ (row & 7)) << 3) + el];
                if (want != got && bad++ < 5) printf("row %d ch %d el %d: want %u got %u\n", row, ch, el, want, got);
            }
    printf("swizzle layout chunk^(row&7), rows 128 B apart: %s (%d mismatches)\n", bad ? "NO" : "YES", bad);
    return bad != 0;
}
