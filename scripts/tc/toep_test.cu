// Standalone bring-up test of the tcgen05 Toeplitz-GEMM lag correlation (toepcorr.cuh).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I passiveradar_b200/csrc -o scripts/tc/toep_test scripts/tc/toep_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "toepcorr.cuh"

using namespace prc::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384;       // complex samples
    const int nlag = argc > 2 ? atoi(argv[2]) : 310;
    const int ranges = argc > 3 ? atoi(argv[3]) : 2;
    const int peek = 10;
    const int nk = (n + 1023) / 1024;
    const long long nx = (long long)nk * 1024;
    const int npass = (2 * (64 + nlag) + 255) / 256;
    const long long slen = nx + (long long)npass * 128;
    const int HT = ((nlag + 9) / 10) * 10;
    printf("n=%d nlag=%d nk=%d npass=%d ranges=%d grid=%d smem=%zu\n", n, nlag, nk, npass, ranges, 2 * npass * ranges, toep_smem_bytes(HT));

    std::vector<float2> ref(n), srv(n);
    srand(1);
    for (int i = 0; i < n; ++i) {
        ref[i] = make_float2((rand() / (float)RAND_MAX - 0.5f) * 2.f, (rand() / (float)RAND_MAX - 0.5f) * 2.f);
        srv[i] = make_float2((rand() / (float)RAND_MAX - 0.5f) * 2.f + 0.7f * ref[i].x, (rand() / (float)RAND_MAX - 0.5f) * 2.f - 0.3f * ref[i].y);
    }
    float2 *dref, *dsrv, *dpart;
    uint16_t *xp[3], *s0p[3], *s1p[3];
    float* dbg;
    CK(cudaMalloc(&dref, n * sizeof(float2))); CK(cudaMalloc(&dsrv, n * sizeof(float2)));
    CK(cudaMemcpy(dref, ref.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dsrv, srv.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    for (int k = 0; k < 3; ++k) {
        CK(cudaMalloc(&xp[k], 2 * nx * 2)); CK(cudaMalloc(&s0p[k], 2 * slen * 2)); CK(cudaMalloc(&s1p[k], 2 * slen * 2));
    }
    const int rows = npass * ranges;
    CK(cudaMalloc(&dpart, (size_t)2 * rows * HT * sizeof(float2)));
    CK(cudaMalloc(&dbg, 128 * 256 * 4));
    CK(cudaMemset(dbg, 0, 128 * 256 * 4));

    prc::tc::bf16_split_kernel<<<(nx + 255) / 256, 256>>>(dref, n, 0, xp[0], xp[1], xp[2], nx, n);
    prc::tc::bf16_split_kernel<<<(slen + 255) / 256, 256>>>(dref, n, 0, s0p[0], s0p[1], s0p[2], slen, slen);
    prc::tc::bf16_split_kernel<<<(slen + 255) / 256, 256>>>(dsrv, n, -peek, s1p[0], s1p[1], s1p[2], slen, slen);
    CK(cudaDeviceSynchronize());

    ToepParams p{};
    for (int k = 0; k < 3; ++k) { p.x[k] = xp[k]; p.s[0][k] = s0p[k]; p.s[1][k] = s1p[k]; }
    p.nk = nk; p.nlag = nlag; p.npass = npass; p.ranges = ranges; p.HT = HT;
    p.partial = dpart; p.debug_tile = dbg;
    long long* dclk; CK(cudaMalloc(&dclk, 512)); CK(cudaMemset(dclk, 0, 512)); p.debug_clk = dclk;
    if (argc > 4 && atoi(argv[4]) > 0) {
        // block mode (CAF-like): x = ref planes, s = srv planes, nblk blocks of kb K-steps, persistent grid of 148 CTAs
        ToepParams q = p;
        q.kb = atoi(argv[4]); q.nblk = nk / q.kb; q.ranges = 0; q.debug_tile = nullptr;
        for (int k = 0; k < 3; ++k) q.s[0][k] = s1p[k];
        float2* dp2; CK(cudaMalloc(&dp2, (size_t)q.nblk * npass * HT * sizeof(float2))); q.partial = dp2;
        CK(cudaFuncSetAttribute(toepcorr_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)toep_smem_bytes(HT)));
        cudaEvent_t a0, a1; cudaEventCreate(&a0); cudaEventCreate(&a1);
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(dclk, 0, 512));
            cudaEventRecord(a0);
            toepcorr_kernel<false, false><<<148, THREADS, toep_smem_bytes(HT)>>>(q, ToepMaps{});
            cudaEventRecord(a1);
            CK(cudaDeviceSynchronize());
        }
        float ms2; cudaEventElapsedTime(&ms2, a0, a1);
        long long hc[64]; CK(cudaMemcpy(hc, dclk, 512, cudaMemcpyDeviceToHost));
        printf("block mode: nblk=%d kb=%d items=%d kernel %.2f us\n", q.nblk, q.kb, q.nblk * npass, ms2 * 1e3);
        for (int it = 0; it < 7; ++it) printf("  item %d: MMA committed at %lld, epilogue done at %lld\n", it, hc[16 + it] - hc[0], hc[24 + it] - hc[0]);
        CK(cudaMemset(dclk, 0, 512));
    }
    const size_t smem = toep_smem_bytes(HT);
    CK(cudaFuncSetAttribute(toepcorr_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 4; ++rep) {
        if (rep == 1) { p.debug_tile = nullptr; }      // rep 0 dumps the tile, the timed reps do not
        cudaEventRecord(e0);
        toepcorr_kernel<true, false><<<2 * npass * ranges, THREADS, smem>>>(p, ToepMaps{});
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel time %.3f us\n", ms * 1e3);
    { long long hc[64]; CK(cudaMemcpy(hc, dclk, 512, cudaMemcpyDeviceToHost));
      printf("unit timing (warp 0, 2nd tile): tmem_ld %lld, +add/STS %lld, +diag loads %lld, +RMW %lld cycles\n", hc[9]-hc[8], hc[10]-hc[9], hc[11]-hc[10], hc[12]-hc[11]);
      printf("CTA0 cycles: setup->first full %lld, ->last commit %lld, ->tmem_full seen %lld, ->epilogue done %lld\n", hc[1]-hc[0], hc[2]-hc[0], hc[3]-hc[0], hc[4]-hc[0]);
      printf("  MMA thread waited on full: %lld cyc; loader waited on empty: %lld cyc, on cp.async groups: %lld cyc (sums over 3 reps)\n", hc[5], hc[6], hc[7]); }

    std::vector<float2> part((size_t)2 * rows * HT);
    std::vector<float> tile(128 * 256);
    CK(cudaMemcpy(part.data(), dpart, part.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tile.data(), dbg, tile.size() * 4, cudaMemcpyDeviceToHost));

    // ---- check 1: raw accumulator tile of CTA 0 (problem 0, pass 0, range 0)
    {
        const int k1 = (int)(((long long)nk * 1) / ranges);
        double maxerr = 0, maxval = 0;
        for (int u = 0; u < 128; ++u)
            for (int v = 0; v < 256; ++v) {
                double acc = 0;
                for (int a = 0; a < k1 * 16; ++a) {
                    const long long iu = (long long)a * 128 + u, iv = (long long)a * 128 + v;
                    const long long qu = iu / 2, qv = iv / 2;
                    const double zu = qu < n ? ((iu & 1) ? ref[qu].y : ref[qu].x) : 0.0;
                    const float2 sv = ref[qv % n];
                    const double zv = (iv & 1) ? sv.y : sv.x;
                    acc += zu * zv;
                }
                maxerr = fmax(maxerr, fabs(acc - tile[u * 256 + v]));
                maxval = fmax(maxval, fabs(acc));
            }
        printf("tile check: max|err|=%.3e max|val|=%.3e rel=%.3e\n", maxerr, maxval, maxerr / maxval);
    }
    // ---- check 2: lag correlations
    for (int prob = 0; prob < 2; ++prob) {
        double maxerr = 0, maxval = 0;
        for (int l = 0; l < nlag; ++l) {
            std::complex<double> want(0, 0);
            const int dmin = prob ? -peek : 0;
            for (int i = 0; i < n; ++i) {
                const float2 xv = ref[i];
                long long j = ((long long)i + dmin + l) % n; if (j < 0) j += n;
                const float2 sv = prob ? srv[j] : ref[j];
                want += std::complex<double>(xv.x, xv.y) * std::conj(std::complex<double>(sv.x, sv.y));
            }
            std::complex<double> got(0, 0);
            for (int r = 0; r < rows; ++r) { const float2 v = part[((size_t)prob * rows + r) * HT + l]; got += std::complex<double>(v.x, v.y); }
            maxerr = fmax(maxerr, std::abs(got - want));
            maxval = fmax(maxval, std::abs(want));
            if (l < 3 || l == nlag - 1) printf("  prob %d lag %3d want (%.4f, %.4f) got (%.4f, %.4f)\n", prob, l, want.real(), want.imag(), got.real(), got.imag());
        }
        printf("problem %d: max|err|=%.3e max|C|=%.3e rel=%.3e\n", prob, maxerr, maxval, maxerr / maxval);
    }
    return 0;
}
