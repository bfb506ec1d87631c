// emul.cu -- runs the per-thread phases of fftcore.cuh thread-by-thread on the CPU and compares with a float64 DFT.
//   nvcc -O2 -std=c++17 -I passiveradar_b200/csrc -o /tmp/fft_emul scripts/fft/emul.cu && /tmp/fft_emul
// No CUDA runtime call is made: this pins the index algebra / twiddle conventions without a GPU.
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>
#include "fftcore.cuh"

using namespace prc::fft;
typedef std::complex<double> cd;

template <int R3>
void make_tw(std::vector<float2>& tw) {
    using G = Geo<R3>;
    tw.resize(G::TW1 + G::TW2);
    for (int i = 0; i < G::TW1; ++i) {
        const int k1 = i / G::T, t = i % G::T;
        const double a = -2.0 * M_PI * (double)((long long)t * k1 % G::L) / G::L;
        tw[i] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int q = 0; q < G::TW2; ++q) {
        const int k2 = q / R3, n3 = q % R3;
        const double a = -2.0 * M_PI * (double)(n3 * k2) / (16 * R3);
        tw[G::TW1 + q] = make_float2((float)cos(a), (float)sin(a));
    }
}

template <int R3>
int run() {
    using G = Geo<R3>;
    std::mt19937 rng(1234 + R3);
    std::normal_distribution<float> nd;
    std::vector<float2> x(G::L), tw;
    for (auto& v : x) v = make_float2(nd(rng), nd(rng));
    make_tw<R3>(tw);
    const float2* tw1 = tw.data();
    const float2* tw2 = tw.data() + G::TW1;
    std::vector<float2> ea(G::EX), eb(G::EX);
    struct Regs { float2 v[16]; };
    std::vector<Regs> regs(G::T);
    // float64 DFT
    std::vector<cd> X(G::L);
    for (int k = 0; k < G::L; ++k) {
        cd s = 0;
        for (int n = 0; n < G::L; ++n) {
            const double a = -2.0 * M_PI * (double)((long long)n * k % G::L) / G::L;
            s += cd(x[n].x, x[n].y) * cd(cos(a), sin(a));
        }
        X[k] = s;
    }
    double xmax = 0;
    for (auto& v : X) xmax = std::max(xmax, std::abs(v));
    // n2p
    for (int t = 0; t < G::T; ++t) for (int n1 = 0; n1 < 16; ++n1) regs[t].v[n1] = x[n1 * G::T + t];
    for (int t = 0; t < G::T; ++t) n2p_pass1<R3>(regs[t].v, t, ea.data(), tw1);
    for (int t = 0; t < G::T; ++t) n2p_pass2<R3>(regs[t].v, t, ea.data(), eb.data(), tw2);
    for (int t = 0; t < G::T; ++t) n2p_pass3<R3>(regs[t].v, t, eb.data());
    double e1 = 0;
    std::vector<int> seen(G::L, 0);
    for (int t = 0; t < G::T; ++t)
        for (int r = 0; r < 16; ++r) {
            const int f = perm_freq<R3>(t, r);
            seen[f]++;
            e1 = std::max(e1, std::abs(cd(regs[t].v[r].x, regs[t].v[r].y) - X[f]));
        }
    int bad = 0;
    for (int f = 0; f < G::L; ++f) bad += seen[f] != 1;
    // p2n applied to the permuted spectrum: y[n] = sum_k X[k] W^{nk}  ==  L * x[(-n) mod L]
    for (int t = 0; t < G::T; ++t) p2n_pass1<R3>(regs[t].v, t, eb.data());
    for (int t = 0; t < G::T; ++t) p2n_pass2<R3>(regs[t].v, t, eb.data(), ea.data(), tw2);
    for (int t = 0; t < G::T; ++t) p2n_pass3<R3>(regs[t].v, t, ea.data(), tw1);
    double e2 = 0, ymax = 0;
    for (int t = 0; t < G::T; ++t)
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = n1 * G::T + t;
            const float2 want = x[(G::L - n) % G::L];
            const cd w((double)want.x * G::L, (double)want.y * G::L);
            ymax = std::max(ymax, std::abs(w));
            e2 = std::max(e2, std::abs(cd(regs[t].v[n1].x, regs[t].v[n1].y) - w));
        }
    printf("R3=%2d L=%4d: n2p err %.3e (rel %.2e), perm bijective: %s, p2n(n2p(x)) err rel %.2e\n", R3, G::L, e1, e1 / xmax,
           bad ? "NO" : "yes", e2 / ymax);
    return (e1 / xmax < 2e-6 && !bad && e2 / ymax < 4e-6) ? 0 : 1;
}

int main() {
    int rc = run<4>() | run<8>() | run<16>();
    printf(rc ? "FAIL\n" : "OK\n");
    return rc;
}
