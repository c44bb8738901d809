// Phase-timeline harness for pair_tc_kernel: compiles kernels_tc.cu with -DSVB_TRACE, runs one launch on synthetic data and
// prints the average duration of each phase of a CTA (clock64 cycles) and the raw timeline of the CTAs of one SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSVB_TRACE -o bench_pairtrace tools/bench_pairtrace.cu -lcuda
#include "../so-vits-svc_b200/csrc/kernels_tc.cu"
#include <cstdio>
#include <vector>
#include <algorithm>

namespace svb { int64_t& launch_counter() { static int64_t c = 0; return c; } }

int main(int argc, char** argv) {
    const int C = argc > 1 ? atoi(argv[1]) : 128;
    const int k = argc > 2 ? atoi(argv[2]) : 11;
    const int dil = argc > 3 ? atoi(argv[3]) : 3;
    const float beta = argc > 4 ? (float)atof(argv[4]) : 0.f;
    const int B = 8;
    const int T = C == 256 ? 6896 : 862 * 512 * 16 / C;   // stage lengths of the 44.1 kHz generator
    float *x, *out, *bias;
    uint8_t* w;
    cudaMalloc(&x, (size_t)B * C * T * 4); cudaMalloc(&out, (size_t)B * C * T * 4);
    cudaMemset(x, 0, (size_t)B * C * T * 4); cudaMemset(out, 0, (size_t)B * C * T * 4);
    const size_t wbytes = (size_t)k * C * C * 2;
    cudaMalloc(&w, 2 * wbytes); cudaMemset(w, 0, 2 * wbytes);
    cudaMalloc(&bias, 2 * C * 4); cudaMemset(bias, 0, 2 * C * 4);
    svb::PairTC a;
    a.x = x; a.out = out; a.B = B; a.C = C; a.T = T; a.k = k; a.dil = dil; a.alpha = 1.f; a.beta = beta;
    a.w1 = w; a.w2 = w + wbytes; a.b1 = bias; a.b2 = bias + C;
    for (int i = 0; i < 2; ++i) if (svb::launch_pair_tc(a, 0)) { printf("launch failed\n"); return 1; }
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) svb::launch_pair_tc(a, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("pair C=%d k=%d dil=%d beta=%g T=%d: %.3f ms per launch (untraced)\n", C, k, dil, beta, T, ms / 5);
    const size_t max_ctas = 1 << 16;
    long long* tr; cudaMalloc(&tr, max_ctas * 16 * 8); cudaMemset(tr, 0, max_ctas * 16 * 8);
    cudaMemcpyToSymbol(svb::g_pair_trace, &tr, sizeof(tr));
    svb::launch_pair_tc(a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(max_ctas * 16);
    cudaMemcpy(h.data(), tr, max_ctas * 16 * 8, cudaMemcpyDeviceToHost);
    double tot = 0, load = 0, m1 = 0, e1d = 0, m2 = 0, e2d = 0; int n = 0;
    for (size_t c = 0; c < max_ctas; ++c) {
        const long long* t = &h[c * 16];
        if (!t[0] || !t[10]) continue;
        ++n; tot += t[10] - t[0]; load += t[1] - t[0]; m1 += t[2] - t[1]; e1d += t[3] - t[2]; m2 += t[4] - t[3]; e2d += t[5] - t[4];
    }
    printf("CTAs %d; avg cycles: total %.0f | load %.0f | load_done->acc1 %.0f | epi1 %.0f | epi1_done->acc2 %.0f | epi2 %.0f\n", n, tot / n, load / n, m1 / n, e1d / n, m2 / n, e2d / n);
    const long long sm_pick = 5;
    std::vector<const long long*> on_sm;
    for (size_t c = 0; c < max_ctas; ++c) { const long long* t = &h[c * 16]; if (t[0] && t[10] && t[15] == sm_pick) on_sm.push_back(t); }
    std::sort(on_sm.begin(), on_sm.end(), [](const long long* a_, const long long* b_) { return a_[0] < b_[0]; });
    if (!on_sm.empty()) {
        const long long z = on_sm[0][0];
        for (size_t i = 0; i < on_sm.size() && i < 6; ++i) {
            const long long* t = on_sm[i];
            printf("  SM5 cta#%zu start %lld load_done %lld | M1 %lld..%lld acc1 %lld E1done %lld | M2 %lld..%lld acc2 %lld E2done %lld end %lld\n", i, t[0] - z, t[1] - z,
                   t[6] - z, t[7] - z, t[2] - z, t[3] - z, t[8] - z, t[9] - z, t[4] - z, t[5] - z, t[10] - z);
        }
    }
    return 0;
}
