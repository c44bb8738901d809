// Phase-timeline harness for the fused ResBlock kernel: compiles kernels_resblock.cu with -DSVB_TRACE, runs one launch on
// synthetic data and prints the average duration of each phase of a CTA (clock64 cycles) plus the makespan per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSVB_TRACE -I so-vits-svc_b200/csrc -o bench_rb tools/bench_rb.cu
#include "../so-vits-svc_b200/csrc/kernels_resblock.cu"
#include <cstdio>
#include <vector>
#include <map>
#include <algorithm>

namespace svb { int64_t& launch_counter() { static int64_t c = 0; return c; } }

int main(int argc, char** argv) {
    const int C = argc > 1 ? atoi(argv[1]) : 32;
    const int k = argc > 2 ? atoi(argv[2]) : 11;
    const int variant = argc > 3 ? atoi(argv[3]) : -1;
    const int B = 8;
    const int T = 862 * 512 / (C == 16 ? 1 : C == 32 ? 2 : 4);
    float *x, *out, *bias;
    uint8_t* w;
    cudaMalloc(&x, (size_t)B * C * T * 4); cudaMalloc(&out, (size_t)B * C * T * 4);
    cudaMemset(x, 0, (size_t)B * C * T * 4); cudaMemset(out, 0, (size_t)B * C * T * 4);
    const size_t wbytes = (size_t)k * C * C * 2;
    cudaMalloc(&w, 6 * wbytes); cudaMemset(w, 0, 6 * wbytes);
    cudaMalloc(&bias, 6 * C * 4); cudaMemset(bias, 0, 6 * C * 4);
    svb::ResblockTC a;
    a.x = x; a.out = out; a.B = B; a.C = C; a.T = T; a.k = k; a.alpha = 1.f / 3; a.beta = 0.f; a.variant = variant;
    for (int q = 0; q < 6; ++q) { a.w[q] = w + q * wbytes; a.bias[q] = bias + q * C; }
    // warm-up without tracing
    for (int i = 0; i < 2; ++i) if (svb::launch_resblock_tc(a, 0)) { printf("launch failed\n"); return 1; }
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) svb::launch_resblock_tc(a, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("C=%d k=%d T=%d: %.3f ms per launch (untraced)\n", C, k, T, ms / 5);
    // traced launch
    const size_t max_ctas = 1 << 16;
    long long* tr; cudaMalloc(&tr, max_ctas * 64 * 8); cudaMemset(tr, 0, max_ctas * 64 * 8);
    cudaMemcpyToSymbol(svb::g_rb_trace, &tr, sizeof(tr));
    svb::launch_resblock_tc(a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(max_ctas * 64);
    cudaMemcpy(h.data(), tr, max_ctas * 64 * 8, cudaMemcpyDeviceToHost);
    double load = 0, total = 0, mma_wait[6] = {0}, mma_issue[6] = {0}, acc_wait[6] = {0}, epi[6] = {0}, mma_exec[6] = {0};
    int n = 0;
    for (size_t c = 0; c < max_ctas; ++c) {
        const long long* t = &h[c * 64];
        if (!t[0] || !t[30]) continue;
        ++n;
        load += t[1] - t[0]; total += t[30] - t[0];
        long long prev_epi = t[1];
        for (int q = 0; q < 6; ++q) {
            const long long acc = t[2 + 4 * q], ed = t[2 + 4 * q + 1], ms_ = t[2 + 4 * q + 2], mi = t[2 + 4 * q + 3];
            mma_wait[q] += ms_ - prev_epi;       // epilogue/loader arrive -> MMA warp running
            mma_issue[q] += mi - ms_;            // issue loop (incl. waits on weights)
            mma_exec[q] += acc - ms_;            // MMA start -> workers see the accumulators
            acc_wait[q] += acc - prev_epi;       // workers idle
            epi[q] += ed - acc;
            prev_epi = ed;
        }
    }
    printf("CTAs traced %d; avg cycles: total %.0f  load %.0f\n", n, total / n, load / n);
    for (int q = 0; q < 6; ++q)
        printf("  conv %d: handoff->MMA start %.0f | issue loop %.0f | MMA start->acc visible %.0f | workers idle %.0f | epilogue %.0f\n", q, mma_wait[q] / n,
               mma_issue[q] / n, mma_exec[q] / n, acc_wait[q] / n, epi[q] / n);
    // raw timeline of the CTAs that ran on one SM (absolute clock64, relative to the first start on that SM)
    const long long sm_pick = argc > 4 ? atoi(argv[4]) : 5;
    std::vector<const long long*> on_sm;
    for (size_t c = 0; c < max_ctas; ++c) { const long long* t = &h[c * 64]; if (t[0] && t[30] && t[63] == sm_pick) on_sm.push_back(t); }
    std::sort(on_sm.begin(), on_sm.end(), [](const long long* a, const long long* b) { return a[0] < b[0]; });
    if (!on_sm.empty()) {
        const long long z = on_sm[0][0];
        for (size_t i = 0; i < on_sm.size() && i < 8; ++i) {
            const long long* t = on_sm[i];
            printf("  SM%lld cta#%zu start %lld load_done %lld |", sm_pick, i, t[0] - z, t[1] - z);
            for (int q = 0; q < 6; ++q) printf(" q%d[mma %lld..%lld acc %lld epi_done %lld]", q, t[2 + 4 * q + 2] - z, t[2 + 4 * q + 3] - z, t[2 + 4 * q] - z, t[2 + 4 * q + 1] - z);
            printf(" end %lld\n", t[30] - z);
        }
    }
    return 0;
}
