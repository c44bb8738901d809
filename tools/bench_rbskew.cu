// Timeline harness for the block-skewed, persistent fused-ResBlock kernel: compiles kernels_resblock.cu with -DSVB_TRACE, runs
// one launch on synthetic data and prints, for the SECOND tile of every CTA (steady state), how long the MMAs of a block take
// from issue to "accumulators visible", how long the workers wait for them, how long an epilogue chain is, and the tile period.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSVB_TRACE -I so-vits-svc_b200/csrc -o bench_rbskew tools/bench_rbskew.cu
#include "../so-vits-svc_b200/csrc/kernels_resblock.cu"
#include <cstdio>
#include <vector>
#include <algorithm>

namespace svb { int64_t& launch_counter() { static int64_t c = 0; return c; } }

int main(int argc, char** argv) {
    const int C = argc > 1 ? atoi(argv[1]) : 32;
    const int k = argc > 2 ? atoi(argv[2]) : 7;
    const int B = 8;
    const int T = 862 * 512 / (C == 16 ? 1 : C == 32 ? 2 : 4);
    const int MB = C == 16 ? 8 : 4;
    float *x, *out, *bias;
    uint8_t* w;
    cudaMalloc(&x, (size_t)B * C * T * 4); cudaMalloc(&out, (size_t)B * C * T * 4);
    cudaMemset(x, 0, (size_t)B * C * T * 4); cudaMemset(out, 0, (size_t)B * C * T * 4);
    const size_t wbytes = (size_t)k * C * C * 2;
    cudaMalloc(&w, 6 * wbytes); cudaMemset(w, 0, 6 * wbytes);
    cudaMalloc(&bias, 6 * C * 4); cudaMemset(bias, 0, 6 * C * 4);
    svb::ResblockTC a;
    a.x = x; a.out = out; a.B = B; a.C = C; a.T = T; a.k = k; a.alpha = 1.f / 3; a.beta = 1.f; a.variant = 2;
    a.dil[0] = 1; a.dil[1] = 3; a.dil[2] = 5;
    for (int d = 0; d < 3; ++d) a.inv[d] = 1.f;
    for (int q = 0; q < 6; ++q) { a.w[q] = w + q * wbytes; a.bias[q] = bias + q * C; }
    for (int i = 0; i < 2; ++i) if (svb::launch_resblock_tc(a, 0)) { printf("launch failed\n"); return 1; }
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) svb::launch_resblock_tc(a, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("C=%d k=%d T=%d: %.3f ms per launch (untraced)\n", C, k, T, ms / 5);
    const size_t max_ctas = 1024;
    long long* tr; cudaMalloc(&tr, max_ctas * 256 * 8); cudaMemset(tr, 0, max_ctas * 256 * 8);
    cudaMemcpyToSymbol(svb::g_rb_trace, &tr, sizeof(tr));
    svb::launch_resblock_tc(a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(max_ctas * 256);
    cudaMemcpy(h.data(), tr, max_ctas * 256 * 8, cudaMemcpyDeviceToHost);
    double exec[6] = {0}, idle[6] = {0}, epi[6] = {0}, gap[6] = {0}, period = 0, react[6] = {0};
    int n = 0, np = 0;
    for (size_t c = 0; c < max_ctas; ++c) {
        const long long* t = &h[c * 256];
        if (!t[0] || !t[(5 * 8 + MB - 1) * 4 + 3]) continue;
        ++n;
        if (t[201] && t[202]) { period += t[202] - t[201]; ++np; }
        for (int q = 0; q < 6; ++q) {
            double ex = 0, id = 0, ep = 0, gp = 0, rc = 0;
            for (int mb = 0; mb < MB; ++mb) {
                const long long* s = t + (q * 8 + mb) * 4;
                ex += s[2] - s[0]; id += s[2] - s[1]; ep += s[3] - s[2];
                if (mb > 0) gp += s[0] - (s - 4)[0];
                // issuer reaction: issue time minus the latest hand-off it depends on (blocks mb-1..mb+1 of conv q-1)
                if (q > 0) {
                    long long dep = 0;
                    for (int d = -1; d <= 1; ++d) if (mb + d >= 0 && mb + d < MB) dep = std::max(dep, t[((q - 1) * 8 + mb + d) * 4 + 3]);
                    rc += s[0] - dep;
                }
            }
            exec[q] += ex / MB; idle[q] += id / MB; epi[q] += ep / MB; gap[q] += gp / (MB - 1); react[q] += rc / MB;
        }
    }
    printf("CTAs traced %d; tile period (issuer, tile 1 -> tile 2) %.0f clk\n", n, np ? period / np : 0.0);
    for (int q = 0; q < 6; ++q)
        printf("  conv %d: issue->visible %.0f | worker wait %.0f | epilogue chain %.0f | issue-to-issue %.0f | issue - last dependency %.0f\n",
               q, exec[q] / n, idle[q] / n, epi[q] / n, gap[q] / n, react[q] / n);
    {   // fine trace of (conv 2, block 2): averages over CTAs, relative to "accumulators visible"
        double f[7] = {0}; int nf = 0;
        for (size_t c = 0; c < max_ctas; ++c) {
            const long long* t = &h[c * 256];
            if (!t[0] || !t[208] || !t[214]) continue;
            const long long v = t[(2 * 8 + 2) * 4 + 2];
            for (int i = 0; i < 7; ++i) f[i] += (double)(t[208 + i] - v);
            ++nf;
        }
        if (nf) printf("  fine (conv 2, block 2; clk after 'visible'): before LDTM %.0f | after wait::ld %.0f | math done %.0f | neighbour wait done %.0f | stores issued %.0f | after fence.proxy.async %.0f | after arrive %.0f\n",
                       f[0] / nf, f[1] / nf, f[2] / nf, f[3] / nf, f[4] / nf, f[5] / nf, f[6] / nf);
    }
    // one CTA in full, times relative to its tile start
    for (size_t c = 0; c < max_ctas; ++c) {
        const long long* t = &h[c * 256];
        if (!t[0] || !t[201]) continue;
        const long long t0 = t[201];
        printf("CTA %zu on SM %lld (times - tile start):\n", c, t[255]);
        for (int q = 0; q < 6; ++q) {
            printf("  q%d:", q);
            for (int mb = 0; mb < MB; ++mb) {
                const long long* s = t + (q * 8 + mb) * 4;
                printf(" [b%d i%lld w%lld v%lld d%lld]", mb, s[0] - t0, s[1] - t0, s[2] - t0, s[3] - t0);
            }
            printf("\n");
        }
        if (c >= 1) break;
    }
    return 0;
}
