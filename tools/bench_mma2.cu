// Round-2 preparation (NOT validated on hardware yet - written without GPU access at the end of round 1): probe of
// tcgen05.mma.cta_group::2 for the conv-as-shifted-GEMM mapping.  A CTA pair (cluster of 2, same TPC) executes M = 256:
// each CTA supplies its own 128 activation rows (A, tap-shifted start) and HALF of the weight rows (B, N/2 x K), so per SM
// the operand fetch drops from 4096 + 32 N to 4096 + 16 N bytes per MMA and the weight ring traffic halves.
//   part 1: correctness of one accumulation chain against an integer CPU reference (both CTAs' accumulators);
//   part 2: cycles per MMA for N in {64, 128, 256} (per CTA pair, both SMs busy).
// Every wait is bounded (clock64 timeout) so that a wrong assumption reports FAIL instead of hanging the GPU.
// Open questions this probe answers: (a) is tcgen05.alloc.cta_group::2 issued by one warp of EACH CTA (assumed here);
// (b) do the A/B descriptors use CTA-local shared addresses in both CTAs (assumed: same offsets in both);
// (c) the layout of B across the pair (assumed: CTA r holds rows [r*N/2, (r+1)*N/2) of the N x K operand).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I so-vits-svc_b200/csrc -o bench_mma2 tools/bench_mma2.cu
#include "../so-vits-svc_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace svb::tc;

__host__ __device__ inline float a_val(int r, int j) { return (float)(((r * 7 + j * 3) % 13) - 6); }
__host__ __device__ inline float b_val(int n, int j) { return (float)(((n * 5 + j) % 7) - 3); }

constexpr int AROWS = 192;                 // 128 rows + tap-shift slack
constexpr long long TIMEOUT_CLK = 200000000LL;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in both CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar) {
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > TIMEOUT_CLK) return false;
    }
    return true;
}

// out: [2 CTAs][128 rows][N] accumulators (part 1), clk: cycles of the issue->completion window seen by the leader
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_mma_kernel(float* out, long long* clk, int* status, int RB, int N, int r0, int n_round, int check) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t b_base = base + AROWS * 128;
    const uint32_t bar = b_base + 128 * 128;
    const uint32_t slot = bar + 16;
    const int K = RB / 2, NH = N / 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    // A: this CTA's 128(+slack) activation rows (global row index = rank*1000 + r keeps the two tiles distinguishable)
    for (int idx = tid; idx < AROWS * (RB / 16); idx += 128) {
        const int r = idx / (RB / 16), ch = idx % (RB / 16);
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) w[e] = pack_h2(a_val(rank * 1000 + r, ch * 8 + 2 * e), a_val(rank * 1000 + r, ch * 8 + 2 * e + 1));
        *reinterpret_cast<uint4*>(sm + swz_offset(r, ch, RB)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // B: this CTA's half of the weight rows
    for (int idx = tid; idx < NH * (RB / 16); idx += 128) {
        const int n = idx / (RB / 16), ch = idx % (RB / 16);
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) w[e] = pack_h2(b_val(rank * NH + n, ch * 8 + 2 * e), b_val(rank * NH + n, ch * 8 + 2 * e + 1));
        *reinterpret_cast<uint4*>(sm + AROWS * 128 + swz_offset(n, ch, RB)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc2(slot, 256); tmem_relinquish2(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // both CTAs' operands and barriers are ready
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - base));
    long long t0 = clock64();
    if (rank == 0 && warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(256, N);
            const uint64_t a_d = make_smem_desc(a_base + r0 * RB, RB, 0);
            const uint64_t b_d = make_smem_desc(b_base, RB, 0);
            for (int it = 0; it < n_round; ++it)
                for (int ks = 0; ks < K / 16; ++ks)
                    umma2_f16(tmem, a_d + (uint64_t)((ks * 32) >> 4), b_d + (uint64_t)((ks * 32) >> 4), idesc, (it > 0 || ks > 0) ? 1u : 0u);
            umma2_commit_multicast(bar);
        }
        __syncwarp();
    }
    const bool ok = mbar_wait_bounded(bar, 0);
    long long t1 = clock64();
    if (!ok) { if (tid == 0) status[rank] = -1; }
    else {
        tc_fence_after();
        if (tid == 0) { status[rank] = 1; clk[rank] = t1 - t0; }
        if (check) {
            const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tl + c0, r);
                tmem_ld_wait();
                for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + 32 * warp + lane) * N + c0 + j] = __uint_as_float(r[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // the peer may still be reading its accumulators
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 256); }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("device: %s SMs=%d\n", prop.name, prop.multiProcessorCount);
    const size_t smem = 1024 + AROWS * 128 + 128 * 128 + 64;
    cudaFuncSetAttribute(pair_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d_out; long long* d_clk; int* d_st;
    cudaMalloc(&d_out, 2 * 128 * 256 * sizeof(float)); cudaMalloc(&d_clk, 2 * sizeof(long long)); cudaMalloc(&d_st, 2 * sizeof(int));
    // ---- part 1: correctness
    const int rbs[2] = {128, 64};
    const int ns[3] = {64, 128, 256};
    const int r0s[3] = {0, 5, 50};
    int bad = 0;
    for (int rb : rbs)
        for (int N : ns)
            for (int r0 : r0s) {
                cudaMemset(d_out, 0, 2 * 128 * 256 * sizeof(float)); cudaMemset(d_st, 0, 2 * sizeof(int));
                pair_mma_kernel<<<2, 128, smem>>>(d_out, d_clk, d_st, rb, N, r0, 1, 1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("RB=%d N=%d r0=%d CUDA error %s\n", rb, N, r0, cudaGetErrorString(e)); return 1; }
                int st[2]; cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
                if (st[0] != 1 || st[1] != 1) { printf("RB=%3d N=%3d r0=%2d TIMEOUT (status %d %d)\n", rb, N, r0, st[0], st[1]); ++bad; continue; }
                std::vector<float> h(2 * 128 * N);
                cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                double maxerr = 0;
                for (int rank = 0; rank < 2; ++rank)
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < N; ++n) {
                            double ref = 0;
                            for (int j = 0; j < rb / 2; ++j) ref += (double)a_val(rank * 1000 + r0 + m, j) * b_val(n, j);
                            const double dd = fabs(ref - h[((size_t)rank * 128 + m) * N + n]);
                            if (dd > maxerr) maxerr = dd;
                        }
                printf("RB=%3d N=%3d r0=%2d maxerr=%g %s\n", rb, N, r0, maxerr, maxerr == 0 ? "OK" : "MISMATCH");
                if (maxerr != 0) ++bad;
            }
    printf("SUMMARY part 1: %d failing configurations\n", bad);
    // ---- part 2: rate (one CTA pair per TPC: grid = SM count)
    for (int N : ns) {
        const int n_round = 512;
        cudaMemset(d_st, 0, 2 * sizeof(int));
        pair_mma_kernel<<<prop.multiProcessorCount & ~1, 128, smem>>>(d_out, d_clk, d_st, 128, N, 0, n_round, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate N=%d CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
        long long c[2]; int st[2];
        cudaMemcpy(c, d_clk, sizeof(c), cudaMemcpyDeviceToHost); cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
        const double n_mma = (double)n_round * 4;     // RB=128 -> 4 k-steps per round
        printf("RATE cta_group::2 N=%3d: %.1f clk per M=256 MMA (status %d %d); cta_group::1 law: %d clk per M=128 MMA\n", N, (double)c[0] / n_mma, st[0], st[1],
               (N / 2 > (4096 + 32 * N) / 128) ? N / 2 : (4096 + 32 * N) / 128);
    }
    return 0;
}
