// Micro-benchmarks of the tcgen05 costs the conv kernels are built around (run on a B200):
//   (1) cycles per tcgen05.mma (M=128, K=16, fp16) for N in {16..256}, operand row width 128/64/32 B, with the MMAs chained
//       on ONE accumulator or rotated over several, from one CTA per SM and from two co-resident CTAs;
//   (2) cycles per tcgen05.ld 32x32b.x16 with 4 / 8 warps reading (TMEM read bandwidth).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_mma bench_mma.cu
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace svb::tc;

constexpr int AROWS = 512;

// NACC accumulators used round-robin (1 = every MMA chained on the same accumulator); vary_a: A row start moves per round
template <int RB, int NACC>
__global__ void __launch_bounds__(128, 2) mma_rate_kernel(long long* out, int N, int n_round, int vary_a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t b_base = base + AROWS * RB;
    const uint32_t bar = b_base + 256 * RB;
    const uint32_t slot = bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (AROWS + 256) * RB / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - base));
    if (warp == 1) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t a_d0 = make_smem_desc(a_base, RB, 0);
        const uint64_t b_d0 = make_smem_desc(b_base, RB, 0);
        constexpr int KSTEPS = RB / 32;
        long long t0 = clock64();
        if (elect_one()) {
            uint32_t shift = 0;
            for (int i = 0; i < n_round; ++i) {
                const uint64_t ar = a_d0 + (uint64_t)((shift * RB) >> 4);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                    for (int acc = 0; acc < NACC; ++acc)
                        umma_f16(tmem + acc * N, ar + (uint64_t)((ks * 32) >> 4), b_d0 + (uint64_t)((ks * 32) >> 4), idesc, 1u);
                if (vary_a) shift = (shift + 5) & 127;
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if (tid == 32) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int RB, int NACC>
void run_rate(long long* d_out, int sms, int ctas, int N) {
    if (NACC * N > 256) return;
    constexpr int KSTEPS = RB / 32;
    const int n_round = 4096 / (KSTEPS * NACC);
    const int n_mma = n_round * KSTEPS * NACC;
    const size_t smem = 1024 + (AROWS + 256) * RB + 64;
    cudaFuncSetAttribute(mma_rate_kernel<RB, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int vary = 0; vary < 2; ++vary) {
        std::vector<long long> h(ctas * sms);
        for (int rep = 0; rep < 2; ++rep) {
            mma_rate_kernel<RB, NACC><<<ctas * sms, 128, smem>>>(d_out, N, n_round, vary);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(h.data(), d_out, sizeof(long long) * ctas * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= h.size();
        printf("MMA ctas/SM=%d RB=%3d N=%3d nacc=%d varyA=%d : %6.1f clk/MMA per CTA (%6.1f per SM)  math-floor %5.1f\n", ctas, RB, N, NACC, vary,
               avg / n_mma, avg / n_mma / ctas, 128.0 * N * 16 / 4096.0);
    }
}
template <int RB>
void run_rb(long long* d_out, int sms, int ctas) {
    const int ns[5] = {16, 32, 64, 128, 256};
    for (int N : ns) {
        run_rate<RB, 1>(d_out, sms, ctas, N);
        run_rate<RB, 2>(d_out, sms, ctas, N);
        run_rate<RB, 4>(d_out, sms, ctas, N);
        run_rate<RB, 8>(d_out, sms, ctas, N);
    }
}

// The conv kernels' real issue pattern: per tap MB row-blocks x KSTEPS k-steps, A start moves by `cd` rows per tap, B walks
// through an 11-tap weight stage, one accumulator per row block.
template <int RB, int MB>
__global__ void __launch_bounds__(128, 2) mma_pattern_kernel(long long* out, int n_tap, int cd) {
    constexpr int C = RB / 2, KSTEPS = RB / 32, SUB = C * RB, ROWS = 128 * MB + 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t b_base = base + ((ROWS * RB + 1023) / 1024) * 1024;
    const uint32_t bar = b_base + 11 * SUB;
    const uint32_t slot = bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (int)((bar - base) / 16); i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - base));
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_f16(128, C);
        long long t0 = clock64();
        if (elect_one()) {
            const uint64_t a_step = (uint64_t)((uint32_t)(cd * RB) >> 4);
            for (int t0_ = 0; t0_ < n_tap; t0_ += 11) {
                uint64_t ad = make_smem_desc(a_base, RB, 0);
                uint64_t bd = make_smem_desc(b_base, RB, 0);
                for (int i = 0; i < 11; ++i) {
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks)
                            umma_f16(tmem + mb * C, ad + (uint64_t)(((uint32_t)(mb * 128) * RB + ks * 32) >> 4), bd + (uint64_t)((ks * 32) >> 4), idesc, 1u);
                    ad += a_step;
                    bd += (uint64_t)(SUB >> 4);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if (tid == 32) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int RB, int MB>
void run_pattern(long long* d_out, int sms) {
    constexpr int C = RB / 2, SUB = C * RB, ROWS = 128 * MB + 64;
    const size_t smem = 1024 + ((ROWS * RB + 1023) / 1024) * 1024 + 11 * SUB + 64;
    cudaFuncSetAttribute(mma_pattern_kernel<RB, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tap = 11 * 40;
    for (int ctas = 1; ctas <= 2; ++ctas)
        for (int cd = 1; cd <= 5; cd += 2) {
            std::vector<long long> h(ctas * sms);
            for (int rep = 0; rep < 2; ++rep) {
                mma_pattern_kernel<RB, MB><<<ctas * sms, 128, smem>>>(d_out, n_tap, cd);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
            }
            cudaMemcpy(h.data(), d_out, sizeof(long long) * ctas * sms, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (auto v : h) avg += (double)v;
            avg /= h.size();
            const double n_mma = (double)n_tap * MB * (RB / 32);
            printf("PATTERN ctas/SM=%d C=%3d MB=%d dil=%d : %6.1f clk/MMA per CTA (%6.1f per SM)\n", ctas, C, MB, cd, avg / n_mma, avg / n_mma / ctas);
        }
}

__global__ void __launch_bounds__(256, 2) tmem_ld_kernel(long long* out, float* sink, int n_iter, int nwarps, int with_st) {
    __shared__ uint32_t slot_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(smem_u32(&slot_s), 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot_s + ((uint32_t)(32 * (warp & 3)) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < n_iter; ++i) {
            uint32_t r[16];
            tmem_ld16(tmem + ((i * 16) & 127) + (warp >> 2) * 128, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += __uint_as_float(r[j]);
            if (with_st) { tmem_st16(tmem + ((i * 16) & 127) + (warp >> 2) * 128, r); }
        }
        if (with_st) tmem_st_wait();
    }
    long long t1 = clock64();
    if (acc == 123.456f) sink[tid] = acc;
    if ((tid & 31) == 0 && warp < nwarps) out[blockIdx.x * 8 + warp] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot_s, 256); }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device: %s SMs=%d\n", prop.name, sms);
    long long* d_out;
    float* d_sink;
    cudaMalloc(&d_out, sizeof(long long) * 8 * 2 * sms);
    cudaMalloc(&d_sink, 4096);
    run_pattern<32, 8>(d_out, sms);
    run_pattern<64, 4>(d_out, sms);
    run_pattern<128, 2>(d_out, sms);
    if (getenv("PATTERN_ONLY")) return 0;
    for (int ctas = 1; ctas <= 2; ++ctas) {
        run_rb<128>(d_out, sms, ctas);
        run_rb<64>(d_out, sms, ctas);
        run_rb<32>(d_out, sms, ctas);
    }
    for (int ctas = 1; ctas <= 2; ++ctas)
        for (int nw = 4; nw <= 8; nw += 4)
            for (int st = 0; st < 2; ++st) {
                const int n_iter = 2048;
                std::vector<long long> h(8 * ctas * sms, 0);
                cudaMemset(d_out, 0, sizeof(long long) * 8 * 2 * sms);
                for (int rep = 0; rep < 2; ++rep) {
                    tmem_ld_kernel<<<ctas * sms, 256>>>(d_out, d_sink, n_iter, nw, st);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h.data(), d_out, sizeof(long long) * 8 * ctas * sms, cudaMemcpyDeviceToHost);
                double mx = 0;
                for (auto v : h) if ((double)v > mx) mx = (double)v;
                const double bytes = (double)n_iter * 2048.0 * nw * ctas;
                printf("TMEM ld.x16%s ctas/SM=%d warps=%d : %.1f clk/iter/warp, %.1f B/clk/SM\n", st ? "+st" : "", ctas, nw, mx / n_iter, bytes / mx);
            }
    return 0;
}
