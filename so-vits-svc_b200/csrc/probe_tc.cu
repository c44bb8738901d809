// Stand-alone probe of the tcgen05 operand-descriptor semantics the pair kernel relies on:
//   (1) K-major swizzled (128B/64B/32B) operands written with plain st.shared using the address-bit swizzle,
//   (2) advancing K inside the swizzle span by adding bytes to the start address,
//   (3) starting the A operand at an ARBITRARY row (tap shift) with base_offset = 0 or (addr>>7)&7.
// Prints one line per configuration with the max |error| against an integer CPU reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tc probe_tc.cu ; run on a B200.
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace svb::tc;

__host__ __device__ inline float a_val(int r, int j) { return (float)(((r * 7 + j * 3) % 13) - 6); }
__host__ __device__ inline float b_val(int n, int j) { return (float)(((n * 5 + j) % 7) - 3); }

constexpr int AROWS = 256;

__global__ void __launch_bounds__(128, 1) probe_kernel(float* out, int RB, int N, int r0, int bo_mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t b_base = base + AROWS * 128;
    const uint32_t bar = b_base + 256 * 128;
    const uint32_t slot = bar + 8;
    const int K = RB / 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < AROWS * (RB / 16); idx += 128) {
        int r = idx / (RB / 16), ch = idx % (RB / 16);
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) w[e] = pack_h2(a_val(r, ch * 8 + 2 * e), a_val(r, ch * 8 + 2 * e + 1));
        *reinterpret_cast<uint4*>(sm + swz_offset(r, ch, RB)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int idx = tid; idx < N * (RB / 16); idx += 128) {
        int n = idx / (RB / 16), ch = idx % (RB / 16);
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) w[e] = pack_h2(b_val(n, ch * 8 + 2 * e), b_val(n, ch * 8 + 2 * e + 1));
        *reinterpret_cast<uint4*>(sm + AROWS * 128 + swz_offset(n, ch, RB)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - base));
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint32_t arow = a_base + r0 * RB;
        const uint32_t bo = bo_mode ? ((arow >> 7) & 7u) : 0u;
        for (int ks = 0; ks < K / 16; ++ks)
            umma_f16(tmem, make_smem_desc(arow + ks * 32, RB, bo), make_smem_desc(b_base + ks * 32, RB, 0), idesc, ks > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tl + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(size_t)(32 * warp + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    const size_t smem = 1024 + AROWS * 128 + 256 * 128 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d_out;
    cudaMalloc(&d_out, 128 * 256 * sizeof(float));
    const int rbs[3] = {128, 64, 32};
    const int ns[3] = {16, 64, 256};
    const int r0s[7] = {0, 1, 3, 5, 8, 13, 50};
    int bad_mode0 = 0, bad_mode1 = 0;
    for (int rb : rbs)
        for (int N : ns)
            for (int r0 : r0s)
                for (int mode = 0; mode < 2; ++mode) {
                    cudaMemset(d_out, 0, 128 * 256 * sizeof(float));
                    probe_kernel<<<1, 128, smem>>>(d_out, rb, N, r0, mode);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("RB=%d N=%d r0=%d mode=%d CUDA error %s\n", rb, N, r0, mode, cudaGetErrorString(e)); return 1; }
                    std::vector<float> h(128 * N);
                    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                    double maxerr = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < N; ++n) {
                            double ref = 0;
                            for (int j = 0; j < rb / 2; ++j) ref += (double)a_val(r0 + m, j) * b_val(n, j);
                            double d = fabs(ref - h[(size_t)m * N + n]);
                            if (d > maxerr) maxerr = d;
                        }
                    printf("RB=%3d N=%3d r0=%2d base_offset_mode=%d maxerr=%g %s\n", rb, N, r0, mode, maxerr, maxerr == 0 ? "OK" : "MISMATCH");
                    if (maxerr != 0) (mode ? bad_mode1 : bad_mode0)++;
                }
    printf("SUMMARY mismatches: base_offset=0 -> %d, base_offset=(addr>>7)&7 -> %d\n", bad_mode0, bad_mode1);
    cudaFree(d_out);
    return 0;
}
