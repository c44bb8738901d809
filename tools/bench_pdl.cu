// Round-2 preparation: what does programmatic dependent launch (PDL) buy for a chain of short dependent kernels like the
// flow's 41 conv launches (~20 us each, under-filled grids)?  Each kernel has a data-independent prologue (shared-memory
// fill standing in for barrier init / TMEM alloc / weight prefetch) and a body that reads the previous kernel's output.
//   mode 0: plain stream-ordered launches
//   mode 1: cudaLaunchKernelEx with programmaticStreamSerializationAllowed, griddepcontrol.launch_dependents at kernel start,
//           griddepcontrol.wait before the first dependent read
// Prints microseconds per kernel for both modes and checks the result of the chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench_pdl tools/bench_pdl.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <bool PDL>
__global__ void __launch_bounds__(256) stage_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w, int n, int prologue_iters) {
    extern __shared__ float sm[];
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // prologue: independent of `in` (weights are constants)
    float acc = 0.f;
    for (int it = 0; it < prologue_iters; ++it)
        for (int i = threadIdx.x; i < 2048; i += 256) { sm[i] = w[(i + it) & 2047]; acc += sm[i]; }
    __syncthreads();
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    // body: depends on the previous kernel's output
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) out[i] = in[i] * 0.5f + sm[i & 2047] + (acc == 12345.f ? 1.f : 0.f);
}

int main() {
    const int n = 1 << 20, chain = 200, grid = 56;
    float *a, *b, *w;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&w, 2048 * 4);
    std::vector<float> hw(2048, 0.25f), ha(n, 1.f);
    cudaMemcpy(w, hw.data(), 2048 * 4, cudaMemcpyHostToDevice);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int prologue = 1; prologue <= 16; prologue *= 4) {
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaMemcpyAsync(a, ha.data(), n * 4, cudaMemcpyHostToDevice, st);
                cudaEventRecord(e0, st);
                float *src = a, *dst = b;
                for (int k = 0; k < chain; ++k) {
                    if (mode == 0) {
                        stage_kernel<false><<<grid, 256, 8192, st>>>(src, dst, w, n, prologue);
                    } else {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 8192; cfg.stream = st;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        const float* s2 = src; float* d2 = dst; const float* w2 = w; int n2 = n, p2 = prologue;
                        cudaLaunchKernelEx(&cfg, stage_kernel<true>, s2, d2, w2, n2, p2);
                    }
                    float* t = src; src = dst; dst = t;
                }
                cudaEventRecord(e1, st);
                cudaError_t e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
                if (rep == 0) {
                    float h0; cudaMemcpy(&h0, src, 4, cudaMemcpyDeviceToHost);
                    // x_{k+1} = x_k/2 + 0.25  ->  converges to 0.5
                    if (h0 < 0.49f || h0 > 0.51f) { printf("mode %d: wrong result %f\n", mode, h0); return 1; }
                }
            }
            printf("prologue_iters=%2d mode=%s: %.2f us per kernel (chain of %d, grid %d)\n", prologue, mode ? "PDL " : "plain", best * 1000.f / chain, chain, grid);
        }
    }
    return 0;
}
