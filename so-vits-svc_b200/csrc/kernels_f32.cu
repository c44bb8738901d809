// fp32 FFMA kernels for sm_100a: the strict-parity path (SVB_PREC_FP32) and the small bandwidth-bound
// pieces shared with the tensor-core path (NSF source, noise convs, conv_post, gate, conditioning GEMV).
//
// Compiled WITHOUT fast-math: the NSF source reproduces the reference's fp32 `(f0*h/sr) % 1`
// bit-for-bit (vdecoder/hifigan/models.py:144) and errors there are amplified by up to 4.4e5 samples.
#include "kernels.h"
#include <math.h>

namespace svb {

int64_t& launch_counter() { static int64_t c = 0; return c; }
int& sticky_launch_error() { static int e = 0; return e; }

// =====================================================================================================
// Generic 1-D convolution, fp32.  Block = 256 threads computes TCO output channels x TT time steps.
//   threads: (cg = tid / TX) selects 8 output channels, (tx = tid % TX) selects 4 time steps tx + TX*j.
//   Shared memory per Cin-chunk of CK channels: x tile [CK][TT+halo] (activation applied while loading)
//   and weight tile [CK][k][TCO]; lanes read consecutive x (conflict-free), weights are warp-broadcast.
// =====================================================================================================
constexpr int CONV_THREADS = 256;
constexpr int CONV_CK = 8;

template <int TCO>
__global__ void __launch_bounds__(CONV_THREADS) conv_f32_kernel(const ConvF32 a) {
    constexpr int G = TCO / 8;                 // channel groups
    constexpr int TX = CONV_THREADS / G;       // threads along time
    constexpr int TT = 4 * TX;                 // time tile
    extern __shared__ float smem[];
    const int halo = (a.k - 1) * a.dil;
    const int xs_pitch = TT + halo;
    float* xs = smem;                                   // [CK][xs_pitch]
    float* ws = smem + CONV_CK * xs_pitch;              // [CK][k][TCO]

    const int tid = threadIdx.x;
    const int cg = tid / TX, tx = tid % TX;
    const int n_cot = (a.Cout + TCO - 1) / TCO;
    const int phase = blockIdx.y / n_cot;
    const int co0 = (blockIdx.y % n_cot) * TCO;
    const int t0 = blockIdx.x * TT;
    const int b = blockIdx.z;
    const float* __restrict__ w = a.w + (long long)phase * a.w_phase_stride;
    const float* __restrict__ xb = a.x + ((long long)b * a.x_ctot + a.x_c0) * (long long)a.Tin;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ci0 = 0; ci0 < a.Cin; ci0 += CONV_CK) {
        // ---- stage x chunk
        for (int idx = tid; idx < CONV_CK * xs_pitch; idx += CONV_THREADS) {
            int c = idx / xs_pitch, j = idx - c * xs_pitch;
            int ti = t0 + j - a.pad_left;
            float v = 0.f;
            if (ci0 + c < a.Cin && ti >= 0 && ti < a.Tin) {
                v = __ldg(xb + (long long)(ci0 + c) * a.Tin + ti);
                if (a.in_act) v = v > 0.f ? v : v * a.in_slope;
            }
            xs[idx] = v;
        }
        // ---- stage weight chunk
        const int wn = CONV_CK * a.k * TCO;
        for (int idx = tid; idx < wn; idx += CONV_THREADS) {
            int co = idx % TCO;
            int rest = idx / TCO;              // c*k + tap
            int c = rest / a.k;
            float v = 0.f;
            if (ci0 + c < a.Cin && co0 + co < a.Cout)
                v = __ldg(w + ((long long)(ci0 * a.k + rest)) * a.Cout + co0 + co);
            ws[idx] = v;
        }
        __syncthreads();
        for (int c = 0; c < CONV_CK; ++c) {
            const float* xrow = xs + c * xs_pitch + tx;
            const float* wrow = ws + (c * a.k) * TCO + cg * 8;
            for (int tap = 0; tap < a.k; ++tap) {
                float xv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) xv[j] = xrow[TX * j + tap * a.dil];
                const float4 w0 = *reinterpret_cast<const float4*>(wrow + tap * TCO);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + tap * TCO + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue
    const int ooff = a.ooff + phase;
    const int len = a.lengths ? a.lengths[b] : 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + cg * 8 + i;
        if (co >= a.Cout) continue;
        float bsum = a.bias ? __ldg(a.bias + co) : 0.f;
        if (a.bias_b) bsum += __ldg(a.bias_b + (long long)b * a.bias_b_stride + a.bias_b_off + co);
        float* yrow = a.y + ((long long)b * a.y_ctot + a.y_c0 + co) * (long long)a.Ty;
        const float* rrow = a.res ? a.res + ((long long)b * a.res_ctot + a.res_c0 + co) * (long long)a.Ty : nullptr;
        const float* btrow = a.bias_t ? a.bias_t + ((long long)b * a.bias_t_ctot + a.bias_t_c0 + co) * (long long)a.Ty : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + tx + TX * j;
            if (t >= a.n_out) continue;
            const long long n = (long long)t * a.ostride + ooff;
            if (n < 0 || n >= a.Ty) continue;
            float v = acc[i][j] + bsum;
            if (btrow) v += __ldg(btrow + n);
            if (rrow) v += __ldg(rrow + n);
            v *= a.alpha;
            if (a.beta != 0.f) v = fmaf(a.beta, yrow[n], v);
            if (a.out_tanh) v = tanhf(v);
            if (n >= len) v = 0.f;
            yrow[n] = v;
        }
    }
}

template <int TCO>
static void launch_conv_t(const ConvF32& a, cudaStream_t st) {
    constexpr int G = TCO / 8, TX = CONV_THREADS / G, TT = 4 * TX;
    const int halo = (a.k - 1) * a.dil;
    size_t smem = sizeof(float) * (size_t)(CONV_CK * (TT + halo) + CONV_CK * a.k * TCO);
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (smem > 48 * 1024 && ensure_dyn_smem(conv_f32_kernel<TCO>, smem, granted)) { sticky_launch_error() = 1; return; }
    dim3 grid((a.n_out + TT - 1) / TT, ((a.Cout + TCO - 1) / TCO) * a.n_phase, a.B);
    conv_f32_kernel<TCO><<<grid, CONV_THREADS, smem, st>>>(a);
    launch_counter()++;
}

void launch_conv_f32(const ConvF32& a, cudaStream_t st) {
    if (a.Cout >= 64) launch_conv_t<64>(a, st);
    else if (a.Cout >= 32) launch_conv_t<32>(a, st);
    else launch_conv_t<16>(a, st);
}

// =====================================================================================================
// WN gate (modules/commons.py:129-136): acts = tanh(a[:, :H]) * sigmoid(a[:, H:])
// =====================================================================================================
__global__ void gate_kernel(const float* __restrict__ a, float* __restrict__ acts, int H, int T, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int t = (int)(i % T);
    long long r = i / T;
    int c = (int)(r % H);
    long long b = r / H;
    const float* base = a + (b * 2 * H) * (long long)T;
    float ta = base[(long long)c * T + t];
    float sa = base[(long long)(c + H) * T + t];
    acts[i] = tanhf(ta) * (1.f / (1.f + expf(-sa)));
}

void launch_gate(const float* a, float* acts, int B, int H, int T, cudaStream_t st) {
    long long total = (long long)B * H * T;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    gate_kernel<<<(unsigned)blocks, threads, 0, st>>>(a, acts, H, T, total);
    launch_counter()++;
}

// =====================================================================================================
// Conditioning GEMV: out[b,co] = bias[co] + W[co,:] . g[b,:]   (one warp per output)
// =====================================================================================================
__global__ void gemv_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ g,
                            float* __restrict__ out, int Cout, int Cin, int total) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= total) return;
    int b = warp / Cout, co = warp % Cout;
    const float* wr = W + (long long)co * Cin;
    const float* gr = g + (long long)b * Cin;
    float s = 0.f;
    for (int i = lane; i < Cin; i += 32) s = fmaf(wr[i], gr[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp] = s + (bias ? bias[co] : 0.f);
}

void launch_gemv(const float* W, const float* bias, const float* g, float* out, int B, int Cout, int Cin, cudaStream_t st) {
    int total = B * Cout;
    int threads = 256;
    int blocks = (total * 32 + threads - 1) / threads;
    gemv_kernel<<<blocks, threads, 0, st>>>(W, bias, g, out, Cout, Cin, total);
    launch_counter()++;
}

// =====================================================================================================
// noise_convs[i]: strided analysis filter of the 1-channel excitation, accumulated into the stage input.
// Block: 128 output steps x all channels.  The excitation window is staged as rows of `s` samples with
// an odd pitch (s+1) so that lanes (consecutive t) hit distinct banks.
// =====================================================================================================
// Block: 64 output steps x 32 channels (256 threads = 64 t x 4 groups of 8 channels).  The excitation window is
// staged as rows of `s` samples with an odd pitch (s+1) so lanes (consecutive t) hit distinct banks; the weight
// slice is staged transposed [kk][co] so each thread reads its 8 channels as two warp-broadcast float4.
__global__ void __launch_bounds__(256) noise_conv_add_kernel(const float* __restrict__ har, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             int Cout, int Tout, int N, int K, int s, int p) {
    extern __shared__ float sm[];
    constexpr int TT = 64, TC = 32;
    const int pitch = s + 1;
    const int rows = (K > 1) ? TT + 1 : TT;
    float* hs = sm;                                   // [rows][pitch]  row r = har[(t0+r)*s - p .. +s)
    float* wsm = sm + (((TT + 1) * pitch + 3) & ~3);  // [K][TC]
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * TT;
    const int co0 = blockIdx.y * TC;
    const float* hb = har + (long long)b * N;
    for (int idx = threadIdx.x; idx < rows * s; idx += blockDim.x) {
        int r = idx / s, c = idx - r * s;
        long long n = (long long)(t0 + r) * s - p + c;
        hs[r * pitch + c] = (n >= 0 && n < N) ? hb[n] : 0.f;
    }
    for (int idx = threadIdx.x; idx < K * TC; idx += blockDim.x) {
        int co = idx / K, kk = idx - co * K;            // coalesced read of w[co][kk]
        wsm[kk * TC + co] = (co0 + co < Cout) ? w[(long long)(co0 + co) * K + kk] : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x % TT, cg = threadIdx.x / TT;
    const int t = t0 + tx;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (co0 + cg * 8 + i < Cout) ? bias[co0 + cg * 8 + i] : 0.f;
    int r = tx, c = 0;
    for (int kk = 0; kk < K; ++kk) {
        const float h = hs[r * pitch + c];
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + kk * TC + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + kk * TC + cg * 8 + 4);
        acc[0] = fmaf(w0.x, h, acc[0]); acc[1] = fmaf(w0.y, h, acc[1]); acc[2] = fmaf(w0.z, h, acc[2]); acc[3] = fmaf(w0.w, h, acc[3]);
        acc[4] = fmaf(w1.x, h, acc[4]); acc[5] = fmaf(w1.y, h, acc[5]); acc[6] = fmaf(w1.z, h, acc[6]); acc[7] = fmaf(w1.w, h, acc[7]);
        if (++c == s) { c = 0; ++r; }
    }
    if (t >= Tout) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + cg * 8 + i;
        if (co < Cout) y[((long long)b * Cout + co) * Tout + t] += acc[i];
    }
}

void launch_noise_conv_add(const float* har, const float* w, const float* bias, float* y,
                           int B, int Cout, int Tout, int N, int K, int s, int p, cudaStream_t st) {
    constexpr int TT = 64, TC = 32;
    size_t smem = sizeof(float) * ((size_t)(((TT + 1) * (s + 1) + 3) & ~3) + (size_t)K * TC);
    dim3 grid((Tout + TT - 1) / TT, (Cout + TC - 1) / TC, B);
    noise_conv_add_kernel<<<grid, 256, smem, st>>>(har, w, bias, y, Cout, Tout, N, K, s, p);
    launch_counter()++;
}

// =====================================================================================================
// SnakeAlias (vdecoder/hifiganwithsnake/alias/act.py:109-129), one (batch, channel, 256-step tile) per
// block:  u = 2x kaiser-sinc upsample of replicate-padded x;  s = u + sin^2(u*e^alpha)/(e^beta + 1e-9);
//         y = 12-tap low-pass of replicate-padded s, decimated by 2.
// =====================================================================================================
__global__ void __launch_bounds__(256) snake_alias_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          const float* __restrict__ ealpha, const float* __restrict__ inv_beta,
                                                          const float* __restrict__ filt, int C, int L) {
    __shared__ float xs[272];
    __shared__ float ss[528];
    __shared__ float f[12];
    const int c = blockIdx.y, b = blockIdx.z;
    const int t0 = blockIdx.x * 256;
    const float* xr = x + ((long long)b * C + c) * L;
    if (threadIdx.x < 12) f[threadIdx.x] = filt[threadIdx.x];
    for (int j = threadIdx.x; j < 266; j += 256) {
        int ti = min(max(t0 - 5 + j, 0), L - 1);
        xs[j] = xr[ti];
    }
    __syncthreads();
    const float ea = ealpha[c], ib = inv_beta[c];
    for (int idx = threadIdx.x; idx < 522; idx += 256) {
        const int m = min(max(2 * t0 - 5 + idx, 0), 2 * L - 1);
        // u[m] = 2 * sum_i xp[i] * f[m + 15 - 2i],  xp[i] = x[clamp(i - 5)],  0 <= m + 15 - 2i <= 11
        const int i_lo = (m + 5) >> 1;            // ceil((m + 4) / 2)
        float u = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int i = i_lo + q;
            const int tap = m + 15 - 2 * i;
            if (tap >= 0 && tap <= 11) {
                const int xi = min(max(i - 5, 0), L - 1) - (t0 - 5);
                u = fmaf(xs[min(max(xi, 0), 265)], f[tap], u);
            }
        }
        u *= 2.f;
        const float sn = sinf(u * ea);
        ss[idx] = u + ib * sn * sn;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= L) return;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) acc = fmaf(f[j], ss[2 * threadIdx.x + j], acc);
    y[((long long)b * C + c) * L + t] = acc;
}

void launch_snake_alias(const float* x, float* y, const float* ealpha, const float* inv_beta, const float* filt,
                        int B, int C, int L, cudaStream_t st) {
    dim3 grid((L + 255) / 256, C, B);
    snake_alias_kernel<<<grid, 256, 0, st>>>(x, y, ealpha, inv_beta, filt, C, L);
    launch_counter()++;
}

// =====================================================================================================
// conv_post: leaky_relu(slope) -> Conv1d(C->1, K, pad (K-1)/2) -> tanh
// =====================================================================================================
// 128 threads x 4 consecutive outputs: the lrelu'd input window of a thread (4 + K - 1 <= 12 samples per channel) is read
// from shared memory with three vector loads and reused for 4*K FMAs, so the kernel is bound by the single HBM pass over
// the stage-4 tensor instead of by shared-memory loads (was 2 LDS per FMA).
constexpr int CP_TT = 512, CP_PITCH = CP_TT + 8, CP_MAXK = 9;
__global__ void __launch_bounds__(128) conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias,
                                                        float* __restrict__ wav, int C, int N, int K, float slope) {
    extern __shared__ __align__(16) float sm[];
    const int pad = (K - 1) / 2;
    float* xs = sm;                  // [C][CP_PITCH]
    float* wsm = sm + C * CP_PITCH;  // [C][K]
    const int b = blockIdx.y, n0 = blockIdx.x * CP_TT;
    const float* xb = x + (long long)b * C * N;
    for (int c = 0; c < C; ++c) {
        const float* xc = xb + (long long)c * N;
        for (int j = threadIdx.x; j < CP_PITCH; j += 128) {
            const int n = n0 + j - pad;
            const float v = (n >= 0 && n < N) ? __ldg(xc + n) : 0.f;
            xs[c * CP_PITCH + j] = fmaxf(v, v * slope);
        }
    }
    for (int idx = threadIdx.x; idx < C * K; idx += 128) wsm[idx] = w[idx];
    __syncthreads();
    const int o = 4 * threadIdx.x;
    float acc[4] = {bias, bias, bias, bias};
    for (int c = 0; c < C; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(xs + c * CP_PITCH + o);
        const float4 a1 = *reinterpret_cast<const float4*>(xs + c * CP_PITCH + o + 4);
        const float4 a2 = *reinterpret_cast<const float4*>(xs + c * CP_PITCH + o + 8);
        const float win[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
        for (int k = 0; k < CP_MAXK; ++k) {
            if (k < K) {
                const float wk = wsm[c * K + k];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wk, win[i + k], acc[i]);
            }
        }
    }
    const int n = n0 + o;
    float* dst = wav + (long long)b * N + n;
    if (n + 3 < N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (n + i < N) dst[i] = tanhf(acc[i]);
    }
}

// Register-window form for the shipped shape (C = 16 channels, K = 7, N % 4 == 0): a thread produces 4 consecutive samples
// and reads, per channel, the aligned 12-sample window [n0-4, n0+8) as three 128-bit loads straight from global memory
// (neighbouring threads share two of them through L1).  Four channels = 12 LDG.128 are in flight per thread, no shared
// memory, no block barrier: the kernel is bound by the single HBM pass over the stage-4 tensor.
template <int C, int K>
__global__ void __launch_bounds__(256) conv_post_vec_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias,
                                                            float* __restrict__ wav, int N, float slope) {
    constexpr int PAD = (K - 1) / 2;
    static_assert(PAD <= 4 && K - 1 - PAD <= 4, "window is [n0-4, n0+8)");
    __shared__ float wsm[C * K];
    for (int i = threadIdx.x; i < C * K; i += 256) wsm[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int n0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (n0 >= N) return;
    const float* __restrict__ xb = x + (size_t)b * C * N;
    float acc[4] = {bias, bias, bias, bias};
    const bool inner = (n0 >= 4) && (n0 + 8 <= N);
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 4) {
        float win[4][12];
        if (inner) {
            float4 v[4][3];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int u = 0; u < 3; ++u) v[cc][u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(c0 + cc) * N + n0 - 4 + 4 * u));
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int u = 0; u < 3; ++u) { win[cc][4 * u] = v[cc][u].x; win[cc][4 * u + 1] = v[cc][u].y; win[cc][4 * u + 2] = v[cc][u].z; win[cc][4 * u + 3] = v[cc][u].w; }
        } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int j = 0; j < 12; ++j) { const int n = n0 - 4 + j; win[cc][j] = (n >= 0 && n < N) ? __ldg(xb + (size_t)(c0 + cc) * N + n) : 0.f; }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
            for (int j = 0; j < 12; ++j) win[cc][j] = fmaxf(win[cc][j], win[cc][j] * slope);   // slope <= 1: LeakyReLU (1 = identity)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float wk = wsm[(c0 + cc) * K + k];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wk, win[cc][4 - PAD + i + k], acc[i]);
            }
        }
    }
    float* dst = wav + (size_t)b * N + n0;
    if (n0 + 3 < N) *reinterpret_cast<float4*>(dst) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
    else
        for (int i = 0; i < 4; ++i) if (n0 + i < N) dst[i] = tanhf(acc[i]);
}

void launch_conv_post(const float* x, const float* w, float bias, float* wav, int B, int C, int N, int K, float slope, cudaStream_t st) {
    if (C == 16 && K == 7 && (N % 4) == 0 && slope <= 1.f && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wav)) & 15u) == 0) {
        dim3 grid((N / 4 + 255) / 256, B);
        conv_post_vec_kernel<16, 7><<<grid, 256, 0, st>>>(x, w, bias, wav, N, slope);
        launch_counter()++;
        return;
    }
    size_t smem = sizeof(float) * ((size_t)C * CP_PITCH + (size_t)C * K);
    dim3 grid((N + CP_TT - 1) / CP_TT, B);
    conv_post_kernel<<<grid, 128, smem, st>>>(x, w, bias, wav, C, N, K, slope);
    launch_counter()++;
}

// =====================================================================================================
// NSF source.  f0 is constant inside a hop (nn.Upsample nearest, vdecoder/hifigan/models.py:330,369), so the
// reference's doubly-wrapped cumsum over N samples (:160-166) collapses to
//     phase[b, hop*F+k, h] = rand_ini[b,h] + sum_{f<F} hop*r[b,f,h] + (k+1)*r[b,F,h]   (mod 1)
// with r = ((f0*(h+1))/sr) % 1 evaluated in fp32 exactly as :144 does, then promoted to fp64.
// Only the T-long prefix sum is serial (kernel 1, fp64); kernel 2 is one thread per output sample.
// =====================================================================================================
__device__ __forceinline__ float rad_value(float f0, int h, float sr) {
    float fn = __fmul_rn(f0, (float)(h + 1));
    float q = __fdiv_rn(fn, sr);
    return fmodf(q, 1.0f);
}

// one warp per (batch, harmonic): lanes own contiguous chunks of frames, fp64 exclusive scan across lanes
__device__ __forceinline__ double rad_frame(const float* __restrict__ f0row, const float* __restrict__ ri, int t, int h, float sr, int rand_in_rate) {
    float r = rad_value(f0row[t], h, sr);
    if (rand_in_rate && t == 0 && h > 0) r = __fadd_rn(r, ri[h]);      // fp32 add like rad_values[:, 0, :] += rand_ini
    return (double)r;
}

__global__ void nsf_phase_kernel(const float* __restrict__ f0, const float* __restrict__ rand_ini, double* __restrict__ phase,
                                 int B, int T, int H, int hop, float sr, int rand_in_rate) {
    const int i = blockIdx.x;                 // b*H + h
    const int lane = threadIdx.x;
    const int b = i / H, h = i % H;
    const int ch = (T + 31) / 32;
    const int t_lo = lane * ch, t_hi = min(T, t_lo + ch);
    const float* f0row = f0 + (long long)b * T;
    const float* ri = rand_ini + b * H;
    double local = 0.0;
    for (int t = t_lo; t < t_hi; ++t) local += rad_frame(f0row, ri, t, h, sr, rand_in_rate) * (double)hop;
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    double acc = (incl - local) + ((h == 0 || rand_in_rate) ? 0.0 : (double)ri[h]);
    acc -= floor(acc);
    for (int t = t_lo; t < t_hi; ++t) {
        phase[((long long)b * T + t) * H + h] = acc;
        acc += rad_frame(f0row, ri, t, h, sr, rand_in_rate) * (double)hop;
        acc -= floor(acc);
    }
}

template <int H>
__global__ void __launch_bounds__(256) nsf_source_kernel(const float* __restrict__ f0, const float* __restrict__ noise,
                                                         const double* __restrict__ phase, const float* __restrict__ lin_w, float lin_b,
                                                         float* __restrict__ har, int T, int hop, float sr, long long N,
                                                         const float* __restrict__ rand_ini, int rand_in_rate) {
    const int b = blockIdx.y;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int F = (int)(n / hop);
    const int k = (int)(n - (long long)F * hop);
    const float f = f0[(long long)b * T + F];
    const float uv = f > 0.f ? 1.f : 0.f;
    const float amp = uv * 0.003f + (1.f - uv) * (0.1f / 3.f);
    const double* ph0 = phase + ((long long)b * T + F) * H;
    const float* nz = noise ? noise + ((long long)b * N + n) * H : nullptr;
    float acc = lin_b;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float rf = rad_value(f, h, sr);
        if (rand_in_rate && F == 0 && h > 0) rf = __fadd_rn(rf, rand_ini[b * H + h]);
        double r = (double)rf;
        double ph = ph0[h] + (double)(k + 1) * r;
        ph -= floor(ph);
        float s = sinpif(2.0f * (float)ph) * 0.1f;
        float v = s * uv + (nz ? amp * nz[h] : 0.f);
        acc = fmaf(lin_w[h], v, acc);
    }
    har[(long long)b * N + n] = tanhf(acc);
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): counter-based, so every thread derives its
// 4 x H normals from (seed, its sample-group index, call number) alone - the "throughput mode" of SURVEY §2a: no [B,N,H]
// noise tensor is written by torch and read back (127 MB at config 2).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// two uniform 32-bit integers -> two independent standard normals (Box-Muller; u1 in (0,1], never 0)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;      // (a + 1) / 2^32
    const float u2 = (float)b * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z0 = r * cs; z1 = r * sn;
}

// Four consecutive samples per thread (hop % 4 == 0, so they share their frame): the 4 x H noise values are H 128-bit loads
// issued up front (the [B,N,H] noise tensor is 9/10 of the kernel's traffic), the frame's phase base and r are read once,
// the waveform leaves as one 128-bit store.
template <int H>
__global__ void __launch_bounds__(256) nsf_source_vec4_kernel(const float* __restrict__ f0, const float* __restrict__ noise,
                                                              const double* __restrict__ phase, const float* __restrict__ lin_w, float lin_b,
                                                              float* __restrict__ har, int T, int hop, float sr, long long N,
                                                              const float* __restrict__ rand_ini, int rand_in_rate,
                                                              int philox, unsigned long long seed) {
    const int b = blockIdx.y;
    const long long n0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (n0 >= N) return;
    float nz[4 * H];
    if (philox) {
        const unsigned long long grp = ((unsigned long long)b * (unsigned long long)N + (unsigned long long)n0) >> 2;
#pragma unroll
        for (int j = 0; j < H; ++j) {                              // call j -> flat noise elements 4j .. 4j+3 ([sample][harmonic] order)
            uint32_t c[4] = {(uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)j, 0x6e736673u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            box_muller(c[0], c[1], nz[4 * j], nz[4 * j + 1]);
            box_muller(c[2], c[3], nz[4 * j + 2], nz[4 * j + 3]);
        }
    } else if (noise) {
        const float4* __restrict__ np = reinterpret_cast<const float4*>(noise + ((long long)b * N + n0) * H);
#pragma unroll
        for (int j = 0; j < H; ++j) { const float4 q = __ldg(np + j); nz[4 * j] = q.x; nz[4 * j + 1] = q.y; nz[4 * j + 2] = q.z; nz[4 * j + 3] = q.w; }
    }
    const int F = (int)(n0 / hop);
    const int k0 = (int)(n0 - (long long)F * hop);
    const float f = f0[(long long)b * T + F];
    const float uv = f > 0.f ? 1.f : 0.f;
    const float amp = uv * 0.003f + (1.f - uv) * (0.1f / 3.f);
    const double* __restrict__ ph0 = phase + ((long long)b * T + F) * H;
    float acc[4] = {lin_b, lin_b, lin_b, lin_b};
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float rf = rad_value(f, h, sr);
        if (rand_in_rate && F == 0 && h > 0) rf = __fadd_rn(rf, rand_ini[b * H + h]);
        // The phase of the FIRST of the four samples is formed in fp64 exactly as in nsf_source_kernel (it may be thousands of
        // cycles); the other three add s*r <= 0.75 cycles in fp32 (<= 2e-7 cycles of rounding), and the sine of the reduced
        // argument in [-pi, pi] is one MUFU (|err| <= 2^-20.9).  Worst case 5e-7 on the waveform against the fp64 closed form
        // (the parity bound is 5e-6); 4x fewer fp64 instructions and ~10 fewer fp32 ones per sample and harmonic - the
        // kernel used to be bound by their issue rate (0.22 of the HBM roof).
        double ph = ph0[h] + (double)(k0 + 1) * (double)rf;
        ph -= floor(ph);
        const float phf = (float)ph;
        const float lws = lin_w[h] * 0.1f * uv;
        const float lwn = lin_w[h] * amp;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float x = fmaf((float)s, rf, phf);
            x -= rintf(x);
            acc[s] = fmaf(lws, __sinf(6.283185307179586f * x), acc[s]);
            if (noise || philox) acc[s] = fmaf(lwn, nz[s * H + h], acc[s]);
        }
    }
    *reinterpret_cast<float4*>(har + (long long)b * N + n0) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
}

void launch_nsf_source(const float* f0, const float* rand_ini, const float* noise, const float* lin_w, float lin_b,
                       double* phase_ws, float* har, int B, int T, int hop, int n_harm, float sr, int rand_in_rate, cudaStream_t st,
                       int philox, unsigned long long seed) {
    nsf_phase_kernel<<<B * n_harm, 32, 0, st>>>(f0, rand_ini, phase_ws, B, T, n_harm, hop, sr, rand_in_rate);
    launch_counter()++;
    long long N = (long long)T * hop;
    const bool vec = (hop % 4 == 0) && ((reinterpret_cast<uintptr_t>(har) & 15u) == 0) && (!noise || (reinterpret_cast<uintptr_t>(noise) & 15u) == 0);
    if (vec && (n_harm == 9 || n_harm == 1)) {
        dim3 gridv((unsigned)((N / 4 + 255) / 256), B);
        if (n_harm == 9)
            nsf_source_vec4_kernel<9><<<gridv, 256, 0, st>>>(f0, noise, phase_ws, lin_w, lin_b, har, T, hop, sr, N, rand_ini, rand_in_rate, philox && !noise, seed);
        else
            nsf_source_vec4_kernel<1><<<gridv, 256, 0, st>>>(f0, noise, phase_ws, lin_w, lin_b, har, T, hop, sr, N, rand_ini, rand_in_rate, philox && !noise, seed);
        launch_counter()++;
        return;
    }
    if (philox && !noise) { sticky_launch_error() = 1; return; }     // the in-kernel generator lives in the vectorised kernel only
    dim3 grid((unsigned)((N + 255) / 256), B);
    if (n_harm == 9)
        nsf_source_kernel<9><<<grid, 256, 0, st>>>(f0, noise, phase_ws, lin_w, lin_b, har, T, hop, sr, N, rand_ini, rand_in_rate);
    else
        nsf_source_kernel<1><<<grid, 256, 0, st>>>(f0, noise, phase_ws, lin_w, lin_b, har, T, hop, sr, N, rand_ini, rand_in_rate);
    launch_counter()++;
}

}  // namespace svb
