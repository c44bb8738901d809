// Fused element-wise tails of the enc_p transformer layers (reference modules/attentions.py:95-106,
// 317-363 and modules/modules.py:23-35).  The GEMMs of the prefix stay on cuBLAS; what is fused here is the chain of
// small memory-bound kernels around them, on time-major [B,L,C] fp32 activations:
//   add_ln_im2col : y = LayerNorm_C(x + r);  cols[b,l,t*C+c] = y[b, l+t-(k-1)/2, c]  (zero outside [0,L))
//                   = residual add + LayerNorm + F.pad + the k shifted views concatenated for the FFN's first conv
//   ffn_tail      : y = LayerNorm_C(x + bias + sum_t ya[b, l+t-(k-1)/2, t*C + c])
//                   = the shift-and-add of the FFN's second conv run as one GEMM against the k stacked tap matrices,
//                     + bias + residual add + LayerNorm
//   rel_softmax   : band add of the relative-key logits + softmax + band extraction of p  (one block per attention row)
//   attn_merge    : p@v + relative-value term, heads merged back to [B,L,H*dk]
// One warp per (b,l) row, mean / variance by two-pass warp reductions in fp32 (torch: same statistics, eps inside sqrt).
#include "kernels.h"
#include "../../include/sovits_b200.h"

namespace svb {
namespace {

constexpr int PF_MAXV = 8;      // values per lane: C <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) add_ln_im2col_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y, float* __restrict__ cols,
                                                            int B, int L, int C, int k) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)B * L) return;
    const int l = (int)(row % L);
    const int nv = (C + 31) / 32;
    float v[PF_MAXV];
    const float* xr = x + row * C;
    const float* rr = r + row * C;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) {
        const int c = lane + 32 * i;
        v[i] = (i < nv && c < C) ? xr[c] + rr[c] : 0.f;
    }
    // channels beyond C contribute 0 to the sums only if they are excluded: handle C % 32 != 0 by masking
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) if (lane + 32 * i < C) s += v[i];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) if (lane + 32 * i < C) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    const int h = (k - 1) / 2;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
            const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
            y[row * C + c] = o;
            if (cols) {
                // y[l] is tap t of output row l - t + h
                for (int t = 0; t < k; ++t) {
                    const int lo = l - t + h;
                    if (lo >= 0 && lo < L) cols[(row + (lo - l)) * (long long)(k * C) + t * C + c] = o;
                }
                // zero padding: taps of this output row that fall outside [0,L)
                for (int t = 0; t < k; ++t) {
                    const int li = l + t - h;
                    if (li < 0 || li >= L) cols[row * (long long)(k * C) + t * C + c] = 0.f;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) ffn_tail_kernel(const float* __restrict__ ya, const float* __restrict__ x, const float* __restrict__ bias,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                       int B, int L, int C, int k) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)B * L) return;
    const int l = (int)(row % L);
    const int h = (k - 1) / 2;
    float v[PF_MAXV];
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) {
        const int c = lane + 32 * i;
        float a = 0.f;
        if (c < C) {
            a = x[row * C + c] + bias[c];
            for (int t = 0; t < k; ++t) {
                const int li = l + t - h;
                if (li >= 0 && li < L) a += ya[(row + (li - l)) * (long long)(k * C) + t * C + c];
            }
        }
        v[i] = a;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) if (lane + 32 * i < C) s += v[i];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) if (lane + 32 * i < C) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < PF_MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) y[row * C + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
}


// scores[row, j] += relk[row, j - i + w] on the (2w+1)-wide band (attentions.py:246-250, the relative-key logits), softmax
// over j (attentions.py:259), probabilities written back in place and the band of p returned for the relative-value term
// (attentions.py:262-266).  One block per row (b,h,i); the row lives in shared memory.
__global__ void __launch_bounds__(256) rel_softmax_kernel(float* __restrict__ scores, const float* __restrict__ relk, float* __restrict__ pband, int L, int w) {
    extern __shared__ float srow[];
    __shared__ float red[8];
    const long long row = blockIdx.x;
    const int i = (int)(row % L);
    const int nb = 2 * w + 1;
    float* sr = scores + row * L;
    const float* rk = relk + row * nb;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float m = -INFINITY;
    for (int j = tid; j < L; j += 256) {
        float v = sr[j];
        const int r = j - i + w;
        if (r >= 0 && r < nb) v += rk[r];
        srow[j] = v;
        m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[wid] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q]);
    __syncthreads();
    float s = 0.f;
    for (int j = tid; j < L; j += 256) { const float e = __expf(srow[j] - m); srow[j] = e; s += e; }
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q];
    const float inv = 1.f / s;
    for (int j = tid; j < L; j += 256) sr[j] = srow[j] * inv;
    if (tid < nb) {
        const int j = i + tid - w;
        pband[row * nb + tid] = (j >= 0 && j < L) ? srow[j] * inv : 0.f;
    }
}

// y[b,l,h*dk+d] = out[b,h,l,d] + sum_r pband[b,h,l,r] * embv[r,d]: the relative-value term (attentions.py:262-266) added to
// p@v and the heads merged back to time-major [B,L,H*dk] (attentions.py:267) in one pass.
__global__ void __launch_bounds__(256) attn_merge_kernel(const float* __restrict__ out, const float* __restrict__ pband, const float* __restrict__ embv,
                                                         float* __restrict__ y, int B, int H, int L, int dk, int nb) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)B * L * H * dk;
    if (idx >= total) return;
    const int d = (int)(idx % dk);
    const int h = (int)((idx / dk) % H);
    const int l = (int)((idx / ((long long)dk * H)) % L);
    const int b = (int)(idx / ((long long)dk * H * L));
    const long long rowp = ((long long)b * H + h) * L + l;
    float a = out[rowp * dk + d];
    for (int r = 0; r < nb; ++r) a = fmaf(pband[rowp * nb + r], embv[r * dk + d], a);
    y[idx] = a;
}

}  // namespace

// ---- channel-major pieces of the own-kernel enc_p (svb_enc_p): LayerNorm over channels of [B,C,T] and the prior sample ----
// Block = 32 frames x 8 warps; warp w holds channels w, w+8, ... of its 32 frames in registers (lanes = frames, so every
// global access is a coalesced 128-byte row segment), statistics are combined across the 8 warps through shared memory.
// modules/modules.py:23-35 (F.layer_norm over the channel axis, eps inside the square root).
constexpr int LN_MAXC = 256;
__global__ void __launch_bounds__(256) ln_cm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float eps, float* __restrict__ y, int C, int T) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int t = blockIdx.x * 32 + lane;
    const bool tv = t < T;
    const float* __restrict__ xb = x + (size_t)b * C * T + (tv ? t : 0);
    float v[LN_MAXC / 8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXC / 8; ++i) {
        const int c = w + 8 * i;
        v[i] = (tv && c < C) ? __ldg(xb + (size_t)c * T) : 0.f;
        s += v[i];
    }
    red[w][lane] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) mean += red[j][lane];
    mean /= (float)C;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXC / 8; ++i) { const int c = w + 8 * i; if (c < C) { const float d = v[i] - mean; q += d * d; } }
    red[w][lane] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) var += red[j][lane];
    const float rstd = rsqrtf(var / (float)C + eps);
    if (!tv) return;
    float* __restrict__ yb = y + (size_t)b * C * T + t;
#pragma unroll
    for (int i = 0; i < LN_MAXC / 8; ++i) {
        const int c = w + 8 * i;
        if (c < C) yb[(size_t)c * T] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
}

void launch_ln_cm(const float* x, const float* gamma, const float* beta, float eps, float* y, int B, int C, int T, cudaStream_t st) {
    dim3 grid((T + 31) / 32, B);
    ln_cm_kernel<<<grid, 256, 0, st>>>(x, gamma, beta, eps, y, C, T);
    launch_counter()++;
}

// z = m + noise * exp(logs) * noice_scale with stats = [m | logs] ([B, 2C, T]) - models.py:158-160 (mask = ones)
__global__ void prior_sample_kernel(const float* __restrict__ stats, const float* __restrict__ noise, float ns, float* __restrict__ z,
                                    float* __restrict__ m_out, float* __restrict__ logs_out, int C, int T, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    const long long r = i / T;
    const int c = (int)(r % C);
    const long long b = r / C;
    const float m = stats[((b * 2 * C) + c) * (long long)T + t];
    const float lg = stats[((b * 2 * C) + C + c) * (long long)T + t];
    z[i] = m + noise[i] * expf(lg) * ns;
    if (m_out) m_out[i] = m;
    if (logs_out) logs_out[i] = lg;
}

void launch_prior_sample(const float* stats, const float* noise, float ns, float* z, float* m_out, float* logs_out, int B, int C, int T, cudaStream_t st) {
    const long long total = (long long)B * C * T;
    prior_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(stats, noise, ns, z, m_out, logs_out, C, T, total);
    launch_counter()++;
}

}  // namespace svb

extern "C" {

int svb_prefix_add_ln_im2col(const float* x, const float* r, const float* gamma, const float* beta, float eps, float* y, float* cols, int B, int L, int C,
                             int k, void* stream) {
    if (!x || !r || !gamma || !beta || !y || B <= 0 || L <= 0 || C <= 0 || C > 32 * svb::PF_MAXV || k < 1 || k > 7) return SVB_ERR_INVALID_ARG;
    const long long rows = (long long)B * L;
    svb::add_ln_im2col_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, r, gamma, beta, eps, y, cols, B, L, C, k);
    svb::launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

int svb_prefix_ffn_tail(const float* ya, const float* x, const float* bias, const float* gamma, const float* beta, float eps, float* y, int B, int L, int C,
                        int k, void* stream) {
    if (!ya || !x || !bias || !gamma || !beta || !y || B <= 0 || L <= 0 || C <= 0 || C > 32 * svb::PF_MAXV || k < 1 || k > 7) return SVB_ERR_INVALID_ARG;
    const long long rows = (long long)B * L;
    svb::ffn_tail_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(ya, x, bias, gamma, beta, eps, y, B, L, C, k);
    svb::launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

int svb_prefix_rel_softmax(float* scores, const float* relk, float* pband, int rows, int L, int window, void* stream) {
    if (!scores || !relk || !pband || rows <= 0 || L <= 0 || window < 0 || 2 * window + 1 > 256 || L > 12000) return SVB_ERR_INVALID_ARG;
    svb::rel_softmax_kernel<<<(unsigned)rows, 256, (size_t)L * sizeof(float), static_cast<cudaStream_t>(stream)>>>(scores, relk, pband, L, window);
    svb::launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

int svb_prefix_attn_merge(const float* out, const float* pband, const float* embv, float* y, int B, int H, int L, int dk, int nb, void* stream) {
    if (!out || !pband || !embv || !y || B <= 0 || H <= 0 || L <= 0 || dk <= 0 || nb <= 0) return SVB_ERR_INVALID_ARG;
    const long long total = (long long)B * L * H * dk;
    svb::attn_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, pband, embv, y, B, H, L, dk, nb);
    svb::launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

}  // extern "C"
