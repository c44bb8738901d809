// Internal launch interface between the C-ABI host code (api.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <functional>

namespace svb {

// Generic 1-D convolution descriptor for the fp32 FFMA kernel (kernels_f32.cu).
//   acc[b,co,t] = sum_{ci,tap} w[ci][tap][co] * act(x[b, x_c0+ci, t + tap*dil - pad_left])
//   n = t*ostride + ooff  (polyphase transposed convs write with a stride)
//   y[b, y_c0+co, n] = mask * maybe_tanh( alpha*(acc + bias[co] + bias_b[b,co] + res[b,res_c0+co,n]) + beta*y_old )
struct ConvF32 {
    const float* x = nullptr; int x_ctot = 0, x_c0 = 0, Cin = 0, Tin = 0;
    const float* w = nullptr;           // [n_phase][Cin][k][Cout]
    const float* bias = nullptr;        // [Cout] or null
    const float* bias_b = nullptr;      // per-batch bias, bias_b[b*bias_b_stride + bias_b_off + co] or null
    int bias_b_stride = 0, bias_b_off = 0;
    const float* bias_t = nullptr;      // time-varying bias [B, bias_t_ctot, Ty] (speaker-mix g) or null
    int bias_t_ctot = 0, bias_t_c0 = 0;
    const float* res = nullptr; int res_ctot = 0, res_c0 = 0;
    float* y = nullptr; int y_ctot = 0, y_c0 = 0, Ty = 0;
    int Cout = 0, k = 1, dil = 1, pad_left = 0;
    int n_out = 0, ostride = 1, ooff = 0;
    int n_phase = 1; long long w_phase_stride = 0;   // phase p: w += p*w_phase_stride, ooff += p
    int in_act = 0; float in_slope = 0.f;            // leaky-relu on the input
    float alpha = 1.f, beta = 0.f;
    int out_tanh = 0;
    const int32_t* lengths = nullptr;                // mask n < lengths[b]
    int B = 1;
};

void launch_conv_f32(const ConvF32& a, cudaStream_t st);

// acts[b,c,t] = tanh(a[b,c,t]) * sigmoid(a[b,c+H,t])   (commons.py:129-136 after the conditioning add)
void launch_gate(const float* a, float* acts, int B, int H, int T, cudaStream_t st);

// out[b,co] = bias[co] + sum_ci W[co][ci] * g[b,ci]     (1x1 conv on a length-1 conditioning vector)
void launch_gemv(const float* W, const float* bias, const float* g, float* out, int B, int Cout, int Cin, cudaStream_t st);

// y[b,co,t] += bias[co] + sum_kk w[co][kk] * har[b, t*s - p + kk]   (vdecoder/hifigan/models.py:343-348,380-382)
void launch_noise_conv_add(const float* har, const float* w, const float* bias, float* y,
                           int B, int Cout, int Tout, int N, int K, int s, int p, cudaStream_t st);

// y = SnakeAlias(x) per channel (vdecoder/hifiganwithsnake/alias/act.py:109-129); ealpha = e^alpha, inv_beta = 1/(e^beta+1e-9)
void launch_snake_alias(const float* x, float* y, const float* ealpha, const float* inv_beta, const float* filt,
                        int B, int C, int L, cudaStream_t st);

// wav[b,n] = tanh(bias + sum_{ci,k} w[ci][k] * lrelu(x[b,ci,n+k-pad], slope))   (conv_post, :390-392)
void launch_conv_post(const float* x, const float* w, float bias, float* wav, int B, int C, int N, int K, float slope, cudaStream_t st);

// NSF harmonic source (vdecoder/hifigan/models.py:250-271,307-320) in the closed form of SURVEY §9.7.
// rand_in_rate = 0: rand_ini is an initial phase (vdecoder/hifigan/models.py:147-150, sample-rate rad_values);
// rand_in_rate = 1: rand_ini is added to the FIRST FRAME's per-sample phase increment (vdecoder/nsf_hifigan/models.py:146-148,
//                   where rad_values is built at frame rate and then nearest-upsampled).
void launch_nsf_source(const float* f0, const float* rand_ini, const float* noise, const float* lin_w, float lin_b,
                       double* phase_ws /*[B,T,H]*/, float* har, int B, int T, int hop, int n_harm, float sr, int rand_in_rate,
                       cudaStream_t st, int philox = 0, unsigned long long seed = 0);   // philox: in-kernel N(0,1) when noise == null

// ---- tensor-core path (kernels_tc.cu) --------------------------------------------------------------
// One ResBlock "pair": y = x + conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 with out = alpha*y + beta*out_old.
struct PairTC {
    const float* x = nullptr;      // [B,C,T] fp32 residual stream in
    float* out = nullptr;          // [B,C,T]
    const void* w1 = nullptr;      // packed fp16 tap images (see pack_tc_weights in api.cu)
    const void* w2 = nullptr;
    const float* b1 = nullptr; const float* b2 = nullptr;
    int B = 1, C = 0, T = 0, k = 3, dil = 1;
    float alpha = 1.f, beta = 0.f;
    float inv = 1.f;               // 1 / (s1*s2): the weight images hold w1*s1, w2*s2 (power-of-two range normalisation), b1 = s1*bias1
    int variant = -1;              // -1: library default (env SVB_TC_VARIANT), else explicit tile variant
    const void* a16_in = nullptr;  // optional fp16 [B][T][C] copy of lrelu(x): the A tile is then loaded by TMA (tensor map)
    void* a16_out = nullptr;       // optional fp16 [B][T][C] copy of lrelu(out) for the next pair
};
bool pair_tc_supports_tma(int C, int variant);
int launch_pair_tc(const PairTC& a, cudaStream_t st);   // returns 0 or a negative status
int launch_pair_tc_multi(const PairTC* av, int nbr, cudaStream_t st);   // up to 3 independent pairs (same C, B, T) in one launch
size_t tc_weight_image_bytes(int C, int k);
// host-side: build the swizzled fp16 image for one conv (w_folded is [Cout][Cin][k] fp32)
void tc_pack_weight_image(const float* w_folded, int C, int k, void* dst_host, float scale = 1.f);

// ---- fused ResBlock1 (three pairs, one branch) for C <= 64 (kernels_resblock.cu) -------------------------------------
struct ResblockTC {
    const float* x = nullptr; float* out = nullptr;     // [B,C,T]
    const void* w[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // c1[d0], c2[d0], c1[d1], c2[d1], c1[d2], c2[d2]
    const float* bias[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int B = 1, C = 0, T = 0, k = 3;
    int dil[3] = {1, 3, 5};
    float alpha = 1.f, beta = 0.f;
    float inv[3] = {1.f, 1.f, 1.f};   // per pair: 1 / (scale of c1's image * scale of c2's image); bias[2d] is pre-multiplied by c1's scale
    int variant = -1;
};
int launch_resblock_tc(const ResblockTC& a, cudaStream_t st);

// ---- generic tensor-core convolution-as-GEMM (kernels_convn.cu): conv_pre, polyphase ups, flow WN layers ----------
struct ConvNSeg {                 // destination of a column range [col0, col1)
    float* y = nullptr; int y_ctot = 0, y_c0 = 0;
    int col0 = 0, col1 = 0;
    float alpha = 1.f, beta = 0.f;
    const float* res = nullptr; int res_ctot = 0, res_c0 = 0;
    int masked = 0;
};
struct ConvNTC {
    const float* x = nullptr; int x_ctot = 0, x_c0 = 0, cin_real = 0, Tin = 0;
    int cinp = 0;                  // padded Cin (template selector): 32,64,128,192,256,512
    // optional strided view of the input (0 = the default [B,C,T] tensor): element (c, t) = x[b*view_bstride + t*view_tstride
    // + c*view_cstride + view_off], zero outside [0, view_limit).  Used to read the 1-channel excitation as overlapping windows.
    long long view_bstride = 0; int view_tstride = 0, view_cstride = 0, view_off = 0; long long view_limit = 0;
    int in_act = 0; float in_slope = 0.f;
    const void* w = nullptr;       // fp16 image [chunk][tap][panel][NC rows][swizzled]
    const float* bias = nullptr;   // per column (column order), nullable
    const float* bias_b = nullptr; int bias_b_stride = 0, bias_b_off = 0;   // per-batch column bias
    int k = 1, dil = 1, pad_left = 0;
    int n_rows = 0;                // output rows (time steps of the GEMM's M dimension)
    int N_total = 0, NC = 0, chunks_per_cta = 1;
    int mode = 0;                  // 0 plain (column == channel), 1 polyphase (column = co*s + phase), 2 tanh*sigmoid gate,
                                   // 3 attention operands: columns = [q | k | v] x heads x dk(96); the epilogue writes fp16 shared-memory
                                   // IMAGES of the attention kernel's tiles (att_q / att_k: [b][head][tile][2 panels][128 rows][128 B],
                                   // att_v: [b][head][tile][2 panels][96 rows][128 B] = V^T), swizzled, ready for 1-D bulk copies
    void* att_q = nullptr; void* att_k = nullptr; void* att_v = nullptr; int att_heads = 0, att_tiles = 0;
    int s = 1, p = 0, Ty = 0;      // polyphase stride / padding; Ty = length of the output time axis
    ConvNSeg seg[2]; int n_seg = 1;
    const int32_t* lengths = nullptr;
    int B = 1;
    // optional fused noise conv (polyphase mode): excitation window har[b, i*noise_stride + noise_w0 + u], u in [0,16)
    const float* har = nullptr; int har_N = 0; int noise_stride = 0, noise_w0 = 0;
    int noise_wide = 0;            // 1: window of up to 80 samples (64-sample + 16-sample panels) instead of 16
    // optional second K chunk (input channels [k2_c0, k2_c0 + cinp) against the image w_k2) accumulated into the same
    // accumulators before the epilogue: a 768-channel input as two operand tiles of 384 (single column chunk only)
    const void* w_k2 = nullptr; int k2_c0 = 0;
    int out_relu = 0;              // plain mode: ReLU after everything else (FFN of enc_p, modules/attentions.py:343-346)
    float acc_scale = 1.f;         // the image holds w * 2^e (range normalisation at pack time); epilogues use acc * acc_scale
    // SnakeAlias fused into the loader (vdecoder/hifiganwithsnake/alias/act.py:109-129): A = SnakeAlias(x) per input channel
    const float* snake_ealpha = nullptr;   // [Cin] e^alpha          (null: plain / LeakyReLU loader)
    const float* snake_invbeta = nullptr;  // [Cin] 1/(e^beta + 1e-9)
    const float* snake_filt = nullptr;     // the 12-tap kaiser-sinc filter
    // time-varying bias (speaker-mix conditioning): bias_t[b, bias_t_c0 + column, row] added per (row, column); plain/gate modes
    const float* bias_t = nullptr; int bias_t_ctot = 0, bias_t_c0 = 0;
};
int launch_convn_tc(const ConvNTC& a, cudaStream_t st);
int convn_mb(int cinp);
int convn_ups_mb(int cinp);
int convn_snake_mb(int cinp);
size_t convn_weight_image_bytes(int cinp, int N_total, int NC, int k, int noise);
void convn_pack_weight_image(int cinp, int N_total, int NC, int k, const std::function<float(int, int, int)>& wcol,
                             const std::function<float(int, int)>* ncol, int noise, void* dst_host);

// ---- one ResidualCouplingLayer (reverse) per launch (kernels_flow.cu) ------------------------------------------------
struct FlowLayerTC {
    float* y = nullptr; int y_ctot = 0, in_c0 = 0, out_c0 = 0;     // [B, 2*half, T] updated in place
    const void* w = nullptr;                                          // block stream built by flow_layer_pack
    const float* bias_gate = nullptr;    // [L][2H] in_layers biases, chunk-permuted (flow_gate_row)
    const float* bias_h = nullptr;       // [L][H]  b_pre + sum_{j<i} b_res_j
    const float* bias_m = nullptr;       // [half]  W_post (sum of the skip biases) + b_post
    const float* gcond = nullptr;        // cond_layer(g) for g[B,gin,1], chunk-permuted, row b at gcond + b*gcond_bstride; or null
    int gcond_bstride = 0;
    const float* gcond_t = nullptr;      // [B][L*2H][T] for time-varying g (speaker mix); or null
    const int32_t* lengths = nullptr;
    int B = 1, T = 0, H = 192, half = 96, L = 4, k = 5;
};
int launch_flow_layer_tc(const FlowLayerTC& a, cudaStream_t st);
size_t flow_layer_image_bytes();
int flow_gate_row(int col);
void flow_layer_pack(const std::function<float(int, int)>& pre, const std::function<float(int, int, int, int)>& inl,
                     const std::function<float(int, int, int)>& rsm, void* dst_host);

// ---- enc_p attention (kernels_attn.cu): windowed relative-position MHA on channel-major [B,C,T] tensors --------------------
struct AttnTC {
    // fp16 tile images written by the q/k/v projection GEMM (ConvNTC mode 3): q_img / k_img [B][heads][tiles][2][128][128 B],
    // v_img [B][heads][tiles][2][96][128 B] (V^T), tiles = ceil(T / 128)
    const void* q_img = nullptr; const void* k_img = nullptr; const void* v_img = nullptr;
    const float* ek = nullptr; const float* ev = nullptr;     // emb_rel_k / emb_rel_v [2*window+1][dk]
    float* out = nullptr; int out_ctot = 0;                   // [B, heads*dk, T]
    const int32_t* lengths = nullptr;
    int B = 1, T = 0, heads = 2, dk = 96, window = 4;
};
int launch_attn_rel_tc(const AttnTC& a, cudaStream_t st);

// channel-major LayerNorm over C of [B,C,T] and the prior sample z = m + noise*exp(logs)*ns (kernels_prefix.cu)
void launch_ln_cm(const float* x, const float* gamma, const float* beta, float eps, float* y, int B, int C, int T, cudaStream_t st);
void launch_prior_sample(const float* stats, const float* noise, float ns, float* z, float* m_out, float* logs_out, int B, int C, int T, cudaStream_t st);

int64_t& launch_counter();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: track what has been granted per (kernel, device).
// `granted` is a function-local static array of SVB_MAX_DEV atomics (one per kernel instantiation); setting the attribute
// twice is idempotent, so concurrent contexts on different devices need no lock.
constexpr int SVB_MAX_DEV = 64;
template <typename K>
inline int ensure_dyn_smem(K kernel, size_t bytes, std::atomic<size_t>* granted) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SVB_MAX_DEV) return -1;
    if (bytes > granted[dev].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return -1;
        granted[dev].store(bytes, std::memory_order_release);
    }
    return 0;
}
// Number of SMs of the current device (persistent kernels size their grid by it); 148 if the query fails.
inline int sm_count() {
    static std::atomic<int> cached[SVB_MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SVB_MAX_DEV) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
int& sticky_launch_error();   // set by void launchers whose set-up failed (read + cleared by check_launch in api.cu)

}  // namespace svb
