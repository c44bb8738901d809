/*
 * sovits_b200.h — C ABI of libsovits_b200.so (sm_100a only; no CPU fallback).
 *
 * The reference (svc-develop-team/so-vits-svc) has no FFI of its own: its "plugin surface" is the
 * Python class models.SynthesizerTrn (models.py:339-532) and the vdecoder Generator it owns
 * (vdecoder/hifigan/models.py:323-403).  Each entry point below replaces one slice of that surface;
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions:
 *   - every function returns 0 (SVB_OK) or a negative svb_status; nothing throws or aborts;
 *   - all activation buffers are CALLER-OWNED DEVICE pointers to contiguous fp32 [B,C,T] tensors
 *     (torch: tensor.data_ptr()); the *_host entry point is the only one taking host pointers;
 *   - every launch goes on the caller's stream; the library never calls cudaDeviceSynchronize
 *     (the *_host variant synchronises the stream it is given, because it returns host data);
 *   - one context per device; a context is not re-entrant (matches the single-threaded callers,
 *     inference/infer_tool.py:116-496);
 *   - weights are handed over once in the reference's own checkpoint layout (weight_g / weight_v
 *     pairs, utils.py:155-187); the library folds weight-norm and packs for its kernels.
 */
#ifndef SOVITS_B200_H
#define SOVITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SVB_API __attribute__((visibility("default")))
#else
#define SVB_API
#endif

typedef struct svb_ctx svb_ctx;

typedef enum svb_status {
    SVB_OK = 0,
    SVB_ERR_INVALID_ARG = -1,
    SVB_ERR_CUDA = -2,
    SVB_ERR_NOT_LOADED = -3,       /* weights not loaded yet                              */
    SVB_ERR_MISSING_TENSOR = -4,   /* a required state_dict key was not supplied          */
    SVB_ERR_SHAPE = -5,            /* tensor shape disagrees with svb_model_cfg           */
    SVB_ERR_UNSUPPORTED = -6,      /* configuration outside the implemented hot path      */
    SVB_ERR_WORKSPACE = -7,        /* caller workspace too small                          */
    SVB_ERR_ARCH = -8              /* device is not sm_100 (B200)                          */
} svb_status;

/* Precision of the convolution arithmetic.
 *   SVB_PREC_FP32 : fp32 FFMA kernels (bit-for-bit independent of tensor cores; strict parity).
 *   SVB_PREC_TC   : tcgen05 tensor-core kernels, fp16 operands (11-bit significand, same as the
 *                   TF32 the reference's cuDNN path uses), fp32 accumulation in TMEM, fp32
 *                   residual stream. */
typedef enum svb_precision { SVB_PREC_FP32 = 0, SVB_PREC_TC = 1 } svb_precision;

/* One named host tensor of the reference state_dict (fp32 or fp16, contiguous). */
typedef struct svb_tensor {
    const char* name;      /* e.g. "dec.ups.0.weight_v" (keys of SynthesizerTrn.state_dict())   */
    const void* data;      /* host pointer                                                      */
    int32_t dtype;         /* 0 = float32, 1 = float16                                           */
    int32_t ndim;
    int64_t shape[4];
} svb_tensor;

/* Mirror of the hps.model fields the path depends on (models.py:344-372, 411-422, 441). */
typedef struct svb_model_cfg {
    int32_t inter_channels;          /* 192 */
    int32_t hidden_channels;         /* 192 */
    int32_t gin_channels;            /* 768 */
    int32_t n_flows;                 /* 4 coupling layers (models.py:22)                   */
    int32_t flow_wn_layers;          /* n_flow_layer -> WN n_layers (models.py:441)        */
    int32_t flow_kernel_size;        /* 5                                                   */
    int32_t upsample_initial_channel;/* 512 */
    int32_t n_upsamples;             /* <= 8 */
    int32_t upsample_rates[8];
    int32_t upsample_kernel_sizes[8];
    int32_t n_resblock_kernels;      /* 3 */
    int32_t resblock_kernel_sizes[4];
    int32_t resblock_dilations[4][3];
    int32_t sampling_rate;           /* 44100 */
    int32_t n_harmonics;             /* 9 (harmonic_num 8 + fundamental)                    */
    int32_t snake;                   /* 1: vdecoder/hifiganwithsnake (SnakeAlias activations), 0: LeakyReLU */
    int32_t num_mels;                /* > 0: the mel-conditioned vocoder vdecoder/nsf_hifigan (no flow, no speaker
                                        conditioning; conv_pre takes num_mels channels; keys without the "dec." prefix) */
    /* prior encoder (pre + enc_p); enc_layers = 0: the prefix stays with the caller (PyTorch)               */
    int32_t ssl_dim;                 /* 768: input channels of `pre` (models.py:400)                                       */
    int32_t enc_layers;              /* 6   (hps.model.n_layers)                                                          */
    int32_t enc_heads;               /* 2   (n_heads)                                                                     */
    int32_t enc_filter;              /* 768 (filter_channels)                                                             */
    int32_t enc_kernel;              /* 3   (kernel_size of the FFN convolutions)                                         */
    int32_t enc_window;              /* 4   (attentions.py:74 window_size)                                                */
} svb_model_cfg;

/* replaces: Svc.load_model's `.to(dev)` of the model (inference/infer_tool.py:189-200) */
SVB_API int svb_create(int device, svb_ctx** out);
SVB_API void svb_destroy(svb_ctx* ctx);

/* replaces: utils.load_checkpoint -> model.load_state_dict (utils.py:155-187) for flow.* and dec.*,
 * plus the per-forward weight_norm recomputation (vdecoder/hifigan/models.py:335-355, F6). */
SVB_API int svb_load_weights(svb_ctx* ctx, const svb_tensor* tensors, int n_tensors, const svb_model_cfg* cfg);

SVB_API int svb_set_precision(svb_ctx* ctx, int precision /* svb_precision */);
SVB_API int svb_get_precision(const svb_ctx* ctx);
/* Kernel-schedule switches (all numerically equivalent paths): "tma" (0/1: pair kernels load their operand tile from an
 * fp16 [B][T][C] copy through a TMA tensor map), "fuse_resblock" (0/1), "fuse_maxc" (largest C using the fused ResBlock),
 * "fuse_flow" (0/1: one kernel per coupling layer).
 * Noise source of the NSF excitation: "philox_noise" (0/1) - when 1, calls that pass noise == NULL draw the N(0,1) harmonic
 * noise in-kernel (Philox4x32-10 keyed by "philox_seed" and the sample index) instead of running noiseless; the waveform is
 * then statistically, not bitwise, equivalent to the reference's torch.randn stream (vdecoder/hifigan/models.py:266). */
SVB_API int svb_set_option(svb_ctx* ctx, const char* name, int value);

/* Device scratch needed by svb_infer_tail for a [B,*,T] call.  Pass ws = NULL to let the library
 * keep its own grow-only workspace. */
SVB_API size_t svb_workspace_bytes(const svb_ctx* ctx, int B, int T);

/* replaces: ResidualCouplingBlock.forward(reverse=True) (models.py:45-52; modules/modules.py:288-307,
 * 110-138).  z_p,z_out: [B,inter,T]; g: [B,gin,gT] with gT in {1,T}; lengths: int32[B] on DEVICE
 * or NULL (= all T, which is what SynthesizerTrn.infer always passes, models.py:503). */
SVB_API int svb_flow_reverse(svb_ctx* ctx, const float* z_p, const float* g, int gT, const int32_t* lengths,
                     float* z_out, int B, int T, void* ws, size_t ws_bytes, void* stream);

/* replaces: f0_upsamp + SourceModuleHnNSF/SineGen (vdecoder/hifigan/models.py:369-372,250-271,307-320).
 * f0: [B,T]; rand_ini: [B,n_harm] U[0,1) (column 0 ignored, :148); noise: [B,N,n_harm] N(0,1) or NULL
 * (NULL = noise-free excitation, for tests); har: [B,N], N = T*prod(upsample_rates). */
SVB_API int svb_nsf_source(svb_ctx* ctx, const float* f0, const float* rand_ini, const float* noise,
                   float* har, int B, int T, void* stream);

/* replaces: Generator.forward after m_source (vdecoder/hifigan/models.py:373-392).
 * z: [B,inter,T] (already masked); har: [B,N]; wav: [B,N]. */
SVB_API int svb_generator(svb_ctx* ctx, const float* z, const float* g, int gT, const float* har,
                  float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream);

/* replaces: vdecoder.nsf_hifigan.models.Generator.forward(mel, f0) (vdecoder/nsf_hifigan/models.py:259-278; callers
 * modules/enhancer.py:106, diffusion/vocoder.py:81-84).  mel: [B,num_mels,T]; f0: [B,T]; wav: [B,N].  Needs a context
 * loaded with svb_model_cfg.num_mels > 0. */
SVB_API int svb_vocoder(svb_ctx* ctx, const float* mel, const float* f0, const float* rand_ini, const float* noise,
                        float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream);

/* replaces: `z = self.flow(z_p, c_mask, g=g, reverse=True); o = self.dec(z * c_mask, g=g, f0=f0)`
 * (models.py:530-531): the three calls above back-to-back on one stream. */
SVB_API int svb_infer_tail(svb_ctx* ctx, const float* z_p, const float* g, int gT, const int32_t* lengths,
                   const float* f0, const float* rand_ini, const float* noise,
                   float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream);

/* replaces: `self.pre(c)` (models.py:400,518: Conv1d(ssl_dim -> hidden, k5, pad 2)).  c: [B,ssl_dim,T] -> x: [B,hidden,T].
 * Needs svb_model_cfg.enc_layers > 0 and the `pre.*` tensors at load time.  Tensor-core arithmetic (fp16 operands). */
SVB_API int svb_pre_conv(svb_ctx* ctx, const float* c, float* x, int B, int T, void* stream);

/* replaces: TextEncoder.forward (models.py:155-162) after the f0-embedding add, with an all-ones mask:
 *   x = enc_(x_in) [attentions.Encoder, modules/attentions.py:73-107: 6 x {rel-pos MHA, LayerNorm, k3 FFN, LayerNorm}];
 *   stats = proj(x);  m, logs = split(stats);  z_p = m + z_noise * exp(logs) * noice_scale.
 * x_in, z_noise, z_p (and the optional m_p, logs_p): [B,hidden|inter,T] fp32 device tensors.  One tcgen05 GEMM launch per
 * projection / FFN convolution, one fused attention kernel per layer (scores never leave the SM). */
SVB_API int svb_enc_p(svb_ctx* ctx, const float* x_in, const float* z_noise, float noice_scale,
                      float* z_p, float* m_p, float* logs_p, int B, int T, void* stream);

/* Same as svb_infer_tail with HOST buffers: copies inputs H2D, runs, copies wav D2H, synchronises. */
SVB_API int svb_infer_tail_host(svb_ctx* ctx, const float* z_p, const float* g, int gT,
                        const float* f0, const float* rand_ini, const float* noise,
                        float* wav, int B, int T);

/* Diagnostics. */
SVB_API const char* svb_strerror(int status);
SVB_API const char* svb_last_error(const svb_ctx* ctx);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
SVB_API int64_t svb_launch_count(const svb_ctx* ctx);
/* number of times a call made in SVB_PREC_TC had to run (part of) its work on the fp32 FFMA kernels (unsupported shapes /
 * conditioning).  The result is still correct but ~15x slower; the first occurrence also prints one line on stderr. */
SVB_API int64_t svb_fallback_count(const svb_ctx* ctx);
/* copies an internal activation ("z","conv_pre","ups0".."ups4","stage0".."stage4") of the LAST
 * svb_generator/svb_infer_tail call into dst (device, fp32, n floats); test hook. */
SVB_API int svb_debug_enable(svb_ctx* ctx, int on);
SVB_API int svb_debug_fetch(svb_ctx* ctx, const char* what, float* dst, size_t n, void* stream);
/* Test / microbenchmark hook: one ResBlock pair (stage, branch j, dilation index d) of the loaded generator on
 * caller device buffers x,out [B,C,L] (scratch [B,C,L] for the fp32 form).  variant >= 0 selects a tensor-core
 * tile variant, variant == -2 the two fp32 FFMA convolutions. */
SVB_API int svb_debug_pair(svb_ctx* ctx, int stage, int j, int d, const float* x, float* out, float* scratch,
                           int B, int L, int variant, float alpha, float beta, void* stream);
/* Same for one whole ResBlock branch j (three pairs) through the fused tensor-core kernel (stages with C <= 64). */
SVB_API int svb_debug_resblock(svb_ctx* ctx, int stage, int j, const float* x, float* out, int B, int L, int variant,
                               float alpha, float beta, void* stream);
/* CUDA-event timers on the launching stream, per kernel family ("pair_tc","pair_f32","flow","nsf_source",
 * "generator"): enable, run, then read the summed device time, launch count and algorithmic FLOPs/bytes
 * (bench.py's roofline).  Re-enabling clears the counters. */
SVB_API int svb_profile_enable(svb_ctx* ctx, int on);
SVB_API int svb_profile_read(svb_ctx* ctx, const char* name, double* total_ms, int64_t* count, double* flops, double* bytes);
/* ---- enc_p prefix helpers: fused element-wise tails of a transformer layer, time-major [B,L,C] fp32,
 * no context needed.  Reference: modules/attentions.py:95-106 (x = norm(x + y)), :317-363 (FFN with k-tap convs),
 * modules/modules.py:23-35 (LayerNorm over channels).
 *   svb_prefix_add_ln_im2col : y = LayerNorm_C(x + r); if cols != NULL also cols[b,l,t*C+c] = y[b, l+t-(k-1)/2, c] (0 outside)
 *   svb_prefix_ffn_tail      : y = LayerNorm_C(x + bias + sum_t ya[b, l+t-(k-1)/2, t*C+c])      (ya: [B,L,k*C])           */
SVB_API int svb_prefix_add_ln_im2col(const float* x, const float* r, const float* gamma, const float* beta, float eps, float* y, float* cols,
                                     int B, int L, int C, int k, void* stream);
SVB_API int svb_prefix_ffn_tail(const float* ya, const float* x, const float* bias, const float* gamma, const float* beta, float eps, float* y,
                                int B, int L, int C, int k, void* stream);

/*   svb_prefix_rel_softmax   : scores[rows,L] (rows = B*heads*L, row i of each [L,L] matrix) += relk on the 2w+1 band, softmax
 *                              over the last axis in place, pband[rows,2w+1] = band of p   (attentions.py:246-266)
 *   svb_prefix_attn_merge    : y[B,L,H*dk] = merge_heads(out[B,H,L,dk] + pband @ embv[2w+1,dk])  (attentions.py:262-267)      */
SVB_API int svb_prefix_rel_softmax(float* scores, const float* relk, float* pband, int rows, int L, int window, void* stream);
SVB_API int svb_prefix_attn_merge(const float* out, const float* pband, const float* embv, float* y, int B, int H, int L, int dk, int nb,
                                  void* stream);

SVB_API const char* svb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SOVITS_B200_H */
