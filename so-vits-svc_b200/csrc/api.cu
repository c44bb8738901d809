// C ABI of libsovits_b200.so: context, weight folding/packing, and the launch schedule of the
// flow -> NSF source -> generator tail.  See include/sovits_b200.h for the contract.
#include "../../include/sovits_b200.h"
#include "kernels.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

using namespace svb;

#define SVB_VERSION "sovits_b200 0.1.0 (sm_100a)"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct ConvW {            // folded fp32 weights of one Conv1d, packed [Cin][k][Cout]
    float* w = nullptr;
    float* b = nullptr;
    int Cin = 0, Cout = 0, k = 0;
    void* w_tc = nullptr;  // tensor-core image (fp16, swizzled), when built: holds w * tc_scale
    float tc_scale = 1.f;  // power-of-two range normalisation of the fp16 image (1 unless max|w| leaves [2^-6, 2^6])
    float* b_tc = nullptr; // bias * tc_scale (what a FIRST conv of a ResBlock pair adds before its LeakyReLU); == b when tc_scale == 1
};

struct ConvNW {           // tensor-core image of one layer for convn_tc_kernel
    void* img = nullptr;
    float* bias = nullptr;   // per column, column order
    int cinp = 0, cin_real = 0, N_total = 0, NC = 0, k = 1, pad_left = 0;
    int noise = 0, noise_stride = 0, noise_w0 = 0;   // fused noise conv (polyphase ups of the narrow stages)
    float acc_scale = 1.f;   // 1 / (power-of-two range normalisation applied to the fp16 image)
};

struct FlowLayer {
    ConvNW pre_tc, post_tc;
    std::vector<ConvNW> in_tc, rs_tc;
    float* cond_w_perm = nullptr;  // cond_layer rows permuted to the gate kernel's column order
    float* cond_b_perm = nullptr;
    ConvW pre, post;
    float* cond_w_nat = nullptr;   // [2H*L][gin] natural layout for the GEMV
    ConvW cond;                    // packed variant for time-varying g
    std::vector<ConvW> in_layers, res_skip;
    int in_c0 = 0, out_c0 = 0;     // which half feeds pre / receives post (Flip folded)
    // fused coupling-layer kernel (kernels_flow.cu): weight block stream, bias sums, conditioning in its chunk order
    void* fused_img = nullptr;
    float* fb_gate = nullptr; float* fb_h = nullptr; float* fb_m = nullptr;
    float* cond_w_perm2 = nullptr; float* cond_b_perm2 = nullptr;   // GEMV form (g[B,gin,1])
    ConvW cond2;                                                     // packed conv form (time-varying g)
};

struct SnakeP {           // SnakeAlias parameters of one activation: e^alpha and 1/(e^beta + 1e-9) per channel
    float* ealpha = nullptr;
    float* inv_beta = nullptr;
    int C = 0;
};

struct Stage {
    SnakeP snake_in;               // dec.snakes[i] (before ups[i])
    std::vector<SnakeP> acts;      // dec.resblocks[3i+j].activations[a] at [j*6 + a]
    float* up_w = nullptr;         // [s][Cin][2][Cout]
    float* up_b = nullptr;
    int Cin = 0, Cout = 0, s = 0, k = 0, p = 0;
    float* noise_w = nullptr;      // [Cout][K]
    float* noise_b = nullptr;
    int noise_K = 0, noise_s = 0, noise_p = 0;
    std::vector<ConvW> c1, c2;     // [branch*3 + d]
    std::vector<ConvNW> c1n, c2n;  // the same convolutions as convn images (Snake generator: SnakeAlias-loader conv kernel)
    ConvNW up_tc;
    ConvNW noise_tc;               // wide noise_convs (stage 0): a 2-tap GEMM over 64-sample excitation rows
};

}  // namespace

namespace {
struct EncLayer {
    ConvNW qkv, o, ffn1, ffn2a, ffn2b;
    float* ek = nullptr; float* ev = nullptr;
    float* g1 = nullptr; float* b1 = nullptr; float* g2 = nullptr; float* b2 = nullptr;
};
struct Prefix {
    bool ok = false;
    int ssl = 0, H = 0, F = 0, heads = 0, k = 0, window = 0, out2 = 0;
    ConvNW pre_a, pre_b, proj;
    std::vector<EncLayer> layers;
};
}  // namespace

struct svb_ctx {
    int device = 0;
    bool loaded = false;
    int precision = SVB_PREC_FP32;
    svb_model_cfg cfg{};
    std::vector<void*> allocs;
    std::vector<FlowLayer> flow;
    ConvW conv_pre;
    ConvNW conv_pre_tc;
    bool flow_tc_ok = false, gen_tc_ok = false;
    float* cond_all_w = nullptr; float* cond_all_b = nullptr;   // cond_layer of all coupling layers stacked (one GEMV per call)
    float* dcond_w_nat = nullptr;  // dec.cond [U][gin]
    float* dcond_b = nullptr;
    ConvW dcond;                   // packed (time-varying g)
    std::vector<Stage> stages;
    SnakeP snake_post;
    float* snake_filt = nullptr;   // the fixed 12-tap kaiser-sinc filter
    float* post_w = nullptr;       // [C][7]
    float post_b = 0.f;
    int post_C = 0, post_K = 7;
    float* lin_w = nullptr;
    float lin_b = 0.f;
    int hop = 1;
    Prefix prefix;                 // pre + enc_p on the library's own kernels (svb_pre_conv / svb_enc_p)
    DevBuf ws_prefix;
    DevBuf ws;                     // library-owned workspace
    DevBuf host_io;                // device staging for svb_infer_tail_host
    bool debug = false;
    std::map<std::string, DevBuf> dbg;
    int opt_tma = 0;            // TMA-fed pair kernels (fp16 operand copies written by the previous pair): 13.5 vs 13.3 ms/step
                                // against the 128-bit thread loader on B200, so off by default; env SVB_TC_TMA overrides
    int opt_fuse_rb = 1;        // fused ResBlock kernel for narrow stages
    int opt_fuse_maxc = 64;     // ... up to this channel count (C = 64 fused: 12.6 vs 13.15 ms/step against the pair chain)
    int opt_philox = 0;         // "throughput mode": calls with noise == NULL draw the harmonic noise in-kernel (Philox4x32-10)
    unsigned long long philox_seed = 52468;
    int opt_merge_branches = 1; // the three ResBlock branches of a wide, short stage in one pair launch per dilation index
    int opt_fuse_flow = 1;      // one kernel per coupling layer (kernels_flow.cu) instead of 10 conv-as-GEMM launches
    int64_t ffma_fallbacks = 0;    // times a "tc" call ran (part of) its work on the fp32 FFMA kernels
    bool warned_fallback = false;
    bool profile = false;
    struct ProfEntry { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; double flops = 0, bytes = 0; };
    std::map<std::string, ProfEntry> prof;
    std::string err;
};

namespace {

// Every entry point switches to the context's device for its own duration and restores the caller's current device on
// every exit path (the reference's Svc(device="cuda:1") never calls torch.cuda.set_device; svb_destroy may run at GC time).
struct DevGuard {
    int prev = -1; bool ok = true;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;                    // nothing to restore
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

int fail(svb_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

// A call in SVB_PREC_TC that cannot use the tensor-core kernels still computes the right answer on the FFMA kernels, ~15x
// slower: count it (svb_fallback_count) and say so once on stderr instead of degrading silently.
void note_fallback(svb_ctx* c, const char* why) {
    c->ffma_fallbacks++;
    if (!c->warned_fallback) {
        c->warned_fallback = true;
        std::fprintf(stderr, "[libsovits_b200] warning: precision=tc but running fp32 FFMA kernels (%s); further occurrences are only counted\n", why);
    }
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, SVB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

struct HostT {
    std::vector<float> v;
    std::vector<int64_t> shape;
};

using TMap = std::unordered_map<std::string, const svb_tensor*>;

int get_tensor(svb_ctx* ctx, const TMap& m, const std::string& name, std::vector<int64_t> want, HostT& out) {
    auto it = m.find(name);
    if (it == m.end()) return fail(ctx, SVB_ERR_MISSING_TENSOR, "missing tensor: " + name);
    const svb_tensor* t = it->second;
    int64_t n = 1;
    out.shape.assign(t->shape, t->shape + t->ndim);
    for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
    int64_t wn = 1;
    for (auto d : want) wn *= d;
    if (n != wn || (int)want.size() != t->ndim) {
        std::string s = "shape mismatch for " + name + ": got [";
        for (int i = 0; i < t->ndim; ++i) s += std::to_string(t->shape[i]) + (i + 1 < t->ndim ? "," : "");
        s += "] want [";
        for (size_t i = 0; i < want.size(); ++i) s += std::to_string(want[i]) + (i + 1 < want.size() ? "," : "");
        return fail(ctx, SVB_ERR_SHAPE, s + "]");
    }
    for (int i = 0; i < t->ndim; ++i)
        if (t->shape[i] != want[i]) return fail(ctx, SVB_ERR_SHAPE, "shape mismatch for " + name);
    out.v.resize(n);
    if (t->dtype == 0) {
        std::memcpy(out.v.data(), t->data, n * sizeof(float));
    } else if (t->dtype == 1) {
        const __half* h = static_cast<const __half*>(t->data);
        for (int64_t i = 0; i < n; ++i) out.v[i] = __half2float(h[i]);
    } else {
        return fail(ctx, SVB_ERR_INVALID_ARG, "unsupported dtype for " + name);
    }
    return SVB_OK;
}

// torch.nn.utils.weight_norm(dim=0): w = g * v / ||v||, norm over all dims but 0 — for ConvTranspose1d dim 0
// is Cin (vdecoder/hifigan/models.py:340-342).
int folded(svb_ctx* ctx, const TMap& m, const std::string& prefix, std::vector<int64_t> shape, HostT& w) {
    HostT g, v;
    int rc = get_tensor(ctx, m, prefix + ".weight_v", shape, v);
    if (rc) return rc;
    rc = get_tensor(ctx, m, prefix + ".weight_g", {shape[0], 1, 1}, g);
    if (rc) return rc;
    int64_t rows = shape[0], cols = (int64_t)v.v.size() / rows;
    w.v.resize(v.v.size());
    w.shape = shape;
    for (int64_t r = 0; r < rows; ++r) {
        double ss = 0;
        for (int64_t c = 0; c < cols; ++c) ss += (double)v.v[r * cols + c] * v.v[r * cols + c];
        // match torch: norm computed in fp32 then g*v/norm in fp32
        float nrm = (float)std::sqrt(ss);
        float gg = g.v[r];
        for (int64_t c = 0; c < cols; ++c) w.v[r * cols + c] = gg * v.v[r * cols + c] / nrm;
    }
    return SVB_OK;
}

int upload(svb_ctx* ctx, const void* src, size_t bytes, void** dst) {
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes ? bytes : 4));
    ctx->allocs.push_back(p);
    if (bytes) CU(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
    *dst = p;
    return SVB_OK;
}

// w: [Cout][Cin][k] (natural Conv1d layout) -> packed [Cin][k][Cout]; optional channel reversals.
int make_conv(svb_ctx* ctx, const std::vector<float>& w, const std::vector<float>& b, int Cout, int Cin, int k,
              bool rev_in, bool rev_out, ConvW& out) {
    std::vector<float> pk((size_t)Cin * k * Cout);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int t = 0; t < k; ++t) {
                int sco = rev_out ? Cout - 1 - co : co;
                int sci = rev_in ? Cin - 1 - ci : ci;
                pk[((size_t)ci * k + t) * Cout + co] = w[((size_t)sco * Cin + sci) * k + t];
            }
    std::vector<float> bb(Cout);
    for (int co = 0; co < Cout; ++co) bb[co] = b[rev_out ? Cout - 1 - co : co];
    out.Cin = Cin; out.Cout = Cout; out.k = k;
    int rc = upload(ctx, pk.data(), pk.size() * sizeof(float), (void**)&out.w);
    if (rc) return rc;
    rc = upload(ctx, bb.data(), bb.size() * sizeof(float), (void**)&out.b);
    if (rc) return rc;
    if (Cin == Cout && (k == 3 || k == 7 || k == 11) && (Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256)) {
        size_t ib = tc_weight_image_bytes(Cin, k);
        std::vector<uint8_t> img(ib);
        // fp16 range: see make_convn.  LeakyReLU is positively homogeneous, so a pair computes
        //   x + conv2'(lrelu(conv1'(lrelu x) + s1 b1)) / (s1 s2) + b2   with conv' = the normalised images.
        float wmax = 0.f;
        for (float v : w) wmax = std::max(wmax, std::fabs(v));
        out.tc_scale = 1.f;
        if (wmax > 0.f && std::isfinite(wmax) && (wmax < 0.015625f || wmax > 64.f)) out.tc_scale = std::exp2(-std::round(std::log2(wmax)));
        tc_pack_weight_image(w.data(), Cin, k, img.data(), out.tc_scale);
        rc = upload(ctx, img.data(), ib, &out.w_tc);
        if (rc) return rc;
        out.b_tc = out.b;
        if (out.tc_scale != 1.f) {
            std::vector<float> bs(Cout);
            for (int co = 0; co < Cout; ++co) bs[co] = bb[co] * out.tc_scale;
            rc = upload(ctx, bs.data(), bs.size() * sizeof(float), (void**)&out.b_tc);
            if (rc) return rc;
        }
    }
    return SVB_OK;
}

int make_convn(svb_ctx* ctx, int cinp, int cin_real, int N_total, int NC, int k, int pad_left,
               const std::function<float(int, int, int)>& wcol, const std::function<float(int)>& bcol, ConvNW& out,
               const std::function<float(int, int)>* ncol = nullptr, int noise_kind = 1, float force_scale = 0.f) {
    out.cinp = cinp; out.cin_real = cin_real; out.N_total = N_total; out.NC = NC; out.k = k; out.pad_left = pad_left;
    out.noise = ncol ? noise_kind : 0;
    const size_t ib = convn_weight_image_bytes(cinp, N_total, NC, k, out.noise);
    std::vector<uint8_t> img(ib);
    // fp16 has a 5-bit exponent (TF32, the reference's CUDA arithmetic, has 8): weights far from 1 would lose bits to
    // subnormals (< 6e-5) or overflow (> 65504).  Normalise the layer by a power of two (exact) when its largest weight
    // leaves [2^-6, 2^6]; the epilogue multiplies the fp32 accumulator by the inverse.
    float wmax = 0.f;
    for (int col = 0; col < N_total; ++col)
        for (int ci = 0; ci < cin_real; ++ci)
            for (int tap = 0; tap < k; ++tap) wmax = std::max(wmax, std::fabs(wcol(col, ci, tap)));
    if (ncol)
        for (int col = 0; col < N_total; ++col)
            for (int u = 0; u < (noise_kind == 2 ? 80 : 16); ++u) wmax = std::max(wmax, std::fabs((*ncol)(col, u)));
    float wscale = 1.f;
    if (wmax > 0.f && std::isfinite(wmax) && (wmax < 0.015625f || wmax > 64.f)) wscale = std::exp2(-std::round(std::log2(wmax)));
    if (force_scale > 0.f) wscale = force_scale;       // K-chunked layers: both images must share one normalisation
    out.acc_scale = 1.f / wscale;
    std::function<float(int, int)> ncol_s;
    if (ncol) ncol_s = [&](int col, int u) { return (*ncol)(col, u) * wscale; };
    convn_pack_weight_image(cinp, N_total, NC, k, [&](int col, int ci, int tap) { return ci < cin_real ? wcol(col, ci, tap) * wscale : 0.f; },
                            ncol ? &ncol_s : nullptr, out.noise, img.data());
    int rc = upload(ctx, img.data(), ib, &out.img);
    if (rc) return rc;
    std::vector<float> bc(N_total);
    for (int c = 0; c < N_total; ++c) bc[c] = bcol(c);
    return upload(ctx, bc.data(), bc.size() * sizeof(float), (void**)&out.bias);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Column chunks per CTA of a conv-as-GEMM launch: as many as possible (the operand tile is staged once and the epilogue of a
// chunk overlaps the MMAs of the next), but never so many that the grid falls under two CTAs per SM - the stage-0 upsampler
// (862 frames -> 56 row tiles) ran on 56 of 148 SMs with all 8 chunks in one CTA.
int pick_chunks_per_cta(int n_rows, int rows_per_cta, int B, int n_chunks) {
    const long long tiles = (long long)((n_rows + rows_per_cta - 1) / rows_per_cta) * B;
    int cpc = n_chunks;
    while (cpc > 1 && tiles * ((n_chunks + cpc - 1) / cpc) < 2 * 148) cpc = (cpc + 1) / 2;
    return cpc;
}

struct WsPlan {
    size_t total = 0;
    size_t off_y, off_h, off_xin, off_acts, off_out, off_gcond, off_gcond_all, off_dgcond, off_phase;
    size_t off_har, off_pre, off_X, off_A, off_Bb, off_T, off_O, off_z, off_S, off_A16, off_B16;
    size_t off_br[6]; size_t br_elems = 0;      // per-branch ping-pong buffers of the branch-merged pair launches
};

// Branch-merged pair launches (launch_pair_tc_multi) pay a memset of the stage output and private buffers; they win where a
// single pair launch leaves the GPU half empty: wide stages (pair kernels, C >= 128) with at most ~3 waves of tiles.
bool merge_branches(int C, int L, int B) {
    if (C < 128) return false;
    // measured (config 2): merging stage 1 as well (12 waves of tiles), with the branches interleaved in launch order, takes the
    // pair kernels from 3.44 to 3.22 ms per step - a k = 3 pair is memory-phase bound, a k = 11 pair MMA bound, side by side on
    // an SM they fill each other's gaps
    static const int max_waves = [] { const char* e = std::getenv("SVB_MERGE_WAVES"); return e ? std::atoi(e) : 64; }();
    const long long tiles = (long long)((L + 245) / 246) * B;
    return tiles <= (long long)max_waves * 148;
}

WsPlan plan_ws(const svb_model_cfg& c, int B, int T, int gT) {
    WsPlan p;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t f = sizeof(float);
    const int H = c.hidden_channels, C = c.inter_channels, L = c.flow_wn_layers;
    size_t BT = (size_t)B * T;
    p.off_z = take(BT * C * f);
    p.off_h = take(BT * H * f);
    p.off_xin = take(BT * 2 * H * f);
    p.off_acts = take(BT * H * f);
    p.off_out = take(BT * H * f);
    p.off_gcond = take((size_t)B * 2 * H * L * (gT > 1 ? T : 1) * f);
    p.off_gcond_all = take((size_t)B * 2 * H * L * (c.n_flows > 0 ? c.n_flows : 1) * f);
    p.off_dgcond = take((size_t)B * c.upsample_initial_channel * (gT > 1 ? T : 1) * f);
    p.off_phase = take(BT * c.n_harmonics * sizeof(double));
    long long hop = 1;
    for (int i = 0; i < c.n_upsamples; ++i) hop *= c.upsample_rates[i];
    p.off_har = take(BT * hop * f);
    p.off_pre = take(BT * c.upsample_initial_channel * f);
    size_t maxel = 0;
    long long len = T;
    for (int i = 0; i < c.n_upsamples; ++i) {
        len *= c.upsample_rates[i];
        size_t el = (size_t)(c.upsample_initial_channel >> (i + 1)) * (size_t)len;
        if (el > maxel) maxel = el;
    }
    maxel *= B;
    p.off_X = take(maxel * f);
    p.off_A = take(maxel * f);
    p.off_Bb = take(maxel * f);
    p.off_T = take(maxel * f);
    p.off_O = take(maxel * f);
    p.off_S = c.snake ? take(maxel * f) : 0;
    // stages whose pair launches have few tiles run their three branches in one launch (run_generator): 3 x 2 private buffers
    {
        long long l2 = T;
        for (int i = 0; i < c.n_upsamples; ++i) {
            l2 *= c.upsample_rates[i];
            const int co = c.upsample_initial_channel >> (i + 1);
            if (merge_branches(co, (int)l2, B)) p.br_elems = std::max(p.br_elems, (size_t)B * co * (size_t)l2);
        }
        for (int q = 0; q < 6; ++q) p.off_br[q] = p.br_elems ? take(p.br_elems * f) : 0;
    }
    p.off_A16 = take(maxel * 2);      // fp16 [B][T][C] copies of lrelu(residual stream) for the TMA-fed pair kernels
    p.off_B16 = take(maxel * 2);
    p.total = o;
    p.off_y = p.off_z;
    return p;
}

int ensure_ws(svb_ctx* ctx, size_t need, void* user_ws, size_t user_bytes, char** base) {
    if (user_ws) {
        if (user_bytes < need) return fail(ctx, SVB_ERR_WORKSPACE, "caller workspace too small");
        *base = static_cast<char*>(user_ws);
        return SVB_OK;
    }
    if (ctx->ws.bytes < need) {
        if (ctx->ws.p) CU(cudaFree(ctx->ws.p));
        ctx->ws.p = nullptr; ctx->ws.bytes = 0;
        CU(cudaMalloc(&ctx->ws.p, need));
        ctx->ws.bytes = need;
    }
    *base = static_cast<char*>(ctx->ws.p);
    return SVB_OK;
}

int dbg_keep(svb_ctx* ctx, const std::string& name, const float* src, size_t n, cudaStream_t st) {
    if (!ctx->debug) return SVB_OK;
    DevBuf& d = ctx->dbg[name];
    if (d.bytes < n * sizeof(float)) {
        if (d.p) CU(cudaFree(d.p));
        CU(cudaMalloc(&d.p, n * sizeof(float)));
        d.bytes = n * sizeof(float);
    }
    CU(cudaMemcpyAsync(d.p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SVB_OK;
}

// CUDA-event timers on the launching stream (svb_profile_enable); read back by svb_profile_read.
struct ProfScope {
    svb_ctx* c; cudaStream_t st; cudaEvent_t e1 = nullptr; svb_ctx::ProfEntry* ent = nullptr;
    ProfScope(svb_ctx* c_, const char* name, cudaStream_t st_, double flops, double bytes) : c(c_), st(st_) {
        if (!c->profile) return;
        ent = &c->prof[name];
        cudaEvent_t e0;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        ent->ev.push_back({e0, e1});
        ent->flops += flops; ent->bytes += bytes;
    }
    ~ProfScope() { if (ent) cudaEventRecord(e1, st); }
};

int check_launch(svb_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SVB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    if (sticky_launch_error()) { sticky_launch_error() = 0; return fail(ctx, SVB_ERR_CUDA, std::string(what) + ": kernel set-up failed (dynamic shared memory opt-in)"); }
    return SVB_OK;
}

// ---------------------------------------------------------------------------------------------- prior encoder (pre + enc_p)
// Packs `pre` and the six transformer layers of enc_p as conv-as-GEMM images (channel-major activations, like the reference):
//   pre   Conv1d(768 -> 192, k5)            two K halves of 384 channels (the operand tile of one CTA holds <= 512 channels)
//   qkv   conv_q | conv_k | conv_v stacked  192 -> 576, the 1/sqrt(dk) of attentions.py:243 folded into the q rows
//   o     conv_o 192 -> 192;  ffn conv_1 192 -> 768 k3 (+ReLU);  conv_2 768 -> 192 k3 in two K halves;  proj 192 -> 384
float shared_pow2_scale(const std::vector<float>& w) {
    float wmax = 0.f;
    for (float v : w) wmax = std::max(wmax, std::fabs(v));
    if (wmax > 0.f && std::isfinite(wmax) && (wmax < 0.015625f || wmax > 64.f)) return std::exp2(-std::round(std::log2(wmax)));
    return 1.f;
}

int load_prefix(svb_ctx* ctx, const TMap& m) {
    const svb_model_cfg& c = ctx->cfg;
    Prefix& P = ctx->prefix;
    const int H = c.hidden_channels, F = c.enc_filter, heads = c.enc_heads, k = c.enc_kernel, S = c.ssl_dim, nl = c.enc_layers;
    const int C2 = 2 * c.inter_channels;
    if (H != 192 || heads < 1 || H % heads || H / heads != 96 || S != 768 || F != 768 || (k != 1 && k != 3 && k != 5) ||
        c.enc_window < 0 || c.enc_window > 4 || (C2 % 192))
        return SVB_OK;                       // shapes the prefix kernels do not serve: the caller keeps the PyTorch prefix
    int rc;
    HostT w, b;
    if ((rc = get_tensor(ctx, m, "pre.weight", {H, S, 5}, w))) return rc;
    if ((rc = get_tensor(ctx, m, "pre.bias", {H}, b))) return rc;
    {
        const std::vector<float> wv = w.v, bv = b.v;
        const float sc = shared_pow2_scale(wv);
        for (int half = 0; half < 2; ++half) {
            if ((rc = make_convn(ctx, 384, 384, H, H, 5, 2,
                                 [&](int col, int ci, int tap) { return wv[((size_t)col * S + half * 384 + ci) * 5 + tap]; },
                                 [&](int col) { return half == 0 ? bv[col] : 0.f; }, half == 0 ? P.pre_a : P.pre_b, nullptr, 1, sc))) return rc;
        }
    }
    P.layers.assign(nl, EncLayer());
    const float qs = 1.0f / std::sqrt((float)(H / heads));
    for (int l = 0; l < nl; ++l) {
        EncLayer& E = P.layers[l];
        const std::string a = "enc_p.enc_.attn_layers." + std::to_string(l) + ".";
        HostT wq, wk, wvv, bq, bk, bvv;
        if ((rc = get_tensor(ctx, m, a + "conv_q.weight", {H, H, 1}, wq)) || (rc = get_tensor(ctx, m, a + "conv_q.bias", {H}, bq)) ||
            (rc = get_tensor(ctx, m, a + "conv_k.weight", {H, H, 1}, wk)) || (rc = get_tensor(ctx, m, a + "conv_k.bias", {H}, bk)) ||
            (rc = get_tensor(ctx, m, a + "conv_v.weight", {H, H, 1}, wvv)) || (rc = get_tensor(ctx, m, a + "conv_v.bias", {H}, bvv))) return rc;
        if ((rc = make_convn(ctx, H, H, 3 * H, H, 1, 0,
                             [&](int col, int ci, int) { return col < H ? wq.v[(size_t)col * H + ci] * qs : (col < 2 * H ? wk.v[(size_t)(col - H) * H + ci] : wvv.v[(size_t)(col - 2 * H) * H + ci]); },
                             [&](int col) { return col < H ? bq.v[col] * qs : (col < 2 * H ? bk.v[col - H] : bvv.v[col - 2 * H]); }, E.qkv))) return rc;
        if ((rc = get_tensor(ctx, m, a + "conv_o.weight", {H, H, 1}, w)) || (rc = get_tensor(ctx, m, a + "conv_o.bias", {H}, b))) return rc;
        {
            const std::vector<float> wv = w.v, bv = b.v;
            if ((rc = make_convn(ctx, H, H, H, H, 1, 0, [&](int col, int ci, int) { return wv[(size_t)col * H + ci]; }, [&](int col) { return bv[col]; }, E.o))) return rc;
        }
        const int nbnd = 2 * c.enc_window + 1, dk = H / heads;
        if ((rc = get_tensor(ctx, m, a + "emb_rel_k", {1, nbnd, dk}, w))) return rc;
        if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&E.ek))) return rc;
        if ((rc = get_tensor(ctx, m, a + "emb_rel_v", {1, nbnd, dk}, w))) return rc;
        if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&E.ev))) return rc;
        for (int n = 0; n < 2; ++n) {
            const std::string np_ = "enc_p.enc_.norm_layers_" + std::to_string(n + 1) + "." + std::to_string(l) + ".";
            if ((rc = get_tensor(ctx, m, np_ + "gamma", {H}, w)) || (rc = get_tensor(ctx, m, np_ + "beta", {H}, b))) return rc;
            if ((rc = upload(ctx, w.v.data(), H * sizeof(float), (void**)(n ? &E.g2 : &E.g1)))) return rc;
            if ((rc = upload(ctx, b.v.data(), H * sizeof(float), (void**)(n ? &E.b2 : &E.b1)))) return rc;
        }
        const std::string f = "enc_p.enc_.ffn_layers." + std::to_string(l) + ".";
        if ((rc = get_tensor(ctx, m, f + "conv_1.weight", {F, H, k}, w)) || (rc = get_tensor(ctx, m, f + "conv_1.bias", {F}, b))) return rc;
        {
            const std::vector<float> wv = w.v, bv = b.v;
            if ((rc = make_convn(ctx, H, H, F, H, k, (k - 1) / 2, [&](int col, int ci, int tap) { return wv[((size_t)col * H + ci) * k + tap]; },
                                 [&](int col) { return bv[col]; }, E.ffn1))) return rc;
        }
        if ((rc = get_tensor(ctx, m, f + "conv_2.weight", {H, F, k}, w)) || (rc = get_tensor(ctx, m, f + "conv_2.bias", {H}, b))) return rc;
        {
            const std::vector<float> wv = w.v, bv = b.v;
            const float sc = shared_pow2_scale(wv);
            for (int half = 0; half < 2; ++half)
                if ((rc = make_convn(ctx, 384, 384, H, H, k, (k - 1) / 2,
                                     [&](int col, int ci, int tap) { return wv[((size_t)col * F + half * 384 + ci) * k + tap]; },
                                     [&](int col) { return half == 0 ? bv[col] : 0.f; }, half == 0 ? E.ffn2a : E.ffn2b, nullptr, 1, sc))) return rc;
        }
    }
    if ((rc = get_tensor(ctx, m, "enc_p.proj.weight", {C2, H, 1}, w)) || (rc = get_tensor(ctx, m, "enc_p.proj.bias", {C2}, b))) return rc;
    {
        const std::vector<float> wv = w.v, bv = b.v;
        if ((rc = make_convn(ctx, H, H, C2, H, 1, 0, [&](int col, int ci, int) { return wv[(size_t)col * H + ci]; }, [&](int col) { return bv[col]; }, P.proj))) return rc;
    }
    P.ssl = S; P.H = H; P.F = F; P.heads = heads; P.k = k; P.window = c.enc_window; P.out2 = C2;
    P.ok = true;
    return SVB_OK;
}

// one plain conv-as-GEMM launch on channel-major tensors: y[B,N,T] = alpha-less (conv_k(x[:, x_c0 : x_c0+cin]) + bias [+ res]) [+ y]
int prefix_conv(svb_ctx* ctx, const ConvNW& W, const float* x, int x_ctot, int x_c0, float* y, const float* res, float beta, int relu,
                int B, int T, cudaStream_t st, const ConvNW* W2 = nullptr, int k2_c0 = 0) {
    ConvNTC a;
    a.x = x; a.x_ctot = x_ctot; a.x_c0 = x_c0; a.cin_real = W.cin_real; a.cinp = W.cinp; a.Tin = T;
    a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = W.k; a.dil = 1; a.pad_left = W.pad_left;
    a.n_rows = T; a.N_total = W.N_total; a.NC = W.NC; a.Ty = T; a.B = B; a.out_relu = relu;
    if (W2) { a.w_k2 = W2->img; a.k2_c0 = k2_c0; }       // second K chunk of the same layer (shared normalisation, zero bias)
    const int n_chunks = (W.N_total + W.NC - 1) / W.NC;
    a.chunks_per_cta = n_chunks >= 4 ? 2 : 1;
    a.seg[0].y = y; a.seg[0].y_ctot = W.N_total; a.seg[0].col0 = 0; a.seg[0].col1 = W.N_total;
    a.seg[0].res = res; a.seg[0].res_ctot = W.N_total; a.seg[0].beta = beta;
    ProfScope ps(ctx, "enc_gemm", st, 2.0 * (double)W.cin_real * (W2 ? 2 : 1) * W.k * W.N_total * (double)T * B,
                 ((double)W.cin_real * (W2 ? 2 : 1) + W.N_total) * (double)T * B * sizeof(float));
    const int rc = launch_convn_tc(a, st);
    return rc ? fail(ctx, rc, "prefix conv launch failed") : SVB_OK;
}

// ---------------------------------------------------------------------------------------------- flow
int run_flow(svb_ctx* ctx, const float* z_p, const float* g, int gT, const int32_t* lengths, float* y,
             int B, int T, char* ws, const WsPlan& pl, cudaStream_t st) {
    const svb_model_cfg& c = ctx->cfg;
    const int H = c.hidden_channels, C = c.inter_channels, L = c.flow_wn_layers, half = C / 2;
    float* h = reinterpret_cast<float*>(ws + pl.off_h);
    float* xin = reinterpret_cast<float*>(ws + pl.off_xin);
    float* acts = reinterpret_cast<float*>(ws + pl.off_acts);
    float* out = reinterpret_cast<float*>(ws + pl.off_out);
    float* gcond = reinterpret_cast<float*>(ws + pl.off_gcond);
    if (y != z_p) CU(cudaMemcpyAsync(y, z_p, (size_t)B * C * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const bool fused = ctx->precision == SVB_PREC_TC && ctx->flow_tc_ok && ctx->opt_fuse_flow && !ctx->flow.empty() && ctx->flow[0].fused_img;
    const bool cond_once = fused && gT == 1 && ctx->cond_all_w;
    float* gcond_all = reinterpret_cast<float*>(ws + pl.off_gcond_all);
    if (cond_once)      // out[b][fl*2HL + col]
        launch_gemv(ctx->cond_all_w, ctx->cond_all_b, g, gcond_all, B, c.n_flows * 2 * H * L, c.gin_channels, st);
    for (int fl = c.n_flows - 1; fused && fl >= 0; --fl) {
        // one kernel per coupling layer; the conditioning cond_layer(g) is a GEMV (g[B,gin,1]; all layers in one launch) or, for
        // speaker-mix g[B,gin,T], a 1x1 convolution written in the gate's chunk order and added per (frame, column) by the gate
        FlowLayer& F = ctx->flow[fl];
        if (cond_once) {
        } else if (gT == 1) {
            launch_gemv(F.cond_w_perm2, F.cond_b_perm2, g, gcond, B, 2 * H * L, c.gin_channels, st);
        } else {
            ConvF32 cg;
            cg.x = g; cg.x_ctot = c.gin_channels; cg.Cin = c.gin_channels; cg.Tin = T;
            cg.w = F.cond2.w; cg.bias = F.cond2.b; cg.Cout = 2 * H * L; cg.k = 1;
            cg.y = gcond; cg.y_ctot = 2 * H * L; cg.Ty = T; cg.n_out = T; cg.B = B;
            launch_conv_f32(cg, st);
        }
        FlowLayerTC a;
        a.y = y; a.y_ctot = C; a.in_c0 = F.in_c0; a.out_c0 = F.out_c0;
        a.w = F.fused_img; a.bias_gate = F.fb_gate; a.bias_h = F.fb_h; a.bias_m = F.fb_m;
        a.gcond = gT == 1 ? (cond_once ? gcond_all + (size_t)fl * 2 * H * L : gcond) : nullptr; a.gcond_t = gT == 1 ? nullptr : gcond;
        a.gcond_bstride = cond_once ? c.n_flows * 2 * H * L : 2 * H * L;
        a.lengths = lengths; a.B = B; a.T = T; a.H = H; a.half = half; a.L = L; a.k = c.flow_kernel_size;
        const int trc = launch_flow_layer_tc(a, st);
        if (trc) return fail(ctx, trc, "fused coupling-layer kernel launch failed");
    }
    if (fused) return check_launch(ctx, "flow");
    const bool use_tc = ctx->precision == SVB_PREC_TC && ctx->flow_tc_ok && gT == 1;
    if (ctx->precision == SVB_PREC_TC && !use_tc) note_fallback(ctx, gT != 1 ? "flow: time-varying conditioning g" : "flow: unsupported layer shapes");
    for (int fl = c.n_flows - 1; use_tc && fl >= 0; --fl) {
        FlowLayer& F = ctx->flow[fl];
        int trc;
        auto base_args = [&](const ConvNW& W, const float* x, int x_ctot, int x_c0) {
            ConvNTC a;
            a.x = x; a.x_ctot = x_ctot; a.x_c0 = x_c0; a.cin_real = W.cin_real; a.cinp = W.cinp; a.Tin = T;
            a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = W.k; a.dil = 1; a.pad_left = W.pad_left;
            a.n_rows = T; a.N_total = W.N_total; a.NC = W.NC; a.chunks_per_cta = 1; a.Ty = T; a.lengths = lengths; a.B = B;
            return a;
        };
        {   // h = pre(x0) * mask
            ConvNTC a = base_args(F.pre_tc, y, C, F.in_c0);
            a.seg[0].y = h; a.seg[0].y_ctot = H; a.seg[0].col0 = 0; a.seg[0].col1 = H; a.seg[0].masked = 1;
            if ((trc = launch_convn_tc(a, st))) return fail(ctx, trc, "convn launch failed (flow pre)");
        }
        launch_gemv(F.cond_w_perm, F.cond_b_perm, g, gcond, B, 2 * H * L, c.gin_channels, st);
        for (int i = 0; i < L; ++i) {
            {   // acts = tanh(.)*sigmoid(.) of in_layers[i](h) + cond
                ConvNTC a = base_args(F.in_tc[i], h, H, 0);
                a.mode = 2; a.bias_b = gcond; a.bias_b_stride = 2 * H * L; a.bias_b_off = 2 * H * i;
                a.seg[0].y = acts; a.seg[0].y_ctot = H;
                if ((trc = launch_convn_tc(a, st))) return fail(ctx, trc, "convn launch failed (flow in_layer)");
            }
            {   // res/skip
                ConvNTC a = base_args(F.rs_tc[i], acts, H, 0);
                if (i < L - 1) {
                    a.n_seg = 2;
                    a.seg[0].y = h; a.seg[0].y_ctot = H; a.seg[0].col0 = 0; a.seg[0].col1 = H; a.seg[0].beta = 1.f; a.seg[0].masked = 1;
                    a.seg[1].y = out; a.seg[1].y_ctot = H; a.seg[1].col0 = H; a.seg[1].col1 = 2 * H; a.seg[1].beta = (i > 0) ? 1.f : 0.f; a.seg[1].masked = 1;
                } else {
                    a.seg[0].y = out; a.seg[0].y_ctot = H; a.seg[0].col0 = 0; a.seg[0].col1 = H; a.seg[0].beta = (i > 0) ? 1.f : 0.f; a.seg[0].masked = 1;
                }
                if ((trc = launch_convn_tc(a, st))) return fail(ctx, trc, "convn launch failed (flow res_skip)");
            }
        }
        {   // x1 = (x1 - post(out)) * mask
            ConvNTC a = base_args(F.post_tc, out, H, 0);
            a.seg[0].y = y; a.seg[0].y_ctot = C; a.seg[0].y_c0 = F.out_c0; a.seg[0].col0 = 0; a.seg[0].col1 = half;
            a.seg[0].alpha = -1.f; a.seg[0].beta = 1.f; a.seg[0].masked = 1;
            if ((trc = launch_convn_tc(a, st))) return fail(ctx, trc, "convn launch failed (flow post)");
        }
    }
    for (int fl = c.n_flows - 1; !use_tc && fl >= 0; --fl) {
        FlowLayer& F = ctx->flow[fl];
        // h = pre(x0) * mask
        ConvF32 a;
        a.x = y; a.x_ctot = C; a.x_c0 = F.in_c0; a.Cin = half; a.Tin = T;
        a.w = F.pre.w; a.bias = F.pre.b; a.Cout = H; a.k = 1;
        a.y = h; a.y_ctot = H; a.Ty = T; a.n_out = T; a.lengths = lengths; a.B = B;
        launch_conv_f32(a, st);
        // conditioning: g -> [B, 2H*L, gT]
        if (gT == 1) {
            launch_gemv(F.cond_w_nat, F.cond.b, g, gcond, B, 2 * H * L, c.gin_channels, st);
        } else {
            ConvF32 cg;
            cg.x = g; cg.x_ctot = c.gin_channels; cg.Cin = c.gin_channels; cg.Tin = T;
            cg.w = F.cond.w; cg.bias = F.cond.b; cg.Cout = 2 * H * L; cg.k = 1;
            cg.y = gcond; cg.y_ctot = 2 * H * L; cg.Ty = T; cg.n_out = T; cg.B = B;
            launch_conv_f32(cg, st);
        }
        for (int i = 0; i < L; ++i) {
            ConvF32 ci;
            ci.x = h; ci.x_ctot = H; ci.Cin = H; ci.Tin = T;
            ci.w = F.in_layers[i].w; ci.bias = F.in_layers[i].b; ci.Cout = 2 * H; ci.k = c.flow_kernel_size;
            ci.pad_left = (c.flow_kernel_size - 1) / 2;
            if (gT == 1) { ci.bias_b = gcond; ci.bias_b_stride = 2 * H * L; ci.bias_b_off = 2 * H * i; }
            else { ci.bias_t = gcond; ci.bias_t_ctot = 2 * H * L; ci.bias_t_c0 = 2 * H * i; }
            ci.y = xin; ci.y_ctot = 2 * H; ci.Ty = T; ci.n_out = T; ci.B = B;
            launch_conv_f32(ci, st);
            launch_gate(xin, acts, B, H, T, st);
            const ConvW& R = F.res_skip[i];
            if (i < L - 1) {
                // residual half: h = (h + rs[:H]) * mask
                ConvF32 r1;
                r1.x = acts; r1.x_ctot = H; r1.Cin = H; r1.Tin = T;
                r1.w = R.w; r1.bias = R.b; r1.Cout = H; r1.k = 1;
                // packed layout is [Cin][1][2H]: restrict to the first H output columns via a strided view
                // (handled by packing res/skip halves separately, see load)
                r1.y = h; r1.y_ctot = H; r1.Ty = T; r1.n_out = T; r1.beta = 1.f; r1.lengths = lengths; r1.B = B;
                launch_conv_f32(r1, st);
                // skip half: out (+)= rs[H:]
                ConvF32 r2 = r1;
                r2.w = R.w + (size_t)H * H; r2.bias = R.b + H;
                r2.y = out; r2.beta = (i > 0) ? 1.f : 0.f;
                launch_conv_f32(r2, st);
            } else {
                ConvF32 r2;
                r2.x = acts; r2.x_ctot = H; r2.Cin = H; r2.Tin = T;
                r2.w = R.w; r2.bias = R.b; r2.Cout = H; r2.k = 1;
                r2.y = out; r2.y_ctot = H; r2.Ty = T; r2.n_out = T; r2.beta = (i > 0) ? 1.f : 0.f;
                r2.lengths = lengths; r2.B = B;
                launch_conv_f32(r2, st);
            }
        }
        // x1 = (x1 - post(out)*mask) * mask   (mean_only => logs = 0)
        ConvF32 p;
        p.x = out; p.x_ctot = H; p.Cin = H; p.Tin = T;
        p.w = F.post.w; p.bias = F.post.b; p.Cout = half; p.k = 1;
        p.y = y; p.y_ctot = C; p.y_c0 = F.out_c0; p.Ty = T; p.n_out = T;
        p.alpha = -1.f; p.beta = 1.f; p.lengths = lengths; p.B = B;
        launch_conv_f32(p, st);
    }
    return check_launch(ctx, "flow");
}

// ---------------------------------------------------------------------------------------------- generator
int run_generator(svb_ctx* ctx, const float* z, const float* g, int gT, const float* har, float* wav,
                  int B, int T, char* ws, const WsPlan& pl, cudaStream_t st) {
    const svb_model_cfg& c = ctx->cfg;
    const int U = c.upsample_initial_channel;
    float* dg = reinterpret_cast<float*>(ws + pl.off_dgcond);
    float* pre = reinterpret_cast<float*>(ws + pl.off_pre);
    float* X = reinterpret_cast<float*>(ws + pl.off_X);
    float* A = reinterpret_cast<float*>(ws + pl.off_A);
    float* Bb = reinterpret_cast<float*>(ws + pl.off_Bb);
    float* Tm = reinterpret_cast<float*>(ws + pl.off_T);
    float* O = reinterpret_cast<float*>(ws + pl.off_O);
    const long long N = (long long)T * ctx->hop;
    int rc;

    // conv_pre(z) + cond(g)
    const bool melv = c.num_mels > 0;
    const int Cpre = melv ? c.num_mels : c.inter_channels;
    ConvF32 cp;
    cp.x = z; cp.x_ctot = Cpre; cp.Cin = Cpre; cp.Tin = T;
    cp.w = ctx->conv_pre.w; cp.bias = ctx->conv_pre.b; cp.Cout = U; cp.k = ctx->conv_pre.k; cp.pad_left = (cp.k - 1) / 2;
    cp.y = pre; cp.y_ctot = U; cp.Ty = T; cp.n_out = T; cp.B = B;
    if (melv) {
        // vdecoder/nsf_hifigan: no speaker conditioning
    } else if (gT == 1) {
        launch_gemv(ctx->dcond_w_nat, ctx->dcond_b, g, dg, B, U, c.gin_channels, st);
        cp.bias_b = dg; cp.bias_b_stride = U; cp.bias_b_off = 0;
    } else {
        ConvF32 cg;
        cg.x = g; cg.x_ctot = c.gin_channels; cg.Cin = c.gin_channels; cg.Tin = T;
        cg.w = ctx->dcond.w; cg.bias = ctx->dcond.b; cg.Cout = U; cg.k = 1;
        cg.y = dg; cg.y_ctot = U; cg.Ty = T; cg.n_out = T; cg.B = B;
        launch_conv_f32(cg, st);
        cp.bias_t = dg; cp.bias_t_ctot = U;
    }
    // The SnakeAlias variant is a time-domain filter around every activation, not an elementwise op.  Tensor-core path: it
    // is computed by the LOADER of the conv kernel (convn_tc_kernel<.., SNAKE>), one launch per convolution; fp32 path: its
    // own kernel in front of the FFMA convolutions.
    const bool snake = c.snake != 0;
    float* Sb = snake ? reinterpret_cast<float*>(ws + pl.off_S) : nullptr;
    const bool gen_tc = ctx->precision == SVB_PREC_TC && ctx->gen_tc_ok;
    if (ctx->precision == SVB_PREC_TC && !gen_tc) note_fallback(ctx, "generator: layer shapes not served by the tensor-core kernels");
    if (gen_tc) {
        const ConvNW& W = ctx->conv_pre_tc;
        ConvNTC a;
        a.x = z; a.x_ctot = Cpre; a.cin_real = W.cin_real; a.cinp = W.cinp; a.Tin = T;
        a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale;
        if (!melv && gT == 1) { a.bias_b = dg; a.bias_b_stride = U; a.bias_b_off = 0; }
        else if (!melv) { a.bias_t = dg; a.bias_t_ctot = U; a.bias_t_c0 = 0; }      // speaker mix: cond(g)[b, :, t] per frame
        a.k = W.k; a.pad_left = W.pad_left; a.n_rows = T; a.N_total = W.N_total; a.NC = W.NC; a.chunks_per_cta = 1; a.Ty = T; a.B = B;
        a.seg[0].y = pre; a.seg[0].y_ctot = U; a.seg[0].col0 = 0; a.seg[0].col1 = U;
        int trc = launch_convn_tc(a, st);
        if (trc) return fail(ctx, trc, "convn launch failed (conv_pre)");
    } else {
        launch_conv_f32(cp, st);
    }
    if ((rc = dbg_keep(ctx, "conv_pre", pre, (size_t)B * U * T, st))) return rc;

    const float* cur = pre;
    int Lin = T;
    const int nk = c.n_resblock_kernels;
    for (int i = 0; i < c.n_upsamples; ++i) {
        Stage& S = ctx->stages[i];
        const int Lout = Lin * S.s;
        // x = ups(lrelu(x, 0.1))  as s polyphase 2-tap convs
        if (snake && !gen_tc) launch_snake_alias(cur, Sb, S.snake_in.ealpha, S.snake_in.inv_beta, ctx->snake_filt, B, S.Cin, Lin, st);
        ConvF32 up;
        up.x = snake ? Sb : cur; up.x_ctot = S.Cin; up.Cin = S.Cin; up.Tin = Lin;
        up.w = S.up_w; up.bias = S.up_b; up.Cout = S.Cout; up.k = 2; up.dil = 1; up.pad_left = 1;
        up.n_phase = S.s; up.w_phase_stride = (long long)S.Cin * 2 * S.Cout;
        up.ostride = S.s; up.ooff = -S.p; up.n_out = Lin + 1;
        up.in_act = snake ? 0 : 1; up.in_slope = 0.1f;
        up.y = X; up.y_ctot = S.Cout; up.Ty = Lout; up.B = B;
        if (gen_tc) {
            const ConvNW& W = S.up_tc;
            ConvNTC a;
            a.x = cur; a.x_ctot = S.Cin; a.cin_real = S.Cin; a.cinp = W.cinp; a.Tin = Lin; a.in_act = snake ? 0 : 1; a.in_slope = 0.1f;
            if (snake) { a.snake_ealpha = S.snake_in.ealpha; a.snake_invbeta = S.snake_in.inv_beta; a.snake_filt = ctx->snake_filt; }
            a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = 2; a.pad_left = 1; a.n_rows = Lin + 1; a.N_total = W.N_total; a.NC = W.NC;
            a.chunks_per_cta = pick_chunks_per_cta(Lin + 1, 128 * (snake ? convn_snake_mb(W.cinp) : convn_ups_mb(W.cinp)), B, (W.N_total + W.NC - 1) / W.NC);
            a.mode = 1; a.s = S.s; a.p = S.p; a.Ty = Lout; a.B = B;
            a.seg[0].y = X; a.seg[0].y_ctot = S.Cout;
            if (W.noise) { a.har = har; a.har_N = (int)N; a.noise_stride = W.noise_stride; a.noise_w0 = W.noise_w0; a.noise_wide = (W.noise == 2); }
            ProfScope ps(ctx, "ups_tc", st, 2.0 * 2.0 * S.Cin * (double)S.Cout * (double)Lout * B,
                         ((double)S.Cin * Lin + (double)S.Cout * Lout + (W.noise ? (double)N : 0.0)) * B * sizeof(float));
            int trc = launch_convn_tc(a, st);
            if (trc) return fail(ctx, trc, "convn launch failed (ups)");
        } else {
            launch_conv_f32(up, st);
        }
        if (gen_tc && !S.up_tc.noise && S.noise_tc.img) {
            const ConvNW& W = S.noise_tc;
            ConvNTC a;
            a.x = har; a.cin_real = 64; a.cinp = 64; a.Tin = Lout + 1;
            a.view_bstride = N; a.view_tstride = S.noise_s; a.view_cstride = 1; a.view_off = -S.noise_p; a.view_limit = N;
            a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = 2; a.pad_left = 0; a.n_rows = Lout; a.N_total = W.N_total; a.NC = W.NC;
            a.chunks_per_cta = (W.N_total + W.NC - 1) / W.NC; a.Ty = Lout; a.B = B;
            a.seg[0].y = X; a.seg[0].y_ctot = S.Cout; a.seg[0].col0 = 0; a.seg[0].col1 = S.Cout; a.seg[0].beta = 1.f;
            int trc = launch_convn_tc(a, st);
            if (trc) return fail(ctx, trc, "convn launch failed (noise conv)");
        } else if (!(gen_tc && S.up_tc.noise)) {
            launch_noise_conv_add(har, S.noise_w, S.noise_b, X, B, S.Cout, Lout, (int)N, S.noise_K, S.noise_s, S.noise_p, st);
        }
        if ((rc = dbg_keep(ctx, "ups" + std::to_string(i), X, (size_t)B * S.Cout * Lout, st))) return rc;
        const int fuse_rb = ctx->opt_fuse_rb;
        const int use_tma = ctx->opt_tma;
        void* a16[2] = {ws + pl.off_A16, ws + pl.off_B16};
        const int fuse_maxc = ctx->opt_fuse_maxc;
        bool merged_done = false;
        if (gen_tc && !snake && nk == 3 && ctx->opt_merge_branches && pl.br_elems >= (size_t)B * S.Cout * (size_t)Lout && merge_branches(S.Cout, Lout, B) &&
            S.c1[0].w_tc && S.c1[3].w_tc && S.c1[6].w_tc) {
            // the three branches of this stage advance together: one launch per dilation index, blockIdx.z = branch; every
            // branch keeps its own ping-pong buffers, the three last pairs accumulate alpha*y into the zeroed stage output
            CU(cudaMemsetAsync(O, 0, (size_t)B * S.Cout * (size_t)Lout * sizeof(float), st));
            int trc = 0;
            for (int d = 0; d < 3 && trc == 0; ++d) {
                PairTC pt[3];
                double fl = 0;
                for (int j = 0; j < 3; ++j) {
                    const int k = c.resblock_kernel_sizes[j];
                    const ConvW& W1 = S.c1[j * 3 + d];
                    const ConvW& W2 = S.c2[j * 3 + d];
                    float* bufj[2] = {reinterpret_cast<float*>(ws + pl.off_br[2 * j]), reinterpret_cast<float*>(ws + pl.off_br[2 * j + 1])};
                    pt[j].x = d == 0 ? X : bufj[(d - 1) & 1];
                    pt[j].out = d == 2 ? O : bufj[d & 1];
                    pt[j].w1 = W1.w_tc; pt[j].w2 = W2.w_tc; pt[j].b1 = W1.b_tc; pt[j].b2 = W2.b;
                    pt[j].inv = 1.f / (W1.tc_scale * W2.tc_scale);
                    pt[j].B = B; pt[j].C = S.Cout; pt[j].T = Lout; pt[j].k = k; pt[j].dil = c.resblock_dilations[j][d];
                    pt[j].alpha = d == 2 ? 1.f / nk : 1.f; pt[j].beta = d == 2 ? 1.f : 0.f;
                    fl += 2.0 * 2.0 * S.Cout * (double)S.Cout * k * (double)Lout * B;
                }
                if (d < 2) {
                    ProfScope ps(ctx, "pair_tc", st, fl, 3 * 3.0 * S.Cout * (double)Lout * B * sizeof(float));
                    trc = launch_pair_tc_multi(pt, 3, st);
                } else {
                    // The last pairs accumulate alpha*y into the stage output.  Three floating-point reductions in arrival order
                    // would make the sum depend on the schedule ((a+b)+c vs (a+c)+b), so only TWO branches reduce into the zeroed
                    // buffer (0 + a + b is exact in either order) and the third adds afterwards in a launch of its own: the
                    // result is bit-reproducible from run to run and independent of the batch an item travels in.
                    // The shortest and the longest kernel share the merged launch (memory-phase bound next to MMA bound), the middle
                    // one follows alone.
                    const double fl1 = 2.0 * 2.0 * S.Cout * (double)S.Cout * c.resblock_kernel_sizes[1] * (double)Lout * B;
                    const PairTC outer[2] = {pt[0], pt[2]};
                    {
                        ProfScope ps(ctx, "pair_tc", st, fl - fl1, 2 * 3.0 * S.Cout * (double)Lout * B * sizeof(float));
                        trc = launch_pair_tc_multi(outer, 2, st);
                    }
                    if (trc == 0) {
                        ProfScope ps(ctx, "pair_tc", st, fl1, 3.0 * S.Cout * (double)Lout * B * sizeof(float));
                        trc = launch_pair_tc_multi(pt + 1, 1, st);
                    }
                }
            }
            if (trc == 0) merged_done = true;
            else if (trc != SVB_ERR_UNSUPPORTED) return fail(ctx, trc, "branch-merged pair launch failed");
        }
        for (int j = 0; j < nk && !merged_done; ++j) {
            const int k = c.resblock_kernel_sizes[j];
            if (gen_tc && snake && !S.c1n.empty()) {
                // Snake ResBlock (vdecoder/hifiganwithsnake/models.py:61-72): x = x + c2(a2(c1(a1(x)))) per dilation, every
                // convolution one tcgen05 launch with the SnakeAlias activation computed by its loader
                const float* src = X;
                float* pp[2] = {A, Bb};
                for (int d = 0; d < 3; ++d) {
                    const int dil = c.resblock_dilations[j][d];
                    const bool last = (d == 2);
                    float* dst = last ? O : pp[d & 1];
                    auto conv = [&](const ConvNW& W, const SnakeP& act, const float* x, int cdil, float* y, const float* res, float alpha, float beta) {
                        ConvNTC a;
                        a.x = x; a.x_ctot = S.Cout; a.cin_real = S.Cout; a.cinp = W.cinp; a.Tin = Lout;
                        a.snake_ealpha = act.ealpha; a.snake_invbeta = act.inv_beta; a.snake_filt = ctx->snake_filt;
                        a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = k; a.dil = cdil; a.pad_left = cdil * (k - 1) / 2;
                        a.n_rows = Lout; a.N_total = W.N_total; a.NC = W.NC; a.chunks_per_cta = (W.N_total + W.NC - 1) / W.NC; a.Ty = Lout; a.B = B;
                        a.seg[0].y = y; a.seg[0].y_ctot = S.Cout; a.seg[0].col0 = 0; a.seg[0].col1 = S.Cout;
                        a.seg[0].res = res; a.seg[0].res_ctot = S.Cout; a.seg[0].alpha = alpha; a.seg[0].beta = beta;
                        return launch_convn_tc(a, st);
                    };
                    const double cflops = 2.0 * S.Cout * (double)S.Cout * k * (double)Lout * B;
                    int trc;
                    {
                        ProfScope ps(ctx, "snake_conv", st, cflops, 2.0 * S.Cout * (double)Lout * B * sizeof(float));
                        trc = conv(S.c1n[j * 3 + d], S.acts[j * 6 + 2 * d], src, dil, Tm, nullptr, 1.f, 0.f);
                    }
                    if (trc) return fail(ctx, trc, "snake conv launch failed (c1)");
                    {
                        ProfScope ps(ctx, "snake_conv", st, cflops, 3.0 * S.Cout * (double)Lout * B * sizeof(float));
                        trc = conv(S.c2n[j * 3 + d], S.acts[j * 6 + 2 * d + 1], Tm, 1, dst, src, last ? 1.f / nk : 1.f, (last && j > 0) ? 1.f : 0.f);
                    }
                    if (trc) return fail(ctx, trc, "snake conv launch failed (c2)");
                    src = dst;
                }
                continue;
            }
            if (ctx->precision == SVB_PREC_TC && !snake && fuse_rb && S.Cout <= fuse_maxc && S.c1[j * 3].w_tc) {
                // narrow stages: the whole ResBlock in one kernel (residual stream in TMEM)
                ResblockTC rb;
                rb.x = X; rb.out = O; rb.B = B; rb.C = S.Cout; rb.T = Lout; rb.k = k;
                for (int d = 0; d < 3; ++d) {
                    rb.dil[d] = c.resblock_dilations[j][d];
                    rb.w[2 * d] = S.c1[j * 3 + d].w_tc; rb.w[2 * d + 1] = S.c2[j * 3 + d].w_tc;
                    rb.bias[2 * d] = S.c1[j * 3 + d].b_tc; rb.bias[2 * d + 1] = S.c2[j * 3 + d].b;
                    rb.inv[d] = 1.f / (S.c1[j * 3 + d].tc_scale * S.c2[j * 3 + d].tc_scale);
                }
                rb.alpha = 1.f / nk; rb.beta = (j > 0) ? 1.f : 0.f;
                const double rflops = 3 * 2.0 * 2.0 * S.Cout * (double)S.Cout * k * (double)Lout * B;
                ProfScope ps(ctx, "resblock_tc", st, rflops, 2.0 * S.Cout * (double)Lout * B * sizeof(float));
                int trc = launch_resblock_tc(rb, st);
                if (trc == 0) continue;
                if (trc != SVB_ERR_UNSUPPORTED) return fail(ctx, trc, "fused ResBlock kernel launch failed");
            }
            const float* src = X;
            float* pp[2] = {A, Bb};
            for (int d = 0; d < 3; ++d) {
                const int dil = c.resblock_dilations[j][d];
                const ConvW& W1 = S.c1[j * 3 + d];
                const ConvW& W2 = S.c2[j * 3 + d];
                const bool last = (d == 2);
                float* dst = last ? O : pp[d & 1];
                const float alpha = last ? 1.f / nk : 1.f;
                const float beta = (last && j > 0) ? 1.f : 0.f;
                bool done = false;
                const double pair_flops = 2.0 * 2.0 * S.Cout * (double)S.Cout * k * (double)Lout * B;
                const double pair_bytes = 3.0 * S.Cout * (double)Lout * B * sizeof(float);   // read x (tile + residual) + write out
                ProfScope ps(ctx, (ctx->precision == SVB_PREC_TC && W1.w_tc && !snake) ? "pair_tc" : "pair_f32", st, pair_flops, pair_bytes);
                if (ctx->precision == SVB_PREC_TC && !snake && W1.w_tc && W2.w_tc) {
                    PairTC pt;
                    pt.x = src; pt.out = dst; pt.w1 = W1.w_tc; pt.w2 = W2.w_tc; pt.b1 = W1.b_tc; pt.b2 = W2.b;
                    pt.inv = 1.f / (W1.tc_scale * W2.tc_scale);
                    pt.B = B; pt.C = S.Cout; pt.T = Lout; pt.k = k; pt.dil = dil; pt.alpha = alpha; pt.beta = beta;
                    if (use_tma && pair_tc_supports_tma(S.Cout, -1)) {
                        // pair d writes lrelu(out) as fp16 [B][T][C]; pair d+1 loads its operand tile from it with TMA
                        if (d > 0) pt.a16_in = a16[(d - 1) & 1];
                        if (!last) pt.a16_out = a16[d & 1];
                    }
                    int trc = launch_pair_tc(pt, st);
                    if (trc == 0) done = true;
                    else if (trc != SVB_ERR_UNSUPPORTED) return fail(ctx, trc, "tensor-core pair kernel launch failed");
                }
                if (!done) {
                    if (snake) { const SnakeP& a1 = S.acts[j * 6 + 2 * d]; launch_snake_alias(src, Sb, a1.ealpha, a1.inv_beta, ctx->snake_filt, B, S.Cout, Lout, st); }
                    ConvF32 c1;
                    c1.x = snake ? Sb : src; c1.x_ctot = S.Cout; c1.Cin = S.Cout; c1.Tin = Lout;
                    c1.w = W1.w; c1.bias = W1.b; c1.Cout = S.Cout; c1.k = k; c1.dil = dil; c1.pad_left = dil * (k - 1) / 2;
                    c1.in_act = snake ? 0 : 1; c1.in_slope = 0.1f;
                    c1.y = Tm; c1.y_ctot = S.Cout; c1.Ty = Lout; c1.n_out = Lout; c1.B = B;
                    launch_conv_f32(c1, st);
                    if (snake) { const SnakeP& a2 = S.acts[j * 6 + 2 * d + 1]; launch_snake_alias(Tm, Sb, a2.ealpha, a2.inv_beta, ctx->snake_filt, B, S.Cout, Lout, st); }
                    ConvF32 c2;
                    c2.x = snake ? Sb : Tm; c2.x_ctot = S.Cout; c2.Cin = S.Cout; c2.Tin = Lout;
                    c2.w = W2.w; c2.bias = W2.b; c2.Cout = S.Cout; c2.k = k; c2.dil = 1; c2.pad_left = (k - 1) / 2;
                    c2.in_act = snake ? 0 : 1; c2.in_slope = 0.1f;
                    c2.res = src; c2.res_ctot = S.Cout;
                    c2.y = dst; c2.y_ctot = S.Cout; c2.Ty = Lout; c2.n_out = Lout; c2.B = B;
                    c2.alpha = alpha; c2.beta = beta;
                    launch_conv_f32(c2, st);
                }
                src = dst;
            }
        }
        if ((rc = dbg_keep(ctx, "stage" + std::to_string(i), O, (size_t)B * S.Cout * Lout, st))) return rc;
        // O becomes the next stage's input; swap roles so the next stage does not overwrite it
        float* tmp = O; O = Tm; Tm = tmp;   // next stage writes its output into the old T buffer
        cur = tmp;
        Lin = Lout;
    }
    {
        ProfScope ps(ctx, "conv_post", st, 0, (double)(ctx->post_C + 1) * (double)Lin * B * sizeof(float));
        if (snake) {
            launch_snake_alias(cur, Sb, ctx->snake_post.ealpha, ctx->snake_post.inv_beta, ctx->snake_filt, B, ctx->post_C, Lin, st);
            launch_conv_post(Sb, ctx->post_w, ctx->post_b, wav, B, ctx->post_C, Lin, ctx->post_K, 1.0f, st);   // slope 1 = no activation
        } else {
            launch_conv_post(cur, ctx->post_w, ctx->post_b, wav, B, ctx->post_C, Lin, ctx->post_K, 0.01f, st);
        }
    }
    return check_launch(ctx, "generator");
}

}  // namespace

// ================================================================================================ ABI
extern "C" {

const char* svb_version(void) { return SVB_VERSION; }

const char* svb_strerror(int s) {
    switch (s) {
        case SVB_OK: return "ok";
        case SVB_ERR_INVALID_ARG: return "invalid argument";
        case SVB_ERR_CUDA: return "CUDA error";
        case SVB_ERR_NOT_LOADED: return "weights not loaded";
        case SVB_ERR_MISSING_TENSOR: return "missing tensor";
        case SVB_ERR_SHAPE: return "tensor shape mismatch";
        case SVB_ERR_UNSUPPORTED: return "unsupported configuration";
        case SVB_ERR_WORKSPACE: return "workspace too small";
        case SVB_ERR_ARCH: return "device is not sm_100";
        default: return "unknown status";
    }
}

const char* svb_last_error(const svb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t svb_launch_count(const svb_ctx*) { return launch_counter(); }
int64_t svb_fallback_count(const svb_ctx* ctx) { return ctx ? ctx->ffma_fallbacks : -1; }

int svb_create(int device, svb_ctx** out) {
    if (!out) return SVB_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return SVB_ERR_CUDA;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return SVB_ERR_CUDA;
    if (p.major != 10) return SVB_ERR_ARCH;   // sm_100a binary only: fail loudly elsewhere
    DevGuard dg(device);
    if (!dg.ok) return SVB_ERR_CUDA;
    svb_ctx* c = new svb_ctx();
    c->device = device;
    if (const char* e = std::getenv("SVB_TC_TMA")) c->opt_tma = std::atoi(e);
    if (const char* e = std::getenv("SVB_FUSE_RESBLOCK")) c->opt_fuse_rb = std::atoi(e);
    if (const char* e = std::getenv("SVB_FUSE_MAXC")) c->opt_fuse_maxc = std::atoi(e);
    if (const char* e = std::getenv("SVB_FUSE_FLOW")) c->opt_fuse_flow = std::atoi(e);
    *out = c;
    return SVB_OK;
}

void svb_destroy(svb_ctx* ctx) {
    if (!ctx) return;
    DevGuard dg(ctx->device);
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->ws.p) cudaFree(ctx->ws.p);
    if (ctx->ws_prefix.p) cudaFree(ctx->ws_prefix.p);
    if (ctx->host_io.p) cudaFree(ctx->host_io.p);
    for (auto& kv : ctx->dbg) if (kv.second.p) cudaFree(kv.second.p);
    delete ctx;
}

int svb_set_precision(svb_ctx* ctx, int precision) {
    if (!ctx || (precision != SVB_PREC_FP32 && precision != SVB_PREC_TC)) return SVB_ERR_INVALID_ARG;
    ctx->precision = precision;
    return SVB_OK;
}
int svb_get_precision(const svb_ctx* ctx) { return ctx ? ctx->precision : SVB_ERR_INVALID_ARG; }

int svb_set_option(svb_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return SVB_ERR_INVALID_ARG;
    const std::string n(name);
    if (n == "tma") ctx->opt_tma = value;
    else if (n == "fuse_resblock") ctx->opt_fuse_rb = value;
    else if (n == "fuse_maxc") ctx->opt_fuse_maxc = value;
    else if (n == "fuse_flow") ctx->opt_fuse_flow = value;
    else if (n == "merge_branches") ctx->opt_merge_branches = value;
    else if (n == "philox_noise") ctx->opt_philox = value;
    else if (n == "philox_seed") ctx->philox_seed = (unsigned long long)(unsigned int)value;
    else return fail(ctx, SVB_ERR_INVALID_ARG, "unknown option " + n);
    return SVB_OK;
}

int svb_debug_enable(svb_ctx* ctx, int on) {
    if (!ctx) return SVB_ERR_INVALID_ARG;
    ctx->debug = on != 0;
    return SVB_OK;
}

// One ResBlock pair (stage, branch j, dilation index d) of the loaded generator on caller buffers [B,C,L]:
// variant >= 0 -> tensor-core tile variant, variant == -2 -> the two fp32 FFMA convs.  Microbenchmark / unit-test hook.
int svb_debug_pair(svb_ctx* ctx, int stage, int j, int d, const float* x, float* out, float* scratch, int B, int L, int variant,
                   float alpha, float beta, void* stream) {
    if (!ctx || !ctx->loaded) return SVB_ERR_NOT_LOADED;
    if (stage < 0 || stage >= ctx->cfg.n_upsamples || j < 0 || j >= 3 || d < 0 || d >= 3 || !x || !out) return SVB_ERR_INVALID_ARG;
    DevGuard dg__(ctx->device);
    if (!dg__.ok) return fail(ctx, SVB_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    Stage& S = ctx->stages[stage];
    const int k = ctx->cfg.resblock_kernel_sizes[j], dil = ctx->cfg.resblock_dilations[j][d];
    const ConvW& W1 = S.c1[j * 3 + d];
    const ConvW& W2 = S.c2[j * 3 + d];
    if (variant >= 0) {
        PairTC pt;
        pt.x = x; pt.out = out; pt.w1 = W1.w_tc; pt.w2 = W2.w_tc; pt.b1 = W1.b_tc; pt.b2 = W2.b;
        pt.inv = 1.f / (W1.tc_scale * W2.tc_scale);
        pt.B = B; pt.C = S.Cout; pt.T = L; pt.k = k; pt.dil = dil; pt.alpha = alpha; pt.beta = beta; pt.variant = variant;
        int rc = launch_pair_tc(pt, st);
        if (rc) return fail(ctx, rc, "pair kernel launch failed");
        return check_launch(ctx, "debug_pair");
    }
    if (!scratch) return SVB_ERR_INVALID_ARG;
    ConvF32 c1;
    c1.x = x; c1.x_ctot = S.Cout; c1.Cin = S.Cout; c1.Tin = L;
    c1.w = W1.w; c1.bias = W1.b; c1.Cout = S.Cout; c1.k = k; c1.dil = dil; c1.pad_left = dil * (k - 1) / 2;
    c1.in_act = 1; c1.in_slope = 0.1f;
    c1.y = scratch; c1.y_ctot = S.Cout; c1.Ty = L; c1.n_out = L; c1.B = B;
    launch_conv_f32(c1, st);
    ConvF32 c2;
    c2.x = scratch; c2.x_ctot = S.Cout; c2.Cin = S.Cout; c2.Tin = L;
    c2.w = W2.w; c2.bias = W2.b; c2.Cout = S.Cout; c2.k = k; c2.dil = 1; c2.pad_left = (k - 1) / 2;
    c2.in_act = 1; c2.in_slope = 0.1f;
    c2.res = x; c2.res_ctot = S.Cout;
    c2.y = out; c2.y_ctot = S.Cout; c2.Ty = L; c2.n_out = L; c2.B = B;
    c2.alpha = alpha; c2.beta = beta;
    launch_conv_f32(c2, st);
    return check_launch(ctx, "debug_pair");
}

// One whole ResBlock branch j of a stage through the fused tensor-core kernel (variant >= 0).
int svb_debug_resblock(svb_ctx* ctx, int stage, int j, const float* x, float* out, int B, int L, int variant,
                       float alpha, float beta, void* stream) {
    if (!ctx || !ctx->loaded) return SVB_ERR_NOT_LOADED;
    if (stage < 0 || stage >= ctx->cfg.n_upsamples || j < 0 || j >= 3 || !x || !out) return SVB_ERR_INVALID_ARG;
    DevGuard dg__(ctx->device);
    if (!dg__.ok) return fail(ctx, SVB_ERR_CUDA, "cudaSetDevice failed");
    Stage& S = ctx->stages[stage];
    ResblockTC rb;
    rb.x = x; rb.out = out; rb.B = B; rb.C = S.Cout; rb.T = L; rb.k = ctx->cfg.resblock_kernel_sizes[j];
    for (int d = 0; d < 3; ++d) {
        rb.dil[d] = ctx->cfg.resblock_dilations[j][d];
        rb.w[2 * d] = S.c1[j * 3 + d].w_tc; rb.w[2 * d + 1] = S.c2[j * 3 + d].w_tc;
        rb.bias[2 * d] = S.c1[j * 3 + d].b_tc; rb.bias[2 * d + 1] = S.c2[j * 3 + d].b;
        rb.inv[d] = 1.f / (S.c1[j * 3 + d].tc_scale * S.c2[j * 3 + d].tc_scale);
    }
    rb.alpha = alpha; rb.beta = beta; rb.variant = variant;
    int rc = launch_resblock_tc(rb, (cudaStream_t)stream);
    if (rc) return fail(ctx, rc, "fused ResBlock kernel launch failed");
    return check_launch(ctx, "debug_resblock");
}

int svb_profile_enable(svb_ctx* ctx, int on) {
    if (!ctx) return SVB_ERR_INVALID_ARG;
    for (auto& kv : ctx->prof)
        for (auto& e : kv.second.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    ctx->prof.clear();
    ctx->profile = on != 0;
    return SVB_OK;
}

int svb_profile_read(svb_ctx* ctx, const char* name, double* total_ms, int64_t* count, double* flops, double* bytes) {
    if (!ctx || !name) return SVB_ERR_INVALID_ARG;
    auto it = ctx->prof.find(name);
    if (it == ctx->prof.end()) return fail(ctx, SVB_ERR_INVALID_ARG, std::string("no profile entry named ") + name);
    double ms = 0;
    for (auto& e : it->second.ev) {
        CU(cudaEventSynchronize(e.second));
        float t = 0;
        CU(cudaEventElapsedTime(&t, e.first, e.second));
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (count) *count = (int64_t)it->second.ev.size();
    if (flops) *flops = it->second.flops;
    if (bytes) *bytes = it->second.bytes;
    return SVB_OK;
}

int svb_debug_fetch(svb_ctx* ctx, const char* what, float* dst, size_t n, void* stream) {
    if (!ctx || !what || !dst) return SVB_ERR_INVALID_ARG;
    auto it = ctx->dbg.find(what);
    if (it == ctx->dbg.end()) return fail(ctx, SVB_ERR_INVALID_ARG, std::string("no debug tap named ") + what);
    if (n * sizeof(float) > it->second.bytes) return fail(ctx, SVB_ERR_INVALID_ARG, "debug tap smaller than requested");
    CU(cudaMemcpyAsync(dst, it->second.p, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SVB_OK;
}

int svb_load_weights(svb_ctx* ctx, const svb_tensor* tensors, int n_tensors, const svb_model_cfg* cfgp) {
    if (!ctx || !tensors || !cfgp || n_tensors <= 0) return SVB_ERR_INVALID_ARG;
    DevGuard dg__(ctx->device);
    if (!dg__.ok) return fail(ctx, SVB_ERR_CUDA, "cudaSetDevice failed");
    const svb_model_cfg& c = *cfgp;
    if (c.num_mels > 0 && c.snake) return fail(ctx, SVB_ERR_UNSUPPORTED, "mel vocoder has no snake variant");
    if (c.n_upsamples < 1 || c.n_upsamples > 8 || c.n_resblock_kernels != 3 || (c.num_mels == 0 && c.n_flows != 4) ||
        (c.inter_channels & 1) || c.flow_wn_layers < 1 || (c.n_harmonics != 9 && c.n_harmonics != 1))
        return fail(ctx, SVB_ERR_UNSUPPORTED, "unsupported model configuration");
    for (int i = 0; i < c.n_upsamples; ++i)
        if (c.upsample_kernel_sizes[i] != 2 * c.upsample_rates[i])
            return fail(ctx, SVB_ERR_UNSUPPORTED, "upsample kernel size must be 2x the rate (polyphase 2-tap form)");
    if ((c.upsample_initial_channel >> c.n_upsamples) < 1 || (c.hidden_channels % 8) || (c.inter_channels % 16))
        return fail(ctx, SVB_ERR_UNSUPPORTED, "channel counts must be multiples of 8");
    // A context may be re-loaded (the Python side re-packs after .to()/.half()/load_state_dict): drop the previous weight
    // images first, and leave the context "not loaded" unless this call succeeds completely.
    if (!ctx->allocs.empty() || ctx->loaded) {
        CU(cudaDeviceSynchronize());
        for (void* p : ctx->allocs) cudaFree(p);
        ctx->allocs.clear();
    }
    ctx->loaded = false;
    ctx->flow.clear(); ctx->stages.clear();
    ctx->conv_pre = ConvW(); ctx->conv_pre_tc = ConvNW(); ctx->dcond = ConvW();
    ctx->dcond_w_nat = nullptr; ctx->dcond_b = nullptr; ctx->post_w = nullptr; ctx->lin_w = nullptr; ctx->snake_filt = nullptr;
    ctx->snake_post = SnakeP();
    ctx->flow_tc_ok = false; ctx->gen_tc_ok = false;
    ctx->cond_all_w = nullptr; ctx->cond_all_b = nullptr;
    ctx->prefix = Prefix();
    TMap m;
    for (int i = 0; i < n_tensors; ++i)
        if (tensors[i].name && tensors[i].data) m[tensors[i].name] = &tensors[i];
    const bool melv = c.num_mels > 0;
    const std::string DP = melv ? "" : "dec.";          // vdecoder/nsf_hifigan checkpoints have no "dec." prefix
    ctx->cfg = c;
    ctx->hop = 1;
    for (int i = 0; i < c.n_upsamples; ++i) ctx->hop *= c.upsample_rates[i];
    const int H = c.hidden_channels, C = c.inter_channels, G = c.gin_channels, L = c.flow_wn_layers, half = C / 2;
    int rc;
    HostT w, b;

    // ---- flow
    ctx->flow.assign(melv ? 0 : c.n_flows, FlowLayer());
    for (int fl = 0; fl < (melv ? 0 : c.n_flows); ++fl) {
        FlowLayer& F = ctx->flow[fl];
        const std::string p = "flow.flows." + std::to_string(2 * fl) + ".";
        const bool odd = (fl & 1) != 0;           // Flip folded into channel order
        F.in_c0 = odd ? half : 0;
        F.out_c0 = odd ? 0 : half;
        if ((rc = get_tensor(ctx, m, p + "pre.weight", {H, half, 1}, w))) return rc;
        if ((rc = get_tensor(ctx, m, p + "pre.bias", {H}, b))) return rc;
        if ((rc = make_conv(ctx, w.v, b.v, H, half, 1, odd, false, F.pre))) return rc;
        const std::vector<float> h_pre_w = w.v, h_pre_b = b.v;
        std::vector<float> h_post_w, h_post_b, h_cond_w, h_cond_b;
        std::vector<std::vector<float>> h_in_w(L), h_in_b(L), h_rs_w(L), h_rs_b(L);
        const bool flow_tc = (H == 192) && (half <= H) && (half % 32 == 0) && (c.flow_kernel_size - 1 <= 8);
        ctx->flow_tc_ok = flow_tc;
        if (flow_tc) {
            const std::vector<float> wv = w.v, bv = b.v;
            if ((rc = make_convn(ctx, H, half, H, H, 1, 0,
                                 [&](int col, int ci, int) { return wv[(size_t)col * half + (odd ? half - 1 - ci : ci)]; },
                                 [&](int col) { return bv[col]; }, F.pre_tc))) return rc;
        }
        if ((rc = get_tensor(ctx, m, p + "post.weight", {half, H, 1}, w))) return rc;
        if ((rc = get_tensor(ctx, m, p + "post.bias", {half}, b))) return rc;
        if ((rc = make_conv(ctx, w.v, b.v, half, H, 1, false, odd, F.post))) return rc;
        h_post_w = w.v; h_post_b = b.v;
        if (flow_tc) {
            const std::vector<float> wv = w.v, bv = b.v;
            if ((rc = make_convn(ctx, H, H, half, half, 1, 0,
                                 [&](int col, int ci, int) { return wv[(size_t)(odd ? half - 1 - col : col) * H + ci]; },
                                 [&](int col) { return bv[odd ? half - 1 - col : col]; }, F.post_tc))) return rc;
        }
        if ((rc = folded(ctx, m, p + "enc.cond_layer", {2 * H * L, G, 1}, w))) return rc;
        if ((rc = get_tensor(ctx, m, p + "enc.cond_layer.bias", {2 * H * L}, b))) return rc;
        if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&F.cond_w_nat))) return rc;
        if ((rc = make_conv(ctx, w.v, b.v, 2 * H * L, G, 1, false, false, F.cond))) return rc;
        h_cond_w = w.v; h_cond_b = b.v;
        // gate kernel column order: chunk c (of H columns) = [tanh channels c*H/2.. | sigmoid channels c*H/2..]
        auto gate_row = [H](int col) { const int cc = col / H, j = col % H, hh = H / 2; return j < hh ? cc * hh + j : H + cc * hh + (j - hh); };
        if (flow_tc) {
            std::vector<float> wp(w.v.size()), bp(b.v.size());
            for (int i = 0; i < L; ++i)
                for (int col = 0; col < 2 * H; ++col) {
                    const int src = 2 * H * i + gate_row(col), dst = 2 * H * i + col;
                    std::memcpy(&wp[(size_t)dst * G], &w.v[(size_t)src * G], sizeof(float) * G);
                    bp[dst] = b.v[src];
                }
            if ((rc = upload(ctx, wp.data(), wp.size() * sizeof(float), (void**)&F.cond_w_perm))) return rc;
            if ((rc = upload(ctx, bp.data(), bp.size() * sizeof(float), (void**)&F.cond_b_perm))) return rc;
            F.in_tc.assign(L, ConvNW());
            F.rs_tc.assign(L, ConvNW());
        }
        F.in_layers.assign(L, ConvW());
        F.res_skip.assign(L, ConvW());
        for (int i = 0; i < L; ++i) {
            const std::string q = p + "enc.in_layers." + std::to_string(i);
            if ((rc = folded(ctx, m, q, {2 * H, H, c.flow_kernel_size}, w))) return rc;
            if ((rc = get_tensor(ctx, m, q + ".bias", {2 * H}, b))) return rc;
            if ((rc = make_conv(ctx, w.v, b.v, 2 * H, H, c.flow_kernel_size, false, false, F.in_layers[i]))) return rc;
            h_in_w[i] = w.v; h_in_b[i] = b.v;
            if (flow_tc) {
                const int kk = c.flow_kernel_size;
                const std::vector<float> wv = w.v, bv = b.v;
                if ((rc = make_convn(ctx, H, H, 2 * H, H, kk, (kk - 1) / 2,
                                     [&](int col, int ci, int tap) { return wv[((size_t)gate_row(col) * H + ci) * kk + tap]; },
                                     [&](int col) { return bv[gate_row(col)]; }, F.in_tc[i]))) return rc;
            }
            const std::string r = p + "enc.res_skip_layers." + std::to_string(i);
            const int co = (i < L - 1) ? 2 * H : H;
            if ((rc = folded(ctx, m, r, {co, H, 1}, w))) return rc;
            if ((rc = get_tensor(ctx, m, r + ".bias", {co}, b))) return rc;
            h_rs_w[i] = w.v; h_rs_b[i] = b.v;
            if (flow_tc) {
                const std::vector<float> wv = w.v, bv = b.v;
                if ((rc = make_convn(ctx, H, H, co, H, 1, 0, [&](int col, int ci, int) { return wv[(size_t)col * H + ci]; },
                                     [&](int col) { return bv[col]; }, F.rs_tc[i]))) return rc;
            }
            if (co == 2 * H) {
                // pack the residual half and the skip half as two consecutive [H][1][H] blocks
                std::vector<float> pk((size_t)2 * H * H), bb(2 * H);
                for (int hsel = 0; hsel < 2; ++hsel)
                    for (int ci = 0; ci < H; ++ci)
                        for (int o = 0; o < H; ++o)
                            pk[(size_t)hsel * H * H + (size_t)ci * H + o] = w.v[(size_t)(hsel * H + o) * H + ci];
                for (int o = 0; o < 2 * H; ++o) bb[o] = b.v[o];
                ConvW& R = F.res_skip[i];
                R.Cin = H; R.Cout = 2 * H; R.k = 1;
                if ((rc = upload(ctx, pk.data(), pk.size() * sizeof(float), (void**)&R.w))) return rc;
                if ((rc = upload(ctx, bb.data(), bb.size() * sizeof(float), (void**)&R.b))) return rc;
            } else {
                if ((rc = make_conv(ctx, w.v, b.v, H, H, 1, false, false, F.res_skip[i]))) return rc;
            }
        }
        if (flow_tc && H == 192 && half == 96 && L == 4 && c.flow_kernel_size == 5) {
            // ---- fused coupling-layer kernel: one block stream + bias sums + conditioning in its own chunk order
            const int kk = 5;
            std::vector<uint8_t> img(flow_layer_image_bytes());
            // post folded into the skip path (fp32 on the host): Wm_i = W_post W_skip_i  [half x H];  channel reversal of the odd
            // layers applied to the rows of W_post
            auto wpost = [&](int co, int ci) { return h_post_w[(size_t)(odd ? half - 1 - co : co) * H + ci]; };
            std::vector<std::vector<float>> wm(L, std::vector<float>((size_t)half * H));
            for (int i = 0; i < L; ++i)
                for (int co = 0; co < half; ++co)
                    for (int ci = 0; ci < H; ++ci) {
                        double acc = 0.0;
                        for (int s = 0; s < H; ++s)
                            acc += (double)wpost(co, s) * (double)h_rs_w[i][(size_t)((i < L - 1 ? H : 0) + s) * H + ci];
                        wm[i][(size_t)co * H + ci] = (float)acc;
                    }
            flow_layer_pack(
                [&](int co, int ci) { return h_pre_w[(size_t)co * half + (odd ? half - 1 - ci : ci)]; },
                [&](int i, int row, int ci, int tap) { return h_in_w[i][((size_t)row * H + ci) * kk + tap]; },
                [&](int i, int row, int ci) {
                    if (row < H) return i < L - 1 ? h_rs_w[i][(size_t)row * H + ci] : 0.f;
                    return wm[i][(size_t)(row - H) * H + ci];
                }, img.data());
            if ((rc = upload(ctx, img.data(), img.size(), &F.fused_img))) return rc;
            std::vector<float> bg((size_t)L * 2 * H), bh((size_t)L * H), bo(H, 0.f), bm(half);
            for (int i = 0; i < L; ++i)
                for (int col = 0; col < 2 * H; ++col) bg[(size_t)i * 2 * H + col] = h_in_b[i][flow_gate_row(col)];
            for (int ch = 0; ch < H; ++ch) {
                float acc = h_pre_b[ch];
                for (int i = 0; i < L; ++i) {
                    bh[(size_t)i * H + ch] = acc;
                    if (i < L - 1) { acc += h_rs_b[i][ch]; bo[ch] += h_rs_b[i][H + ch]; }
                    else bo[ch] += h_rs_b[i][ch];
                }
            }
            for (int co = 0; co < half; ++co) {
                double acc = h_post_b[odd ? half - 1 - co : co];
                for (int s = 0; s < H; ++s) acc += (double)wpost(co, s) * (double)bo[s];
                bm[co] = (float)acc;
            }
            if ((rc = upload(ctx, bg.data(), bg.size() * sizeof(float), (void**)&F.fb_gate))) return rc;
            if ((rc = upload(ctx, bh.data(), bh.size() * sizeof(float), (void**)&F.fb_h))) return rc;
            if ((rc = upload(ctx, bm.data(), bm.size() * sizeof(float), (void**)&F.fb_m))) return rc;
            std::vector<float> wp(h_cond_w.size()), bpc(h_cond_b.size());
            for (int i = 0; i < L; ++i)
                for (int col = 0; col < 2 * H; ++col) {
                    const int src = 2 * H * i + flow_gate_row(col), dst = 2 * H * i + col;
                    std::memcpy(&wp[(size_t)dst * G], &h_cond_w[(size_t)src * G], sizeof(float) * G);
                    bpc[dst] = h_cond_b[src];
                }
            if ((rc = upload(ctx, wp.data(), wp.size() * sizeof(float), (void**)&F.cond_w_perm2))) return rc;
            if ((rc = upload(ctx, bpc.data(), bpc.size() * sizeof(float), (void**)&F.cond_b_perm2))) return rc;
            if ((rc = make_conv(ctx, wp, bpc, 2 * H * L, G, 1, false, false, F.cond2))) return rc;
        }
    }

    if (!ctx->flow.empty() && ctx->flow[0].fused_img) {
        // conditioning of all coupling layers as ONE GEMV per call: rows [fl][2H*L] in each layer's chunk order
        const size_t per = (size_t)2 * H * L;
        std::vector<float> wa(ctx->flow.size() * per * G), ba(ctx->flow.size() * per);
        for (size_t fl = 0; fl < ctx->flow.size(); ++fl) {
            CU(cudaMemcpy(&wa[fl * per * G], ctx->flow[fl].cond_w_perm2, per * G * sizeof(float), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&ba[fl * per], ctx->flow[fl].cond_b_perm2, per * sizeof(float), cudaMemcpyDeviceToHost));
        }
        if ((rc = upload(ctx, wa.data(), wa.size() * sizeof(float), (void**)&ctx->cond_all_w))) return rc;
        if ((rc = upload(ctx, ba.data(), ba.size() * sizeof(float), (void**)&ctx->cond_all_b))) return rc;
    }

    // ---- generator
    const int U = c.upsample_initial_channel;
    const int Cpre = melv ? c.num_mels : C;             // conv_pre input channels: mel bins or the latent z
    if ((rc = folded(ctx, m, DP + "conv_pre", {U, Cpre, 7}, w))) return rc;
    if ((rc = get_tensor(ctx, m, DP + "conv_pre.bias", {U}, b))) return rc;
    if ((rc = make_conv(ctx, w.v, b.v, U, Cpre, 7, false, false, ctx->conv_pre))) return rc;
    ctx->gen_tc_ok = (Cpre == 192 || Cpre == 128 || Cpre == 256) && (U % 256 == 0);
    for (int i = 0; i < c.n_upsamples; ++i) {
        const int ci_ = U >> i, n_ = (U >> (i + 1)) * c.upsample_rates[i];
        if (!(ci_ == 512 || ci_ == 256 || ci_ == 128 || ci_ == 64 || ci_ == 32) || (n_ % 32)) ctx->gen_tc_ok = false;
    }
    if (ctx->gen_tc_ok) {
        const std::vector<float> wv = w.v, bv = b.v;
        if ((rc = make_convn(ctx, Cpre, Cpre, U, 256 / convn_mb(Cpre), 7, 3, [&](int col, int ci, int tap) { return wv[((size_t)col * Cpre + ci) * 7 + tap]; },
                             [&](int col) { return bv[col]; }, ctx->conv_pre_tc))) return rc;
    }
    if (!melv) {
        if ((rc = get_tensor(ctx, m, "dec.cond.weight", {U, G, 1}, w))) return rc;
        if ((rc = get_tensor(ctx, m, "dec.cond.bias", {U}, b))) return rc;
        if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&ctx->dcond_w_nat))) return rc;
        if ((rc = make_conv(ctx, w.v, b.v, U, G, 1, false, false, ctx->dcond))) return rc;
        ctx->dcond_b = ctx->dcond.b;
    }
    ctx->stages.assign(c.n_upsamples, Stage());
    for (int i = 0; i < c.n_upsamples; ++i) {
        Stage& S = ctx->stages[i];
        S.Cin = U >> i; S.Cout = U >> (i + 1); S.s = c.upsample_rates[i]; S.k = c.upsample_kernel_sizes[i];
        S.p = (S.k - S.s + 1) / 2;
        const std::string p = DP + "ups." + std::to_string(i);
        if ((rc = folded(ctx, m, p, {S.Cin, S.Cout, S.k}, w))) return rc;    // ConvTranspose1d: [Cin][Cout][k]
        if ((rc = get_tensor(ctx, m, p + ".bias", {S.Cout}, b))) return rc;
        std::vector<float> pk((size_t)S.s * S.Cin * 2 * S.Cout);
        for (int ph = 0; ph < S.s; ++ph)
            for (int ci = 0; ci < S.Cin; ++ci)
                for (int co = 0; co < S.Cout; ++co) {
                    const size_t base = ((size_t)ph * S.Cin + ci) * 2 * S.Cout;
                    pk[base + 0 * S.Cout + co] = w.v[((size_t)ci * S.Cout + co) * S.k + ph + S.s];  // tap 0 <-> x[i0-1]
                    pk[base + 1 * S.Cout + co] = w.v[((size_t)ci * S.Cout + co) * S.k + ph];        // tap 1 <-> x[i0]
                }
        if ((rc = upload(ctx, pk.data(), pk.size() * sizeof(float), (void**)&S.up_w))) return rc;
        if ((rc = upload(ctx, b.v.data(), b.v.size() * sizeof(float), (void**)&S.up_b))) return rc;
        const std::vector<float> upw_host = w.v, upb_host = b.v;
        // noise conv
        int stride = 1;
        for (int q = i + 1; q < c.n_upsamples; ++q) stride *= c.upsample_rates[q];
        const std::string np_ = DP + "noise_convs." + std::to_string(i);
        if (i + 1 < c.n_upsamples) { S.noise_K = 2 * stride; S.noise_s = stride; S.noise_p = (stride + 1) / 2; }
        else { S.noise_K = 1; S.noise_s = 1; S.noise_p = 0; }
        if ((rc = get_tensor(ctx, m, np_ + ".weight", {S.Cout, 1, S.noise_K}, w))) return rc;
        if ((rc = get_tensor(ctx, m, np_ + ".bias", {S.Cout}, b))) return rc;
        if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&S.noise_w))) return rc;
        if ((rc = upload(ctx, b.v.data(), b.v.size() * sizeof(float), (void**)&S.noise_b))) return rc;
        if (ctx->gen_tc_ok) {
            // polyphase GEMM columns: col = co*s + phase; tap 0 multiplies x[i-1] (kernel index phase+s), tap 1 x[i] (index phase).
            // Narrow stages also absorb noise_convs[i]: window u = phase*s' + kk of har[i*s*s' - p*s' - p_n + u] (<= 16 wide).
            const std::vector<float> nwv = w.v, nbv = b.v;
            const int s_ = S.s, Co = S.Cout, kk = S.k;
            const int ntot = Co * s_;
            int nc = 256 / (c.snake ? convn_mb(S.Cin) : convn_ups_mb(S.Cin));
            if (nc > ntot) nc = ntot;
            const int sp = S.noise_s, Kn = S.noise_K, pn_ = S.noise_p;
            const int nwin = (s_ - 1) * sp + Kn;                   // excitation window of one output row
            const int noise_kind = nwin <= 16 ? 1 : (nwin <= 80 ? 2 : 0);
            const bool fuse_noise = noise_kind != 0 && S.Cin <= 256;
            std::function<float(int, int)> ncol = [&](int col, int u) {
                const int co = col / s_, ph = col % s_;
                const int q = u - ph * sp;
                return (q >= 0 && q < Kn) ? nwv[(size_t)co * Kn + q] : 0.f;
            };
            if ((rc = make_convn(ctx, S.Cin, S.Cin, ntot, nc, 2, 1,
                                 [&](int col, int ci, int tap) { const int co = col / s_, ph = col % s_; return upw_host[((size_t)ci * Co + co) * kk + (tap == 0 ? ph + s_ : ph)]; },
                                 [&](int col) { return upb_host[col / s_] + (fuse_noise ? nbv[col / s_] : 0.f); }, S.up_tc,
                                 fuse_noise ? &ncol : nullptr, noise_kind))) return rc;
            S.up_tc.noise_stride = s_ * sp;
            S.up_tc.noise_w0 = -S.p * sp - pn_;
            if (!fuse_noise && Kn == 2 * sp && sp == 64 && (Co % 32) == 0) {
                // y[co,t] = sum_kk wn[co,kk] har[t*64 - 32 + kk]: rows H[t] = har[64t-32 .. +64) are a strided view of the
                // excitation; the 128-tap filter is two 64-wide taps on consecutive rows -> convn with Cin = 64, k = 2
                if ((rc = make_convn(ctx, 64, 64, Co, Co <= 128 ? Co : 128, 2, 0,
                                     [&](int col, int ci, int tap) { return nwv[(size_t)col * Kn + tap * 64 + ci]; },
                                     [&](int col) { return nbv[col]; }, S.noise_tc))) return rc;
            }
        }
        S.c1.assign(9, ConvW());
        S.c2.assign(9, ConvW());
        const bool snake_tc = c.snake && ctx->gen_tc_ok && (S.Cout == 256 || S.Cout == 128 || S.Cout == 64 || S.Cout == 32 || S.Cout == 16);
        if (snake_tc) { S.c1n.assign(9, ConvNW()); S.c2n.assign(9, ConvNW()); }
        for (int j = 0; j < 3; ++j) {
            const int k = c.resblock_kernel_sizes[j];
            const std::string r = DP + "resblocks." + std::to_string(i * 3 + j) + ".";
            for (int d = 0; d < 3; ++d) {
                if ((rc = folded(ctx, m, r + "convs1." + std::to_string(d), {S.Cout, S.Cout, k}, w))) return rc;
                if ((rc = get_tensor(ctx, m, r + "convs1." + std::to_string(d) + ".bias", {S.Cout}, b))) return rc;
                if ((rc = make_conv(ctx, w.v, b.v, S.Cout, S.Cout, k, false, false, S.c1[j * 3 + d]))) return rc;
                auto as_convn = [&](ConvNW& dstw) {
                    const std::vector<float> wv = w.v, bv = b.v;
                    const int Cc = S.Cout;
                    return make_convn(ctx, Cc, Cc, Cc, Cc, k, 0 /* pad_left is set per launch (dilation) */,
                                      [&](int col, int ci, int tap) { return wv[((size_t)col * Cc + ci) * k + tap]; },
                                      [&](int col) { return bv[col]; }, dstw);
                };
                if (snake_tc && (rc = as_convn(S.c1n[j * 3 + d]))) return rc;
                if ((rc = folded(ctx, m, r + "convs2." + std::to_string(d), {S.Cout, S.Cout, k}, w))) return rc;
                if ((rc = get_tensor(ctx, m, r + "convs2." + std::to_string(d) + ".bias", {S.Cout}, b))) return rc;
                if ((rc = make_conv(ctx, w.v, b.v, S.Cout, S.Cout, k, false, false, S.c2[j * 3 + d]))) return rc;
                if (snake_tc && (rc = as_convn(S.c2n[j * 3 + d]))) return rc;
            }
        }
    }
    if (c.snake) {
        auto load_snake = [&](const std::string& prefix, int Cc, SnakeP& out) -> int {
            HostT al, be;
            int r2;
            if ((r2 = get_tensor(ctx, m, prefix + "act.alpha", {Cc}, al))) return r2;
            if ((r2 = get_tensor(ctx, m, prefix + "act.beta", {Cc}, be))) return r2;
            std::vector<float> ea(Cc), ib(Cc);
            for (int q = 0; q < Cc; ++q) { ea[q] = std::exp(al.v[q]); ib[q] = 1.0f / (std::exp(be.v[q]) + 0.000000001f); }
            out.C = Cc;
            if ((r2 = upload(ctx, ea.data(), Cc * sizeof(float), (void**)&out.ealpha))) return r2;
            return upload(ctx, ib.data(), Cc * sizeof(float), (void**)&out.inv_beta);
        };
        for (int i = 0; i < c.n_upsamples; ++i) {
            Stage& S = ctx->stages[i];
            if ((rc = load_snake("dec.snakes." + std::to_string(i) + ".", S.Cin, S.snake_in))) return rc;
            S.acts.assign(18, SnakeP());
            for (int j = 0; j < 3; ++j)
                for (int a2 = 0; a2 < 6; ++a2)
                    if ((rc = load_snake("dec.resblocks." + std::to_string(i * 3 + j) + ".activations." + std::to_string(a2) + ".", S.Cout, S.acts[j * 6 + a2]))) return rc;
        }
        if ((rc = load_snake("dec.snake_post.", U >> c.n_upsamples, ctx->snake_post))) return rc;
        if ((rc = get_tensor(ctx, m, "dec.snake_post.upsample.filter", {1, 1, 12}, w))) return rc;
        if ((rc = upload(ctx, w.v.data(), 12 * sizeof(float), (void**)&ctx->snake_filt))) return rc;
    }
    ctx->post_C = U >> c.n_upsamples;
    if ((rc = folded(ctx, m, DP + "conv_post", {1, ctx->post_C, 7}, w))) return rc;
    if ((rc = get_tensor(ctx, m, DP + "conv_post.bias", {1}, b))) return rc;
    ctx->post_b = b.v[0];
    if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&ctx->post_w))) return rc;
    if ((rc = get_tensor(ctx, m, DP + "m_source.l_linear.weight", {1, c.n_harmonics}, w))) return rc;
    if ((rc = get_tensor(ctx, m, DP + "m_source.l_linear.bias", {1}, b))) return rc;
    ctx->lin_b = b.v[0];
    if ((rc = upload(ctx, w.v.data(), w.v.size() * sizeof(float), (void**)&ctx->lin_w))) return rc;
    if (!melv && c.enc_layers > 0 && m.count("pre.weight") && m.count("enc_p.proj.weight")) {
        if ((rc = load_prefix(ctx, m))) return rc;
    }
    ctx->loaded = true;
    return SVB_OK;
}

size_t svb_workspace_bytes(const svb_ctx* ctx, int B, int T) {
    if (!ctx || !ctx->loaded || B < 1 || T < 1) return 0;
    return plan_ws(ctx->cfg, B, T, T).total;   // sized for time-varying g (the larger case)
}

#define PRECHECK()                                                                     \
    if (!ctx) return SVB_ERR_INVALID_ARG;                                              \
    if (!ctx->loaded) return fail(ctx, SVB_ERR_NOT_LOADED, "svb_load_weights first");  \
    if (B < 1 || T < 1) return fail(ctx, SVB_ERR_INVALID_ARG, "B and T must be >= 1"); \
    DevGuard dg__(ctx->device);                                                        \
    if (!dg__.ok) return fail(ctx, SVB_ERR_CUDA, "cudaSetDevice failed");

int svb_flow_reverse(svb_ctx* ctx, const float* z_p, const float* g, int gT, const int32_t* lengths,
                     float* z_out, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    PRECHECK();
    if (ctx->cfg.num_mels > 0) return fail(ctx, SVB_ERR_UNSUPPORTED, "mel vocoder contexts have no flow");
    if (!z_p || !g || !z_out || (gT != 1 && gT != T)) return fail(ctx, SVB_ERR_INVALID_ARG, "bad flow arguments");
    WsPlan pl = plan_ws(ctx->cfg, B, T, gT);
    char* base;
    int rc = ensure_ws(ctx, pl.total, ws, ws_bytes, &base);
    if (rc) return rc;
    return run_flow(ctx, z_p, g, gT, lengths, z_out, B, T, base, pl, (cudaStream_t)stream);
}

int svb_nsf_source(svb_ctx* ctx, const float* f0, const float* rand_ini, const float* noise,
                   float* har, int B, int T, void* stream) {
    PRECHECK();
    if (!f0 || !rand_ini || !har) return fail(ctx, SVB_ERR_INVALID_ARG, "bad source arguments");
    WsPlan pl = plan_ws(ctx->cfg, B, T, 1);
    char* base;
    int rc = ensure_ws(ctx, pl.total, nullptr, 0, &base);
    if (rc) return rc;
    launch_nsf_source(f0, rand_ini, noise, ctx->lin_w, ctx->lin_b, reinterpret_cast<double*>(base + pl.off_phase), har,
                      B, T, ctx->hop, ctx->cfg.n_harmonics, (float)ctx->cfg.sampling_rate, ctx->cfg.num_mels > 0 ? 1 : 0, (cudaStream_t)stream,
                      ctx->opt_philox, ctx->philox_seed);
    return check_launch(ctx, "nsf_source");
}

int svb_generator(svb_ctx* ctx, const float* z, const float* g, int gT, const float* har,
                  float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    PRECHECK();
    if (!z || (!g && ctx->cfg.num_mels == 0) || !har || !wav || (gT != 1 && gT != T)) return fail(ctx, SVB_ERR_INVALID_ARG, "bad generator arguments");
    WsPlan pl = plan_ws(ctx->cfg, B, T, gT);
    char* base;
    int rc = ensure_ws(ctx, pl.total, ws, ws_bytes, &base);
    if (rc) return rc;
    return run_generator(ctx, z, g, gT, har, wav, B, T, base, pl, (cudaStream_t)stream);
}

int svb_vocoder(svb_ctx* ctx, const float* mel, const float* f0, const float* rand_ini, const float* noise,
                float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    PRECHECK();
    if (ctx->cfg.num_mels <= 0) return fail(ctx, SVB_ERR_UNSUPPORTED, "context was not loaded as a mel vocoder (num_mels == 0)");
    if (!mel || !f0 || !rand_ini || !wav) return fail(ctx, SVB_ERR_INVALID_ARG, "bad vocoder arguments");
    WsPlan pl = plan_ws(ctx->cfg, B, T, 1);
    char* base;
    int rc = ensure_ws(ctx, pl.total, ws, ws_bytes, &base);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* har = reinterpret_cast<float*>(base + pl.off_har);
    launch_nsf_source(f0, rand_ini, noise, ctx->lin_w, ctx->lin_b, reinterpret_cast<double*>(base + pl.off_phase), har,
                      B, T, ctx->hop, ctx->cfg.n_harmonics, (float)ctx->cfg.sampling_rate, 1, st, ctx->opt_philox, ctx->philox_seed);
    if ((rc = dbg_keep(ctx, "har", har, (size_t)B * T * ctx->hop, st))) return rc;
    return run_generator(ctx, mel, nullptr, 1, har, wav, B, T, base, pl, st);
}

int svb_infer_tail(svb_ctx* ctx, const float* z_p, const float* g, int gT, const int32_t* lengths,
                   const float* f0, const float* rand_ini, const float* noise,
                   float* wav, int B, int T, void* ws, size_t ws_bytes, void* stream) {
    PRECHECK();
    if (ctx->cfg.num_mels > 0) return fail(ctx, SVB_ERR_UNSUPPORTED, "mel vocoder contexts: use svb_vocoder");
    if (!z_p || !g || !f0 || !rand_ini || !wav || (gT != 1 && gT != T)) return fail(ctx, SVB_ERR_INVALID_ARG, "bad tail arguments");
    WsPlan pl = plan_ws(ctx->cfg, B, T, gT);
    char* base;
    int rc = ensure_ws(ctx, pl.total, ws, ws_bytes, &base);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* z = reinterpret_cast<float*>(base + pl.off_z);
    float* har = reinterpret_cast<float*>(base + pl.off_har);
    {
        ProfScope ps(ctx, "flow", st, 0, 0);
        // per frame: 4 coupling layers x (pre 96x192 + 4 x (k5 192x384 + 192x384|192) + post 192x96) MACs (14.156 MFLOP/frame)
        ProfScope ps2(ctx, ctx->precision == SVB_PREC_TC ? "flow_tc" : "flow_f32", st, 14.156e6 * (double)B * T, 4.0 * 1152.0 * (double)B * T);
        if ((rc = run_flow(ctx, z_p, g, gT, lengths, z, B, T, base, pl, st))) return rc;
    }
    if ((rc = dbg_keep(ctx, "z", z, (size_t)B * ctx->cfg.inter_channels * T, st))) return rc;
    {
        const double N = (double)T * ctx->hop;
        ProfScope ps(ctx, "nsf_source", st, 0, (double)B * (T * 4.0 + N * 4.0 + (noise ? N * 4.0 * ctx->cfg.n_harmonics : 0.0)));
        launch_nsf_source(f0, rand_ini, noise, ctx->lin_w, ctx->lin_b, reinterpret_cast<double*>(base + pl.off_phase), har,
                          B, T, ctx->hop, ctx->cfg.n_harmonics, (float)ctx->cfg.sampling_rate, ctx->cfg.num_mels > 0 ? 1 : 0, st,
                          ctx->opt_philox, ctx->philox_seed);
    }
    if ((rc = dbg_keep(ctx, "har", har, (size_t)B * T * ctx->hop, st))) return rc;
    ProfScope ps(ctx, "generator", st, 0, 0);
    return run_generator(ctx, z, g, gT, har, wav, B, T, base, pl, st);
}

int svb_pre_conv(svb_ctx* ctx, const float* c, float* x, int B, int T, void* stream) {
    PRECHECK();
    if (!ctx->prefix.ok) return fail(ctx, SVB_ERR_UNSUPPORTED, "the prior encoder was not loaded (enc_layers = 0, missing pre./enc_p. tensors or unsupported shapes)");
    if (!c || !x) return fail(ctx, SVB_ERR_INVALID_ARG, "bad pre_conv arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const Prefix& P = ctx->prefix;
    int rc;
    if ((rc = prefix_conv(ctx, P.pre_a, c, P.ssl, 0, x, nullptr, 0.f, 0, B, T, st, &P.pre_b, 384))) return rc;     // K = 768 as 2 x 384
    return check_launch(ctx, "pre_conv");
}

int svb_enc_p(svb_ctx* ctx, const float* x_in, const float* z_noise, float noice_scale, float* z_p, float* m_p, float* logs_p,
              int B, int T, void* stream) {
    PRECHECK();
    if (!ctx->prefix.ok) return fail(ctx, SVB_ERR_UNSUPPORTED, "the prior encoder was not loaded (enc_layers = 0, missing pre./enc_p. tensors or unsupported shapes)");
    if (!x_in || !z_noise || !z_p) return fail(ctx, SVB_ERR_INVALID_ARG, "bad enc_p arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const Prefix& P = ctx->prefix;
    const size_t BT = (size_t)B * T, f = sizeof(float);
    // fp16 tile images of q, k (32 KB per 128-row tile and head) and V^T (24 KB) for the attention kernel
    const int tiles = (T + 127) / 128;
    const size_t n_img = (size_t)B * P.heads * tiles;
    const size_t img_bytes = align_up(n_img * 32768, 1024) * 2 + align_up(n_img * 24576, 1024);
    const size_t need = img_bytes + 4 * align_up(BT * P.H * f, 256) + align_up(BT * P.F * f, 256) + align_up(BT * P.out2 * f, 256);
    if (ctx->ws_prefix.bytes < need) {
        if (ctx->ws_prefix.p) { CU(cudaStreamSynchronize(st)); CU(cudaFree(ctx->ws_prefix.p)); }
        ctx->ws_prefix.p = nullptr; ctx->ws_prefix.bytes = 0;
        CU(cudaMalloc(&ctx->ws_prefix.p, need));
        ctx->ws_prefix.bytes = need;
        // rows of the last tile beyond T are never written by the projection GEMM: they must be finite (p = 0 multiplies them)
        CU(cudaMemsetAsync(ctx->ws_prefix.p, 0, need, st));
    }
    char* wp = static_cast<char*>(ctx->ws_prefix.p);
    char* q_img = wp; char* k_img = q_img + align_up(n_img * 32768, 1024); char* v_img = k_img + align_up(n_img * 32768, 1024);
    wp += img_bytes;
    auto take = [&](size_t bytes) { char* r = wp; wp += align_up(bytes, 256); return reinterpret_cast<float*>(r); };
    float* A = take(BT * P.H * f);
    float* Y = take(BT * P.H * f);
    float* X1 = take(BT * P.H * f);
    float* X = take(BT * P.H * f);
    float* Hd = take(BT * P.F * f);
    float* St = take(BT * P.out2 * f);
    int rc;
    const float* cur = x_in;
    ProfScope ps(ctx, "enc_p", st, 0, 0);
    for (size_t l = 0; l < P.layers.size(); ++l) {
        const EncLayer& E = P.layers[l];
        {   // q | k | v projection; the epilogue writes the attention kernel's fp16 operand tiles directly
            ConvNTC a;
            const ConvNW& W = E.qkv;
            a.x = cur; a.x_ctot = P.H; a.cin_real = W.cin_real; a.cinp = W.cinp; a.Tin = T;
            a.w = W.img; a.bias = W.bias; a.acc_scale = W.acc_scale; a.k = 1; a.pad_left = 0;
            a.n_rows = T; a.N_total = W.N_total; a.NC = W.NC; a.Ty = T; a.B = B;
            // one wave: 7 x 8 x 3 = 168 one-CTA-per-SM blocks would need two; with two chunks per CTA the grid is 112
            a.chunks_per_cta = ((T + 127) / 128) * B * 3 > 148 ? 2 : 1;
            a.mode = 3; a.att_q = q_img; a.att_k = k_img; a.att_v = v_img; a.att_heads = P.heads; a.att_tiles = tiles;
            ProfScope ps(ctx, "enc_gemm", st, 2.0 * P.H * 3.0 * P.H * (double)T * B, 0);
            if ((rc = launch_convn_tc(a, st))) return fail(ctx, rc, "qkv projection launch failed");
        }
        AttnTC at;
        at.q_img = q_img; at.k_img = k_img; at.v_img = v_img;
        at.ek = E.ek; at.ev = E.ev; at.out = A; at.out_ctot = P.H; at.B = B; at.T = T; at.heads = P.heads; at.dk = P.H / P.heads; at.window = P.window;
        {
            // 2 x (S = QK^T twice, P V once): 3 x 2 x T x T x dk per (batch, head)
            ProfScope pa(ctx, "enc_attn", st, 3.0 * 2.0 * (double)T * T * at.dk * P.heads * B, 4.0 * P.H * (double)T * B * sizeof(float));
            if ((rc = launch_attn_rel_tc(at, st))) return fail(ctx, rc, "attention kernel launch failed");
        }
        if ((rc = prefix_conv(ctx, E.o, A, P.H, 0, Y, cur, 0.f, 0, B, T, st))) return rc;            // x + conv_o(attn)
        launch_ln_cm(Y, E.g1, E.b1, 1e-5f, X1, B, P.H, T, st);
        if ((rc = prefix_conv(ctx, E.ffn1, X1, P.H, 0, Hd, nullptr, 0.f, 1, B, T, st))) return rc;   // relu(conv_1)
        if ((rc = prefix_conv(ctx, E.ffn2a, Hd, P.F, 0, Y, X1, 0.f, 0, B, T, st, &E.ffn2b, 384))) return rc;   // x1 + conv_2, K = 768 as 2 x 384
        launch_ln_cm(Y, E.g2, E.b2, 1e-5f, X, B, P.H, T, st);
        cur = X;
    }
    if ((rc = prefix_conv(ctx, P.proj, cur, P.H, 0, St, nullptr, 0.f, 0, B, T, st))) return rc;
    launch_prior_sample(St, z_noise, noice_scale, z_p, m_p, logs_p, B, P.out2 / 2, T, st);
    return check_launch(ctx, "enc_p");
}

int svb_infer_tail_host(svb_ctx* ctx, const float* z_p, const float* g, int gT,
                        const float* f0, const float* rand_ini, const float* noise,
                        float* wav, int B, int T) {
    PRECHECK();
    if (!z_p || !g || !f0 || !rand_ini || !wav || (gT != 1 && gT != T)) return fail(ctx, SVB_ERR_INVALID_ARG, "bad tail arguments");
    const svb_model_cfg& c = ctx->cfg;
    const size_t N = (size_t)T * ctx->hop;
    const size_t n_zp = (size_t)B * c.inter_channels * T, n_g = (size_t)B * c.gin_channels * gT, n_f0 = (size_t)B * T,
                 n_ri = (size_t)B * c.n_harmonics, n_nz = noise ? (size_t)B * N * c.n_harmonics : 0, n_wav = (size_t)B * N;
    // every sub-buffer starts on a 256-byte boundary: the kernels pick their 128-bit code paths by pointer alignment, and a
    // host call must run exactly the kernels a device call with torch-allocated tensors runs (bit-identical results)
    auto al = [](size_t n) { return (n + 63) & ~(size_t)63; };
    const size_t need = (al(n_zp) + al(n_g) + al(n_f0) + al(n_ri) + al(n_nz) + al(n_wav)) * sizeof(float);
    if (ctx->host_io.bytes < need) {
        if (ctx->host_io.p) CU(cudaFree(ctx->host_io.p));
        ctx->host_io.p = nullptr; ctx->host_io.bytes = 0;
        CU(cudaMalloc(&ctx->host_io.p, need));
        ctx->host_io.bytes = need;
    }
    float* d = static_cast<float*>(ctx->host_io.p);
    float* d_zp = d; d += al(n_zp);
    float* d_g = d; d += al(n_g);
    float* d_f0 = d; d += al(n_f0);
    float* d_ri = d; d += al(n_ri);
    float* d_nz = d; d += al(n_nz);
    float* d_wav = d;
    cudaStream_t st = 0;
    CU(cudaMemcpyAsync(d_zp, z_p, n_zp * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_g, g, n_g * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_f0, f0, n_f0 * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ri, rand_ini, n_ri * sizeof(float), cudaMemcpyHostToDevice, st));
    if (noise) CU(cudaMemcpyAsync(d_nz, noise, n_nz * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = svb_infer_tail(ctx, d_zp, d_g, gT, nullptr, d_f0, d_ri, noise ? d_nz : nullptr, d_wav, B, T, nullptr, 0, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(wav, d_wav, n_wav * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SVB_OK;
}

}  // extern "C"
