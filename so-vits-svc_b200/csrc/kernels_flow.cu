// One ResidualCouplingLayer (reverse) per launch on tcgen05 (sm_100a) - north_star kernel (i):
//     x0, x1 = split(y);  h = pre(x0) * mask;  out = 0
//     for i in 0..3:  a = in_layers[i](h) + cond_i(g);  acts = tanh(a[:H]) * sigmoid(a[H:])
//                     rs = res_skip[i](acts);  h = (h + rs[:H]) * mask (i < 3);  out += rs[H:] (all of rs for i = 3)
//     m = post(out * mask) * mask;  x1 = (x1 - m) * mask                      (mean_only: logs = 0)
// (modules/modules.py:288-307 ResidualCouplingLayer, :110-138 WN, modules/commons.py:129-136 gate; Flip folded into the
// channel order of pre / post at pack time)  H = 192 hidden channels, 96-channel halves, k = 5, 4 WN layers.
//
// One CTA owns 128 consecutive frames of one utterance for the WHOLE layer: 112 output frames + 8 halo frames per side
// (4 WN layers x (k-1)/2 = 2 frames each).  Nothing but x0 (read), x1 (read-modify-write) and the weights touches global
// memory; the former schedule was 10 launches with h / acts / out (5.3 MB each at config 2) round-tripping between them.
//
// `post` is folded into the skip path: m = post(sum_i skip_i) = sum_i (W_post W_skip_i) acts_i + (W_post b_skip + b_post),
// so the kernel accumulates m (96 columns) directly instead of `out` (192) and needs no post GEMM; the products
// W_post W_skip_i are formed in fp32 on the host at load time (mask = row validity commutes with the 1x1 maps).
//
//   TMEM (480 of 512 columns x 128 lanes, lane = frame of the tile):
//     [  0,192)  h    fp32 residual stream of the WN stack.  Written by the pre GEMM, then the res halves of res_skip are
//                     ACCUMULATED into it by the MMAs themselves (no epilogue pass for the residual add).
//     [192,288)  m    fp32 accumulator of post(skip), zeroed once, accumulated by the res_skip MMAs of all four layers.
//     [288,384) [384,480)  two buffers for one 96-column chunk of the in_layer pre-activations (48 tanh + 48 sigmoid
//                     channels): the gate epilogue of chunk c runs while the MMAs of chunk c+1 fill the other buffer.
//   Shared memory: H16 = fp16 operand copy of h with zero pad rows above / below (3 K-panels x 136 rows x 128 B;
//     a k5 tap is a row offset of the A descriptor); ACTS = fp16 gate output, one 128-row panel per chunk (48 channels in
//     128-byte rows); an 8-stage ring of 12 KB weight blocks ([96 output rows][64 input channels], host-swizzled) fed by
//     1-D bulk TMA copies.  The x0 operand tile aliases ACTS.
//   Weights: ONE linear stream of 292 blocks per coupling layer in exactly the order the MMA warp consumes them:
//     pre (4) | per WN layer: in(0) in(1) rs(0) in(2) rs(1) in(3) rs(2) rs(3)   with in(c) = 5 taps x 3 panels (15 blocks)
//     and rs(c) = 3 row blocks of the combined [res 192 | post.skip 96] matrix against the 48 channels of chunk c.
//   Biases: res / pre biases are never added in TMEM; the running sums are added when h is converted to fp16
//     (bias_h[i] = b_pre + sum_{j<i} b_res_j), the folded skip bias when m is consumed.
//
// Warp roles: warps 0-7 stage / epilogue (warp w <-> TMEM lanes 32 (w%4), column half w/4), warp 8 owns TMEM and issues
// every MMA from one elected lane, warp 9 streams the weight blocks.  All hand-offs are mbarriers in one linear dependency
// chain, so the protocol cannot deadlock.
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/sovits_b200.h"

#include <cstring>
#include <vector>

namespace svb {

using namespace tc;

namespace {

constexpr int FL_THREADS = 320;
constexpr int FL_NWORK = 256;
constexpr int FL_H = 192, FL_HALF = 96, FL_L = 4, FL_K = 5;
constexpr int FL_HALO = FL_L * (FL_K - 1) / 2;       // 8 frames per side
constexpr int FL_TOUT = 128 - 2 * FL_HALO;           // 112 output frames per tile
constexpr int FL_PAD = (FL_K - 1) / 2;               // zero rows above / below the H16 tile
constexpr int FL_HROWS = 136;                        // 128 + FL_PAD rows above + 6 below (a multiple of 8: every K-panel must
                                                     // start on a 1024-byte boundary, the swizzle is a function of the address)
constexpr int FL_RB = 128;                           // operand row bytes (64 fp16 channels, SWIZZLE_128B)
constexpr int FL_HPANEL = FL_HROWS * FL_RB;          // 17408
constexpr int FL_APANEL = 128 * FL_RB;               // 16384
constexpr int FL_NB = 96;                            // rows of a weight block = columns of every MMA
constexpr int FL_BLOCK = FL_NB * FL_RB;              // one weight block: 96 rows x 64 channels = 12 KB
constexpr int FL_NSTAGE = 8;                         // 96 KB of weights in flight
constexpr int FL_NCHUNK = 4, FL_CH = 48;             // gate chunks per WN layer, channels per chunk
constexpr int FL_NBLK_PRE = 4, FL_NBLK_IN = 15, FL_NBLK_RS = 3;
constexpr int FL_NBLK = FL_NBLK_PRE + FL_L * FL_NCHUNK * (FL_NBLK_IN + FL_NBLK_RS);          // 292
constexpr int COL_H = 0, COL_M = 192, COL_X = 288;   // xin buffer b at COL_X + 96 b

// shared memory map (bytes from the 1024-aligned base)
constexpr uint32_t OFF_H16 = 0;                                          // 3 panels x 136 rows
constexpr uint32_t OFF_ACTS = ((3 * FL_HPANEL + 1023) / 1024) * 1024;    // 52224
static_assert(FL_HPANEL % 1024 == 0 && FL_APANEL % 1024 == 0, "operand panels must be 1024-byte aligned");
constexpr uint32_t OFF_RING = OFF_ACTS + FL_NCHUNK * FL_APANEL;          // + 65536
constexpr uint32_t OFF_BAR = OFF_RING + FL_NSTAGE * FL_BLOCK;            // + 65536
constexpr uint32_t OFF_BIAS = OFF_BAR + 256;
// floats: gate bias [L][2H] | bias_h [L][H] | bias_m [HALF]
constexpr uint32_t N_BIAS = FL_L * 2 * FL_H + FL_L * FL_H + FL_HALF;
static_assert(1024 + OFF_BIAS + N_BIAS * 4 <= 227 * 1024, "flow kernel shared memory");
constexpr size_t FL_SMEM = 1024 + OFF_BIAS + N_BIAS * 4;

struct FlowParams {
    float* y; int y_ctot, in_c0, out_c0;
    const uint8_t* w;
    const float* bias_gate; const float* bias_h; const float* bias_m;
    const float* gcond; int gcond_bstride;   // per-utterance conditioning (chunk-permuted) [L*2H] at gcond + b*stride, or null
    const float* gcond_t;    // [B][L*2H][T] time-varying conditioning (speaker mix), or null
    const int32_t* lengths;
    int T;
};

// tanh(a) * sigmoid(b) = (1 - 2/(1 + e^{2a})) / (1 + e^{-b}) through MUFU.EX2 / MUFU.RCP (relative error ~1e-6; the result is
// rounded to fp16 right after)
__device__ __forceinline__ float gate_act(float a, float b) {
    const float ea = __expf(2.f * a);
    const float eb = __expf(-b);
    const float th = 1.f - __fdividef(2.f, 1.f + ea);
    return __fdividef(th, 1.f + eb);
}

__global__ void __launch_bounds__(FL_THREADS, 1) flow_layer_kernel(const FlowParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t h16_base = base + OFF_H16, acts_base = base + OFF_ACTS, ring_base = base + OFF_RING;
    const uint32_t bar_base = base + OFF_BAR;
    static_assert(FL_NSTAGE <= 8, "barrier block holds 8 ring stages");
    const uint32_t bar_full = bar_base;                      // [8]
    const uint32_t bar_empty = bar_base + 64;                // [8]
    const uint32_t bar_a0 = bar_base + 128;                  // x0 operand tile staged          (256 arrivals)
    const uint32_t bar_hfin = bar_base + 136;                // h / out accumulators final      (tcgen05.commit)
    const uint32_t bar_h = bar_base + 144;                   // H16 (or OUT16) operand written  (256 arrivals)
    const uint32_t bar_x = bar_base + 152;                   // [2] xin buffer complete         (tcgen05.commit)
    const uint32_t bar_xfree = bar_base + 168;               // [2] gate epilogue done: ACTS panel written, xin buffer free (256 arrivals)
    const uint32_t tmem_slot = bar_base + 192;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 192);
    float* sb_gate = reinterpret_cast<float*>(sm + OFF_BIAS);
    float* sb_h = sb_gate + FL_L * 2 * FL_H;
    float* sb_m = sb_h + FL_L * FL_H;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FL_TOUT - FL_HALO;           // frame of tile row 0
    const int T = p.T;
    const int len = p.lengths ? min(p.lengths[b], T) : T;

    if (tid == 0) {
        for (int s = 0; s < FL_NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_a0, FL_NWORK);
        mbar_init(bar_hfin, 1);
        mbar_init(bar_h, FL_NWORK);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_x + 8 * s, 1); mbar_init(bar_xfree + 8 * s, FL_NWORK); }
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = tid; i < (int)N_BIAS; i += FL_THREADS) {
        float v;
        if (i < FL_L * 2 * FL_H) {
            v = __ldg(p.bias_gate + i);
            if (p.gcond) v += __ldg(p.gcond + (size_t)b * p.gcond_bstride + i);
        } else if (i < FL_L * 2 * FL_H + FL_L * FL_H) v = __ldg(p.bias_h + (i - FL_L * 2 * FL_H));
        else v = __ldg(p.bias_m + (i - FL_L * 2 * FL_H - FL_L * FL_H));
        sb_gate[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9) {
        // ------------------------------------------------------------ weight producer: 223 blocks, strictly in order
        for (int blk = 0; blk < FL_NBLK; ++blk) {
            const int s = blk % FL_NSTAGE;
            if (blk >= FL_NSTAGE) mbar_wait(bar_empty + 8 * s, ((blk / FL_NSTAGE) - 1) & 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8 * s, FL_BLOCK);
                bulk_g2s(ring_base + s * FL_BLOCK, p.w + (size_t)blk * FL_BLOCK, FL_BLOCK, bar_full + 8 * s);
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer (one elected lane runs the whole program)
        if (elect_one()) {
            uint32_t blk = 0;
            // one ring block against an A operand: `ksteps` K=16 steps, N columns at TMEM column `dcol`
            auto issue_block = [&](uint64_t adesc, uint32_t dcol, int ncols, int ksteps, uint32_t acc_first) {
                const uint32_t s = blk % FL_NSTAGE;
                mbar_wait(bar_full + 8 * s, (blk / FL_NSTAGE) & 1u);
                tc_fence_after();
                const uint64_t bd = make_smem_desc(ring_base + s * FL_BLOCK, FL_RB, 0);
                const uint32_t idesc = make_idesc_f16(128, ncols);
                for (int ks = 0; ks < ksteps; ++ks)
                    umma_f16(tmem_base + dcol, adesc + (uint64_t)((ks * 32) >> 4), bd + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : acc_first);
                umma_commit(bar_empty + 8 * s);
                ++blk;
            };
            const uint64_t a0_desc = make_smem_desc(acts_base, FL_RB, 0);          // x0 tile aliases ACTS: panel p at + p*APANEL
            const uint64_t h_desc = make_smem_desc(h16_base, FL_RB, 0);            // row 0 of H16 = tile row -2
            const uint64_t acts_desc = make_smem_desc(acts_base, FL_RB, 0);
            uint32_t n_h = 0, n_xf[2] = {0u, 0u};
            auto in_chunk = [&](int c) {                       // 5 taps x 3 K-panels of H16 -> xin buffer c & 1
                for (int tap = 0; tap < FL_K; ++tap)
                    for (int pn = 0; pn < 3; ++pn)
                        issue_block(h_desc + (uint64_t)((uint32_t)(pn * FL_HPANEL + tap * FL_RB) >> 4), COL_X + (c & 1) * FL_NB, FL_NB, 4, (tap | pn) ? 1u : 0u);
                umma_commit(bar_x + 8 * (c & 1));
            };
            auto rs_chunk = [&](int c) {                       // [res 192 | post.skip 96] += R[:, chunk c] acts(chunk c)
                mbar_wait(bar_xfree + 8 * (c & 1), n_xf[c & 1] & 1u); ++n_xf[c & 1];    // gate(c) done: ACTS panel c written, buffer free
                tc_fence_after();
                for (int nb = 0; nb < 3; ++nb)
                    issue_block(acts_desc + (uint64_t)((uint32_t)(c * FL_APANEL) >> 4), COL_H + nb * FL_NB, FL_NB, 3, 1u);
            };
            // pre: h = W_pre x0   (K = 96: panel 0 has 4 K-steps, panel 1 two; N = 192 as 2 x 96)
            mbar_wait(bar_a0, 0);
            tc_fence_after();
            for (int pn = 0; pn < 2; ++pn)
                for (int nb = 0; nb < 2; ++nb)
                    issue_block(a0_desc + (uint64_t)((uint32_t)(pn * FL_APANEL) >> 4), COL_H + nb * FL_NB, FL_NB, pn ? 2 : 4, pn ? 1u : 0u);
            umma_commit(bar_hfin);
            for (int i = 0; i < FL_L; ++i) {
                mbar_wait(bar_h, n_h & 1u); ++n_h;                 // H16 of layer i written (and, for i = 0, m zeroed)
                tc_fence_after();
                in_chunk(0);
                in_chunk(1);
                rs_chunk(0);
                in_chunk(2);
                rs_chunk(1);
                in_chunk(3);
                rs_chunk(2);
                rs_chunk(3);
                umma_commit(bar_hfin);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ workers
        const int q = warp & 3, hsel = warp >> 2;
        const int row = 32 * q + lane;                         // tile row == TMEM lane
        const int t = t0 + row;
        const bool tv = (t >= 0) && (t < len);                 // inside the utterance (mask = 1)
        const uint32_t keep = tv ? 0xffffffffu : 0u;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const float* __restrict__ yb = p.y + (size_t)b * p.y_ctot * T;

        // zero the pad rows of H16 (rows [0,2) and [130,136) of every panel; never written again)
        constexpr int NPADROWS = FL_HROWS - 128;
        for (int i = tid; i < 3 * NPADROWS * 8; i += FL_NWORK) {
            const int pn = i / (NPADROWS * 8), rem = i % (NPADROWS * 8);
            const int rr = rem / 8, ch = rem % 8;
            const int r = rr < FL_PAD ? rr : (128 + rr);
            *reinterpret_cast<uint4*>(sm + OFF_H16 + pn * FL_HPANEL + swz_offset(r, ch, FL_RB)) = make_uint4(0, 0, 0, 0);
        }
        // x0 operand tile (aliases ACTS): 96 channels = panel 0 (64) + half of panel 1; this thread: 48 channels of its row
        {
            const float* __restrict__ xt = yb + (size_t)(p.in_c0 + hsel * 48) * T + (tv ? t : 0);
            const uint32_t phase = swz_phase(row, FL_RB);
#pragma unroll
            for (int c0 = 0; c0 < 48; c0 += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = tv ? __ldg(xt + (size_t)(c0 + j) * T) : 0.f;
                const int ch = hsel * 48 + c0;                 // channel of v[0]
                uint8_t* prow = sm + OFF_ACTS + (ch / 64) * FL_APANEL + row * FL_RB;
                store_chunk8(prow, phase, (ch % 64) / 8, v, 0xffffffffu);
                store_chunk8(prow, phase, (ch % 64) / 8 + 1, v + 8, 0xffffffffu);
            }
        }
        fence_proxy_async();
        mbar_arrive(bar_a0);
        // m = 0 (the res_skip MMAs only ever accumulate into it)
        {
            uint32_t z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0u;
#pragma unroll
            for (int cc = 0; cc < 48; cc += 16) tmem_st16(tlane + COL_M + hsel * 48 + cc, z);
            tmem_st_wait();
        }

        uint32_t n_hfin = 0, n_x[2] = {0u, 0u};
        // fp32 TMEM columns [col0 + hsel*96, +96) + bias -> masked fp16 operand rows of H16 (row index shifted by FL_PAD)
        auto to_h16 = [&](uint32_t col0, const float* __restrict__ bias) {
            const uint32_t phase = swz_phase(row + FL_PAD, FL_RB);
#pragma unroll 1
            for (int cc = 0; cc < 96; cc += 32) {
                uint32_t r0[16], r1[16];
                tmem_ld16(tlane + col0 + hsel * 96 + cc, r0);
                tmem_ld16(tlane + col0 + hsel * 96 + cc + 16, r1);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t* rr = g ? r1 : r0;
                    const int ch = hsel * 96 + cc + 16 * g;
                    float v[16];
#pragma unroll
                    for (int j4 = 0; j4 < 16; j4 += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(bias + ch + j4);
                        v[j4 + 0] = __uint_as_float(rr[j4 + 0]) + bq.x;
                        v[j4 + 1] = __uint_as_float(rr[j4 + 1]) + bq.y;
                        v[j4 + 2] = __uint_as_float(rr[j4 + 2]) + bq.z;
                        v[j4 + 3] = __uint_as_float(rr[j4 + 3]) + bq.w;
                    }
                    uint8_t* prow = sm + OFF_H16 + (ch / 64) * FL_HPANEL + (row + FL_PAD) * FL_RB;
                    store_chunk8(prow, phase, (ch % 64) / 8, v, keep);
                    store_chunk8(prow, phase, (ch % 64) / 8 + 1, v + 8, keep);
                }
            }
        };

#pragma unroll 1
        for (int i = 0; i < FL_L; ++i) {
            mbar_wait(bar_hfin, n_hfin & 1u); ++n_hfin;
            tc_fence_after();
            to_h16(COL_H, sb_h + i * FL_H);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_h);
#pragma unroll 1
            for (int c = 0; c < FL_NCHUNK; ++c) {
                const int xb_ = c & 1;
                mbar_wait(bar_x + 8 * xb_, n_x[xb_] & 1u); ++n_x[xb_];
                tc_fence_after();
                // chunk c: buffer cols [0,48) = tanh pre-activations of channels 48c.., [48,96) = sigmoid pre-activations;
                // this thread: 24 of the 48 channels of its row
                const float* __restrict__ gb = sb_gate + i * 2 * FL_H + c * FL_NB;
                const float* __restrict__ gt = p.gcond_t ? p.gcond_t + ((size_t)b * (FL_L * 2 * FL_H) + i * 2 * FL_H + c * FL_NB) * (size_t)T + (tv ? t : 0) : nullptr;
                const uint32_t phase = swz_phase(row, FL_RB);
                uint8_t* prow = sm + OFF_ACTS + c * FL_APANEL + row * FL_RB;
                const int j0 = hsel * 24;
                const uint32_t xcol = COL_X + xb_ * FL_NB;
                uint32_t ra[24], rb[24];
                tmem_ld16(tlane + xcol + j0, reinterpret_cast<uint32_t(&)[16]>(ra[0]));
                tmem_ld8(tlane + xcol + j0 + 16, reinterpret_cast<uint32_t(&)[8]>(ra[16]));
                tmem_ld16(tlane + xcol + FL_CH + j0, reinterpret_cast<uint32_t(&)[16]>(rb[0]));
                tmem_ld8(tlane + xcol + FL_CH + j0 + 16, reinterpret_cast<uint32_t(&)[8]>(rb[16]));
                tmem_ld_wait();
                float v[24];
#pragma unroll
                for (int j = 0; j < 24; ++j) {
                    float ta = __uint_as_float(ra[j]) + gb[j0 + j];
                    float sa = __uint_as_float(rb[j]) + gb[FL_CH + j0 + j];
                    if (gt && tv) { ta += __ldg(gt + (size_t)(j0 + j) * T); sa += __ldg(gt + (size_t)(FL_CH + j0 + j) * T); }
                    v[j] = gate_act(ta, sa);
                }
#pragma unroll
                for (int g = 0; g < 3; ++g) store_chunk8(prow, phase, j0 / 8 + g, v + 8 * g, 0xffffffffu);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_xfree + 8 * xb_);
            }
        }
        // m = post(skip) complete once the res_skip MMAs of the last layer have landed
        mbar_wait(bar_hfin, n_hfin & 1u); ++n_hfin;
        tc_fence_after();
        {
            // x1 = (x1 - (post(out) + b_post)) * mask on the 112 interior frames; this thread: 48 of the 96 channels
            const bool wr = (row >= FL_HALO) && (row < 128 - FL_HALO) && (t < T) && (t >= 0);
            float* __restrict__ yo = p.y + ((size_t)b * p.y_ctot + p.out_c0 + hsel * 48) * (size_t)T + (wr ? t : 0);
#pragma unroll
            for (int cc = 0; cc < 48; cc += 16) {
                uint32_t r[16];
                float xo[16];
                tmem_ld16(tlane + COL_M + hsel * 48 + cc, r);
                if (wr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) xo[j] = yo[(size_t)(cc + j) * T];
                }
                tmem_ld_wait();
                if (wr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float m = __uint_as_float(r[j]) + sb_m[hsel * 48 + cc + j];
                        yo[(size_t)(cc + j) * T] = tv ? (xo[j] - m) : 0.f;
                    }
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

size_t flow_layer_image_bytes() { return (size_t)FL_NBLK * FL_BLOCK; }

// gate column order of in_layers for this kernel: chunk c (96 columns) = [tanh channels 48c..48c+47 | sigmoid channels 48c..]
int flow_gate_row(int col) { const int cc = col / FL_NB, j = col % FL_NB; return j < FL_CH ? cc * FL_CH + j : FL_H + cc * FL_CH + (j - FL_CH); }

// Build the block stream.  Accessors return folded fp32 weights in the kernel's channel conventions:
//   pre(co, ci)          co < 192, ci < 96
//   in(i, row, ci, tap)  row < 384 in NATURAL order (tanh rows 0..191, sigmoid rows 192..383)
//   rsm(i, row, ci)      row < 288: rows 0..191 = res rows of res_skip[i] (zero for the last layer), rows 192..287 =
//                        (W_post W_skip_i)[row-192] (for the last layer W_post W_rs), acting on acts channel ci
void flow_layer_pack(const std::function<float(int, int)>& pre, const std::function<float(int, int, int, int)>& inl,
                     const std::function<float(int, int, int)>& rsm, void* dst_host) {
    uint8_t* dst = static_cast<uint8_t*>(dst_host);
    std::memset(dst, 0, flow_layer_image_bytes());
    size_t blk = 0;
    auto put = [&](int n, int cc, float v) {
        const __half h = __float2half_rn(v);
        const uint32_t off = tc::swz_offset((uint32_t)n, (uint32_t)(cc / 8), FL_RB) + (cc % 8) * 2;
        std::memcpy(dst + blk * FL_BLOCK + off, &h, 2);
    };
    for (int pn = 0; pn < 2; ++pn)
        for (int nb = 0; nb < 2; ++nb, ++blk)
            for (int n = 0; n < FL_NB; ++n)
                for (int cc = 0; cc < 64; ++cc) {
                    const int ci = pn * 64 + cc;
                    if (ci < FL_HALF) put(n, cc, pre(nb * FL_NB + n, ci));
                }
    for (int i = 0; i < FL_L; ++i) {
        auto put_in = [&](int c) {
            for (int tap = 0; tap < FL_K; ++tap)
                for (int pn = 0; pn < 3; ++pn, ++blk)
                    for (int n = 0; n < FL_NB; ++n)
                        for (int cc = 0; cc < 64; ++cc) put(n, cc, inl(i, flow_gate_row(c * FL_NB + n), pn * 64 + cc, tap));
        };
        auto put_rs = [&](int c) {
            for (int nb = 0; nb < 3; ++nb, ++blk)
                for (int n = 0; n < FL_NB; ++n)
                    for (int cc = 0; cc < FL_CH; ++cc) put(n, cc, rsm(i, nb * FL_NB + n, c * FL_CH + cc));
        };
        put_in(0); put_in(1); put_rs(0); put_in(2); put_rs(1); put_in(3); put_rs(2); put_rs(3);
    }
}

int launch_flow_layer_tc(const FlowLayerTC& a, cudaStream_t st) {
    if (a.H != FL_H || a.half != FL_HALF || a.L != FL_L || a.k != FL_K) return SVB_ERR_UNSUPPORTED;
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (ensure_dyn_smem(flow_layer_kernel, FL_SMEM, granted)) return SVB_ERR_CUDA;
    FlowParams p;
    p.y = a.y; p.y_ctot = a.y_ctot; p.in_c0 = a.in_c0; p.out_c0 = a.out_c0;
    p.w = static_cast<const uint8_t*>(a.w);
    p.bias_gate = a.bias_gate; p.bias_h = a.bias_h; p.bias_m = a.bias_m;
    p.gcond = a.gcond; p.gcond_bstride = a.gcond_bstride > 0 ? a.gcond_bstride : FL_L * 2 * FL_H;
    p.gcond_t = a.gcond_t; p.lengths = a.lengths; p.T = a.T;
    dim3 grid((a.T + FL_TOUT - 1) / FL_TOUT, a.B);
    flow_layer_kernel<<<grid, FL_THREADS, FL_SMEM, st>>>(p);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

}  // namespace svb
