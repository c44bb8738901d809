// Windowed relative-position self-attention of enc_p on tcgen05 (sm_100a): one kernel per layer, scores never leave the SM.
//   reference: modules/attentions.py:198-239 (MultiHeadAttention.attention, heads_share, window_size 4), :246-267 (relative
//   keys / values), :275-303 (the pad/reshape skewing that this kernel replaces by direct band indexing).
//     scores[i,j] = q_i . k_j  (q pre-scaled by 1/sqrt(dk))  +  q_i . E_k[j-i+w]          for |j-i| <= w
//     p = softmax_j(scores)   (keys j >= length are filled with -1e4 like masked_fill, :253)
//     out_i = sum_j p[i,j] v_j  +  sum_{|m|<=w} p[i,i+m] E_v[m+w]
// Layout: q, k, v, out are channel-major [B, C, T] fp32 (the reference's layout); head h owns channels [h*dk, (h+1)*dk).
//
// One CTA = one (128-query tile, head, utterance).  Two passes over the key tiles (128 keys each):
//   pass 1: S = Q K^T on the tensor cores (M = 128 queries, N = 128 keys, K = dk = 96), the softmax warps keep a running
//           row maximum and sum of exponentials (scalars per row, no accumulator to rescale);
//   pass 2: S again, P = exp(S - max) / sum -> fp16 -> shared memory, O += P V on the tensor cores (N = 96, K = 128).
// Recomputing S costs 2 x 2.3 GFLOP per layer at config 2 - nothing next to the [B,2,T,T] fp32 score tensor (47 MB) the
// cuBLAS formulation wrote, re-read for the softmax and re-read again for P V.
//
// Operands arrive ready-made: the q/k/v projection GEMM (convn epilogue mode 3) writes fp16 shared-memory IMAGES of this
// kernel's tiles (Q / K: [128 rows][96 channels] as two swizzled K-major panels; V^T: [96 channels][128 keys]), so a tile is
// ONE 1-D bulk TMA copy issued by a producer thread - no per-thread global loads or conversions here (versions 1 and 2
// staged the fp32 tensors with the worker warps: 93-117 us per layer at config 2, bound by those load round trips).
// Warp roles (320 threads): warps 0-7 softmax / epilogue (warp w <-> TMEM lanes 32 (w%4), column half w/4), warp 8 streams
// the tiles (3-stage K ring, 2-stage V ring), warp 9 owns TMEM and issues the MMAs; S is double-buffered in TMEM.
//   TMEM  S0 [0,128)  S1 [128,256)  O [256,352).
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/sovits_b200.h"

namespace svb {

using namespace tc;

namespace {

constexpr int AT_THREADS = 320;          // 8 softmax warps + producer warp + MMA warp
constexpr int AT_NSOFT = 256;
constexpr int AT_KST = 3;                // K ring stages
constexpr int AT_DK = 96;
constexpr int AT_RB = 128;
constexpr int AT_PANEL = 128 * AT_RB;    // 128 rows x 64 fp16
constexpr int AT_VPANEL = AT_DK * AT_RB; // V^T: 96 rows (channels) x 64 keys
constexpr int AT_MAXW = 4;               // window <= 4 -> <= 9 band entries
constexpr uint32_t OFFA_Q = 0;                              // 2 panels (64 + 32 channels)
constexpr uint32_t OFFA_K = OFFA_Q + 2 * AT_PANEL;         // AT_KST buffers x 2 panels
constexpr uint32_t OFFA_P = OFFA_K + AT_KST * 2 * AT_PANEL;   // 2 panels (128 keys)
constexpr uint32_t OFFA_V = OFFA_P + 2 * AT_PANEL;         // 2 buffers x 2 panels x 96 rows
constexpr uint32_t OFFA_BAR = OFFA_V + 4 * AT_VPANEL;
constexpr uint32_t OFFA_F = OFFA_BAR + 160;                // floats: relk logits [128][9] | pband [128][9] | stats [2][128][2] | E_k, E_v [9][96] each
constexpr uint32_t AT_NFLOAT = 128 * 9 * 2 + 2 * 128 * 2 + 2 * 9 * AT_DK;
constexpr size_t AT_SMEM = 1024 + OFFA_F + AT_NFLOAT * 4;
constexpr int COL_S = 0, COL_O = 256;

struct AttnParams {
    const uint8_t* q_img; const uint8_t* k_img; const uint8_t* v_img;   // fp16 tile images [b][head][tile][...] (convn mode 3)
    int heads, tiles;
    const float* ek; const float* ev;                  // [2w+1][dk]
    float* out; int out_ctot;
    const int32_t* lengths;
    int T, window;
};

__global__ void __launch_bounds__(AT_THREADS, 1) attn_rel_kernel(const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t bb = base + OFFA_BAR;
    const uint32_t bar_q = bb;               // Q tile landed                                    (bulk-copy transaction bytes)
    const uint32_t bar_kfull = bb + 8;       // [3] K tile landed                                (tx)
    const uint32_t bar_kfree = bb + 32;      // [3] S-MMA that read the K buffer complete        (tcgen05.commit)
    const uint32_t bar_vfull = bb + 56;      // [2] V^T tile landed                              (tx)
    const uint32_t bar_vfree = bb + 72;      // [2] O-MMA that read the V buffer complete        (commit)
    const uint32_t bar_sfull = bb + 88;      // [2] S accumulator complete                       (commit)
    const uint32_t bar_sfree = bb + 104;     // [2] S accumulator drained by the softmax warps   (256)
    const uint32_t bar_pfull = bb + 120;     // P tile written                                   (256)
    const uint32_t bar_pfree = bb + 128;     // O-MMA that read P complete                       (commit)
    const uint32_t tmem_slot = bb + 136;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + OFFA_BAR + 136);
    float* s_relk = reinterpret_cast<float*>(sm + OFFA_F);   // [128][9]
    float* s_pband = s_relk + 128 * 9;                        // [128][9]
    float* s_stat = s_pband + 128 * 9;                        // [2 halves][128][max, sum]
    float* s_ek = s_stat + 2 * 128 * 2;                       // [9][96]
    float* s_ev = s_ek + 9 * AT_DK;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, h = blockIdx.y;
    const int i0 = blockIdx.x * 128;
    const int T = p.T, w = p.window, nb = 2 * w + 1;
    const int len = p.lengths ? min(p.lengths[b], T) : T;
    const int nt = (T + 127) / 128;

    if (tid == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < AT_KST; ++s) { mbar_init(bar_kfull + 8 * s, 1); mbar_init(bar_kfree + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_vfull + 8 * s, 1); mbar_init(bar_vfree + 8 * s, 1);
            mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sfree + 8 * s, AT_NSOFT);
        }
        mbar_init(bar_pfull, AT_NSOFT);
        mbar_init(bar_pfree, 1);
        fence_barrier_init();
    }
    if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = tid; i < 9 * AT_DK; i += AT_THREADS) { s_ek[i] = i < nb * AT_DK ? __ldg(p.ek + i) : 0.f; s_ev[i] = i < nb * AT_DK ? __ldg(p.ev + i) : 0.f; }
    for (int i = tid; i < 128 * 9; i += AT_THREADS) { s_pband[i] = 0.f; s_relk[i] = 0.f; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const size_t bh = ((size_t)b * p.heads + h) * (size_t)p.tiles;             // first tile image of this (b, head)

    if (warp == 9) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint64_t q_desc = make_smem_desc(base + OFFA_Q, AT_RB, 0);
            const uint64_t p_desc = make_smem_desc(base + OFFA_P, AT_RB, 0);
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
            constexpr uint32_t idesc_o = make_idesc_f16(128, AT_DK);
            mbar_wait(bar_q, 0);
            tc_fence_after();
            uint32_t n_sf[2] = {0u, 0u}, n_v[2] = {0u, 0u}, n_p = 0;
            int g = 0;                                        // S tiles issued so far (both passes): K stage = g % 3, S buffer = g & 1
            auto s_mma = [&]() {                              // S_g = Q K^T : K = 96 = panel 0 (4 K-steps) + half of panel 1 (2)
                const int sb = g & 1, ks_ = g % AT_KST;
                mbar_wait(bar_kfull + 8 * ks_, (uint32_t)(g / AT_KST) & 1u);
                if (g >= 2) { mbar_wait(bar_sfree + 8 * sb, n_sf[sb] & 1u); ++n_sf[sb]; }
                tc_fence_after();
                const uint64_t k_desc = make_smem_desc(base + OFFA_K + ks_ * 2 * AT_PANEL, AT_RB, 0);
                for (int pn = 0; pn < 2; ++pn)
                    for (int ks = 0; ks < (pn ? 2 : 4); ++ks)
                        umma_f16(tmem_base + COL_S + sb * 128, q_desc + (uint64_t)((uint32_t)(pn * AT_PANEL + ks * 32) >> 4),
                                 k_desc + (uint64_t)((uint32_t)(pn * AT_PANEL + ks * 32) >> 4), idesc_s, (pn | ks) ? 1u : 0u);
                umma_commit(bar_sfull + 8 * sb);
                umma_commit(bar_kfree + 8 * ks_);
                ++g;
            };
            auto o_mma = [&](int jt) {                        // O += P_jt V_jt
                const int vb = jt & 1;
                mbar_wait(bar_vfull + 8 * vb, n_v[vb] & 1u); ++n_v[vb];
                mbar_wait(bar_pfull, n_p & 1u); ++n_p;
                tc_fence_after();
                const uint64_t v_desc = make_smem_desc(base + OFFA_V + vb * 2 * AT_VPANEL, AT_RB, 0);
                for (int pn = 0; pn < 2; ++pn)
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tmem_base + COL_O, p_desc + (uint64_t)((uint32_t)(pn * AT_PANEL + ks * 32) >> 4),
                                 v_desc + (uint64_t)((uint32_t)(pn * AT_VPANEL + ks * 32) >> 4), idesc_o, (jt > 0 || pn || ks) ? 1u : 0u);
                umma_commit(bar_pfree);
                umma_commit(bar_vfree + 8 * vb);
            };
            for (int jt = 0; jt < nt; ++jt) s_mma();                          // pass 1
            for (int jt = 0; jt < nt; ++jt) {                                 // pass 2: S runs one tile ahead of O
                if (jt == 0) s_mma();
                if (jt + 1 < nt) s_mma();
                o_mma(jt);
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ------------------------------------------------------------ tile producer: one elected lane, 1-D bulk copies of tile images
        constexpr uint32_t QK_BYTES = 2 * AT_PANEL, V_BYTES = 2 * AT_VPANEL;
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_q, QK_BYTES);
            bulk_g2s(base + OFFA_Q, p.q_img + (bh + blockIdx.x) * (size_t)QK_BYTES, QK_BYTES, bar_q);
            for (int g = 0; g < 2 * nt; ++g) {
                const int jt = g < nt ? g : g - nt, ks_ = g % AT_KST;
                if (g >= AT_KST) mbar_wait(bar_kfree + 8 * ks_, (uint32_t)(g / AT_KST - 1) & 1u);
                mbar_arrive_expect_tx(bar_kfull + 8 * ks_, QK_BYTES);
                bulk_g2s(base + OFFA_K + ks_ * QK_BYTES, p.k_img + (bh + jt) * (size_t)QK_BYTES, QK_BYTES, bar_kfull + 8 * ks_);
                if (g >= nt) {
                    const int vb = jt & 1;
                    if (jt >= 2) mbar_wait(bar_vfree + 8 * vb, (uint32_t)(jt / 2 - 1) & 1u);
                    mbar_arrive_expect_tx(bar_vfull + 8 * vb, V_BYTES);
                    bulk_g2s(base + OFFA_V + vb * V_BYTES, p.v_img + (bh + jt) * (size_t)V_BYTES, V_BYTES, bar_vfull + 8 * vb);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ softmax / epilogue warps
        const int q4 = warp & 3, hsel = warp >> 2;
        const int row = 32 * q4 + lane;                    // query row of the tile == TMEM lane
        const int ti = i0 + row;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q4) << 16);
        const uint32_t phase = swz_phase(row, AT_RB);
        mbar_wait(bar_q, 0);                               // Q tile landed
        {   // relative-key logits q_i . E_k[d] from the fp16 q row (this thread: 48 of the 96 channels), halves combined in smem
            float dots[2 * AT_MAXW + 1];
#pragma unroll
            for (int d = 0; d < 2 * AT_MAXW + 1; ++d) dots[d] = 0.f;
#pragma unroll
            for (int cc = 0; cc < 48; cc += 8) {
                const int ch = hsel * 48 + cc;
                const uint4 raw4 = *reinterpret_cast<const uint4*>(sm + OFFA_Q + (ch / 64) * AT_PANEL + row * AT_RB + ((((uint32_t)(ch % 64) / 8u) ^ phase) << 4));
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw4);
                float qv[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f2 = __half22float2(h2[e]); qv[2 * e] = f2.x; qv[2 * e + 1] = f2.y; }
#pragma unroll
                for (int d = 0; d < 2 * AT_MAXW + 1; ++d) {
                    float acc = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc = fmaf(qv[e], s_ek[d * AT_DK + ch + e], acc);
                    dots[d] += acc;
                }
            }
#pragma unroll
            for (int d = 0; d < 2 * AT_MAXW + 1; ++d) atomicAdd(&s_relk[row * 9 + d], dots[d]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const float* __restrict__ relk = s_relk + row * 9; // indexed by the (runtime) band position: stays in shared memory
        uint32_t n_s[2] = {0u, 0u}, n_pf = 0;
        float m_run = -3.0e38f, l_run = 0.f, m_row = 0.f, inv_l = 0.f;
        int g = 0;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                // combine the two column halves of every row: global maximum and sum
                s_stat[(hsel * 128 + row) * 2] = m_run; s_stat[(hsel * 128 + row) * 2 + 1] = l_run;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float m0 = s_stat[row * 2], l0 = s_stat[row * 2 + 1], m1 = s_stat[(128 + row) * 2], l1 = s_stat[(128 + row) * 2 + 1];
                m_row = fmaxf(m0, m1);
                inv_l = 1.f / (l0 * __expf(m0 - m_row) + l1 * __expf(m1 - m_row));
            }
            for (int jt = 0; jt < nt; ++jt, ++g) {
                const int sb = g & 1;
                mbar_wait(bar_sfull + 8 * sb, n_s[sb] & 1u); ++n_s[sb];
                tc_fence_after();
                if (pass == 1 && jt > 0) { mbar_wait(bar_pfree, n_pf & 1u); ++n_pf; }     // O-MMA of the previous tile has read P
                const int jbase = jt * 128 + hsel * 64;
                float tmax = -3.0e38f;
#pragma unroll 1
                for (int cc = 0; cc < 64; cc += 32) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(tlane + COL_S + sb * 128 + hsel * 64 + cc, r0);
                    tmem_ld16(tlane + COL_S + sb * 128 + hsel * 64 + cc + 16, r1);
                    tmem_ld_wait();
                    float sv[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int j = jbase + cc + e;
                        float s = __uint_as_float(e < 16 ? r0[e] : r1[e - 16]);
                        const int d = j - ti + w;
                        if ((unsigned)d < (unsigned)nb) s += relk[d];
                        if (j >= len) s = -1.0e4f;                  // masked_fill(mask == 0, -1e4)
                        if (j >= T) s = -3.0e38f;                   // beyond the sequence: not part of the softmax
                        sv[e] = s;
                    }
                    if (pass == 0) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) tmax = fmaxf(tmax, sv[e]);
                        const float m_new = fmaxf(m_run, tmax);
                        float acc = 0.f;
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc += __expf(sv[e] - m_new);
                        l_run = l_run * __expf(m_run - m_new) + acc;
                        m_run = m_new;
                    } else {
                        float pv[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            pv[e] = __expf(sv[e] - m_row) * inv_l;
                            const int d = jbase + cc + e - ti + w;
                            if ((unsigned)d < (unsigned)nb && jbase + cc + e < T) s_pband[row * 9 + d] = pv[e];
                        }
                        uint8_t* prow = sm + OFFA_P + hsel * AT_PANEL + row * AT_RB;
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) store_chunk8(prow, phase, cc / 8 + gq, pv + 8 * gq, 0xffffffffu);
                    }
                }
                tc_fence_before();
                if (pass == 1) { fence_proxy_async(); mbar_arrive(bar_pfull); }
                mbar_arrive(bar_sfree + 8 * sb);
            }
        }
        // final: O complete once the last O-MMA has landed
        mbar_wait(bar_pfree, n_pf & 1u); ++n_pf;
        tc_fence_after();
        asm volatile("bar.sync 1, 256;" ::: "memory");            // band probabilities of both halves visible
        {
            const bool wr = ti < T;
            float pb[2 * AT_MAXW + 1];
#pragma unroll
            for (int d = 0; d < 2 * AT_MAXW + 1; ++d) pb[d] = d < nb ? s_pband[row * 9 + d] : 0.f;
            float* __restrict__ ob = p.out + ((size_t)b * p.out_ctot + (size_t)h * AT_DK + hsel * 48) * (size_t)T + (wr ? ti : 0);
#pragma unroll
            for (int cc = 0; cc < 48; cc += 16) {
                uint32_t r[16];
                tmem_ld16(tlane + COL_O + hsel * 48 + cc, r);
                tmem_ld_wait();
                if (wr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float o = __uint_as_float(r[j]);
                        const int c = hsel * 48 + cc + j;
#pragma unroll
                        for (int d = 0; d < 2 * AT_MAXW + 1; ++d) o = fmaf(pb[d], s_ev[d * AT_DK + c], o);
                        ob[(size_t)(cc + j) * T] = o;
                    }
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

int launch_attn_rel_tc(const AttnTC& a, cudaStream_t st) {
    if (a.dk != AT_DK || a.window < 0 || a.window > AT_MAXW || a.heads < 1 || a.T < 1) return SVB_ERR_UNSUPPORTED;
    if (!a.q_img || !a.k_img || !a.v_img) return SVB_ERR_INVALID_ARG;
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (ensure_dyn_smem(attn_rel_kernel, AT_SMEM, granted)) return SVB_ERR_CUDA;
    AttnParams p;
    p.q_img = static_cast<const uint8_t*>(a.q_img); p.k_img = static_cast<const uint8_t*>(a.k_img); p.v_img = static_cast<const uint8_t*>(a.v_img);
    p.heads = a.heads; p.tiles = (a.T + 127) / 128;
    p.ek = a.ek; p.ev = a.ev; p.out = a.out; p.out_ctot = a.out_ctot;
    p.lengths = a.lengths; p.T = a.T; p.window = a.window;
    dim3 grid((a.T + 127) / 128, a.heads, a.B);
    attn_rel_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(p);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

}  // namespace svb
