// Fused HiFiGAN ResBlock1 on tcgen05 (sm_100a) for the narrow stages (C <= 64), vdecoder/hifigan/models.py:60-67:
//     for d in (d0, d1, d2):  x = x + c2_d( lrelu( c1_d( lrelu(x) ) + b1 ) ) + b2        (c1_d dilation d, c2_d dilation 1)
//     out = alpha * x + beta * out_old                                                     (the xs / num_kernels sum, :383-389)
// One CTA owns a tile of 128*MB time steps for the WHOLE block: the fp32 residual stream x lives in TMEM
// (tcgen05.st / tcgen05.ld) next to the fp32 accumulators, the fp16 MMA operand tile lives in shared memory and is
// rewritten in place by the epilogues, so HBM sees one read of x and one write of out per element instead of the
// nine read/read/write passes of the pair-by-pair schedule.  Rows near the tile edge become invalid as the six
// convolutions eat their receptive field (halo = sum_d (d+1)(k-1)/2 per side: 12 / 36 / 60 for k = 3 / 7 / 11);
// tiles overlap by 2*halo and only the interior is written.
//
// Same warp roles and operand conventions as pair_tc_kernel (kernels_tc.cu): row r of the tile <-> time tt0 + r for
// every buffer, a tap offset is a row offset of the A descriptor (the tile is padded by PAD rows of zeros on both
// sides so negative offsets stay inside the buffer).
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/sovits_b200.h"

#include <cstdlib>

namespace svb {

using namespace tc;

namespace {

constexpr int RBK_THREADS = 320;
constexpr int RBK_NWORK = 256;
constexpr int RBK_PAD = 32;          // >= max |tap offset| = 5*5 = 25

// Phase tracing for tools/bench_rb.cu (compiled only with -DSVB_TRACE): 64 clock64() slots per CTA.
#ifdef SVB_TRACE
__device__ long long* g_rb_trace = nullptr;
#define RB_TRACE(slot) do { if (g_rb_trace) g_rb_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + (slot)] = clock64(); } while (0)
// block-skewed kernel (tools/bench_rbskew.cu): 256 slots per CTA, [q][mb][what] of the CTA's SECOND tile (steady state);
// what = 0 issuer about to issue MMA(q, mb), 1 worker starts waiting for the accumulators, 2 sees them, 3 has handed the block on
#define SK_TRACE(it, q, mb, w) do { if (g_rb_trace && (it) == 1) g_rb_trace[(size_t)blockIdx.x * 256 + (((q) * 8 + (mb)) * 4 + (w))] = clock64(); } while (0)
// fine-grained epilogue trace of ONE block (conv 2, block 2, second tile): slots 208 + n
#define SK_FINE(it, q, mb, n) do { if (g_rb_trace && (it) == 1 && (q) == 2 && (mb) == 2 && q4 == 0 && lane == 0) g_rb_trace[(size_t)blockIdx.x * 256 + 208 + (n)] = clock64(); } while (0)
#else
#define RB_TRACE(slot) do { } while (0)
#define SK_TRACE(it, q, mb, w) do { } while (0)
#define SK_FINE(it, q, mb, n) do { } while (0)
#endif

struct ResblockParams {
    const float* x; float* out;
    const uint8_t* w[6];             // c1[d0], c2[d0], c1[d1], c2[d1], c1[d2], c2[d2] tensor-core images
    const float* bias[6];
    int T, k, halo;
    int dil[3];
    float alpha, beta;
    float inv[3];                    // per pair 1/(s1*s2) of the range-normalised weight images
    uint32_t epoch; int skew_clk;    // first-wave de-phasing (tc_common.cuh)
    int red_old;                     // beta == 1 handled with red.global.add instead of load + store
    int stage_bytes, nstage;         // block-skewed kernel: weight ring geometry (a stage holds one whole conv)
    int tiles_per_item, B;           // block-skewed kernel: persistent tile list (item-major)
    int pf_q;                        // conv index at which the next tile is prefetched into L2 (-1: off)
    int park_ns;                     // block-skewed kernel: suspend-time hint of the barrier waits (0 = poll)
    int dual;                        // block-skewed kernel: two MMA issuers (warp 8: even row blocks, warp 9: odd ones + the weight ring)
};
__device__ unsigned long long g_rb_ticket[256];

template <int C, int STAGE_KB>
struct RBGeom {
    static constexpr int RB = C * 2;                 // operand row bytes (C <= 64: one K-panel)
    static constexpr int KSTEPS = C / 16;
    static constexpr int SUB = C * RB;               // one tap's weight block
    static constexpr int SPC_RAW = STAGE_KB * 1024 / SUB;
    static constexpr int SPC = SPC_RAW < 1 ? 1 : (SPC_RAW > 11 ? 11 : SPC_RAW);
    static constexpr int STAGE_BYTES = ((SPC * SUB + 1023) / 1024) * 1024;
};

template <int C, int MB, int STAGE_KB>
constexpr size_t resblock_smem_bytes() {
    using G = RBGeom<C, STAGE_KB>;
    return 1024 + (size_t)(128 * MB + 2 * RBK_PAD) * G::RB + 2 * (size_t)G::STAGE_BYTES + 256 + 6 * C * 4;
}

template <int C, int MB, int STAGE_KB, int MINB>
__global__ void __launch_bounds__(RBK_THREADS, MINB) resblock_tc_kernel(const ResblockParams p) {
    using G = RBGeom<C, STAGE_KB>;
    constexpr int R1 = 128 * MB;
    constexpr int AROWS = R1 + 2 * RBK_PAD;
    constexpr int TMEM_COLS = 2 * MB * C;
    constexpr int ACC0 = MB * C;                      // accumulator column base (residual at 0)
    static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
    static_assert(C == 16 || C == 32 || C == 64, "fused ResBlock kernel serves C <= 64");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t ring_base = base + AROWS * G::RB;
    const uint32_t bar_base = ring_base + 2 * G::STAGE_BYTES;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 16, bar_a = bar_base + 32, bar_acc = bar_base + 40;
    const uint32_t tmem_slot = bar_base + 48;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
    float* sbias = reinterpret_cast<float*>(sm + (bar_base + 256 - base));     // [6][C]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int k = p.k;
    const int TOUT = R1 - 2 * p.halo;
    const int tt0 = blockIdx.x * TOUT - p.halo;       // time of tile row 0
    const float* __restrict__ xb = p.x + (size_t)b * C * p.T;
    float* __restrict__ ob = p.out + (size_t)b * C * p.T;

    if (tid == 0) {
        dephase_first_wave(g_rb_ticket, p.epoch, p.skew_clk, MINB);
        RB_TRACE(0);
#ifdef SVB_TRACE
        { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); if (g_rb_trace) g_rb_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + 63] = smid; }
#endif
        for (int s = 0; s < 2; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_a, RBK_NWORK);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    for (int i = tid; i < 6 * C; i += RBK_THREADS) sbias[i] = __ldg(p.bias[i / C] + (i % C));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9) {
        // ------------------------------------------------------------ weight producer: six convolutions back to back
        int chunk = 0;
        for (int q = 0; q < 6; ++q) {
            const uint8_t* wsrc = p.w[q];
            for (int sb0 = 0; sb0 < k; sb0 += G::SPC, ++chunk) {
                const int s = chunk & 1;
                if (chunk >= 2) mbar_wait(bar_empty + 8 * s, ((chunk >> 1) - 1) & 1);
                const int nsb = (k - sb0) < G::SPC ? (k - sb0) : G::SPC;
                const uint32_t bytes = (uint32_t)nsb * G::SUB;
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                    bulk_g2s(ring_base + s * G::STAGE_BYTES, wsrc + (size_t)sb0 * G::SUB, bytes, bar_full + 8 * s);
                }
                __syncwarp();
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp stays converged; ONE elected lane runs the issue loop of a conv (waits on the weight ring
        // included).  The loop body is kept to the UTCHMMAs plus two descriptor increments: every extra scalar
        // instruction here is serialised latency in front of an asynchronous 40-clk MMA (measured: the former
        // div/mod + R2UR heavy body cost ~480 clk per tap and capped the tensor pipe at ~45 %).
        constexpr uint32_t idesc = make_idesc_f16(128, C);
        const int h = (k - 1) / 2;
        const uint32_t nchunk = (uint32_t)((k + G::SPC - 1) / G::SPC);     // ring chunks per conv
        for (int q = 0; q < 6; ++q) {
            mbar_wait(bar_a, q & 1);
            tc_fence_after();
            if (lane == 0) RB_TRACE(2 + 4 * q + 2);
            if (elect_one()) {
                const int cd = (q & 1) ? 1 : p.dil[q >> 1];
                uint64_t ad = make_smem_desc(a_base + (uint32_t)(RBK_PAD - h * cd) * G::RB, G::RB, 0);
                const uint64_t a_step = (uint64_t)((uint32_t)(cd * G::RB) >> 4);
                uint32_t chunk = (uint32_t)q * nchunk;
                uint32_t acc = 0u;
                for (int tap0 = 0; tap0 < k; tap0 += G::SPC, ++chunk) {
                    const uint32_t s = chunk & 1u;
                    mbar_wait(bar_full + 8 * s, (chunk >> 1) & 1u);
                    tc_fence_after();
                    uint64_t bd = make_smem_desc(ring_base + s * G::STAGE_BYTES, G::RB, 0);
                    const int n = (k - tap0) < G::SPC ? (k - tap0) : G::SPC;
                    for (int i = 0; i < n; ++i) {
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                            for (int ks = 0; ks < G::KSTEPS; ++ks)
                                umma_f16(tmem_base + ACC0 + mb * C, ad + (uint64_t)(((uint32_t)(mb * 128) * G::RB + ks * 32) >> 4),
                                         bd + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : acc);
                        }
                        acc = 1u;
                        ad += a_step;
                        bd += (uint64_t)(G::SUB >> 4);
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_acc);
            }
            __syncwarp();
            if (lane == 0) RB_TRACE(2 + 4 * q + 3);
        }
    } else {
        // ------------------------------------------------------------ workers
        const int q4 = warp & 3, hsel = warp >> 2;
        const int rib = 32 * q4 + lane;
        constexpr int CH = C / 2;                      // channels per warp-half
        constexpr int CG = CH < 16 ? CH : 16;          // columns per tcgen05.ld/st
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q4) << 16);

        // zero the PAD rows above and below the tile (never written again)
        for (int i = tid; i < 2 * RBK_PAD * (G::RB / 16); i += RBK_NWORK) {
            const int rr = i / (G::RB / 16), ch = i % (G::RB / 16);
            const int row = rr < RBK_PAD ? rr : (R1 + rr);
            *reinterpret_cast<uint4*>(sm + swz_offset(row, ch, G::RB)) = make_uint4(0, 0, 0, 0);
        }
        // (0) load x: fp32 -> TMEM residual, lrelu -> fp16 operand tile
        const int cbase = hsel * CH;
        static_assert(MB % 2 == 0, "loader batches two row blocks");
#pragma unroll 1
        for (int mb0 = 0; mb0 < MB; mb0 += 2) {        // two row blocks (2 x CH loads per thread) in flight at a time
            float v[2][CH];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int t = tt0 + (mb0 + u) * 128 + rib;
                const bool valid = (t >= 0) && (t < p.T);
                const float* __restrict__ xt = xb + (valid ? t : 0) + (size_t)cbase * p.T;
#pragma unroll
                for (int j = 0; j < CH; ++j) v[u][j] = valid ? __ldg(xt + (size_t)j * p.T) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int mb = mb0 + u;
                const int row = mb * 128 + rib;
                uint8_t* prow = sm + (row + RBK_PAD) * G::RB;
                const uint32_t phase = swz_phase(row + RBK_PAD, G::RB);
#pragma unroll
                for (int cc = 0; cc < CH; cc += CG) {
                    const int c0 = cbase + cc;
                    uint32_t r[16];
#pragma unroll
                    for (int j = 0; j < CG; ++j) r[j] = __float_as_uint(v[u][cc + j]);
                    if (CG == 16) tmem_st16(tlane + mb * C + c0, r);
                    else tmem_st8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r));
                    float w[16];
#pragma unroll
                    for (int j = 0; j < CG; ++j) w[j] = lrelu01(v[u][cc + j]);
                    store_chunk8(prow, phase, c0 / 8, w, 0xffffffffu);
                    if (CG == 16) store_chunk8(prow, phase, c0 / 8 + 1, w + 8, 0xffffffffu);
                }
            }
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(bar_a);
        if (tid == 0) RB_TRACE(1);

#pragma unroll 1
        for (int q = 0; q < 6; ++q) {
            mbar_wait(bar_acc, q & 1);
            tc_fence_after();
            if (tid == 0) RB_TRACE(2 + 4 * q + 0);
            const float* __restrict__ bq_ = sbias + q * C;
            const float inv_q = p.inv[q >> 1];
            if ((q & 1) == 0) {
                // ---- first conv of a pair: mid = lrelu(acc + b1) -> operand tile (two row blocks per TMEM round trip)
#pragma unroll 1
                for (int mb0 = 0; mb0 < MB; mb0 += 2) {
                    uint32_t r[2][CH];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int cc = 0; cc < CH; cc += CG) {
                            if (CG == 16) tmem_ld16(tlane + ACC0 + (mb0 + u) * C + cbase + cc, reinterpret_cast<uint32_t(&)[16]>(r[u][cc]));
                            else tmem_ld8(tlane + ACC0 + (mb0 + u) * C + cbase + cc, reinterpret_cast<uint32_t(&)[8]>(r[u][cc]));
                        }
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int row = (mb0 + u) * 128 + rib;
                        const int t = tt0 + row;
                        const uint32_t keep = ((t >= 0) && (t < p.T)) ? 0xffffffffu : 0u;
                        uint8_t* prow = sm + (row + RBK_PAD) * G::RB;
                        const uint32_t phase = swz_phase(row + RBK_PAD, G::RB);
#pragma unroll
                        for (int cc = 0; cc < CH; cc += CG) {
                            const int c0 = cbase + cc;
                            float v[16];
#pragma unroll
                            for (int j4 = 0; j4 < CG; j4 += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(bq_ + c0 + j4);
                                v[j4 + 0] = lrelu01(__uint_as_float(r[u][cc + j4 + 0]) + bb.x);
                                v[j4 + 1] = lrelu01(__uint_as_float(r[u][cc + j4 + 1]) + bb.y);
                                v[j4 + 2] = lrelu01(__uint_as_float(r[u][cc + j4 + 2]) + bb.z);
                                v[j4 + 3] = lrelu01(__uint_as_float(r[u][cc + j4 + 3]) + bb.w);
                            }
                            store_chunk8(prow, phase, c0 / 8, v, keep);
                            if (CG == 16) store_chunk8(prow, phase, c0 / 8 + 1, v + 8, keep);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_a);
                if (tid == 0) RB_TRACE(2 + 4 * q + 1);
            } else if (q < 5) {
                // ---- second conv of pair 0/1: x <- x + acc + b2 (TMEM), operand tile <- lrelu(x)
#pragma unroll 1
                for (int mb = 0; mb < MB; ++mb) {
                    const int row = mb * 128 + rib;
                    const int t = tt0 + row;
                    const uint32_t keep = ((t >= 0) && (t < p.T)) ? 0xffffffffu : 0u;
                    uint8_t* prow = sm + (row + RBK_PAD) * G::RB;
                    const uint32_t phase = swz_phase(row + RBK_PAD, G::RB);
#pragma unroll
                    for (int cc = 0; cc < CH; cc += CG) {
                        const int c0 = cbase + cc;
                        uint32_t r[16], xr[16];
                        if (CG == 16) { tmem_ld16(tlane + ACC0 + mb * C + c0, r); tmem_ld16(tlane + mb * C + c0, xr); }
                        else { tmem_ld8(tlane + ACC0 + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r)); tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(xr)); }
                        tmem_ld_wait();
                        float v[16];
#pragma unroll
                        for (int j4 = 0; j4 < CG; j4 += 4) {
                            const float4 bb = *reinterpret_cast<const float4*>(bq_ + c0 + j4);
                            v[j4 + 0] = fmaf(__uint_as_float(r[j4 + 0]), inv_q, bb.x) + __uint_as_float(xr[j4 + 0]);
                            v[j4 + 1] = fmaf(__uint_as_float(r[j4 + 1]), inv_q, bb.y) + __uint_as_float(xr[j4 + 1]);
                            v[j4 + 2] = fmaf(__uint_as_float(r[j4 + 2]), inv_q, bb.z) + __uint_as_float(xr[j4 + 2]);
                            v[j4 + 3] = fmaf(__uint_as_float(r[j4 + 3]), inv_q, bb.w) + __uint_as_float(xr[j4 + 3]);
                        }
#pragma unroll
                        for (int j = 0; j < CG; ++j) xr[j] = __float_as_uint(v[j]);
                        if (CG == 16) tmem_st16(tlane + mb * C + c0, xr);
                        else tmem_st8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(xr));
#pragma unroll
                        for (int j = 0; j < CG; ++j) v[j] = lrelu01(v[j]);
                        store_chunk8(prow, phase, c0 / 8, v, keep);
                        if (CG == 16) store_chunk8(prow, phase, c0 / 8 + 1, v + 8, keep);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_a);
                if (tid == 0) RB_TRACE(2 + 4 * q + 1);
            } else {
                // ---- last conv: out = alpha*(x + acc + b2) + beta*out_old for the interior rows
                const bool red_old = p.red_old != 0;             // out += y as a fire-and-forget reduction (SVB_RB_RED=1): no read of out
                const bool has_beta = p.beta != 0.f && !red_old;
#pragma unroll 1
                for (int mb = 0; mb < MB; ++mb) {
                    const int row = mb * 128 + rib;
                    const int t = tt0 + row;
                    const bool wr = (t < p.T) && (row >= p.halo) && (row < R1 - p.halo);
                    float* __restrict__ ot = ob + (wr ? t : 0);
#pragma unroll
                    for (int cc = 0; cc < CH; cc += CG) {
                        const int c0 = cbase + cc;
                        uint32_t r[16], xr[16];
                        float oo[16];
                        if (CG == 16) { tmem_ld16(tlane + ACC0 + mb * C + c0, r); tmem_ld16(tlane + mb * C + c0, xr); }
                        else { tmem_ld8(tlane + ACC0 + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r)); tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(xr)); }
                        if (wr && has_beta) {
#pragma unroll
                            for (int j = 0; j < CG; ++j) oo[j] = ot[(size_t)(c0 + j) * p.T];
                        }
                        tmem_ld_wait();
                        if (wr) {
#pragma unroll
                            for (int j4 = 0; j4 < CG; j4 += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(bq_ + c0 + j4);
                                const float b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int j = j4 + e;
                                    float y = p.alpha * (fmaf(__uint_as_float(r[j]), inv_q, b4[e]) + __uint_as_float(xr[j]));
                                    if (has_beta) y = fmaf(p.beta, oo[j], y);
                                    if (red_old) atomicAdd(ot + (size_t)(c0 + j) * p.T, y);
                                    else ot[(size_t)(c0 + j) * p.T] = y;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (tid == 0) RB_TRACE(2 + 4 * 5 + 1);
        tc_fence_before();
    }

    __syncthreads();
    if (tid == 0) RB_TRACE(30);
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}


// ---------------------------------------------------------------------------------------------------------------------
// Block-skewed variant (variant 2, the default; a ring stage holds one whole conv: two stages except C = 64, k = 11).  Same data layout and math as
// resblock_tc_kernel, but the hand-off between the MMA issuer and the epilogue warps is per 128-row block instead of per
// conv: MMA(block b, conv q+1) only needs the epilogues of blocks b-1, b, b+1 of conv q (its taps reach at most 25 rows
// into the neighbours), and epilogue(b, q) only needs MMA(b, q) and MMA(b+1, q) (the latter still reads rows of block b).
// The issuer therefore runs one to two blocks ahead of the epilogue warps and the tensor pipe works while accumulators
// are being drained, inside ONE CTA - the overlap no longer depends on a second CTA being in a different phase.
// Barriers: bar_a[b] completes once per round (round 0 = loader, round q+1 = epilogue of conv q), bar_acc[b] once per conv.
// Every wait of round q only depends on arrivals of round q-1 of the other side, so the protocol cannot deadlock
// (tests/test_skew_protocol.py models it with one or two issuers and single or paired epilogues).
// Roles (320 threads): warps 0-3 / 4-7 = two epilogue groups owning the even / odd row blocks; warp 8 = MMA issuer; warp 9 =
// weight ring and, in dual mode (the default), the second issuer: warp 8 issues the even blocks, warp 9 the odd ones, each
// waiting for the previous round of blocks b-1, b, b+1 itself (the MMAs of one thread retire in order, which is what lets an
// epilogue wait for acc[b] and acc[b+1] only).  PAIR: a group takes two of its blocks (b, b+2) per barrier round trip.
// The kernel is persistent: CTA c walks the tiles c, c + gridDim.x, ...; the loads of the next tile's block b are issued by
// the thread that finishes the last epilogue of block b.
// ---------------------------------------------------------------------------------------------------------------------
template <int C, int MB, int STAGE_KB, int MINB, bool PAIR>
__global__ void __launch_bounds__(RBK_THREADS, MINB) resblock_skew_kernel(const ResblockParams p) {
    using G = RBGeom<C, STAGE_KB>;
    constexpr int R1 = 128 * MB;
    constexpr int AROWS = R1 + 2 * RBK_PAD;
    constexpr int TMEM_COLS = 2 * MB * C;
    constexpr int ACC0 = MB * C;
    constexpr int NBG = MB / 2;                      // row blocks per worker group
    static_assert(MB >= 2 && MB <= 8 && MB % 2 == 0, "block-skewed ResBlock kernel: 2..8 row blocks");
    static_assert(!PAIR || MB % 4 == 0, "paired epilogues take the blocks of a group two at a time");
    static_assert(C == 16 || C == 32 || C == 64, "block-skewed ResBlock kernel serves C <= 64");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t ring_base = base + AROWS * G::RB;
    const uint32_t bar_base = ring_base + (uint32_t)(p.nstage * p.stage_bytes);
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 16, bar_a = bar_base + 32, bar_acc = bar_base + 32 + 8 * MB;
    const uint32_t tmem_slot = bar_base + 192;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
    float* sbias = reinterpret_cast<float*>(sm + (bar_base + 256 - base));     // [6][C]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.k;
    const int TOUT = R1 - 2 * p.halo;
    const uint32_t park_ns = (uint32_t)p.park_ns;      // > 0: waiting threads are parked by the hardware instead of polling
    auto WAIT = [park_ns](uint32_t bar, uint32_t parity) { if (park_ns) mbar_wait_park(bar, parity, park_ns); else mbar_wait(bar, parity); };
    // Persistent tile loop: CTA c owns the tiles c, c + gridDim.x, ... of the (item, tile) list.  The load of the next tile's
    // block b is issued by the thread that has just finished the last epilogue of block b, so the x read, the output
    // reduction and the first MMAs of the next tile overlap the tail of this one inside the CTA (before: load + final
    // epilogue = 16 k of a 53 k clk tile with the tensor pipe idle unless the co-resident CTA happened to be in an MMA phase).
    const int ntile = p.tiles_per_item * p.B;
    const int niter = (ntile - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        dephase_first_wave(g_rb_ticket, p.epoch, p.skew_clk, MINB);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, p.dual ? 2 : 1); }
        for (int m = 0; m < MB; ++m) { mbar_init(bar_a + 8 * m, RBK_NWORK / 2); mbar_init(bar_acc + 8 * m, 1); }
        fence_barrier_init();
    }
    // The first x tile of a worker (its row of every block its group owns, all channels) is requested before the set-up
    // barrier: one HBM latency per tile instead of one per block.
    float xin[NBG * C];
    if (warp < 8) {
        const int grp0 = warp >> 2, rib0 = 32 * (warp & 3) + lane;
        const int g0 = (int)blockIdx.x;
        const int tt00 = (g0 % p.tiles_per_item) * TOUT - p.halo;
        const float* __restrict__ xb0 = p.x + (size_t)(g0 / p.tiles_per_item) * C * p.T;
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            const int t = tt00 + (grp0 + 2 * i) * 128 + rib0;
            const bool valid = (t >= 0) && (t < p.T);
            const float* __restrict__ xt = xb0 + (valid ? t : 0);
#pragma unroll
            for (int c = 0; c < C; ++c) xin[i * C + c] = valid ? __ldg(xt + (size_t)c * p.T) : 0.f;
        }
    }
    if (warp == 8) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    for (int i = tid; i < 6 * C; i += RBK_THREADS) sbias[i] = __ldg(p.bias[i / C] + (i % C));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9 && !p.dual) {
        // ------------------------------------------------------------ weight producer: one ring stage per conv
        // (with a single stage - C = 64, k = 11 - conv q+1 can only be fetched once conv q has been consumed)
        const int ns = p.nstage;
        const int nconv = 6 * niter;
        for (int gq = 0; gq < nconv; ++gq) {
            const int q = gq % 6;
            const int s = gq % ns, use = gq / ns;
            if (use >= 1) WAIT(bar_empty + 8 * s, (use - 1) & 1);
            const uint32_t bytes = (uint32_t)k * G::SUB;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                bulk_g2s(ring_base + s * p.stage_bytes, p.w[q], bytes, bar_full + 8 * s);
            }
            __syncwarp();
        }
    } else if (warp == 8 || warp == 9) {
        // ------------------------------------------------------------ MMA issuer(s): block-major inside a conv
        // One issuing thread spends ~700 clk of serial scalar work per row block (barrier polls, descriptor arithmetic, the
        // commit) on top of the MMAs themselves - measured with tools/bench_rbskew.cu: issue-to-issue 850 / 1000 / 1200 clk
        // for k = 3 / 7 / 11 at C = 16 while the epilogue groups sit waiting for accumulators.  MMAs into different
        // accumulators are independent, so in dual mode two threads share the work: warp 8 issues the even row blocks, warp
        // 9 the odd ones and, between its blocks, keeps the weight ring filled (non-blocking polls of the empty barrier).
        // Each issuer waits for the previous round of blocks mb-1, mb, mb+1 itself; a ring stage is released by both commits.
        constexpr uint32_t idesc = make_idesc_f16(128, C);
        const int h = (k - 1) / 2;
        const int first = p.dual ? warp - 8 : 0, step = p.dual ? 2 : 1;
        const bool feeds = p.dual && warp == 9;
        if (elect_one()) {
            const int nconv = 6 * niter;
            const int ns = p.nstage;
            const uint32_t wbytes = (uint32_t)k * G::SUB;
            int next_load = 0;                                   // dual mode, warp 9: next conv whose weights have to be requested
            auto try_refill = [&](bool block) {                  // conv L goes into stage L % ns once conv L - ns has been consumed
                if (next_load >= nconv) return;
                const int L = next_load;
                const uint32_t sl = (uint32_t)(L % ns);
                if (L >= ns) {
                    const uint32_t bar = bar_empty + 8 * sl, par = (uint32_t)(L / ns - 1) & 1u;
                    if (block) WAIT(bar, par);
                    else if (!mbar_try_wait(bar, par)) return;
                }
                mbar_arrive_expect_tx(bar_full + 8 * sl, wbytes);
                bulk_g2s(ring_base + sl * p.stage_bytes, p.w[L % 6], wbytes, bar_full + 8 * sl);
                ++next_load;
            };
            if (feeds) { try_refill(true); if (ns > 1) try_refill(true); }
            int q = 0;
            for (int gq = 0; gq < nconv; ++gq, q = (q == 5 ? 0 : q + 1)) {
                const uint32_t s = (uint32_t)(gq % ns), par = (uint32_t)q & 1u;
                WAIT(bar_full + 8 * s, (uint32_t)(gq / ns) & 1u);
                const int cd = (q & 1) ? 1 : p.dil[q >> 1];
                const uint32_t a_step = (uint32_t)(cd * G::RB) >> 4;
                const uint64_t a_q = make_smem_desc(a_base + (uint32_t)(RBK_PAD - h * cd) * G::RB, G::RB, 0);
                const uint64_t b_q = make_smem_desc(ring_base + s * (uint32_t)p.stage_bytes, G::RB, 0);
                const uint32_t a_hi = (uint32_t)(a_q >> 32), b_hi = (uint32_t)(b_q >> 32);
#pragma unroll 1
                for (int mb = first; mb < MB; mb += step) {
                    if (feeds && next_load <= gq + ns - 1) try_refill(false);
                    if (step == 1) {
                        if (mb == 0) { WAIT(bar_a, par); WAIT(bar_a + 8, par); }
                        else if (mb + 1 < MB) WAIT(bar_a + 8 * (mb + 1), par);
                    } else {
                        if (mb == 1) WAIT(bar_a, par);                       // (block mb - 1 was covered by this thread's previous block otherwise)
                        WAIT(bar_a + 8 * mb, par);
                        if (mb + 1 < MB) WAIT(bar_a + 8 * (mb + 1), par);
                    }
                    tc_fence_after();
                    SK_TRACE(gq / 6, q, mb, 0);
#ifdef SVB_TRACE
                    if (g_rb_trace && q == 0 && mb == 0 && gq / 6 < 8) g_rb_trace[(size_t)blockIdx.x * 256 + 200 + gq / 6] = clock64();   // tile starts
                    if (g_rb_trace && gq == 0 && mb == 0) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); g_rb_trace[(size_t)blockIdx.x * 256 + 255] = smid; }
#endif
                    // descriptors advance in their low words only (start address field; the tile and the ring stay far below its 14 bits)
                    uint32_t ad = (uint32_t)a_q + (((uint32_t)(mb * 128) * G::RB) >> 4);
                    uint32_t bd = (uint32_t)b_q;
                    uint32_t acc = 0u;
                    for (int tap = 0; tap < k; ++tap) {
#pragma unroll
                        for (int ks = 0; ks < G::KSTEPS; ++ks)
                            umma_f16_split(tmem_base + ACC0 + mb * C, ad + (uint32_t)((ks * 32) >> 4), a_hi, bd + (uint32_t)((ks * 32) >> 4), b_hi, idesc,
                                           (ks > 0) ? 1u : acc);
                        acc = 1u;
                        ad += a_step;
                        bd += (uint32_t)(G::SUB >> 4);
                    }
                    umma_commit(bar_acc + 8 * mb);
                }
                umma_commit(bar_empty + 8 * s);
                // the next conv's weights must have been requested before anybody waits for them (single-stage ring: only now,
                // once this conv has been consumed by both issuers)
                if (feeds && next_load <= gq + 1) try_refill(true);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ workers
        // Two worker groups of four warps: group g owns the row blocks mb = g, g+2, ... and all C channels of its row.  A
        // block's epilogue is one dependent chain (barrier -> tcgen05.ld -> math -> st.shared -> fences -> arrive) of several
        // hundred cycles; two blocks in flight halve that critical path.
        const int q4 = warp & 3, grp = warp >> 2;
        const int rib = 32 * q4 + lane;
        constexpr int CH = C;
        constexpr int CG = 16;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q4) << 16);

        for (int i = tid; i < 2 * RBK_PAD * (G::RB / 16); i += RBK_NWORK) {
            const int rr = i / (G::RB / 16), ch = i % (G::RB / 16);
            const int row = rr < RBK_PAD ? rr : (R1 + rr);
            *reinterpret_cast<uint4*>(sm + swz_offset(row, ch, G::RB)) = make_uint4(0, 0, 0, 0);
        }
        // x rows of one block: fp32 -> TMEM residual, lrelu -> fp16 operand rows, then hand the block to the issuer
        auto put_block = [&](int mb, const float* v) {
            const int row = mb * 128 + rib;
            uint8_t* prow = sm + (row + RBK_PAD) * G::RB;
            const uint32_t phase = swz_phase(row + RBK_PAD, G::RB);
#pragma unroll
            for (int cc = 0; cc < CH; cc += CG) {
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < CG; ++j) r[j] = __float_as_uint(v[cc + j]);
                tmem_st16(tlane + mb * C + cc, r);
                float w[16];
#pragma unroll
                for (int j = 0; j < CG; ++j) w[j] = lrelu01(v[cc + j]);
                store_chunk8(prow, phase, cc / 8, w, 0xffffffffu);
                store_chunk8(prow, phase, cc / 8 + 1, w + 8, 0xffffffffu);
            }
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 8 * mb);
        };
#pragma unroll
        for (int i = 0; i < NBG; ++i) put_block(grp + 2 * i, xin + i * C);

        const bool red_old = p.red_old != 0;
        const bool ld_old = p.beta != 0.f && !red_old;
#pragma unroll 1
        for (int it = 0; it < niter; ++it) {
            const int g = (int)blockIdx.x + it * (int)gridDim.x;
            const int tt0 = (g % p.tiles_per_item) * TOUT - p.halo;
            float* __restrict__ ob = p.out + (size_t)(g / p.tiles_per_item) * C * p.T;
            const bool has_next = it + 1 < niter;
            const int gn = g + (int)gridDim.x;
            const int tt0n = (gn % p.tiles_per_item) * TOUT - p.halo;
            const float* __restrict__ xbn = p.x + (size_t)(has_next ? gn / p.tiles_per_item : 0) * C * p.T;
#pragma unroll 1
            for (int q = 0; q < 6; ++q) {
                const uint32_t par = (uint32_t)q & 1u;
                const float* __restrict__ bq_ = sbias + q * C;
                const float inv_q = p.inv[q >> 1];
                if (PAIR && q < 5) {
                    // Two row blocks of the group per barrier round trip (mb0 and mb0 + 2).  A block epilogue is a chain of
                    // long-latency steps (barrier polls ~200 clk each, tcgen05.ld ~150, the proxy fence ~100, the arrive ~100:
                    // measured with the fine trace of tools/bench_rbskew.cu) around ~16 values of arithmetic per thread; with two
                    // blocks per trip the chain is paid once for twice the data.
#pragma unroll 1
                    for (int mb0 = grp; mb0 < MB; mb0 += 4) {
                        if (q == p.pf_q && has_next) {
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int tw = tt0n + (mb0 + 2 * u) * 128 + 32 * q4;
#pragma unroll
                                for (int c = lane; c < 2 * C; c += 32) {
                                    const int tp = tw + ((c >= C) ? 31 : 0);
                                    if (tp >= 0 && tp < p.T) prefetch_l2(xbn + (size_t)(c % C) * p.T + tp);
                                }
                            }
                        }
                        if (q4 == 0 && lane == 0) { SK_TRACE(it, q, mb0, 1); SK_TRACE(it, q, mb0 + 2, 1); }
                        WAIT(bar_acc + 8 * mb0, par);
                        WAIT(bar_acc + 8 * (mb0 + 2), par);
                        tc_fence_after();
                        if (q4 == 0 && lane == 0) { SK_TRACE(it, q, mb0, 2); SK_TRACE(it, q, mb0 + 2, 2); }
                        uint32_t srow[2], keep[2];
                        bool interior = true;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int mb = mb0 + 2 * u;
                            const int row = mb * 128 + rib;
                            const int t = tt0 + row;
                            keep[u] = ((t >= 0) && (t < p.T)) ? 0xffffffffu : 0u;
                            srow[u] = a_base + (uint32_t)(row + RBK_PAD) * G::RB;
                            interior = interior && (tt0 + mb * 128 >= 0) && (tt0 + mb * 128 + 127 < p.T);
                        }
                        const uint32_t phase = swz_phase(rib + RBK_PAD, G::RB);      // 128 rows = a whole number of swizzle periods
                        auto wait_neighbours = [&]() {
                            // MMA(mb + 1, q) still reads operand rows of block mb: wait for it before the first store (the MMAs of
                            // one issuer retire in order, so blocks mb0 - 1 .. mb0 + 3 are all done after these two polls)
                            WAIT(bar_acc + 8 * (mb0 + 1), par);
                            if (mb0 + 3 < MB) WAIT(bar_acc + 8 * (mb0 + 3), par);
                        };
                        if ((q & 1) == 0) {
                            // mid = lrelu(acc + b1): 16 columns of both blocks per TMEM round trip
#pragma unroll
                            for (int cc = 0; cc < CH; cc += 16) {
                                uint32_t r[2][16], hq[2][8];
#pragma unroll
                                for (int u = 0; u < 2; ++u) tmem_ld16(tlane + ACC0 + (mb0 + 2 * u) * C + cc, r[u]);
                                tmem_ld_wait();
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
#pragma unroll
                                    for (int j4 = 0; j4 < 16; j4 += 4) {
                                        const float4 bb = *reinterpret_cast<const float4*>(bq_ + cc + j4);
                                        float v0 = __uint_as_float(r[u][j4 + 0]), v1 = __uint_as_float(r[u][j4 + 1]);
                                        float v2 = __uint_as_float(r[u][j4 + 2]), v3 = __uint_as_float(r[u][j4 + 3]);
                                        add2(v0, v1, bb.x, bb.y);
                                        add2(v2, v3, bb.z, bb.w);
                                        hq[u][j4 / 2] = lrelu_pack2(v0, v1);
                                        hq[u][j4 / 2 + 1] = lrelu_pack2(v2, v3);
                                    }
                                }
                                if (!interior) {
#pragma unroll
                                    for (int u = 0; u < 2; ++u)
#pragma unroll
                                        for (int j = 0; j < 8; ++j) hq[u][j] &= keep[u];
                                }
                                if (cc == 0) wait_neighbours();
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    sts128(srow[u] + ((((uint32_t)(cc / 8)) ^ phase) << 4), hq[u][0], hq[u][1], hq[u][2], hq[u][3]);
                                    sts128(srow[u] + ((((uint32_t)(cc / 8 + 1)) ^ phase) << 4), hq[u][4], hq[u][5], hq[u][6], hq[u][7]);
                                }
                            }
                        } else {
                            // x <- x + acc/s + b2 (TMEM), operand rows <- lrelu(x): 8 columns of both blocks per round trip (the
                            // residual doubles the registers per column)
#pragma unroll
                            for (int cc = 0; cc < CH; cc += 8) {
                                uint32_t r[2][8], xr[2][8], hq[2][4];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    tmem_ld8(tlane + ACC0 + (mb0 + 2 * u) * C + cc, r[u]);
                                    tmem_ld8(tlane + (mb0 + 2 * u) * C + cc, xr[u]);
                                }
                                tmem_ld_wait();
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
#pragma unroll
                                    for (int j4 = 0; j4 < 8; j4 += 4) {
                                        const float4 bb = *reinterpret_cast<const float4*>(bq_ + cc + j4);
                                        float v0, v1, v2, v3;
                                        fma2(v0, v1, __uint_as_float(r[u][j4 + 0]), __uint_as_float(r[u][j4 + 1]), inv_q, inv_q, bb.x, bb.y);
                                        fma2(v2, v3, __uint_as_float(r[u][j4 + 2]), __uint_as_float(r[u][j4 + 3]), inv_q, inv_q, bb.z, bb.w);
                                        add2(v0, v1, __uint_as_float(xr[u][j4 + 0]), __uint_as_float(xr[u][j4 + 1]));
                                        add2(v2, v3, __uint_as_float(xr[u][j4 + 2]), __uint_as_float(xr[u][j4 + 3]));
                                        xr[u][j4 + 0] = __float_as_uint(v0); xr[u][j4 + 1] = __float_as_uint(v1);
                                        xr[u][j4 + 2] = __float_as_uint(v2); xr[u][j4 + 3] = __float_as_uint(v3);
                                        hq[u][j4 / 2] = lrelu_pack2(v0, v1);
                                        hq[u][j4 / 2 + 1] = lrelu_pack2(v2, v3);
                                    }
                                    tmem_st8(tlane + (mb0 + 2 * u) * C + cc, xr[u]);
                                }
                                if (!interior) {
#pragma unroll
                                    for (int u = 0; u < 2; ++u)
#pragma unroll
                                        for (int j = 0; j < 4; ++j) hq[u][j] &= keep[u];
                                }
                                if (cc == 0) wait_neighbours();
#pragma unroll
                                for (int u = 0; u < 2; ++u)
                                    sts128(srow[u] + ((((uint32_t)(cc / 8)) ^ phase) << 4), hq[u][0], hq[u][1], hq[u][2], hq[u][3]);
                            }
                        }
                        if (q & 1) tmem_st_wait();
                        tc_fence_before();
                        fence_proxy_async();
                        mbar_arrive(bar_a + 8 * mb0);
                        mbar_arrive(bar_a + 8 * (mb0 + 2));
                        if (q4 == 0 && lane == 0) { SK_TRACE(it, q, mb0, 3); SK_TRACE(it, q, mb0 + 2, 3); }
                    }
                    continue;
                }
#pragma unroll 1
                for (int mb = grp; mb < MB; mb += 2) {
                    float xn[C];
                    if (q == p.pf_q && has_next) {
                        // pull the next tile's lines of this warp into L2 two convolutions ahead (no registers held): the
                        // demand loads below then pay an L2 hit instead of a loaded-DRAM round trip
                        const int tw = tt0n + mb * 128 + 32 * q4;                 // first time step of this warp's rows
#pragma unroll
                        for (int c = lane; c < 2 * C; c += 32) {
                            const int tp = tw + ((c >= C) ? 31 : 0);              // the 32 steps may straddle two 128-byte lines
                            if (tp >= 0 && tp < p.T) prefetch_l2(xbn + (size_t)(c % C) * p.T + tp);
                        }
                    }
                    if (q == 5 && has_next) {
                        // next tile, same block: the loads fly while this block's output is reduced into HBM
                        const int tn = tt0n + mb * 128 + rib;
                        const bool valid = (tn >= 0) && (tn < p.T);
#pragma unroll
                        for (int c = 0; c < C; ++c) xn[c] = 0.f;
                        if (valid) {
                            const float* __restrict__ xt = xbn + tn;          // 32-bit element offsets (C * T < 2^31): one IMAD.WIDE per load
#pragma unroll
                            for (int c = 0; c < C; ++c) xn[c] = ldg_nc_v(xt + c * p.T);
                        }
                    }
                    if (q4 == 0 && lane == 0) SK_TRACE(it, q, mb, 1);
                    WAIT(bar_acc + 8 * mb, par);
                    // MMA(mb+1, q) still reads operand rows of this block: that wait is only needed before the first shared-memory
                    // store of the epilogue, so it is taken after the TMEM loads and the arithmetic (off the critical path)
                    const bool wait_nb = (q < 5) && (mb + 1 < MB);
                    tc_fence_after();
                    if (q4 == 0 && lane == 0) SK_TRACE(it, q, mb, 2);
                    const int row = mb * 128 + rib;
                    const int t = tt0 + row;
                    const uint32_t keep = ((t >= 0) && (t < p.T)) ? 0xffffffffu : 0u;
                    uint8_t* prow = sm + (row + RBK_PAD) * G::RB;
                    const uint32_t phase = swz_phase(row + RBK_PAD, G::RB);
                    // whole block inside [0, T): no row has to be zeroed (uniform over the CTA; only the first / last tile of an item fails it)
                    const bool interior = (tt0 + mb * 128 >= 0) && (tt0 + mb * 128 + 127 < p.T);
                    const uint32_t srow = a_base + (uint32_t)(row + RBK_PAD) * G::RB;
                    if ((q & 1) == 0) {
                        // first conv of a pair: mid = lrelu(acc + b1) -> operand rows (two column groups per TMEM round trip)
#pragma unroll
                        for (int cc = 0; cc < CH; cc += 2 * CG) {
                            uint32_t r0[16], r1[16];
                            const bool two = (cc + CG) < CH;
                            SK_FINE(it, q, mb, 0);
                            tmem_ld16(tlane + ACC0 + mb * C + cc, r0);
                            if (two) tmem_ld16(tlane + ACC0 + mb * C + cc + CG, r1);
                            tmem_ld_wait();
                            SK_FINE(it, q, mb, 1);
#pragma unroll
                            for (int gg = 0; gg < 2; ++gg) {
                                if (gg == 1 && !two) break;
                                const int c0 = cc + gg * CG;
                                const uint32_t* rr = gg ? r1 : r0;
                                uint32_t hq[8];
#pragma unroll
                                for (int j4 = 0; j4 < CG; j4 += 4) {
                                    const float4 bb = *reinterpret_cast<const float4*>(bq_ + c0 + j4);
                                    float v0 = __uint_as_float(rr[j4 + 0]), v1 = __uint_as_float(rr[j4 + 1]);
                                    float v2 = __uint_as_float(rr[j4 + 2]), v3 = __uint_as_float(rr[j4 + 3]);
                                    add2(v0, v1, bb.x, bb.y);
                                    add2(v2, v3, bb.z, bb.w);
                                    hq[j4 / 2] = lrelu_pack2(v0, v1);
                                    hq[j4 / 2 + 1] = lrelu_pack2(v2, v3);
                                }
                                if (!interior) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) hq[j] &= keep;
                                }
                                if (cc == 0 && gg == 0) SK_FINE(it, q, mb, 2);
                                if (cc == 0 && gg == 0 && wait_nb) WAIT(bar_acc + 8 * (mb + 1), par);
                                if (cc == 0 && gg == 0) SK_FINE(it, q, mb, 3);
                                sts128(srow + ((((uint32_t)(c0 / 8)) ^ phase) << 4), hq[0], hq[1], hq[2], hq[3]);
                                sts128(srow + ((((uint32_t)(c0 / 8 + 1)) ^ phase) << 4), hq[4], hq[5], hq[6], hq[7]);
                            }
                        }
                        SK_FINE(it, q, mb, 4);
                        tc_fence_before();
                        fence_proxy_async();
                        SK_FINE(it, q, mb, 5);
                        mbar_arrive(bar_a + 8 * mb);
                        SK_FINE(it, q, mb, 6);
                    } else if (q < 5) {
                        // second conv of pair 0/1: x <- x + acc + b2 (TMEM), operand rows <- lrelu(x)
#pragma unroll
                        for (int cc = 0; cc < CH; cc += CG) {
                            uint32_t r[16], xr[16];
                            tmem_ld16(tlane + ACC0 + mb * C + cc, r);
                            tmem_ld16(tlane + mb * C + cc, xr);
                            tmem_ld_wait();
                            uint32_t hq[8];
#pragma unroll
                            for (int j4 = 0; j4 < CG; j4 += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(bq_ + cc + j4);
                                float v0, v1, v2, v3;
                                fma2(v0, v1, __uint_as_float(r[j4 + 0]), __uint_as_float(r[j4 + 1]), inv_q, inv_q, bb.x, bb.y);
                                fma2(v2, v3, __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]), inv_q, inv_q, bb.z, bb.w);
                                add2(v0, v1, __uint_as_float(xr[j4 + 0]), __uint_as_float(xr[j4 + 1]));
                                add2(v2, v3, __uint_as_float(xr[j4 + 2]), __uint_as_float(xr[j4 + 3]));
                                xr[j4 + 0] = __float_as_uint(v0); xr[j4 + 1] = __float_as_uint(v1);
                                xr[j4 + 2] = __float_as_uint(v2); xr[j4 + 3] = __float_as_uint(v3);
                                hq[j4 / 2] = lrelu_pack2(v0, v1);
                                hq[j4 / 2 + 1] = lrelu_pack2(v2, v3);
                            }
                            tmem_st16(tlane + mb * C + cc, xr);
                            if (!interior) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) hq[j] &= keep;
                            }
                            if (cc == 0 && wait_nb) WAIT(bar_acc + 8 * (mb + 1), par);
                            sts128(srow + ((((uint32_t)(cc / 8)) ^ phase) << 4), hq[0], hq[1], hq[2], hq[3]);
                            sts128(srow + ((((uint32_t)(cc / 8 + 1)) ^ phase) << 4), hq[4], hq[5], hq[6], hq[7]);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        fence_proxy_async();
                        mbar_arrive(bar_a + 8 * mb);
                    } else {
                        // last conv: out = alpha*(x + acc + b2) (+ beta*out_old) for the interior rows
                        const bool wr = (t >= 0) && (t < p.T) && (row >= p.halo) && (row < R1 - p.halo);
                        float* __restrict__ ot = ob + (wr ? t : 0);
#pragma unroll
                        for (int cc = 0; cc < CH; cc += CG) {
                            uint32_t r[16], xr[16];
                            float oo[16];
                            tmem_ld16(tlane + ACC0 + mb * C + cc, r);
                            tmem_ld16(tlane + mb * C + cc, xr);
                            if (wr && ld_old) {
#pragma unroll
                                for (int j = 0; j < CG; ++j) oo[j] = ot[(cc + j) * p.T];
                            }
                            tmem_ld_wait();
                            if (wr) {
#pragma unroll
                                for (int j4 = 0; j4 < CG; j4 += 4) {
                                    const float4 bb = *reinterpret_cast<const float4*>(bq_ + cc + j4);
                                    float y[4];
                                    fma2(y[0], y[1], __uint_as_float(r[j4 + 0]), __uint_as_float(r[j4 + 1]), inv_q, inv_q, bb.x, bb.y);
                                    fma2(y[2], y[3], __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]), inv_q, inv_q, bb.z, bb.w);
                                    add2(y[0], y[1], __uint_as_float(xr[j4 + 0]), __uint_as_float(xr[j4 + 1]));
                                    add2(y[2], y[3], __uint_as_float(xr[j4 + 2]), __uint_as_float(xr[j4 + 3]));
                                    mul2(y[0], y[1], y[0], y[1], p.alpha, p.alpha);
                                    mul2(y[2], y[3], y[2], y[3], p.alpha, p.alpha);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const int j = j4 + e;
                                        if (ld_old) y[e] = fmaf(p.beta, oo[j], y[e]);
                                        if (red_old) atomicAdd(ot + (cc + j) * p.T, y[e]);
                                        else ot[(cc + j) * p.T] = y[e];
                                    }
                                }
                            }
                        }
                        if (has_next) {
                            // rows of this block are still read by the last MMAs of blocks mb-1 .. mb+1 of this tile
                            if (mb + 1 < MB) WAIT(bar_acc + 8 * (mb + 1), par);
                            tc_fence_after();
                            put_block(mb, xn);
                        }
                    }
                    if (q4 == 0 && lane == 0) SK_TRACE(it, q, mb, 3);
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

uint32_t g_rb_epoch = 0;         // launch serial number shared by all instantiations (they share g_rb_ticket)

int rb_env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}

template <int C, int MB, int STAGE_KB, int MINB, bool SKEW, bool PAIR = false>
int launch_resblock_t(const ResblockTC& a, cudaStream_t st) {
    constexpr size_t smem = resblock_smem_bytes<C, MB, STAGE_KB>();
    static_assert((smem + 1024) * MINB <= 228 * 1024, "fused ResBlock kernel shared memory exceeds the SM budget");
    auto kernel = [] {
        if constexpr (SKEW) return resblock_skew_kernel<C, MB, STAGE_KB, MINB, PAIR>;
        else return resblock_tc_kernel<C, MB, STAGE_KB, MINB>;
    }();
    // block-skewed kernel: a ring stage holds one whole conv (k taps); two stages when they fit next to the operand tile
    using GG = RBGeom<C, STAGE_KB>;
    const int stage_bytes = ((a.k * GG::SUB + 1023) / 1024) * 1024;
    const size_t fixed = 1024 + (size_t)(128 * MB + 2 * RBK_PAD) * GG::RB + 256 + 6 * C * 4;
    const size_t budget = (size_t)(MINB == 1 ? 227 : 113) * 1024;
    const int nstage = (fixed + 2 * (size_t)stage_bytes <= budget) ? 2 : 1;
    const size_t smem_run = SKEW ? fixed + (size_t)nstage * stage_bytes : smem;
    if (SKEW && smem_run > budget) return SVB_ERR_UNSUPPORTED;
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (ensure_dyn_smem(kernel, smem_run, granted)) return SVB_ERR_CUDA;
    ResblockParams p;
    p.x = a.x; p.out = a.out;
    for (int q = 0; q < 6; ++q) { p.w[q] = static_cast<const uint8_t*>(a.w[q]); p.bias[q] = a.bias[q]; }
    p.T = a.T; p.k = a.k; p.alpha = a.alpha; p.beta = a.beta;
    for (int d = 0; d < 3; ++d) p.inv[d] = a.inv[d];
    int halo = 0;
    for (int d = 0; d < 3; ++d) { p.dil[d] = a.dil[d]; halo += (a.dil[d] + 1) * (a.k - 1) / 2; }
    p.halo = halo;
    {
        // first-wave de-phasing (tc_common.cuh): tile period estimate = six conv cycles (MMA at the shared-pipe rate +
        // epilogue) + load / store phases
        static const int env_skew = rb_env_int("SVB_RB_SKEW", -1);
        static const int env_red = rb_env_int("SVB_RB_RED", 1);
        p.red_old = (env_red && a.beta == 1.f) ? 1 : 0;
        p.epoch = ++g_rb_epoch;
        const int mma_clk = (C <= 32 ? 40 : 48) * a.k * MB * (C / 16);
        const int grid_ctas = (int)(((a.T + (128 * MB - 2 * halo) - 1) / (128 * MB - 2 * halo)) * a.B);
        p.skew_clk = (MINB < 2 || grid_ctas < 4 * 148 * MINB) ? 0 : (env_skew >= 0 ? env_skew : 6 * (mma_clk + 2000) + 15000);
    }
    const int TOUT = 128 * MB - 2 * halo;
    if (TOUT < 64) return SVB_ERR_UNSUPPORTED;
    dim3 grid((a.T + TOUT - 1) / TOUT, a.B);
    p.stage_bytes = stage_bytes; p.nstage = nstage; p.dual = 0; p.park_ns = 0;
    p.tiles_per_item = (int)grid.x; p.B = a.B;
    if (SKEW) {
        // persistent: one CTA per resident slot walks the (item, tile) list; SVB_RB_PERSIST=0 launches one CTA per tile
        static const int env_persist = rb_env_int("SVB_RB_PERSIST", 1);
        static const int env_pf = rb_env_int("SVB_RB_PF", 3);
        static const int env_dual = rb_env_int("SVB_RB_DUAL", 1);
        static const int env_park = rb_env_int("SVB_RB_PARK", 0);      // measured: parking (300 / 2000 ns) 3.13 vs polling 3.07 ms per step
        p.pf_q = env_pf;
        p.dual = env_dual;
        p.park_ns = env_park;
        const int ntile = (int)grid.x * a.B;
        const int slots = sm_count() * MINB;
        grid = dim3(env_persist ? (ntile < slots ? ntile : slots) : ntile, 1);
    }
    kernel<<<grid, RBK_THREADS, smem_run, st>>>(p);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

}  // namespace

// variant 0: one CTA/SM with all 512 TMEM columns (largest tile); variant 1: two CTAs/SM with 256 columns each.
int launch_resblock_tc(const ResblockTC& a, cudaStream_t st) {
    if (!(a.k == 3 || a.k == 7 || a.k == 11)) return SVB_ERR_UNSUPPORTED;
    for (int d = 0; d < 3; ++d)
        if (a.dil[d] * (a.k - 1) / 2 > RBK_PAD - 1 || a.dil[d] < 1) return SVB_ERR_UNSUPPORTED;
    static const int env_variant = rb_env_int("SVB_RB_VARIANT", -1);
    int variant = a.variant >= 0 ? a.variant : env_variant;
    if (variant < 0) variant = 2;     // measured: block-skewed hand-off 11.64 vs 11.95 ms/step
    if (variant == 2) {               // block-skewed hand-off (resblock_skew_kernel): two CTAs/SM for C <= 32, one for C = 64
        switch (a.C) {
            case 16: {
                static const int env_pair = rb_env_int("SVB_RB_PAIR16", 1);     // two row blocks per epilogue round trip
                if (env_pair) return launch_resblock_t<16, 8, 8, 2, true, true>(a, st);
                return launch_resblock_t<16, 8, 8, 2, true>(a, st);
            }
            case 32: {
                // 1024-row tiles, one CTA/SM: 1 = every kernel size, 2 = k = 11 only (the MMA-bound branch gains from the smaller
                // halo share, 120 of 1024 rows instead of 120 of 512: resblock_tc 3.04 vs 3.07 ms/step; the epilogue-bound k = 3 / 7
                // keep two CTAs per SM)
                static const int env_c32 = rb_env_int("SVB_RB_C32_ONE", 2);
                static const int env_pair = rb_env_int("SVB_RB_PAIR32", 1);
                if (env_c32 == 1 || (env_c32 == 2 && a.k == 11)) return launch_resblock_t<32, 8, 22, 1, true, true>(a, st);
                if (env_pair) return launch_resblock_t<32, 4, 22, 2, true, true>(a, st);
                return launch_resblock_t<32, 4, 22, 2, true>(a, st);
            }
            case 64: return launch_resblock_t<64, 4, 32, 1, true>(a, st);
            default: return SVB_ERR_UNSUPPORTED;
        }
    }
    if (variant == 1) {
        switch (a.C) {
            case 16: return launch_resblock_t<16, 8, 8, 2, false>(a, st);
            case 32: return launch_resblock_t<32, 4, 22, 2, false>(a, st);
            case 64: return launch_resblock_t<64, 4, 32, 1, false>(a, st);     // 2 CTAs/SM would leave 256-2*120 = 16 useful rows
            default: return SVB_ERR_UNSUPPORTED;
        }
    }
    switch (a.C) {
        case 16: return launch_resblock_t<16, 16, 8, 1, false>(a, st);
        case 32: return launch_resblock_t<32, 8, 32, 1, false>(a, st);
        case 64: return launch_resblock_t<64, 4, 32, 1, false>(a, st);
        default: return SVB_ERR_UNSUPPORTED;
    }
}

}  // namespace svb
