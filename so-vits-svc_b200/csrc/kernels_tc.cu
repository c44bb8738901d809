// tcgen05 tensor-core kernels (sm_100a): the HiFiGAN ResBlock "pair"
//     out = alpha * ( x + conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 ) + beta * out_old
// (vdecoder/hifigan/models.py:60-67, one iteration of the loop; 92.5 % of the path's FLOPs).
//
// Mapping onto the 5th-gen tensor cores: a k-tap dilated Conv1d is k shifted channel-mixing GEMMs.
//   M = 128 time steps (TMEM lanes), N = C_out (TMEM columns), K = C_in per tap.
//   A = activations, staged in shared memory as [time][channel] fp16 rows (K-major, hardware swizzle);
//       a tap shift of s samples is a ROW offset of the A descriptor, so one staged tile feeds all taps.
//   B = folded weights per tap [C_out][C_in] fp16, pre-swizzled on the host and streamed from L2 with
//       1-D bulk TMA copies through an mbarrier ring.
//   D = fp32 accumulators in TMEM; epilogue warps read them with tcgen05.ld (thread == time step),
//       apply bias + LeakyReLU, convert to fp16 and write the second conv's A tile in place.
// The residual stream stays fp32 in HBM/L2; only MMA operands are fp16 (11-bit significand, the same
// as the TF32 operands of the reference's cuDNN path).
//
// Warp roles (320 threads): warps 0-7 stage x / run epilogues (warp w owns TMEM lanes 32*(w%4)..+31 and
// half (w/4) of the channels), warp 8 allocates TMEM and issues MMAs (one elected lane), warp 9 streams
// weights.
#include "kernels.h"
#include "tc_common.cuh"
#include <cuda.h>
#include "../../include/sovits_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace svb {

using namespace tc;

namespace {

constexpr int TC_THREADS = 320;
constexpr int NWORK = 256;           // worker threads (8 warps)
constexpr int NSTAGE = 2;
constexpr int TMA_NBOX = 3;          // the A tile arrives as 3 row-boxes (box rows must be <= 256 and a multiple of 8)
template <int MB> struct TileRows { static constexpr int HALO = (MB == 4) ? 64 : 56; static constexpr int AROWS = 128 * MB + HALO; };

template <int C, int STAGE_KB = 32>
struct TCGeom {
    static constexpr int CPP = C < 64 ? C : 64;         // channels per K-panel
    static constexpr int NP = C / CPP;                  // K-panels
    static constexpr int RB = CPP * 2;                  // operand row bytes: 128 / 64 / 32
    static constexpr int KSTEPS = CPP / 16;             // MMAs (K=16) per panel row
    static constexpr int SUB = C * RB;                  // bytes of one (tap, panel) weight block: C_out rows
    static constexpr int SPC_RAW = STAGE_KB * 1024 / SUB;   // sub-blocks per ring chunk
    static constexpr int SPC = SPC_RAW < 1 ? 1 : (SPC_RAW > 11 * NP ? 11 * NP : SPC_RAW);
    static constexpr int STAGE_BYTES = ((SPC * SUB + 1023) / 1024) * 1024;
};

struct PairParams {
    const float* x; float* out;
    const uint8_t* w1; const uint8_t* w2;
    const float* b1; const float* b2;
    int T, k, dil;
    float alpha, beta, inv;
    __half* a16_out;     // optional second output: fp16 lrelu(out) in [B][T][C] (the next pair's TMA-loadable operand)
    int vec4;            // x / out 16-byte aligned and T % 4 == 0: the loader uses 128-bit loads along time
    int prefetch;        // vec4 loader: 1 = L2 prefetch of the whole tile before the demand loads, 2 = also of the slot's next tile
    int prefetch_ahead;  // resident CTAs of the launch (distance, in launch order, to the tile that follows in this slot)
    int red_out;         // loader pre-writes alpha*x (+ beta*old for beta == 1) into out, epilogue 2 only adds (RED): no x re-read
    uint32_t epoch; int dephase_clk;   // first-wave start skew (tc_common.cuh: dephase_first_wave)
};
// Up to three independent pairs (the three ResBlock branches of a stage at the same dilation index) in ONE launch:
// blockIdx.z selects the branch.  A wide stage with few tiles (stage 0: 232 tiles of C = 256 on 148 SMs = 1.57 waves per
// launch, i.e. 2 waves of time) is then scheduled as 696 CTAs of mixed length (k = 3 / 7 / 11) = 4.7 waves of the mean.
struct PairParamsN { PairParams br[3]; int nbr; };
__device__ unsigned long long g_pair_ticket[256];
uint32_t g_pair_epoch = 0;

// Phase tracing for tools/bench_pairtrace.cu (compiled only with -DSVB_TRACE): 16 clock64() slots per CTA.
#ifdef SVB_TRACE
__device__ long long* g_pair_trace = nullptr;
#define PAIR_TRACE(slot) do { if (g_pair_trace) g_pair_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
#else
#define PAIR_TRACE(slot) do { } while (0)
#endif

template <int C, int MB, int STAGE_KB>
constexpr size_t pair_smem_bytes() {
    using G = TCGeom<C, STAGE_KB>;
    return 1024 /*align slack*/ + (size_t)G::NP * TileRows<MB>::AROWS * G::RB + (size_t)NSTAGE * G::STAGE_BYTES + 256 + 2 * C * 4;
}

template <int C, int MB, int STAGE_KB, int MINB, bool TMA_IN>
__global__ void __launch_bounds__(TC_THREADS, MINB) pair_tc_kernel(const __grid_constant__ PairParamsN pn_, const __grid_constant__ CUtensorMap tmap) {
    // branches interleaved along x (CTAs are dispatched x-fastest): neighbours in launch order - and therefore the CTAs that
    // share an SM - belong to different branches, so a k = 3 pair (memory phases dominate) runs next to a k = 11 pair (MMA
    // phases dominate) instead of the launch going through one homogeneous branch after the other
    const int bz = (int)blockIdx.x % pn_.nbr, bx = (int)blockIdx.x / pn_.nbr;
    const PairParams& p = pn_.br[bz];
    if (bx * (128 * MB - (p.k - 1)) >= p.T) return;       // this branch has fewer tiles than the widest one
    using G = TCGeom<C, STAGE_KB>;
    constexpr int R1 = 128 * MB;
    constexpr int AROWS = TileRows<MB>::AROWS;
    constexpr int APANEL = AROWS * G::RB;
    constexpr int TMEM_COLS = (MB * C) < 32 ? 32 : (MB * C);
    static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;                                   // A tile (A1, then A2 in place)
    const uint32_t ring_base = base + G::NP * APANEL;               // weight ring
    const uint32_t bar_base = ring_base + NSTAGE * G::STAGE_BYTES;  // barriers
    const uint32_t bar_full = bar_base;                             // [NSTAGE]
    const uint32_t bar_empty = bar_base + 8 * NSTAGE;               // [NSTAGE]
    const uint32_t bar_a = bar_base + 16 * NSTAGE;                  // A tile ready (256 arrivals), 2 phases
    const uint32_t bar_acc = bar_a + 8;                             // accumulators ready, 2 phases
    const uint32_t bar_tma = bar_acc + 8;                           // A1 tile landed (TMA transaction bytes)
    const uint32_t tmem_slot = bar_acc + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
    float* sbias = reinterpret_cast<float*>(sm + (bar_base + 256 - base));     // [2][C]: b1 | b2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int k = p.k, dil = p.dil;
    const int TOUT = R1 - (k - 1);
    const int t0 = bx * TOUT;
    const int h2 = (k - 1) / 2, h1 = dil * (k - 1) / 2;
    const int tA0 = t0 - h2 - h1;          // time of A1 row 0
    const int tM0 = t0 - h2;               // time of mid (conv1 output / A2) row 0
    const int RA1 = R1 + (k - 1) * dil;    // rows of A1 the MMAs touch
    const float* __restrict__ xb = p.x + (size_t)b * C * p.T;
    float* __restrict__ ob = p.out + (size_t)b * C * p.T;

    // ---------------------------------------------------------------- setup
    if (tid == 0) {
        dephase_first_wave(g_pair_ticket, p.epoch, p.dephase_clk, MINB);
        PAIR_TRACE(0);
#ifdef SVB_TRACE
        { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); if (g_pair_trace) g_pair_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = smid; }
#endif
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_a, NWORK);
        mbar_init(bar_acc, 1);
        mbar_init(bar_tma, 1);
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = tid; i < 2 * C; i += TC_THREADS) sbias[i] = i < C ? __ldg(p.b1 + i) : __ldg(p.b2 + i - C);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9) {
        // ------------------------------------------------------------ weight producer (and A-tile TMA issuer)
        // whole warp converged, one elected lane issues (see elect_one() in tc_common.cuh)
        if (TMA_IN) {
            // the previous pair already wrote lrelu(x) as fp16 [B][T][C]: the tensor map delivers it swizzled, rows
            // outside [0,T) zero-filled, straight into the A-operand layout - no thread touches the data
            constexpr int BOX = AROWS / TMA_NBOX;
            if (elect_one()) {
                tma_prefetch_desc(&tmap);
                mbar_arrive_expect_tx(bar_tma, (uint32_t)(G::NP * AROWS * G::RB));
                for (int pn = 0; pn < G::NP; ++pn)
                    for (int bx = 0; bx < TMA_NBOX; ++bx)
                        tma_load_3d(a_base + pn * APANEL + bx * BOX * G::RB, &tmap, pn * G::CPP, tA0 + bx * BOX, b, bar_tma);
            }
            __syncwarp();
        }
        {
            const int total_sb = k * G::NP;
            int chunk = 0;
            for (int conv = 0; conv < 2; ++conv) {
                const uint8_t* wsrc = conv ? p.w2 : p.w1;
                for (int sb0 = 0; sb0 < total_sb; sb0 += G::SPC, ++chunk) {
                    const int s = chunk % NSTAGE;
                    if (chunk >= NSTAGE) mbar_wait(bar_empty + 8 * s, ((chunk / NSTAGE) - 1) & 1);
                    const int nsb = (total_sb - sb0) < G::SPC ? (total_sb - sb0) : G::SPC;
                    const uint32_t bytes = (uint32_t)nsb * G::SUB;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                        bulk_g2s(ring_base + s * G::STAGE_BYTES, wsrc + (size_t)sb0 * G::SUB, bytes, bar_full + 8 * s);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer
        // The warp stays converged; ONE elected lane runs the whole issue loop of a conv, ring waits included.  The
        // body is the UTCHMMAs plus descriptor increments - no div/mod, no per-MMA register->uniform moves: scalar work
        // here is serialised latency in front of asynchronous MMAs (it used to cap the tensor pipe at ~45 %).
        constexpr uint32_t idesc = make_idesc_f16(128, C);
        const int total_sb = k * G::NP;
        const uint32_t nchunk = (uint32_t)((total_sb + G::SPC - 1) / G::SPC);
        for (int conv = 0; conv < 2; ++conv) {
            if (TMA_IN) { if (conv == 0) mbar_wait(bar_tma, 0); else mbar_wait(bar_a, 0); }
            else mbar_wait(bar_a, conv);
            tc_fence_after();
            if (lane == 0) PAIR_TRACE(6 + 2 * conv);
            if (elect_one()) {
                const int cd = conv ? 1 : dil;
                uint64_t a_tap = make_smem_desc(a_base, G::RB, 0);            // A descriptor of (tap, panel 0)
                const uint64_t a_step = (uint64_t)((uint32_t)(cd * G::RB) >> 4);
                uint32_t chunk = (uint32_t)conv * nchunk;
                uint32_t acc = 0u;
                int pn = 0;
                for (int sb0 = 0; sb0 < total_sb; sb0 += G::SPC, ++chunk) {
                    const uint32_t s = chunk % NSTAGE;
                    mbar_wait(bar_full + 8 * s, (chunk / NSTAGE) & 1u);
                    tc_fence_after();
                    uint64_t bd = make_smem_desc(ring_base + s * G::STAGE_BYTES, G::RB, 0);
                    const int n = (total_sb - sb0) < G::SPC ? (total_sb - sb0) : G::SPC;
                    for (int i = 0; i < n; ++i) {
                        const uint64_t ad = a_tap + (uint64_t)pn * (uint64_t)(APANEL >> 4);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                            for (int ks = 0; ks < G::KSTEPS; ++ks)
                                umma_f16(tmem_base + mb * C, ad + (uint64_t)(((uint32_t)(mb * 128) * G::RB + ks * 32) >> 4),
                                         bd + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : acc);
                        }
                        acc = 1u;
                        bd += (uint64_t)(G::SUB >> 4);
                        if (++pn == G::NP) { pn = 0; a_tap += a_step; }
                    }
                    umma_commit(bar_empty + 8 * s);                           // frees the ring stage once these MMAs retire
                }
                umma_commit(bar_acc);                                         // accumulators of this conv are complete
            }
            __syncwarp();
            if (lane == 0) PAIR_TRACE(7 + 2 * conv);
        }
    } else {
        // ------------------------------------------------------------ workers (warps 0-7)
        // (1) stage A1 = lrelu(x) tile as fp16 [time][channel], zero outside [0,T)  (skipped when the tile comes by TMA)
        if (!TMA_IN && p.vec4) {
            // 128-bit loads along time: a lane owns an aligned group of 4 time steps x 8 channels (8 LDG.128 in flight,
            // 32 KB per CTA), transposes in registers and writes four 16-byte operand chunks.  A warp item is GPI
            // consecutive groups x CHK chunks, so every load instruction covers GPI*16 contiguous bytes per channel.
            // When red_out is set the same registers also initialise out = alpha*x (+ old), so that epilogue 2 only
            // has to ADD the convolution result and never waits on a global load.
            constexpr int CHK = (C / 8) < 4 ? (C / 8) : 4;
            constexpr int GPI = 32 / CHK;
            constexpr int NCB = (C / 8) / CHK;
            const int sh = tA0 & 3;                       // tA0 - sh is a multiple of 4 (two's complement, tA0 may be negative)
            const int tAa = tA0 - sh;
            const int NG = (RA1 + sh + 3) >> 2;
            const int NGO = (NG + GPI - 1) / GPI;
            const int g_l = lane % GPI, ch_l = lane / GPI;
            const bool acc_old = p.beta != 0.f;
            if (p.prefetch) {
                // The item loop below is a chain of dependent round trips (8 loads -> convert -> store, ~5 per thread): pull the
                // whole tile into L2 first, so that one DRAM latency is paid up front and the demand loads are L2 hits.
                const int nline = (4 * NG + 31) / 32 + 1;                 // 128-byte lines per channel row (tAa is 16-byte aligned only)
                for (int i = tid; i < C * nline; i += NWORK) {
                    const int c = i / nline, l = i - c * nline;
                    const int tp = tAa + 32 * l;
                    if (tp >= 0 && tp < p.T) prefetch_l2(xb + (size_t)c * p.T + tp);
                }
            }
#pragma unroll 1
            for (int item = warp; item < NGO * NCB; item += NWORK / 32) {
                const int go = item % NGO, cb = item / NGO;
                const int g = go * GPI + g_l;
                const int c0 = (cb * CHK + ch_l) * 8;
                const int t4 = tAa + 4 * g;
                const bool gv = (g < NG) && (t4 >= 0) && (t4 < p.T);
                const float* __restrict__ src = xb + (size_t)c0 * p.T + (gv ? t4 : 0);
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gv ? __ldg(reinterpret_cast<const float4*>(src + (size_t)j * p.T)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < NG) {
                    uint8_t* pan = sm + (c0 / G::CPP) * APANEL;
                    const int chunk = (c0 % G::CPP) / 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = 4 * g - sh + i;
                        if (r >= 0 && r < RA1) {
                            float w8[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float e = i == 0 ? v[j].x : (i == 1 ? v[j].y : (i == 2 ? v[j].z : v[j].w));
                                w8[j] = lrelu01(e);
                            }
                            store_chunk8(pan + r * G::RB, swz_phase(r, G::RB), chunk, w8, 0xffffffffu);
                        }
                    }
                }
                if (p.red_out && gv) {
                    const int o = t4 - t0;
                    float* __restrict__ dst = ob + (size_t)c0 * p.T + t4;
                    if (o >= 0 && o + 3 < TOUT) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { v[j].x *= p.alpha; v[j].y *= p.alpha; v[j].z *= p.alpha; v[j].w *= p.alpha; }
                        if (acc_old) {
                            float4 od[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) od[j] = *reinterpret_cast<const float4*>(dst + (size_t)j * p.T);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[j].x += od[j].x; v[j].y += od[j].y; v[j].z += od[j].z; v[j].w += od[j].w; }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + (size_t)j * p.T) = v[j];
                    } else if (o + 3 >= 0 && o < TOUT) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (o + i >= 0 && o + i < TOUT) {
                                    float* d = dst + (size_t)j * p.T + i;
                                    *d = acc_old ? (*d + p.alpha * e[i]) : (p.alpha * e[i]);
                                }
                        }
                    }
                }
            }
            if (p.red_out) __threadfence();              // out = alpha*x must be in L2 before epilogue 2 adds to it
        }
        for (int r = tid; !TMA_IN && !p.vec4 && r < RA1; r += NWORK) {
            const int t = tA0 + r;
            const bool valid = (t >= 0) && (t < p.T);
            const float* __restrict__ xt = xb + (valid ? t : 0);
            const uint32_t phase = swz_phase(r, G::RB);
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = valid ? __ldg(xt + (size_t)(c0 + j) * p.T) : 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = lrelu01(v[j]);
                uint8_t* prow = sm + (c0 / G::CPP) * APANEL + r * G::RB;
                store_chunk8(prow, phase, (c0 % G::CPP) / 8, v, 0xffffffffu);
                store_chunk8(prow, phase, (c0 % G::CPP) / 8 + 1, v + 8, 0xffffffffu);
            }
        }
        if (!TMA_IN) {
            fence_proxy_async();
            mbar_arrive(bar_a);
        }
        if (tid == 0) PAIR_TRACE(1);
        if (!TMA_IN && p.vec4 && p.prefetch > 1) {
            // The workers now idle until conv1's accumulators arrive: pull the tile of the CTA that will take this slot next
            // into L2.  CTAs are dispatched in linear order (x fastest), `ahead` = resident slots of the launch, so the tile
            // `ahead` positions further on is (about) the one that starts when this CTA retires; a wrong guess costs nothing
            // but the prefetch.
            const int lin = (int)blockIdx.x + (int)gridDim.x * (int)blockIdx.y + p.prefetch_ahead;
            const int fy = lin / (int)gridDim.x, fx = lin - fy * (int)gridDim.x;
            if (fy < (int)gridDim.y) {
                const PairParams& pf_ = pn_.br[fx % pn_.nbr];
                const int fbx = fx / pn_.nbr;
                const int ft0 = fbx * (R1 - (pf_.k - 1));
                if (ft0 < pf_.T) {
                    const int fA0 = ft0 - (pf_.k - 1) / 2 - pf_.dil * (pf_.k - 1) / 2;
                    const int frows = R1 + (pf_.k - 1) * pf_.dil;
                    const int nline = (frows + 31) / 32 + 1;
                    const float* __restrict__ fb = pf_.x + (size_t)fy * C * pf_.T;
                    for (int i = tid; i < C * nline; i += NWORK) {
                        const int c = i / nline, l = i - c * nline;
                        const int tp = (fA0 & ~3) + 32 * l;
                        if (tp >= 0 && tp < pf_.T) prefetch_l2(fb + (size_t)c * pf_.T + tp);
                    }
                }
            }
        }

        const int q = warp & 3, hsel = warp >> 2;
        const int rib = 32 * q + lane;                 // row inside a 128-row block == TMEM lane
        constexpr int CH = C / 2;                      // channels handled by this warp-half
        constexpr int CG = CH < 16 ? CH : 16;          // columns per tcgen05.ld
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const int cbase = hsel * CH;

        // (2) epilogue 1: mid = conv1 + b1 -> lrelu -> fp16 -> A2 (in place), zero outside [0,T)
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (tid == 0) PAIR_TRACE(2);
#pragma unroll 1
        for (int mb = 0; mb < MB; ++mb) {
            const int row = mb * 128 + rib;
            const int tm = tM0 + row;
            const uint32_t keep = ((tm >= 0) && (tm < p.T)) ? 0xffffffffu : 0u;
            const uint32_t phase = swz_phase(row, G::RB);
#pragma unroll
            for (int cc = 0; cc < CH; cc += 2 * CG) {          // two column groups per TMEM round trip
                uint32_t r0[16], r1[16];
                const int c0 = cbase + cc, c1 = c0 + CG;
                const bool two = (cc + CG) < CH;
                if (CG == 16) { tmem_ld16(tlane + mb * C + c0, r0); if (two) tmem_ld16(tlane + mb * C + c1, r1); }
                else tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r0));
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (g == 1 && !two) break;
                    const int cg0 = g ? c1 : c0;
                    const uint32_t* rr = g ? r1 : r0;
                    uint32_t hq[8];                     // packed fp32x2 arithmetic (bit-identical to the scalar form, half the issue slots)
#pragma unroll
                    for (int j4 = 0; j4 < CG; j4 += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + cg0 + j4);
                        float v0 = __uint_as_float(rr[j4 + 0]), v1 = __uint_as_float(rr[j4 + 1]);
                        float v2 = __uint_as_float(rr[j4 + 2]), v3 = __uint_as_float(rr[j4 + 3]);
                        add2(v0, v1, bq.x, bq.y);
                        add2(v2, v3, bq.z, bq.w);
                        hq[j4 / 2] = lrelu_pack2(v0, v1) & keep;
                        hq[j4 / 2 + 1] = lrelu_pack2(v2, v3) & keep;
                    }
                    uint8_t* prow = sm + (cg0 / G::CPP) * APANEL + row * G::RB;
                    const uint32_t ch0 = (uint32_t)((cg0 % G::CPP) / 8);
                    *reinterpret_cast<uint4*>(prow + ((ch0 ^ phase) << 4)) = make_uint4(hq[0], hq[1], hq[2], hq[3]);
                    if (CG == 16) *reinterpret_cast<uint4*>(prow + (((ch0 + 1) ^ phase) << 4)) = make_uint4(hq[4], hq[5], hq[6], hq[7]);
                }
            }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(bar_a);
        if (tid == 0) PAIR_TRACE(3);

        // (3) epilogue 2: out = alpha*(conv2 + b2 + x) + beta*out_old
        const bool has_beta = p.beta != 0.f;
        // residual loads are software-pipelined one column group ahead; the first group is requested BEFORE waiting for
        // conv2, so its latency hides behind the MMAs (each group used to expose a full global-load round trip)
        const bool add_old = p.beta == 1.f;            // xs accumulation: out += y as a fire-and-forget reduction, out is never read
        const bool pf = !p.red_out && (!has_beta || add_old);
        constexpr int QPB = CH / CG;                   // column groups per 128-row block (per warp half)
        constexpr int NQ = MB * QPB;
        float xa[16], xc[16];
        auto issue_x = [&](int qq, float (&buf)[16]) {
            const int o_ = (qq / QPB) * 128 + rib, t_ = t0 + o_;
            const bool v_ = (o_ < TOUT) && (t_ < p.T);
            const float* __restrict__ xs_ = xb + (v_ ? t_ : 0) + (size_t)(cbase + (qq % QPB) * CG) * p.T;
#pragma unroll
            for (int j = 0; j < CG; ++j) buf[j] = v_ ? __ldg(xs_ + (size_t)j * p.T) : 0.f;
        };
        if (pf) issue_x(0, xa);
        mbar_wait(bar_acc, 1);
        tc_fence_after();
        if (tid == 0) PAIR_TRACE(4);
        if (pf) {
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq) {
                const int mb = qq / QPB, c0 = cbase + (qq % QPB) * CG;
                const int o = mb * 128 + rib;
                const int t = t0 + o;
                const bool valid = (o < TOUT) && (t < p.T);
                float* __restrict__ ot = ob + (valid ? t : 0);
                uint32_t r[16];
                if (CG == 16) tmem_ld16(tlane + mb * C + c0, r);
                else tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r));
                if (qq + 1 < NQ) { if (qq & 1) issue_x(qq + 1, xa); else issue_x(qq + 1, xc); }
                float (&xr)[16] = (qq & 1) ? xc : xa;
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j4 = 0; j4 < CG; j4 += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + C + c0 + j4);
                        float y[4];
                        fma2(y[0], y[1], __uint_as_float(r[j4 + 0]), __uint_as_float(r[j4 + 1]), p.inv, p.inv, bq.x, bq.y);
                        fma2(y[2], y[3], __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]), p.inv, p.inv, bq.z, bq.w);
                        add2(y[0], y[1], xr[j4 + 0], xr[j4 + 1]);
                        add2(y[2], y[3], xr[j4 + 2], xr[j4 + 3]);
                        mul2(y[0], y[1], y[0], y[1], p.alpha, p.alpha);
                        mul2(y[2], y[3], y[2], y[3], p.alpha, p.alpha);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 + e;
                            if (add_old) atomicAdd(ot + (size_t)(c0 + j) * p.T, y[e]);
                            else ot[(size_t)(c0 + j) * p.T] = y[e];
                            xr[j] = y[e];
                        }
                    }
                    if (p.a16_out) {
#pragma unroll
                        for (int j = 0; j < CG; ++j) xr[j] = lrelu01(xr[j]);
                        uint4* dst = reinterpret_cast<uint4*>(p.a16_out + ((size_t)b * p.T + t) * C + c0);
                        dst[0] = make_uint4(pack_h2(xr[0], xr[1]), pack_h2(xr[2], xr[3]), pack_h2(xr[4], xr[5]), pack_h2(xr[6], xr[7]));
                        if (CG == 16) dst[1] = make_uint4(pack_h2(xr[8], xr[9]), pack_h2(xr[10], xr[11]), pack_h2(xr[12], xr[13]), pack_h2(xr[14], xr[15]));
                    }
                }
            }
        }
        if (p.red_out) {
            // out already holds alpha*x (+ old): add alpha*(conv2 + b2) with fire-and-forget reductions
#pragma unroll 1
            for (int mb = 0; mb < MB; ++mb) {
                const int o = mb * 128 + rib;
                const int t = t0 + o;
                const bool valid = (o < TOUT) && (t < p.T);
                float* __restrict__ ot = ob + (valid ? t : 0);
#pragma unroll
                for (int cc = 0; cc < CH; cc += 2 * CG) {          // two column groups per TMEM round trip
                    uint32_t r0[16], r1[16];
                    const int c0 = cbase + cc, c1 = c0 + CG;
                    const bool two = (cc + CG) < CH;
                    if (CG == 16) { tmem_ld16(tlane + mb * C + c0, r0); if (two) tmem_ld16(tlane + mb * C + c1, r1); }
                    else tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r0));
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (g == 1 && !two) break;
                            const int cg0 = g ? c1 : c0;
                            const uint32_t* rr = g ? r1 : r0;
#pragma unroll
                            for (int j4 = 0; j4 < CG; j4 += 4) {
                                const float4 bq = *reinterpret_cast<const float4*>(sbias + C + cg0 + j4);
                                atomicAdd(ot + (size_t)(cg0 + j4 + 0) * p.T, p.alpha * fmaf(__uint_as_float(rr[j4 + 0]), p.inv, bq.x));
                                atomicAdd(ot + (size_t)(cg0 + j4 + 1) * p.T, p.alpha * fmaf(__uint_as_float(rr[j4 + 1]), p.inv, bq.y));
                                atomicAdd(ot + (size_t)(cg0 + j4 + 2) * p.T, p.alpha * fmaf(__uint_as_float(rr[j4 + 2]), p.inv, bq.z));
                                atomicAdd(ot + (size_t)(cg0 + j4 + 3) * p.T, p.alpha * fmaf(__uint_as_float(rr[j4 + 3]), p.inv, bq.w));
                            }
                        }
                    }
                }
            }
        }
#pragma unroll 1
        for (int mb = 0; !p.red_out && !pf && mb < MB; ++mb) {
            const int o = mb * 128 + rib;
            const int t = t0 + o;
            const bool valid = (o < TOUT) && (t < p.T);
            const float* __restrict__ xt = xb + (valid ? t : 0);
            float* __restrict__ ot = ob + (valid ? t : 0);
#pragma unroll
            for (int cc = 0; cc < CH; cc += CG) {
                const int c0 = cbase + cc;
                uint32_t r[16];
                float xr[16], oo[16];
                if (CG == 16) tmem_ld16(tlane + mb * C + c0, r);
                else tmem_ld8(tlane + mb * C + c0, reinterpret_cast<uint32_t(&)[8]>(r));
                if (valid) {
#pragma unroll
                    for (int j = 0; j < CG; ++j) xr[j] = __ldg(xt + (size_t)(c0 + j) * p.T);
                    if (has_beta) {
#pragma unroll
                        for (int j = 0; j < CG; ++j) oo[j] = ot[(size_t)(c0 + j) * p.T];
                    }
                }
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j4 = 0; j4 < CG; j4 += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + C + c0 + j4);
                        const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 + e;
                            float y = p.alpha * (fmaf(__uint_as_float(r[j]), p.inv, bb[e]) + xr[j]);
                            if (has_beta) y = fmaf(p.beta, oo[j], y);
                            ot[(size_t)(c0 + j) * p.T] = y;
                            xr[j] = lrelu01(y);
                        }
                    }
                    if (p.a16_out) {
                        uint4* dst = reinterpret_cast<uint4*>(p.a16_out + ((size_t)b * p.T + t) * C + c0);
                        dst[0] = make_uint4(pack_h2(xr[0], xr[1]), pack_h2(xr[2], xr[3]), pack_h2(xr[4], xr[5]), pack_h2(xr[6], xr[7]));
                        if (CG == 16) dst[1] = make_uint4(pack_h2(xr[8], xr[9]), pack_h2(xr[10], xr[11]), pack_h2(xr[12], xr[13]), pack_h2(xr[14], xr[15]));
                    }
                }
            }
        }
        if (tid == 0) PAIR_TRACE(5);
        tc_fence_before();
    }

    // ---------------------------------------------------------------- teardown
    __syncthreads();
    if (tid == 0) PAIR_TRACE(10);
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s ? std::atoi(s) : dflt;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// fp16 activation copy [B][T][C] viewed as a rank-3 tensor (C innermost); box = one 64-channel panel x box_rows time steps
int make_a16_tmap(CUtensorMap* out, const void* base, int B, int T, int C, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return SVB_ERR_CUDA;
    const int cpp = C < 64 ? C : 64;
    cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)cpp, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = cpp * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cpp * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : SVB_ERR_CUDA;
}

template <int C, int MB, int STAGE_KB, int MINB, bool TMA_IN>
int launch_pair_t2(const PairTC* av, int nbr, cudaStream_t st) {
    const PairTC& a = av[0];
    constexpr size_t smem = pair_smem_bytes<C, MB, STAGE_KB>();
    static_assert(smem * MINB + 1024 * MINB <= 228 * 1024, "pair kernel shared memory exceeds the SM budget");
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (ensure_dyn_smem(pair_tc_kernel<C, MB, STAGE_KB, MINB, TMA_IN>, smem, granted)) return SVB_ERR_CUDA;
    alignas(64) CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (TMA_IN) {
        if (make_a16_tmap(&tmap, a.a16_in, a.B, a.T, C, TileRows<MB>::AROWS / TMA_NBOX) != 0) return SVB_ERR_CUDA;
    }
    PairParamsN pn;
    int gx = 0;
    for (int z = 0; z < 3; ++z) {
    const PairTC& a = av[z < nbr ? z : 0];
    PairParams& p = pn.br[z];
    p.x = a.x; p.out = a.out;
    p.w1 = static_cast<const uint8_t*>(a.w1); p.w2 = static_cast<const uint8_t*>(a.w2);
    p.b1 = a.b1; p.b2 = a.b2; p.T = a.T; p.k = a.k; p.dil = a.dil; p.alpha = a.alpha; p.beta = a.beta; p.inv = a.inv;
    p.a16_out = static_cast<__half*>(a.a16_out);
    {
        static const int env_vec4 = env_int("SVB_PAIR_VEC4", 1), env_red = env_int("SVB_PAIR_RED", 0)   /* measured: L2 reductions cost 0.7 ms/step, off */;
        const bool aligned = (a.T % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0);
        static const int env_pf = env_int("SVB_PAIR_PF", 1);     // measured: 2 (also the slot's next tile) 3.50 vs 3.37 ms/step - 47 MB of
                                                                 // prefetched lines per wave compete with the live tiles for L2
        p.vec4 = (!TMA_IN && env_vec4 && aligned) ? 1 : 0;
        p.prefetch = env_pf;
        p.prefetch_ahead = sm_count() * MINB;
        p.red_out = (p.vec4 && env_red && !a.a16_out && a.x != a.out && (a.beta == 0.f || a.beta == 1.f)) ? 1 : 0;
        // tile period estimate (clk) for the first-wave de-phasing: two MMA phases at the shared-pipe rate + memory phases
        static const int env_dephase = env_int("SVB_PAIR_DEPHASE", -1);
        const int mma_clk = (C >= 128 ? C / 2 : (C == 64 ? 48 : 40)) * a.k * (C / 16) * MB;
        const int grid_ctas = (int)(((a.T + (128 * MB - (a.k - 1)) - 1) / (128 * MB - (a.k - 1))) * a.B);
        p.epoch = z == 0 ? ++g_pair_epoch : pn.br[0].epoch;
        p.dephase_clk = (MINB < 2 || grid_ctas < 4 * 148 * MINB) ? 0 : (env_dephase >= 0 ? env_dephase : 2 * mma_clk + 24000);
    }
    const int TOUT = 128 * MB - (a.k - 1);
    gx = std::max(gx, (a.T + TOUT - 1) / TOUT);
    }
    pn.nbr = nbr;
    dim3 grid(gx * nbr, a.B, 1);
    pair_tc_kernel<C, MB, STAGE_KB, MINB, TMA_IN><<<grid, TC_THREADS, smem, st>>>(pn, tmap);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

template <int C, int MB, int STAGE_KB, int MINB>
int launch_pair_t(const PairTC* av, int nbr, cudaStream_t st) {
    // the TMA-fed variant exists for the tiles whose row count splits into three 8-row-aligned boxes (MB = 2, 4) and C >= 64
    if constexpr ((MB == 2 || MB == 4) && C >= 64) {
        if (av[0].a16_in) return nbr == 1 ? launch_pair_t2<C, MB, STAGE_KB, MINB, true>(av, 1, st) : SVB_ERR_UNSUPPORTED;
    } else {
        if (av[0].a16_in) return SVB_ERR_UNSUPPORTED;
    }
    return launch_pair_t2<C, MB, STAGE_KB, MINB, false>(av, nbr, st);
}

}  // namespace

bool pair_tc_supports_tma(int C, int variant) { (void)variant; return C >= 64; }   // every tile variant for C >= 64 has MB in {2,4}

size_t tc_weight_image_bytes(int C, int k) { return (size_t)C * C * k * 2; }

// w_folded: [Cout][Cin][k] fp32  ->  image [tap][panel][Cout row][swizzled Cin halves]
void tc_pack_weight_image(const float* w, int C, int k, void* dst_host, float scale) {
    const int CPP = C < 64 ? C : 64, NP = C / CPP, RB = CPP * 2, SUB = C * RB;
    uint8_t* dst = static_cast<uint8_t*>(dst_host);
    for (int tap = 0; tap < k; ++tap)
        for (int pn = 0; pn < NP; ++pn) {
            uint8_t* blk = dst + (size_t)(tap * NP + pn) * SUB;
            for (int n = 0; n < C; ++n)
                for (int cc = 0; cc < CPP; ++cc) {
                    const int ci = pn * CPP + cc;
                    const float v = w[((size_t)n * C + ci) * k + tap] * scale;
                    const __half h = __float2half_rn(v);
                    const uint32_t off = tc::swz_offset((uint32_t)n, (uint32_t)(cc / 8), (uint32_t)RB) + (cc % 8) * 2;
                    std::memcpy(blk + off, &h, 2);
                }
        }
}

// Tile variants.  variant 0: one CTA per SM with the largest tile (weights amortised over 128*MB rows);
// variant 1: two CTAs per SM (<= 113 KB smem, <= 256 TMEM columns, <= 102 registers each) so that one CTA's
// load / epilogue phases overlap the other's MMA phases.  C=256 needs all 512 TMEM columns and stays 1 CTA/SM.
int launch_pair_tc(const PairTC& a, cudaStream_t st) { return launch_pair_tc_multi(&a, 1, st); }

// nbr <= 3 pairs with the same C, B, T (different k / dilation / weights / buffers) in one launch
int launch_pair_tc_multi(const PairTC* av, int nbr, cudaStream_t st) {
    if (nbr < 1 || nbr > 3) return SVB_ERR_INVALID_ARG;
    for (int z = 0; z < nbr; ++z) {
        const PairTC& q = av[z];
        if (!(q.k == 3 || q.k == 7 || q.k == 11) || (q.k - 1) * q.dil > 50) return SVB_ERR_UNSUPPORTED;
        if (q.C != av[0].C || q.B != av[0].B || q.T != av[0].T || (nbr > 1 && (q.a16_in || q.a16_out))) return SVB_ERR_INVALID_ARG;
    }
    const PairTC& a = av[0];
    static const int env_variant = env_int("SVB_TC_VARIANT", -1);
    int variant = a.variant >= 0 ? a.variant : env_variant;
    if (variant < 0) variant = 1;     // measured on B200 (profiles/r01/bench_pair_sweep_epi.log): two CTAs/SM win for every C that allows it
    if (variant == 1) {
        switch (a.C) {
            case 16: return launch_pair_t<16, 16, 8, 2>(av, nbr, st);
            case 32: return launch_pair_t<32, 8, 22, 2>(av, nbr, st);
            case 64: return launch_pair_t<64, 4, 16, 2>(av, nbr, st);
            case 128: return launch_pair_t<128, 2, 16, 2>(av, nbr, st);
            case 256: return launch_pair_t<256, 2, 32, 1>(av, nbr, st);
            default: return SVB_ERR_UNSUPPORTED;
        }
    }
    switch (a.C) {
        case 16: return launch_pair_t<16, 16, 32, 1>(av, nbr, st);
        case 32: return launch_pair_t<32, 8, 32, 1>(av, nbr, st);
        case 64: return launch_pair_t<64, 4, 32, 1>(av, nbr, st);
        case 128: return launch_pair_t<128, 4, 32, 1>(av, nbr, st);
        case 256: return launch_pair_t<256, 2, 32, 1>(av, nbr, st);
        default: return SVB_ERR_UNSUPPORTED;
    }
}

}  // namespace svb
