// Generic tcgen05 convolution-as-GEMM kernel (sm_100a) for the layers around the ResBlocks:
//   conv_pre (192->512, k7), the polyphase transposed convolutions ups[i] (a stride-s
//   ConvTranspose1d with k = 2s is s interleaved 2-tap channel mixes), and the flow's WN layers
//   (pre 96->192, in_layers 192->384 k5 with the tanh*sigmoid gate fused, res_skip 192->384, post 192->96).
//
// Structure: the activation tile A = act(x)[rows][Cin] (fp16, K-major, hardware swizzle) is staged ONCE per CTA and
// stays resident; the output columns (N_total = C_out, or C_out*s for the polyphase form) are processed in chunks
// of NC columns with (when a CTA owns several chunks) DOUBLE-BUFFERED TMEM accumulators, so the epilogue of chunk j
// overlaps the MMAs of chunk j+1.  Weights stream through a 2-stage mbarrier ring of 1-D bulk TMA copies.
//
// Polyphase mode can also absorb the stage's noise_convs[i] (vdecoder/hifigan/models.py:343-348,380-382): for the
// narrow stages the excitation window of an output row is <= 16 samples, so it is staged as one extra 16-wide K panel
// (fp16, 32-byte swizzled rows) and the strided analysis filter becomes ONE more MMA per chunk against a banded
// (Toeplitz-expanded) weight block - no separate read-modify-write pass over the stage tensor.
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/sovits_b200.h"

#include <cstring>

namespace svb {

using namespace tc;

namespace {

constexpr int CN_THREADS = 320;
constexpr int CN_NWORK = 256;
constexpr int CN_HALO_MIN = 8;        // A-tile rows beyond the 128*MB output rows: max(8, roundup8((k-1)*dil)), <= 56 (k11, d5)
constexpr int CN_HALO_MAX = 56;
constexpr int CN_NOISE_RB = 32;       // noise panel: 16 fp16 per row
constexpr int SNK_RUN = 16;           // SnakeAlias loader: outputs per (channel, run) task
constexpr int SNK_WIN = SNK_RUN + 12; // x samples a run needs: [t0-6, t0+21]

template <int CINP>
struct CNGeom {
    static constexpr int CPP = CINP < 64 ? CINP : 64;
    static constexpr int NP = CINP / CPP;
    static constexpr int RB = CPP * 2;
    static constexpr int KSTEPS = CPP / 16;
    // SnakeAlias loader: channels staged per pass (fp32 scratch [SNK_GCH][pitch]); 16 where the operand tile leaves less room
    static constexpr int SNK_GCH = CINP < 32 ? CINP : ((CINP == 64 || CINP >= 512) ? 16 : 32);
};

struct ConvNDev {                     // launch-time geometry (host computed)
    int stage_bytes, tmem_cols, bufcols, blkcols, nbuf;
    int arows;                        // A-tile rows per panel = 128*MB + halo
    int tout;                         // output rows a CTA produces (128*MB; SnakeAlias loader: 128*MB - (k-1)*dil so that the
                                      // rows it has to activate are exactly 128*MB = 8*MB runs per channel, one per worker warp)
    int xs_pitch; uint32_t off_xs;    // SnakeAlias loader: fp32 staging rows [<=32 channels][xs_pitch] (odd pitch)
    int noise_np;                     // noise panels (0, 1 or 2)
    int noise_rb[2];                  // their row bytes (128 -> 64 samples, 32 -> 16 samples)
    uint32_t off_noise[2], off_ring, off_bar, off_bias;   // byte offsets from the 1024-aligned smem base
};

// ---- SnakeAlias for one (channel, run of SNK_RUN rows) task, everything in registers ---------------------------------
//   u[2a]   = 2 (f1 x[a+2] + f3 x[a+1] + f5 x[a] + f7 x[a-1] + f9 x[a-2] + f11 x[a-3])        (zero-stuffed 2x upsampling,
//   u[2a+1] = 2 (f0 x[a+3] + f2 x[a+2] + f4 x[a+1] + f6 x[a] + f8 x[a-1] + f10 x[a-2])         alias/resample.py:35-54)
//   s[m]    = u[m] + sin^2(e^alpha u[m]) / (e^beta + 1e-9)                                      (alias/act.py SnakeBeta)
//   y[t]    = sum_j f[j] s[clamp(2t + j - 5, 0, 2L-1)]                                          (alias/filter.py:93-109)
// with x replicate-padded (indices clamped when the window is staged).  sin^2(z) = (1 - cos 2z)/2 through MUFU.COS
// (absolute error ~6e-8 |2z|, far inside the fp16 operand rounding that follows).
// xw[j] = x[t0 - 6 + j]; returns y[t0 + i] in y[i].  EDGE: the tile touches t < 3 or t > L - 4, where s indices clamp
// (s0 = s[0], sL = s[2L-1] are supplied) and rows outside [0, L) are zero (the conv's own zero padding).
template <bool EDGE, typename Emit>
__device__ __forceinline__ void snake_run(const float (&xw)[SNK_WIN], const float (&f)[12], float ea2, float hib, int t0, int L,
                                          float s0, float sL, Emit&& emit) {
    // f holds 2 x the filter taps (the gain of the zero-stuffing upsampler folded in, exact); the low-pass therefore returns
    // 2 y, and the factor 1/2 rides on the accumulator scale of the convolution that consumes the operand
    float s[12];
    auto s_val = [&](int jj) -> float {
        // m = 2*t0 - 5 + jj;  jj even -> m odd (a = t0 - 3 + jj/2), jj odd -> m even (a = t0 - 2 + (jj-1)/2)
        float u;
        if ((jj & 1) == 0) {
            const int q = jj / 2 + 3;          // xw index of x[a]
            u = f[0] * xw[q + 3];
            u = fmaf(f[2], xw[q + 2], u); u = fmaf(f[4], xw[q + 1], u); u = fmaf(f[6], xw[q], u);
            u = fmaf(f[8], xw[q - 1], u); u = fmaf(f[10], xw[q - 2], u);
        } else {
            const int q = (jj - 1) / 2 + 4;
            u = f[1] * xw[q + 2];
            u = fmaf(f[3], xw[q + 1], u); u = fmaf(f[5], xw[q], u); u = fmaf(f[7], xw[q - 1], u);
            u = fmaf(f[9], xw[q - 2], u); u = fmaf(f[11], xw[q - 3], u);
        }
        float v = fmaf(hib, 1.f - __cosf(ea2 * u), u);      // u + (1/beta) * (1 - cos(2 e^alpha u)) / 2
        if (EDGE) {
            const int m = 2 * t0 - 5 + jj;
            v = m < 0 ? s0 : (m > 2 * L - 1 ? sL : v);
        }
        return v;
    };
#pragma unroll
    for (int jj = 0; jj < 10; ++jj) s[jj] = s_val(jj);
#pragma unroll
    for (int i = 0; i < SNK_RUN; ++i) {
        s[10] = s_val(2 * i + 10);
        s[11] = s_val(2 * i + 11);
        float acc = f[0] * s[0];
#pragma unroll
        for (int j = 1; j < 12; ++j) acc = fmaf(f[j], s[j], acc);       // = 2 y[t0 + i]: launch_convn_tc halves acc_scale (exact)
        if (EDGE) { const int t = t0 + i; if (t < 0 || t >= L) acc = 0.f; }
        emit(i, acc);
#pragma unroll
        for (int j = 0; j < 10; ++j) s[j] = s[j + 2];
    }
}

// MODE is the epilogue (ConvNTC::mode) as a template parameter: the small latency-bound launches of the prior encoder execute
// every instruction once per CTA, and ncu showed "no instruction" (fetch) stalls second only to scoreboard waits with all four
// epilogues in one body; VIEW = strided-view loader (stage-0 noise conv) compiled in.
template <int CINP, int MB, int MINB, bool SNAKE, int MODE, bool VIEW>
__global__ void __launch_bounds__(CN_THREADS, MINB) convn_tc_kernel(const ConvNTC a, const ConvNDev d) {
    using G = CNGeom<CINP>;
    constexpr int R1 = 128 * MB;
    const int AROWS = d.arows;
    const int APANEL = AROWS * G::RB;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t ring_base = base + d.off_ring;
    const uint32_t bar_base = base + d.off_bar;
    const uint32_t bar_full = bar_base;                  // [2]
    const uint32_t bar_empty = bar_base + 16;            // [2]
    const uint32_t bar_a = bar_base + 32;                // A tile staged (256 arrivals)
    const uint32_t bar_tfull = bar_base + 40;            // [2] accumulator buffer complete
    const uint32_t bar_tempty = bar_base + 56;           // [2] accumulator buffer drained (256 arrivals)
    const uint32_t tmem_slot = bar_base + 72;
    const uint32_t bar_adone = bar_base + 96;            // MMAs of a K chunk complete: the operand tile may be restaged
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
    float* sbias = reinterpret_cast<float*>(sm + d.off_bias);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * d.tout;                   // first output row of this tile
    const int NC = a.NC;
    const int n_chunks = (a.N_total + NC - 1) / NC;
    const int c_lo = blockIdx.z * a.chunks_per_cta;
    const int c_hi = min(n_chunks, c_lo + a.chunks_per_cta);
    const int SUB = NC * G::RB;                           // bytes of one (tap, panel) weight block
    const int n_reg = a.k * G::NP;                        // regular sub-blocks per chunk
    const int n_sb = n_reg + d.noise_np;
    const int SUBN0 = NC * d.noise_rb[0], SUBN1 = NC * d.noise_rb[1];   // bytes of the noise weight blocks
    const size_t chunk_bytes = (size_t)n_reg * SUB + (d.noise_np > 0 ? SUBN0 : 0) + (d.noise_np > 1 ? SUBN1 : 0);
    auto sb_bytes = [&](int idx) -> uint32_t { return idx < n_reg ? (uint32_t)SUB : (idx == n_reg ? (uint32_t)SUBN0 : (uint32_t)SUBN1); };
    const int RA = d.tout + (a.k - 1) * a.dil;            // A rows that feed valid output rows
    const int nkc = a.w_k2 ? 2 : 1;                       // K chunks (operand tiles staged one after the other)

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, CN_NWORK);
        }
        mbar_init(bar_a, CN_NWORK);
        mbar_init(bar_adone, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(tmem_slot, d.tmem_cols); tmem_relinquish(); }
    // per-column bias (+ per-utterance conditioning bias) of this CTA's chunks -> shared memory
    for (int col = c_lo * NC + tid; col < min(c_hi * NC, a.N_total); col += CN_THREADS) {
        float v = a.bias ? __ldg(a.bias + col) : 0.f;
        if (a.bias_b) v += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + a.bias_b_off + col);
        sbias[col - c_lo * NC] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9) {
        // ------------------------------------------------------------ weight producer (whole warp converged, elected lane issues)
        int ring = 0;
        for (int kc = 0; kc < nkc; ++kc)
        for (int c = c_lo; c < c_hi; ++c) {
            const uint8_t* wsrc = static_cast<const uint8_t*>(kc ? a.w_k2 : a.w) + (size_t)c * chunk_bytes;
            int idx = 0;
            uint32_t off = 0;
            while (idx < n_sb) {
                uint32_t bytes = 0;
                while (idx < n_sb) {
                    const uint32_t sbb = sb_bytes(idx);
                    if (bytes && bytes + sbb > (uint32_t)d.stage_bytes) break;
                    bytes += sbb; ++idx;
                }
                const int s = ring & 1;
                if (ring >= 2) { if (SNAKE) mbar_wait_sleep(bar_empty + 8 * s, ((ring >> 1) - 1) & 1); else mbar_wait(bar_empty + 8 * s, ((ring >> 1) - 1) & 1); }
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                    bulk_g2s(ring_base + s * d.stage_bytes, wsrc + off, bytes, bar_full + 8 * s);
                }
                __syncwarp();
                off += bytes; ++ring;
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer
        // Warp converged, ONE elected lane runs the whole issue loop (ring / accumulator-buffer waits included); the
        // body is UTCHMMAs plus descriptor increments only (see pair_tc_kernel for why).
        const uint32_t idesc = make_idesc_f16(128, NC);
        if (SNAKE) mbar_wait_sleep(bar_a, 0); else mbar_wait(bar_a, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint64_t a_step = (uint64_t)((uint32_t)(a.dil * G::RB) >> 4);
            const uint64_t sub16 = (uint64_t)((uint32_t)SUB >> 4);
            uint32_t ring = 0;
            for (int kc = 0; kc < nkc; ++kc)
            for (int c = c_lo; c < c_hi; ++c) {
                if (kc > 0) { mbar_wait(bar_a, kc & 1); tc_fence_after(); }     // operand tile of this K chunk staged
                const int u = c - c_lo, buf = (d.nbuf > 1) ? (u & 1) : 0;
                if (u >= d.nbuf) { mbar_wait(bar_tempty + 8 * buf, ((u / d.nbuf) - 1) & 1); tc_fence_after(); }
                const uint32_t dcol = tmem_base + buf * d.bufcols;
                uint64_t a_tap = make_smem_desc(a_base, G::RB, 0);
                int pn = 0;
                uint32_t acc = kc > 0 ? 1u : 0u;
                int idx = 0;
                while (idx < n_sb) {
                    // same greedy grouping as the producer
                    const int g0 = idx;
                    uint32_t bytes = 0;
                    while (idx < n_sb) {
                        const uint32_t sbb = sb_bytes(idx);
                        if (bytes && bytes + sbb > (uint32_t)d.stage_bytes) break;
                        bytes += sbb; ++idx;
                    }
                    const uint32_t s = ring & 1u;
                    mbar_wait(bar_full + 8 * s, (ring >> 1) & 1u);
                    tc_fence_after();
                    uint64_t bd = make_smem_desc(ring_base + s * d.stage_bytes, G::RB, 0);
                    const int g_reg = (idx < n_reg ? idx : n_reg);
                    for (int sb = g0; sb < g_reg; ++sb) {
                        const uint64_t ad = a_tap + (uint64_t)pn * (uint64_t)(APANEL >> 4);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                            for (int ks = 0; ks < G::KSTEPS; ++ks)
                                umma_f16(dcol + mb * d.blkcols, ad + (uint64_t)(((uint32_t)(mb * 128) * G::RB + ks * 32) >> 4),
                                         bd + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : acc);
                        }
                        acc = 1u;
                        bd += sub16;
                        if (++pn == G::NP) { pn = 0; a_tap += a_step; }
                    }
                    if (MODE == 1 && idx > n_reg) {
                        // excitation (noise_convs) panels: narrower rows, own swizzle width
                        uint32_t boff = (uint32_t)(g_reg > g0 ? (g_reg - g0) : 0) * (uint32_t)SUB;
                        for (int sb = (g0 > n_reg ? g0 : n_reg); sb < idx; ++sb) {
                            const int pnn = sb - n_reg;
                            const uint32_t rbn = (uint32_t)d.noise_rb[pnn];
                            const uint64_t a_d0 = make_smem_desc(base + d.off_noise[pnn], rbn, 0);
                            const uint64_t b_d0 = make_smem_desc(ring_base + s * d.stage_bytes + boff, rbn, 0);
                            for (int mb = 0; mb < MB; ++mb)
                                for (uint32_t ks = 0; ks < rbn / 32u; ++ks)
                                    umma_f16(dcol + mb * d.blkcols, a_d0 + (uint64_t)(((uint32_t)(mb * 128) * rbn + ks * 32u) >> 4),
                                             b_d0 + (uint64_t)((ks * 32u) >> 4), idesc, (ks > 0) ? 1u : acc);
                            acc = 1u;
                            boff += sb_bytes(sb);
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                    ++ring;
                }
                if (kc + 1 < nkc) umma_commit(bar_adone); else umma_commit(bar_tfull + 8 * buf);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ workers
        // (1) stage A = act(x)[rows][Cin] as fp16; row r <-> input index i0 - pad_left + r
        for (int kc = 0; kc < nkc; ++kc) {
        const int xc0 = kc ? a.k2_c0 : a.x_c0;
        if (kc > 0) mbar_wait(bar_adone, (kc - 1) & 1);                     // the MMAs have consumed the previous operand tile
        if constexpr (SNAKE) {
            // SnakeAlias loader.  Per pass of GCH channels: (i) the fp32 window x[c][ti - 6 .. ] of the tile is staged in
            // shared memory with coalesced loads (indices clamped = replicate padding), (ii) every lane takes one channel
            // and a run of SNK_RUN rows: window -> registers, 2x upsample -> snake -> low-pass/decimate in registers,
            // (iii) the fp16 results go to the swizzled operand tile (32 lanes = 32 consecutive channels of one row).
            // The global loads of pass p+1 are issued (into registers) BEFORE the arithmetic of pass p, so their latency
            // hides behind it; the tile has exactly 128*MB activated rows = 8*MB runs, one round of work per warp and pass.
            constexpr int GCH = G::SNK_GCH;
            constexpr int LPR = 32 / GCH;                      // runs handled side by side in one warp (2 when GCH = 16)
            constexpr int CPW = GCH / 8;                       // channels a warp stages per pass
            constexpr int XLC = 128 * MB + 12;                 // staged samples per channel (RA = 128*MB rows)
            constexpr int NQ = (XLC + 31) / 32;                // strides of 32 samples per channel row
            float* xs = reinterpret_cast<float*>(sm + d.off_xs);
            const int XP = d.xs_pitch;
            constexpr int n_runs = 8 * MB;
            const int tlo = i0 - a.pad_left - 6;               // time of staged sample 0
            const int L = a.Tin;
            const bool edge = (tlo + 6 < 3) || (tlo + XLC > L - 4);
            float f[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) f[j] = 2.f * __ldg(a.snake_filt + j);       // 2 x taps, see snake_run
            const float* __restrict__ xb = a.x + ((size_t)b * a.x_ctot + xc0) * (size_t)L;
            const int cl = lane % GCH, rsel = lane / GCH;
            // rows past the activated ones are read by the MMAs of the tile's unused output rows: keep them finite (zero)
            for (int idx = tid; idx < G::NP * (AROWS - n_runs * SNK_RUN) * (G::RB / 16); idx += CN_NWORK) {
                const int per = (AROWS - n_runs * SNK_RUN) * (G::RB / 16);
                const int pn = idx / per, rem = idx % per;
                const int r = n_runs * SNK_RUN + rem / (G::RB / 16), ch = rem % (G::RB / 16);
                *reinterpret_cast<uint4*>(sm + pn * APANEL + swz_offset(r, ch, G::RB)) = make_uint4(0, 0, 0, 0);
            }
            float nxt[CPW][NQ];                                // the next pass's samples, in flight
            // interior tiles (every staged index inside [0, L)) read with constant offsets from one pointer per channel; only
            // the first / last tiles of a sequence pay for the clamps (ncu: the staging was 17 % of the executed instructions)
            const bool ld_inner = (tlo >= 0) && (tlo + 32 * NQ <= L);
            auto prefetch = [&](int cg0) {
#pragma unroll
                for (int cc = 0; cc < CPW; ++cc) {
                    const int c = cg0 + warp + 8 * cc;
                    const bool cv = c < a.cin_real;
                    const float* __restrict__ xc = xb + (size_t)(cv ? c : 0) * L;
                    if (ld_inner) {
                        const float* __restrict__ xl = xc + tlo + lane;
#pragma unroll
                        for (int u = 0; u < NQ; ++u) nxt[cc][u] = cv ? __ldg(xl + 32 * u) : 0.f;
                    } else {
#pragma unroll
                        for (int u = 0; u < NQ; ++u) {
                            const int ti = min(max(tlo + lane + 32 * u, 0), L - 1);
                            nxt[cc][u] = cv ? __ldg(xc + ti) : 0.f;
                        }
                    }
                }
            };
            prefetch(0);
#pragma unroll 1
            for (int cg0 = 0; cg0 < CINP; cg0 += GCH) {
#pragma unroll
                for (int cc = 0; cc < CPW; ++cc)
#pragma unroll
                    for (int u = 0; u < NQ; ++u)
                        if (lane + 32 * u < XLC) xs[(warp + 8 * cc) * XP + lane + 32 * u] = nxt[cc][u];
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (cg0 + GCH < CINP) prefetch(cg0 + GCH);
                const int c = cg0 + cl;
                float ea2 = 2.f * __ldg(a.snake_ealpha + min(c, a.cin_real - 1));
                asm volatile("" : "+f"(ea2));          // keep the product: otherwise every sample pays an extra u + u
                const float hib = 0.5f * __ldg(a.snake_invbeta + min(c, a.cin_real - 1));
                const float* __restrict__ xrow = xs + cl * XP;
                float s0 = 0.f, sL = 0.f;
                if (edge) {
                    // s[0] and s[2L-1] from the clamped signal (only their own tile(s) read them)
                    auto xat = [&](int t) { const int q = min(max(t, 0), L - 1) - tlo; return (q >= 0 && q < XLC) ? xrow[q] : 0.f; };
                    float u0 = f[1] * xat(2) + f[3] * xat(1) + (f[5] + f[7] + f[9] + f[11]) * xat(0);
                    float uL = (f[0] + f[2] + f[4] + f[6]) * xat(L - 1) + f[8] * xat(L - 2) + f[10] * xat(L - 3);
                    s0 = fmaf(hib, 1.f - __cosf(ea2 * u0), u0);
                    sL = fmaf(hib, 1.f - __cosf(ea2 * uL), uL);
                }
                // operand address of (row r0 + i, channel c): the row's swizzle phase depends on i only (r0 is a multiple of 16)
                uint8_t* const pch = sm + (c / G::CPP) * APANEL + (c % 8) * 2;
                const uint32_t chunk = (uint32_t)(c % G::CPP) / 8u;
#pragma unroll 1
                for (int run = warp * LPR + rsel; run < n_runs; run += 8 * LPR) {
                    float xw[SNK_WIN];
#pragma unroll
                    for (int j = 0; j < SNK_WIN; ++j) xw[j] = xrow[run * SNK_RUN + j];
                    const int t0 = tlo + 6 + run * SNK_RUN;
                    uint8_t* const prun = pch + run * (SNK_RUN * G::RB);
                    auto emit = [&](int i, float v) {
                        const uint32_t ph = ((uint32_t)(i * G::RB) >> 7) & (uint32_t)(G::RB / 16 - 1);     // compile-time per i
                        uint16_t hv;
                        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hv) : "f"(v));
                        *reinterpret_cast<uint16_t*>(prun + i * G::RB + ((chunk ^ ph) << 4)) = hv;
                    };
                    if (edge) snake_run<true>(xw, f, ea2, hib, t0, L, s0, sL, emit);
                    else snake_run<false>(xw, f, ea2, hib, t0, L, s0, sL, emit);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        } else if (VIEW && a.view_tstride == CINP && a.view_cstride == 1 && CINP == 64 && (a.view_off & 3) == 0 && (a.view_bstride & 3) == 0 && !a.in_act) {
            // Strided view whose rows tile a CONTIGUOUS range (row r = samples [r*64 + off, +64) of a 1-channel signal: the
            // stage-0 noise_convs as a 2-tap GEMM over 64-sample rows).  Element (r, c) = flat[r*64 + c]: load the range with
            // coalesced 128-bit loads along the flat index and drop each group of 4 samples into its half-chunk of the tile.
            const float* __restrict__ xb = a.x + (size_t)b * a.view_bstride;
            const long long flat0 = (long long)(i0 - a.pad_left) * 64 + a.view_off;
            for (int g = tid; g < RA * 16; g += CN_NWORK) {
                const int r = g >> 4, c = (g & 15) * 4;
                const long long idx = flat0 + (long long)g * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= 0 && idx + 3 < a.view_limit) v = __ldg(reinterpret_cast<const float4*>(xb + idx));
                else {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = (idx + u >= 0 && idx + u < a.view_limit) ? __ldg(xb + idx + u) : 0.f;
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
                const uint32_t phase = swz_phase(r, G::RB);
                uint8_t* pdst = sm + r * G::RB + ((((uint32_t)c >> 3) ^ phase) << 4) + (c & 4) * 2;
                *reinterpret_cast<uint2*>(pdst) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
            }
        } else {
            const bool view = VIEW && a.view_tstride != 0;
            const float* __restrict__ xb = view ? a.x + (size_t)b * a.view_bstride : a.x + ((size_t)b * a.x_ctot + xc0) * (size_t)a.Tin;
            if (!view && kc == 0) {
                // The loader below is a chain of 2-8 dependent global round trips (16-48 loads in flight per thread): request every
                // line of the tile - all K chunks - from DRAM up front, so that the demand loads are L2 hits (a warp covers 32
                // consecutive rows = one 128-byte line per channel; same idea as the pair kernel's tile prefetch)
                const int ti0 = i0 - a.pad_left;
                const int nl = (RA + 31) / 32 + 1;
                for (int kk2 = 0; kk2 < nkc; ++kk2) {
                    const float* __restrict__ xk = a.x + ((size_t)b * a.x_ctot + (kk2 ? a.k2_c0 : a.x_c0)) * (size_t)a.Tin;
                    for (int idx = tid; idx < a.cin_real * nl; idx += CN_NWORK) {
                        const int c = idx / nl, l = idx - c * nl;
                        const int tp = min(max(ti0 + 32 * l, 0), a.Tin - 1);
                        prefetch_l2(xk + (size_t)c * a.Tin + tp);
                    }
                }
            }
            // one-block tiles (MB = 1) have ~130 rows for 256 loader threads: two threads share a row (half of the channels
            // each) and keep 48 loads in flight, so the tile costs 2-4 global round trips instead of 6-12 (the small GEMMs of
            // enc_p are a chain of such latencies: ~31 us per launch before, of which the loader was about a third)
            constexpr int NSPLIT = (MB == 1 && (CINP % 32) == 0 && CINP >= 192) ? 2 : 1;
            constexpr int RP = CN_NWORK / NSPLIT, CSPAN = CINP / NSPLIT;
            const int cbeg = (tid / RP) * CSPAN;
            for (int r = tid % RP; r < RA; r += RP) {
                const int ti = i0 - a.pad_left + r;
                const bool rv = (ti >= 0) && (ti < a.Tin);
                const float* __restrict__ xt = xb + (rv && !view ? ti : 0);
                const long long vbase = (long long)ti * a.view_tstride + a.view_off;
                const uint32_t phase = swz_phase(r, G::RB);
#pragma unroll NSPLIT == 2 ? 3 : 2
                for (int c0 = cbeg; c0 < cbeg + CSPAN; c0 += 16) {
                    float v[16];
                    if (view) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const long long idx = vbase + (long long)(c0 + j) * a.view_cstride;
                            v[j] = (rv && (c0 + j) < a.cin_real && idx >= 0 && idx < a.view_limit) ? __ldg(xb + idx) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = (rv && (c0 + j) < a.cin_real) ? __ldg(xt + (size_t)(c0 + j) * a.Tin) : 0.f;
                    }
                    if (a.in_act) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], a.in_slope * v[j]);
                    }
                    uint8_t* prow = sm + (c0 / G::CPP) * APANEL + r * G::RB;
                    store_chunk8(prow, phase, (c0 % G::CPP) / 8, v, 0xffffffffu);
                    store_chunk8(prow, phase, (c0 % G::CPP) / 8 + 1, v + 8, 0xffffffffu);
                }
            }
        }
        {
            if (MODE == 1 && d.noise_np) {
                // excitation window of output row i: har[i*noise_stride + noise_w0 + u]; panel 0 holds u in [0, rb0/2),
                // panel 1 the next rb1/2 samples
                const float* __restrict__ hb = a.har + (size_t)b * a.har_N;
                int u0 = 0;
                for (int pnn = 0; pnn < d.noise_np; ++pnn) {
                    const int rbn = d.noise_rb[pnn];
                    const int groups = rbn / 32;                       // 16-sample groups per row
                    for (int it = tid; it < R1 * groups; it += CN_NWORK) {
                        const int r = it / groups, gq = it % groups;
                        const long long h0 = (long long)(i0 + r) * a.noise_stride + a.noise_w0 + u0 + 16 * gq;
                        float v[16];
#pragma unroll
                        for (int uu = 0; uu < 16; ++uu) {
                            const long long hi = h0 + uu;
                            v[uu] = (hi >= 0 && hi < a.har_N) ? __ldg(hb + hi) : 0.f;
                            if (SNAKE) v[uu] *= 2.f;           // same factor as the 2 x activation the SnakeAlias loader stages
                        }
                        uint8_t* prow = sm + d.off_noise[pnn] + r * rbn;
                        const uint32_t phase = swz_phase(r, rbn);
                        store_chunk8(prow, phase, 2 * gq, v, 0xffffffffu);
                        store_chunk8(prow, phase, 2 * gq + 1, v + 8, 0xffffffffu);
                    }
                    u0 += rbn / 2;
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_a);
        }
        }   // K chunks
        // (2) per-chunk epilogues
        const int q = warp & 3, hsel = warp >> 2;
        const int rib = 32 * q + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const int len = a.lengths ? a.lengths[b] : 0x7fffffff;
        for (int c = c_lo; c < c_hi; ++c) {
            const int u = c - c_lo, buf = (d.nbuf > 1) ? (u & 1) : 0;
            if (SNAKE) mbar_wait_sleep(bar_tfull + 8 * buf, (u / d.nbuf) & 1, 128); else mbar_wait(bar_tfull + 8 * buf, (u / d.nbuf) & 1);
            tc_fence_after();
            const int col_base = c * NC;
            const int ncols = min(NC, a.N_total - col_base);
            const float* __restrict__ cb_ = sbias + (col_base - c_lo * NC);
#pragma unroll 1
            for (int mb = 0; mb < MB; ++mb) {
                const int i = i0 + mb * 128 + rib;                 // output row
                const bool rowok = (i < a.n_rows) && (mb * 128 + rib < d.tout);
                const uint32_t tcol = tlane + buf * d.bufcols + mb * d.blkcols;
                if constexpr (MODE == 3) {
                    // ---- attention operand images (dk = 96: panel 0 = channels 0..63, panel 1 = 64..95 in 128-byte rows)
                    const int w_lo = hsel * (ncols / 2), w_hi = (hsel + 1) * (ncols / 2);      // ncols = heads*96, a multiple of 32
                    const int tile = i >> 7, trow = i & 127;
#pragma unroll 1
                    for (int j0 = w_lo; j0 < w_hi; j0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tcol + j0, r);
                        tmem_ld_wait();
                        if (!rowok) continue;
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r[j]), a.acc_scale, cb_[j0 + j]);
                        const int col0 = col_base + j0;                       // global column: [q | k | v] blocks of heads*96
                        const int hd96 = a.att_heads * 96;
                        const int which = col0 / hd96, hc = col0 % hd96, head = hc / 96, ch = hc % 96;     // 16 | 96: a group never straddles
                        const size_t bh = ((size_t)b * a.att_heads + head) * a.att_tiles + tile;
                        if (which < 2) {
                            uint8_t* img = static_cast<uint8_t*>(which == 0 ? a.att_q : a.att_k) + bh * (size_t)(2 * 128 * 128);
                            uint8_t* prow = img + (ch / 64) * (128 * 128) + trow * 128;
                            const uint32_t phase = (uint32_t)trow & 7u;
                            store_chunk8(prow, phase, (ch % 64) / 8, v, 0xffffffffu);
                            store_chunk8(prow, phase, (ch % 64) / 8 + 1, v + 8, 0xffffffffu);
                        } else {
                            uint8_t* img = static_cast<uint8_t*>(a.att_v) + bh * (size_t)(2 * 96 * 128);
                            const int kp = trow >> 6, kc = (trow & 63) >> 3, ko = (trow & 7) * 2;       // key panel, 8-key chunk, byte in chunk
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int c = ch + j;
                                const __half hv = __float2half_rn(fminf(fmaxf(v[j], -65504.f), 65504.f));
                                *reinterpret_cast<__half*>(img + kp * (96 * 128) + c * 128 + (((uint32_t)kc ^ ((uint32_t)c & 7u)) << 4) + ko) = hv;
                            }
                        }
                    }
                } else if constexpr (MODE == 2) {
                    // ---- gate: cols [0,NC/2) = tanh pre-activations, [NC/2,NC) = sigmoid pre-activations of the same channels
                    const int hc = NC / 2;
                    const int j_lo = hsel * (hc / 2), j_hi = (hsel + 1) * (hc / 2);
                    float* __restrict__ yb = a.seg[0].y + ((size_t)b * a.seg[0].y_ctot + a.seg[0].y_c0 + c * hc) * (size_t)a.Ty + (rowok ? i : 0);
                    // speaker-mix conditioning: gcond[b, bias_t_c0 + column, row] in the same (permuted) column order as the image
                    const float* __restrict__ bt = a.bias_t ? a.bias_t + ((size_t)b * a.bias_t_ctot + a.bias_t_c0 + col_base) * (size_t)a.Ty + (rowok ? i : 0) : nullptr;
#pragma unroll 1
                    for (int j0 = j_lo; j0 < j_hi; j0 += 16) {
                        uint32_t ra[16], rb[16];
                        tmem_ld16(tcol + j0, ra);
                        tmem_ld16(tcol + hc + j0, rb);
                        tmem_ld_wait();
                        if (rowok) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float ta = fmaf(__uint_as_float(ra[j]), a.acc_scale, cb_[j0 + j]);
                                float sa = fmaf(__uint_as_float(rb[j]), a.acc_scale, cb_[hc + j0 + j]);
                                if (bt) { ta += __ldg(bt + (size_t)(j0 + j) * a.Ty); sa += __ldg(bt + (size_t)(hc + j0 + j) * a.Ty); }
                                yb[(size_t)(j0 + j) * a.Ty] = tanhf(ta) * (1.f / (1.f + expf(-sa)));
                            }
                        }
                    }
                } else if constexpr (MODE == 1) {
                    // ---- polyphase: column = co*s + phase; output index n = i*s + phase - p (s is 2 or 8)
                    const ConvNSeg& sg = a.seg[0];
                    const int n16 = ncols / 16;
                    const int w_lo = hsel * ((n16 + 1) / 2) * 16, w_hi = hsel ? ncols : ((n16 + 1) / 2) * 16;
                    const long long n0 = (long long)i * a.s - a.p;             // output index of phase 0
                    const bool inner = rowok && n0 >= 0 && n0 + a.s <= a.Ty;
#pragma unroll 1
                    for (int j0 = w_lo; j0 < w_hi; j0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tcol + j0, r);
                        tmem_ld_wait();
                        if (!rowok) continue;
                        float v[16];
#pragma unroll
                        for (int j4 = 0; j4 < 16; j4 += 4) {
                            const float4 bq = *reinterpret_cast<const float4*>(cb_ + j0 + j4);
                            v[j4 + 0] = sg.alpha * fmaf(__uint_as_float(r[j4 + 0]), a.acc_scale, bq.x);
                            v[j4 + 1] = sg.alpha * fmaf(__uint_as_float(r[j4 + 1]), a.acc_scale, bq.y);
                            v[j4 + 2] = sg.alpha * fmaf(__uint_as_float(r[j4 + 2]), a.acc_scale, bq.z);
                            v[j4 + 3] = sg.alpha * fmaf(__uint_as_float(r[j4 + 3]), a.acc_scale, bq.w);
                        }
                        const int col0 = col_base + j0;
                        if (a.s == 8) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float* dst = sg.y + ((size_t)b * sg.y_ctot + sg.y_c0 + (col0 >> 3) + e) * (size_t)a.Ty + n0;
                                if (inner && sg.beta == 0.f) {
                                    *reinterpret_cast<float4*>(dst) = make_float4(v[8 * e], v[8 * e + 1], v[8 * e + 2], v[8 * e + 3]);
                                    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[8 * e + 4], v[8 * e + 5], v[8 * e + 6], v[8 * e + 7]);
                                } else {
#pragma unroll
                                    for (int ph = 0; ph < 8; ++ph) {
                                        const long long n = n0 + ph;
                                        if (n >= 0 && n < a.Ty) dst[ph] = (sg.beta != 0.f) ? fmaf(sg.beta, dst[ph], v[8 * e + ph]) : v[8 * e + ph];
                                    }
                                }
                            }
                        } else {   // s == 2
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float* dst = sg.y + ((size_t)b * sg.y_ctot + sg.y_c0 + (col0 >> 1) + e) * (size_t)a.Ty + n0;
#pragma unroll
                                for (int ph = 0; ph < 2; ++ph) {
                                    const long long n = n0 + ph;
                                    if (inner || (n >= 0 && n < a.Ty)) dst[ph] = (sg.beta != 0.f) ? fmaf(sg.beta, dst[ph], v[2 * e + ph]) : v[2 * e + ph];
                                }
                            }
                        }
                    }
                } else {
                    // ---- plain: one column per output channel; up to two destination segments
                    const int n16 = ncols / 16;                                            // ncols is a multiple of 16
                    const int w_lo = hsel * ((n16 + 1) / 2) * 16, w_hi = hsel ? ncols : ((n16 + 1) / 2) * 16;
                    // the residual of column group j0+16 is requested while group j0 is processed (one exposed global round
                    // trip per row block instead of one per 16 columns)
                    float rnext[16];
                    auto issue_res = [&](int j0) {
                        const int col0 = col_base + j0;
                        const ConvNSeg& sg = (a.n_seg > 1 && col0 >= a.seg[1].col0) ? a.seg[1] : a.seg[0];
                        if (rowok && sg.res && j0 < w_hi) {
                            const float* __restrict__ rb_ = sg.res + ((size_t)b * sg.res_ctot + sg.res_c0 + (col0 - sg.col0)) * (size_t)a.Ty + i;
#pragma unroll
                            for (int j = 0; j < 16; ++j) rnext[j] = __ldg(rb_ + (size_t)j * a.Ty);
                        }
                    };
                    issue_res(w_lo);
#pragma unroll 1
                    for (int j0 = w_lo; j0 < w_hi; j0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tcol + j0, r);
                        const int col0 = col_base + j0;
                        const ConvNSeg& sg = (a.n_seg > 1 && col0 >= a.seg[1].col0) ? a.seg[1] : a.seg[0];
                        float rr[16], oo[16];
                        const float* __restrict__ bt = (a.bias_t && rowok) ? a.bias_t + ((size_t)b * a.bias_t_ctot + a.bias_t_c0 + col0) * (size_t)a.Ty + i : nullptr;
                        float* __restrict__ yb = sg.y + ((size_t)b * sg.y_ctot + sg.y_c0 + (col0 - sg.col0)) * (size_t)a.Ty + (rowok ? i : 0);
#pragma unroll
                        for (int j = 0; j < 16; ++j) rr[j] = rnext[j];
                        issue_res(j0 + 16);
                        if (rowok && sg.beta != 0.f) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) oo[j] = yb[(size_t)j * a.Ty];
                        }
                        tmem_ld_wait();
                        if (rowok) {
                            const bool dead = sg.masked && i >= len;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float o = fmaf(__uint_as_float(r[j]), a.acc_scale, cb_[j0 + j]);
                                if (bt) o += __ldg(bt + (size_t)j * a.Ty);
                                if (sg.res) o += rr[j];
                                o *= sg.alpha;
                                if (sg.beta != 0.f) o = fmaf(sg.beta, oo[j], o);
                                if (a.out_relu) o = fmaxf(o, 0.f);
                                yb[(size_t)j * a.Ty] = dead ? 0.f : o;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * buf);
        }
    }

    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, d.tmem_cols); }
}

int pow2ceil(int v) { int p = 32; while (p < v) p <<= 1; return p; }

template <int CINP, int MB, int MINB, bool SNAKE, int MODE, bool VIEW>
int launch_convn_m(const ConvNTC& a, cudaStream_t st) {
    using G = CNGeom<CINP>;
    int halo = ((a.k - 1) * a.dil + 7) & ~7;
    if (halo < CN_HALO_MIN) halo = CN_HALO_MIN;
    if (halo > CN_HALO_MAX) return SVB_ERR_UNSUPPORTED;
    const int n_chunks = (a.N_total + a.NC - 1) / a.NC;
    const int cpc = a.chunks_per_cta < 1 ? 1 : a.chunks_per_cta;
    ConvNDev d;
    d.noise_np = a.har ? (a.noise_wide ? 2 : 1) : 0;
    d.noise_rb[0] = a.noise_wide ? 128 : CN_NOISE_RB;
    d.noise_rb[1] = CN_NOISE_RB;
    d.nbuf = cpc > 1 ? 2 : 1;
    d.blkcols = pow2ceil(a.NC);
    d.bufcols = MB * d.blkcols;
    d.tmem_cols = pow2ceil(d.nbuf * d.bufcols);
    if (d.tmem_cols > 512) return SVB_ERR_UNSUPPORTED;
    const int SUB = a.NC * G::RB;
    int stage = SUB > 16384 ? 32768 : 16384;
    if (SUB > stage) return SVB_ERR_UNSUPPORTED;
    d.stage_bytes = stage;
    d.arows = 128 * MB + halo;
    d.tout = SNAKE ? 128 * MB - (a.k - 1) * a.dil : 128 * MB;
    if (d.tout < 64) return SVB_ERR_UNSUPPORTED;
    uint32_t off = (uint32_t)G::NP * d.arows * G::RB;
    off = (off + 1023u) & ~1023u;
    d.xs_pitch = 0; d.off_xs = 0;
    if (SNAKE) {
        const int xl = 128 * MB + 12;                          // RA = 128*MB rows -> 8*MB runs of 16
        d.xs_pitch = xl | 1;                                   // odd pitch: lanes (channels) hit distinct banks
        d.off_xs = off;
        off += (uint32_t)G::SNK_GCH * d.xs_pitch * 4;
        off = (off + 1023u) & ~1023u;
    }
    for (int pnn = 0; pnn < 2; ++pnn) {
        d.off_noise[pnn] = off;
        if (pnn < d.noise_np) off += 128 * MB * d.noise_rb[pnn];
        off = (off + 1023u) & ~1023u;
    }
    d.off_ring = off;
    off += 2 * stage;
    d.off_bar = off;
    off += 256;
    d.off_bias = off;
    off += (uint32_t)cpc * a.NC * 4;
    const size_t smem = 1024 + off;
    if (smem > 227 * 1024) return SVB_ERR_UNSUPPORTED;
    static std::atomic<size_t> granted[SVB_MAX_DEV];
    if (ensure_dyn_smem(convn_tc_kernel<CINP, MB, MINB, SNAKE, MODE, VIEW>, smem, granted)) return SVB_ERR_CUDA;
    dim3 grid((a.n_rows + d.tout - 1) / d.tout, a.B, (n_chunks + cpc - 1) / cpc);
    convn_tc_kernel<CINP, MB, MINB, SNAKE, MODE, VIEW><<<grid, CN_THREADS, smem, st>>>(a, d);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

// which epilogues exist per operand width: plain everywhere; polyphase for the upsamplers (512..32); gate (unfused flow) and
// attention images (q/k/v projection) for the 192-channel layers; the strided view only for the 64-channel stage-0 noise conv
template <int CINP, int MB, int MINB, bool SNAKE>
int launch_convn_t(const ConvNTC& a, cudaStream_t st) {
    constexpr bool POLY = (CINP == 512 || CINP == 256 || CINP == 128 || CINP == 64 || CINP == 32);
    if (a.mode == 0) {
        if constexpr (CINP == 64 && !SNAKE) { if (a.view_tstride != 0) return launch_convn_m<CINP, MB, MINB, SNAKE, 0, true>(a, st); }
        if (a.view_tstride != 0) return SVB_ERR_UNSUPPORTED;
        return launch_convn_m<CINP, MB, MINB, SNAKE, 0, false>(a, st);
    }
    if (a.view_tstride != 0) return SVB_ERR_UNSUPPORTED;
    if (a.mode == 1) { if constexpr (POLY) return launch_convn_m<CINP, MB, MINB, SNAKE, 1, false>(a, st); }
    if (a.mode == 2) { if constexpr (CINP == 192 && !SNAKE) return launch_convn_m<CINP, MB, MINB, SNAKE, 2, false>(a, st); }
    if (a.mode == 3) { if constexpr (CINP == 192 && !SNAKE) return launch_convn_m<CINP, MB, MINB, SNAKE, 3, false>(a, st); }
    return SVB_ERR_UNSUPPORTED;
}

}  // namespace

// experiment (SVB_UPS_MB4=1): 512-row tiles for the last two upsamplers, whose 256-row CTAs move only 66 KB each.  Measured:
// ups_tc 0.874 vs 0.852 ms/step with 256-row tiles - the per-CTA set-up is not what bounds them; default off
static int narrow_mb() { static const int v = [] { const char* e = std::getenv("SVB_UPS_MB4"); return (e ? std::atoi(e) : 0) ? 4 : 2; }(); return v; }
int convn_mb(int cinp) { return cinp >= 384 ? 1 : (cinp == 192 ? 1 : 2); }
int convn_ups_mb(int cinp) { return cinp <= 64 ? narrow_mb() : convn_mb(cinp); }      // tiles of the polyphase (mode 1) launches
// SnakeAlias-loader tiles: one 128-row block for the wide stages so that two CTAs share an SM (one CTA's CUDA-core loader
// phase overlaps the other's MMA / epilogue phases); C = 256/512 need the whole shared memory for the operand tile.
int convn_snake_mb(int cinp) { return cinp >= 128 ? 1 : 2; }

int launch_convn_tc(const ConvNTC& a, cudaStream_t st) {
    const bool snake = a.snake_ealpha != nullptr;
    const int mb = snake ? convn_snake_mb(a.cinp) : (a.mode == 1 ? convn_ups_mb(a.cinp) : convn_mb(a.cinp));
    if (a.NC % 16 || a.NC > 256 / mb || a.N_total % 16) return SVB_ERR_UNSUPPORTED;
    if (a.mode == 1 && !(a.s == 2 || a.s == 8)) return SVB_ERR_UNSUPPORTED;
    if (a.mode == 2 && (a.NC % 64)) return SVB_ERR_UNSUPPORTED;
    if (a.mode == 3 && (!a.att_q || !a.att_k || !a.att_v || a.att_heads < 1 || a.NC != a.att_heads * 96 || a.N_total != 3 * a.NC)) return SVB_ERR_INVALID_ARG;
    if (a.w_k2 && (a.N_total > a.NC || a.har || snake)) return SVB_ERR_UNSUPPORTED;      // K chunking: one column chunk, plain loader
    if (snake) {
        if (a.view_tstride != 0 || a.in_act || !a.snake_invbeta || !a.snake_filt) return SVB_ERR_INVALID_ARG;
        ConvNTC h = a;
        h.acc_scale = 0.5f * a.acc_scale;        // the SnakeAlias loader stages 2 x the activation (snake_run)
        switch (a.cinp) {
            case 512: return launch_convn_t<512, 1, 1, true>(h, st);
            case 256: return launch_convn_t<256, 1, 1, true>(h, st);
            case 128: return launch_convn_t<128, 1, 2, true>(h, st);
            case 64: return launch_convn_t<64, 2, 2, true>(h, st);
            case 32: return launch_convn_t<32, 2, 2, true>(h, st);
            case 16: return launch_convn_t<16, 2, 2, true>(h, st);
            default: return SVB_ERR_UNSUPPORTED;
        }
    }
    switch (a.cinp) {
        case 512: return launch_convn_t<512, 1, 1, false>(a, st);
        case 384: return launch_convn_t<384, 1, 1, false>(a, st);
        case 256: return launch_convn_t<256, 2, 1, false>(a, st);
        case 192: return launch_convn_t<192, 1, 1, false>(a, st);
        case 128: return launch_convn_t<128, 2, 2, false>(a, st);
        case 64: return mb == 4 ? launch_convn_t<64, 4, 2, false>(a, st) : launch_convn_t<64, 2, 2, false>(a, st);
        case 32: return mb == 4 ? launch_convn_t<32, 4, 2, false>(a, st) : launch_convn_t<32, 2, 2, false>(a, st);
        default: return SVB_ERR_UNSUPPORTED;
    }
}

// noise: 0 none, 1 one 16-sample panel, 2 a 64-sample + a 16-sample panel (window <= 80)
size_t convn_weight_image_bytes(int cinp, int N_total, int NC, int k, int noise) {
    const size_t nb = noise == 0 ? 0 : (noise == 1 ? (size_t)NC * CN_NOISE_RB : (size_t)NC * (128 + CN_NOISE_RB));
    return (size_t)((N_total + NC - 1) / NC) * ((size_t)NC * cinp * k * 2 + nb);
}

// wcol(col, ci, tap) -> folded weight value; ncol(col, u) -> banded noise weight (u in [0,16)) or null.
// Image layout per chunk: [tap][panel][NC rows][swizzled Cin halves] then (optionally) [NC rows][16 halves, 32-byte rows].
void convn_pack_weight_image(int cinp, int N_total, int NC, int k, const std::function<float(int, int, int)>& wcol,
                             const std::function<float(int, int)>* ncol, int noise, void* dst_host) {
    const int CPP = cinp < 64 ? cinp : 64, NP = cinp / CPP, RB = CPP * 2;
    const int n_chunks = (N_total + NC - 1) / NC;
    uint8_t* dst = static_cast<uint8_t*>(dst_host);
    const size_t SUB = (size_t)NC * RB;
    const int nrb[2] = {noise == 2 ? 128 : CN_NOISE_RB, CN_NOISE_RB};
    const int nnp = ncol ? (noise == 2 ? 2 : 1) : 0;
    const size_t chunk_bytes = (size_t)k * NP * SUB + (nnp > 0 ? (size_t)NC * nrb[0] : 0) + (nnp > 1 ? (size_t)NC * nrb[1] : 0);
    for (int c = 0; c < n_chunks; ++c) {
        uint8_t* cbase = dst + (size_t)c * chunk_bytes;
        for (int tap = 0; tap < k; ++tap)
            for (int pn = 0; pn < NP; ++pn) {
                uint8_t* blk = cbase + ((size_t)tap * NP + pn) * SUB;
                for (int n = 0; n < NC; ++n)
                    for (int cc = 0; cc < CPP; ++cc) {
                        const int col = c * NC + n;
                        const float v = col < N_total ? wcol(col, pn * CPP + cc, tap) : 0.f;
                        const __half h = __float2half_rn(v);
                        const uint32_t off = tc::swz_offset((uint32_t)n, (uint32_t)(cc / 8), (uint32_t)RB) + (cc % 8) * 2;
                        std::memcpy(blk + off, &h, 2);
                    }
            }
        uint8_t* blk = cbase + (size_t)k * NP * SUB;
        int u0 = 0;
        for (int pnn = 0; pnn < nnp; ++pnn) {
            const int rbn = nrb[pnn], width = rbn / 2;
            for (int n = 0; n < NC; ++n)
                for (int uu = 0; uu < width; ++uu) {
                    const int col = c * NC + n;
                    const float v = col < N_total ? (*ncol)(col, u0 + uu) : 0.f;
                    const __half h = __float2half_rn(v);
                    const uint32_t off = tc::swz_offset((uint32_t)n, (uint32_t)(uu / 8), (uint32_t)rbn) + (uu % 8) * 2;
                    std::memcpy(blk + off, &h, 2);
                }
            blk += (size_t)NC * rbn;
            u0 += width;
        }
    }
}

}  // namespace svb
