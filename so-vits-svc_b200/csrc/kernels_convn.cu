// Generic tcgen05 convolution-as-GEMM kernel (sm_100a) for the layers around the ResBlocks:
//   conv_pre (192->512, k7), the polyphase transposed convolutions ups[i] (a stride-s
//   ConvTranspose1d with k = 2s is s interleaved 2-tap channel mixes), and the flow's WN layers
//   (pre 96->192, in_layers 192->384 k5 with the tanh*sigmoid gate fused, res_skip 192->384, post 192->96).
//
// Structure: the activation tile A = act(x)[rows][Cin] (fp16, K-major, hardware swizzle) is staged ONCE per CTA and
// stays resident; the output columns (N_total = C_out, or C_out*s for the polyphase form) are processed in chunks
// of NC <= 256/MB columns with DOUBLE-BUFFERED TMEM accumulators, so the epilogue of chunk j overlaps the MMAs of
// chunk j+1.  Weights stream through a 2-stage mbarrier ring of 1-D bulk TMA copies, like the pair kernel.
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/sovits_b200.h"

#include <cstring>

namespace svb {

using namespace tc;

namespace {

constexpr int CN_THREADS = 320;
constexpr int CN_NWORK = 256;
constexpr int CN_NSTAGE = 2;
constexpr int CN_STAGE_BYTES = 32768;
constexpr int CN_BUFCOLS = 256;      // TMEM columns per accumulator buffer (2 buffers = 512)

template <int CINP>
struct CNGeom {
    static constexpr int CPP = CINP < 64 ? CINP : 64;
    static constexpr int NP = CINP / CPP;
    static constexpr int RB = CPP * 2;
    static constexpr int KSTEPS = CPP / 16;
};

template <int CINP, int MB>
constexpr size_t convn_smem_bytes(int halo_rows_max) {
    return 1024 + (size_t)CNGeom<CINP>::NP * (128 * MB + halo_rows_max) * CNGeom<CINP>::RB + (size_t)CN_NSTAGE * CN_STAGE_BYTES + 256;
}

constexpr int CN_HALO = 8;   // >= max (k-1)*dil over the layers served here (k7 d1 -> 6), multiple of 8

template <int CINP, int MB>
__global__ void __launch_bounds__(CN_THREADS, 1) convn_tc_kernel(const ConvNTC a) {
    using G = CNGeom<CINP>;
    constexpr int R1 = 128 * MB;
    constexpr int AROWS = R1 + CN_HALO;
    constexpr int APANEL = AROWS * G::RB;
    constexpr int BLKCOLS = CN_BUFCOLS / MB;            // TMEM column stride between the MB row-blocks of a buffer

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t a_base = base;
    const uint32_t ring_base = base + G::NP * APANEL;
    const uint32_t bar_base = ring_base + CN_NSTAGE * CN_STAGE_BYTES;
    const uint32_t bar_full = bar_base;                  // [2]
    const uint32_t bar_empty = bar_base + 16;            // [2]
    const uint32_t bar_a = bar_base + 32;                // A tile staged (256 arrivals)
    const uint32_t bar_tfull = bar_base + 40;            // [2] accumulator buffer complete (1 arrival: commit)
    const uint32_t bar_tempty = bar_base + 56;           // [2] accumulator buffer drained (256 arrivals)
    const uint32_t tmem_slot = bar_base + 72;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * R1;                       // first output row of this tile
    const int NC = a.NC;
    const int n_chunks = (a.N_total + NC - 1) / NC;
    const int c_lo = blockIdx.z * a.chunks_per_cta;
    const int c_hi = min(n_chunks, c_lo + a.chunks_per_cta);
    const int SUB = NC * G::RB;                           // bytes of one (tap, panel) weight block
    const int SPC = max(1, CN_STAGE_BYTES / SUB);         // sub-blocks per ring chunk
    const int sb_per_chunk = a.k * G::NP;
    const int RA = R1 + (a.k - 1) * a.dil;                // A rows touched

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, CN_NWORK);
        }
        mbar_init(bar_a, CN_NWORK);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 9) {
        // ------------------------------------------------------------ weight producer
        if (lane == 0) {
            int ring = 0;
            for (int c = c_lo; c < c_hi; ++c) {
                const uint8_t* wsrc = static_cast<const uint8_t*>(a.w) + (size_t)c * sb_per_chunk * SUB;
                for (int sb0 = 0; sb0 < sb_per_chunk; sb0 += SPC, ++ring) {
                    const int s = ring & 1;
                    if (ring >= 2) mbar_wait(bar_empty + 8 * s, ((ring >> 1) - 1) & 1);
                    const int nsb = min(SPC, sb_per_chunk - sb0);
                    const uint32_t bytes = (uint32_t)nsb * SUB;
                    mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                    bulk_g2s(ring_base + s * CN_STAGE_BYTES, wsrc + (size_t)sb0 * SUB, bytes, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, NC);
            mbar_wait(bar_a, 0);
            tc_fence_after();
            int ring = 0;
            for (int c = c_lo; c < c_hi; ++c) {
                const int u = c - c_lo, buf = u & 1;
                if (u >= 2) { mbar_wait(bar_tempty + 8 * buf, ((u >> 1) - 1) & 1); tc_fence_after(); }
                const uint32_t dcol = tmem_base + buf * CN_BUFCOLS;
                for (int sb = 0; sb < sb_per_chunk; ++sb) {
                    const int s = ring & 1;
                    const int within = sb % SPC;
                    if (within == 0) { mbar_wait(bar_full + 8 * s, (ring >> 1) & 1); tc_fence_after(); }
                    const int tap = sb / G::NP, pn = sb % G::NP;
                    const uint32_t a0 = a_base + pn * APANEL + (uint32_t)(tap * a.dil) * G::RB;
                    const uint64_t a_d0 = make_smem_desc(a0, G::RB, 0);
                    const uint64_t b_d0 = make_smem_desc(ring_base + s * CN_STAGE_BYTES + within * SUB, G::RB, 0);
                    const uint32_t acc0 = (sb > 0) ? 1u : 0u;
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                        for (int ks = 0; ks < G::KSTEPS; ++ks) {
                            const uint64_t ad = a_d0 + (uint64_t)(((uint32_t)(mb * 128) * G::RB + ks * 32) >> 4);
                            const uint64_t bd = b_d0 + (uint64_t)((ks * 32) >> 4);
                            umma_f16(dcol + mb * BLKCOLS, ad, bd, idesc, (ks > 0) ? 1u : acc0);
                        }
                    }
                    if (within == SPC - 1 || sb == sb_per_chunk - 1) { umma_commit(bar_empty + 8 * s); ++ring; }
                }
                umma_commit(bar_tfull + 8 * buf);
            }
        }
    } else {
        // ------------------------------------------------------------ workers
        // (1) stage A = act(x)[rows][Cin] as fp16; row r <-> input index i0 - pad_left + r
        {
            const float* __restrict__ xb = a.x + ((size_t)b * a.x_ctot + a.x_c0) * (size_t)a.Tin;
            for (int r = tid; r < RA; r += CN_NWORK) {
                const int ti = i0 - a.pad_left + r;
                const bool rv = (ti >= 0) && (ti < a.Tin);
#pragma unroll 1
                for (int c0 = 0; c0 < CINP; c0 += 16) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = (rv && (c0 + j) < a.cin_real) ? __ldg(xb + (size_t)(c0 + j) * a.Tin + ti) : 0.f;
                    if (a.in_act) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : a.in_slope * v[j];
                    }
                    const int pn = c0 / G::CPP, ch0 = (c0 % G::CPP) / 8;
                    uint8_t* prow = sm + pn * APANEL;
                    *reinterpret_cast<uint4*>(prow + swz_offset(r, ch0, G::RB)) =
                        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    *reinterpret_cast<uint4*>(prow + swz_offset(r, ch0 + 1, G::RB)) =
                        make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_a);
        }
        // (2) per-chunk epilogues
        const int q = warp & 3, hsel = warp >> 2;
        const int rib = 32 * q + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * q) << 16);
        const int len = a.lengths ? a.lengths[b] : 0x7fffffff;
        for (int c = c_lo; c < c_hi; ++c) {
            const int u = c - c_lo, buf = u & 1;
            mbar_wait(bar_tfull + 8 * buf, (u >> 1) & 1);
            tc_fence_after();
            const int col_base = c * NC;
            const int ncols = min(NC, a.N_total - col_base);
#pragma unroll 1
            for (int mb = 0; mb < MB; ++mb) {
                const int i = i0 + mb * 128 + rib;                 // output row
                const uint32_t tcol = tlane + buf * CN_BUFCOLS + mb * BLKCOLS;
                if (a.mode == 2) {
                    // ---- gate: cols [0,NC/2) = tanh pre-activations, [NC/2,NC) = sigmoid pre-activations of the same channels
                    const int hc = NC / 2;
                    const int j_lo = hsel * (hc / 2), j_hi = (hsel + 1) * (hc / 2);
                    for (int j0 = j_lo; j0 < j_hi; j0 += 16) {
                        uint32_t ra[16], rb[16];
                        tmem_ld16(tcol + j0, ra);
                        tmem_ld16(tcol + hc + j0, rb);
                        tmem_ld_wait();
                        if (i < a.n_rows) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int ca = col_base + j0 + j, cb = col_base + hc + j0 + j;
                                float ta = __uint_as_float(ra[j]) + __ldg(a.bias + ca);
                                float sa = __uint_as_float(rb[j]) + __ldg(a.bias + cb);
                                if (a.bias_b) {
                                    ta += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + a.bias_b_off + ca);
                                    sa += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + a.bias_b_off + cb);
                                }
                                const float g = tanhf(ta) * (1.f / (1.f + expf(-sa)));
                                const int ch = c * hc + j0 + j;
                                a.seg[0].y[((size_t)b * a.seg[0].y_ctot + a.seg[0].y_c0 + ch) * (size_t)a.Ty + i] = g;
                            }
                        }
                    }
                } else {
                    const int w_lo = hsel * (ncols / 2), w_hi = (hsel + 1) * (ncols / 2);   // ncols is a multiple of 32
                    for (int j0 = w_lo; j0 < w_hi; j0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tcol + j0, r);
                        tmem_ld_wait();
                        const int col0 = col_base + j0;
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            v[j] = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + col0 + j) : 0.f);
                            if (a.bias_b) v[j] += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + a.bias_b_off + col0 + j);
                        }
                        if (a.mode == 1) {
                            // ---- polyphase: column = co*s + phase; output index n = i*s + phase - p
                            const ConvNSeg& sg = a.seg[0];
                            if (i < a.n_rows) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int col = col0 + j;
                                    const int co = col / a.s, ph = col - co * a.s;
                                    const long long n = (long long)i * a.s + ph - a.p;
                                    if (n >= 0 && n < a.Ty) {
                                        float* dst = sg.y + ((size_t)b * sg.y_ctot + sg.y_c0 + co) * (size_t)a.Ty + n;
                                        float o = sg.alpha * v[j];
                                        if (sg.beta != 0.f) o = fmaf(sg.beta, *dst, o);
                                        *dst = o;
                                    }
                                }
                            }
                        } else {
                            // ---- plain: one column per output channel; up to two destination segments
                            const ConvNSeg& sg = (a.n_seg > 1 && col0 >= a.seg[1].col0) ? a.seg[1] : a.seg[0];
                            if (i < a.n_rows) {
                                float rr[16], oo[16];
                                const size_t rowoff = (size_t)i;
                                if (sg.res) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        rr[j] = __ldg(sg.res + ((size_t)b * sg.res_ctot + sg.res_c0 + (col0 + j - sg.col0)) * (size_t)a.Ty + rowoff);
                                }
                                if (sg.beta != 0.f) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        oo[j] = sg.y[((size_t)b * sg.y_ctot + sg.y_c0 + (col0 + j - sg.col0)) * (size_t)a.Ty + rowoff];
                                }
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    float o = v[j];
                                    if (sg.res) o += rr[j];
                                    o *= sg.alpha;
                                    if (sg.beta != 0.f) o = fmaf(sg.beta, oo[j], o);
                                    if (sg.masked && i >= len) o = 0.f;
                                    sg.y[((size_t)b * sg.y_ctot + sg.y_c0 + (col0 + j - sg.col0)) * (size_t)a.Ty + rowoff] = o;
                                }
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
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int CINP, int MB>
int launch_convn_t(const ConvNTC& a, cudaStream_t st) {
    constexpr size_t smem = convn_smem_bytes<CINP, MB>(CN_HALO);
    static_assert(smem <= 227 * 1024, "convn kernel shared memory exceeds 227 KB");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(convn_tc_kernel<CINP, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return SVB_ERR_CUDA;
        attr_set = true;
    }
    const int n_chunks = (a.N_total + a.NC - 1) / a.NC;
    dim3 grid((a.n_rows + 128 * MB - 1) / (128 * MB), a.B, (n_chunks + a.chunks_per_cta - 1) / a.chunks_per_cta);
    convn_tc_kernel<CINP, MB><<<grid, CN_THREADS, smem, st>>>(a);
    launch_counter()++;
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

}  // namespace

int convn_mb(int cinp) { return cinp >= 512 ? 1 : (cinp == 192 ? 1 : 2); }

int launch_convn_tc(const ConvNTC& a, cudaStream_t st) {
    if ((a.k - 1) * a.dil > CN_HALO || a.NC % 32 || a.NC > 256 / convn_mb(a.cinp) || a.N_total % 32) return SVB_ERR_UNSUPPORTED;
    switch (a.cinp) {
        case 512: return launch_convn_t<512, 1>(a, st);
        case 256: return launch_convn_t<256, 2>(a, st);
        case 192: return launch_convn_t<192, 1>(a, st);
        case 128: return launch_convn_t<128, 2>(a, st);
        case 64: return launch_convn_t<64, 2>(a, st);
        case 32: return launch_convn_t<32, 2>(a, st);
        default: return SVB_ERR_UNSUPPORTED;
    }
}

size_t convn_weight_image_bytes(int cinp, int N_total, int NC, int k) { return (size_t)((N_total + NC - 1) / NC) * NC * cinp * k * 2; }

// wcol(col, ci, tap) -> folded weight value; image layout [chunk][tap][panel][NC rows][swizzled Cin halves]
void convn_pack_weight_image(int cinp, int N_total, int NC, int k, const std::function<float(int, int, int)>& wcol, void* dst_host) {
    const int CPP = cinp < 64 ? cinp : 64, NP = cinp / CPP, RB = CPP * 2;
    const int n_chunks = (N_total + NC - 1) / NC;
    uint8_t* dst = static_cast<uint8_t*>(dst_host);
    const size_t SUB = (size_t)NC * RB;
    for (int c = 0; c < n_chunks; ++c)
        for (int tap = 0; tap < k; ++tap)
            for (int pn = 0; pn < NP; ++pn) {
                uint8_t* blk = dst + ((size_t)(c * k + tap) * NP + pn) * SUB;
                for (int n = 0; n < NC; ++n)
                    for (int cc = 0; cc < CPP; ++cc) {
                        const int col = c * NC + n;
                        const float v = col < N_total ? wcol(col, pn * CPP + cc, tap) : 0.f;
                        const __half h = __float2half_rn(v);
                        const uint32_t off = tc::swz_offset((uint32_t)n, (uint32_t)(cc / 8), (uint32_t)RB) + (cc % 8) * 2;
                        std::memcpy(blk + off, &h, 2);
                    }
            }
}

}  // namespace svb
