// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: mbarrier, 1-D bulk TMA copies,
// tcgen05 alloc / mma / commit / ld, and shared-memory matrix descriptors.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (same fields CUTLASS' cute/arch/mma_sm100_desc.hpp exposes).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace svb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes or `ns` elapse, instead of
// coming back after the (short) default limit.  ncu on the fused ResBlock kernel: 13.7 % of ALL issued warp instructions were
// the YIELD / TRYWAIT / BRA triples of polling warps, taking issue slots from the epilogue warps that bound the kernel.
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity, uint32_t ns = 2000) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(ns)
            : "memory");
    } while (!ok);
}
// Long waits (a role that idles for a whole phase of the others): poll with a back-off so the spinning warp does not eat the
// issue slots of the warps doing the work (ncu on the SnakeAlias-loader conv: 11 % of all issued instructions were try_wait
// / branch pairs of the idle MMA and producer warps).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns = 256) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 1-D bulk copy global -> shared (TMA engine, completes on an mbarrier) -----------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- 3-D tiled TMA load (tensor map) global -> shared, completes on an mbarrier; coordinates innermost first
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// One lane of a fully converged warp (always the same one).  Issuing tcgen05.mma / commit / bulk copies under
// `if (elect_one())` instead of `if (lane == 0)` lets ptxas keep the operands in uniform registers and emit the
// UTCHMMA stream back to back; with a lane test it wraps EVERY instruction in an ELECT / BRA.U.ANY loop
// (~15 issue slots per MMA, measured ~100 clk per MMA on B200 - longer than an N<=128 MMA itself).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- phase skew of co-resident CTAs ------------------------------------------------------------------
// Two CTAs that start together on one SM run the same phase sequence in lockstep: both issue MMAs at the same time
// (sharing the tensor pipe) and both run their loads / epilogues at the same time (pipe idle) - measured with
// tools/bench_rb.cu.  Sharing the pipe preserves whatever offset the two CTAs have, so the CTA that occupies the "odd"
// slot of its SM delays its first MMA phase by about half a conv cycle, which puts the pair in anti-phase.
// sm_ticket() returns how many CTAs of this launch (identified by `epoch`) started on this SM before the caller; its
// parity identifies the slot as long as the two slots retire alternately.  The counter array needs no reset.
__device__ __forceinline__ uint32_t sm_ticket(unsigned long long* ctr, uint32_t epoch) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* c = ctr + (smid & 255u);
    while (true) {
        const unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(c);
        if ((uint32_t)(old >> 32) == epoch) return (uint32_t)atomicAdd(c, 1ull);
        if (atomicCAS(c, old, ((unsigned long long)epoch << 32) | 1ull) == old) return 0u;
    }
}
__device__ __forceinline__ void spin_clocks(int clk) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)clk) { __nanosleep(64); }
}
// Launch-wide de-phasing.  All CTAs of a wave start together, take the same time per phase and therefore hit HBM together
// (every SM loads its tile at once, then every SM sits in its MMA phase with HBM idle) - measured: the load phase of the
// pair kernel is bandwidth-bound only because of that burst.  The first-wave CTAs (ticket < ctas_per_sm on their SM) wait
// a golden-ratio-hashed fraction of one tile period before starting; later CTAs inherit the offset of the slot they
// replace, so the whole launch runs as a steady flow and the block scheduler balances the tail.
__device__ __forceinline__ void dephase_first_wave(unsigned long long* ctr, uint32_t epoch, int period_clk, uint32_t ctas_per_sm) {
    if (period_clk <= 0) return;
    const uint32_t ticket = sm_ticket(ctr, epoch);
    if (ticket >= ctas_per_sm) return;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const uint32_t h = ((smid * ctas_per_sm + ticket) * 0x9E3779B1u) >> 16;          // 16 well-spread bits
    spin_clocks((int)(((unsigned long long)h * (unsigned long long)period_clk) >> 16));
}

// Read-only global load the compiler may not move across barrier waits (prefetch of the next tile).
__device__ __forceinline__ float ldg_nc_v(const float* ptr) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(ptr));
    return v;
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// ---- tcgen05 -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same MMA with the descriptors given as (low word, high word): only the low word (start address >> 4, LBO) changes from tap
// to tap / K step to K step, so an issue loop can advance 32-bit values and keep the constant high words (SBO, version,
// layout) out of the per-MMA arithmetic (the 64-bit form costs IADD3 + IMAD.X + two R2UR per descriptor and MMA).
__device__ __forceinline__ void umma_f16_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs (of this thread) have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {    // 32 lanes x 16 columns (warp collective)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {   // registers -> 32 lanes x 16 columns
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------
// K-major operand whose rows are `row_bytes` (128/64/32) long and stored densely with the matching
// hardware swizzle (SWIZZLE_128B / 64B / 32B): 16-byte chunk index XOR (address bits [7,10) & mask).
//   start address  bits [0,14)   (>>4)
//   LBO            bits [16,30)  (unused for swizzled K-major; 1 like CUTLASS)
//   SBO            bits [32,46)  (>>4) byte distance between 8-row groups = 8*row_bytes
//   version        bits [46,48)  = 1 on sm_100
//   base offset    bits [49,52)
//   layout type    bits [61,64)  2 = 128B, 4 = 64B, 6 = 32B swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t row_bytes, uint32_t base_offset) {
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7u) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bits[4,6)=1), A=B=f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of 16-byte chunk `chunk` of row `row` inside a swizzled operand whose base is aligned to the
// swizzle repeat (8 rows): the swizzle is a pure function of the address bits.
__host__ __device__ inline uint32_t swz_offset(uint32_t row, uint32_t chunk, uint32_t row_bytes) {
    uint32_t a = row * row_bytes + chunk * 16u;
    uint32_t mask = row_bytes / 16u - 1u;
    return a ^ (((a >> 7) & mask) << 4);
}

// fp32 pair -> packed fp16x2 (a in the low half), round-to-nearest, SATURATING to +-65504: the reference's TF32 operands
// cannot overflow, so an out-of-range activation must not turn the waveform into inf/NaN here either.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float lrelu01(float v) { return fmaxf(v, 0.1f * v); }   // LeakyReLU(0.1): max(v, 0.1 v)

// ---- packed fp32 pairs (FADD2 / FMUL2 / FFMA2): element-wise IEEE round-to-nearest, bit-identical to the scalar
// instructions, one issue slot for two values.  The fused-ResBlock epilogues are bound by instruction issue (~7 thread
// instructions per accumulator element, tools/bench_rbskew.cu), so every halved instruction is time.
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {          // (a0, a1) += (b0, b1)
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {   // a*b + c
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\tfma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
// LeakyReLU(0.1) of a pair, packed to fp16x2 (same arithmetic as pack_h2(lrelu01(a), lrelu01(b)))
__device__ __forceinline__ uint32_t lrelu_pack2(float a, float b) {
    float ma, mb;
    mul2(ma, mb, a, b, 0.1f, 0.1f);
    return pack_h2(fmaxf(a, ma), fmaxf(b, mb));
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Write 8 fp16 values (one 16-byte chunk `chunk` of a swizzled operand row) given the row's base pointer
// (operand base + row*row_bytes) and the row's swizzle phase ((row*row_bytes) >> 7) & mask.  `keep` is all-ones or 0.
__device__ __forceinline__ void store_chunk8(uint8_t* row_ptr, uint32_t phase, int chunk, const float* v, uint32_t keep) {
    uint4 q = make_uint4(pack_h2(v[0], v[1]) & keep, pack_h2(v[2], v[3]) & keep, pack_h2(v[4], v[5]) & keep, pack_h2(v[6], v[7]) & keep);
    *reinterpret_cast<uint4*>(row_ptr + ((((uint32_t)chunk) ^ phase) << 4)) = q;
}
__device__ __forceinline__ uint32_t swz_phase(uint32_t row, uint32_t row_bytes) {
    return ((row * row_bytes) >> 7) & (row_bytes / 16u - 1u);
}

}  // namespace tc
}  // namespace svb
