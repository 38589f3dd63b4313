// sm_100a primitives for the tensor-core feature branch: tcgen05 (MMA, TMEM alloc / load, commit), mbarrier,
// bulk async copies and the shared-memory matrix / instruction descriptors.  Inline PTX only; no CUTLASS.
//
// Operand layout used by the kernels: rows of 64 bf16 channels (128 B) per pixel position with the 128-byte swizzle --
// exactly what a TMA tiled load of a channels-last bf16 tensor produces.
//  * read K-major  (rows = positions, K = channels): forward / data-gradient GEMMs; a convolution tap is a change of the
//    start address by whole rows (the swizzle is a function of the absolute address, so no re-layout is needed);
//  * read MN-major (rows = channels,  K = positions): weight-gradient GEMM, same tile, no transpose.
// (The no-swizzle "plane" layout of make_desc() was the first candidate; tools/microbench/umma_probe*.cu compare them.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (64 bit) ----------------------------------------------------------------------
// [0,14) start address >> 4 | [16,30) leading-dimension byte offset >> 4 | [32,46) stride-dimension byte offset >> 4 |
// [46,48) version = 1 (sm_100) | [49,52) base offset = 0 | [61,64) layout type: 0 = no swizzle.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// 128-byte swizzle ([row][64 bf16] rows of 128 B, 8-row atoms of 1024 B; the XOR is a function of the ABSOLUTE shared
// address: 16-byte chunk index ^= (address >> 7) & 7, so a start address shifted by whole rows needs no base offset --
// measured, tools/microbench/umma_probe2.cu).  K-major: SBO = 1024 (8-row groups), LBO unused (1).  MN-major: LBO = byte
// distance between 64-element atoms along M/N, SBO = 1024 (8 K-positions).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_desc(saddr, lbo_bytes, sbo_bytes) | (2ull << 61);
}
// Only the start address changes between the MMAs of a tile: patch the low 14 bits.
__device__ __forceinline__ uint64_t desc_at(uint64_t desc_hi_bits, uint32_t saddr) {
    return desc_hi_bits | (uint64_t)((saddr >> 4) & 0x3FFF);
}

// ---- instruction descriptor (32 bit), kind::f16 with bf16 operands and fp32 accumulation -----------------------------
// [4,6) D format 1 = f32 | [7,10) A format 1 = bf16 | [10,13) B format 1 = bf16 | [15] A major (0 = K, 1 = MN) |
// [16] B major | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Operand formats of kind::f16 are per operand: format code 0 = fp16, 1 = bf16.  `fmt` bit 0: A is fp16, bit 1: B is fp16.
__host__ __device__ constexpr uint32_t idesc_with_formats(uint32_t idesc_bf16, uint32_t fmt) {
    return idesc_bf16 & ~(((fmt & 1u) << 7) | (((fmt >> 1) & 1u) << 10));
}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {      // one full warp; result lands in *smem_slot
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {         // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads, bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA pairs (cta_group::2): one MMA spans two SMs of a cluster; M = 256 (128 accumulator rows per CTA), each CTA
// supplies its own A rows and HALF of the B rows from its own shared memory at the same offsets, so the B operand
// (the weights) costs half the shared-memory capacity and half the operand bandwidth per SM.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;       // shared::cluster address -> the same offset in the pair's even CTA
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {   // one warp of EACH CTA of the pair, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in the pair's EVEN (leader) CTA, from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
// TMA tiled load whose completion bytes are counted on the LEADER CTA's barrier (data lands in the issuing CTA)
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                   "r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMEM -> registers: 32 lanes (this warp's quarter) x 32 consecutive columns; thread = lane, register j = column j.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error code, never as a hung GPU.  Returns false on time-out.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 26); ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// One lane of a fully active warp (the compiler then knows a single thread runs the guarded code: no per-lane replay
// loop around tcgen05.mma, and warp-uniform operands stay in uniform registers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- bulk async copy global -> shared (TMA engine, 1-D), completion counted in bytes on an mbarrier -------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- TMA: 4-D tiled tensor load global -> shared (box described by a CUtensorMap), completion in bytes on an mbarrier.
// Coordinates are signed, innermost first; elements outside the tensor arrive as zeros (= the convolution's zero padding).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- bf16 split: x = hi + lo (+ ~2^-17 |x|) ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16_rn_bits(float x) {             // round-to-nearest-even bf16, as the upper 16 bits
    uint32_t u = __float_as_uint(x);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u & 0xFFFF0000u;
}
__device__ __forceinline__ void split_bf16(float x, uint32_t& hi_bits, uint32_t& lo_bits) {
    hi_bits = bf16_rn_bits(x);
    lo_bits = bf16_rn_bits(x - __uint_as_float(hi_bits));
}
// Hardware conversion of two floats to one packed bf16x2 word (round to nearest even): `a` lands in the LOW half.
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// two floats -> packed fp16x2 (round to nearest even, saturating to +-65504 instead of inf): `a` in the LOW half
__device__ __forceinline__ uint32_t cvt_f16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// hi = bf16x2(a, b); lo = bf16x2(a - float(hi.a), b - float(hi.b))
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt_bf16x2(a, b);
    lo = cvt_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}
// pack two bf16 (given as fp32 bit patterns whose low halves are zero) into one 32-bit word: e0 in the low half
__device__ __forceinline__ uint32_t pack_bf16(uint32_t e0_bits, uint32_t e1_bits) { return (e0_bits >> 16) | e1_bits; }

}  // namespace umma
