// rcf_conv64.cu -- the 64 -> 64 channel 3x3 convolution of flow_feat_before_agg (reference
// models/flow_aggregation_head_with_residual.py:89-91) as an implicit GEMM on the 5th-generation tensor cores:
// TMA tiled loads (halo tile of a channels-last bf16 tensor, zero fill = the convolution's padding) -> tcgen05.mma (bf16
// operands from shared memory, fp32 accumulators in tensor memory) -> tcgen05.ld epilogue.  Weights are brought in once
// per CTA by the bulk-copy engine; CTAs are persistent (one per SM) and warp-specialised.  The same kernel computes the
// data gradient (weights packed transposed + flipped).  Geometry: rcf_conv64.cuh.
//
// Activations travel between the library's kernels as TWO bf16 tensors, x ~ hi + lo (2^-17 relative): the producer
// kernels (conv stem, pooling backward) write them directly, so this kernel's producer is a single TMA-issuing thread.
//
// Precision ("NPROD", the number of bf16 products per fp32 product):
//   3  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi                              -> fp32-grade (~1e-5)
//   2  A_hi*W_hi + A_hi*W_lo: weights to 2^-17, activations rounded to bf16  -> the TF32-class default
//   1  A_hi*W_hi                                                       -> autocast (bf16) class
// The two weight words of an output channel sit side by side in the N dimension (N = 128: columns 0-63 hi, 64-127 lo),
// so NPROD = 2 costs ONE M128 x N128 x K16 MMA per tap and K-step -- the shape at which the tensor pipe is no longer
// starved by shared-memory operand bandwidth (N = 64 is: 48 clk instead of 32 per MMA, tools/microbench/umma_probe3.cu).
//
// Pipeline per CTA:  the TMA thread loads tile i+1 into the other A buffer while the MMA thread issues tile i and the
// epilogue warps (8: TMEM lane quarter x column half) drain the accumulators of tile i-1 from the other TMEM stage.
#include <cuda.h>

#include "rcf_common.cuh"
#include "rcf_conv64.cuh"
#include "rcf_umma.cuh"

namespace {
using namespace umma;

struct Conv64Args {
    Conv64Geom g;
    float* out;             // [nimg][H][W][64] fp32
    const uint8_t* wpack;   // C64_W_BYTES, see k_conv64_pack
    int* status;            // device word: set to 1 when a barrier wait timed out (protocol bug), never read on the hot path
    int debug;              // measurement switches: 1 no epilogue stores, 4 no MMAs (results invalid)
    uint32_t fmt;           // operand formats: bit 0 activations are fp16, bit 1 packed weights are fp16 (else bf16)
    long long* trace;       // optional: clock64 stamps of CTA 0's pipeline events, [tile][8] (tools/trace_conv64.py)
};
#define C64_TRACE(slot) do { if (a.trace && blockIdx.x == 0 && it < 64 && lane == 0) a.trace[it * 8 + (slot)] = clock64(); } while (0)

struct Bars {
    uint64_t full[C64_NA], empty[C64_NA], tfull[C64_NT], tempty[C64_NT], wbar;
    uint32_t tmem_base, abort_flag;
};

// every wait is bounded: a protocol bug must come back as an error code, not as a hung GPU
__device__ __forceinline__ bool wait_or_abort(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag) {
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

// ---- weight packing ------------------------------------------------------------------------------------------------------
// w [64 co][64 ci][3][3] fp32 -> per tap a K-major B tile of 128 rows x 64 k (bf16, 128-byte swizzle):
//   forward  (transpose_flip = 0): row n <-> co, k <-> ci, tap (ty,tx):      out[q] = sum W[co][ci][ty][tx] in[q + (ty-1)Wp + (tx-1)]
//   data grad (transpose_flip = 1): row n <-> ci, k <-> co, tap (2-ty,2-tx): din[q] = sum W[co][ci][2-ty][2-tx] dout[q + (ty-1)Wp + (tx-1)]
// rows 0-63 hold the bf16 "hi" word of the weight, rows 64-127 the "lo" word.
// transpose_flip = 2: both, the data-gradient image C64_WPACK_TOTAL_BYTES after the forward one (blockIdx.y selects).
// w_f16: the "hi" words are fp16 (11-bit significand, the single-product TF32-class mode), the "lo" words are zero.
__global__ void k_conv64_pack(const float* __restrict__ w, uint8_t* __restrict__ out, int transpose_flip, int w_f16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 9 * 64 * 64) return;
    if (transpose_flip == 2) {
        transpose_flip = (int)blockIdx.y;
        out += (size_t)blockIdx.y * C64_WPACK_TOTAL_BYTES;
    }
    const int k = idx & 63, n = (idx >> 6) & 63, tap = idx >> 12;
    float v;
    if (!transpose_flip) v = w[(n * 64 + k) * 9 + tap];
    else v = w[(k * 64 + n) * 9 + (8 - tap)];
    uint32_t hi, lo;
    split_bf16(v, hi, lo);
    if (w_f16) { hi = (cvt_f16x2(v, 0.0f) & 0xffffu) << 16; lo = 0u; }
    const int chunk = ((k >> 3) ^ (n & 7)) << 4;     // both row n and row 64 + n have (row & 7) == (n & 7)
    uint8_t* base = out + tap * C64_TAP_BYTES + chunk + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(base + n * 128) = (uint16_t)(hi >> 16);
    *reinterpret_cast<uint16_t*>(base + (64 + n) * 128) = (uint16_t)(lo >> 16);
    // CTA-pair images (rcf_conv64_pair.cu): rank 0 region Y = hi rows, rank 1 region Y = lo rows; region Z of rank n / 32 =
    // hi rows of output channels [32 rank, +32)
    uint8_t* pair = out + C64_W_BYTES;
    const int inrow = chunk + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(pair + tap * 8192 + n * 128 + inrow) = (uint16_t)(hi >> 16);
    *reinterpret_cast<uint16_t*>(pair + C64_PAIR_IMAGE_BYTES + tap * 8192 + n * 128 + inrow) = (uint16_t)(lo >> 16);
    *reinterpret_cast<uint16_t*>(pair + (n >> 5) * C64_PAIR_IMAGE_BYTES + C64_PAIR_Y_BYTES + tap * 4096 + (n & 31) * 128 + inrow) =
        (uint16_t)(hi >> 16);
}

// fp32 channels-last -> (hi, lo) bf16 channels-last; the library's own producers write the pair directly, this is for
// callers that hold fp32 activations (tests, rcf_conv64 used stand-alone).
__global__ void k_split_bf16(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    uint32_t h0, h1, l0, l1;
    split_bf16x2(v.x, v.y, h0, l0);
    split_bf16x2(v.z, v.w, h1, l1);
    hi[i] = make_uint2(h0, h1);
    if (lo) lo[i] = make_uint2(l0, l1);
}

// ---- the convolution -----------------------------------------------------------------------------------------------------
template <int NPROD>
__global__ void __launch_bounds__(C64_THREADS, 1)
k_conv64(const Conv64Args a, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* const sW = smem;
    uint8_t* const sA = smem + C64_W_BYTES;            // C64_NA tile buffers
    Bars* const bars = reinterpret_cast<Bars*>(sA + C64_NA * C64_ABUF_BYTES);
    constexpr int NBUF = NPROD == 3 ? 1 : C64_NA;     // NPROD 3: buffer 1 holds the "lo" words of the tile in buffer 0
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Conv64Geom& g = a.g;

    if (tid == 0) {
        for (int i = 0; i < C64_NA; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < C64_NT; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], C64_EPI_WARPS); }
        mbar_init(&bars->wbar, 1);
        bars->abort_flag = (smem_u32(smem) & 1023u) ? 1u : 0u;     // the swizzled images assume a 1024-aligned base
        mbar_init_fence();
    }
    if (warp == C64_MMA_WARP) tmem_alloc<512>(&bars->tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    volatile uint32_t* abort_flag = &bars->abort_flag;
    const int tiles_per_img = g.tiles_x * g.tiles_y;

    if (warp == C64_MMA_WARP) {
        // ================= MMA issue (+ one-time weight load) =================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wbar, C64_W_BYTES);
            for (int t = 0; t < 9; ++t) bulk_g2s(sW + t * C64_TAP_BYTES, a.wpack + t * C64_TAP_BYTES, C64_TAP_BYTES, &bars->wbar);
        }
        __syncwarp();
        wait_or_abort(&bars->wbar, 0, abort_flag);
        const uint32_t IDESC64 = idesc_with_formats(make_idesc_bf16(128, 64, 0, 0), a.fmt);
        const uint32_t IDESC128 = idesc_with_formats(make_idesc_bf16(128, 128, 0, 0), a.fmt);
        const uint64_t wdesc = make_desc_sw128(smem_u32(sW), 16, 1024);
        const uint32_t Wp8 = (uint32_t)g.Wp * 8;                     // one tile row, in 16-byte units
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int b = it % NBUF, ph = (it / NBUF) & 1, ts = it % C64_NT, tph = (it / C64_NT) & 1;
            wait_or_abort(&bars->full[b], ph, abort_flag);
            C64_TRACE(0);
            wait_or_abort(&bars->tempty[ts], tph ^ 1, abort_flag);
            C64_TRACE(1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t adesc = make_desc_sw128(smem_u32(sA + b * C64_ABUF_BYTES), 16, 1024);
                const uint64_t adesc_lo = make_desc_sw128(smem_u32(sA + C64_ABUF_BYTES), 16, 1024);
                if (!(a.debug & 4)) {
                    const uint32_t dcol = tmem + ts * 128;
                    const uint64_t aj = adesc, ajl = adesc_lo;
#pragma unroll
                    for (int ty = 0; ty < 3; ++ty) {
                        const uint64_t ar = aj + (uint64_t)(ty * Wp8), arl = ajl + (uint64_t)(ty * Wp8);
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t aoff = tx * 8 + ks * 2, boff = (ty * 3 + tx) * (C64_TAP_BYTES >> 4) + ks * 2;
                                const uint32_t acc = (ty | tx | ks) != 0;
                                if (NPROD == 1) mma_bf16(dcol, ar + aoff, wdesc + boff, IDESC64, acc);
                                else mma_bf16(dcol, ar + aoff, wdesc + boff, IDESC128, acc);
                                if (NPROD == 3) mma_bf16(dcol, arl + aoff, wdesc + boff, IDESC64, 1);
                            }
                    }
                }
                mma_commit(&bars->empty[b]);      // the staged tile may be overwritten once these MMAs have read it
                mma_commit(&bars->tfull[ts]);     // ... and the accumulators are complete
            }
            __syncwarp();
            C64_TRACE(2);
        }
    } else if (warp == C64_TMA_WARP) {
        // ================= producer: one thread, one (NPROD 3: two) TMA tiled load(s) per tile =================
        if (lane == 0) {
            tma_prefetch_desc(&tm_hi);
            if (NPROD == 3) tma_prefetch_desc(&tm_lo);
            const uint32_t bytes = (uint32_t)g.npos * 128u * (NPROD == 3 ? 2u : 1u);
            int it = 0;
            for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
                const int b = it % NBUF, ph = (it / NBUF) & 1;
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
                wait_or_abort(&bars->empty[b], ph ^ 1, abort_flag);
                C64_TRACE(3);
                if (a.trace && blockIdx.x == 0 && it < 64) { long long ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns)); a.trace[it * 8 + 6] = ns; }
                mbar_arrive_expect_tx(&bars->full[b], bytes);
                tma_load_4d(sA + b * C64_ABUF_BYTES, &tm_hi, 0, txi * g.TW - 1, tyi * g.TR - 1, img, &bars->full[b]);
                if (NPROD == 3) tma_load_4d(sA + C64_ABUF_BYTES, &tm_lo, 0, txi * g.TW - 1, tyi * g.TR - 1, img, &bars->full[b]);
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> channels-last fp32 =================
        const int quarter = warp & 3, c0 = (warp >> 2) * 32;         // TMEM lanes [32 quarter, +32), output channels [c0, c0 + 32)
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int ts = it % C64_NT, tph = (it / C64_NT) & 1;
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
            float* dst = a.out + (size_t)img * g.H * g.W * 64;
            wait_or_abort(&bars->tfull[ts], tph, abort_flag);
            if (warp == 0) C64_TRACE(4);
            fence_after_sync();
            {
                // this lane's accumulator row <-> output pixel (or -1: halo column / outside the tile or image)
                const int q = g.Wp + 1 + quarter * 32 + lane;
                const int r = (int)__umulhi((uint32_t)q, g.wp_magic) - 1, x = q - (r + 1) * g.Wp - 1;
                const int y = tyi * g.TR + r, xx = txi * g.TW + x;
                const int mypix = (x >= 0 && x < g.TW && r < g.TR && y < g.H && xx < g.W && !(a.debug & 1)) ? y * g.W + xx : -1;
                // after the register <-> lane exchange below, float4 i of lane t belongs to row 4i + (t & 3)
                int pix[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pix[i] = __shfl_sync(0xffffffffu, mypix, 4 * i + (lane & 3));
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + ts * 128 + c0;
                uint32_t v[32];
                tmem_ld32(taddr, v);
                if (NPROD >= 2) {
                    uint32_t u[32];
                    tmem_ld32(taddr + 64, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
                } else {
                    tmem_ld_wait();
                }
                if (warp == 0) C64_TRACE(7);
                // Lane = row, register = channel  ->  swap lane bits 4,3,2 with register bits 4,3,2: lane t then holds, in
                // registers 4i..4i+3, channels c0 + 4(t>>2) .. +3 of row 4i + (t&3); a store instruction covers 4 rows x
                // 128 contiguous bytes (4 full lines) instead of 32 rows x 16 bytes (32 lines).
#pragma unroll
                for (int bit = 4; bit >= 2; --bit) {
                    const int m = 1 << bit;
                    const bool up = (lane & m) != 0;
#pragma unroll
                    for (int xr = 0; xr < 32; ++xr) {
                        if (xr & m) continue;
                        const uint32_t send = up ? v[xr] : v[xr | m];
                        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
                        if (up) v[xr] = recv; else v[xr | m] = recv;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (pix[i] >= 0)
                        *reinterpret_cast<float4*>(dst + (size_t)pix[i] * 64 + c0 + 4 * (lane >> 2)) =
                            make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                        __uint_as_float(v[4 * i + 3]));
            }
            fence_before_sync();
            __syncwarp();
            if (warp == 0) C64_TRACE(5);
            if (lane == 0) mbar_arrive(&bars->tempty[ts]);
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == C64_MMA_WARP) tmem_dealloc<512>(tmem);
    if (tid == 0 && bars->abort_flag && a.status) *a.status = 1;
}

int g_conv64_attr_done = 0;
int g_conv64_debug = 0;
int g_conv64_pair = 1;          // CTA-pair kernel (rcf_conv64_pair.cu) instead of the one-CTA kernel
long long* g_conv64_trace = nullptr;
__device__ int g_conv64_status;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// Tensor map of a channels-last bf16 tensor [nimg][H][W][64] with a (64, bw, bh, 1) box and the 128-byte swizzle.
int rcf_make_tmap_nhwc64(CUtensorMap* tm, const void* base, int nimg, int H, int W, int bw, int bh) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return (int)cudaErrorNotSupported;
    const cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? RCF_OK : (int)cudaErrorInvalidValue;
}

int rcf_conv64_pair_launch(const void* in_hi, const void* in_lo, const void* wpack_pair, float* out, int nimg, int H, int W,
                           int nprod, uint32_t fmt, cudaStream_t s);

int* rcf_conv64_status_addr() {
    static int* addr = nullptr;          // resolved once (outside any stream capture of later calls)
    if (!addr && cudaGetSymbolAddress((void**)&addr, g_conv64_status) != cudaSuccess) addr = nullptr;
    return addr;
}
int rcf_conv64_debug_flags() { return g_conv64_debug; }

// ---- C ABI ------------------------------------------------------------------------------------------------------------------
extern "C" {

RCF_API int rcf_conv64_pack_weights(const float* w, void* wpack, int transpose_flip, void* stream) {
    if (!w || !wpack) return RCF_ERR_NULL;
    const int w_f16 = (transpose_flip & RCF_CONV64_W_F16) ? 1 : 0;
    transpose_flip &= 0xff;
    if (((uintptr_t)wpack & 15) != 0) return RCF_ERR_ALIGN;
    if (transpose_flip < 0 || transpose_flip > 2) return RCF_ERR_MODE;
    const dim3 grid((9 * 64 * 64 + 255) / 256, transpose_flip == 2 ? 2 : 1);
    k_conv64_pack<<<grid, 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)wpack, transpose_flip, w_f16);
    return (int)cudaGetLastError();
}

RCF_API int rcf_split_bf16(const float* x, void* hi, void* lo, size_t n, void* stream) {
    if (!x || !hi) return RCF_ERR_NULL;
    if (n % 4 || (((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15)) return RCF_ERR_ALIGN;
    if (n == 0) return RCF_OK;
    k_split_bf16<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n / 4);
    return (int)cudaGetLastError();
}

RCF_API int rcf_conv64_forward(const void* in_hi, const void* in_lo, const void* wpack, float* out, int nimg, int H, int W,
                               int nprod, void* stream) {
    const uint32_t fmt = ((nprod & RCF_CONV64_A_F16) ? 1u : 0u) | ((nprod & RCF_CONV64_W_F16) ? 2u : 0u);
    nprod &= 0xff;
    if (fmt && nprod != 1) return RCF_ERR_MODE;          // fp16 operands are the single-product mode
    if (!in_hi || !wpack || !out || (nprod == 3 && !in_lo)) return RCF_ERR_NULL;
    if (nimg < 1 || H < 1 || W < 1 || (long long)nimg * H * W > (1ll << 31)) return RCF_ERR_SHAPE;
    if (nprod < 1 || nprod > 3) return RCF_ERR_MODE;
    if ((((uintptr_t)in_hi | (uintptr_t)in_lo | (uintptr_t)out | (uintptr_t)wpack) & 15) != 0) return RCF_ERR_ALIGN;
    if (!g_conv64_attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_conv64<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv64<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv64<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        g_conv64_attr_done = 1;
    }
    if (g_conv64_pair)
        return rcf_conv64_pair_launch(in_hi, in_lo, (const uint8_t*)wpack + C64_W_BYTES, out, nimg, H, W, nprod, fmt, (cudaStream_t)stream);
    Conv64Args a;
    a.g = conv64_make_geom(nimg, H, W);
    a.out = out; a.wpack = (const uint8_t*)wpack; a.debug = g_conv64_debug; a.trace = g_conv64_trace; a.fmt = fmt;
    a.status = rcf_conv64_status_addr();
    if (!a.status) return (int)cudaErrorInvalidSymbol;
    alignas(64) CUtensorMap tm_hi, tm_lo;
    int e = rcf_make_tmap_nhwc64(&tm_hi, in_hi, nimg, H, W, a.g.Wp, a.g.TR + 2);
    if (e != RCF_OK) return e;
    e = rcf_make_tmap_nhwc64(&tm_lo, nprod == 3 ? in_lo : in_hi, nimg, H, W, a.g.Wp, a.g.TR + 2);
    if (e != RCF_OK) return e;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.ntiles < nsm ? a.g.ntiles : nsm;
    cudaStream_t s = (cudaStream_t)stream;
    if (nprod == 1) k_conv64<1><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a, tm_hi, tm_lo);
    else if (nprod == 2) k_conv64<2><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a, tm_hi, tm_lo);
    else k_conv64<3><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a, tm_hi, tm_lo);
    return (int)cudaGetLastError();
}

void rcf_conv64_set_debug(int v) { g_conv64_debug = v; }
void rcf_conv64_set_pair(int v) { g_conv64_pair = v ? 1 : 0; }
// Measurement hook: device buffer of 64 x 8 int64 that CTA 0 of the following conv launches fills with clock64 stamps
// (per tile: 0 tile landed, 1 TMEM stage free, 2 MMAs issued, 3 A buffer free, 4 accumulators complete, 5 epilogue done).
RCF_API int rcf_debug_conv64_trace(void* buf) { g_conv64_trace = (long long*)buf; return RCF_OK; }

// Test hook (synchronises): 1 if any tcgen05 kernel of this process hit a barrier time-out since the last call.
RCF_API int rcf_debug_conv64_status(void) {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_conv64_status, sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_conv64_status, &zero, sizeof(int));
    return v;
}

}  // extern "C"
