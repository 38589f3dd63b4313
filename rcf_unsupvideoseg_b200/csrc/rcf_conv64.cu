// rcf_conv64.cu -- the 64 -> 64 channel 3x3 convolution of flow_feat_before_agg (reference
// models/flow_aggregation_head_with_residual.py:89-91) as an implicit GEMM on the 5th-generation tensor cores:
// tcgen05.mma (bf16 operands from shared memory, fp32 accumulators in tensor memory), weights brought in once per CTA by
// the bulk-copy engine, persistent warp-specialised CTAs (one per SM).  The same kernel computes the data gradient
// (weights packed transposed + flipped).  Geometry: rcf_conv64.cuh.
//
// Precision ("NPROD", the number of bf16 products per fp32 product):
//   3  x ~ x_hi + x_lo (two bf16 words, 2^-17 relative):  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi   -> fp32-grade (~1e-5)
//   2  A_hi*W_hi + A_hi*W_lo: weights to 2^-17, activations rounded to bf16                      -> the TF32-class default
//   1  A_hi*W_hi                                                                                 -> autocast (bf16) class
// The two weight words of an output channel sit side by side in the N dimension (N = 128: columns 0-63 hi, 64-127 lo),
// so NPROD = 2 costs ONE M128 x N128 x K16 MMA per tap and K-step -- the shape at which the tensor pipe is no longer
// starved by shared-memory operand bandwidth (N = 64 is: 48 clk instead of 32 per MMA, tools/microbench/umma_probe3.cu).
//
// Pipeline per CTA:  producers (7 warps) stage tile i+1 into the other A buffer while the MMA warp issues tile i and the
// epilogue warps (4, one per TMEM lane quarter) drain the accumulators of tile i-1 from the other TMEM stage.
#include "rcf_common.cuh"
#include "rcf_conv64.cuh"
#include "rcf_umma.cuh"

namespace {
using namespace umma;

struct Conv64Args {
    Conv64Geom g;
    const float* in;        // [nimg][H][W][64] fp32, channels-last
    float* out;             // [nimg][H][W][64] fp32
    const uint8_t* wpack;   // C64_W_BYTES, see k_conv64_pack
    int* status;            // device word: set to 1 when a barrier wait timed out (protocol bug), never read on the hot path
};

struct Bars {
    uint64_t full[2], empty[2], tfull[2], tempty[2], wbar;
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
__global__ void k_conv64_pack(const float* __restrict__ w, uint8_t* __restrict__ out, int transpose_flip) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 9 * 64 * 64) return;
    const int k = idx & 63, n = (idx >> 6) & 63, tap = idx >> 12;
    float v;
    if (!transpose_flip) v = w[(n * 64 + k) * 9 + tap];
    else v = w[(k * 64 + n) * 9 + (8 - tap)];
    uint32_t hi, lo;
    split_bf16(v, hi, lo);
    const int chunk = ((k >> 3) ^ (n & 7)) << 4;     // both row n and row 64 + n have (row & 7) == (n & 7)
    uint8_t* base = out + tap * C64_TAP_BYTES + chunk + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(base + n * 128) = (uint16_t)(hi >> 16);
    *reinterpret_cast<uint16_t*>(base + (64 + n) * 128) = (uint16_t)(lo >> 16);
}

// ---- the convolution -----------------------------------------------------------------------------------------------------
template <int NPROD>
__global__ void __launch_bounds__(C64_THREADS, 1) k_conv64(const Conv64Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* const sW = smem;
    uint8_t* const sA0 = smem + C64_W_BYTES;
    uint8_t* const sA1 = sA0 + C64_ABUF_BYTES;
    Bars* const bars = reinterpret_cast<Bars*>(sA1 + C64_ABUF_BYTES);
    constexpr int NBUF = NPROD == 3 ? 1 : 2;          // NPROD 3: buffer 1 holds the "lo" words of the tile in buffer 0
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Conv64Geom& g = a.g;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->full[i], C64_PROD_WARPS);
            mbar_init(&bars->empty[i], 1);
            mbar_init(&bars->tfull[i], 1);
            mbar_init(&bars->tempty[i], C64_EPI_WARPS);
        }
        mbar_init(&bars->wbar, 1);
        bars->abort_flag = (smem_u32(smem) & 1023u) ? 1u : 0u;     // the swizzled weight image assumes a 1024-aligned base
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
        constexpr uint32_t IDESC64 = make_idesc_bf16(128, 64, 0, 0), IDESC128 = make_idesc_bf16(128, 128, 0, 0);
        const uint64_t wdesc = make_desc_sw128(smem_u32(sW), 16, 1024);
        const uint32_t Wp8 = (uint32_t)g.Wp * 8;                     // one tile row, in 16-byte units
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int b = it % NBUF, ph = (it / NBUF) & 1, ts = it & 1, tph = (it >> 1) & 1;
            wait_or_abort(&bars->full[b], ph, abort_flag);
            wait_or_abort(&bars->tempty[ts], tph ^ 1, abort_flag);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t adesc = make_desc_sw128(smem_u32(b ? sA1 : sA0), 16, 1024);
                const uint64_t adesc_lo = make_desc_sw128(smem_u32(sA1), 16, 1024);
                for (int j = 0; j < g.nmt; ++j) {
                    const uint32_t dcol = tmem + ts * 256 + j * 128;
                    const uint64_t aj = adesc + (uint64_t)(j * 128 * 8), ajl = adesc_lo + (uint64_t)(j * 128 * 8);
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
        }
    } else if (warp >= C64_PROD_WARP0) {
        // ================= producers: fp32 channels-last pixels -> bf16 (hi [, lo]) swizzled rows =================
        const int pt = tid - C64_PROD_WARP0 * 32;
        const int nitems = g.npos * 8;                                 // (position, 8-channel chunk)
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int b = it % NBUF, ph = (it / NBUF) & 1;
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
            const int y0 = tyi * g.TR - 1, x0 = txi * g.TW - 1;
            const float* src = a.in + (size_t)img * g.H * g.W * 64;
            wait_or_abort(&bars->empty[b], ph ^ 1, abort_flag);
            const uint32_t dst_hi = smem_u32(b ? sA1 : sA0), dst_lo = smem_u32(sA1);
            for (int i0 = pt; i0 < nitems; i0 += 4 * C64_PROD_THREADS) {
                float4 v[4][2];
                int q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = i0 + u * C64_PROD_THREADS;
                    q[u] = idx >> 3;
                    const int c = idx & 7, r = q[u] / g.Wp, x = q[u] - r * g.Wp, y = y0 + r, xx = x0 + x;
                    v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < nitems && y >= 0 && y < g.H && xx >= 0 && xx < g.W) {
                        const float4* p = reinterpret_cast<const float4*>(src + ((size_t)y * g.W + xx) * 64 + c * 8);
                        v[u][0] = __ldg(p); v[u][1] = __ldg(p + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = i0 + u * C64_PROD_THREADS;
                    if (idx >= nitems) break;
                    const float f[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) split_bf16(f[e], hi[e], lo[e]);
                    const uint32_t row = dst_hi + q[u] * 128;
                    const uint32_t off = (((idx & 7) ^ ((row >> 7) & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + off), "r"(pack_bf16(hi[0], hi[1])),
                                 "r"(pack_bf16(hi[2], hi[3])), "r"(pack_bf16(hi[4], hi[5])), "r"(pack_bf16(hi[6], hi[7])) : "memory");
                    if (NPROD == 3) {
                        const uint32_t rowl = dst_lo + q[u] * 128;
                        const uint32_t offl = (((idx & 7) ^ ((rowl >> 7) & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowl + offl), "r"(pack_bf16(lo[0], lo[1])),
                                     "r"(pack_bf16(lo[2], lo[3])), "r"(pack_bf16(lo[4], lo[5])), "r"(pack_bf16(lo[6], lo[7])) : "memory");
                    }
                }
            }
            fence_async_smem();              // generic-proxy stores -> visible to the tensor core's operand reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->full[b]);
        }
    } else {
        // ================= epilogue: TMEM -> registers -> channels-last fp32 =================
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int ts = it & 1, tph = (it >> 1) & 1;
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
            float* dst = a.out + (size_t)img * g.H * g.W * 64;
            wait_or_abort(&bars->tfull[ts], tph, abort_flag);
            fence_after_sync();
            for (int j = 0; j < g.nmt; ++j) {
                const int q = g.Wp + 1 + 128 * j + warp * 32 + lane;
                const int r = q / g.Wp - 1, x = q - (r + 1) * g.Wp - 1;
                const int y = tyi * g.TR + r, xx = txi * g.TW + x;
                const bool valid = x >= 0 && x < g.TW && r < g.TR && y < g.H && xx < g.W;
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ts * 256 + j * 128;
                float* o = dst + ((size_t)y * g.W + xx) * 64;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    if (NPROD >= 2) {
                        uint32_t u[32];
                        tmem_ld32(taddr + 64 + c0, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
                    } else {
                        tmem_ld_wait();
                    }
                    if (valid) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(o + c0 + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                                  __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tempty[ts]);
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == C64_MMA_WARP) tmem_dealloc<512>(tmem);
    if (tid == 0 && bars->abort_flag && a.status) *a.status = 1;
}

int g_conv64_attr_done = 0;
__device__ int g_conv64_status;

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------------------------------
extern "C" {

RCF_API int rcf_conv64_pack_weights(const float* w, void* wpack, int transpose_flip, void* stream) {
    if (!w || !wpack) return RCF_ERR_NULL;
    if (((uintptr_t)wpack & 15) != 0) return RCF_ERR_ALIGN;
    k_conv64_pack<<<(9 * 64 * 64 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)wpack, transpose_flip);
    return (int)cudaGetLastError();
}

RCF_API int rcf_conv64_forward(const float* in, const void* wpack, float* out, int nimg, int H, int W, int nprod, void* stream) {
    if (!in || !wpack || !out) return RCF_ERR_NULL;
    if (nimg < 1 || H < 1 || W < 1 || (long long)nimg * H * W > (1ll << 31)) return RCF_ERR_SHAPE;
    if (nprod < 1 || nprod > 3) return RCF_ERR_MODE;
    if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)wpack) & 15) != 0) return RCF_ERR_ALIGN;
    if (!g_conv64_attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_conv64<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv64<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv64<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C64_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        g_conv64_attr_done = 1;
    }
    Conv64Args a;
    a.g = conv64_make_geom(nimg, H, W);
    a.in = in; a.out = out; a.wpack = (const uint8_t*)wpack;
    static int* status_addr = nullptr;          // resolved once (outside any stream capture of later calls)
    if (!status_addr) {
        const cudaError_t e = cudaGetSymbolAddress((void**)&status_addr, g_conv64_status);
        if (e != cudaSuccess) return (int)e;
    }
    a.status = status_addr;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.ntiles < nsm ? a.g.ntiles : nsm;
    cudaStream_t s = (cudaStream_t)stream;
    if (nprod == 1) k_conv64<1><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a);
    else if (nprod == 2) k_conv64<2><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a);
    else k_conv64<3><<<grid, C64_THREADS, C64_SMEM_BYTES, s>>>(a);
    return (int)cudaGetLastError();
}

// Test hook (synchronises): 1 if any tcgen05 kernel of this process hit a barrier time-out since the last call.
RCF_API int rcf_debug_conv64_status(void) {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_conv64_status, sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_conv64_status, &zero, sizeof(int));
    return v;
}

}  // extern "C"
