// rcf_conv64_wgrad.cu -- weight gradient of the 64 -> 64 channel 3x3 convolution of flow_feat_before_agg (reference
// models/flow_aggregation_head_with_residual.py:89-91; replaces the cuDNN wgrad kernel autograd runs for it) on tcgen05:
//     dW[co][ci][ty][tx] = sum over images, pixels (y, x) of  dOut[y][x][co] * In[y + ty - 1][x + tx - 1][ci]
// as GEMMs whose K dimension is the PIXEL axis.  Both operands are the channels-last bf16 (hi, lo) tensors the library's
// producers write; TMA tiled loads bring row segments into shared memory ([position][64 ch] rows of 128 B, 128-byte
// swizzle) and the tensor core reads them MN-major (rows = channels, K = 16 consecutive positions), no transpose.
//
// One MMA covers SIX taps:   M = 128 = (ty, ty+1) x 64 ci   -- two "atoms" of the A operand one tile row apart (LBO),
//                            N = 192 = 3 shifts x 64 co      -- three atoms of the B operand ONE position apart (LBO = 128 B):
//   D[(a,ci)][(u,co)] += sum_k In_tile[row r + ty0 + a][c + k + 2][ci] * dOut_tile[row r][c + k + u][co],   tx = 2 - u.
// Shifting dOut instead of In is legal because the chunks c tile the whole (zero-padded) image row: the sum over all
// chunks of all row segments runs over every position of the row, and the TMA zero fill outside the image supplies the
// two zero positions the shifted windows need at both ends.  Two such MMAs ((ty 0,1) and (ty 1,2); ty = 1 is computed
// twice and dropped once) give all nine taps with 10 KB of operand reads per 96 clk of tensor time -- under the 128 B/clk
// of shared-memory bandwidth that starves the M128 x N64 shape (48 instead of 32 clk, tools/microbench/umma_probe3.cu).
//
// Precision: nprod 3: In_hi*dOut_hi + In_lo*dOut_hi + In_hi*dOut_lo (fp32-grade); 2: In_hi*(dOut_hi + dOut_lo); 1: hi*hi.
// Accumulators (2 x 128 lanes x 192 columns fp32) stay in tensor memory for the CTA's whole tile list and are written
// once, as per-CTA partials that a second kernel sums in a fixed order (bit-reproducible).
#include <cuda.h>

#include "rcf_common.cuh"
#include "rcf_umma.cuh"
#include "rcf_internal.h"

int rcf_make_tmap_nhwc64(CUtensorMap* tm, const void* base, int nimg, int H, int W, int bw, int bh);
int* rcf_conv64_status_addr();

namespace {
using namespace umma;

constexpr int WG_THREADS = 320;       // warps 0-7 final drain (TMEM lane quarter x accumulator), warp 8 MMA issue, warp 9 TMA
constexpr int WG_MMA_WARP = 8, WG_TMA_WARP = 9;
constexpr int WG_STAGES = 2;
constexpr int WG_PART_FLOATS = 2 * 128 * 192;

struct WgradGeom {
    int nimg, H, W;
    int L, TR;                 // tile: TR output rows x L positions of the padded row (L % 16 == 0)
    int tiles_x, tiles_y, ntiles;
    int xrow_bytes;            // (L + 2) * 128: one tile row of either operand
    int xbuf_bytes, gbuf_bytes;   // one (hi or lo) In tile / dOut tile, rounded up to 1024
    int stage_bytes;
};

struct WgradArgs {
    WgradGeom g;
    float* part;               // [gridDim.x][2][128][192]
    int* status;
    uint32_t fmt;              // operand formats: bit 0 the layer input x (A operand) is fp16, bit 1 the gradient (B) is fp16
};

struct WgBars {
    uint64_t full[WG_STAGES], empty[WG_STAGES], done;
    uint32_t tmem_base, abort_flag;
};

__device__ __forceinline__ bool wg_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag) {
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

template <int NPROD>
__global__ void __launch_bounds__(WG_THREADS, 1)
k_conv64_wgrad(const WgradArgs a, const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
               const __grid_constant__ CUtensorMap tm_gh, const __grid_constant__ CUtensorMap tm_gl) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const WgradGeom& g = a.g;
    // stage layout: [In hi][In lo (NPROD 3)][dOut hi][dOut lo (NPROD >= 2)]
    constexpr int NX = NPROD == 3 ? 2 : 1, NG = NPROD >= 2 ? 2 : 1;
    WgBars* const bars = reinterpret_cast<WgBars*>(smem + WG_STAGES * g.stage_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->done, 1);
        bars->abort_flag = (smem_u32(smem) & 1023u) ? 1u : 0u;
        mbar_init_fence();
    }
    if (warp == WG_MMA_WARP) tmem_alloc<512>(&bars->tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    volatile uint32_t* abort_flag = &bars->abort_flag;
    const int tiles_per_img = g.tiles_x * g.tiles_y;
    const int nchunk = g.L >> 4;

    if (warp == WG_MMA_WARP) {
        const uint32_t IDESC = idesc_with_formats(make_idesc_bf16(128, 192, 1, 1), a.fmt);
        int it = 0;
        for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
            const int s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
            wg_wait(&bars->full[s], ph, abort_flag);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t sbase = smem_u32(smem + s * g.stage_bytes);
                const uint32_t xh = sbase, xl = sbase + g.xbuf_bytes;
                const uint32_t gh = sbase + NX * g.xbuf_bytes, gl = gh + g.gbuf_bytes;
                // A: two 64-channel atoms one tile row apart; B: three 64-channel atoms one position (128 B) apart
                const uint64_t adesc_hi = make_desc_sw128(0, (uint32_t)g.xrow_bytes, 1024);
                const uint64_t bdesc_hi = make_desc_sw128(0, 128, 1024);
                for (int r = 0; r < g.TR; ++r)
                    for (int c = 0; c < nchunk; ++c) {
                        const uint32_t xoff = (uint32_t)(r * g.xrow_bytes + (c * 16 + 2) * 128), goff = (uint32_t)(r * g.xrow_bytes + c * 16 * 128);
                        const uint32_t first = (it | r | c) == 0 ? 0u : 1u;
#pragma unroll
                        for (int pr = 0; pr < NPROD; ++pr) {
                            // products: 0 In_hi*dOut_hi; 1 In_hi*dOut_lo (NPROD 2) or In_lo*dOut_hi (NPROD 3); 2 In_hi*dOut_lo
                            const uint32_t xs = (NPROD == 3 && pr == 1) ? xl : xh;
                            const uint32_t gs = ((NPROD == 2 && pr == 1) || pr == 2) ? gl : gh;
                            const uint64_t bd = desc_at(bdesc_hi, gs + goff);
                            const uint32_t acc = pr ? 1u : first;
                            mma_bf16(tmem, desc_at(adesc_hi, xs + xoff), bd, IDESC, acc);
                            mma_bf16(tmem + 192, desc_at(adesc_hi, xs + xoff + g.xrow_bytes), bd, IDESC, acc);
                        }
                    }
                mma_commit(&bars->empty[s]);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(&bars->done);
        __syncwarp();
    } else if (warp == WG_TMA_WARP) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_xh); tma_prefetch_desc(&tm_gh);
            const uint32_t xbytes = (uint32_t)(g.TR + 2) * g.xrow_bytes, gbytes = (uint32_t)g.TR * g.xrow_bytes;
            int it = 0;
            for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
                const int s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
                const int x0 = txi * g.L, y0 = tyi * g.TR;
                uint8_t* sb = smem + s * g.stage_bytes;
                wg_wait(&bars->empty[s], ph ^ 1, abort_flag);
                mbar_arrive_expect_tx(&bars->full[s], NX * xbytes + NG * gbytes);
                tma_load_4d(sb, &tm_xh, 0, x0 - 3, y0 - 1, img, &bars->full[s]);
                if (NPROD == 3) tma_load_4d(sb + g.xbuf_bytes, &tm_xl, 0, x0 - 3, y0 - 1, img, &bars->full[s]);
                tma_load_4d(sb + NX * g.xbuf_bytes, &tm_gh, 0, x0 - 2, y0, img, &bars->full[s]);
                if (NPROD >= 2) tma_load_4d(sb + NX * g.xbuf_bytes + g.gbuf_bytes, &tm_gl, 0, x0 - 2, y0, img, &bars->full[s]);
            }
        }
    } else {
        // final drain: warp w reads TMEM lanes [32 (w & 3), +32) of accumulator (w >> 2)
        wg_wait(&bars->done, 0, abort_flag);
        fence_after_sync();
        const int quarter = warp & 3, accn = warp >> 2;
        float* dst = a.part + ((size_t)blockIdx.x * 2 + accn) * 128 * 192 + (size_t)(quarter * 32 + lane) * 192;
        const bool any = blockIdx.x < g.ntiles;          // a CTA without tiles never initialised its accumulators
#pragma unroll 1
        for (int c0 = 0; c0 < 192; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + accn * 192 + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 4)
                *reinterpret_cast<float4*>(dst + c0 + e) = any ? make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                             __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]))
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_before_sync();
    }

    fence_before_sync();
    __syncthreads();
    if (warp == WG_MMA_WARP) tmem_dealloc<512>(tmem);
    if (tid == 0 && bars->abort_flag && a.status) *a.status = 1;
}

// dW[co][ci][ty][tx] = sum over CTAs (fixed order) of the partial accumulators:
//   accumulator 0, lane a*64 + ci, column u*64 + co  ->  (ty = a,     tx = 2 - u)
//   accumulator 1, lane a*64 + ci, column u*64 + co  ->  (ty = a + 1, tx = 2 - u); its a = 0 half duplicates ty = 1 and is dropped
__global__ void __launch_bounds__(256) k_conv64_wgrad_reduce(const float* __restrict__ part, int nparts, float* __restrict__ dw,
                                                             const float* __restrict__ gmax, int nfd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per output element, co fastest within a warp
    if (idx >= 64 * 64 * 9) return;                                 // (64 * 64 * 9 is a multiple of 32: whole warps leave)
    const int co = idx & 63, ci = (idx >> 6) & 63, tap = idx >> 12;
    const int ty = tap / 3, tx = tap - ty * 3;
    const int accn = ty == 2 ? 1 : 0, arow = ty == 2 ? 1 : ty;
    const size_t off = ((size_t)accn * 128 + arow * 64 + ci) * 192 + (2 - tx) * 64 + co;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 3 < nparts; p += 4) {
        s0 += __ldg(part + (size_t)p * WG_PART_FLOATS + off);
        s1 += __ldg(part + (size_t)(p + 1) * WG_PART_FLOATS + off);
        s2 += __ldg(part + (size_t)(p + 2) * WG_PART_FLOATS + off);
        s3 += __ldg(part + (size_t)(p + 3) * WG_PART_FLOATS + off);
    }
    for (; p < nparts; ++p) s0 += __ldg(part + (size_t)p * WG_PART_FLOATS + off);
    // fp16 gradient words carry a power-of-two scale: divide it out (exact).  After the loads, so the sum loop is unchanged.
    const float inv = gmax ? 1.0f / rcf_grad_scale(gmax, nfd) : 1.0f;
    dw[(co * 64 + ci) * 9 + tap] = ((s0 + s1) + (s2 + s3)) * inv;
}

WgradGeom wgrad_geom(int nimg, int H, int W, int nprod) {
    WgradGeom best = {};
    long long best_cost = -1;
    const int nx = nprod == 3 ? 2 : 1, ng = nprod >= 2 ? 2 : 1;
    for (int L = 16; L <= 96; L += 16)
        for (int TR = 1; TR <= 16; ++TR) {
            const int xrow = (L + 2) * 128;
            const int xbuf = ((TR + 2) * xrow + 1023) & ~1023, gbuf = (TR * xrow + 1023) & ~1023;
            const int stage = nx * xbuf + ng * gbuf;
            if (WG_STAGES * stage + 256 > 227 * 1024 || TR + 2 > 256) continue;
            const long long tx = (W + 2 + L - 1) / L, ty = (H + TR - 1) / TR;
            // MMA work ~ positions; loaded bytes matter second (halo rows / overlap columns)
            const long long cost = tx * ty * ((long long)TR * L * 16 + (long long)stage / 64);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best.L = L; best.TR = TR; best.tiles_x = (int)tx; best.tiles_y = (int)ty;
                best.xrow_bytes = xrow; best.xbuf_bytes = xbuf; best.gbuf_bytes = gbuf; best.stage_bytes = stage;
            }
        }
    best.nimg = nimg; best.H = H; best.W = W;
    best.ntiles = nimg * best.tiles_x * best.tiles_y;
    return best;
}

int wgrad_grid(int ntiles) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    return ntiles < nsm ? ntiles : nsm;
}

template <int NPROD>
int launch_wgrad(const WgradArgs& a, const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& gh, const CUtensorMap& gl,
                 int grid, size_t smem, cudaStream_t s) {
    static bool attr_done = false;
    if (!attr_done) {
        const cudaError_t e = cudaFuncSetAttribute(k_conv64_wgrad<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    k_conv64_wgrad<NPROD><<<grid, WG_THREADS, smem, s>>>(a, xh, xl, gh, gl);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

RCF_API int rcf_conv64_wgrad_workspace_bytes(int nimg, int H, int W, size_t* bytes) {
    if (!bytes) return RCF_ERR_NULL;
    if (nimg < 1 || H < 1 || W < 1) return RCF_ERR_SHAPE;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    *bytes = (size_t)nsm * WG_PART_FLOATS * sizeof(float);
    return RCF_OK;
}

RCF_API int rcf_conv64_wgrad(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* ws,
                             int nimg, int H, int W, int nprod, void* stream) {
    return rcf_conv64_wgrad_ex(x_hi, x_lo, g_hi, g_lo, dw, ws, nimg, H, W, nprod, nullptr, 0, stream);
}

}  // extern "C"

int rcf_conv64_wgrad_ex(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* ws, int nimg,
                        int H, int W, int nprod, const float* gmax, int nfd, void* stream) {
    const uint32_t fmt = ((nprod & RCF_CONV64_A_F16) ? 1u : 0u) | ((nprod & RCF_CONV64_W_F16) ? 2u : 0u);
    nprod &= 0xff;
    if (fmt && nprod != 1) return RCF_ERR_MODE;          // fp16 operands are the single-product mode
    if (!x_hi || !g_hi || !dw || !ws || (nprod == 3 && !x_lo) || (nprod >= 2 && !g_lo)) return RCF_ERR_NULL;
    if (nimg < 1 || H < 1 || W < 1 || (long long)nimg * H * W > (1ll << 31)) return RCF_ERR_SHAPE;
    if (nprod < 1 || nprod > 3) return RCF_ERR_MODE;
    if ((((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)g_hi | (uintptr_t)g_lo | (uintptr_t)dw | (uintptr_t)ws) & 15) != 0) return RCF_ERR_ALIGN;
    WgradArgs a;
    a.g = wgrad_geom(nimg, H, W, nprod);
    a.part = (float*)ws;
    a.fmt = fmt;
    a.status = rcf_conv64_status_addr();
    if (!a.status) return (int)cudaErrorInvalidSymbol;
    alignas(64) CUtensorMap xh, xl, gh, gl;
    int e = rcf_make_tmap_nhwc64(&xh, x_hi, nimg, H, W, a.g.L + 2, a.g.TR + 2);
    if (e == RCF_OK) e = rcf_make_tmap_nhwc64(&xl, nprod == 3 ? x_lo : x_hi, nimg, H, W, a.g.L + 2, a.g.TR + 2);
    if (e == RCF_OK) e = rcf_make_tmap_nhwc64(&gh, g_hi, nimg, H, W, a.g.L + 2, a.g.TR);
    if (e == RCF_OK) e = rcf_make_tmap_nhwc64(&gl, nprod >= 2 ? g_lo : g_hi, nimg, H, W, a.g.L + 2, a.g.TR);
    if (e != RCF_OK) return e;
    const int grid = wgrad_grid(a.g.ntiles);
    const size_t smem = (size_t)WG_STAGES * a.g.stage_bytes + 256;
    cudaStream_t s = (cudaStream_t)stream;
    if (nprod == 1) e = launch_wgrad<1>(a, xh, xl, gh, gl, grid, smem, s);
    else if (nprod == 2) e = launch_wgrad<2>(a, xh, xl, gh, gl, grid, smem, s);
    else e = launch_wgrad<3>(a, xh, xl, gh, gl, grid, smem, s);
    if (e != RCF_OK) return e;
    k_conv64_wgrad_reduce<<<(64 * 64 * 9 + 255) / 256, 256, 0, s>>>((const float*)ws, grid, dw, gmax, nfd);
    return (int)cudaGetLastError();
}
