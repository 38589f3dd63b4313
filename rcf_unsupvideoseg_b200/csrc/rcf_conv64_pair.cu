// rcf_conv64_pair.cu -- the tcgen05 3x3 conv (rcf_conv64.cu) on CTA PAIRS: clusters of two CTAs on the two SMs of a TPC
// issue ONE tcgen05.mma.cta_group::2 per tap and K-step, M = 256 (128 output positions per CTA).  Each CTA keeps its own
// activation tile and only HALF of the weight rows in shared memory:
//   * per-SM operand traffic of an M256 x N128 x K16 MMA: 4 KB (A) + 2 KB (half of B) per 64 clk = 96 B/clk, under the
//     128 B/clk of shared-memory bandwidth that the one-CTA kernel saturates (4 + 4 KB per 64 clk, plus the TMA fills);
//   * 72 KB instead of 144 KB of weights per SM: room for a 4-stage ring of activation tiles (and a 2-stage ring of
//     (hi, lo) tiles in the fp32-grade mode, which the one-CTA kernel has to run single-buffered).
// Per-CTA weight image (k_conv64_pack): region Y = 9 taps x 64 rows (rank 0: the bf16 "hi" words of the 64 output
// channels, rank 1: the "lo" words) -> the N = 128 MMA A_hi x [W_hi | W_lo];  region Z = 9 taps x 32 rows (hi words of output
// channels [32 rank, +32)) -> the N = 64 MMA of the plain-bf16 mode and of the third product A_lo x W_hi.
//
// Protocol: every CTA's TMA thread loads its own tile and counts the bytes on the LEADER's (even CTA's) full barrier; the
// leader's MMA thread waits for both tiles, issues, and commits with a multicast arrive on both CTAs' empty / accumulator-
// full barriers; the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
#include <cuda.h>

#include "rcf_common.cuh"
#include "rcf_conv64.cuh"
#include "rcf_umma.cuh"

int rcf_make_tmap_nhwc64(CUtensorMap* tm, const void* base, int nimg, int H, int W, int bw, int bh);
int* rcf_conv64_status_addr();
int rcf_conv64_debug_flags();

namespace {
using namespace umma;

constexpr int P_MAXSTAGE = 5, P_NT = 4;
// 18 warps: two groups of 8 epilogue warps (TMEM lane quarter x column half) that take alternate tiles -- one group cannot
// drain a tile (TMEM loads, register/lane exchange, stores: ~2700 clk of dependent latency) in the ~2300 clk the pair's MMAs
// take -- then the MMA-issue warp and the TMA warp.
constexpr int P_EPI_GROUPS = 2, P_EPI_WARPS = 8 * P_EPI_GROUPS, P_MMA_WARP = P_EPI_WARPS, P_TMA_WARP = P_EPI_WARPS + 1;
constexpr int P_THREADS = (P_EPI_WARPS + 2) * 32;

struct PairArgs {
    Conv64Geom g;
    float* out;
    const uint8_t* wpack;      // pair images: [rank][C64_PAIR_IMAGE_BYTES]
    int* status;
    int npairs;                // ceil(ntiles / 2)
    int debug;                 // measurement switch: 1 no epilogue stores
    uint32_t fmt;              // operand formats: bit 0 activations fp16, bit 1 weights fp16 (else bf16)
};

struct PairBars {
    uint64_t full[P_MAXSTAGE], empty[P_MAXSTAGE], tfull[P_NT], tempty[P_NT], wbar, wpeer;
    uint32_t tmem_base, abort_flag;
};

__device__ __forceinline__ bool pwait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag) {
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

template <int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
k_conv64_pair(const PairArgs a, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // weights: Y (72 KB) for NPROD >= 2, Z (36 KB) for NPROD 1 and 3
    constexpr int WY = NPROD >= 2 ? C64_PAIR_Y_BYTES : 0, WZ = NPROD != 2 ? C64_PAIR_Z_BYTES : 0;
    constexpr int NA = NPROD == 3 ? 2 : (NPROD == 2 ? 4 : 5);          // ring stages
    constexpr int STAGE = (NPROD == 3 ? 2 : 1) * C64_ABUF_BYTES;        // NPROD 3: hi tile then lo tile
    constexpr int NCOL = NPROD == 1 ? 64 : 128;                        // accumulator columns per stage
    uint8_t* const sY = smem;
    uint8_t* const sZ = smem + WY;
    uint8_t* const sA = smem + WY + WZ;
    PairBars* const bars = reinterpret_cast<PairBars*>(sA + NA * STAGE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Conv64Geom& g = a.g;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < P_NT; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 2 * 8); }
        mbar_init(&bars->wbar, 1);
        mbar_init(&bars->wpeer, 1);
        bars->abort_flag = (smem_u32(smem) & 1023u) ? 1u : 0u;
        mbar_init_fence();
    }
    if (warp == P_MMA_WARP) tmem_alloc_pair<512>(&bars->tmem_base);
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();                       // both CTAs' barriers are initialised before anything remote touches them
    fence_after_sync();
    const uint32_t tmem = bars->tmem_base;
    volatile uint32_t* abort_flag = &bars->abort_flag;
    const int tiles_per_img = g.tiles_x * g.tiles_y;

    if (warp == P_MMA_WARP) {
        // ===== weights of this CTA (both ranks), then the leader issues the MMAs of the pair =====
        if (lane == 0) {
            const uint8_t* img = a.wpack + (size_t)rank * C64_PAIR_IMAGE_BYTES;
            mbar_arrive_expect_tx(&bars->wbar, WY + WZ);
            if (WY) for (int t = 0; t < 9; ++t) bulk_g2s(sY + t * 8192, img + t * 8192, 8192, &bars->wbar);
            if (WZ) for (int t = 0; t < 9; ++t) bulk_g2s(sZ + t * 4096, img + C64_PAIR_Y_BYTES + t * 4096, 4096, &bars->wbar);
        }
        __syncwarp();
        pwait(&bars->wbar, 0, abort_flag);
        // the peer's weights must have landed too before the leader's MMAs read them: the odd CTA reports on the leader's barrier
        if (rank == 1 && lane == 0)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(&bars->wpeer) & PEER_BIT_MASK) : "memory");
        if (rank == 0) {
            pwait(&bars->wpeer, 0, abort_flag);
            const uint32_t IDESC64 = idesc_with_formats(make_idesc_bf16(256, 64, 0, 0), a.fmt);
            const uint32_t IDESC128 = idesc_with_formats(make_idesc_bf16(256, 128, 0, 0), a.fmt);
            const uint64_t ydesc = make_desc_sw128(smem_u32(sY), 16, 1024), zdesc = make_desc_sw128(smem_u32(sZ), 16, 1024);
            const uint32_t Wp8 = (uint32_t)g.Wp * 8;
            int it = 0;
            for (int pair = cluster; pair < a.npairs; pair += nclusters, ++it) {
                const int b = it % NA, ph = (it / NA) & 1, ts = it % P_NT, tph = (it / P_NT) & 1;
                pwait(&bars->full[b], ph, abort_flag);
                pwait(&bars->tempty[ts], tph ^ 1, abort_flag);
                fence_after_sync();
                if (elect_one()) {
                    const uint64_t adesc = make_desc_sw128(smem_u32(sA + b * STAGE), 16, 1024);
                    const uint64_t adesc_lo = make_desc_sw128(smem_u32(sA + b * STAGE + C64_ABUF_BYTES), 16, 1024);
                    const uint32_t dcol = tmem + ts * NCOL;
#pragma unroll
                    for (int ty = 0; ty < 3; ++ty) {
                        const uint64_t ar = adesc + (uint64_t)(ty * Wp8), arl = adesc_lo + (uint64_t)(ty * Wp8);
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t aoff = tx * 8 + ks * 2, tap = ty * 3 + tx;
                                const uint32_t yoff = tap * (8192 >> 4) + ks * 2, zoff = tap * (4096 >> 4) + ks * 2;
                                const uint32_t acc = (ty | tx | ks) != 0;
                                if (NPROD == 1) mma_bf16_pair(dcol, ar + aoff, zdesc + zoff, IDESC64, acc);
                                else mma_bf16_pair(dcol, ar + aoff, ydesc + yoff, IDESC128, acc);
                                // third product: A_lo x W_hi; the hi columns of the N = 128 layout are 0-63 (CTA 0's rows),
                                // and an N = 64 MMA over the Z rows writes columns 0-31 (CTA 0) and 32-63 (CTA 1): same channels
                                if (NPROD == 3) mma_bf16_pair(dcol, arl + aoff, zdesc + zoff, IDESC64, 1);
                            }
                    }
                    mma_commit_pair(&bars->empty[b]);
                    mma_commit_pair(&bars->tfull[ts]);
                }
                __syncwarp();
            }
        }
    } else if (warp == P_TMA_WARP) {
        // ===== producer: this CTA's tile, bytes counted on the leader's full barrier =====
        if (lane == 0) {
            tma_prefetch_desc(&tm_hi);
            const uint32_t bytes = (uint32_t)g.npos * 128u * (NPROD == 3 ? 2u : 1u);
            int it = 0;
            for (int pair = cluster; pair < a.npairs; pair += nclusters, ++it) {
                const int b = it % NA, ph = (it / NA) & 1;
                const int tile = pair * 2 + (int)rank;
                int img = tile / tiles_per_img;
                const int trem = tile - img * tiles_per_img;
                const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
                if (tile >= g.ntiles) img = g.nimg;                       // odd tail: a box outside the tensor = zeros
                pwait(&bars->empty[b], ph ^ 1, abort_flag);
                if (rank == 0) mbar_arrive_expect_tx(&bars->full[b], 2 * bytes);   // both CTAs' tiles
                tma_load_4d_pair(sA + b * STAGE, &tm_hi, 0, txi * g.TW - 1, tyi * g.TR - 1, img, &bars->full[b]);
                if (NPROD == 3)
                    tma_load_4d_pair(sA + b * STAGE + C64_ABUF_BYTES, &tm_lo, 0, txi * g.TW - 1, tyi * g.TR - 1, img, &bars->full[b]);
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM -> registers -> channels-last fp32 =====
        const int quarter = warp & 3, c0 = ((warp >> 2) & 1) * 32, group = warp >> 3;
        const int debug = a.debug;
        int it = 0;
        for (int pair = cluster; pair < a.npairs; pair += nclusters, ++it) {
            if ((it % P_EPI_GROUPS) != group) continue;
            const int ts = it % P_NT, tph = (it / P_NT) & 1;
            const int tile = pair * 2 + (int)rank;
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int tyi = trem / g.tiles_x, txi = trem - tyi * g.tiles_x;
            float* dst = a.out + (size_t)img * g.H * g.W * 64;
            pwait(&bars->tfull[ts], tph, abort_flag);
            fence_after_sync();
            {
                const int q = g.Wp + 1 + quarter * 32 + lane;
                const int r = (int)__umulhi((uint32_t)q, g.wp_magic) - 1, x = q - (r + 1) * g.Wp - 1;
                const int y = tyi * g.TR + r, xx = txi * g.TW + x;
                const int mypix = (tile < g.ntiles && x >= 0 && x < g.TW && r < g.TR && y < g.H && xx < g.W && !(debug & 1)) ? y * g.W + xx : -1;
                int pix[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pix[i] = __shfl_sync(0xffffffffu, mypix, 4 * i + (lane & 3));
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + ts * NCOL + c0;
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
                // accumulators are in registers: the TMEM stage can be handed back before the stores
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bars->tempty[ts]);
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
        }
    }

    fence_before_sync();
    __syncthreads();
    cluster_sync_all();                       // nobody leaves while the peer may still signal this CTA's barriers
    if (warp == P_MMA_WARP) tmem_dealloc_pair<512>(tmem);
    if (tid == 0 && bars->abort_flag && a.status) *a.status = 1;
}

template <int NPROD>
int launch_pair(const PairArgs& a, const CUtensorMap& hi, const CUtensorMap& lo, int grid, cudaStream_t s) {
    constexpr int WY = NPROD >= 2 ? C64_PAIR_Y_BYTES : 0, WZ = NPROD != 2 ? C64_PAIR_Z_BYTES : 0;
    constexpr int NA = NPROD == 3 ? 2 : (NPROD == 2 ? 4 : 5);
    constexpr int STAGE = (NPROD == 3 ? 2 : 1) * C64_ABUF_BYTES;
    constexpr size_t smem = WY + WZ + NA * STAGE + 256;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static bool attr_done = false;
    if (!attr_done) {
        const cudaError_t e = cudaFuncSetAttribute(k_conv64_pair<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    k_conv64_pair<NPROD><<<grid, P_THREADS, smem, s>>>(a, hi, lo);
    return (int)cudaGetLastError();
}

}  // namespace

int rcf_conv64_pair_launch(const void* in_hi, const void* in_lo, const void* wpack_pair, float* out, int nimg, int H, int W,
                           int nprod, uint32_t fmt, cudaStream_t s) {
    PairArgs a;
    a.fmt = fmt;
    a.g = conv64_make_geom(nimg, H, W);
    a.out = out; a.wpack = (const uint8_t*)wpack_pair;
    a.npairs = (a.g.ntiles + 1) / 2;
    a.debug = rcf_conv64_debug_flags();
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
    int nclusters = nsm / 2;
    if (a.npairs < nclusters) nclusters = a.npairs;
    const int grid = 2 * nclusters;
    if (nprod == 1) return launch_pair<1>(a, tm_hi, tm_lo, grid, s);
    if (nprod == 2) return launch_pair<2>(a, tm_hi, tm_lo, grid, s);
    return launch_pair<3>(a, tm_hi, tm_lo, grid, s);
}
