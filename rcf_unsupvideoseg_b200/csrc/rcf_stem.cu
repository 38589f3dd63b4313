// rcf_stem.cu -- the first layer of flow_feat_before_agg as hand-written kernels:
//     act = LeakyReLU(conv_{ks x ks, pad (ks-1)/2}(clamp(flow)) + bias)          reference :84-88, :155-156, :246
// A 2-channel ks x ks convolution has 2*ks*ks = 18 inputs per output: it is bandwidth-bound (8 B/px in, 4*Cf B/px out),
// not a tensor-core shape, and cuDNN runs it (plus a separate bias add, LeakyReLU, LeakyReLU backward, a weight-gradient
// GEMM and a bias-gradient reduction) as six launches with layout conversions around them.  Here:
//   k_stem_fwd  : clamp + conv + bias + LeakyReLU in one pass; output channels-last ([N,H,W,Cf]), which is what the
//                 second (cuDNN tensor-core) convolution wants.  A thread owns 4 output channels (72 weights in registers
//                 for ks = 3) and walks the pixels of a 32x8 tile whose clamped flow (+halo) is staged in shared memory.
//   k_stem_bwd  : LeakyReLU backward + weight gradient + bias gradient in one pass over (act, dact); per-thread
//                 accumulators, fixed-order combination inside the CTA, per-CTA partials reduced by k_stem_bwd_final
//                 in fp64 => bit-reproducible.  No input gradient: the RAFT flow carries no grad.
#include "rcf_common.cuh"
#include "rcf_umma.cuh"
#include "rcf_internal.h"

namespace {

struct StemK {
    const float* flow[2];
    long long flow_bs[2];
    int ndir, B, H, W, P, Cf;
    float clamp_t, slope;
    const float* w;      // [Cf][2][KS][KS]
    const float* b;      // [Cf]
    float* act;          // [N][P][Cf]
    uint32_t* act_hi;    // tensor-core path, optional: the activation as bf16 pairs [N][P][Cf/2] (hi word; lo word = act - hi)
    uint32_t* act_lo;    //   instead of the fp32 map: what the tcgen05 conv (rcf_conv64.cu) loads by TMA
    int act_f16;         // act_hi holds IEEE fp16 words (saturating) instead of bf16: the single-product TF32-class conv mode
    const float* act_in; // backward: forward output (sign of the pre-activation) -- CUDA-core path
    uint32_t* sign_out;  // tensor-core path: [N][P][Cf/32] bit f%32 of word f/32 set <=> pre-activation of channel f is <= 0
    const uint32_t* sign_in;
    const float* dact;   // [N][P][Cf]
    float* part;         // [gridDim.x][Cf*(NT+1)]
    int ntiles;          // N * tiles per image
};

// A CTA works on a TW x TH pixel tile; the clamped flow tile plus its (KS-1)/2 halo (zero outside the frame = the
// conv's zero padding) is staged once in shared memory and every thread reads its 2*KS*KS taps from there
// (the Cf/4 threads of a pixel hit the same words: broadcast).
constexpr int STEM_TW = 32, STEM_TH = 8;

template <int KS>
__device__ __forceinline__ void stage_tile(const StemK& a, const float* __restrict__ fl, int x0, int y0,
                                           float (*tile)[STEM_TH + KS - 1][STEM_TW + KS - 1]) {
    constexpr int R = (KS - 1) / 2, SW = STEM_TW + KS - 1, SH = STEM_TH + KS - 1;
    for (int i = threadIdx.x; i < 2 * SH * SW; i += RCF_BLOCK) {
        const int c = i / (SH * SW), r = i - c * SH * SW;
        const int ly = r / SW, lx = r - ly * SW;
        const int y = y0 + ly - R, x = x0 + lx - R;
        float v = 0.0f;
        if (y >= 0 && y < a.H && x >= 0 && x < a.W) v = clamp_flow(__ldg(fl + (long long)c * a.P + (long long)y * a.W + x), a.clamp_t);
        tile[c][ly][lx] = v;
    }
}

template <int KS>
__device__ __forceinline__ void read_taps(const float (*tile)[STEM_TH + KS - 1][STEM_TW + KS - 1], int ly, int lx,
                                          float (&tap)[2 * KS * KS]) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) tap[(c * KS + dy) * KS + dx] = tile[c][ly + dy][lx + dx];
}

template <int KS>
__global__ void __launch_bounds__(RCF_BLOCK) k_stem_fwd(const StemK a) {
    constexpr int NT = 2 * KS * KS;
    __shared__ float tile[2][STEM_TH + KS - 1][STEM_TW + KS - 1];
    const int n = blockIdx.z;                       // image = dir * B + b
    const int dir = n / a.B, b = n - dir * a.B;
    const int tid = threadIdx.x, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int x0 = blockIdx.x * STEM_TW, y0 = blockIdx.y * STEM_TH;
    const float* __restrict__ fl = (dir ? a.flow[1] : a.flow[0]) + (long long)b * (dir ? a.flow_bs[1] : a.flow_bs[0]);
    stage_tile<KS>(a, fl, x0, y0, tile);
    float w[4][NT], bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bias[j] = __ldg(a.b + c4 * 4 + j);
#pragma unroll
        for (int t = 0; t < NT; ++t) w[j][t] = __ldg(a.w + (size_t)(c4 * 4 + j) * NT + t);
    }
    __syncthreads();
    float4* __restrict__ out = reinterpret_cast<float4*>(a.act + (long long)n * a.P * Cf) + c4;
#pragma unroll 2
    for (int p = grp; p < STEM_TW * STEM_TH; p += groups) {
        const int ly = p / STEM_TW, lx = p - ly * STEM_TW;
        const int y = y0 + ly, x = x0 + lx;
        if (y < a.H && x < a.W) {
            float tap[NT];
            read_taps<KS>(tile, ly, lx, tap);
            float o[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fmaf(w[j][t], tap[t], o[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = o[j] >= 0.0f ? o[j] : a.slope * o[j];
            out[((long long)y * a.W + x) * nf4] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(RCF_BLOCK) k_stem_bwd(const StemK a) {
    constexpr int NT = 2 * KS * KS, NO = NT + 1;
    __shared__ float tile[2][STEM_TH + KS - 1][STEM_TW + KS - 1];
    extern __shared__ float red[];                   // [RCF_WARPS][Cf * NO]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int tx = (a.W + STEM_TW - 1) / STEM_TW, ty = (a.H + STEM_TH - 1) / STEM_TH;
    float dw[4][NT], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        db[j] = 0.0f;
#pragma unroll
        for (int t = 0; t < NT; ++t) dw[j][t] = 0.0f;
    }
    for (int tl = blockIdx.x; tl < a.ntiles; tl += gridDim.x) {      // fixed tile -> CTA assignment (reproducible)
        const int n = tl / (tx * ty), r = tl - n * tx * ty;
        const int y0 = (r / tx) * STEM_TH, x0 = (r - (r / tx) * tx) * STEM_TW;
        const int dir = n / a.B, b = n - dir * a.B;
        const float* __restrict__ fl = (dir ? a.flow[1] : a.flow[0]) + (long long)b * (dir ? a.flow_bs[1] : a.flow_bs[0]);
        __syncthreads();                               // previous tile fully consumed
        stage_tile<KS>(a, fl, x0, y0, tile);
        __syncthreads();
        const float4* __restrict__ ap = reinterpret_cast<const float4*>(a.act_in + (long long)n * a.P * Cf) + c4;
        const float4* __restrict__ gp = reinterpret_cast<const float4*>(a.dact + (long long)n * a.P * Cf) + c4;
#pragma unroll 2
        for (int p = grp; p < STEM_TW * STEM_TH; p += groups) {
            const int ly = p / STEM_TW, lx = p - ly * STEM_TW;
            const int y = y0 + ly, x = x0 + lx;
            if (y < a.H && x < a.W) {
                const long long o = ((long long)y * a.W + x) * nf4;
                const float4 av = __ldg(ap + o), gv = __ldg(gp + o);
                float dpre[4] = {gv.x * (av.x >= 0.0f ? 1.0f : a.slope), gv.y * (av.y >= 0.0f ? 1.0f : a.slope),
                                 gv.z * (av.z >= 0.0f ? 1.0f : a.slope), gv.w * (av.w >= 0.0f ? 1.0f : a.slope)};
                float tap[NT];
                read_taps<KS>(tile, ly, lx, tap);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    db[j] += dpre[j];
#pragma unroll
                    for (int t = 0; t < NT; ++t) dw[j][t] = fmaf(dpre[j], tap[t], dw[j][t]);
                }
            }
        }
    }
    // lanes l, l + nf4, ... of a warp own the same channels: combine them with shuffles (nf4 <= 32, power of two)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        for (int o = 16; o >= nf4; o >>= 1) db[j] += __shfl_xor_sync(0xffffffffu, db[j], o);
#pragma unroll
        for (int t = 0; t < NT; ++t)
            for (int o = 16; o >= nf4; o >>= 1) dw[j][t] += __shfl_xor_sync(0xffffffffu, dw[j][t], o);
    }
    // every warp now holds all channel quads in its first nf4 lanes: combine the 8 warps through shared memory
    if (lane < nf4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* r = red + (size_t)warp * Cf * NO + (size_t)(lane * 4 + j) * NO;
#pragma unroll
            for (int t = 0; t < NT; ++t) r[t] = dw[j][t];
            r[NT] = db[j];
        }
    }
    __syncthreads();
    for (int o = tid; o < Cf * NO; o += RCF_BLOCK) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < RCF_WARPS; ++wi) v += red[(size_t)wi * Cf * NO + o];
        a.part[(size_t)blockIdx.x * Cf * NO + o] = v;
    }
}


// ================================================================================================================
// Tensor-core path for Cf == 64 (the reference default, configs/*/rcf_stage*.yaml never change it).
// The CUDA-core kernels above are issue-bound: 72 FFMA + 18 LDS per thread per pixel (measured 225 us forward / 340 us
// backward at 2x2x480x854 against HBM floors of 65 / 130 us).  Both passes are small GEMMs over the pixel axis,
//     forward   act[p, f]  = sum_t tap[p, t] * w[t, f]          [P x 19] . [19 x 64]   (tap 18 == 1: the bias)
//     backward  dw [f, t]  = sum_p dpre[p, f] * tap[p, t]       [64 x P] . [P x 19]    (tap 18 == 1: the bias gradient)
// so they run on mma.sync.m16n8k8 TF32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi, fp32 accumulate: ~2^-21 relative,
// i.e. fp32-grade -- plain TF32 would break the 1e-4 parity bar).  That is 36 MMA per 16 pixels x 32 channels instead of
// 576 FFMA.  This is NOT a tcgen05 shape: K is 19 and the kernels are HBM-bound once the issue pressure is gone.
// Rows / columns of the MMA tiles are PERMUTED channel indices so that every lane owns 8 consecutive channels of a
// pixel: 128-bit global loads / stores in the native channels-last layout, no shuffles, no shared-memory transpose.
// The forward also emits one bit per (pixel, channel) -- the sign of the pre-activation -- and the backward reads those
// 8 B/px instead of re-reading the 256 B/px activation map.
// ================================================================================================================
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d += a * b with a, b given as (hi, lo) TF32 pairs; small terms first
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(d, al, bh[0], bh[1]);
    mma_tf32(d, ah, bl[0], bl[1]);
    mma_tf32(d, ah, bh[0], bh[1]);
}

// cheap split for values produced in registers: hi = truncation (1 LOP3), lo = x - hi (exact), lo rounded to TF32 by
// adding half an ulp to its bit pattern (the MMA ignores the low 13 bits).  |x - hi - lo'| <= 2^-21 |x|.
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}

// Shared-memory word of tap `ti` relative to a pixel's tile position: (offset, multiplier of the pixel offset).
// ti < NT: the conv tap; ti == NT: a constant 1 (bias column); above: a constant 0 (padding of the K / N dimension).
template <int KS>
__device__ __forceinline__ void tap_addr(int ti, int& off, int& mul) {
    constexpr int NT = 2 * KS * KS, SW = STEM_TW + KS - 1, SH = STEM_TH + KS - 1;
    if (ti < NT) {
        const int c = ti / (KS * KS), r = ti - c * KS * KS, dy = r / KS, dx = r - dy * KS;
        off = (c * SH + dy) * SW + dx; mul = 1;
    } else {
        off = 2 * SH * SW + (ti == NT ? 0 : 1); mul = 0;
    }
}

constexpr int STEM_MMA_CF = 64;

// The clamped flow tile (+halo, zero outside the frame) is kept in shared memory already split into TF32 (hi, lo)
// words, so the MMA operands that come from it cost two LDS and no ALU work.  The loads of the NEXT tile are issued
// before the current tile is computed and committed to shared memory afterwards (software pipeline: the HBM latency of
// this read-once input is hidden behind the MMAs instead of being paid at a barrier).
// tile index -> (image, tile origin) without integer divisions (they cost ~25 instructions each, per thread and tile):
// float reciprocal + one correction step, exact for the < 2^22 tiles a launch can have
struct TileDecode {
    int tx, tpi;
    float inv_tx, inv_tpi;
    __device__ __forceinline__ TileDecode(int tx_, int ty_) : tx(tx_), tpi(tx_ * ty_), inv_tx(1.0f / (float)tx_), inv_tpi(1.0f / (float)(tx_ * ty_)) {}
    static __device__ __forceinline__ int fdiv(int x, int d, float inv) {
        int q = (int)((float)x * inv);
        q -= (q * d > x) ? 1 : 0;
        q += ((q + 1) * d <= x) ? 1 : 0;
        return q;
    }
    __device__ __forceinline__ void operator()(int tl, int& n, int& y0, int& x0) const {
        n = fdiv(tl, tpi, inv_tpi);
        const int r = tl - n * tpi, yi = fdiv(r, tx, inv_tx);
        y0 = yi * STEM_TH; x0 = (r - yi * tx) * STEM_TW;
    }
};

template <int KS>
struct StemTile {
    static constexpr int SW = STEM_TW + KS - 1, SH = STEM_TH + KS - 1, TSZ = 2 * SH * SW, R = (KS - 1) / 2;
    static constexpr int NE = (TSZ + RCF_BLOCK - 1) / RCF_BLOCK;
    float pre[NE];
    __device__ __forceinline__ void fetch(const StemK& a, int n, int y0, int x0) {
        const int dir = n >= a.B ? 1 : 0, b = n - dir * a.B;
        const float* __restrict__ fl = (dir ? a.flow[1] : a.flow[0]) + (long long)b * (dir ? a.flow_bs[1] : a.flow_bs[0]);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int i = threadIdx.x + e * RCF_BLOCK;
            const int c = i / (SH * SW), q = i - c * SH * SW;
            const int ly = q / SW, lx = q - ly * SW;
            const int y = y0 + ly - R, x = x0 + lx - R;
            float v = 0.0f;
            if (i < TSZ && y >= 0 && y < a.H && x >= 0 && x < a.W) v = __ldg(fl + (long long)c * a.P + (long long)y * a.W + x);
            pre[e] = v;
        }
    }
    template <int NP>
    __device__ __forceinline__ void commit(const StemK& a, uint32_t* Thi, uint32_t* Tlo) const {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int i = threadIdx.x + e * RCF_BLOCK;
            if (i < TSZ) {
                if constexpr (NP == 1) Thi[i] = to_tf32(clamp_flow(pre[e], a.clamp_t));
                else split_tf32(clamp_flow(pre[e], a.clamp_t), Thi[i], Tlo[i]);
            }
        }
    }
};

__device__ __forceinline__ void mma_tf32_k4(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// sign bits of four fp32 values ("v <= 0", -0.0 excepted) as a 4-bit field: top bytes of (bits - 1) packed with two
// byte permutes, masked to the sign positions and gathered with one multiply.
__device__ __forceinline__ uint32_t nonpos4(float v0, float v1, float v2, float v3) {
    const uint32_t u0 = __float_as_uint(v0) - 1u, u1 = __float_as_uint(v1) - 1u;
    const uint32_t u2 = __float_as_uint(v2) - 1u, u3 = __float_as_uint(v3) - 1u;
    const uint32_t lo = __byte_perm(u0, u1, 0x0073);       // byte0 = top(u0), byte1 = top(u1)
    const uint32_t hi = __byte_perm(u2, u3, 0x7300);       // byte2 = top(u2), byte3 = top(u3)
    const uint32_t w = __byte_perm(lo, hi, 0x7610) & 0x80808080u;   // sign of value i at bit 8i+7
    return (w * 0x00204081u) >> 28;                          // bit i = sign of value i
}

// Forward.  CTA = 8 warps on a 32x8 tile = 16 M-tiles of 16 pixels (half a tile row); warp w computes channels
// [32*(w&1), +32) of M-tiles (w>>1) + 4i.  Column c of N-tile j is channel 32*half + 16*(j>>1) + 4*(c>>1) + 2*(j&1) + (c&1):
// lane (g = lane/4, t = lane%4) ends up with channels 4t..4t+3 (N-tiles 0,1) and 16+4t..16+4t+3 (N-tiles 2,3) of the
// pixels g and g+8, so each 128-bit store instruction of the warp writes 64 contiguous bytes per pixel (full sectors).
// K = taps + 1 (bias column): full k8 steps plus one k4 step for a remainder <= 4 (3x3: 19 = 8 + 8 + 3).
// The scheduler's issue slots are shared by the MMAs and everything else (measured: MMA time + other-instruction time
// add up), so the epilogue is kept short: LeakyReLU as max(v, slope*v), sign bits by byte permutes, one address per tile.
// NP = TF32 products per fp32 product: 3 (hi*hi + hi*lo + lo*hi, fp32-grade) or 1 (plain TF32 with round-to-nearest
// operands -- what cuDNN runs for this layer under torch.backends.cudnn.allow_tf32; a third of the MMAs, half the LDS).
template <int KS, int NP>
__global__ void __launch_bounds__(RCF_BLOCK, (KS <= 3) ? 2 : 1) k_stem_fwd_mma(const StemK a) {
    using TL = StemTile<KS>;
    constexpr int NT = 2 * KS * KS, NK = NT + 1, SW = TL::SW, TSZ = TL::TSZ;
    constexpr int REM = NK % 8;
    constexpr bool K4 = REM > 0 && REM <= 4;
    constexpr int NK8 = K4 ? NK / 8 : (NK + 7) / 8;
    constexpr int NK8A = NK8 > 0 ? NK8 : 1;
    constexpr int Cf = STEM_MMA_CF;
    __shared__ uint32_t Thi[TSZ + 2], Tlo[TSZ + 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int half = warp & 1, mset = warp >> 1;
    const int tx = (a.W + STEM_TW - 1) / STEM_TW, ty = (a.H + STEM_TH - 1) / STEM_TH;

    auto wval = [&](int ch, int k) -> float {
        return k < NT ? __ldg(a.w + (size_t)ch * NT + k) : (k == NT ? __ldg(a.b + ch) : 0.0f);
    };
    uint32_t bh[4][NK8A][2], bl[4][NK8A][2], b4h[4], b4l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = half * 32 + (j >> 1) * 16 + (g >> 1) * 4 + 2 * (j & 1) + (g & 1);
#pragma unroll
        for (int s = 0; s < NK8; ++s)
#pragma unroll
            for (int e = 0; e < 2; ++e) split_tf32(wval(ch, 8 * s + t + 4 * e), bh[j][s][e], bl[j][s][e]);
        if (K4) split_tf32(wval(ch, 8 * NK8 + t), b4h[j], b4l[j]);
    }
    int offA[NK8A][2], mulA[NK8A][2], off4 = 0, mul4 = 0;
#pragma unroll
    for (int s = 0; s < NK8; ++s)
#pragma unroll
        for (int e = 0; e < 2; ++e) tap_addr<KS>(8 * s + t + 4 * e, offA[s][e], mulA[s][e]);
    if (K4) tap_addr<KS>(8 * NK8 + t, off4, mul4);
    if (tid == 0) { Thi[TSZ] = __float_as_uint(1.0f); Tlo[TSZ] = 0u; Thi[TSZ + 1] = 0u; Tlo[TSZ + 1] = 0u; }

    TL st;
    const TileDecode dec(tx, ty);
    int tl = blockIdx.x, n = 0, y0 = 0, x0 = 0, nn = 0, ny0 = 0, nx0 = 0;
    if (tl < a.ntiles) { dec(tl, nn, ny0, nx0); st.fetch(a, nn, ny0, nx0); }
    for (; tl < a.ntiles; tl += gridDim.x) {
        n = nn; y0 = ny0; x0 = nx0;
        __syncthreads();                               // previous tile fully consumed
        st.template commit<NP>(a, Thi, Tlo);
        __syncthreads();
        if (tl + (int)gridDim.x < a.ntiles) {          // next tile's loads in flight while this tile is computed
            dec(tl + gridDim.x, nn, ny0, nx0);
            st.fetch(a, nn, ny0, nx0);
        }
        float* const img = a.act + (long long)n * a.P * Cf + half * 32 + 4 * t;
        uint32_t* const simg = a.sign_out ? a.sign_out + (long long)n * a.P * 2 + half : nullptr;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int mt = mset + 4 * i, ly = mt >> 1, lx0 = (mt & 1) * 16;
            const int y = y0 + ly;
            if (y >= a.H || x0 + lx0 >= a.W) continue;          // warp-uniform
            const int pix0 = ly * SW + lx0 + g, pix1 = pix0 + 8;
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) c[j][q] = 0.0f;
#pragma unroll
            for (int s = 0; s < NK8; ++s) {
                const int i0 = offA[s][0] + pix0 * mulA[s][0], i1 = offA[s][0] + pix1 * mulA[s][0];
                const int i2 = offA[s][1] + pix0 * mulA[s][1], i3 = offA[s][1] + pix1 * mulA[s][1];
                const uint32_t ah[4] = {Thi[i0], Thi[i1], Thi[i2], Thi[i3]};
                if constexpr (NP == 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32(c[j], ah, bh[j][s][0], bh[j][s][1]);
                } else {
                    const uint32_t al[4] = {Tlo[i0], Tlo[i1], Tlo[i2], Tlo[i3]};
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_3xtf32(c[j], ah, al, bh[j][s], bl[j][s]);
                }
            }
            if (K4) {
                const int i0 = off4 + pix0 * mul4, i1 = off4 + pix1 * mul4;
                const uint32_t ah0 = Thi[i0], ah1 = Thi[i1];
                if constexpr (NP == 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32_k4(c[j], ah0, ah1, b4h[j]);
                } else {
                    const uint32_t al0 = Tlo[i0], al1 = Tlo[i1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        mma_tf32_k4(c[j], al0, al1, b4h[j]);
                        mma_tf32_k4(c[j], ah0, ah1, b4l[j]);
                        mma_tf32_k4(c[j], ah0, ah1, b4h[j]);
                    }
                }
            }
            const int x = x0 + lx0 + g;
            const int pofs = y * a.W + x;                        // pixel of row g; row g+8 is 8 pixels further
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // o[0..3]: channels 4t..4t+3, o[4..7]: 16+4t..16+4t+3 (of this warp's 32)
                const float v[8] = {c[0][2 * h], c[0][2 * h + 1], c[1][2 * h], c[1][2 * h + 1],
                                    c[2][2 * h], c[2][2 * h + 1], c[3][2 * h], c[3][2 * h + 1]};
                uint32_t word = (nonpos4(v[0], v[1], v[2], v[3]) << (4 * t)) | (nonpos4(v[4], v[5], v[6], v[7]) << (16 + 4 * t));
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                if (x + 8 * h < a.W) {
                    const float sl = a.slope;                    // 0 <= slope <= 1 (checked by the launcher): lrelu = max(v, slope*v)
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[e], sl * v[e]);
                    if (a.act_hi) {                              // bf16 (hi, lo) pair: 4 channels = 2 words, +16 channels = +8 words
                        const long long wofs = ((long long)n * a.P + pofs + 8 * h) * (Cf / 2) + half * 16 + 2 * t;
                        uint32_t hw[4], lw[4];
                        if (a.act_lo) {                          // warp-uniform
#pragma unroll
                            for (int e = 0; e < 4; ++e) umma::split_bf16x2(o[2 * e], o[2 * e + 1], hw[e], lw[e]);
                            *reinterpret_cast<uint2*>(a.act_lo + wofs) = make_uint2(lw[0], lw[1]);
                            *reinterpret_cast<uint2*>(a.act_lo + wofs + 8) = make_uint2(lw[2], lw[3]);
                        } else if (a.act_f16) {                  // fp16 operand (11-bit significand), saturating
#pragma unroll
                            for (int e = 0; e < 4; ++e) hw[e] = umma::cvt_f16x2(o[2 * e], o[2 * e + 1]);
                        } else {                                 // plain bf16 operand: one conversion per channel pair
#pragma unroll
                            for (int e = 0; e < 4; ++e) hw[e] = umma::cvt_bf16x2(o[2 * e], o[2 * e + 1]);
                        }
                        *reinterpret_cast<uint2*>(a.act_hi + wofs) = make_uint2(hw[0], hw[1]);
                        *reinterpret_cast<uint2*>(a.act_hi + wofs + 8) = make_uint2(hw[2], hw[3]);
                    } else {
                        float* dst = img + (long long)(pofs + 8 * h) * Cf;
                        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(dst + 16) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                    if (t == 0 && simg) simg[(long long)(pofs + 8 * h) * 2] = word;
                }
            }
        }
    }
}

// Backward.  M = channels, N = taps (+ the ones column that yields the bias gradient), K = pixels.  A tile has 32
// groups of 8 consecutive pixels; warp w accumulates channels [32*(w&1), +32) over the groups (w>>1) + 4i.  Row
// r = g + 8h of M-tile m is channel 32*half + 4g + 2m + h: a lane needs channels 4g..4g+3 of the pixels t and t+4 of the
// group = one 128-bit load each, and the 8 lanes of a pixel read 128 contiguous bytes.  The loads of four groups
// (8 x LDG.128 + 8 sign words per lane) are issued before the first group is consumed.
template <int KS, int NP>
__global__ void __launch_bounds__(RCF_BLOCK, (KS <= 3) ? 2 : 1) k_stem_bwd_mma(const StemK a) {
    rcf_pdl_prologue();
    using TL = StemTile<KS>;
    constexpr int NT = 2 * KS * KS, NO = NT + 1, NJ = (NO + 7) / 8, NOP = NJ * 8, SW = TL::SW, TSZ = TL::TSZ;
    constexpr int Cf = STEM_MMA_CF;
    constexpr int GB = 4;                              // groups per load batch
    __shared__ uint32_t Thi[TSZ + 2], Tlo[TSZ + 2];
    extern __shared__ float red[];                   // [RCF_WARPS][32 * NOP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int half = warp & 1, gset = warp >> 1;
    const int tx = (a.W + STEM_TW - 1) / STEM_TW, ty = (a.H + STEM_TH - 1) / STEM_TH;

    float acc[2][NJ][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.0f;
    int offB[NJ], mulB[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) tap_addr<KS>(8 * j + g, offB[j], mulB[j]);
    if (tid == 0) { Thi[TSZ] = __float_as_uint(1.0f); Tlo[TSZ] = 0u; Thi[TSZ + 1] = 0u; Tlo[TSZ + 1] = 0u; }

    TL st;
    const TileDecode dec(tx, ty);
    int tl = blockIdx.x, n = 0, y0 = 0, x0 = 0, nn = 0, ny0 = 0, nx0 = 0;
    if (tl < a.ntiles) { dec(tl, nn, ny0, nx0); st.fetch(a, nn, ny0, nx0); }
    const long long rstride = (long long)a.W * Cf;     // one image row of dact, in floats
    for (; tl < a.ntiles; tl += gridDim.x) {      // fixed tile -> CTA assignment (reproducible)
        n = nn; y0 = ny0; x0 = nx0;
        __syncthreads();
        st.template commit<NP>(a, Thi, Tlo);
        __syncthreads();
        if (tl + (int)gridDim.x < a.ntiles) {
            dec(tl + gridDim.x, nn, ny0, nx0);
            st.fetch(a, nn, ny0, nx0);
        }
        // this warp always works on the column block lx0 = 8 * gset of the tile and walks its 8 rows (group q = gset + 4 * ly)
        const int lx0 = gset * 8, xA = x0 + lx0 + t;
        const bool inx[2] = {xA < a.W, xA + 4 < a.W};
        const float* const dcol = a.dact + ((long long)n * a.P + (long long)y0 * a.W + xA) * Cf + half * 32 + 4 * g;
        const uint32_t* const scol = a.sign_in + ((long long)n * a.P + (long long)y0 * a.W + xA) * 2 + half;
#pragma unroll 1
        for (int i0 = 0; i0 < 8; i0 += GB) {
            float4 d4[GB][2];
            uint32_t sg[GB][2];
#pragma unroll
            for (int u = 0; u < GB; ++u) {
                const bool iny = y0 + i0 + u < a.H;
                const float* dp = dcol + (long long)(i0 + u) * rstride;
                const uint32_t* sp = scol + (long long)(i0 + u) * a.W * 2;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool in = iny && inx[e];
                    d4[u][e] = in ? __ldg(reinterpret_cast<const float4*>(dp + 4 * e * Cf)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sg[u][e] = in ? __ldg(sp + 8 * e) : 0u;
                }
            }
#pragma unroll
            for (int u = 0; u < GB; ++u) {
                const int ly = i0 + u;
                if (y0 + ly >= a.H || x0 + lx0 >= a.W) continue;          // warp-uniform
                float v[2][4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t sb = sg[u][e] >> (4 * g);
                    v[e][0] = (sb & 1u) ? d4[u][e].x * a.slope : d4[u][e].x;
                    v[e][1] = (sb & 2u) ? d4[u][e].y * a.slope : d4[u][e].y;
                    v[e][2] = (sb & 4u) ? d4[u][e].z * a.slope : d4[u][e].z;
                    v[e][3] = (sb & 8u) ? d4[u][e].w * a.slope : d4[u][e].w;
                }
                const int pA = ly * SW + lx0 + t;
                uint32_t bh[NJ][2], bl[NJ][2];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int j0 = offB[j] + pA * mulB[j], j1 = offB[j] + (pA + 4) * mulB[j];
                    bh[j][0] = Thi[j0]; bh[j][1] = Thi[j1];
                    if constexpr (NP != 1) { bl[j][0] = Tlo[j0]; bl[j][1] = Tlo[j1]; }
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if constexpr (NP == 1) {
                        const uint32_t ah[4] = {to_tf32(v[0][2 * m]), to_tf32(v[0][2 * m + 1]), to_tf32(v[1][2 * m]), to_tf32(v[1][2 * m + 1])};
#pragma unroll
                        for (int j = 0; j < NJ; ++j) mma_tf32(acc[m][j], ah, bh[j][0], bh[j][1]);
                    } else {
                        uint32_t ah[4], al[4];
                        split_tf32_fast(v[0][2 * m], ah[0], al[0]);
                        split_tf32_fast(v[0][2 * m + 1], ah[1], al[1]);
                        split_tf32_fast(v[1][2 * m], ah[2], al[2]);
                        split_tf32_fast(v[1][2 * m + 1], ah[3], al[3]);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) mma_3xtf32(acc[m][j], ah, al, bh[j], bl[j]);
                    }
                }
            }
        }
    }
    // combine the warps of each channel half in warp order through shared memory; part[cta][f * NO + tap]
    float* mine = red + (size_t)warp * 32 * NOP;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int fl = 4 * g + 2 * m + h;                 // channel within the half
                mine[fl * NOP + 8 * j + 2 * t] = acc[m][j][2 * h];
                mine[fl * NOP + 8 * j + 2 * t + 1] = acc[m][j][2 * h + 1];
            }
    __syncthreads();
    for (int o = tid; o < Cf * NO; o += RCF_BLOCK) {
        const int f = o / NO, tp = o - f * NO;
        const int hf = f >> 5, fl = f & 31;
        float s = 0.0f;
#pragma unroll
        for (int wi = 0; wi < RCF_WARPS / 2; ++wi) s += red[(size_t)(2 * wi + hf) * 32 * NOP + fl * NOP + tp];
        a.part[(size_t)blockIdx.x * Cf * NO + o] = s;
    }
}

__global__ void __launch_bounds__(256) k_stem_bwd_final(const float* __restrict__ part, int nparts, int Cf, int NT,
                                                        float* __restrict__ dw, float* __restrict__ db,
                                                        const float* __restrict__ gmax, int nfd) {
    rcf_pdl_prologue();
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // warp per output f*(NT+1)+t
    if (o >= Cf * (NT + 1)) return;
    // dact arrived multiplied by the fp16 gradient scale (a power of two): divide it out of the sums (exact)
    const double inv = gmax ? 1.0 / (double)rcf_grad_scale(gmax, nfd) : 1.0;
    const double v = warp_sum_strided(part + o, nparts, (long long)Cf * (NT + 1), lane) * inv;
    if (lane == 0) {
        const int f = o / (NT + 1), t = o - f * (NT + 1);
        if (t < NT) dw[(size_t)f * NT + t] = (float)v;
        else db[f] = (float)v;
    }
}

int stem_check(int ndir, int B, int H, int W, int Cf, int ks) {
    if (ndir < 1 || ndir > 2 || B < 1 || H < 1 || W < 1) return RCF_ERR_SHAPE;
    if ((H + STEM_TH - 1) / STEM_TH > 65535) return RCF_ERR_SHAPE;
    if ((long long)ndir * B > 65535 || (long long)H * W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if (ks != 1 && ks != 3 && ks != 5) return RCF_ERR_UNSUPPORTED;
    if (Cf < 4 || Cf % 4 || Cf > 128 || 256 % (Cf / 4)) return RCF_ERR_UNSUPPORTED;
    return RCF_OK;
}

int stem_grid_bwd(int ntiles) { return ntiles < 2 * 148 ? ntiles : 2 * 148; }

void fill(StemK& a, const float* const* flow, const int64_t* bs, int ndir, int B, int H, int W, int Cf, float clamp_t,
          float slope) {
    a.flow[0] = flow[0]; a.flow[1] = ndir > 1 ? flow[1] : flow[0];
    a.flow_bs[0] = bs[0]; a.flow_bs[1] = ndir > 1 ? bs[1] : bs[0];
    a.ndir = ndir; a.B = B; a.H = H; a.W = W; a.P = H * W; a.Cf = Cf;
    a.clamp_t = clamp_t; a.slope = slope;
    a.ntiles = ndir * B * ((W + STEM_TW - 1) / STEM_TW) * ((H + STEM_TH - 1) / STEM_TH);
}

}  // namespace

#define RCF_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" int rcf_stem_workspace_bytes(int ndir, int B, int H, int W, int Cf, int ks, size_t* bytes) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!bytes) return RCF_ERR_NULL;
    const int ntiles = ndir * B * ((W + STEM_TW - 1) / STEM_TW) * ((H + STEM_TH - 1) / STEM_TH);
    *bytes = (size_t)stem_grid_bwd(ntiles) * Cf * (2 * ks * ks + 1) * sizeof(float);
    return RCF_OK;
}

extern "C" int rcf_stem_forward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                int Cf, int ks, const float* w, const float* b, float clamp_t, float slope, float* act,
                                uint32_t* sign, void* stream) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!flow || !flow_bstride || !flow[0] || (ndir > 1 && !flow[1]) || !w || !b || !act) return RCF_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(act) & 15u) return RCF_ERR_ALIGN;
    if (sign && (Cf != STEM_MMA_CF || !(slope >= 0.0f && slope <= 1.0f))) return RCF_ERR_UNSUPPORTED;
    StemK a{};
    fill(a, flow, flow_bstride, ndir, B, H, W, Cf, clamp_t, slope);
    a.w = w; a.b = b; a.act = act; a.sign_out = sign;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (Cf == STEM_MMA_CF && slope >= 0.0f && slope <= 1.0f) {            // tensor-core path, persistent CTAs (weights split once per CTA)
        const int g = stem_grid_bwd(a.ntiles);
        switch (ks) {
            case 1: k_stem_fwd_mma<1, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 3: k_stem_fwd_mma<3, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 5: k_stem_fwd_mma<5, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
        }
        RCF_CUDA(cudaGetLastError());
        return RCF_OK;
    }
    dim3 grid((W + STEM_TW - 1) / STEM_TW, (H + STEM_TH - 1) / STEM_TH, ndir * B);
    switch (ks) {
        case 1: k_stem_fwd<1><<<grid, RCF_BLOCK, 0, s>>>(a); break;
        case 3: k_stem_fwd<3><<<grid, RCF_BLOCK, 0, s>>>(a); break;
        case 5: k_stem_fwd<5><<<grid, RCF_BLOCK, 0, s>>>(a); break;
    }
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}

// Same layer, output as the bf16 (hi, lo) pair the tcgen05 conv consumes (Cf = 64 only; act_lo may be NULL).
extern "C" int rcf_stem_forward_bf16(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                     int ks, const float* w, const float* b, float clamp_t, float slope, void* act_hi,
                                     void* act_lo, uint32_t* sign, int nprod, void* stream) {
    const int out_f16 = (nprod & RCF_STEM_OUT_F16) ? 1 : 0;
    nprod &= 0xff;
    if (out_f16 && act_lo) return RCF_ERR_MODE;           // the fp16 word stands alone (no lo word)
    const int v = stem_check(ndir, B, H, W, STEM_MMA_CF, ks);
    if (v != RCF_OK) return v;
    if (!flow || !flow_bstride || !flow[0] || (ndir > 1 && !flow[1]) || !w || !b || !act_hi) return RCF_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(act_hi) | reinterpret_cast<uintptr_t>(act_lo)) & 15u) return RCF_ERR_ALIGN;
    if (!(slope >= 0.0f && slope <= 1.0f)) return RCF_ERR_UNSUPPORTED;
    StemK a{};
    fill(a, flow, flow_bstride, ndir, B, H, W, STEM_MMA_CF, clamp_t, slope);
    a.w = w; a.b = b; a.act = nullptr; a.sign_out = sign;
    a.act_hi = static_cast<uint32_t*>(act_hi); a.act_lo = static_cast<uint32_t*>(act_lo);
    a.act_f16 = out_f16;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int g = stem_grid_bwd(a.ntiles);
    if (nprod >= 3) {
        switch (ks) {
            case 1: k_stem_fwd_mma<1, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 3: k_stem_fwd_mma<3, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 5: k_stem_fwd_mma<5, 3><<<g, RCF_BLOCK, 0, s>>>(a); break;
        }
    } else {
        switch (ks) {
            case 1: k_stem_fwd_mma<1, 1><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 3: k_stem_fwd_mma<3, 1><<<g, RCF_BLOCK, 0, s>>>(a); break;
            case 5: k_stem_fwd_mma<5, 1><<<g, RCF_BLOCK, 0, s>>>(a); break;
        }
    }
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}

template <int KS, int NP>
static int launch_stem_bwd_mma(const StemK& a, int g, cudaStream_t s) {
    constexpr int NJ = (2 * KS * KS + 1 + 7) / 8;
    const size_t smem = (size_t)RCF_WARPS * 32 * NJ * 8 * sizeof(float);
    RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd_mma<KS, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return (int)rcf_launch(k_stem_bwd_mma<KS, NP>, g, RCF_BLOCK, smem, s, rcf_pdl_enabled(), a);
}

extern "C" int rcf_stem_backward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                 int Cf, int ks, float clamp_t, float slope, const float* act, const uint32_t* sign,
                                 const float* dact, float* dw, float* db, void* ws, int nprod, void* stream) {
    return rcf_stem_backward_ex(flow, flow_bstride, ndir, B, H, W, Cf, ks, clamp_t, slope, act, sign, dact, dw, db, ws, nprod,
                                nullptr, 0, stream);
}

int rcf_stem_backward_ex(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W, int Cf, int ks,
                         float clamp_t, float slope, const float* act, const uint32_t* sign, const float* dact, float* dw,
                         float* db, void* ws, int nprod, const float* gmax, int nfd, void* stream) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!flow || !flow_bstride || !flow[0] || (ndir > 1 && !flow[1]) || (!act && !sign) || !dact || !dw || !db || !ws)
        return RCF_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(dact)) & 15u) return RCF_ERR_ALIGN;
    if (sign && Cf != STEM_MMA_CF) return RCF_ERR_UNSUPPORTED;
    StemK a{};
    fill(a, flow, flow_bstride, ndir, B, H, W, Cf, clamp_t, slope);
    a.act_in = act; a.sign_in = sign; a.dact = dact; a.part = static_cast<float*>(ws);
    const int g = stem_grid_bwd(a.ntiles);
    const int NT = 2 * ks * ks;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (sign) {                         // tensor-core path (sign bits written by rcf_stem_forward)
        int e = RCF_OK;
        if (nprod >= 3) {
            switch (ks) {
                case 1: e = launch_stem_bwd_mma<1, 3>(a, g, s); break;
                case 3: e = launch_stem_bwd_mma<3, 3>(a, g, s); break;
                case 5: e = launch_stem_bwd_mma<5, 3>(a, g, s); break;
            }
        } else {
            switch (ks) {
                case 1: e = launch_stem_bwd_mma<1, 1>(a, g, s); break;
                case 3: e = launch_stem_bwd_mma<3, 1>(a, g, s); break;
                case 5: e = launch_stem_bwd_mma<5, 1>(a, g, s); break;
            }
        }
        if (e != RCF_OK) return e;
    } else {
        const size_t smem = (size_t)RCF_WARPS * Cf * (NT + 1) * sizeof(float);
        if (smem > 220 * 1024) return RCF_ERR_UNSUPPORTED;
        const bool big = smem > 40 * 1024;       // (plus the static tile) needs the opt-in shared-memory limit
        switch (ks) {
            case 1:
                if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_stem_bwd<1><<<g, RCF_BLOCK, smem, s>>>(a);
                break;
            case 3:
                if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_stem_bwd<3><<<g, RCF_BLOCK, smem, s>>>(a);
                break;
            case 5:
                if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_stem_bwd<5><<<g, RCF_BLOCK, smem, s>>>(a);
                break;
        }
        RCF_CUDA(cudaGetLastError());
    }
    const int nout = Cf * (NT + 1);
    RCF_CUDA(rcf_launch(k_stem_bwd_final, (nout * 32 + 255) / 256, 256, 0, s, rcf_pdl_enabled(), (const float*)a.part, g, Cf, NT, dw, db,
                        gmax, nfd));
    return RCF_OK;
}
