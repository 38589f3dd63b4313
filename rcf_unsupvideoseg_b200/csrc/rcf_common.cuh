// rcf_common.cuh -- shared definitions of the sm_100a RCF motion-loss kernels.
//
// Math reference: SURVEY.md 8(a)-math; reference Python lines cited per kernel.
// Notation: fd = frame-direction index = dir * B + b; P = H*W pixels; K segments; D coordinate
// features (0 / 2 / 5); Cf pooled feature channels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rcf_loss.h"

#define RCF_BLOCK 256
#define RCF_WARPS (RCF_BLOCK / 32)

// pixels handled by one CTA of each streaming kernel
#define RCF_CHUNK_MOM 4096   // pass 1 (moments); the affine fit (D == 2) takes 8192: see rcf_chunk_mom
#define RCF_CHUNK_LOSS 2048  // pass 2 (reconstruct + loss + gradient moments)
#define RCF_CHUNK_BWD 1024   // backward (no reduction)

#define RCF_HD __host__ __device__ __forceinline__

// ---- sizes of the small per-segment records ------------------------------------------------
// pass-1 statistics per (fd,k): [0] S, [1..2] sum m*F_c, [3..3+D) sum m*u_d,
//   [3+D .. 3+3D) sum m*F_c*u_d (c*D+d), [3+3D ..) sum m*u_d*u_e packed d<=e
// pixels per CTA of pass 1: with the affine fit a CTA ends with a 48-accumulator reduction (~150 instructions per thread);
// twice the pixels per CTA halves that overhead: C2 affine 0.459 -> 0.443 ms, FBMS K=3 0.388 -> 0.373 ms; K = 8 (64-bit
// packs, two segment groups) got slower with it (1.68 -> 1.81 ms) and keeps 4096
RCF_HD constexpr int rcf_chunk_mom(int D, int K) { return (D == 2 && K <= 4) ? 2 * RCF_CHUNK_MOM : RCF_CHUNK_MOM; }
// Pass 2 with the affine fit carries 1 + 2K + 4K accumulators through a warp/CTA reduction per chunk: large frames take
// 4x longer chunks (K <= 4: the wider kernels have no registers for the loop state) so that epilogue (and the coefficient prologue) amortise; the training shapes keep the short one.
RCF_HD int rcf_chunk_loss(int D, int K, int P, int nfd) {
    return (D == 2 && K <= 4 && (long long)((P + 8191) / 8192) * nfd >= 4 * 148) ? 8192 : RCF_CHUNK_LOSS;
}
RCF_HD constexpr int rcf_ns(int D) { return D == 0 ? 1 : 3 + 3 * D + D * (D + 1) / 2; }
// pass-2 gradient moments per fd: [0] sum phi, [1 + c*K + k] sum w_c m_k,
//   D > 0: [1 + 2K + (k*2+c)*D + d] sum w_c m_k (u_d - mu_kd)
//   D = 0: [1 + 2K + k] sum m_k  (the mask sums S_k ride along for free: with theta supplied and no affine fit
//          nothing in pass 2 depends on S_k, so the forward is a single pass and S_k is only needed by the backward)
RCF_HD constexpr int rcf_gm(int K, int D) { return D == 0 ? 1 + 3 * K : 1 + 2 * K + 2 * K * D; }
// forward coefficients per (fd,k) (fp32): theta[2], A[2][D], mu_u[D]
RCF_HD constexpr int rcf_cf(int D) { return 2 + 3 * D; }
// backward coefficients per (fd,k) (fp32), all pre-divided by S_k:
//   muF[2], B[2][D], Csym[D(D+1)/2] (off-diagonals doubled), mubar_u[D], c0
RCF_HD constexpr int rcf_cb(int D) { return 3 + 3 * D + D * (D + 1) / 2; }
// fp64 per-segment state kept for backward: S, mu_u[D], mu_F[2], SFu[2D], Suu[D*D], Sinv[D*D], A[2D]
RCF_HD constexpr int rcf_segd(int D) { return 3 + 5 * D + 2 * D * D; }
RCF_HD constexpr int rcf_sym_idx(int D, int d, int e) { return d * D - d * (d - 1) / 2 + (e - d); }

// Pixels per CTA of the channels-last pooling kernels.  A CTA's time is a chain of dependent global round trips (mask
// tile, batches of feature loads, partial store), so large frames take long chunks (the prologue/epilogue amortise over
// more batches: 1024 px forward / 512 px backward) and the 96x96 / 48x48 training shapes take short ones in the forward
// (halved while the grid would not give every SM ~4 CTAs: 12.1 -> 6.0 us at 8x2x48x48).  The backward does not benefit
// from going below 256 (22.4 -> 24.5 us at 8x2x96x96 with 64-pixel CTAs: its prologue -- coefficient pack and two tiles --
// outweighs the shorter pixel loop).
RCF_HD int rcf_pool_chunk_nhwc(int P, int nfd) {
    int c = 1024;
    while (c > 128 && (long long)((P + c - 1) / c) * nfd < 4 * 148) c >>= 1;
    return c;
}
RCF_HD int rcf_pool_tp_nhwc(int P, int nfd) { return ((long long)((P + 511) / 512) * nfd >= 4 * 148) ? 512 : 256; }
RCF_HD int rcf_pool_chunk(int K, int nhwc, int P, int nfd) { return nhwc ? rcf_pool_chunk_nhwc(P, nfd) : (K <= 4 ? 2048 : 1024); }

// ---- memory plan ------------------------------------------------------------------------------
struct RcfLayout {
    int nfd, P, ns, gm, cf, cb, segd;
    int nchunk1, nchunk2, nchunkb, nchunkp;
    int chunk2;   // pixels per CTA of pass 2
    int pool_sums;   // S_k partials come from k_pool_nhwc (nchunk1 == nchunkp), no k_moments launch
    // ctx (bytes offsets)
    size_t c_segd, c_coef, c_mlp, c_gm, c_bytes;
    // ws
    size_t w_part1, w_partp, w_part2, w_coefb, w_gscale, w_poolbar, w_dh, w_thbar, w_dbpart, w_dbfd, w_poolsum, w_cnt, w_gmax, w_bytes;
    int nblkpb;   // CTAs per frame-direction of the channels-last pooling backward (pooltp pixels each)
    int poolchunk, pooltp;
};

static inline size_t rcf_align256(size_t x) { return (x + 255) & ~(size_t)255; }

static inline RcfLayout rcf_make_layout(const RcfDesc& d) {
    RcfLayout L;
    L.nfd = d.ndir * d.B;
    L.P = d.H * d.W;
    L.ns = rcf_ns(d.D);
    L.gm = rcf_gm(d.K, d.D);
    L.cf = rcf_cf(d.D);
    L.cb = rcf_cb(d.D);
    L.segd = rcf_segd(d.D);
    L.nchunk1 = (L.P + rcf_chunk_mom(d.D, d.K) - 1) / rcf_chunk_mom(d.D, d.K);
    // Drop-in head without the affine fit: pass 1 would only produce S_k = sum of the masks, and the channels-last pooling
    // kernel has every mask tile in shared memory anyway -- it writes the per-chunk mask sums itself (its own chunking),
    // k_moments is not launched.
    L.pool_sums = (d.theta_mode == 1 && d.D == 0 && d.feat_nhwc && d.Cf > 0) ? 1 : 0;
    L.chunk2 = rcf_chunk_loss(d.D, d.K, L.P, L.nfd);
    L.nchunk2 = (L.P + L.chunk2 - 1) / L.chunk2;
    L.nchunkb = (L.P + RCF_CHUNK_BWD - 1) / RCF_CHUNK_BWD;
    const int pc = rcf_pool_chunk(d.K, d.feat_nhwc, L.P, L.nfd);
    L.nchunkp = (L.P + pc - 1) / pc;
    if (L.pool_sums) L.nchunk1 = L.nchunkp;
    const size_t nseg = (size_t)L.nfd * d.K;
    size_t o = 0;
    L.c_segd = o; o = rcf_align256(o + nseg * L.segd * sizeof(double));
    L.c_coef = o; o = rcf_align256(o + nseg * L.cf * sizeof(float));
    L.c_mlp = o;  o = rcf_align256(o + nseg * 2 * (size_t)d.Cf * sizeof(double));
    L.c_gm = o;   o = rcf_align256(o + (size_t)L.nfd * L.gm * sizeof(double));
    L.c_bytes = o ? o : 256;
    o = 0;
    L.w_part1 = o;   o = rcf_align256(o + nseg * L.ns * L.nchunk1 * sizeof(float));
    L.w_partp = o;   o = rcf_align256(o + nseg * d.Cf * L.nchunkp * sizeof(float));
    L.w_part2 = o;   o = rcf_align256(o + (size_t)L.nfd * L.gm * L.nchunk2 * sizeof(float));
    L.w_coefb = o;   o = rcf_align256(o + nseg * L.cb * sizeof(float));
    L.w_gscale = o;  o = rcf_align256(o + (size_t)L.nfd * sizeof(float));
    L.w_poolbar = o; o = rcf_align256(o + nseg * d.Cf * sizeof(float));
    L.w_dh = o;      o = rcf_align256(o + nseg * d.Cf * sizeof(double));
    L.w_thbar = o;   o = rcf_align256(o + nseg * 2 * sizeof(double));
    L.pooltp = rcf_pool_tp_nhwc(L.P, L.nfd);
    L.poolchunk = pc;
    L.nblkpb = (L.P + L.pooltp - 1) / L.pooltp;
    L.w_dbpart = o;  o = rcf_align256(o + (d.feat_nhwc ? (size_t)L.nfd * L.nblkpb * d.Cf * sizeof(float) : 0));   // per-CTA bias-gradient partials
    L.w_dbfd = o;    o = rcf_align256(o + (d.feat_nhwc ? (size_t)L.nfd * 16 * d.Cf * sizeof(double) : 0));   // <= 16 splits per fd
    L.w_poolsum = o; o = rcf_align256(o + nseg * d.Cf * sizeof(double));
    L.w_cnt = o;     o = rcf_align256(o + 64);      // arrival counters of the "last CTA finishes the reduction" kernels
    L.w_gmax = o;    o = rcf_align256(o + (size_t)L.nfd * sizeof(float));   // per frame-direction max |dPool/S| (fp16 gradient scale)
    L.w_bytes = o ? o : 256;
    return L;
}

// ---- kernel argument block (passed by value) ---------------------------------------------------
struct RcfK {
    int B, K, H, W, P, Cf, D, ndir, nfd;
    int robust, unbounded, theta_mode;
    float eps, q;
    float scale;        // residual_adjustment_scale (1 when unbounded)
    float ex2_scale;    // 2*log2(e)/pred_div : tanh(r/div) = 1 - 2/(1+2^(r*ex2_scale))
    float dres_scale;   // scale/pred_div (1 when unbounded)
    float clamp_t;      // <0: none
    float inv_n;
    float cy, cx, sy, sx;  // coordinate centring / scaling (u = ((row-cy)*sy, (col-cx)*sx))
    float feat_slope;      // LeakyReLU slope applied to feat on load (1 = none)
    int feat_nhwc;         // feat / dfeat are channels-last
    const float* mask[2];
    const float* flow[2];
    const float* resid[2];
    const float* feat[2];
    const float* theta[2];
    long long mask_bs[2], flow_bs[2], resid_bs[2], feat_bs[2];
    const float *w1, *b1, *w2, *b2;
    const float* feat_bias;   // [Cf] or null (channels-last only)
    // ctx
    double* segd;
    float* coef;
    double* mlp;
    double* gm;
    // ws
    float* part1;
    float* partp;
    float* part2;
    float* coefb;
    float* gscale;
    float* poolbar;
    double* dh;
    double* thbar;
    int nchunk1, nchunk2, nchunkb, nchunkp;
    int chunk2;    // pixels per CTA of pass 2 (a multiple of RCF_CHUNK_LOSS)
    int pool_sums; // k_pool_nhwc also writes the mask-sum partials of pass 1 (see rcf_make_layout)
    int l2_hints;  // pass 2 streams flow/residual with an L2 evict-first policy (keeps the masks resident)
    int pdl;          // launch with programmatic stream serialization (kernels call rcf_pdl_prologue() first)
    int mlp_smem;     // segment kernels stage the MLP weights in shared memory (Cf % 4 == 0 and Cf <= 128)
    int single_pass;  // theta supplied and D == 0: no pass 1; S_k comes out of pass 2 (k_finalize stores it)
    // forward outputs
    float* loss;
    float* vis_gt; float* vis_pred; float* vis_agg; float* vis_res; float* vis_aff;
    long long vis_bs, vis_ds;
    float vis_scale[2];
    // backward
    const float* grad_loss;
    int grad_total;           // grad_loss is one float (gradient of the total), not one per direction
    float* dmask[2]; float* dresid[2]; float* dfeat[2]; float* dtheta[2];
    uint32_t* dfeat_hi[2]; uint32_t* dfeat_lo[2];   // channels-last dfeat as bf16 pairs (hi word, lo word) instead of fp32
    long long dmask_bs[2], dresid_bs[2], dfeat_bs[2];
    float *dw1, *db1, *dw2, *db2;
    float* dfeat_bias;        // [Cf] or null
    float* dbpart;            // ws: [nfd][nblkpb][Cf]
    double* dbfd;             // ws: [nfd * S][Cf], S <= 16 splits of the partial rows
    double* poolsum;          // ws: [nfd][Cf*K] pooled sums (un-normalised), reduced over chunks by k_pool_reduce
    float* gmax;              // ws: [nfd] max |poolbar| per frame-direction, written by k_segment_bwd; source of rcf_grad_scale
    int dfeat_f16;            // dfeat_hi is written as fp16 words scaled by rcf_grad_scale(gmax) (RcfDesc.dfeat_f16)
    unsigned int* cnt;        // ws: arrival counters -- [1] k_bias_grad_fd (zeroed by k_pool_bwd_nhwc); [0] unused (folding k_loss_sum into
                              // k_finalize the same way was measured: +0.7 us on the C2 loss-core step, so that pair stays two launches)
    int nblkpb;
    int poolchunk, pooltp;    // pixels per CTA of k_pool_nhwc / k_pool_bwd_nhwc
};

#ifdef __CUDACC__
// ---- programmatic dependent launch ----------------------------------------------------------------------------------
// The library's kernels form dependent chains of 6-7 launches per call, most of them microseconds long; inside a CUDA
// graph the hand-over between two of them costs ~1.5-2 us.  Launched with the programmatic-stream-serialization attribute
// a kernel may be scheduled while its predecessor drains; every kernel therefore starts with rcf_pdl_prologue():
// `launch_dependents` (lets the successor's CTAs take the SM slots this grid no longer needs; it fires once every CTA
// of this grid has executed it or exited) followed by `wait` (blocks until the predecessor has completed and its writes are
// visible).  Nothing is read or written before the wait, so results are unchanged; without the attribute both are no-ops.
__device__ __forceinline__ void rcf_pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Power-of-two scale that brings the feature-map gradient into fp16's normal range: dG[p][f] = dact * sum_k c[f][k] M[k][p]
// with sum_k M = 1 and |dact| <= 1, so |dG| <= max |c| =: m; s = 2^(12 - exponent(m)) puts the largest element in
// [2^11, 2^12) (fp16 overflows at 2^16: a factor 16 of headroom for masks that do not sum to one; the conversion
// saturates beyond it) and keeps 11 significant bits for everything down to 2^-26 of it.  Every consumer
// recomputes s from the same nfd floats, so all of them agree bit for bit; scaling by s and by 1/s is exact.
// Call with all 32 lanes of a warp.
__device__ __forceinline__ float rcf_grad_scale(const float* __restrict__ gmax, int nfd) {
    float m = 0.0f;
    for (int i = (int)(threadIdx.x & 31); i < nfd; i += 32) m = fmaxf(m, __ldcg(gmax + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (!(m > 0.0f) || !(m < 3.0e38f)) return 1.0f;          // zero / inf / nan gradient: no scaling
    int e = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126;   // m = f * 2^e, f in [0.5, 1)  (subnormal m: e = -126, fine)
    e = 12 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return __uint_as_float((uint32_t)(e + 127) << 23);
}

// "The last CTA to arrive finishes the job": every CTA publishes its partial results, then one thread takes a ticket; the
// CTA that draws the last ticket sees all partials (fence + atomic) and runs the final, fixed-order reduction itself, so
// the separate single-CTA launch that used to follow disappears.  The counter is zeroed by the PRECEDING kernel of the
// same stream (a plain store by one thread), so graph replays and back-to-back calls need no memset.  Returns true in
// every thread of the last CTA.  Partials written by other CTAs must then be read with __ldcg (L1 is not coherent).
__device__ __forceinline__ bool rcf_last_cta(unsigned int* counter, unsigned int total) {
    __shared__ int s_last;
    __syncthreads();                                   // this CTA's partials are written
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == total - 1u) ? 1 : 0;
        __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t rcf_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int pdl,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------------
template <int PX> struct Pack;
template <> struct Pack<4> {
    static __device__ __forceinline__ void ld(float (&v)[4], const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void ld_rw(float (&v)[4], const float* p) {   // coherent (buffer is also written)
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
// streaming loads that should not displace L2-resident data (evict-first policy, no L1 allocation)
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void ld_evict_first(float (&v)[4], const float* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ld_evict_first(float (&v)[2], const float* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v[0]), "=f"(v[1]) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ld_evict_first(float (&v)[1], const float* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[0]) : "l"(p), "l"(pol));
}

template <> struct Pack<2> {   // 64-bit accesses: used for K > 4 where four pixels per thread would not fit in registers
    static __device__ __forceinline__ void ld(float (&v)[2], const float* p) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void ld_rw(float (&v)[2], const float* p) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[2]) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct Pack<1> {
    static __device__ __forceinline__ void ld(float (&v)[1], const float* p) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void ld_rw(float (&v)[1], const float* p) { v[0] = *p; }
    static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic warp reduction of a vector of N accumulators in ~N shuffles instead of 5N:
// recursive halving (each lane keeps half of its values and sends the other half to its partner at
// lane distance OFF), then plain butterflies once a lane is down to one value.
template <int C, int OFF>
__device__ __forceinline__ float warp_reduce_pow2(const float (&v)[C], int lane) {
    if constexpr (C == 1) {
        float x = v[0];
#pragma unroll
        for (int o = OFF; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    } else {
        float h[C / 2];
        const bool upper = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < C / 2; ++i) {
            const float send = upper ? v[i] : v[i + C / 2];
            const float keep = upper ? v[i + C / 2] : v[i];
            h[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        return warp_reduce_pow2<C / 2, OFF / 2>(h, lane);
    }
}
// Reduce acc[0..N) over the warp and store total i to dst[i] (one lane per value writes).
template <int N, int BASE = 0>
__device__ __forceinline__ void warp_reduce_store(const float (&acc)[N], int lane, float* dst) {
    if constexpr (BASE < N) {
        constexpr int REM = N - BASE;
        constexpr int C = REM > 16 ? 32 : REM > 8 ? 16 : REM > 4 ? 8 : REM > 2 ? 4 : REM > 1 ? 2 : 1;
        constexpr int SH = C == 32 ? 0 : C == 16 ? 1 : C == 8 ? 2 : C == 4 ? 3 : C == 2 ? 4 : 5;
        float v[C];
#pragma unroll
        for (int i = 0; i < C; ++i) v[i] = (BASE + i < N) ? acc[BASE + i] : 0.0f;
        const float tot = warp_reduce_pow2<C, 16>(v, lane);
        const int idx = lane >> SH;
        if ((lane & ((1 << SH) - 1)) == 0 && BASE + idx < N) dst[BASE + idx] = tot;
        warp_reduce_store<N, BASE + 32>(acc, lane, dst);
    }
}

// One warp sums n floats at base[i*stride] in fp64: lanes stride over i with four independent chains (loads in flight),
// then a butterfly.  The order depends on n only => bit-reproducible.  Loads bypass L1 (written by other SMs / kernels).
__device__ __forceinline__ double warp_sum_strided(const float* __restrict__ base, int n, long long stride, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = lane;
    for (; i + 96 < n; i += 128) {
        a0 += (double)__ldcg(base + (long long)i * stride);
        a1 += (double)__ldcg(base + (long long)(i + 32) * stride);
        a2 += (double)__ldcg(base + (long long)(i + 64) * stride);
        a3 += (double)__ldcg(base + (long long)(i + 96) * stride);
    }
    for (; i < n; i += 32) a0 += (double)__ldcg(base + (long long)i * stride);
    return warp_sum_d((a0 + a1) + (a2 + a3));
}

// ---- packed fp32 pairs (sm_100 FFMA2): one issue slot does two independent round-to-nearest FMAs, bit-identical to
// two FFMAs; used where a kernel is issue-bound on fp32 multiply-accumulate (pass-1 moments).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh(r/div) with absolute error ~3e-7 (2 MUFU + 3 FP32 ops); saturates cleanly to +-1.
__device__ __forceinline__ float tanh_scaled(float r, float ex2_scale) {
    const float e = fast_ex2(r * ex2_scale);
    return fmaf(-2.0f, fast_rcp(1.0f + e), 1.0f);
}
__device__ __forceinline__ float clamp_flow(float f, float t) {
    return t >= 0.0f ? fminf(fmaxf(f, -t), t) : f;
}

// coordinates of PX consecutive flat pixels starting at (row, col): y = (row - cy) * sy, x = (col - cx) * sx, the row
// advancing where the pack crosses the end of a line.  col - cx is formed as (col - cx) + j [- W]: every term is a
// multiple of 0.5 far below 2^24, so the sum is exact and the result equals ((float)(col + j) - cx) * sx bit for bit,
// at 4-5 instructions per pixel instead of two int->float conversions plus the wrap bookkeeping.
template <int PX>
__device__ __forceinline__ void px_coords_rc(int row, int col, const RcfK& a, float (&y)[PX], float (&x)[PX]) {
    const float y0 = ((float)row - a.cy) * a.sy, y1 = ((float)(row + 1) - a.cy) * a.sy;
    const float base = (float)col - a.cx, wf = (float)a.W;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const bool wrap = col + j >= a.W;            // PX <= W on every path that uses packs
        float c = base + (float)j;
        if (wrap) c -= wf;
        y[j] = wrap ? y1 : y0;
        x[j] = c * a.sx;
    }
}
// ... starting at flat pixel p
template <int PX>
__device__ __forceinline__ void px_coords(int p, const RcfK& a, float (&y)[PX], float (&x)[PX]) {
    const int row = p / a.W;
    px_coords_rc<PX>(row, p - row * a.W, a, y, x);
}
template <int D>
__device__ __forceinline__ void px_feats(float y, float x, float (&u)[D > 0 ? D : 1]) {
    if (D >= 2) { u[0] = y; u[1] = x; }
    if (D == 5) { u[2] = y * y; u[3] = x * x; u[4] = y * x; }
}

// loss value and derivative weight of one residual d = F - pred (reference :359-368)
__device__ __forceinline__ void loss_terms(float d, const RcfK& a, float& phi, float& w) {
    const float ad = fabsf(d);
    const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
    if (a.robust) {
        const float t = ad + a.eps;
        phi = fast_ex2(a.q * fast_lg2(t));
        w = a.q * phi * fast_rcp(t) * sg;
    } else {
        phi = ad;
        w = sg;
    }
}
#endif  // __CUDACC__

int rcf_pdl_enabled();   // RCF_OPT_PDL (rcf_capi.cu)

// launchers (one translation unit each)
cudaError_t rcf_launch_moments(const RcfK& a, bool vec, cudaStream_t s);
cudaError_t rcf_launch_pool(const RcfK& a, bool vec, cudaStream_t s);
cudaError_t rcf_launch_segment_fwd(const RcfK& a, cudaStream_t s);
cudaError_t rcf_launch_loss(const RcfK& a, bool vec, cudaStream_t s);
cudaError_t rcf_launch_finalize(const RcfK& a, cudaStream_t s);
cudaError_t rcf_launch_segment_bwd(const RcfK& a, cudaStream_t s);
cudaError_t rcf_launch_pool_bwd(const RcfK& a, bool vec, cudaStream_t s);
cudaError_t rcf_launch_bwd(const RcfK& a, bool vec, cudaStream_t s);
