// rcf_pool.cu -- masked pooling of the conv feature map and its backward ("Scope H" in SURVEY.md 8(d)).
//
// Forward (reference :251-256): Pool[f,k] = sum_p G[f,p] * M[k,p] / S_k.  The reference materialises
// G[:, :, None] * Mn[:, None] = [B,Cf,K,H,W] (6.7 GB at B=16, K=4, 480x854) and sums it; here G is
// read exactly once and the mask tile of the CTA is staged in shared memory so every warp (one feature
// channel at a time) reuses it.  Cf x K is a skinny contraction (K <= 8): bandwidth-bound, CUDA cores.
//
// Backward: dG[f,p] = sum_k c[f,k] M[k,p]   and   dM[k,p] += sum_f c[f,k] G[f,p],  c = Poolbar / S.
// Thread-per-pixel-pack mapping: G is read once, dG written once, the dM term is accumulated in registers
// on top of what the streaming backward (k_bwd, launched before) already stored.
#include <type_traits>
#ifndef RCF_POOL_BWD_BATCH
#define RCF_POOL_BWD_BATCH 8
#endif
#include "rcf_common.cuh"
#include "rcf_umma.cuh"

// blockIdx.z splits the feature channels (more CTAs for small frames: 96x96 training shapes would otherwise
// launch only 80 CTAs on 148 SMs).  A warp handles FB feature channels at once so that one shared-memory read of
// the mask tile feeds FB global loads (LDS bandwidth, not HBM, was the limiter with FB = 1).
template <int K, int PX, int CHUNK>
__global__ void __launch_bounds__(RCF_BLOCK) k_pool(const RcfK a, int f_per_cta) {
    rcf_pdl_prologue();
    constexpr int FB = 4;
    __shared__ __align__(16) float ms[K][CHUNK];
    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, Cf = a.Cf;
    const int p0 = chunk * CHUNK;
    const int f_begin = blockIdx.z * f_per_cta;
    const int f_end = min(Cf, f_begin + f_per_cta);
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ feat = a.feat[dir] + (long long)b * a.feat_bs[dir];
    const float slope = a.feat_slope;     // fused LeakyReLU of the last conv (1 = feat already activated)

#pragma unroll
    for (int k = 0; k < K; ++k) {
        for (int i = tid * PX; i < CHUNK; i += RCF_BLOCK * PX) {
            float v[PX];
            if (p0 + i < P) Pack<PX>::ld(v, mask + (long long)k * P + p0 + i);
            else {
#pragma unroll
                for (int j = 0; j < PX; ++j) v[j] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) ms[k][i + j] = v[j];
        }
    }
    __syncthreads();

    for (int f0 = f_begin + warp * FB; f0 < f_end; f0 += RCF_WARPS * FB) {
        float acc[FB * K];
#pragma unroll
        for (int i = 0; i < FB * K; ++i) acc[i] = 0.0f;
#pragma unroll 2
        for (int i = lane * PX; i < CHUNK; i += 32 * PX) {
            if (p0 + i < P) {
                float gv[FB][PX];
#pragma unroll
                for (int q = 0; q < FB; ++q) {
                    if (f0 + q < f_end) {
                        Pack<PX>::ld(gv[q], feat + (long long)(f0 + q) * P + p0 + i);
#pragma unroll
                        for (int j = 0; j < PX; ++j) gv[q][j] = gv[q][j] >= 0.0f ? gv[q][j] : slope * gv[q][j];
                    } else {
#pragma unroll
                        for (int j = 0; j < PX; ++j) gv[q][j] = 0.0f;
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float mv[PX];
#pragma unroll
                    for (int j = 0; j < PX; ++j) mv[j] = ms[k][i + j];
#pragma unroll
                    for (int q = 0; q < FB; ++q)
#pragma unroll
                        for (int j = 0; j < PX; ++j) acc[q * K + k] = fmaf(gv[q][j], mv[j], acc[q * K + k]);
                }
            }
        }
        // FB*K totals; partp layout [fd][f*K + k][chunk]
        __shared__ float red[RCF_WARPS][FB * K];
        warp_reduce_store<FB * K>(acc, lane, red[warp]);
        __syncwarp();
        if (lane < FB * K) {
            const int q = lane / K, k = lane - q * K;
            if (f0 + q < f_end)
                a.partp[((size_t)fd * Cf * K + (size_t)(f0 + q) * K + k) * a.nchunkp + chunk] = red[warp][lane];
        }
        __syncwarp();
    }
}

template <int K, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_pool_bwd(const RcfK a) {
    rcf_pdl_prologue();
    extern __shared__ float cs[];   // [Cf*K]
    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int tid = threadIdx.x;
    const int P = a.P, Cf = a.Cf;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ feat = a.feat[dir] + (long long)b * a.feat_bs[dir];
    float* __restrict__ dmask = a.dmask[dir] ? a.dmask[dir] + (long long)b * a.dmask_bs[dir] : nullptr;
    float* __restrict__ dfeat = a.dfeat[dir] ? a.dfeat[dir] + (long long)b * a.dfeat_bs[dir] : nullptr;

    for (int i = tid; i < Cf * K; i += RCF_BLOCK) cs[i] = a.poolbar[(size_t)fd * Cf * K + i];
    __syncthreads();
    const float slope = a.feat_slope;

    const int p = (blockIdx.x * RCF_BLOCK + tid) * PX;
    if (p >= P) return;
    float m[K][PX], dm[K][PX];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        Pack<PX>::ld(m[k], mask + (long long)k * P + p);
        if (dmask) {   // accumulate onto what k_bwd wrote earlier on this stream
            Pack<PX>::ld_rw(dm[k], dmask + (long long)k * P + p);
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) dm[k][j] = 0.0f;
        }
    }
#pragma unroll 4
    for (int f = 0; f < Cf; ++f) {
        float gv[PX], dg[PX];
        Pack<PX>::ld(gv, feat + (long long)f * P + p);
        float dact[PX];      // d lrelu(pre) / d pre
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            dact[j] = gv[j] >= 0.0f ? 1.0f : slope;
            gv[j] *= dact[j];
            dg[j] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float c = cs[f * K + k];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                dm[k][j] = fmaf(c, gv[j], dm[k][j]);
                dg[j] = fmaf(c, m[k][j], dg[j]);
            }
        }
        if (dfeat) {
#pragma unroll
            for (int j = 0; j < PX; ++j) dg[j] *= dact[j];
            Pack<PX>::st(dfeat + (long long)f * P + p, dg);
        }
    }
    if (dmask) {
#pragma unroll
        for (int k = 0; k < K; ++k) Pack<PX>::st(dmask + (long long)k * P + p, dm[k]);
    }
}

// ---- channels-last feature map (what cuDNN's tensor-core convolutions produce without layout transposes) ---------
// Element (f, p) of a frame sits at p*Cf + f.  A thread owns 4 consecutive channels (one float4) and walks pixels;
// 256/(Cf/4) pixel groups work side by side, so a warp reads 2 pixels = 512 contiguous bytes (Cf = 64).
// Forward: per-thread accumulators acc[4 channels][K], no shuffles; the groups are combined through shared memory.
template <int K>
__global__ void __launch_bounds__(RCF_BLOCK) k_pool_nhwc(const RcfK a) {
    rcf_pdl_prologue();
    const int CHUNK = a.poolchunk;
    extern __shared__ float sm[];
    float* msT = sm;                         // [CHUNK][K]   mask tile, pixel-major
    float* redn = sm + CHUNK * K;            // [groups][Cf*K]
    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int chunk = blockIdx.x;
    const int tid = threadIdx.x;
    const int P = a.P, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int p0 = chunk * CHUNK;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ feat = a.feat[dir] + (long long)b * a.feat_bs[dir];
    const float slope = a.feat_slope;
    float bias[4] = {0.0f, 0.0f, 0.0f, 0.0f};       // bias of the last conv, added on load (conv runs bias-free)
    if (a.feat_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bias[j] = __ldg(a.feat_bias + c4 * 4 + j);
    }

    // The first batch of feature loads is requested BEFORE the mask tile is staged: the tile's global round trip and
    // the barrier then overlap with it instead of leaving the CTA without a load in flight for the whole prologue.
    constexpr int UB = 8;
    const int pend = min(CHUNK, P - p0);
    const float4* __restrict__ gp = reinterpret_cast<const float4*>(feat + (long long)p0 * Cf) + c4;
    const int nf4s = nf4;    // float4 stride between consecutive pixels
    float4 gq[UB];
    auto load_batch = [&](int pb) {
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int p = pb + u * groups;
            gq[u] = (p < pend) ? __ldg(gp + (long long)p * nf4s) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    };
    load_batch(grp);
    for (int i = tid; i < CHUNK * K; i += RCF_BLOCK) {
        const int k = i / CHUNK, p = i - k * CHUNK;                  // coalesced along pixels per plane
        msT[p * K + k] = (p0 + p < P) ? __ldg(mask + (long long)k * P + p0 + p) : 0.0f;
    }
    __syncthreads();
    if (a.pool_sums && (tid >> 5) < K) {          // S_k partial of this chunk (reference :242-243): warp k sums column k
        const int k = tid >> 5, lane = tid & 31;
        float v0 = 0.0f, v1 = 0.0f;
        int p = lane;
        for (; p + 32 < CHUNK; p += 64) { v0 += msT[p * K + k]; v1 += msT[(p + 32) * K + k]; }
        for (; p < CHUNK; p += 32) v0 += msT[p * K + k];
        const float v = warp_sum(v0 + v1);
        if (lane == 0) a.part1[((size_t)fd * K + k) * a.nchunk1 + chunk] = v;
    }

    float acc[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[j][k] = 0.0f;
    // Explicit batches: the 8 feature loads of a batch are issued before the first one is consumed (nvcc does not
    // hoist them out of a predicated, runtime-bounded loop on its own: one load in flight per thread = 3.3 TB/s).
    for (int pb = grp; pb < pend; pb += groups * UB) {
        if (pb != grp) load_batch(pb);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int p = pb + u * groups;
            if (p < pend) {
                float gv[4] = {gq[u].x + bias[0], gq[u].y + bias[1], gq[u].z + bias[2], gq[u].w + bias[3]};
#pragma unroll
                for (int j = 0; j < 4; ++j) gv[j] = gv[j] >= 0.0f ? gv[j] : slope * gv[j];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float m = msT[p * K + k];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][k] = fmaf(gv[j], m, acc[j][k]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) redn[grp * Cf * K + (c4 * 4 + j) * K + k] = acc[j][k];
    __syncthreads();
    for (int o = tid; o < Cf * K; o += RCF_BLOCK) {
        float v = 0.0f;
        for (int g2 = 0; g2 < groups; ++g2) v += redn[g2 * Cf * K + o];      // fixed order
        a.partp[((size_t)fd * Cf * K + o) * a.nchunkp + chunk] = v;
    }
}

// Backward, channels-last: dG[p][f] = dact * sum_k c[f][k] M[k][p] (float4 store) and dM[k][p] += sum_f c[f][k] G[p][f],
// the latter reduced over the Cf/4 threads of a pixel with shuffles and accumulated in a shared tile that is written
// back to the NCHW gradient planes with coalesced stores.
template <int K>
__global__ void __launch_bounds__(RCF_BLOCK, 3) k_pool_bwd_nhwc(const RcfK a) {
    rcf_pdl_prologue();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.cnt[1] = 0u;      // arrival counter of k_bias_grad_fd (next launch)
    const int TP = a.pooltp;                 // pixels per CTA
    extern __shared__ float sm[];
    float* msT = sm;                         // [TP][K]
    float* dms = sm + TP * K;                // [TP][K]
    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int tid = threadIdx.x;
    const int P = a.P, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int p0 = blockIdx.x * TP;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ feat = a.feat[dir] + (long long)b * a.feat_bs[dir];
    float* dmask = a.dmask[dir] ? a.dmask[dir] + (long long)b * a.dmask_bs[dir] : nullptr;
    float* __restrict__ dfeat = a.dfeat[dir] ? a.dfeat[dir] + (long long)b * a.dfeat_bs[dir] : nullptr;
    const float slope = a.feat_slope;

    const int pend = min(TP, P - p0);
    const int iters = (TP + groups - 1) / groups;      // uniform trip count: shuffles below stay convergent
    const float4* __restrict__ gp = reinterpret_cast<const float4*>(feat + (long long)p0 * Cf) + c4;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    // (Requesting the first feature batch here, ahead of the tile staging, was measured: the eight live float4 push the
    // 80-register kernel into heavy spilling, 192 -> 300 us.  The forward kernel, which has the registers, does it.)
    constexpr int PBW = RCF_POOL_BWD_BATCH;          // feature loads in flight per thread on the fast path
    const bool fast = (K == 4) && (nf4 == 16) && TP % groups == 0 && (iters % PBW) == 0;
    for (int i = tid; i < TP * K; i += RCF_BLOCK) {
        const int k = i / TP, p = i - k * TP;
        const bool in = p0 + p < P;
        msT[p * K + k] = in ? __ldg(mask + (long long)k * P + p0 + p) : 0.0f;
        dms[p * K + k] = (in && dmask) ? dmask[(long long)k * P + p0 + p] : 0.0f;    // accumulate onto k_bwd's result
    }
    float c[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) c[j][k] = a.poolbar[(size_t)fd * Cf * K + (c4 * 4 + j) * K + k];
    float bias[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dbias[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (a.feat_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bias[j] = __ldg(a.feat_bias + c4 * 4 + j);
    }
    // fp16 gradient words (RcfDesc.dfeat_f16): dG is multiplied by the power of two rcf_grad_scale derives from k_segment_bwd's
    // per-frame-direction maxima before it is rounded; every consumer divides the same factor out again (exact)
    __shared__ float s_gs;
    if (tid < 32) {
        const float v = a.dfeat_f16 ? rcf_grad_scale(a.gmax, a.nfd) : 1.0f;
        if (tid == 0) s_gs = v;
    }
    __syncthreads();
    const float gs = s_gs;

    float4* __restrict__ dgp = dfeat ? reinterpret_cast<float4*>(dfeat + (long long)p0 * Cf) + c4 : nullptr;
    // bf16 (hi, lo) pair output (what the tcgen05 conv kernels load by TMA): 4 channels = one uint2 per word tensor
    uint2* __restrict__ dgh = a.dfeat_hi[dir] ? reinterpret_cast<uint2*>(a.dfeat_hi[dir] + ((long long)b * a.dfeat_bs[dir] + (long long)p0 * Cf) / 2) + c4 : nullptr;
    uint2* __restrict__ dgl = a.dfeat_lo[dir] ? reinterpret_cast<uint2*>(a.dfeat_lo[dir] + ((long long)b * a.dfeat_bs[dir] + (long long)p0 * Cf) / 2) + c4 : nullptr;
    auto store_pair = [&](int p, const float (&dg)[4]) {
        if (a.dfeat_f16) {
            dgh[(long long)p * nf4] = make_uint2(umma::cvt_f16x2(dg[0] * gs, dg[1] * gs), umma::cvt_f16x2(dg[2] * gs, dg[3] * gs));
        } else if (dgl) {
            uint32_t h0, h1, l0, l1;
            umma::split_bf16x2(dg[0], dg[1], h0, l0);
            umma::split_bf16x2(dg[2], dg[3], h1, l1);
            dgh[(long long)p * nf4] = make_uint2(h0, h1);
            dgl[(long long)p * nf4] = make_uint2(l0, l1);
        } else {
            dgh[(long long)p * nf4] = make_uint2(umma::cvt_bf16x2(dg[0], dg[1]), umma::cvt_bf16x2(dg[2], dg[3]));
        }
    };
    // fast reduction when 16 lanes share a pixel and K == 4 (Cf = 64, the reference default): recursive halving
    // (2 + 1 shuffles) then two butterflies, instead of 4 x 4 butterflies
    const int kown = ((c4 >> 3) & 1) * 2 + ((c4 >> 2) & 1);
    bool done = false;
    if constexpr (K == 4) {
        if (fast) {
            // Reference default (Cf = 64, K = 4): batches of 8 pixels per thread, all eight feature loads in flight before
            // the first is consumed; the mask quadruple is one LDS.128; no per-iteration slow-path branches.
            done = true;
            const bool up8 = (c4 & 8) != 0, up4 = (c4 & 4) != 0, writer = (c4 & 3) == 0;
            // MODE: which gradient tensors this launch writes (uniform over the grid) -- 0 none, 1 fp32, 2 bf16 hi,
            // 3 bf16 (hi, lo), 4 fp32 + pair.  Resolved once per CTA: the pixel loop carries no pointer tests.
            auto sweep = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
                for (int it0 = 0; it0 < iters; it0 += PBW) {
                    float4 gq[PBW];
#pragma unroll
                    for (int u = 0; u < PBW; ++u) {
                        const int p = grp + (it0 + u) * groups;
                        gq[u] = (p < pend) ? __ldg(gp + (long long)p * nf4) : zero4;
                    }
#pragma unroll
                    for (int u = 0; u < PBW; ++u) {
                        const int p = grp + (it0 + u) * groups;          // < TP by construction
                        const bool live = p < pend;                       // dead pixels: the mask tile holds zeros => dg = 0
                        const float4 m4 = *reinterpret_cast<const float4*>(msT + p * 4);
                        const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
                        float gv[4] = {gq[u].x + bias[0], gq[u].y + bias[1], gq[u].z + bias[2], gq[u].w + bias[3]}, dg[4], part[4];
                        float dact[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dact[j] = gv[j] >= 0.0f ? 1.0f : slope;
                            gv[j] *= dact[j];
                            dg[j] = 0.0f;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            part[k] = 0.0f;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                part[k] = fmaf(c[j][k], gv[j], part[k]);
                                dg[j] = fmaf(c[j][k], mv[k], dg[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dg[j] *= dact[j];
                            dbias[j] += dg[j];
                        }
                        if (live) {
                            if constexpr (MODE == 1 || MODE == 4) dgp[(long long)p * nf4] = make_float4(dg[0], dg[1], dg[2], dg[3]);
                            if constexpr (MODE == 2) {
                                dgh[(long long)p * nf4] = make_uint2(umma::cvt_bf16x2(dg[0], dg[1]), umma::cvt_bf16x2(dg[2], dg[3]));
                            }
                            if constexpr (MODE == 5) {        // fp16 words of the scaled gradient
                                dgh[(long long)p * nf4] = make_uint2(umma::cvt_f16x2(dg[0] * gs, dg[1] * gs), umma::cvt_f16x2(dg[2] * gs, dg[3] * gs));
                            }
                            if constexpr (MODE == 3 || MODE == 4) {
                                uint32_t h0, h1, l0, l1;
                                umma::split_bf16x2(dg[0], dg[1], h0, l0);
                                umma::split_bf16x2(dg[2], dg[3], h1, l1);
                                dgh[(long long)p * nf4] = make_uint2(h0, h1);
                                dgl[(long long)p * nf4] = make_uint2(l0, l1);
                            }
                        }
                        const float s0 = up8 ? part[0] : part[2], s1 = up8 ? part[1] : part[3];
                        const float k0 = up8 ? part[2] : part[0], k1 = up8 ? part[3] : part[1];
                        const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
                        const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
                        float v = (up4 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up4 ? a0 : a1, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        if (live && writer) dms[p * 4 + kown] += v;        // 4 lanes per pixel, one per k
                    }
                }
            };
            const int mode = dgp ? ((dgh && dgl) ? 4 : 1) : (dgh ? (a.dfeat_f16 ? 5 : (dgl ? 3 : 2)) : 0);
            if (dgp && dgh && (!dgl || a.dfeat_f16)) done = false;          // fp32 + hi only: not a combination any caller asks for
            else if (mode == 0) sweep(std::integral_constant<int, 0>{});
            else if (mode == 1) sweep(std::integral_constant<int, 1>{});
            else if (mode == 2) sweep(std::integral_constant<int, 2>{});
            else if (mode == 3) sweep(std::integral_constant<int, 3>{});
            else if (mode == 5) sweep(std::integral_constant<int, 5>{});
            else sweep(std::integral_constant<int, 4>{});
        }
    }
    if (!done) {
    // software pipeline: the loads of the next two pixels are in flight while the current one is processed
    float4 g0 = (grp < pend) ? __ldg(gp + (long long)grp * nf4) : zero4;
    float4 g1 = (grp + groups < pend) ? __ldg(gp + (long long)(grp + groups) * nf4) : zero4;
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
        const int p = grp + it * groups;
        const bool live = p < pend;
        const int p2 = p + 2 * groups;
        const float4 g2 = (p2 < pend) ? __ldg(gp + (long long)p2 * nf4) : zero4;
        const float4 g = g0;
        g0 = g1; g1 = g2;
        float part[K];
#pragma unroll
        for (int k = 0; k < K; ++k) part[k] = 0.0f;
        if (live) {
            float gv[4] = {g.x + bias[0], g.y + bias[1], g.z + bias[2], g.w + bias[3]}, dact[4], dg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dact[j] = gv[j] >= 0.0f ? 1.0f : slope;
                gv[j] *= dact[j];
                dg[j] = 0.0f;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float m = msT[p * K + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    part[k] = fmaf(c[j][k], gv[j], part[k]);
                    dg[j] = fmaf(c[j][k], m, dg[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dg[j] *= dact[j];
                dbias[j] += dg[j];
            }
            if (dgp) dgp[(long long)p * nf4] = make_float4(dg[0], dg[1], dg[2], dg[3]);
            if (dgh) store_pair(p, dg);
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
            for (int o = nf4 >> 1; o > 0; o >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
        if (live && c4 == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) dms[p * K + k] += part[k];   // one group per pixel: no conflicts
        }
    }
    }
    __syncthreads();
    if (dmask) {
        for (int i = tid; i < TP * K; i += RCF_BLOCK) {
            const int k = i / TP, p = i - k * TP;
            if (p0 + p < P) dmask[(long long)k * P + p0 + p] = dms[p * K + k];
        }
    }
    if (a.dfeat_bias) {
        // per-CTA partial of the bias gradient: combine the pixel groups in fixed order (tiles are free to reuse now)
        __syncthreads();
        float* red = sm;                       // [groups][Cf] = 1024 floats (the launch sizes shared memory for it)
        for (int j = 0; j < 4; ++j) red[grp * Cf + c4 * 4 + j] = dbias[j];
        __syncthreads();
        for (int f = tid; f < Cf; f += RCF_BLOCK) {
            float v = 0.0f;
            for (int g2 = 0; g2 < groups; ++g2) v += red[g2 * Cf + f];
            a.dbpart[((size_t)fd * a.nblkpb + blockIdx.x) * Cf + f] = v;
        }
    }
}

// Chunk partials of the pooling -> pooled sums, one warp per (frame-direction, channel, segment): thousands of warps
// instead of one CTA per frame-direction walking hundreds of chunks serially inside k_segment_fwd.
__global__ void __launch_bounds__(256) k_pool_reduce(const RcfK a) {
    rcf_pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nout = a.nfd * a.Cf * a.K;
    if (warp >= nout) return;
    const double v = warp_sum_strided(a.partp + (size_t)warp * a.nchunkp, a.nchunkp, 1, lane);
    if (lane == 0) a.poolsum[warp] = v;
}

// bias gradient of the last conv: two fixed-order levels over the per-CTA partials of k_pool_bwd_nhwc.
// Level 1, grid (S, nfd): a CTA owns a contiguous range of the nblkpb partial rows of one frame-direction; thread (f, r)
// walks rows r, r+R, ... of that range (R = 256/Cf), so a warp reads 128 contiguous bytes per row and keeps four loads in
// flight (the first version gave one warp per (fd, f) a stride-Cf walk over all rows: 17-31 us at 480x854).
__global__ void __launch_bounds__(256) k_bias_grad_fd(const RcfK a) {
    rcf_pdl_prologue();
    __shared__ double red[256];
    const int Cf = a.Cf, R = 256 / Cf, f = threadIdx.x % Cf, r = threadIdx.x / Cf;
    const int S = gridDim.x, s = blockIdx.x, fd = blockIdx.y;
    const int per = (a.nblkpb + S - 1) / S, lo = s * per, hi = min(lo + per, a.nblkpb);
    const float* __restrict__ src = a.dbpart + (size_t)fd * a.nblkpb * Cf + f;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int blk = lo + r;
    for (; blk + 3 * R < hi; blk += 4 * R) {
        const float v0 = __ldcg(src + (size_t)blk * Cf), v1 = __ldcg(src + (size_t)(blk + R) * Cf);
        const float v2 = __ldcg(src + (size_t)(blk + 2 * R) * Cf), v3 = __ldcg(src + (size_t)(blk + 3 * R) * Cf);
        a0 += (double)v0; a1 += (double)v1; a2 += (double)v2; a3 += (double)v3;
    }
    for (; blk < hi; blk += R) a0 += (double)__ldcg(src + (size_t)blk * Cf);
    red[threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (r == 0 && threadIdx.x < Cf) {
        double v = 0.0;
        for (int q = 0; q < R; ++q) v += red[q * Cf + f];
        a.dbfd[((size_t)fd * S + s) * Cf + f] = v;
    }
    // Level 2 by the last CTA to arrive (counter zeroed by k_pool_bwd_nhwc): sum the nfd * S rows of level 1 the same way.
    if (!rcf_last_cta(a.cnt + 1, gridDim.x * gridDim.y)) return;
    const int nrows = gridDim.x * gridDim.y;
    double v = 0.0;
    for (int row = r; row < nrows; row += R) v += __ldcg(a.dbfd + (size_t)row * Cf + f);
    red[threadIdx.x] = v;
    __syncthreads();
    if (r == 0 && threadIdx.x < Cf) {
        double t = 0.0;
        for (int q = 0; q < R; ++q) t += red[q * Cf + f];
        a.dfeat_bias[f] = (float)t;
    }
}

template <int K>
static cudaError_t launch_pool_nhwc_k(const RcfK& a, bool, cudaStream_t s) {
    const int groups = RCF_BLOCK / (a.Cf / 4);
    const size_t smem = ((size_t)a.poolchunk * K + (size_t)groups * a.Cf * K) * sizeof(float);
    dim3 grid(a.nchunkp, a.nfd), block(RCF_BLOCK);
    if (smem > 48 * 1024) {     // K = 7, 8 with 1024-pixel chunks: above the default dynamic shared-memory limit
        static bool allowed = false;
        if (!allowed) {
            const cudaError_t e = cudaFuncSetAttribute(k_pool_nhwc<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
            allowed = true;
        }
    }
    rcf_launch(k_pool_nhwc<K>, grid, block, smem, s, a.pdl, a);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_pool_bwd_nhwc_k(const RcfK& a, bool, cudaStream_t s) {
    dim3 grid(a.nblkpb, a.nfd), block(RCF_BLOCK);
    const size_t tile = (size_t)2 * a.pooltp * K, red = 1024;     // mask + dM tiles; [groups][Cf] = 1024 floats for the bias partial
    const size_t smem = (tile > red ? tile : red) * sizeof(float);
    if (smem > 48 * 1024) {
        static bool allowed = false;
        if (!allowed) {
            const cudaError_t e = cudaFuncSetAttribute(k_pool_bwd_nhwc<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
            allowed = true;
        }
    }
    rcf_launch(k_pool_bwd_nhwc<K>, grid, block, smem, s, a.pdl, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !a.dfeat_bias) return e;
    int S = (64 + a.nfd - 1) / a.nfd; S = S < 1 ? 1 : (S > 16 ? 16 : S);
    rcf_launch(k_bias_grad_fd, dim3(S, a.nfd), 256, 0, s, a.pdl, a);      // its last CTA also does the second level
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_pool_k(const RcfK& a, bool vec, cudaStream_t s) {
    constexpr int CHUNK = K <= 4 ? 2048 : 1024;
    // split the channels over blockIdx.z until the grid has a few CTAs per SM (each CTA keeps >= 4 channels/warp-pass)
    int fsplit = 1;
    const long long base = (long long)a.nchunkp * a.nfd;
    while (base * fsplit < 4 * 148 && a.Cf / (fsplit * 2) >= 4) fsplit *= 2;
    const int f_per_cta = (a.Cf + fsplit - 1) / fsplit;
    dim3 grid(a.nchunkp, a.nfd, (a.Cf + f_per_cta - 1) / f_per_cta), block(RCF_BLOCK);
    if (vec) rcf_launch(k_pool<K, 4, CHUNK>, grid, block, 0, s, a.pdl, a, f_per_cta);
    else rcf_launch(k_pool<K, 1, CHUNK>, grid, block, 0, s, a.pdl, a, f_per_cta);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_pool_bwd_k(const RcfK& a, bool vec, cudaStream_t s) {
    // small frames: one pixel per thread gives 4x the CTAs (the loop over Cf channels is serial per thread)
    if (vec && (long long)((a.P + RCF_BLOCK * 4 - 1) / (RCF_BLOCK * 4)) * a.nfd < 2 * 148) vec = false;
    const int px = vec ? 4 : 1;
    dim3 grid((a.P + RCF_BLOCK * px - 1) / (RCF_BLOCK * px), a.nfd), block(RCF_BLOCK);
    const size_t smem = (size_t)a.Cf * K * sizeof(float);
    if (vec) rcf_launch(k_pool_bwd<K, 4>, grid, block, smem, s, a.pdl, a);
    else rcf_launch(k_pool_bwd<K, 1>, grid, block, smem, s, a.pdl, a);
    return cudaGetLastError();
}

#define RCF_K_SWITCH(fn)                                   \
    switch (a.K) {                                         \
        case 1: return fn<1>(a, vec, s);                   \
        case 2: return fn<2>(a, vec, s);                   \
        case 3: return fn<3>(a, vec, s);                   \
        case 4: return fn<4>(a, vec, s);                   \
        case 5: return fn<5>(a, vec, s);                   \
        case 6: return fn<6>(a, vec, s);                   \
        case 7: return fn<7>(a, vec, s);                   \
        case 8: return fn<8>(a, vec, s);                   \
    }                                                      \
    return cudaErrorInvalidValue;

static cudaError_t launch_pool_dispatch(const RcfK& a, bool vec, cudaStream_t s) {
    if (a.feat_nhwc) { RCF_K_SWITCH(launch_pool_nhwc_k) }
    RCF_K_SWITCH(launch_pool_k)
}
cudaError_t rcf_launch_pool(const RcfK& a, bool vec, cudaStream_t s) {
    const cudaError_t e = launch_pool_dispatch(a, vec, s);
    if (e != cudaSuccess) return e;
    const int nout = a.nfd * a.Cf * a.K;
    rcf_launch(k_pool_reduce, (nout * 32 + 255) / 256, 256, 0, s, a.pdl, a);
    return cudaGetLastError();
}
cudaError_t rcf_launch_pool_bwd(const RcfK& a, bool vec, cudaStream_t s) {
    if (a.feat_nhwc) { RCF_K_SWITCH(launch_pool_bwd_nhwc_k) }
    RCF_K_SWITCH(launch_pool_bwd_k)
}
