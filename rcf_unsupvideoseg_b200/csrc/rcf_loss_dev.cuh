// rcf_loss_dev.cuh -- pass-2 tile body (shared by k_loss and the fused forward kernel).
#pragma once
#include "rcf_common.cuh"

template <int K, int D, int PX, bool VIS>
__device__ __forceinline__ void loss_tile(const RcfK& a, int fd, int chunk, float* cf, float (*red)[rcf_gm(K, D)]) {
    constexpr int CF = rcf_cf(D);
    constexpr int GM = rcf_gm(K, D);
    constexpr int ITER = RCF_CHUNK_LOSS / (RCF_BLOCK * PX);
    constexpr int DD = D > 0 ? D : 1;

    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ flow = a.flow[dir] + (long long)b * a.flow_bs[dir];
    const float* __restrict__ resid = a.resid[dir] + (long long)b * a.resid_bs[dir];

    if (D == 0 && a.single_pass) {      // theta supplied: the coefficient pack {theta_0k, theta_1k} straight from theta [B,2,K]
        for (int i = tid; i < K * CF; i += RCF_BLOCK) cf[i] = __ldg(a.theta[dir] + (size_t)b * 2 * K + (i & 1) * K + (i >> 1));
    } else {
        for (int i = tid; i < K * CF; i += RCF_BLOCK) cf[i] = __ldcg(a.coef + (size_t)fd * K * CF + i);
    }
    __syncthreads();

    const long long vis_off = (long long)b * a.vis_bs + (long long)dir * a.vis_ds;

    float acc[GM];
#pragma unroll
    for (int s = 0; s < GM; ++s) acc[s] = 0.0f;

    const int p0 = chunk * RCF_CHUNK_LOSS;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = p0 + (it * RCF_BLOCK + tid) * PX;
        if (p < P) {
            float m[K][PX], r[2][K][PX], f[2][PX];
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::ld(m[k], mask + (long long)k * P + p);
            if (a.l2_hints) {
                const unsigned long long pol = l2_policy_evict_first();
                ld_evict_first(f[0], flow + p, pol);
                ld_evict_first(f[1], flow + P + p, pol);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < K; ++k) ld_evict_first(r[c][k], resid + (long long)(c * K + k) * P + p, pol);
            } else {
                Pack<PX>::ld(f[0], flow + p);
                Pack<PX>::ld(f[1], flow + P + p);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < K; ++k) Pack<PX>::ld(r[c][k], resid + (long long)(c * K + k) * P + p);
            }
            float y[PX], x[PX];
            if constexpr (D > 0) px_coords<PX>(p, a, y, x);

            float pred[2][PX], agg[2][PX], aff[2][PX];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < PX; ++j) { pred[c][j] = 0.0f; agg[c][j] = 0.0f; aff[c][j] = 0.0f; }

            // ---- phase 1: reconstruct (segment-outer) -------------------------------------------
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float ck[CF];
#pragma unroll
                for (int i = 0; i < CF; ++i) ck[i] = cf[k * CF + i];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float mk = m[k][j];
                    float u[DD];
                    if constexpr (D > 0) px_feats<D>(y[j], x[j], u);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float t = a.unbounded ? r[c][k][j] : tanh_scaled(r[c][k][j], a.ex2_scale);
                        if constexpr (VIS) {
                            agg[c][j] = fmaf(mk, ck[c], agg[c][j]);
                            pred[c][j] = fmaf(mk, t, pred[c][j]);          // residual part (unscaled)
                            if constexpr (D > 0) {
                                float av = 0.0f;
#pragma unroll
                                for (int d = 0; d < D; ++d) av = fmaf(ck[2 + c * D + d], u[d] - ck[2 + 2 * D + d], av);
                                aff[c][j] = fmaf(mk, av, aff[c][j]);
                            }
                        } else {
                            float q = fmaf(a.scale, t, ck[c]);
                            if constexpr (D > 0) {
#pragma unroll
                                for (int d = 0; d < D; ++d) q = fmaf(ck[2 + c * D + d], u[d] - ck[2 + 2 * D + d], q);
                            }
                            pred[c][j] = fmaf(mk, q, pred[c][j]);
                        }
                    }
                }
            }
            // ---- loss value and derivative weights (w overwrites f) -------------------------------
            float o_gt[2][PX], o_res[2][PX];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float fc = clamp_flow(f[c][j], a.clamp_t);
                    if constexpr (VIS) {
                        o_gt[c][j] = fc;
                        o_res[c][j] = pred[c][j] * a.scale;
                        pred[c][j] = agg[c][j] + aff[c][j] + o_res[c][j];
                    }
                    float phi, w;
                    loss_terms(fc - pred[c][j], a, phi, w);
                    acc[0] += phi;
                    f[c][j] = w;
                }
            // ---- phase 2: gradient moments (segment-outer) --------------------------------------
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float mu[DD];
                if constexpr (D > 0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) mu[d] = cf[k * CF + 2 + 2 * D + d];
                }
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float wm0 = f[0][j] * m[k][j], wm1 = f[1][j] * m[k][j];
                    acc[1 + k] += wm0;
                    acc[1 + K + k] += wm1;
                    if constexpr (D == 0) acc[1 + 2 * K + k] += m[k][j];      // S_k (see rcf_gm)
                    if constexpr (D > 0) {
                        float u[DD];
                        px_feats<D>(y[j], x[j], u);
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const float v = u[d] - mu[d];
                            acc[1 + 2 * K + (k * 2 + 0) * D + d] = fmaf(wm0, v, acc[1 + 2 * K + (k * 2 + 0) * D + d]);
                            acc[1 + 2 * K + (k * 2 + 1) * D + d] = fmaf(wm1, v, acc[1 + 2 * K + (k * 2 + 1) * D + d]);
                        }
                    }
                }
            }
            if constexpr (VIS) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float sc = a.vis_scale[c];
                    const long long o = vis_off + (long long)c * P + p;
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        o_gt[c][j] *= sc; pred[c][j] *= sc; agg[c][j] *= sc; o_res[c][j] *= sc; aff[c][j] *= sc;
                    }
                    if (a.vis_gt) Pack<PX>::st(a.vis_gt + o, o_gt[c]);
                    if (a.vis_pred) Pack<PX>::st(a.vis_pred + o, pred[c]);
                    if (a.vis_agg) Pack<PX>::st(a.vis_agg + o, agg[c]);
                    if (a.vis_res) Pack<PX>::st(a.vis_res + o, o_res[c]);
                    if (D > 0 && a.vis_aff) Pack<PX>::st(a.vis_aff + o, aff[c]);
                }
            }
        }
    }

    warp_reduce_store<GM>(acc, lane, red[warp]);
    __syncthreads();
    for (int i = tid; i < GM; i += RCF_BLOCK) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < RCF_WARPS; ++w) v += red[w][i];
        a.part2[((size_t)fd * GM + i) * a.nchunk2 + chunk] = v;
    }
}

