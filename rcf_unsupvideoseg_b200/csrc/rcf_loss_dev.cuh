// rcf_loss_dev.cuh -- pass-2 tile body of k_loss.
#pragma once
#include "rcf_common.cuh"

template <int K, int D, int PX, bool VIS>
__device__ __forceinline__ void loss_tile(const RcfK& a, int fd, int chunk, float* cf, float (*red)[rcf_gm(K, D)]) {
    constexpr int CF = rcf_cf(D);
    constexpr int GM = rcf_gm(K, D);
    constexpr int ITER0 = RCF_CHUNK_LOSS / (RCF_BLOCK * PX);     // iterations of the short chunk (unrolled)
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
    // Affine fit without visualisation outputs: ~200 instructions per pixel in scalar form (measured 85 M warp
    // instructions, issue active 54 %, 147 us against a 112 us HBM floor), so the arithmetic is packed two ways with
    // FFMA2 / FMUL2 / FADD2 (two IEEE round-to-nearest results per issue slot, one operand may be a broadcast scalar):
    //   phase 1 over PIXEL pairs (the two halves of a 128-bit load are neighbours in the register file),
    //   phase 2 over the two FLOW COMPONENTS (w_0, w_1 of a pixel), whose accumulators are pairs as well.
    // Same products and the same summation order per accumulator as the scalar code.
    constexpr bool PACKED = (D == 2) && !VIS && (PX % 2 == 0);
    f32x2 accw[PACKED ? K : 1], accv[PACKED ? K * 2 : 1];
#pragma unroll
    for (int i = 0; i < (PACKED ? K : 1); ++i) accw[i] = 0ull;
#pragma unroll
    for (int i = 0; i < (PACKED ? K * 2 : 1); ++i) accv[i] = 0ull;

    const int p0 = chunk * a.chunk2;
    const int outer = (D == 2 && K <= 4) ? a.chunk2 / RCF_CHUNK_LOSS : 1;      // see rcf_chunk_loss
    const bool hints = a.l2_hints != 0;
    float r[2][K][PX];
    auto load_resid = [&](int p) {
        if (hints) {
            const unsigned long long pol = l2_policy_evict_first();
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) ld_evict_first(r[c][k], resid + (long long)(c * K + k) * P + p, pol);
        } else {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) Pack<PX>::ld(r[c][k], resid + (long long)(c * K + k) * P + p);
        }
    };
    // PACKED: the residual registers are dead once phase 1 has turned them into the prediction, so the residual packs
    // of the NEXT tile are requested right there and fly during the loss terms and phase 2 (split-phase prefetch).
    constexpr bool PREFETCH = PACKED && K <= 4;       // K > 4 has no registers to spare (spills)
    if constexpr (PREFETCH) {
        if (p0 + tid * PX < P) load_resid(p0 + tid * PX);
    }
#pragma unroll 1
    for (int ot = 0; ot < outer; ++ot) {
#pragma unroll
    for (int it0 = 0; it0 < ITER0; ++it0) {
        const int it = ot * ITER0 + it0;
        const int p = p0 + (it * RCF_BLOCK + tid) * PX;
        [[maybe_unused]] const int pn = p + RCF_BLOCK * PX;
        [[maybe_unused]] const bool more = (it + 1 < outer * ITER0) && pn < P;
        if (p < P) {
            float m[K][PX], f[2][PX];
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::ld(m[k], mask + (long long)k * P + p);
            if (hints) {
                const unsigned long long pol = l2_policy_evict_first();
                ld_evict_first(f[0], flow + p, pol);
                ld_evict_first(f[1], flow + P + p, pol);
            } else {
                Pack<PX>::ld(f[0], flow + p);
                Pack<PX>::ld(f[1], flow + P + p);
            }
            if constexpr (!PREFETCH) load_resid(p);
            float y[PX], x[PX];
            if constexpr (D > 0) px_coords<PX>(p, a, y, x);

            if constexpr (PACKED) {
                constexpr int PP = PX / 2;
                f32x2 up[2][PP], predp[2][PP];
#pragma unroll
                for (int h = 0; h < PP; ++h) {
                    up[0][h] = pack2(y[2 * h], y[2 * h + 1]);
                    up[1][h] = pack2(x[2 * h], x[2 * h + 1]);
                    predp[0][h] = predp[1][h] = 0ull;
                }
                const f32x2 one2 = pack2(1.0f, 1.0f), mtwo2 = pack2(-2.0f, -2.0f), sc2 = pack2(a.scale, a.scale);
                const f32x2 es2 = pack2(a.ex2_scale, a.ex2_scale);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float ck[CF];
#pragma unroll
                    for (int i = 0; i < CF; ++i) ck[i] = cf[k * CF + i];
                    const f32x2 nmu0 = pack2(-ck[6], -ck[6]), nmu1 = pack2(-ck[7], -ck[7]);
#pragma unroll
                    for (int h = 0; h < PP; ++h) {
                        const f32x2 v0 = add2(up[0][h], nmu0), v1 = add2(up[1][h], nmu1);
                        const f32x2 mp = pack2(m[k][2 * h], m[k][2 * h + 1]);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            f32x2 t = pack2(r[c][k][2 * h], r[c][k][2 * h + 1]);
                            if (!a.unbounded) {
                                float e0, e1;
                                unpack2(mul2(t, es2), e0, e1);
                                float q0, q1;
                                unpack2(add2(pack2(fast_ex2(e0), fast_ex2(e1)), one2), q0, q1);
                                t = fma2(mtwo2, pack2(fast_rcp(q0), fast_rcp(q1)), one2);
                            }
                            f32x2 q = fma2(sc2, t, pack2(ck[c], ck[c]));
                            q = fma2(pack2(ck[2 + c * 2], ck[2 + c * 2]), v0, q);
                            q = fma2(pack2(ck[3 + c * 2], ck[3 + c * 2]), v1, q);
                            predp[c][h] = fma2(mp, q, predp[c][h]);
                        }
                    }
                }
                if constexpr (PREFETCH) {
                    if (more) load_resid(pn);
                }
                // loss value and derivative weights: w_c of pixel j replaces f[c][j]
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int h = 0; h < PP; ++h) {
                        float pr[2];
                        unpack2(predp[c][h], pr[0], pr[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float fc = clamp_flow(f[c][2 * h + e], a.clamp_t);
                            float phi, w;
                            loss_terms(fc - pr[e], a, phi, w);
                            acc[0] += phi;
                            f[c][2 * h + e] = w;
                        }
                    }
                // gradient moments, the two flow components as one pair
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float mu0 = cf[k * CF + 6], mu1 = cf[k * CF + 7];
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        const float mk = m[k][j];
                        const f32x2 wm = mul2(pack2(f[0][j], f[1][j]), pack2(mk, mk));
                        const float v0 = y[j] - mu0, v1 = x[j] - mu1;
                        accw[k] = add2(accw[k], wm);
                        accv[k * 2 + 0] = fma2(wm, pack2(v0, v0), accv[k * 2 + 0]);
                        accv[k * 2 + 1] = fma2(wm, pack2(v1, v1), accv[k * 2 + 1]);
                    }
                }
            } else {
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
            }   // !PACKED
        }
    }
    }
    if constexpr (PACKED) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            unpack2(accw[k], acc[1 + k], acc[1 + K + k]);
            unpack2(accv[k * 2 + 0], acc[1 + 2 * K + (k * 2 + 0) * 2 + 0], acc[1 + 2 * K + (k * 2 + 1) * 2 + 0]);
            unpack2(accv[k * 2 + 1], acc[1 + 2 * K + (k * 2 + 0) * 2 + 1], acc[1 + 2 * K + (k * 2 + 1) * 2 + 1]);
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

