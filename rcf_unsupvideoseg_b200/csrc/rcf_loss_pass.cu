// rcf_loss_pass.cu -- pass 2: fused reconstruct + residual + robust norm + loss + gradient moments.
//
// Per pixel (reference lines in brackets):
//   agg_c  = sum_k theta_ck m_k                                   [:260-265]
//   aff_c  = sum_k m_k A_kc.(u - mu_k)                            [:223-231]
//   res_c  = s * sum_k tanh(r_ck / div) m_k   (or sum_k r_ck m_k) [:279-286, :302-303]
//   pred_c = agg_c + aff_c + res_c                                [:288, :304]
//   phi(F_c - pred_c) summed for the mean                         [:359-368]
// and, in the same pass, the "gradient moments" sum_p w_c m_k and sum_p w_c m_k (u - mu_k)
// (w = phi'), which is everything the per-segment backward needs, so the
// backward never has to reduce anything.  Optionally writes the visualisation flows [:370-395].
//
// HBM-bound: algorithmic bytes per pixel = 4K + 8 + 8K read (+ up to 40 written when vis is on).
#include "rcf_common.cuh"

template <int K, int D, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_loss(const RcfK a) {
    constexpr int CF = rcf_cf(D);
    constexpr int GM = rcf_gm(K, D);
    constexpr int ITER = RCF_CHUNK_LOSS / (RCF_BLOCK * PX);
    constexpr int DD = D > 0 ? D : 1;
    __shared__ float cf[K * CF];
    __shared__ float red[RCF_WARPS][GM];

    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ flow = a.flow[dir] + (long long)b * a.flow_bs[dir];
    const float* __restrict__ resid = a.resid[dir] + (long long)b * a.resid_bs[dir];

    for (int i = tid; i < K * CF; i += RCF_BLOCK) cf[i] = a.coef[(size_t)fd * K * CF + i];
    __syncthreads();

    const bool want_vis = (a.vis_gt != nullptr) | (a.vis_pred != nullptr) | (a.vis_agg != nullptr) |
                          (a.vis_res != nullptr) | (a.vis_aff != nullptr);
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
            Pack<PX>::ld(f[0], flow + p);
            Pack<PX>::ld(f[1], flow + P + p);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < K; ++k) Pack<PX>::ld(r[c][k], resid + (long long)(c * K + k) * P + p);
            float y[PX], x[PX];
            if constexpr (D > 0) px_coords<PX>(p, a, y, x);

            float o_gt[2][PX], o_pred[2][PX], o_agg[2][PX], o_res[2][PX], o_aff[2][PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float u[DD];
                if constexpr (D > 0) px_feats<D>(y[j], x[j], u);
                float agg[2] = {0.0f, 0.0f}, aff[2] = {0.0f, 0.0f}, res[2] = {0.0f, 0.0f};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float mk = m[k][j];
                    const float* ck = cf + k * CF;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        agg[c] = fmaf(mk, ck[c], agg[c]);
                        if constexpr (D > 0) {
                            float av = 0.0f;
#pragma unroll
                            for (int d = 0; d < D; ++d) av = fmaf(ck[2 + c * D + d], u[d] - ck[2 + 2 * D + d], av);
                            aff[c] = fmaf(mk, av, aff[c]);
                        }
                        const float t = a.unbounded ? r[c][k][j] : tanh_scaled(r[c][k][j], a.ex2_scale);
                        res[c] = fmaf(mk, t, res[c]);
                    }
                }
                float w[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float fc = clamp_flow(f[c][j], a.clamp_t);
                    res[c] *= a.scale;
                    const float pred = agg[c] + aff[c] + res[c];
                    float phi;
                    loss_terms(fc - pred, a, phi, w[c]);
                    acc[0] += phi;
                    o_gt[c][j] = fc; o_pred[c][j] = pred; o_agg[c][j] = agg[c]; o_res[c][j] = res[c]; o_aff[c][j] = aff[c];
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float wm = w[c] * m[k][j];
                        acc[1 + c * K + k] += wm;
                        if constexpr (D > 0) {
#pragma unroll
                            for (int d = 0; d < D; ++d)
                                acc[1 + 2 * K + (k * 2 + c) * D + d] =
                                    fmaf(wm, u[d] - cf[k * CF + 2 + 2 * D + d], acc[1 + 2 * K + (k * 2 + c) * D + d]);
                        }
                    }
                }
            }
            if (want_vis) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float sc = a.vis_scale[c];
                    const long long o = vis_off + (long long)c * P + p;
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        o_gt[c][j] *= sc; o_pred[c][j] *= sc; o_agg[c][j] *= sc; o_res[c][j] *= sc; o_aff[c][j] *= sc;
                    }
                    if (a.vis_gt) Pack<PX>::st(a.vis_gt + o, o_gt[c]);
                    if (a.vis_pred) Pack<PX>::st(a.vis_pred + o, o_pred[c]);
                    if (a.vis_agg) Pack<PX>::st(a.vis_agg + o, o_agg[c]);
                    if (a.vis_res) Pack<PX>::st(a.vis_res + o, o_res[c]);
                    if (D > 0 && a.vis_aff) Pack<PX>::st(a.vis_aff + o, o_aff[c]);
                }
            }
        }
    }

#pragma unroll
    for (int s = 0; s < GM; ++s) {
        const float v = warp_sum(acc[s]);
        if (lane == 0) red[warp][s] = v;
    }
    __syncthreads();
    for (int i = tid; i < GM; i += RCF_BLOCK) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < RCF_WARPS; ++w) v += red[w][i];
        a.part2[((size_t)fd * GM + i) * a.nchunk2 + chunk] = v;
    }
}

template <int K, int D>
static cudaError_t launch_kd(const RcfK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunk2, a.nfd), block(RCF_BLOCK);
    if (vec) k_loss<K, D, 4><<<grid, block, 0, s>>>(a);
    else k_loss<K, D, 1><<<grid, block, 0, s>>>(a);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_k(const RcfK& a, bool vec, cudaStream_t s) {
    switch (a.D) {
        case 0: return launch_kd<K, 0>(a, vec, s);
        case 2: return launch_kd<K, 2>(a, vec, s);
        case 5: return launch_kd<K, 5>(a, vec, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t rcf_launch_loss(const RcfK& a, bool vec, cudaStream_t s) {
    switch (a.K) {
        case 1: return launch_k<1>(a, vec, s);
        case 2: return launch_k<2>(a, vec, s);
        case 3: return launch_k<3>(a, vec, s);
        case 4: return launch_k<4>(a, vec, s);
        case 5: return launch_k<5>(a, vec, s);
        case 6: return launch_k<6>(a, vec, s);
        case 7: return launch_k<7>(a, vec, s);
        case 8: return launch_k<8>(a, vec, s);
    }
    return cudaErrorInvalidValue;
}
