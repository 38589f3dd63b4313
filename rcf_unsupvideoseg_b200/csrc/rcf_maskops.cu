// rcf_maskops.cu -- caller-side mask preparation fused into one pass each way.
//
// Reference (models/rcf_model.py): the segmentation logits [B, I, K, H, W] become the soft masks the motion loss
// consumes, and the same tensor feeds the entropy regulariser of stage 1 (configs/rcf/rcf_stage1.yaml:67, w_entropy 0.05):
//   :433      all_pred_mask     = F.softmax(all_pred_mask, dim=2)
//   :434      log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)       (sic: log-softmax OF the probabilities)
//   :376-378  entropy           = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
// ATen runs this as softmax, log_softmax, mul, sum, mean (+ their five backward kernels, + the add that merges the two
// gradient streams into the softmax backward).  Here:
//   k_mask_fwd : one read of the logits -> masks written once + per-CTA partials of every loss (fixed-order reduction:
//                k_mask_frame sums a frame's partials, k_mask_final sums the frames)
//   k_mask_bwd : one read of (masks, dL/dmasks) -> dL/dlogits, with the entropy gradient folded in:
//        ls = log_softmax(m), q = exp(ls), Sm = sum_k m_k
//        dE/dm_j    = -(ls_j + m_j - q_j * Sm) / Npix
//        gm_j       = gmask_j + gE * dE/dm_j
//        dlogit_j   = m_j * (gm_j - sum_k m_k gm_k)
// The same read of the mask tensor also feeds the other caller-side mask losses the shipped configs enable:
//   compactness (models/compactness_head.py:33-56; STv2 stage 1, w_compactness 1.0): soft centroid of one channel per
//       frame and mean(((y-yc)^2 + (x-xc)^2) * m) -- four moment sums per frame in the forward pass; the dependence of the
//       centroid on m cancels in the gradient (sum_q m_q (y_q - yc) = 0), so dL/dm_p = ((y_p-yc)^2 + (x_p-xc)^2) / Npix;
//   PL / CRF loss (models/rcf_model.py:380-408; stages 2.1 / 2.2): positive/negative weighted MSE between the object
//       channel and a (optionally thresholded) target mask.
// HBM-bound: 8K B/px forward, 12K B/px backward (fp32), 128-bit accesses, no intermediates.
#include "rcf_common.cuh"

namespace {

// per-CTA partial sums: entropy, compactness moments (S, Sy, Sx, Sq) of the compact channel, PL/CRF squared errors (pos, neg),
// sharpening loss
constexpr int MASK_NP = 8;

struct MaskK {
    const float* logits;
    float* masks;
    float* part;         // [nframes * nchunk][MASK_NP]
    float* losses;       // [4]: entropy, compactness, pl/crf, sharpen
    float* fstats;       // [nframes][2]: (y_center, x_center) of the compact channel (forward -> backward)
    double* ftot;        // ws tail: [nframes][8] per-frame loss terms (5 used)
    const float* gmasks; // may be null
    const float* glosses;// device [4] (d/d entropy, d/d compactness, d/d pl, d/d sharpen), may be null
    float* dlogits;
    const float* target; // PL / CRF masks [nframes, P] or null
    int nframes, P, H, W, nchunk;
    int compact_ch;      // -1: off
    int pl_ch;           // object channel (PL / CRF loss), -1: off
    int pl_binarize;     // target = target > pl_th
    float pl_th, pl_wpos, pl_wneg;
    int sharpen_mode;    // 0 off, 1 KL to the sharpened masks (rcf_model.py:370-373), 2 object-aware hinge (:362-369)
    int sharpen_ch;      // object channel of mode 2
    float t_sharpen, inv_t;
    int K;
    float inv_npix;      // 1 / (nframes * P)
    float inv_h, inv_w;
};

constexpr int MASK_CHUNK = 2048;

template <int K, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_mask_fwd(const MaskK a) {
    asm volatile("griddepcontrol.launch_dependents;");      // k_mask_final may be scheduled while this grid drains
    constexpr int ITER = MASK_CHUNK / (RCF_BLOCK * PX);
    __shared__ float red[RCF_WARPS][MASK_NP];
    const int fr = blockIdx.y, tid = threadIdx.x;
    const float* __restrict__ lg = a.logits + (size_t)fr * K * a.P;
    float* __restrict__ mk = a.masks + (size_t)fr * K * a.P;
    const float* __restrict__ tg = a.target ? a.target + (size_t)fr * a.P : nullptr;
    float acc[MASK_NP];
#pragma unroll
    for (int i = 0; i < MASK_NP; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = blockIdx.x * MASK_CHUNK + (it * RCF_BLOCK + tid) * PX;
        if (p < a.P) {
            float x[K][PX], t[PX];
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::ld(x[k], lg + (size_t)k * a.P + p);
            if (tg) Pack<PX>::ld(t, tg + p);
            int row = p / a.W, col = p - row * a.W;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float mx = x[0][j];
#pragma unroll
                for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k][j]);
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) { x[k][j] = __expf(x[k][j] - mx); s += x[k][j]; }
                const float inv = 1.0f / s;
                float s2 = 0.0f, dot = 0.0f, sm = 0.0f, mc = 0.0f, mo = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float m = x[k][j] * inv;
                    x[k][j] = m;
                    s2 += __expf(m);              // m in [0, 1]: no max subtraction needed
                    dot = fmaf(m, m, dot);
                    sm += m;
                    mc = (k == a.compact_ch) ? m : mc;
                    mo = (k == a.pl_ch) ? m : mo;
                }
                // -(sum_k m_k (m_k - lse)) = lse * sum m - sum m^2
                acc[0] += __logf(s2) * sm - dot;
                if (a.compact_ch >= 0) {          // compactness_head.py:33-56: soft count and first/second moments
                    const float yy = (float)row * a.inv_h, xx = (float)col * a.inv_w;
                    acc[1] += mc;
                    acc[2] = fmaf(mc, yy, acc[2]);
                    acc[3] = fmaf(mc, xx, acc[3]);
                    acc[4] = fmaf(mc, fmaf(yy, yy, xx * xx), acc[4]);
                }
                if (tg) {                          // rcf_model.py:380-408: weighted MSE towards the PL / CRF mask
                    const float tv = a.pl_binarize ? (t[j] > a.pl_th ? 1.0f : 0.0f) : t[j];
                    const float d = tv - mo;
                    const float dp = fmaxf(d, 0.0f), dn = fminf(d, 0.0f);
                    acc[5] = fmaf(dp, dp, acc[5]);
                    acc[6] = fmaf(dn, dn, acc[6]);
                }
                if (a.sharpen_mode == 1) {
                    // target = sharpen(m, T) = m^(1/T) / sum m^(1/T) (utils/loss_utils.py:105-108, detached);
                    // F.kl_div(log_softmax(m), target, 'none') = target * (log target - (m - lse)), averaged over all elements
                    const float lse = __logf(s2);
                    float lg[K], spw = 0.0f;
#pragma unroll
                    for (int k = 0; k < K; ++k) { lg[k] = __logf(x[k][j]) * a.inv_t; spw += __expf(lg[k]); }
                    const float lsp = __logf(spw);
                    float term = 0.0f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float lt = lg[k] - lsp;                      // log target_k
                        const float tk = __expf(lt);
                        term += (tk > 0.0f) ? tk * (lt - (x[k][j] - lse)) : 0.0f;
                    }
                    acc[7] += term;
                } else if (a.sharpen_mode == 2) {
                    // hinge on |m_obj - max of the other (detached) channels| (rcf_model.py:362-369)
                    float mobj = 0.0f, mx = 0.0f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        mobj = (k == a.sharpen_ch) ? x[k][j] : mobj;
                        mx = (k == a.sharpen_ch) ? mx : fmaxf(mx, x[k][j]);
                    }
                    acc[7] += fmaxf(a.t_sharpen - fabsf(mobj - mx), 0.0f);
                }
                if (++col == a.W) { col = 0; ++row; }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::st(mk + (size_t)k * a.P + p, x[k]);
        }
    }
    warp_reduce_store<MASK_NP>(acc, tid & 31, red[tid >> 5]);
    __syncthreads();
    if (tid < MASK_NP) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < RCF_WARPS; ++w) v += red[w][tid];
        a.part[((size_t)fr * a.nchunk + blockIdx.x) * MASK_NP + tid] = v;
    }
}

// Level 1, one CTA per frame: fp64 sums of that frame's per-CTA partials (thread = chunk, fixed-order block reduction),
// the compactness centroid of the frame, and the frame's four loss terms -> ftot[frame][4].
__global__ void __launch_bounds__(256) k_mask_frame(const MaskK a) {
    rcf_pdl_prologue();
    __shared__ double red[8][MASK_NP];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double v[MASK_NP];
#pragma unroll
    for (int i = 0; i < MASK_NP; ++i) v[i] = 0.0;
    for (int c = tid; c < a.nchunk; c += 256) {
        const float* q = a.part + ((size_t)f * a.nchunk + c) * MASK_NP;
#pragma unroll
        for (int i = 0; i < MASK_NP; ++i) v[i] += (double)__ldcg(q + i);
    }
#pragma unroll
    for (int i = 0; i < MASK_NP; ++i) v[i] = warp_sum_d(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < MASK_NP; ++i) red[warp][i] = v[i];
    }
    __syncthreads();
    if (tid == 0) {
        double t[MASK_NP];
#pragma unroll
        for (int i = 0; i < MASK_NP; ++i) {
            t[i] = 0.0;
            for (int w = 0; w < 8; ++w) t[i] += red[w][i];
        }
        double comp = 0.0;
        if (a.compact_ch >= 0) {
            const double yc = t[2] / t[1], xc = t[3] / t[1];
            comp = t[4] - (t[2] * t[2] + t[3] * t[3]) / t[1];               // sum m ((y-yc)^2 + (x-xc)^2)
            if (a.fstats) { a.fstats[2 * f] = (float)yc; a.fstats[2 * f + 1] = (float)xc; }
        }
        double* o = a.ftot + (size_t)f * 8;
        o[0] = t[0]; o[1] = comp; o[2] = t[5]; o[3] = t[6]; o[4] = t[7];
    }
}

// Level 2, one warp: the frames in a fixed order.
__global__ void __launch_bounds__(32) k_mask_final(const MaskK a) {
    rcf_pdl_prologue();
    const int lane = threadIdx.x;
    double t[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int f = lane; f < a.nframes; f += 32) {
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] += a.ftot[(size_t)f * 8 + i];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) t[i] = warp_sum_d(t[i]);
    if (lane == 0) {
        a.losses[0] = (float)(t[0] * (double)a.inv_npix);
        a.losses[1] = (float)(t[1] * (double)a.inv_npix);
        a.losses[2] = (float)((t[2] * (double)a.pl_wpos + t[3] * (double)a.pl_wneg) * (double)a.inv_npix);
        // KL: mean over all B*I*K*H*W elements; hinge: mean over B*I*H*W pixels
        a.losses[3] = (float)(t[4] * (double)a.inv_npix / (a.sharpen_mode == 1 ? (double)a.K : 1.0));
    }
}

template <int K, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_mask_bwd(const MaskK a) {
    constexpr int ITER = MASK_CHUNK / (RCF_BLOCK * PX);
    const int fr = blockIdx.y, tid = threadIdx.x;
    const float* __restrict__ mk = a.masks + (size_t)fr * K * a.P;
    const float* __restrict__ gm = a.gmasks ? a.gmasks + (size_t)fr * K * a.P : nullptr;
    const float* __restrict__ tg = a.target ? a.target + (size_t)fr * a.P : nullptr;
    float* __restrict__ dl = a.dlogits + (size_t)fr * K * a.P;
    const float ge = a.glosses ? -__ldg(a.glosses) * a.inv_npix : 0.0f;
    const float gc = (a.glosses && a.compact_ch >= 0) ? __ldg(a.glosses + 1) * a.inv_npix : 0.0f;
    const float gp = (a.glosses && tg) ? -2.0f * __ldg(a.glosses + 2) * a.inv_npix : 0.0f;
    const float gsh = (a.glosses && a.sharpen_mode) ? __ldg(a.glosses + 3) * a.inv_npix / (a.sharpen_mode == 1 ? (float)K : 1.0f) : 0.0f;
    const float yc = (a.compact_ch >= 0 && a.fstats) ? __ldg(a.fstats + 2 * fr) : 0.0f;
    const float xc = (a.compact_ch >= 0 && a.fstats) ? __ldg(a.fstats + 2 * fr + 1) : 0.0f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = blockIdx.x * MASK_CHUNK + (it * RCF_BLOCK + tid) * PX;
        if (p < a.P) {
            float m[K][PX], g[K][PX], t[PX];
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::ld(m[k], mk + (size_t)k * a.P + p);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (gm) Pack<PX>::ld(g[k], gm + (size_t)k * a.P + p);
                else {
#pragma unroll
                    for (int j = 0; j < PX; ++j) g[k][j] = 0.0f;
                }
            }
            float mob[PX];      // the object channel, re-read (L1 hit): selecting m[pl_ch] from registers makes nvcc spill m[][] to a
                                // local array and index it (80-128 B stack frame for K > 4)
            if (tg) { Pack<PX>::ld(t, tg + p); Pack<PX>::ld(mob, mk + (size_t)a.pl_ch * a.P + p); }
            int row = p / a.W, col = p - row * a.W;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float s2 = 0.0f, sm = 0.0f;
                float e[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { e[k] = __expf(m[k][j]); s2 += e[k]; sm += m[k][j]; }
                const float lse = __logf(s2), inv2 = 1.0f / s2;
                // extra per-channel gradient terms (compactness on compact_ch, PL / CRF on pl_ch)
                const float dy = (float)row * a.inv_h - yc, dx = (float)col * a.inv_w - xc;
                const float gcomp = gc * fmaf(dy, dy, dx * dx);
                float gpl = 0.0f;
                if (tg) {
                    const float mo = mob[j];
                    const float tv = a.pl_binarize ? (t[j] > a.pl_th ? 1.0f : 0.0f) : t[j];
                    const float d = tv - mo;
                    gpl = gp * (a.pl_wpos * fmaxf(d, 0.0f) + a.pl_wneg * fminf(d, 0.0f));
                }
                // sharpening: KL -> d/dm_j = -(target_j - q_j) / Nel (target detached); hinge -> -sign(m_obj - mx) on the object channel
                float tsh[K];
                float ghinge = 0.0f;
                if (a.sharpen_mode == 1) {
                    float spw = 0.0f;
#pragma unroll
                    for (int k = 0; k < K; ++k) { tsh[k] = __expf(__logf(m[k][j]) * a.inv_t); spw += tsh[k]; }
                    const float isp = 1.0f / spw;
#pragma unroll
                    for (int k = 0; k < K; ++k) tsh[k] = -gsh * (tsh[k] * isp - e[k] * inv2);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) tsh[k] = 0.0f;
                    if (a.sharpen_mode == 2) {
                        float mobj = 0.0f, mx = 0.0f;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            mobj += (k == a.sharpen_ch) ? e[k] : 0.0f;           // exp is monotone: compare exp(m) to stay in registers
                            mx = (k == a.sharpen_ch) ? mx : fmaxf(mx, e[k]);
                        }
                        const float d = __logf(mobj) - __logf(fmaxf(mx, 1.0f));      // = m_obj - max_other (exp(0) = 1 when K == 1)
                        ghinge = (a.t_sharpen - fabsf(d) > 0.0f) ? -gsh * ((d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f)) : 0.0f;
                    }
                }
                float dot = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float mkj = m[k][j];
                    const float dE = (mkj - lse) + mkj - e[k] * inv2 * sm;     // ls_j + m_j - q_j * Sm
                    float gk = fmaf(ge, dE, g[k][j]) + tsh[k];
                    gk += (k == a.sharpen_ch && a.sharpen_mode == 2) ? ghinge : 0.0f;
                    gk += (k == a.compact_ch) ? gcomp : 0.0f;
                    gk += (k == a.pl_ch) ? gpl : 0.0f;
                    g[k][j] = gk;
                    dot = fmaf(mkj, gk, dot);
                }
#pragma unroll
                for (int k = 0; k < K; ++k) g[k][j] = m[k][j] * (g[k][j] - dot);
                if (++col == a.W) { col = 0; ++row; }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::st(dl + (size_t)k * a.P + p, g[k]);
        }
    }
}

template <int K>
cudaError_t launch_fwd(const MaskK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunk, a.nframes);
    if (vec) k_mask_fwd<K, 4><<<grid, RCF_BLOCK, 0, s>>>(a);
    else k_mask_fwd<K, 1><<<grid, RCF_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}
template <int K>
cudaError_t launch_bwd(const MaskK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunk, a.nframes);
    if (vec) k_mask_bwd<K, 4><<<grid, RCF_BLOCK, 0, s>>>(a);
    else k_mask_bwd<K, 1><<<grid, RCF_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}

int mask_check(int nframes, int K, int H, int W) {
    if (nframes < 1 || nframes > 65535 || K < 1 || H < 1 || W < 1 || (long long)H * W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if (K > RCF_MAX_K) return RCF_ERR_UNSUPPORTED;
    return RCF_OK;
}
bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

void fill_cfg(MaskK& a, const RcfMaskCfg& c) {
    a.nframes = c.nframes; a.H = c.H; a.W = c.W; a.P = c.H * c.W;
    a.nchunk = (a.P + MASK_CHUNK - 1) / MASK_CHUNK;
    a.compact_ch = (c.compact_channel >= 0 && c.compact_channel < c.K) ? c.compact_channel : -1;
    a.pl_ch = (c.pl_channel >= 0 && c.pl_channel < c.K) ? c.pl_channel : -1;
    a.pl_binarize = c.pl_threshold != -1.0f;      // reference: `if self.pl_mask_pos_th != -1`
    a.pl_th = c.pl_threshold; a.pl_wpos = c.pl_pos_weight; a.pl_wneg = c.pl_neg_weight;
    a.inv_npix = (float)(1.0 / ((double)c.nframes * (double)a.P));
    a.inv_h = 1.0f / (float)c.H; a.inv_w = 1.0f / (float)c.W;
    a.K = c.K;
    a.sharpen_mode = (c.sharpen_mode == 1 || (c.sharpen_mode == 2 && c.sharpen_channel >= 0 && c.sharpen_channel < c.K)) ? c.sharpen_mode : 0;
    a.sharpen_ch = a.sharpen_mode == 2 ? c.sharpen_channel : -1;
    a.t_sharpen = c.t_sharpen; a.inv_t = c.t_sharpen > 0.0f ? 1.0f / c.t_sharpen : 1.0f;
}

#define MASK_K_SWITCH(fn, ...)                     \
    switch (K) {                                   \
        case 1: e = fn<1>(__VA_ARGS__); break;     \
        case 2: e = fn<2>(__VA_ARGS__); break;     \
        case 3: e = fn<3>(__VA_ARGS__); break;     \
        case 4: e = fn<4>(__VA_ARGS__); break;     \
        case 5: e = fn<5>(__VA_ARGS__); break;     \
        case 6: e = fn<6>(__VA_ARGS__); break;     \
        case 7: e = fn<7>(__VA_ARGS__); break;     \
        case 8: e = fn<8>(__VA_ARGS__); break;     \
    }

}  // namespace

extern "C" int rcf_mask_prep_workspace_floats(int nframes, int P, size_t* nfloats) {
    if (!nfloats) return RCF_ERR_NULL;
    if (nframes < 1 || P < 1) return RCF_ERR_SHAPE;
    // per-CTA partials (rounded up to an even count so the fp64 tail is 8-byte aligned) + per-frame totals [nframes][4] fp64
    size_t np = (size_t)nframes * ((P + MASK_CHUNK - 1) / MASK_CHUNK) * MASK_NP;
    np += np & 1;
    *nfloats = np + (size_t)nframes * 16;
    return RCF_OK;
}

extern "C" int rcf_mask_losses_forward(const RcfMaskCfg* cfg, const float* logits, const float* target, float* masks,
                                       float* losses, float* frame_stats, float* ws, void* stream) {
    if (!cfg) return RCF_ERR_NULL;
    const int K = cfg->K;
    const int v = mask_check(cfg->nframes, K, cfg->H, cfg->W);
    if (v != RCF_OK) return v;
    if (!logits || !masks || !losses || !ws) return RCF_ERR_NULL;
    MaskK a{};
    // a channel index outside [0, K) is an error (only -1 means "this loss is off"): a mistyped config must not turn into
    // a silently zero loss
    if (cfg->compact_channel < -1 || cfg->compact_channel >= cfg->K || cfg->pl_channel < -1 || cfg->pl_channel >= cfg->K ||
        (cfg->sharpen_mode == 2 && (cfg->sharpen_channel < 0 || cfg->sharpen_channel >= cfg->K)) || cfg->sharpen_mode < 0 ||
        cfg->sharpen_mode > 2)
        return RCF_ERR_SHAPE;
    fill_cfg(a, *cfg);
    if (a.compact_ch >= 0 && !frame_stats) return RCF_ERR_NULL;
    a.logits = logits; a.masks = masks; a.losses = losses; a.part = ws; a.fstats = frame_stats;
    {
        size_t np = (size_t)a.nframes * a.nchunk * MASK_NP;
        np += np & 1;
        if (reinterpret_cast<uintptr_t>(ws) & 7u) return RCF_ERR_ALIGN;
        a.ftot = reinterpret_cast<double*>(ws + np);
    }
    a.target = (a.pl_ch >= 0) ? target : nullptr;
    if (a.pl_ch >= 0 && !target) return RCF_ERR_NULL;
    const bool vec = a.P % 4 == 0 && al16(logits) && al16(masks) && (!a.target || al16(a.target));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    MASK_K_SWITCH(launch_fwd, a, vec, s)
    if (e != cudaSuccess) return (int)e;
    e = rcf_launch(k_mask_frame, a.nframes, 256, 0, s, rcf_pdl_enabled(), a);
    if (e != cudaSuccess) return (int)e;
    return (int)rcf_launch(k_mask_final, 1, 32, 0, s, rcf_pdl_enabled(), a);
}

extern "C" int rcf_mask_losses_backward(const RcfMaskCfg* cfg, const float* masks, const float* target,
                                        const float* grad_masks, const float* grad_losses, const float* frame_stats,
                                        float* dlogits, void* stream) {
    if (!cfg) return RCF_ERR_NULL;
    const int K = cfg->K;
    const int v = mask_check(cfg->nframes, K, cfg->H, cfg->W);
    if (v != RCF_OK) return v;
    if (!masks || !dlogits) return RCF_ERR_NULL;
    MaskK a{};
    fill_cfg(a, *cfg);
    a.masks = const_cast<float*>(masks); a.gmasks = grad_masks; a.glosses = grad_losses; a.dlogits = dlogits;
    a.fstats = const_cast<float*>(frame_stats);
    a.target = (a.pl_ch >= 0) ? target : nullptr;
    if (a.pl_ch >= 0 && !target) return RCF_ERR_NULL;
    if (a.compact_ch >= 0 && grad_losses && !frame_stats) return RCF_ERR_NULL;
    const bool vec = a.P % 4 == 0 && al16(masks) && al16(dlogits) && (!grad_masks || al16(grad_masks)) && (!a.target || al16(a.target));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    MASK_K_SWITCH(launch_bwd, a, vec, s)
    return (int)e;
}
