// rcf_maskops.cu -- caller-side mask preparation fused into one pass each way.
//
// Reference (models/rcf_model.py): the segmentation logits [B, I, K, H, W] become the soft masks the motion loss
// consumes, and the same tensor feeds the entropy regulariser of stage 1 (configs/rcf/rcf_stage1.yaml:67, w_entropy 0.05):
//   :433      all_pred_mask     = F.softmax(all_pred_mask, dim=2)
//   :434      log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)       (sic: log-softmax OF the probabilities)
//   :376-378  entropy           = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
// ATen runs this as softmax, log_softmax, mul, sum, mean (+ their five backward kernels, + the add that merges the two
// gradient streams into the softmax backward).  Here:
//   k_mask_fwd : one read of the logits -> masks written once + per-CTA entropy partials (fixed-order reduction)
//   k_mask_bwd : one read of (masks, dL/dmasks) -> dL/dlogits, with the entropy gradient folded in:
//        ls = log_softmax(m), q = exp(ls), Sm = sum_k m_k
//        dE/dm_j    = -(ls_j + m_j - q_j * Sm) / Npix
//        gm_j       = gmask_j + gE * dE/dm_j
//        dlogit_j   = m_j * (gm_j - sum_k m_k gm_k)
// HBM-bound: 8K B/px forward, 12K B/px backward (fp32), 128-bit accesses, no intermediates.
#include "rcf_common.cuh"

namespace {

struct MaskK {
    const float* logits;
    float* masks;
    float* part;         // [nframes * nchunk]
    float* entropy;      // [1]
    const float* gmasks; // may be null
    const float* gent;   // device scalar, may be null
    float* dlogits;
    int nframes, P, nchunk;
    float inv_npix;      // 1 / (nframes * P)
};

constexpr int MASK_CHUNK = 2048;

template <int K, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_mask_fwd(const MaskK a) {
    asm volatile("griddepcontrol.launch_dependents;");      // k_mask_entropy_final may be scheduled while this grid drains
    constexpr int ITER = MASK_CHUNK / (RCF_BLOCK * PX);
    __shared__ float red[RCF_WARPS];
    const int fr = blockIdx.y, tid = threadIdx.x;
    const float* __restrict__ lg = a.logits + (size_t)fr * K * a.P;
    float* __restrict__ mk = a.masks + (size_t)fr * K * a.P;
    float ent = 0.0f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = blockIdx.x * MASK_CHUNK + (it * RCF_BLOCK + tid) * PX;
        if (p < a.P) {
            float x[K][PX];
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::ld(x[k], lg + (size_t)k * a.P + p);
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float mx = x[0][j];
#pragma unroll
                for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k][j]);
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) { x[k][j] = __expf(x[k][j] - mx); s += x[k][j]; }
                const float inv = 1.0f / s;
                float s2 = 0.0f, dot = 0.0f, sm = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float m = x[k][j] * inv;
                    x[k][j] = m;
                    s2 += __expf(m);              // m in [0, 1]: no max subtraction needed
                    dot = fmaf(m, m, dot);
                    sm += m;
                }
                // -(sum_k m_k (m_k - lse)) = lse * sum m - sum m^2
                ent += __logf(s2) * sm - dot;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) Pack<PX>::st(mk + (size_t)k * a.P + p, x[k]);
        }
    }
    ent = warp_sum(ent);
    if ((tid & 31) == 0) red[tid >> 5] = ent;
    __syncthreads();
    if (tid == 0) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < RCF_WARPS; ++w) v += red[w];
        a.part[(size_t)fr * a.nchunk + blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) k_mask_entropy_final(const MaskK a) {
    rcf_pdl_prologue();
    // one CTA: fp64 sum of the per-CTA partials in a fixed order
    __shared__ double red[8];
    const int n = a.nframes * a.nchunk, tid = threadIdx.x;
    double v = 0.0;
    for (int i = tid; i < n; i += 256) v += (double)__ldcg(a.part + i);
    v = warp_sum_d(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        a.entropy[0] = (float)(t * (double)a.inv_npix);
    }
}

template <int K, int PX>
__global__ void __launch_bounds__(RCF_BLOCK) k_mask_bwd(const MaskK a) {
    constexpr int ITER = MASK_CHUNK / (RCF_BLOCK * PX);
    const int fr = blockIdx.y, tid = threadIdx.x;
    const float* __restrict__ mk = a.masks + (size_t)fr * K * a.P;
    const float* __restrict__ gm = a.gmasks ? a.gmasks + (size_t)fr * K * a.P : nullptr;
    float* __restrict__ dl = a.dlogits + (size_t)fr * K * a.P;
    const float ge = a.gent ? -__ldg(a.gent) * a.inv_npix : 0.0f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int p = blockIdx.x * MASK_CHUNK + (it * RCF_BLOCK + tid) * PX;
        if (p < a.P) {
            float m[K][PX], g[K][PX];
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
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float s2 = 0.0f, sm = 0.0f;
                float e[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { e[k] = __expf(m[k][j]); s2 += e[k]; sm += m[k][j]; }
                const float lse = __logf(s2), inv2 = 1.0f / s2;
                float dot = 0.0f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float mkj = m[k][j];
                    const float dE = (mkj - lse) + mkj - e[k] * inv2 * sm;     // ls_j + m_j - q_j * Sm
                    g[k][j] = fmaf(ge, dE, g[k][j]);
                    dot = fmaf(mkj, g[k][j], dot);
                }
#pragma unroll
                for (int k = 0; k < K; ++k) g[k][j] = m[k][j] * (g[k][j] - dot);
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

int mask_check(int nframes, int K, long long P) {
    if (nframes < 1 || nframes > 65535 || K < 1 || P < 1 || P > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if (K > RCF_MAX_K) return RCF_ERR_UNSUPPORTED;
    return RCF_OK;
}
bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
    *nfloats = (size_t)nframes * ((P + MASK_CHUNK - 1) / MASK_CHUNK);
    return RCF_OK;
}

extern "C" int rcf_mask_prep_forward(const float* logits, float* masks, float* entropy, float* ws, int nframes, int K, int P,
                                     void* stream) {
    const int v = mask_check(nframes, K, P);
    if (v != RCF_OK) return v;
    if (!logits || !masks || !entropy || !ws) return RCF_ERR_NULL;
    MaskK a{};
    a.logits = logits; a.masks = masks; a.entropy = entropy; a.part = ws;
    a.nframes = nframes; a.P = P; a.nchunk = (P + MASK_CHUNK - 1) / MASK_CHUNK;
    a.inv_npix = (float)(1.0 / ((double)nframes * (double)P));
    const bool vec = P % 4 == 0 && al16(logits) && al16(masks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    MASK_K_SWITCH(launch_fwd, a, vec, s)
    if (e != cudaSuccess) return (int)e;
    return (int)rcf_launch(k_mask_entropy_final, 1, 256, 0, s, rcf_pdl_enabled(), a);
}

extern "C" int rcf_mask_prep_backward(const float* masks, const float* grad_masks, const float* grad_entropy, float* dlogits,
                                      int nframes, int K, int P, void* stream) {
    const int v = mask_check(nframes, K, P);
    if (v != RCF_OK) return v;
    if (!masks || !dlogits) return RCF_ERR_NULL;
    MaskK a{};
    a.masks = const_cast<float*>(masks); a.gmasks = grad_masks; a.gent = grad_entropy; a.dlogits = dlogits;
    a.nframes = nframes; a.P = P; a.nchunk = (P + MASK_CHUNK - 1) / MASK_CHUNK;
    a.inv_npix = (float)(1.0 / ((double)nframes * (double)P));
    const bool vec = P % 4 == 0 && al16(masks) && al16(dlogits) && (!grad_masks || al16(grad_masks));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaErrorInvalidValue;
    MASK_K_SWITCH(launch_bwd, a, vec, s)
    return (int)e;
}
