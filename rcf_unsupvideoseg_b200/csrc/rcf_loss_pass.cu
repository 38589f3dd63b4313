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
// backward never has to reduce anything.  VIS instantiations also write the visualisation flows [:370-395].
//
// Loop order is segment-outer / pixel-inner so that one segment's coefficients are live at a time
// (register pressure decides occupancy here: see DESIGN.md).
// HBM-bound: algorithmic bytes per pixel = 4K + 8 + 8K read (+ up to 40 written when VIS).
#include "rcf_loss_dev.cuh"

template <int K, int D, int PX, bool VIS>
__global__ void __launch_bounds__(RCF_BLOCK, (D <= 2) ? 2 : 1) k_loss(const RcfK a) {
    rcf_pdl_prologue();
    __shared__ float cf[K * rcf_cf(D)];
    __shared__ float red[RCF_WARPS][rcf_gm(K, D)];
    // Traverse in the REVERSE order of pass 1 (k_moments runs fd 0..n-1, chunks ascending): the masks read last by
    // pass 1 are still in the 126 MB L2 and are consumed first.  k_bwd then runs in ascending order again, i.e. the
    // reverse of this kernel, for the same reason.
    loss_tile<K, D, PX, VIS>(a, a.nfd - 1 - (int)blockIdx.y, a.nchunk2 - 1 - (int)blockIdx.x, cf, red);
}

template <int K, int D>
static cudaError_t launch_kd(const RcfK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunk2, a.nfd), block(RCF_BLOCK);
    const bool vis = a.vis_gt || a.vis_pred || a.vis_agg || a.vis_res || a.vis_aff;
    constexpr int VPX = K <= 4 ? 4 : 2;     // pixels per thread on the vector path (register budget)
    if (vis) {
        if (vec) rcf_launch(k_loss<K, D, VPX, true>, grid, block, 0, s, a.pdl, a);
        else rcf_launch(k_loss<K, D, 1, true>, grid, block, 0, s, a.pdl, a);
    } else {
        if (vec) rcf_launch(k_loss<K, D, VPX, false>, grid, block, 0, s, a.pdl, a);
        else rcf_launch(k_loss<K, D, 1, false>, grid, block, 0, s, a.pdl, a);
    }
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
