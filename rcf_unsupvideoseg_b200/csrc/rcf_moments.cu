// rcf_moments.cu -- pass 1: per-(frame-direction, segment) weighted moment sums.
//
// Replaces, in one coalesced read of the mask (and of the flow when an affine fit is on):
//   * mask.sum over pixels                                   reference :242-243, :168
//   * mu_F, mu_omega, sigma_F_omega, sigma_omega_omega sums   reference :173-209
// (the reference materialises [B,K,HW,2,2] einsum temporaries; here only K*NS scalars per CTA leave
// the SM).  Raw moments are taken in a centred, [-1,1]-scaled coordinate basis; the segment kernel
// converts them to the de-meaned covariances in fp64.
//
// HBM-bound: algorithmic bytes per pixel = 4K (+8 with the affine fit).
#include "rcf_moments_dev.cuh"

template <int K, int D, int PX>
__global__ void __launch_bounds__(RCF_BLOCK, (D == 2) ? 2 : 1) k_moments(const RcfK a) {
    rcf_pdl_prologue();
    __shared__ float red[RCF_WARPS][K * rcf_ns(D)];
    moments_tile<K, D, PX>(a, blockIdx.y, blockIdx.x, red);
}

template <int K, int D>
static cudaError_t launch_kd(const RcfK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunk1, a.nfd), block(RCF_BLOCK);
    if (vec) rcf_launch(k_moments<K, D, 4>, grid, block, 0, s, a.pdl, a);
    else rcf_launch(k_moments<K, D, 1>, grid, block, 0, s, a.pdl, a);
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

cudaError_t rcf_launch_moments(const RcfK& a, bool vec, cudaStream_t s) {
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
