// rcf_segment_dev.cuh -- per-segment device routines of rcf_segment.cu.
#pragma once
#include "rcf_common.cuh"

#define RCF_SEG_MAXSTAT (RCF_MAX_K * 33)

// Sum `nchunk` fp32 partials of each of `nstat` statistics ([stat][chunk] layout) in fp64.  LPS lanes cooperate on
// one statistic (LPS = smallest power of two >= nchunk, capped at 32) so that small frames (few chunks) reduce
// 32/LPS statistics per warp at once instead of paying one global-load latency per statistic.  The summation order
// depends on nchunk only => bit-reproducible.  Loads bypass L1 (partials were written by other SMs).
__device__ inline void reduce_partials(const float* __restrict__ part, int nstat, int nchunk, double* out) {
    int lps = 1;
    while (lps < nchunk && lps < 32) lps <<= 1;
    const int tid = threadIdx.x;
    const int sub = tid & (lps - 1);
    const int per_pass = blockDim.x / lps;
    constexpr int U = 4;                               // statistics in flight per thread (independent load chains)
    for (int s0 = 0; s0 < nstat; s0 += U * per_pass) { // uniform trip count: shuffles stay convergent
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s0 + u * per_pass + tid / lps;
            v[u] = 0.0;
            if (s < nstat) {
                const float* p = part + (size_t)s * nchunk;
                for (int c = sub; c < nchunk; c += lps) v[u] += (double)__ldcg(p + c);
            }
        }
        for (int o = lps >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s0 + u * per_pass + tid / lps;
            if (sub == 0 && s < nstat) out[s] = v[u];
        }
    }
}

// inverse of a symmetric positive definite D x D matrix via Cholesky (row-major in/out)
template <int D>
__device__ inline void spd_inverse(const double* a, double* inv) {
    double L[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) L[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = a[j * D + j];
#pragma unroll
        for (int t = 0; t < j; ++t) s -= L[j][t] * L[j][t];
        const double dj = sqrt(s);
        L[j][j] = dj;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double v = a[i * D + j];
#pragma unroll
            for (int t = 0; t < j; ++t) v -= L[i][t] * L[j][t];
            L[i][j] = v / dj;
        }
    }
    // Linv (lower) by forward substitution, then inv = Linv^T Linv
    double Li[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) Li[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        Li[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double v = 0.0;
#pragma unroll
            for (int t = j; t < i; ++t) v -= L[i][t] * Li[t][j];
            Li[i][j] = v / L[i][i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < D; ++t) v += Li[t][i] * Li[t][j];
            inv[i * D + j] = v;
        }
}

// segd record: [0] S, [1..1+D) mu_u, [1+D..3+D) mu_F, SFu[2D], Suu[D*D], Sinv[D*D], A[2D]
template <int D>
__device__ inline void seg_affine_fwd(const double* st, double* sd) {
    const double S = st[0];
    sd[0] = S;
    if constexpr (D > 0) {
        const double inv = 1.0 / S;
        double mu[D], muF[2], SFu[2][D], Suu[D * D], Sinv[D * D];
        muF[0] = st[1] * inv; muF[1] = st[2] * inv;
#pragma unroll
        for (int d = 0; d < D; ++d) mu[d] = st[3 + d] * inv;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < D; ++d) SFu[c][d] = st[3 + D + c * D + d] * inv - muF[c] * mu[d];
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int e = d; e < D; ++e) {
                const double v = st[3 + 3 * D + rcf_sym_idx(D, d, e)] * inv - mu[d] * mu[e];
                Suu[d * D + e] = v; Suu[e * D + d] = v;
            }
        spd_inverse<D>(Suu, Sinv);
        double* o = sd + 1;
#pragma unroll
        for (int d = 0; d < D; ++d) *o++ = mu[d];
        *o++ = muF[0]; *o++ = muF[1];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < D; ++d) *o++ = SFu[c][d];
#pragma unroll
        for (int i = 0; i < D * D; ++i) *o++ = Suu[i];
#pragma unroll
        for (int i = 0; i < D * D; ++i) *o++ = Sinv[i];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double v = 0.0;
#pragma unroll
                for (int e = 0; e < D; ++e) v += SFu[c][e] * Sinv[e * D + d];
                *o++ = v;   // A[c][d]
            }
    }
}

