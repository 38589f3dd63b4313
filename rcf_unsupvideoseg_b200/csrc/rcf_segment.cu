// rcf_segment.cu -- the tiny per-(frame-direction, segment) kernels, all in fp64 registers/smem.
//
//   k_segment_fwd : reduce pass-1 partials (fixed order => bit-reproducible), de-mean the moments,
//                   solve the D x D normal equations (reference :198-217, torch.linalg.solve),
//                   pooled features / S (:251-256) and the segment MLP flow_feat_after_agg (:95-101, :258),
//                   emit the fp32 coefficient pack pass 2 consumes.
//   k_finalize    : reduce pass-2 partials -> per-fd loss sums and gradient moments.
//   k_loss_sum    : loss[dir] = inv_n * sum_b (mean of :361-368).
//   k_segment_bwd : per-segment backward incl. MLP backward; emits the fp32
//                   coefficient pack of the streaming backward and d_theta.
//   k_mlp_param_grad : dW1, db1, dW2, db2 summed over all segments in a fixed order.
// One CTA per frame-direction; cost is O(K * (NS*nchunk + Cf^2)) -- microseconds.
#include "rcf_segment_dev.cuh"


template <int D>
__global__ void __launch_bounds__(RCF_BLOCK) k_segment_fwd(const RcfK a) {
    constexpr int NS = rcf_ns(D), SEGD = rcf_segd(D), CF = rcf_cf(D);
    __shared__ double stat[RCF_SEG_MAXSTAT];
    __shared__ double theta_s[2 * RCF_MAX_K];
    extern __shared__ double dyn[];            // pool[Cf*K], h[Cf*K]   (theta_mode 1)
    const int fd = blockIdx.x, K = a.K, Cf = a.Cf, tid = threadIdx.x;
    const int dir = fd / a.B, b = fd - dir * a.B;

    if (a.single_pass) {          // D == 0, theta supplied: S_k is produced by pass 2 (k_finalize overwrites segd[0])
        if (tid < K * NS) stat[tid] = 1.0;
    } else {
        reduce_partials(a.part1 + (size_t)fd * K * NS * a.nchunk1, K * NS, a.nchunk1, stat);
    }
    __syncthreads();
    double* sd = a.segd + (size_t)fd * K * SEGD;
    if (tid < K) seg_affine_fwd<D>(stat + tid * NS, sd + tid * SEGD);

    if (a.theta_mode == 1) {
        double* pool = dyn;              // [f*K + k]
        double* hpre = dyn + Cf * K;     // [i*K + k]
        for (int i = tid; i < Cf * K; i += blockDim.x)          // chunk partials were summed by k_pool_reduce
            pool[i] = __ldcg(a.poolsum + (size_t)fd * Cf * K + i) / stat[(i % K) * NS];
        __syncthreads();
        double* mlp = a.mlp + (size_t)fd * K * 2 * Cf;   // [k][0: pool | 1: hpre][Cf]
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int i = t / K, k = t - i * K;
            double v = (double)a.b1[i];
            const float* __restrict__ w = a.w1 + (size_t)i * Cf;
#pragma unroll 8
            for (int j = 0; j < Cf; ++j) v += (double)__ldg(w + j) * pool[j * K + k];
            hpre[t] = v;
            mlp[(size_t)k * 2 * Cf + Cf + i] = v;
            mlp[(size_t)k * 2 * Cf + i] = pool[t];
        }
        __syncthreads();
        if (tid < 2 * K) {
            const int c = tid / K, k = tid - c * K;
            double v = (double)a.b2[c];
            const float* __restrict__ w = a.w2 + (size_t)c * Cf;
#pragma unroll 8
            for (int i = 0; i < Cf; ++i) {
                const double h = hpre[i * K + k];
                v += (double)__ldg(w + i) * (h >= 0.0 ? h : 0.1 * h);
            }
            theta_s[c * K + k] = v;
        }
    } else {
        if (tid < 2 * K) theta_s[tid] = (double)a.theta[dir][(size_t)b * 2 * K + tid];   // [B,2,K]
    }
    __syncthreads();
    if (tid < K) {
        float* cf = a.coef + ((size_t)fd * K + tid) * CF;
        const double* s = sd + tid * SEGD;
        cf[0] = (float)theta_s[tid];
        cf[1] = (float)theta_s[K + tid];
        if constexpr (D > 0) {
            const double* A = s + 3 + 3 * D + 2 * D * D;
#pragma unroll
            for (int i = 0; i < 2 * D; ++i) cf[2 + i] = (float)A[i];
#pragma unroll
            for (int d = 0; d < D; ++d) cf[2 + 2 * D + d] = (float)s[1 + d];
        }
    }
}

__global__ void __launch_bounds__(RCF_BLOCK) k_finalize(const RcfK a, int GM) {
    const int fd = blockIdx.x;
    __shared__ double out[1 + 3 * RCF_MAX_K + 2 * RCF_MAX_K * 5];
    reduce_partials(a.part2 + (size_t)fd * GM * a.nchunk2, GM, a.nchunk2, out);
    __syncthreads();
    for (int i = threadIdx.x; i < GM; i += blockDim.x) a.gm[(size_t)fd * GM + i] = out[i];
    if (a.single_pass && threadIdx.x < a.K)      // D == 0: segd record is {S, -, -}
        a.segd[((size_t)fd * a.K + threadIdx.x) * rcf_segd(0)] = out[1 + 2 * a.K + threadIdx.x];
}

__global__ void k_loss_sum(const RcfK a, int GM) {
    // one warp per direction; lanes stride over the batch, fp64 butterfly; loss[ndir] = fp32 sum of the directions (:397)
    const int dir = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float ldir[2];
    if (dir < a.ndir) {
        double v = 0.0;
        for (int b = lane; b < a.B; b += 32) v += a.gm[(size_t)(dir * a.B + b) * GM];
        v = warp_sum_d(v);
        if (lane == 0) { ldir[dir] = (float)(v * (double)a.inv_n); a.loss[dir] = ldir[dir]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.loss[a.ndir] = a.ndir == 2 ? ldir[0] + ldir[1] : ldir[0];
}

template <int D>
__global__ void __launch_bounds__(RCF_BLOCK) k_segment_bwd(const RcfK a) {
    constexpr int SEGD = rcf_segd(D), CB = rcf_cb(D);
    __shared__ double thbar[2 * RCF_MAX_K];      // [c*K + k]
    __shared__ double inner[RCF_MAX_K];          // <SFubar,SFu> + <Suubar,Suu>
    __shared__ double poolterm[RCF_MAX_K];
    extern __shared__ double dyn[];              // dh[Cf*K], pbar[Cf*K]
    const int fd = blockIdx.x, K = a.K, Cf = a.Cf, tid = threadIdx.x;
    const int dir = fd / a.B, b = fd - dir * a.B;
    const int GM = rcf_gm(K, D);
    const double gs = -(double)a.grad_loss[a.grad_total ? 0 : dir] * (double)a.inv_n;
    const double* gm = a.gm + (size_t)fd * GM;
    const double* sd = a.segd + (size_t)fd * K * SEGD;
    if (tid == 0) a.gscale[fd] = (float)gs;

    if (tid < K) {
        const int k = tid;
        const double* s = sd + k * SEGD;
        const double S = s[0];
        const double tb0 = gs * gm[1 + k], tb1 = gs * gm[1 + K + k];
        thbar[k] = tb0; thbar[K + k] = tb1;
        float* cb = a.coefb + ((size_t)fd * K + k) * CB;
        double in = 0.0;
        if constexpr (D > 0) {
            const double* muF = s + 1 + D;
            const double* SFu = s + 3 + D;
            const double* Suu = s + 3 + 3 * D;
            const double* Sinv = s + 3 + 3 * D + D * D;
            const double* A = s + 3 + 3 * D + 2 * D * D;
            double Ab[2][D], SFb[2][D], Sub[D][D], mub[D];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int d = 0; d < D; ++d) Ab[c][d] = gs * gm[1 + 2 * K + (k * 2 + c) * D + d];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    double v = 0.0;
#pragma unroll
                    for (int e = 0; e < D; ++e) v += Ab[c][e] * Sinv[e * D + d];
                    SFb[c][d] = v;
                    in += v * SFu[c * D + d];
                }
#pragma unroll
            for (int d = 0; d < D; ++d) {
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    const double v = -(A[d] * SFb[0][e] + A[D + d] * SFb[1][e]);
                    Sub[d][e] = v;
                    in += v * Suu[d * D + e];
                }
                mub[d] = -(A[d] * tb0 + A[D + d] * tb1);
            }
            const double iS = 1.0 / S;
            cb[0] = (float)muF[0]; cb[1] = (float)muF[1];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int d = 0; d < D; ++d) cb[2 + c * D + d] = (float)(SFb[c][d] * iS);
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int e = d; e < D; ++e)
                    cb[2 + 2 * D + rcf_sym_idx(D, d, e)] = (float)((d == e ? Sub[d][d] : Sub[d][e] + Sub[e][d]) * iS);
#pragma unroll
            for (int d = 0; d < D; ++d) cb[2 + 2 * D + D * (D + 1) / 2 + d] = (float)(mub[d] * iS);
        } else {
            cb[0] = 0.0f; cb[1] = 0.0f;
        }
        inner[k] = in;
        poolterm[k] = 0.0;
        if (a.theta_mode == 0 && a.dtheta[dir]) {
            float* dt = a.dtheta[dir] + (size_t)b * 2 * K;
            dt[k] = (float)tb0; dt[K + k] = (float)tb1;
        }
        a.thbar[((size_t)fd * K + k) * 2 + 0] = tb0;
        a.thbar[((size_t)fd * K + k) * 2 + 1] = tb1;
    }
    __syncthreads();

    if (a.theta_mode == 1) {
        double* dh = dyn;             // [i*K + k]
        double* pbar = dyn + Cf * K;  // [j*K + k]
        const double* mlp = a.mlp + (size_t)fd * K * 2 * Cf;
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int i = t / K, k = t - i * K;
            const double hp = mlp[(size_t)k * 2 * Cf + Cf + i];
            const double v = ((double)a.w2[i] * thbar[k] + (double)a.w2[Cf + i] * thbar[K + k]) * (hp >= 0.0 ? 1.0 : 0.1);
            dh[t] = v;
            a.dh[((size_t)fd * K + k) * Cf + i] = v;
        }
        __syncthreads();
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int k = t / Cf, j = t - k * Cf;       // consecutive threads -> consecutive j (coalesced w1 reads)
            double v = 0.0;
#pragma unroll 8
            for (int i = 0; i < Cf; ++i) v += (double)__ldg(a.w1 + (size_t)i * Cf + j) * dh[i * K + k];
            pbar[j * K + k] = v;
            a.poolbar[((size_t)fd * Cf + j) * K + k] = (float)(v / sd[k * SEGD]);
        }
        __syncthreads();
        if (tid < K) {
            double v = 0.0;
            for (int j = 0; j < Cf; ++j) v += pbar[j * K + tid] * mlp[(size_t)tid * 2 * Cf + j];
            poolterm[tid] = v;
        }
        __syncthreads();
    }
    if (tid < K) {
        float* cb = a.coefb + ((size_t)fd * K + tid) * CB;
        cb[CB - 1] = (float)(-(inner[tid] + poolterm[tid]) / sd[tid * SEGD]);
    }
}

// grid = Cf + 1 blocks of 256 threads.  Block i < Cf: row i of dW1 and db1[i].  Block Cf: dW2, db2.
// The segment sum is split over G = 256/64 thread groups (fixed partition) and combined through shared memory
// in group order: deterministic, and 4x shorter dependent chains than one thread per output.
__global__ void __launch_bounds__(256) k_mlp_param_grad(const RcfK a) {
    const int Cf = a.Cf, K = a.K, nseg = a.nfd * K;
    const int i = blockIdx.x;
    const double* __restrict__ dh = a.dh;
    const double* __restrict__ mlp = a.mlp;
    const double* __restrict__ thbar = a.thbar;
    __shared__ double part[256];
    const int nout = (i < Cf) ? Cf + 1 : 2 * Cf + 2;        // outputs of this block (weights + bias)
    int groups = 256 / ((nout + 31) / 32 * 32);
    if (groups < 1) groups = 1;
    const int lanes = 256 / groups;                          // threads per group (>= nout when groups > 1)
    const int grp = threadIdx.x / lanes, t0 = threadIdx.x - grp * lanes;
    const int s_lo = (int)((long long)nseg * grp / groups), s_hi = (int)((long long)nseg * (grp + 1) / groups);
    for (int o0 = 0; o0 < nout; o0 += lanes) {
        const int o = o0 + t0;
        double v = 0.0;
        if (o < nout && grp < groups) {
            if (i < Cf) {
                if (o < Cf) {
#pragma unroll 8
                    for (int s = s_lo; s < s_hi; ++s) v += __ldg(dh + (size_t)s * Cf + i) * __ldg(mlp + (size_t)s * 2 * Cf + o);
                } else {
#pragma unroll 8
                    for (int s = s_lo; s < s_hi; ++s) v += __ldg(dh + (size_t)s * Cf + i);
                }
            } else if (o < 2 * Cf) {
                const int c = o / Cf, j = o - c * Cf;
#pragma unroll 8
                for (int s = s_lo; s < s_hi; ++s) {
                    const double hp = __ldg(mlp + (size_t)s * 2 * Cf + Cf + j);
                    v += __ldg(thbar + (size_t)s * 2 + c) * (hp >= 0.0 ? hp : 0.1 * hp);
                }
            } else {
                const int c = o - 2 * Cf;
#pragma unroll 8
                for (int s = s_lo; s < s_hi; ++s) v += __ldg(thbar + (size_t)s * 2 + c);
            }
        }
        part[threadIdx.x] = v;
        __syncthreads();
        if (grp == 0 && o < nout) {
            double tot = 0.0;
            for (int g = 0; g < groups; ++g) tot += part[g * lanes + t0];
            if (i < Cf) {
                if (o < Cf) a.dw1[(size_t)i * Cf + o] = (float)tot;
                else a.db1[i] = (float)tot;
            } else if (o < 2 * Cf) {
                a.dw2[o] = (float)tot;
            } else {
                a.db2[o - 2 * Cf] = (float)tot;
            }
        }
        __syncthreads();
    }
}

cudaError_t rcf_launch_segment_fwd(const RcfK& a, cudaStream_t s) {
    const size_t dyn = a.theta_mode == 1 ? (size_t)2 * a.Cf * a.K * sizeof(double) : 0;
    switch (a.D) {
        case 0: k_segment_fwd<0><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        case 2: k_segment_fwd<2><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        case 5: k_segment_fwd<5><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t rcf_launch_finalize(const RcfK& a, cudaStream_t s) {
    const int GM = rcf_gm(a.K, a.D);
    k_finalize<<<a.nfd, RCF_BLOCK, 0, s>>>(a, GM);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_loss_sum<<<1, 64, 0, s>>>(a, GM);
    return cudaGetLastError();
}

cudaError_t rcf_launch_segment_bwd(const RcfK& a, cudaStream_t s) {
    const size_t dyn = a.theta_mode == 1 ? (size_t)2 * a.Cf * a.K * sizeof(double) : 0;
    switch (a.D) {
        case 0: k_segment_bwd<0><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        case 2: k_segment_bwd<2><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        case 5: k_segment_bwd<5><<<a.nfd, RCF_BLOCK, dyn, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (a.theta_mode == 1 && a.dw1 && a.db1 && a.dw2 && a.db2) {
        k_mlp_param_grad<<<a.Cf + 1, 256, 0, s>>>(a);
        e = cudaGetLastError();
    }
    return e;
}
