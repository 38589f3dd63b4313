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
    rcf_pdl_prologue();
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
        // The MLP weights are staged in shared memory with coalesced 128-bit loads issued BEFORE the partial sums are
        // reduced (one global round trip in total instead of one per 8 multiply-adds); rows padded to Cf+1 words.
        float* w1s = reinterpret_cast<float*>(dyn + 2 * Cf * K);          // [Cf][Cf+1]   (only when a.mlp_smem)
        float* w2s = w1s + Cf * (Cf + 1);                                   // [2][Cf]
        if (a.mlp_smem) {
            for (int q0 = tid; q0 < Cf * Cf / 4; q0 += 4 * blockDim.x) {      // 4 loads in flight per thread
                float4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * blockDim.x;
                    if (q < Cf * Cf / 4) w[u] = __ldg(reinterpret_cast<const float4*>(a.w1) + q);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * blockDim.x;
                    if (q < Cf * Cf / 4) {
                        const int i = (q * 4) / Cf, j = q * 4 - i * Cf;
                        float* d = w1s + i * (Cf + 1) + j;
                        d[0] = w[u].x; d[1] = w[u].y; d[2] = w[u].z; d[3] = w[u].w;
                    }
                }
            }
            for (int q = tid; q < 2 * Cf; q += blockDim.x) w2s[q] = __ldg(a.w2 + q);
        }
        for (int i = tid; i < Cf * K; i += blockDim.x)          // chunk partials were summed by k_pool_reduce
            pool[i] = __ldcg(a.poolsum + (size_t)fd * Cf * K + i) / stat[(i % K) * NS];
        __syncthreads();
        double* mlp = a.mlp + (size_t)fd * K * 2 * Cf;   // [k][0: pool | 1: hpre][Cf]
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int i = t / K, k = t - i * K;
            double v = (double)a.b1[i];
            if (a.mlp_smem) {
                const float* __restrict__ w = w1s + i * (Cf + 1);
#pragma unroll 8
                for (int j = 0; j < Cf; ++j) v += (double)w[j] * pool[j * K + k];
            } else {
                const float* __restrict__ w = a.w1 + (size_t)i * Cf;
#pragma unroll 8
                for (int j = 0; j < Cf; ++j) v += (double)__ldg(w + j) * pool[j * K + k];
            }
            hpre[t] = v;
            mlp[(size_t)k * 2 * Cf + Cf + i] = v;
            mlp[(size_t)k * 2 * Cf + i] = pool[t];
        }
        __syncthreads();
        // theta[c][k]: one warp per output, lanes stride over the hidden units, fp64 butterfly (fixed order)
        for (int o = tid >> 5; o < 2 * K; o += blockDim.x >> 5) {
            const int c = o / K, k = o - c * K, lane = tid & 31;
            double v = 0.0;
            for (int i = lane; i < Cf; i += 32) {
                const double h = hpre[i * K + k];
                const float w = a.mlp_smem ? w2s[c * Cf + i] : __ldg(a.w2 + (size_t)c * Cf + i);
                v += (double)w * (h >= 0.0 ? h : 0.1 * h);
            }
            v = warp_sum_d(v);
            if (lane == 0) theta_s[c * K + k] = v + (double)a.b2[c];
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
    rcf_pdl_prologue();
    const int fd = blockIdx.x;
    __shared__ double out[1 + 3 * RCF_MAX_K + 2 * RCF_MAX_K * 5];
    reduce_partials(a.part2 + (size_t)fd * GM * a.nchunk2, GM, a.nchunk2, out);
    __syncthreads();
    for (int i = threadIdx.x; i < GM; i += blockDim.x) a.gm[(size_t)fd * GM + i] = out[i];
    if (a.single_pass && threadIdx.x < a.K)      // D == 0: segd record is {S, -, -}
        a.segd[((size_t)fd * a.K + threadIdx.x) * rcf_segd(0)] = out[1 + 2 * a.K + threadIdx.x];
}

__global__ void k_loss_sum(const RcfK a, int GM) {
    rcf_pdl_prologue();
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
    rcf_pdl_prologue();
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
        float* w1s = reinterpret_cast<float*>(dyn + 2 * Cf * K);          // [Cf][Cf]   (only when a.mlp_smem)
        if (a.mlp_smem)
            for (int q0 = tid; q0 < Cf * Cf / 4; q0 += 4 * blockDim.x) {      // 4 loads in flight per thread
                float4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q0 + u * blockDim.x < Cf * Cf / 4) w[u] = __ldg(reinterpret_cast<const float4*>(a.w1) + q0 + u * blockDim.x);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q0 + u * blockDim.x < Cf * Cf / 4) reinterpret_cast<float4*>(w1s)[q0 + u * blockDim.x] = w[u];
            }
        const double* mlp = a.mlp + (size_t)fd * K * 2 * Cf;
        __shared__ unsigned int s_pbmax;
        float pbmax = 0.0f;
        if (tid == 0) s_pbmax = 0u;
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int i = t / K, k = t - i * K;
            const double hp = mlp[(size_t)k * 2 * Cf + Cf + i];
            const double v = ((double)a.w2[i] * thbar[k] + (double)a.w2[Cf + i] * thbar[K + k]) * (hp >= 0.0 ? 1.0 : 0.1);
            dh[t] = v;
            a.dh[((size_t)fd * K + k) * Cf + i] = v;
        }
        __syncthreads();
        for (int t = tid; t < Cf * K; t += blockDim.x) {
            const int k = t / Cf, j = t - k * Cf;       // consecutive threads -> consecutive j (coalesced / conflict-free w1 reads)
            double v = 0.0;
            if (a.mlp_smem) {
#pragma unroll 8
                for (int i = 0; i < Cf; ++i) v += (double)w1s[i * Cf + j] * dh[i * K + k];
            } else {
#pragma unroll 8
                for (int i = 0; i < Cf; ++i) v += (double)__ldg(a.w1 + (size_t)i * Cf + j) * dh[i * K + k];
            }
            pbar[j * K + k] = v;
            const float pb = (float)(v / sd[k * SEGD]);
            a.poolbar[((size_t)fd * Cf + j) * K + k] = pb;
            pbmax = fmaxf(pbmax, fabsf(pb));
        }
        // max |poolbar| of this frame-direction (bounds the feature-map gradient: see rcf_grad_scale); non-negative floats
        // order like their bit patterns, so one shared-memory atomicMax on the bits does it
        atomicMax(&s_pbmax, __float_as_uint(pbmax));
        __syncthreads();
        if (tid == 0) a.gmax[fd] = __uint_as_float(s_pbmax);
        for (int k = tid >> 5; k < K; k += blockDim.x >> 5) {      // one warp per segment, fp64 butterfly
            const int lane = tid & 31;
            double v = 0.0;
            for (int j = lane; j < Cf; j += 32) v += pbar[j * K + k] * mlp[(size_t)k * 2 * Cf + j];
            v = warp_sum_d(v);
            if (lane == 0) poolterm[k] = v;
        }
        __syncthreads();
    }
    if (tid < K) {
        float* cb = a.coefb + ((size_t)fd * K + tid) * CB;
        cb[CB - 1] = (float)(-(inner[tid] + poolterm[tid]) / sd[tid * SEGD]);
    }
}

// Parameter gradients of the segment MLP: tiny fp64 GEMMs over the segment axis s = fd*K + k (nseg = nfd*K terms),
//   dW1[i][j] = sum_s dh[s][i] * pool[s][j],  db1[i] = sum_s dh[s][i]                      (blocks 0 .. Cf/RB-1: RB rows each)
//   dW2[c][j] = sum_s thbar[s][c] * lrelu(hpre[s][j]),  db2[c] = sum_s thbar[s][c]         (last block)
// The operands are staged through shared memory in chunks of SC segments with coalesced loads (one global round trip
// per chunk instead of one per term) and every output is accumulated by ONE thread in segment order: deterministic.
constexpr int MPG_RB = 4, MPG_SC = 64;
__global__ void __launch_bounds__(256) k_mlp_param_grad(const RcfK a) {
    rcf_pdl_prologue();
    const int Cf = a.Cf, nseg = a.nfd * a.K, tid = threadIdx.x;
    const bool w2blk = blockIdx.x == gridDim.x - 1;
    extern __shared__ double sm[];                 // right[SC][Cf] | left[SC][RB]
    double* right = sm;
    double* left = sm + MPG_SC * Cf;
    const int i0 = blockIdx.x * MPG_RB;
    const int nrow = w2blk ? 2 : min(MPG_RB, Cf - i0);
    constexpr int MAXO = (MPG_RB * (RCF_MAX_CF + 1) + 255) / 256;     // outputs per thread (weights + bias column)
    double acc[MAXO];
#pragma unroll
    for (int u = 0; u < MAXO; ++u) acc[u] = 0.0;
    const int nout = nrow * (Cf + 1);
    for (int s0 = 0; s0 < nseg; s0 += MPG_SC) {
        const int ns = min(MPG_SC, nseg - s0);
        __syncthreads();
        double lv = 0.0;                                  // ns * nrow <= 256: one element per thread
        if (tid < ns * nrow) {
            const int s = tid / nrow, r = tid - s * nrow;
            lv = w2blk ? __ldcg(a.thbar + (size_t)(s0 + s) * 2 + r) : __ldcg(a.dh + (size_t)(s0 + s) * Cf + i0 + r);
        }
        for (int q0 = tid; q0 < ns * Cf; q0 += 8 * 256) {  // 8 loads in flight per thread
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u * 256;
                if (q < ns * Cf) {
                    const int s = q / Cf, j = q - s * Cf;
                    v[u] = __ldcg(a.mlp + (size_t)(s0 + s) * 2 * Cf + (w2blk ? Cf : 0) + j);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u * 256;
                if (q < ns * Cf) right[q] = (w2blk && v[u] < 0.0) ? 0.1 * v[u] : v[u];
            }
        }
        if (tid < ns * nrow) left[(tid / nrow) * MPG_RB + (tid - (tid / nrow) * nrow)] = lv;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAXO; ++u) {
            const int o = tid + u * 256;
            if (o < nout) {
                const int r = o / (Cf + 1), j = o - r * (Cf + 1);
                double v = acc[u];
                if (j < Cf) {
#pragma unroll 4
                    for (int s = 0; s < ns; ++s) v += left[s * MPG_RB + r] * right[s * Cf + j];
                } else {
#pragma unroll 4
                    for (int s = 0; s < ns; ++s) v += left[s * MPG_RB + r];
                }
                acc[u] = v;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < MAXO; ++u) {
        const int o = tid + u * 256;
        if (o < nout) {
            const int r = o / (Cf + 1), j = o - r * (Cf + 1);
            if (w2blk) {
                if (j < Cf) a.dw2[(size_t)r * Cf + j] = (float)acc[u]; else a.db2[r] = (float)acc[u];
            } else {
                if (j < Cf) a.dw1[(size_t)(i0 + r) * Cf + j] = (float)acc[u]; else a.db1[i0 + r] = (float)acc[u];
            }
        }
    }
}

// dynamic shared memory of the segment kernels: pool/hpre (or dh/pbar) in fp64 + the staged MLP weights
static size_t seg_dyn_bytes(const RcfK& a) {
    if (a.theta_mode != 1) return 0;
    size_t b = (size_t)2 * a.Cf * a.K * sizeof(double);
    if (a.mlp_smem) b += ((size_t)a.Cf * (a.Cf + 1) + 2 * a.Cf) * sizeof(float);
    return b;
}

template <typename Kern>
static cudaError_t seg_allow_smem(Kern kern, size_t dyn) {
    if (dyn <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
}

cudaError_t rcf_launch_segment_fwd(const RcfK& a, cudaStream_t s) {
    const size_t dyn = seg_dyn_bytes(a);
    cudaError_t e0 = cudaSuccess;
    switch (a.D) {
        case 0: e0 = seg_allow_smem(k_segment_fwd<0>, dyn); break;
        case 2: e0 = seg_allow_smem(k_segment_fwd<2>, dyn); break;
        case 5: e0 = seg_allow_smem(k_segment_fwd<5>, dyn); break;
    }
    if (e0 != cudaSuccess) return e0;
    switch (a.D) {
        case 0: rcf_launch(k_segment_fwd<0>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        case 2: rcf_launch(k_segment_fwd<2>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        case 5: rcf_launch(k_segment_fwd<5>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t rcf_launch_finalize(const RcfK& a, cudaStream_t s) {
    const int GM = rcf_gm(a.K, a.D);
    rcf_launch(k_finalize, a.nfd, RCF_BLOCK, 0, s, a.pdl, a, GM);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    rcf_launch(k_loss_sum, 1, 64, 0, s, a.pdl, a, GM);
    return cudaGetLastError();
}

cudaError_t rcf_launch_segment_bwd(const RcfK& a, cudaStream_t s) {
    const size_t dyn = seg_dyn_bytes(a);
    cudaError_t e0 = cudaSuccess;
    switch (a.D) {
        case 0: e0 = seg_allow_smem(k_segment_bwd<0>, dyn); break;
        case 2: e0 = seg_allow_smem(k_segment_bwd<2>, dyn); break;
        case 5: e0 = seg_allow_smem(k_segment_bwd<5>, dyn); break;
    }
    if (e0 != cudaSuccess) return e0;
    switch (a.D) {
        case 0: rcf_launch(k_segment_bwd<0>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        case 2: rcf_launch(k_segment_bwd<2>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        case 5: rcf_launch(k_segment_bwd<5>, a.nfd, RCF_BLOCK, dyn, s, a.pdl, a); break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (a.theta_mode == 1 && a.dw1 && a.db1 && a.dw2 && a.db2) {
        const size_t sm = (size_t)MPG_SC * (a.Cf + MPG_RB) * sizeof(double);
        e = seg_allow_smem(k_mlp_param_grad, sm);
        if (e != cudaSuccess) return e;
        rcf_launch(k_mlp_param_grad, (a.Cf + MPG_RB - 1) / MPG_RB + 1, 256, sm, s, a.pdl, a);
        e = cudaGetLastError();
    }
    return e;
}
