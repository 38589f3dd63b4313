// rcf_backward_pass.cu -- the single streaming backward pass: recompute, then write dM and dR.
//
// Nothing per-pixel is saved from the forward (the reference's autograd keeps ~20 full-resolution
// tensors alive between :242 and :368).  Per pixel this kernel re-evaluates tanh / pred / phi' and
// emits (gs = -gbar/N):
//   g_c      = gs * w_c
//   dR_ck    = g_c * (s/div) * (1 - T_ck^2) * m_k                      (or g_c * m_k when unbounded)
//   dM_k     = sum_c g_c (theta_ck + A_kc.v_k + s T_ck)
//              + f_k^T B_k v_k + v_k^T C_k v_k + mubar_k.v_k + c0_k    (B, C, mubar, c0 pre-divided by S_k)
//              [the pooled-feature term is added afterwards by k_pool_bwd]
// HBM-bound: algorithmic bytes per pixel = (4K + 8 + 8K) read + (4K + 8K) written.
#include "rcf_common.cuh"

template <int K, int D, int PX>
__global__ void __launch_bounds__(RCF_BLOCK, (D <= 2) ? 2 : 1) k_bwd(const RcfK a) {
    rcf_pdl_prologue();
    constexpr int CF = rcf_cf(D);
    constexpr int CB = rcf_cb(D);
    constexpr int ITER = RCF_CHUNK_BWD / (RCF_BLOCK * PX);
    constexpr int DD = D > 0 ? D : 1;
    __shared__ float cf[K * CF];
    __shared__ float cb[K * CB];

    const int fd = blockIdx.y;
    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int chunk = blockIdx.x;
    const int tid = threadIdx.x;
    const int P = a.P;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ flow = a.flow[dir] + (long long)b * a.flow_bs[dir];
    const float* __restrict__ resid = a.resid[dir] + (long long)b * a.resid_bs[dir];
    float* __restrict__ dmask = a.dmask[dir] ? a.dmask[dir] + (long long)b * a.dmask_bs[dir] : nullptr;
    float* __restrict__ dresid = a.dresid[dir] ? a.dresid[dir] + (long long)b * a.dresid_bs[dir] : nullptr;

    if (D == 0 && a.single_pass) {      // theta supplied, single-pass forward: no k_segment_fwd ran, coef = theta
        for (int i = tid; i < K * CF; i += RCF_BLOCK) cf[i] = __ldg(a.theta[dir] + (size_t)b * 2 * K + (i & 1) * K + (i >> 1));
    } else {
        for (int i = tid; i < K * CF; i += RCF_BLOCK) cf[i] = a.coef[(size_t)fd * K * CF + i];
    }
    for (int i = tid; i < K * CB; i += RCF_BLOCK) cb[i] = a.coefb[(size_t)fd * K * CB + i];
    __syncthreads();
    const float gs = a.gscale[fd];

    const int p0 = chunk * RCF_CHUNK_BWD;
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

            if constexpr (D == 2 && PX % 2 == 0) {
                // Affine fit: every quantity is per pixel, so the whole body runs on PIXEL pairs (the halves of a
                // 128-bit load are register neighbours) with FFMA2 / FMUL2 / FADD2 and broadcast coefficients: the
                // scalar form is 130 M warp instructions (issue active 54 %) for a kernel that should wait on HBM only.
                // Same operations in the same order per pixel as the scalar code below.
                constexpr int PP = PX / 2;
                const f32x2 one2 = pack2(1.0f, 1.0f), mtwo2 = pack2(-2.0f, -2.0f), sc2 = pack2(a.scale, a.scale);
                const f32x2 es2 = pack2(a.ex2_scale, a.ex2_scale), ds2 = pack2(a.dres_scale, a.dres_scale);
                f32x2 up[2][PP], gp[2][PP], tp[2][K][PP], fp[2][PP];
#pragma unroll
                for (int h = 0; h < PP; ++h) {
                    up[0][h] = pack2(y[2 * h], y[2 * h + 1]);
                    up[1][h] = pack2(x[2 * h], x[2 * h + 1]);
                    gp[0][h] = gp[1][h] = 0ull;
                }
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
                                float e0, e1, q0, q1;
                                unpack2(mul2(t, es2), e0, e1);
                                unpack2(add2(pack2(fast_ex2(e0), fast_ex2(e1)), one2), q0, q1);
                                t = fma2(mtwo2, pack2(fast_rcp(q0), fast_rcp(q1)), one2);
                            }
                            tp[c][k][h] = t;
                            f32x2 q = fma2(sc2, t, pack2(ck[c], ck[c]));
                            q = fma2(pack2(ck[2 + c * 2], ck[2 + c * 2]), v0, q);
                            q = fma2(pack2(ck[3 + c * 2], ck[3 + c * 2]), v1, q);
                            gp[c][h] = fma2(mp, q, gp[c][h]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int h = 0; h < PP; ++h) {
                        float pr[2], fc[2], gw[2];
                        unpack2(gp[c][h], pr[0], pr[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            fc[e] = clamp_flow(f[c][2 * h + e], a.clamp_t);
                            float phi, w;
                            loss_terms(fc[e] - pr[e], a, phi, w);
                            gw[e] = gs * w;
                        }
                        fp[c][h] = pack2(fc[0], fc[1]);
                        gp[c][h] = pack2(gw[0], gw[1]);
                    }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float ck[CF], bk[CB];
#pragma unroll
                    for (int i = 0; i < CF; ++i) ck[i] = cf[k * CF + i];
#pragma unroll
                    for (int i = 0; i < CB; ++i) bk[i] = cb[k * CB + i];
                    // bk: muF[2], B[2][2], Csym[3], mubar[2], c0
                    const f32x2 nmu0 = pack2(-ck[6], -ck[6]), nmu1 = pack2(-ck[7], -ck[7]);
                    const f32x2 nb0 = pack2(-bk[0], -bk[0]), nb1 = pack2(-bk[1], -bk[1]);
                    float dm[PX], dr0[PX], dr1[PX];
#pragma unroll
                    for (int h = 0; h < PP; ++h) {
                        const f32x2 mp = pack2(m[k][2 * h], m[k][2 * h + 1]);
                        const f32x2 t0 = tp[0][k][h], t1 = tp[1][k][h];
                        const f32x2 v0 = add2(up[0][h], nmu0), v1 = add2(up[1][h], nmu1);
                        f32x2 q0 = fma2(sc2, t0, pack2(ck[0], ck[0])), q1 = fma2(sc2, t1, pack2(ck[1], ck[1]));
                        q0 = fma2(pack2(ck[2], ck[2]), v0, q0);
                        q1 = fma2(pack2(ck[4], ck[4]), v0, q1);
                        q0 = fma2(pack2(ck[3], ck[3]), v1, q0);
                        q1 = fma2(pack2(ck[5], ck[5]), v1, q1);
                        const f32x2 ff0 = add2(fp[0][h], nb0), ff1 = add2(fp[1][h], nb1);
                        f32x2 acc = pack2(bk[CB - 1], bk[CB - 1]);
                        f32x2 lin = fma2(ff0, pack2(bk[2], bk[2]), fma2(ff1, pack2(bk[4], bk[4]), pack2(bk[9], bk[9])));
                        lin = fma2(pack2(bk[6], bk[6]), v0, lin);
                        lin = fma2(pack2(bk[7], bk[7]), v1, lin);
                        acc = fma2(lin, v0, acc);
                        lin = fma2(ff0, pack2(bk[3], bk[3]), fma2(ff1, pack2(bk[5], bk[5]), pack2(bk[10], bk[10])));
                        lin = fma2(pack2(bk[8], bk[8]), v1, lin);
                        acc = fma2(lin, v1, acc);
                        unpack2(fma2(gp[0][h], q0, fma2(gp[1][h], q1, acc)), dm[2 * h], dm[2 * h + 1]);
                        if (a.unbounded) {
                            unpack2(mul2(gp[0][h], mp), dr0[2 * h], dr0[2 * h + 1]);
                            unpack2(mul2(gp[1][h], mp), dr1[2 * h], dr1[2 * h + 1]);
                        } else {
                            const f32x2 nt0 = mul2(t0, pack2(-1.0f, -1.0f)), nt1 = mul2(t1, pack2(-1.0f, -1.0f));
                            unpack2(mul2(mul2(mul2(gp[0][h], ds2), fma2(nt0, t0, one2)), mp), dr0[2 * h], dr0[2 * h + 1]);
                            unpack2(mul2(mul2(mul2(gp[1][h], ds2), fma2(nt1, t1, one2)), mp), dr1[2 * h], dr1[2 * h + 1]);
                        }
                    }
                    if (dmask) Pack<PX>::st(dmask + (long long)k * P + p, dm);
                    if (dresid) {
                        Pack<PX>::st(dresid + (long long)k * P + p, dr0);
                        Pack<PX>::st(dresid + (long long)(K + k) * P + p, dr1);
                    }
                }
                continue;
            }
            // ---- phase 1: recompute T (kept in r) and pred, segment-outer ------------------------
            float g[2][PX];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < PX; ++j) g[c][j] = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float ck[CF];
#pragma unroll
                for (int i = 0; i < CF; ++i) ck[i] = cf[k * CF + i];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    float u[DD];
                    if constexpr (D > 0) px_feats<D>(y[j], x[j], u);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float t = a.unbounded ? r[c][k][j] : tanh_scaled(r[c][k][j], a.ex2_scale);
                        r[c][k][j] = t;
                        float q = fmaf(a.scale, t, ck[c]);
                        if constexpr (D > 0) {
#pragma unroll
                            for (int d = 0; d < D; ++d) q = fmaf(ck[2 + c * D + d], u[d] - ck[2 + 2 * D + d], q);
                        }
                        g[c][j] = fmaf(m[k][j], q, g[c][j]);
                    }
                }
            }
            // ---- g_c = gs * phi'(F_c - pred_c); f keeps the clamped flow -------------------------
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float fc = clamp_flow(f[c][j], a.clamp_t);
                    f[c][j] = fc;
                    float phi, w;
                    loss_terms(fc - g[c][j], a, phi, w);
                    g[c][j] = gs * w;
                }
            // ---- phase 2: gradients, segment-outer; each segment's outputs are stored at once ----
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float ck[CF], bk[CB];
#pragma unroll
                for (int i = 0; i < CF; ++i) ck[i] = cf[k * CF + i];
#pragma unroll
                for (int i = 0; i < CB; ++i) bk[i] = cb[k * CB + i];
                float dm[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float mk = m[k][j];
                    const float t0 = r[0][k][j], t1 = r[1][k][j];
                    float q0 = fmaf(a.scale, t0, ck[0]), q1 = fmaf(a.scale, t1, ck[1]);
                    float acc = bk[CB - 1];
                    if constexpr (D > 0) {
                        // bk: muF[2], B[2][D], Csym[D(D+1)/2], mubar[D], c0
                        float u[DD], v[DD];
                        px_feats<D>(y[j], x[j], u);
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            v[d] = u[d] - ck[2 + 2 * D + d];
                            q0 = fmaf(ck[2 + d], v[d], q0);
                            q1 = fmaf(ck[2 + D + d], v[d], q1);
                        }
                        const float ff0 = f[0][j] - bk[0], ff1 = f[1][j] - bk[1];
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            float lin = fmaf(ff0, bk[2 + d], fmaf(ff1, bk[2 + D + d], bk[2 + 2 * D + D * (D + 1) / 2 + d]));
#pragma unroll
                            for (int e = d; e < D; ++e) lin = fmaf(bk[2 + 2 * D + rcf_sym_idx(D, d, e)], v[e], lin);
                            acc = fmaf(lin, v[d], acc);
                        }
                    }
                    dm[j] = fmaf(g[0][j], q0, fmaf(g[1][j], q1, acc));
                    if (a.unbounded) {
                        r[0][k][j] = g[0][j] * mk;
                        r[1][k][j] = g[1][j] * mk;
                    } else {
                        r[0][k][j] = g[0][j] * a.dres_scale * fmaf(-t0, t0, 1.0f) * mk;
                        r[1][k][j] = g[1][j] * a.dres_scale * fmaf(-t1, t1, 1.0f) * mk;
                    }
                }
                if (dmask) Pack<PX>::st(dmask + (long long)k * P + p, dm);
                if (dresid) {
                    Pack<PX>::st(dresid + (long long)k * P + p, r[0][k]);
                    Pack<PX>::st(dresid + (long long)(K + k) * P + p, r[1][k]);
                }
            }
        }
    }
}

template <int K, int D>
static cudaError_t launch_kd(const RcfK& a, bool vec, cudaStream_t s) {
    dim3 grid(a.nchunkb, a.nfd), block(RCF_BLOCK);
    constexpr int VPX = K <= 4 ? 4 : 2;     // pixels per thread on the vector path (register budget)
    if (vec) rcf_launch(k_bwd<K, D, VPX>, grid, block, 0, s, a.pdl, a);
    else rcf_launch(k_bwd<K, D, 1>, grid, block, 0, s, a.pdl, a);
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

cudaError_t rcf_launch_bwd(const RcfK& a, bool vec, cudaStream_t s) {
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
