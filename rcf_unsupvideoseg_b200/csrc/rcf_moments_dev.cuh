// rcf_moments_dev.cuh -- pass-1 tile body of k_moments.
#pragma once
#include <type_traits>
#include "rcf_common.cuh"

// Segments are processed in groups of KG (as many as keep KG*NS accumulators in registers); inside a
// group the loop is pixel-outer so that all KG mask loads (+ the flow) of an iteration are in flight
// together and the warp reductions happen once per CTA, not once per segment.
template <int K, int D, int PX>
__device__ __forceinline__ void moments_tile(const RcfK& a, int fd, int chunk, float (*red)[K * rcf_ns(D)]) {
    constexpr int NS = rcf_ns(D);
    constexpr int CHUNK = rcf_chunk_mom(D, K);
    constexpr int ITER = CHUNK / (RCF_BLOCK * PX);
    constexpr int UNR = (CHUNK > RCF_CHUNK_MOM) ? 2 : ITER;     // the long affine chunk is unrolled by 2 (register budget of 3 CTAs/SM)
    constexpr int DD = D > 0 ? D : 1;
    constexpr int KG = (D == 5) ? 1 : (K < 4 ? K : 4);

    const int dir = fd / a.B;
    const int b = fd - dir * a.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ mask = a.mask[dir] + (long long)b * a.mask_bs[dir];
    const float* __restrict__ flow = a.flow[dir] + (long long)b * a.flow_bs[dir];
    const int P = a.P;
    const int p0 = chunk * CHUNK;

#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += KG) {
        float acc[KG * NS];
#pragma unroll
        for (int i = 0; i < KG * NS; ++i) acc[i] = 0.0f;
        if constexpr (D == 2) {
            // Affine fit.  The K x 12 outer-product accumulation is issue-bound as scalar FFMAs (measured: 44 M warp
            // instructions, issue active 57 %, 4.2 TB/s), so the 12 sums of a segment are kept as 6 packed pairs and
            // updated with FFMA2 (two round-to-nearest FMAs per issue slot, the mask value as the broadcast operand):
            //   P0 = (F0, F1)   P1 = (u0, u1)   P2 = F0*P1   P3 = F1*P1   P4 = u0*P1   P5 = (u1*u1, 1)
            // Every pair is produced in place (no register moves); same products and the same summation order per
            // accumulator as the scalar chain, so the results are bit-identical to it.
            f32x2 acc2[KG * 6];
#pragma unroll
            for (int i = 0; i < KG * 6; ++i) acc2[i] = 0ull;
            // Software pipeline: the loads of tile it+1 (KG mask packs + 2 flow packs per thread) are issued before tile
            // it is consumed.  Without it a warp alternates "issue loads / wait ~1 us / compute", and the 16 resident
            // warps do not cover the HBM latency.  Chunks that lie entirely inside the frame (all but the last one) run
            // without any bounds logic; the last chunk loads zeros for out-of-range pixels (their products vanish).
            // (row, col) of the thread's pack advance incrementally: one integer division per thread, not per pack.
            constexpr int STEP = RCF_BLOCK * PX;
            const int drow = STEP / a.W, dcol = STEP - drow * a.W;
            auto sweep = [&](auto guard_tag) {
                constexpr bool GUARD = decltype(guard_tag)::value;
                auto load = [&](int it, float (&m)[KG][PX], float (&f0)[PX], float (&f1)[PX]) {
                    const int p = p0 + it * STEP + tid * PX;
                    const bool in = !GUARD || p < P;
#pragma unroll
                    for (int k = 0; k < KG; ++k) {
                        if (in && k0 + k < K) Pack<PX>::ld(m[k], mask + (long long)(k0 + k) * P + p);
                        else {
#pragma unroll
                            for (int j = 0; j < PX; ++j) m[k][j] = 0.0f;
                        }
                    }
                    if (in) {
                        Pack<PX>::ld(f0, flow + p);
                        Pack<PX>::ld(f1, flow + P + p);
                    } else {
#pragma unroll
                        for (int j = 0; j < PX; ++j) f0[j] = f1[j] = 0.0f;
                    }
                };
                int row = (p0 + tid * PX) / a.W, col = (p0 + tid * PX) - row * a.W;
                float mc[KG][PX], f0c[PX], f1c[PX];
                load(0, mc, f0c, f1c);
#pragma unroll
                for (int it = 0; it < ITER; ++it) {
                    float mn[KG][PX], f0n[PX], f1n[PX];
                    if (it + 1 < ITER) load(it + 1, mn, f0n, f1n);
                    float y[PX], x[PX];
                    px_coords_rc<PX>(row, col, a, y, x);
                    row += drow; col += dcol;
                    if (col >= a.W) { col -= a.W; ++row; }
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        const float F0 = clamp_flow(f0c[j], a.clamp_t), F1 = clamp_flow(f1c[j], a.clamp_t);
                        f32x2 zp[6];
                        zp[0] = pack2(F0, F1);
                        zp[1] = pack2(y[j], x[j]);
                        zp[2] = mul2(pack2(F0, F0), zp[1]);
                        zp[3] = mul2(pack2(F1, F1), zp[1]);
                        zp[4] = mul2(pack2(y[j], y[j]), zp[1]);
                        zp[5] = pack2(x[j] * x[j], 1.0f);
#pragma unroll
                        for (int k = 0; k < KG; ++k) {
                            const f32x2 mm = pack2(mc[k][j], mc[k][j]);
#pragma unroll
                            for (int i = 0; i < 6; ++i) acc2[k * 6 + i] = fma2(mm, zp[i], acc2[k * 6 + i]);
                        }
                    }
                    if (it + 1 < ITER) {
#pragma unroll
                        for (int j = 0; j < PX; ++j) {
#pragma unroll
                            for (int k = 0; k < KG; ++k) mc[k][j] = mn[k][j];
                            f0c[j] = f0n[j];
                            f1c[j] = f1n[j];
                        }
                    }
                }
            };
            if (p0 + CHUNK <= P) sweep(std::false_type{});
            else sweep(std::true_type{});
#pragma unroll
            for (int k = 0; k < KG; ++k) {
                float* o = acc + k * NS;       // (1, F0, F1, u0, u1, F0u0, F0u1, F1u0, F1u1, u0u0, u0u1, u1u1)
                unpack2(acc2[k * 6 + 0], o[1], o[2]);
                unpack2(acc2[k * 6 + 1], o[3], o[4]);
                unpack2(acc2[k * 6 + 2], o[5], o[6]);
                unpack2(acc2[k * 6 + 3], o[7], o[8]);
                unpack2(acc2[k * 6 + 4], o[9], o[10]);
                unpack2(acc2[k * 6 + 5], o[11], o[0]);
            }
        } else
        if constexpr (D == 0) {
            // mask sums only: issue every load of the tile (ITER*K 128-bit loads per thread) before the first add,
            // so the tile runs at full memory-level parallelism
            float mm[ITER][KG][PX];
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const int p = p0 + (it * RCF_BLOCK + tid) * PX;
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    if (p < P) Pack<PX>::ld(mm[it][k], mask + (long long)(k0 + k) * P + p);
                    else {
#pragma unroll
                        for (int j = 0; j < PX; ++j) mm[it][k][j] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < ITER; ++it)
#pragma unroll
                for (int k = 0; k < KG; ++k)
#pragma unroll
                    for (int j = 0; j < PX; ++j) acc[k] += mm[it][k][j];
        } else {
#pragma unroll UNR
        for (int it = 0; it < ITER; ++it) {
            const int p = p0 + (it * RCF_BLOCK + tid) * PX;
            if (p < P) {
                float m[KG][PX];
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    if (k0 + k < K) Pack<PX>::ld(m[k], mask + (long long)(k0 + k) * P + p);
                    else {
#pragma unroll
                        for (int j = 0; j < PX; ++j) m[k][j] = 0.0f;
                    }
                }
                if constexpr (D == 0) {
#pragma unroll
                    for (int k = 0; k < KG; ++k)
#pragma unroll
                        for (int j = 0; j < PX; ++j) acc[k] += m[k][j];
                } else {
                    float f0[PX], f1[PX], y[PX], x[PX];
                    Pack<PX>::ld(f0, flow + p);
                    Pack<PX>::ld(f1, flow + P + p);
                    px_coords<PX>(p, a, y, x);
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        // z = (1, F0, F1, u, F0*u, F1*u, u_d*u_e): shared by all segments, then acc += m_k * z
                        float z[NS], u[DD];
                        px_feats<D>(y[j], x[j], u);
                        z[0] = 1.0f;
                        z[1] = clamp_flow(f0[j], a.clamp_t);
                        z[2] = clamp_flow(f1[j], a.clamp_t);
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            z[3 + d] = u[d];
                            z[3 + D + d] = z[1] * u[d];
                            z[3 + 2 * D + d] = z[2] * u[d];
#pragma unroll
                            for (int e = d; e < D; ++e) z[3 + 3 * D + rcf_sym_idx(D, d, e)] = u[d] * u[e];
                        }
#pragma unroll
                        for (int k = 0; k < KG; ++k) {
                            acc[k * NS] += m[k][j];
#pragma unroll
                            for (int s2 = 1; s2 < NS; ++s2) acc[k * NS + s2] = fmaf(m[k][j], z[s2], acc[k * NS + s2]);
                        }
                    }
                }
            }
        }
        }   // D > 0
        // one vector reduction per group; rows past K (padding of the last group) are never stored
        if (k0 + KG <= K) {
            warp_reduce_store<KG * NS>(acc, lane, &red[warp][k0 * NS]);
        } else {
#pragma unroll
            for (int k = 0; k < KG; ++k) {
                if (k0 + k < K) {
                    float one[NS];
#pragma unroll
                    for (int s2 = 0; s2 < NS; ++s2) one[s2] = acc[k * NS + s2];
                    warp_reduce_store<NS>(one, lane, &red[warp][(k0 + k) * NS]);
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < K * NS; i += RCF_BLOCK) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < RCF_WARPS; ++w) v += red[w][i];
        a.part1[((size_t)fd * (K * NS) + i) * a.nchunk1 + chunk] = v;
    }
}

