// rcf_stem.cu -- the first layer of flow_feat_before_agg as hand-written kernels:
//     act = LeakyReLU(conv_{ks x ks, pad (ks-1)/2}(clamp(flow)) + bias)          reference :84-88, :155-156, :246
// A 2-channel ks x ks convolution has 2*ks*ks = 18 inputs per output: it is bandwidth-bound (8 B/px in, 4*Cf B/px out),
// not a tensor-core shape, and cuDNN runs it (plus a separate bias add, LeakyReLU, LeakyReLU backward, a weight-gradient
// GEMM and a bias-gradient reduction) as six launches with layout conversions around them.  Here:
//   k_stem_fwd  : clamp + conv + bias + LeakyReLU in one pass; output channels-last ([N,H,W,Cf]), which is what the
//                 second (cuDNN tensor-core) convolution wants.  A thread owns 4 output channels (72 weights in registers
//                 for ks = 3) and walks the pixels of a 32x8 tile whose clamped flow (+halo) is staged in shared memory.
//   k_stem_bwd  : LeakyReLU backward + weight gradient + bias gradient in one pass over (act, dact); per-thread
//                 accumulators, fixed-order combination inside the CTA, per-CTA partials reduced by k_stem_bwd_final
//                 in fp64 => bit-reproducible.  No input gradient: the RAFT flow carries no grad.
#include "rcf_common.cuh"

namespace {

struct StemK {
    const float* flow[2];
    long long flow_bs[2];
    int ndir, B, H, W, P, Cf;
    float clamp_t, slope;
    const float* w;      // [Cf][2][KS][KS]
    const float* b;      // [Cf]
    float* act;          // [N][P][Cf]
    const float* act_in; // backward: forward output (sign of the pre-activation)
    const float* dact;   // [N][P][Cf]
    float* part;         // [gridDim.x][Cf*(NT+1)]
    int ntiles;          // N * tiles per image
};

// A CTA works on a TW x TH pixel tile; the clamped flow tile plus its (KS-1)/2 halo (zero outside the frame = the
// conv's zero padding) is staged once in shared memory and every thread reads its 2*KS*KS taps from there
// (the Cf/4 threads of a pixel hit the same words: broadcast).
constexpr int STEM_TW = 32, STEM_TH = 8;

template <int KS>
__device__ __forceinline__ void stage_tile(const StemK& a, const float* __restrict__ fl, int x0, int y0,
                                           float (*tile)[STEM_TH + KS - 1][STEM_TW + KS - 1]) {
    constexpr int R = (KS - 1) / 2, SW = STEM_TW + KS - 1, SH = STEM_TH + KS - 1;
    for (int i = threadIdx.x; i < 2 * SH * SW; i += RCF_BLOCK) {
        const int c = i / (SH * SW), r = i - c * SH * SW;
        const int ly = r / SW, lx = r - ly * SW;
        const int y = y0 + ly - R, x = x0 + lx - R;
        float v = 0.0f;
        if (y >= 0 && y < a.H && x >= 0 && x < a.W) v = clamp_flow(__ldg(fl + (long long)c * a.P + (long long)y * a.W + x), a.clamp_t);
        tile[c][ly][lx] = v;
    }
}

template <int KS>
__device__ __forceinline__ void read_taps(const float (*tile)[STEM_TH + KS - 1][STEM_TW + KS - 1], int ly, int lx,
                                          float (&tap)[2 * KS * KS]) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) tap[(c * KS + dy) * KS + dx] = tile[c][ly + dy][lx + dx];
}

template <int KS>
__global__ void __launch_bounds__(RCF_BLOCK) k_stem_fwd(const StemK a) {
    constexpr int NT = 2 * KS * KS;
    __shared__ float tile[2][STEM_TH + KS - 1][STEM_TW + KS - 1];
    const int n = blockIdx.z;                       // image = dir * B + b
    const int dir = n / a.B, b = n - dir * a.B;
    const int tid = threadIdx.x, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int x0 = blockIdx.x * STEM_TW, y0 = blockIdx.y * STEM_TH;
    const float* __restrict__ fl = a.flow[dir] + (long long)b * a.flow_bs[dir];
    stage_tile<KS>(a, fl, x0, y0, tile);
    float w[4][NT], bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bias[j] = __ldg(a.b + c4 * 4 + j);
#pragma unroll
        for (int t = 0; t < NT; ++t) w[j][t] = __ldg(a.w + (size_t)(c4 * 4 + j) * NT + t);
    }
    __syncthreads();
    float4* __restrict__ out = reinterpret_cast<float4*>(a.act + (long long)n * a.P * Cf) + c4;
#pragma unroll 2
    for (int p = grp; p < STEM_TW * STEM_TH; p += groups) {
        const int ly = p / STEM_TW, lx = p - ly * STEM_TW;
        const int y = y0 + ly, x = x0 + lx;
        if (y < a.H && x < a.W) {
            float tap[NT];
            read_taps<KS>(tile, ly, lx, tap);
            float o[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fmaf(w[j][t], tap[t], o[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = o[j] >= 0.0f ? o[j] : a.slope * o[j];
            out[((long long)y * a.W + x) * nf4] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(RCF_BLOCK) k_stem_bwd(const StemK a) {
    constexpr int NT = 2 * KS * KS, NO = NT + 1;
    __shared__ float tile[2][STEM_TH + KS - 1][STEM_TW + KS - 1];
    extern __shared__ float red[];                   // [RCF_WARPS][Cf * NO]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Cf = a.Cf;
    const int nf4 = Cf >> 2, groups = RCF_BLOCK / nf4;
    const int c4 = tid % nf4, grp = tid / nf4;
    const int tx = (a.W + STEM_TW - 1) / STEM_TW, ty = (a.H + STEM_TH - 1) / STEM_TH;
    float dw[4][NT], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        db[j] = 0.0f;
#pragma unroll
        for (int t = 0; t < NT; ++t) dw[j][t] = 0.0f;
    }
    for (int tl = blockIdx.x; tl < a.ntiles; tl += gridDim.x) {      // fixed tile -> CTA assignment (reproducible)
        const int n = tl / (tx * ty), r = tl - n * tx * ty;
        const int y0 = (r / tx) * STEM_TH, x0 = (r - (r / tx) * tx) * STEM_TW;
        const int dir = n / a.B, b = n - dir * a.B;
        const float* __restrict__ fl = a.flow[dir] + (long long)b * a.flow_bs[dir];
        __syncthreads();                               // previous tile fully consumed
        stage_tile<KS>(a, fl, x0, y0, tile);
        __syncthreads();
        const float4* __restrict__ ap = reinterpret_cast<const float4*>(a.act_in + (long long)n * a.P * Cf) + c4;
        const float4* __restrict__ gp = reinterpret_cast<const float4*>(a.dact + (long long)n * a.P * Cf) + c4;
#pragma unroll 2
        for (int p = grp; p < STEM_TW * STEM_TH; p += groups) {
            const int ly = p / STEM_TW, lx = p - ly * STEM_TW;
            const int y = y0 + ly, x = x0 + lx;
            if (y < a.H && x < a.W) {
                const long long o = ((long long)y * a.W + x) * nf4;
                const float4 av = __ldg(ap + o), gv = __ldg(gp + o);
                float dpre[4] = {gv.x * (av.x >= 0.0f ? 1.0f : a.slope), gv.y * (av.y >= 0.0f ? 1.0f : a.slope),
                                 gv.z * (av.z >= 0.0f ? 1.0f : a.slope), gv.w * (av.w >= 0.0f ? 1.0f : a.slope)};
                float tap[NT];
                read_taps<KS>(tile, ly, lx, tap);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    db[j] += dpre[j];
#pragma unroll
                    for (int t = 0; t < NT; ++t) dw[j][t] = fmaf(dpre[j], tap[t], dw[j][t]);
                }
            }
        }
    }
    // lanes l, l + nf4, ... of a warp own the same channels: combine them with shuffles (nf4 <= 32, power of two)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        for (int o = 16; o >= nf4; o >>= 1) db[j] += __shfl_xor_sync(0xffffffffu, db[j], o);
#pragma unroll
        for (int t = 0; t < NT; ++t)
            for (int o = 16; o >= nf4; o >>= 1) dw[j][t] += __shfl_xor_sync(0xffffffffu, dw[j][t], o);
    }
    // every warp now holds all channel quads in its first nf4 lanes: combine the 8 warps through shared memory
    if (lane < nf4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* r = red + (size_t)warp * Cf * NO + (size_t)(lane * 4 + j) * NO;
#pragma unroll
            for (int t = 0; t < NT; ++t) r[t] = dw[j][t];
            r[NT] = db[j];
        }
    }
    __syncthreads();
    for (int o = tid; o < Cf * NO; o += RCF_BLOCK) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < RCF_WARPS; ++wi) v += red[(size_t)wi * Cf * NO + o];
        a.part[(size_t)blockIdx.x * Cf * NO + o] = v;
    }
}

__global__ void __launch_bounds__(256) k_stem_bwd_final(const float* __restrict__ part, int nparts, int Cf, int NT,
                                                        float* __restrict__ dw, float* __restrict__ db) {
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;   // warp per output f*(NT+1)+t
    if (o >= Cf * (NT + 1)) return;
    const double v = warp_sum_strided(part + o, nparts, (long long)Cf * (NT + 1), lane);
    if (lane == 0) {
        const int f = o / (NT + 1), t = o - f * (NT + 1);
        if (t < NT) dw[(size_t)f * NT + t] = (float)v;
        else db[f] = (float)v;
    }
}

int stem_check(int ndir, int B, int H, int W, int Cf, int ks) {
    if (ndir < 1 || ndir > 2 || B < 1 || H < 1 || W < 1) return RCF_ERR_SHAPE;
    if ((H + STEM_TH - 1) / STEM_TH > 65535) return RCF_ERR_SHAPE;
    if ((long long)ndir * B > 65535 || (long long)H * W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if (ks != 1 && ks != 3 && ks != 5) return RCF_ERR_UNSUPPORTED;
    if (Cf < 4 || Cf % 4 || Cf > 128 || 256 % (Cf / 4)) return RCF_ERR_UNSUPPORTED;
    return RCF_OK;
}

int stem_grid_bwd(int ntiles) { return ntiles < 2 * 148 ? ntiles : 2 * 148; }

void fill(StemK& a, const float* const* flow, const int64_t* bs, int ndir, int B, int H, int W, int Cf, float clamp_t,
          float slope) {
    a.flow[0] = flow[0]; a.flow[1] = ndir > 1 ? flow[1] : flow[0];
    a.flow_bs[0] = bs[0]; a.flow_bs[1] = ndir > 1 ? bs[1] : bs[0];
    a.ndir = ndir; a.B = B; a.H = H; a.W = W; a.P = H * W; a.Cf = Cf;
    a.clamp_t = clamp_t; a.slope = slope;
    a.ntiles = ndir * B * ((W + STEM_TW - 1) / STEM_TW) * ((H + STEM_TH - 1) / STEM_TH);
}

}  // namespace

#define RCF_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" int rcf_stem_workspace_bytes(int ndir, int B, int H, int W, int Cf, int ks, size_t* bytes) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!bytes) return RCF_ERR_NULL;
    const int ntiles = ndir * B * ((W + STEM_TW - 1) / STEM_TW) * ((H + STEM_TH - 1) / STEM_TH);
    *bytes = (size_t)stem_grid_bwd(ntiles) * Cf * (2 * ks * ks + 1) * sizeof(float);
    return RCF_OK;
}

extern "C" int rcf_stem_forward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                int Cf, int ks, const float* w, const float* b, float clamp_t, float slope, float* act,
                                void* stream) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!flow || !flow_bstride || !flow[0] || (ndir > 1 && !flow[1]) || !w || !b || !act) return RCF_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(act) & 15u) return RCF_ERR_ALIGN;
    StemK a{};
    fill(a, flow, flow_bstride, ndir, B, H, W, Cf, clamp_t, slope);
    a.w = w; a.b = b; a.act = act;
    dim3 grid((W + STEM_TW - 1) / STEM_TW, (H + STEM_TH - 1) / STEM_TH, ndir * B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (ks) {
        case 1: k_stem_fwd<1><<<grid, RCF_BLOCK, 0, s>>>(a); break;
        case 3: k_stem_fwd<3><<<grid, RCF_BLOCK, 0, s>>>(a); break;
        case 5: k_stem_fwd<5><<<grid, RCF_BLOCK, 0, s>>>(a); break;
    }
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}

extern "C" int rcf_stem_backward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                 int Cf, int ks, float clamp_t, float slope, const float* act, const float* dact,
                                 float* dw, float* db, void* ws, void* stream) {
    const int v = stem_check(ndir, B, H, W, Cf, ks);
    if (v != RCF_OK) return v;
    if (!flow || !flow_bstride || !flow[0] || (ndir > 1 && !flow[1]) || !act || !dact || !dw || !db || !ws) return RCF_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(dact)) & 15u) return RCF_ERR_ALIGN;
    StemK a{};
    fill(a, flow, flow_bstride, ndir, B, H, W, Cf, clamp_t, slope);
    a.act_in = act; a.dact = dact; a.part = static_cast<float*>(ws);
    const int g = stem_grid_bwd(a.ntiles);
    const int NT = 2 * ks * ks;
    const size_t smem = (size_t)RCF_WARPS * Cf * (NT + 1) * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (smem > 220 * 1024) return RCF_ERR_UNSUPPORTED;
    const bool big = smem > 40 * 1024;       // (plus the static tile) needs the opt-in shared-memory limit
    switch (ks) {
        case 1:
            if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_stem_bwd<1><<<g, RCF_BLOCK, smem, s>>>(a);
            break;
        case 3:
            if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_stem_bwd<3><<<g, RCF_BLOCK, smem, s>>>(a);
            break;
        case 5:
            if (big) RCF_CUDA(cudaFuncSetAttribute(k_stem_bwd<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_stem_bwd<5><<<g, RCF_BLOCK, smem, s>>>(a);
            break;
    }
    RCF_CUDA(cudaGetLastError());
    const int nout = Cf * (NT + 1);
    k_stem_bwd_final<<<(nout * 32 + 255) / 256, 256, 0, s>>>(a.part, g, Cf, NT, dw, db);
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}
