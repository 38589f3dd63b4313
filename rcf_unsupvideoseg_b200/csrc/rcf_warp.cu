// rcf_warp.cu -- sampling kernels of the reference's utils/warp_utils.py (used by the AMD baseline:
// models/amd/pwc_lite.py:199, models/amd/flow_loss.py:73-79), as sm_100a kernels behind the C ABI.
//
//   k_flow_warp_fwd : flow_warp (:84-94) = grid_sample(x, norm_grid(mesh + flow), bilinear, align_corners=True)
//                     with padding 'zeros' or 'border'.  align_corners=True makes the normalise/un-normalise
//                     round trip the identity, so the sampling position is simply (col + fx, row + fy).
//   k_flow_warp_bwd : gradient w.r.t. the flow (gather) and w.r.t. the sampled tensor (bilinear scatter, atomics
//                     -- same as ATen's grid_sampler backward).
//   k_corr_splat / k_corr_finish : get_corresponding_map (:27-81), bilinear forward splat with scatter_add.
//                     Accumulated in 64-bit fixed point (2^-32 resolution) so the result is independent of the
//                     order of the atomics: bit-reproducible, unlike scatter_add_ on floats.
// One thread per pixel, channel loop inside; gathers ride on L2 (flows are smooth).  HBM/L2-bound.
#include "rcf_common.cuh"

namespace {

struct Tap {
    int x0, y0;          // north-west corner
    float wx, wy;        // fractional position inside the cell
    float sx, sy;        // d(position)/d(flow): 0 where 'border' clamped the coordinate
};

// Sampling position = (col + fx, row + fy).  The integer cell and the fractional weights are derived from the
// FLOW's own floor/frac (exact), not from the rounded sum: at column 800 an fp32 sum has an ulp of 6e-5 pixels,
// which is what limits the reference's own fp32 grid_sample.
__device__ __forceinline__ Tap make_tap(int col, int row, float fx, float fy, int H, int W, int border) {
    Tap t;
    t.sx = 1.0f; t.sy = 1.0f;
    const float ffx = floorf(fx), ffy = floorf(fy);
    t.wx = fx - ffx; t.wy = fy - ffy;
    // clamp the integer part so that (int) conversion cannot overflow for wild flows
    t.x0 = col + (int)fminf(fmaxf(ffx, -2.0e9f), 2.0e9f - 4096.0f - (float)col) ;
    t.y0 = row + (int)fminf(fmaxf(ffy, -2.0e9f), 2.0e9f - 4096.0f - (float)row);
    if (border) {   // clip_coordinates + zero gradient where clipped (position <= 0 or >= size-1)
        if (t.x0 < 0 || (t.x0 == 0 && t.wx == 0.0f)) { t.x0 = 0; t.wx = 0.0f; t.sx = 0.0f; }
        else if (t.x0 >= W - 1) { t.x0 = W - 1; t.wx = 0.0f; t.sx = 0.0f; }
        if (t.y0 < 0 || (t.y0 == 0 && t.wy == 0.0f)) { t.y0 = 0; t.wy = 0.0f; t.sy = 0.0f; }
        else if (t.y0 >= H - 1) { t.y0 = H - 1; t.wy = 0.0f; t.sy = 0.0f; }
    }
    return t;
}

__device__ __forceinline__ bool inb(int x, int y, int H, int W) { return x >= 0 && x < W && y >= 0 && y < H; }

__global__ void __launch_bounds__(256) k_flow_warp_fwd(const float* __restrict__ x, const float* __restrict__ flow,
                                                       float* __restrict__ out, int C, int H, int W, int border) {
    const int P = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= P) return;
    const int row = p / W, col = p - row * W;
    const float* fl = flow + (size_t)b * 2 * P;
    const Tap t = make_tap(col, row, __ldg(fl + p), __ldg(fl + P + p), H, W, border);
    const float w00 = (1.0f - t.wx) * (1.0f - t.wy), w10 = t.wx * (1.0f - t.wy);
    const float w01 = (1.0f - t.wx) * t.wy, w11 = t.wx * t.wy;
    const bool b00 = inb(t.x0, t.y0, H, W), b10 = inb(t.x0 + 1, t.y0, H, W);
    const bool b01 = inb(t.x0, t.y0 + 1, H, W), b11 = inb(t.x0 + 1, t.y0 + 1, H, W);
    const int o00 = t.y0 * W + t.x0;
    const float* xb = x + (size_t)b * C * P;
    float* ob = out + (size_t)b * C * P;
    for (int c = 0; c < C; ++c) {
        const float* xc = xb + (size_t)c * P;
        float v = 0.0f;
        if (b00) v = fmaf(w00, __ldg(xc + o00), v);
        if (b10) v = fmaf(w10, __ldg(xc + o00 + 1), v);
        if (b01) v = fmaf(w01, __ldg(xc + o00 + W), v);
        if (b11) v = fmaf(w11, __ldg(xc + o00 + W + 1), v);
        ob[(size_t)c * P + p] = v;
    }
}

__global__ void __launch_bounds__(256) k_flow_warp_bwd(const float* __restrict__ x, const float* __restrict__ flow,
                                                       const float* __restrict__ gout, float* gx, float* __restrict__ gflow,
                                                       int C, int H, int W, int border) {
    const int P = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= P) return;
    const int row = p / W, col = p - row * W;
    const float* fl = flow + (size_t)b * 2 * P;
    const Tap t = make_tap(col, row, __ldg(fl + p), __ldg(fl + P + p), H, W, border);
    const float w00 = (1.0f - t.wx) * (1.0f - t.wy), w10 = t.wx * (1.0f - t.wy);
    const float w01 = (1.0f - t.wx) * t.wy, w11 = t.wx * t.wy;
    const bool b00 = inb(t.x0, t.y0, H, W), b10 = inb(t.x0 + 1, t.y0, H, W);
    const bool b01 = inb(t.x0, t.y0 + 1, H, W), b11 = inb(t.x0 + 1, t.y0 + 1, H, W);
    const int o00 = t.y0 * W + t.x0;
    const float* xb = x + (size_t)b * C * P;
    const float* gb = gout + (size_t)b * C * P;
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(gb + (size_t)c * P + p);
        if (gflow) {
            const float* xc = xb + (size_t)c * P;
            const float v00 = b00 ? __ldg(xc + o00) : 0.0f, v10 = b10 ? __ldg(xc + o00 + 1) : 0.0f;
            const float v01 = b01 ? __ldg(xc + o00 + W) : 0.0f, v11 = b11 ? __ldg(xc + o00 + W + 1) : 0.0f;
            gix = fmaf(g, (v10 - v00) * (1.0f - t.wy) + (v11 - v01) * t.wy, gix);
            giy = fmaf(g, (v01 - v00) * (1.0f - t.wx) + (v11 - v10) * t.wx, giy);
        }
        if (gx) {
            float* gc = gx + ((size_t)b * C + c) * P;
            if (b00) atomicAdd(gc + o00, w00 * g);
            if (b10) atomicAdd(gc + o00 + 1, w10 * g);
            if (b01) atomicAdd(gc + o00 + W, w01 * g);
            if (b11) atomicAdd(gc + o00 + W + 1, w11 * g);
        }
    }
    if (gflow) {
        float* gf = gflow + (size_t)b * 2 * P;
        gf[p] = gix * t.sx;
        gf[P + p] = giy * t.sy;
    }
}

#define RCF_FIX_SCALE 4294967296.0f   // 2^32

__global__ void __launch_bounds__(256) k_corr_splat(const float* __restrict__ coords, unsigned long long* acc,
                                                    int H, int W) {
    const int P = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= P) return;
    const float xx = __ldg(coords + (size_t)b * 2 * P + p), yy = __ldg(coords + (size_t)b * 2 * P + P + p);
    const float x1 = floorf(xx), y1 = floorf(yy);
    const float wx1 = xx - x1, wy1 = yy - y1;     // weight of the "ceil" corner along each axis
    if (!(fabsf(x1) < 1e9f) || !(fabsf(y1) < 1e9f)) return;   // NaN / inf coordinates contribute nothing
    const int xi = (int)x1, yi = (int)y1;
    unsigned long long* ab = acc + (size_t)b * P;
    const float w[4] = {(1.0f - wx1) * (1.0f - wy1), wx1 * (1.0f - wy1), (1.0f - wx1) * wy1, wx1 * wy1};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int cx = xi + (q & 1), cy = yi + (q >> 1);
        if (inb(cx, cy, H, W)) atomicAdd(ab + cy * W + cx, (unsigned long long)(w[q] * RCF_FIX_SCALE));
    }
}

__global__ void __launch_bounds__(256) k_corr_finish(const unsigned long long* __restrict__ acc, float* __restrict__ out,
                                                     size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)((double)acc[i] * (1.0 / 4294967296.0));
}

int check_dims(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || B > 65535) return RCF_ERR_SHAPE;
    if ((long long)H * W > 0x7fffffffLL) return RCF_ERR_SHAPE;
    return RCF_OK;
}

}  // namespace

#define RCF_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" int rcf_flow_warp_forward(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                     int pad_border, void* stream) {
    if (!x || !flow || !out) return RCF_ERR_NULL;
    const int v = check_dims(B, C, H, W);
    if (v != RCF_OK) return v;
    dim3 grid((H * W + 255) / 256, B);
    k_flow_warp_fwd<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, flow, out, C, H, W, pad_border ? 1 : 0);
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}

extern "C" int rcf_flow_warp_backward(const float* x, const float* flow, const float* grad_out, float* grad_x,
                                      float* grad_flow, int B, int C, int H, int W, int pad_border, void* stream) {
    if (!x || !flow || !grad_out) return RCF_ERR_NULL;
    const int v = check_dims(B, C, H, W);
    if (v != RCF_OK) return v;
    if (!grad_x && !grad_flow) return RCF_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (grad_x) RCF_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)B * C * H * W * sizeof(float), s));
    dim3 grid((H * W + 255) / 256, B);
    k_flow_warp_bwd<<<grid, 256, 0, s>>>(x, flow, grad_out, grad_x, grad_flow, C, H, W, pad_border ? 1 : 0);
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}

extern "C" int rcf_corresponding_map(const float* coords, float* out, void* scratch_u64, int B, int H, int W,
                                     void* stream) {
    if (!coords || !out || !scratch_u64) return RCF_ERR_NULL;
    const int v = check_dims(B, 1, H, W);
    if (v != RCF_OK) return v;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)B * H * W;
    RCF_CUDA(cudaMemsetAsync(scratch_u64, 0, n * sizeof(unsigned long long), s));
    dim3 grid((H * W + 255) / 256, B);
    k_corr_splat<<<grid, 256, 0, s>>>(coords, static_cast<unsigned long long*>(scratch_u64), H, W);
    RCF_CUDA(cudaGetLastError());
    k_corr_finish<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const unsigned long long*>(scratch_u64), out, n);
    RCF_CUDA(cudaGetLastError());
    return RCF_OK;
}
