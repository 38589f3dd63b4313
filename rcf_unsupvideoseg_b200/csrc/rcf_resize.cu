// rcf_resize.cu -- bilinear resize of NCHW fp32 planes, forward and (deterministic) backward.
//
// Reference call sites (input staging):
//   models/flow_aggregation_head_with_residual.py:271-273, :294-296
//       all_pred_residual = F.interpolate(all_pred_residual, self.mask_size, mode='bilinear')   (align_corners=False)
//   models/rcf_model.py:438-442   mmseg `resize(gt_*_flows, size=mask_size, mode='bilinear', align_corners=...)`
// Both end in ATen's upsample_bilinear2d: per output pixel
//   src = align ? scale*dst : max(scale*(dst+0.5)-0.5, 0),  i0 = (int)src,  i1 = i0 + (i0 < in-1),  l1 = src-i0,  l0 = 1-l1
//   out = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11),      scale = align ? (in-1)/(out-1) : in/out   (fp32)
// ATen's forward launches 9 CTAs for the DAVIS training shape (8x8x48x48 -> 96x96: 36 us on a B200) and its backward is a
// zero-fill plus float atomics (order-dependent).  Here: one launch covers both directions' tensors, 128-bit stores, and
// the backward is a GATHER (each input pixel sums the output pixels whose footprint contains it, in a fixed order), so
// gradients are bit-reproducible and need no zero-fill.
#include "rcf_common.cuh"

namespace {

struct ResizeK {
    const float* src[2];
    float* dst[2];
    int h, w, H, W;          // input (h,w) -> output (H,W) of the FORWARD op
    float sy, sx;            // source-index scales
    int align;
};

__device__ __forceinline__ float src_index(float scale, int dst, int align) {
    if (align) return scale * (float)dst;
    const float s = scale * ((float)dst + 0.5f) - 0.5f;
    return s < 0.0f ? 0.0f : s;
}

struct Tap1 { int i0, i1; float l0, l1; };
__device__ __forceinline__ Tap1 make_tap1(float scale, int dst, int n_in, int align) {
    Tap1 t;
    const float s = src_index(scale, dst, align);
    t.i0 = min((int)s, n_in - 1);
    t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
    t.l1 = s - (float)t.i0;
    t.l0 = 1.0f - t.l1;
    return t;
}

// PX consecutive output pixels (flat index within the plane) per thread.
template <int PX>
__global__ void __launch_bounds__(256) k_resize_fwd(const ResizeK a) {
    const int P = a.H * a.W;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (p >= P) return;
    const size_t plane = blockIdx.y;
    const float* __restrict__ in = (blockIdx.z ? a.src[1] : a.src[0]) + plane * (size_t)a.h * a.w;
    float* __restrict__ out = (blockIdx.z ? a.dst[1] : a.dst[0]) + plane * (size_t)P;
    int y = p / a.W, x = p - y * a.W;
    Tap1 ty = make_tap1(a.sy, y, a.h, a.align);
    float v[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const Tap1 tx = make_tap1(a.sx, x, a.w, a.align);
        const float* r0 = in + (size_t)ty.i0 * a.w;
        const float* r1 = in + (size_t)ty.i1 * a.w;
        v[j] = ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1))
             + ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
        if (++x == a.W) { x = 0; ++y; if (j + 1 < PX && y < a.H) ty = make_tap1(a.sy, y, a.h, a.align); }
    }
    Pack<PX>::st(out + p, v);
}

// Conservative range of output indices whose taps can touch input index j (every candidate is then tested exactly).
__device__ __forceinline__ void cand_range(float scale, int j, int n_out, int align, int& lo, int& hi) {
    if (!(scale > 1e-12f)) { lo = 0; hi = n_out - 1; return; }
    const float inv = 1.0f / scale;
    float a, b;
    if (align) { a = ((float)j - 1.0f) * inv; b = ((float)j + 1.0f) * inv; }
    else { a = ((float)j - 0.5f) * inv - 0.5f; b = ((float)j + 1.5f) * inv - 0.5f; }
    a = fminf(fmaxf(a - 2.0f, 0.0f), (float)(n_out - 1));
    b = fminf(fmaxf(b + 2.0f, 0.0f), (float)(n_out - 1));
    lo = (int)a; hi = (int)b;
}

__device__ __forceinline__ float tap_weight(const Tap1& t, int j) {
    return (t.i0 == j ? t.l0 : 0.0f) + (t.i1 == j ? t.l1 : 0.0f);
}

// One thread per INPUT pixel (of the forward op): dst = grad_in [h,w], src = grad_out [H,W].
__global__ void __launch_bounds__(256) k_resize_bwd(const ResizeK a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.h * a.w) return;
    const size_t plane = blockIdx.y;
    const float* __restrict__ g = (blockIdx.z ? a.src[1] : a.src[0]) + plane * (size_t)a.H * a.W;
    float* __restrict__ gi = (blockIdx.z ? a.dst[1] : a.dst[0]) + plane * (size_t)a.h * a.w;
    const int j = p / a.w, i = p - j * a.w;
    int ylo, yhi, xlo, xhi;
    cand_range(a.sy, j, a.H, a.align, ylo, yhi);
    cand_range(a.sx, i, a.W, a.align, xlo, xhi);
    // shrink the column range to the taps that really touch column i (weights are re-evaluated per row otherwise)
    while (xlo <= xhi && tap_weight(make_tap1(a.sx, xlo, a.w, a.align), i) == 0.0f) ++xlo;
    while (xhi >= xlo && tap_weight(make_tap1(a.sx, xhi, a.w, a.align), i) == 0.0f) --xhi;
    float acc = 0.0f;
    for (int y = ylo; y <= yhi; ++y) {
        const float wy = tap_weight(make_tap1(a.sy, y, a.h, a.align), j);
        if (wy == 0.0f) continue;
        const float* row = g + (size_t)y * a.W;
        float racc = 0.0f;
        for (int x = xlo; x <= xhi; ++x)
            racc = fmaf(tap_weight(make_tap1(a.sx, x, a.w, a.align), i), __ldg(row + x), racc);
        acc = fmaf(wy, racc, acc);
    }
    gi[p] = acc;
}

// Exact 2x upsampling without align_corners (the DAVIS configs: residual 48x48 -> mask_size 96x96): the transpose of the
// interpolation is a fixed 4x4 stencil -- input index j receives the outputs 2j-1, 2j, 2j+1, 2j+2 with weights
// 1/4, 3/4, 3/4, 1/4; at the borders the clamped source index folds the missing neighbour's 1/4 into the 3/4 tap.
// No searches, no float->int conversions, 16 independent loads per thread.
__device__ __forceinline__ void taps_2x(int j, int n_in, int (&idx)[4], float (&wt)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { idx[q] = 2 * j - 1 + q; wt[q] = (q == 0 || q == 3) ? 0.25f : 0.75f; }
    if (j == 0) { idx[0] = 0; wt[0] = 0.0f; wt[1] = 1.0f; }
    if (j == n_in - 1) { idx[3] = 2 * j + 1; wt[3] = 0.0f; wt[2] = 1.0f; }
}

// Forward counterpart: one thread per INPUT pixel (r, j) writes the 2x2 output block it anchors from its 3x3
// neighbourhood (clamped at the borders, where the clamped source index puts the whole weight on the border pixel):
//   even output index 2j   = 1/4 in(j-1) + 3/4 in(j)        (j = 0: in(0))
//   odd  output index 2j+1 = 3/4 in(j)   + 1/4 in(j+1)      (j = n-1: in(n-1))
// 9 loads, 4 outputs, two 64-bit stores; no float->int conversions.
__global__ void __launch_bounds__(256) k_resize_fwd_2x(const ResizeK a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.h * a.w) return;
    const size_t plane = blockIdx.y;
    const float* __restrict__ in = (blockIdx.z ? a.src[1] : a.src[0]) + plane * (size_t)a.h * a.w;
    float* __restrict__ out = (blockIdx.z ? a.dst[1] : a.dst[0]) + plane * (size_t)a.H * a.W;
    const int r = p / a.w, j = p - r * a.w;
    const int rm = max(r - 1, 0), rp = min(r + 1, a.h - 1), jm = max(j - 1, 0), jp = min(j + 1, a.w - 1);
    float v[3][3];
    const int rows[3] = {rm, r, rp}, cols[3] = {jm, j, jp};
#pragma unroll
    for (int y = 0; y < 3; ++y)
#pragma unroll
        for (int x = 0; x < 3; ++x) v[y][x] = __ldg(in + (size_t)rows[y] * a.w + cols[x]);
    // horizontal pass (ATen order: the x interpolation first, then y): even column, odd column for each of the 3 rows
    float he[3], ho[3];
#pragma unroll
    for (int y = 0; y < 3; ++y) {
        he[y] = (j == 0) ? v[y][1] : 0.25f * v[y][0] + 0.75f * v[y][1];
        ho[y] = (j == a.w - 1) ? v[y][1] : 0.75f * v[y][1] + 0.25f * v[y][2];
    }
    const float e0 = (r == 0) ? he[1] : 0.25f * he[0] + 0.75f * he[1];
    const float o0 = (r == 0) ? ho[1] : 0.25f * ho[0] + 0.75f * ho[1];
    const float e1 = (r == a.h - 1) ? he[1] : 0.75f * he[1] + 0.25f * he[2];
    const float o1 = (r == a.h - 1) ? ho[1] : 0.75f * ho[1] + 0.25f * ho[2];
    float* o = out + (size_t)(2 * r) * a.W + 2 * j;
    *reinterpret_cast<float2*>(o) = make_float2(e0, o0);
    *reinterpret_cast<float2*>(o + a.W) = make_float2(e1, o1);
}

__global__ void __launch_bounds__(256) k_resize_bwd_2x(const ResizeK a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.h * a.w) return;
    const size_t plane = blockIdx.y;
    const float* __restrict__ g = (blockIdx.z ? a.src[1] : a.src[0]) + plane * (size_t)a.H * a.W;
    float* __restrict__ gi = (blockIdx.z ? a.dst[1] : a.dst[0]) + plane * (size_t)a.h * a.w;
    const int j = p / a.w, i = p - j * a.w;
    int yi[4], xi[4];
    float wy[4], wx[4];
    taps_2x(j, a.h, yi, wy);
    taps_2x(i, a.w, xi, wx);
    float v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = __ldg(g + (size_t)yi[r] * a.W + xi[c]);
    float acc = 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float racc = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) racc = fmaf(wx[c], v[r][c], racc);
        acc = fmaf(wy[r], racc, acc);
    }
    gi[p] = acc;
}

// RAFT flows as stored on disk / produced by the loader: [N, h, w, C] (HWC, C = 2; dataset/data.py:114-133), optionally
// scaled per channel (FlowTransform.scale_flow, dataset/transforms.py:831-851), transposed to NCHW (np.transpose(flow,
// (2,0,1)), :850) and resized to mask_size (models/rcf_model.py:438-442) -- three passes in the reference, one here.
// One thread per output pixel; the C channels of a source pixel are adjacent, so the four taps are four short vector reads.
struct StageK {
    const float* src;     // [N, h, w, C]
    float* dst;           // [N, C, H, W]
    int C, h, w, H, W, align;
    float sy, sx;
    float cscale[4];
};

__global__ void __launch_bounds__(256) k_flow_stage_hwc(const StageK a) {
    const int P = a.H * a.W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const size_t n = blockIdx.y;
    const float* __restrict__ in = a.src + n * (size_t)a.h * a.w * a.C;
    float* __restrict__ out = a.dst + n * (size_t)a.C * P;
    const int y = p / a.W, x = p - y * a.W;
    const Tap1 ty = make_tap1(a.sy, y, a.h, a.align), tx = make_tap1(a.sx, x, a.w, a.align);
    const float* p00 = in + ((size_t)ty.i0 * a.w + tx.i0) * a.C;
    const float* p01 = in + ((size_t)ty.i0 * a.w + tx.i1) * a.C;
    const float* p10 = in + ((size_t)ty.i1 * a.w + tx.i0) * a.C;
    const float* p11 = in + ((size_t)ty.i1 * a.w + tx.i1) * a.C;
    for (int c = 0; c < a.C; ++c) {
        const float v = ty.l0 * (tx.l0 * __ldg(p00 + c) + tx.l1 * __ldg(p01 + c))
                      + ty.l1 * (tx.l0 * __ldg(p10 + c) + tx.l1 * __ldg(p11 + c));
        out[(size_t)c * P + p] = v * a.cscale[c];
    }
}

float host_scale(int n_in, int n_out, int align) {
    if (align) return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
    return (float)n_in / (float)n_out;
}

int check(const void* const* a, const void* const* b, int nten, long long planes, int h, int w, int H, int W) {
    if (!a || !b) return RCF_ERR_NULL;
    if (nten < 1 || nten > 2 || planes < 1 || planes > 65535 || h < 1 || w < 1 || H < 1 || W < 1) return RCF_ERR_SHAPE;
    if ((long long)h * w > 0x7fffffffLL / 4 || (long long)H * W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    for (int t = 0; t < nten; ++t) {
        if (!a[t] || !b[t]) return RCF_ERR_NULL;
        if ((reinterpret_cast<uintptr_t>(a[t]) & 3u) || (reinterpret_cast<uintptr_t>(b[t]) & 3u)) return RCF_ERR_ALIGN;
    }
    return RCF_OK;
}

}  // namespace

extern "C" int rcf_resize_bilinear_forward(const float* const* in, float* const* out, int nten, int planes, int h, int w,
                                           int H, int W, int align_corners, void* stream) {
    const int v = check(reinterpret_cast<const void* const*>(in), reinterpret_cast<const void* const*>(out), nten, planes, h, w, H, W);
    if (v != RCF_OK) return v;
    ResizeK a{};
    bool vec = ((long long)H * W) % 4 == 0;
    for (int t = 0; t < nten; ++t) {
        a.src[t] = in[t]; a.dst[t] = out[t];
        if (reinterpret_cast<uintptr_t>(out[t]) & 15u) vec = false;
    }
    a.h = h; a.w = w; a.H = H; a.W = W; a.align = align_corners ? 1 : 0;
    a.sy = host_scale(h, H, a.align); a.sx = host_scale(w, W, a.align);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int P = H * W;
    bool al8 = true;
    for (int t = 0; t < nten; ++t) al8 = al8 && !(reinterpret_cast<uintptr_t>(out[t]) & 7u);
    if (!a.align && H == 2 * h && W == 2 * w && al8)
        k_resize_fwd_2x<<<dim3((h * w + 255) / 256, planes, nten), 256, 0, s>>>(a);
    else if (vec) k_resize_fwd<4><<<dim3((P / 4 + 255) / 256, planes, nten), 256, 0, s>>>(a);
    else k_resize_fwd<1><<<dim3((P + 255) / 256, planes, nten), 256, 0, s>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int rcf_resize_bilinear_backward(const float* const* grad_out, float* const* grad_in, int nten, int planes, int h,
                                            int w, int H, int W, int align_corners, void* stream) {
    const int v = check(reinterpret_cast<const void* const*>(grad_out), reinterpret_cast<const void* const*>(grad_in), nten, planes, h, w, H, W);
    if (v != RCF_OK) return v;
    ResizeK a{};
    for (int t = 0; t < nten; ++t) { a.src[t] = grad_out[t]; a.dst[t] = grad_in[t]; }
    a.h = h; a.w = w; a.H = H; a.W = W; a.align = align_corners ? 1 : 0;
    a.sy = host_scale(h, H, a.align); a.sx = host_scale(w, W, a.align);
    const dim3 grid((h * w + 255) / 256, planes, nten);
    if (!a.align && H == 2 * h && W == 2 * w && h >= 2 && w >= 2)
        k_resize_bwd_2x<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    else
        k_resize_bwd<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int rcf_flow_stage_hwc(const float* in, float* out, int N, int C, int h, int w, int H, int W, int align_corners,
                                  const float* channel_scale_host, void* stream) {
    if (!in || !out) return RCF_ERR_NULL;
    if (N < 1 || N > 65535 || C < 1 || C > 4 || h < 1 || w < 1 || H < 1 || W < 1) return RCF_ERR_SHAPE;
    if ((long long)h * w > 0x7fffffffLL / 16 || (long long)H * W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3u) return RCF_ERR_ALIGN;
    StageK a{};
    a.src = in; a.dst = out; a.C = C; a.h = h; a.w = w; a.H = H; a.W = W; a.align = align_corners ? 1 : 0;
    a.sy = host_scale(h, H, a.align); a.sx = host_scale(w, W, a.align);
    for (int c = 0; c < 4; ++c) a.cscale[c] = (channel_scale_host && c < C) ? channel_scale_host[c] : 1.0f;
    k_flow_stage_hwc<<<dim3((H * W + 255) / 256, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}
