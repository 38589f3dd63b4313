// rcf_head.cu -- the default head (64 feature channels, 3x3 second conv) as one C call each way.
//
// Pure host code: it chains the library's own entry points (stem -> weight pack -> tcgen05 conv -> rcf_forward;
// rcf_backward -> data gradient -> weight gradient -> stem backward) on one stream, so the Python mirror
// (fused_head.py) crosses the ctypes boundary twice per step instead of nine times.  Reference lines: the body of
// FlowAggregationHeadWithResidual.forward for one call, models/flow_aggregation_head_with_residual.py:312-399.
#include <cstdint>

#include "rcf_loss.h"
#include "rcf_internal.h"

namespace {
constexpr int HEAD_CF = 64;

// Conv precision of the head (the `nprod` argument, chosen by the caller from torch's settings):
//   3  fp32-grade: bf16 (hi, lo) pairs of both operands, three products per fp32 product
//   2  TF32-class (torch's default, allow_tf32): ONE product of IEEE fp16 operands everywhere (11-bit significands, 2^-11 per
//      operand -- what TF32 keeps; tcgen05 kind::f16 wants both operands in the same format, mixing fp16 with bf16 is an
//      illegal instruction).  Activations and weights fit fp16's range as they are (the stem saturates at +-65504); the
//      feature-map gradient does not, so rcf_backward writes it multiplied by a power of two chosen on the device from
//      its magnitude (RcfDesc.dfeat_f16, rcf_grad_scale) and the weight-gradient / stem reductions divide it out again
//   1  autocast: plain bf16 operands everywhere
struct HeadModes { int stem, pack, fwd, dgrad, wgrad; bool a_lo, g_lo, g_f16; };
HeadModes head_modes(int nprod) {
    if (nprod >= 3) return {3, 2, 3, 3, 3, true, true, false};
    if (nprod == 2) {
        const int f = RCF_CONV64_A_F16 | RCF_CONV64_W_F16;
        return {2 | RCF_STEM_OUT_F16, 2 | RCF_CONV64_W_F16, 1 | f, 1 | f, 1 | f, false, false, true};
    }
    return {1, 2, 1, 1, 1, false, false, false};
}

int check_head(const RcfDesc* d, const RcfInputs* in, const RcfHeadBuffers* hb) {
    if (!d || !in || !hb) return RCF_ERR_NULL;
    if (d->Cf != HEAD_CF || d->theta_mode != 1 || !d->feat_nhwc) return RCF_ERR_MODE;
    if (d->ndir < 1 || d->ndir > 2) return RCF_ERR_SHAPE;
    const long long img = (long long)d->H * d->W * HEAD_CF;
    for (int i = 0; i < d->ndir; ++i)
        if (d->feat_bstride[i] != img) return RCF_ERR_SHAPE;
    return RCF_OK;
}
}  // namespace

extern "C" int rcf_head_forward(const RcfDesc* desc, const RcfInputs* in, const float* cw1, const float* cb1, const float* cw2,
                                int ks, float stem_slope, int nprod, int resid_h, int resid_w, const RcfHeadBuffers* hb, float* loss,
                                void* ctx, void* ws, const RcfVisOut* vis, void* stream) {
    int rc = check_head(desc, in, hb);
    if (rc != RCF_OK) return rc;
    if (!cw1 || !cb1 || !cw2 || !hb->a_hi || !hb->sign || !hb->wpack || !hb->feat) return RCF_ERR_NULL;
    const int ndir = desc->ndir, B = desc->B, H = desc->H, W = desc->W;
    const long long img = (long long)H * W * HEAD_CF;
    const HeadModes m = head_modes(nprod);
    if (m.a_lo && !hb->a_lo) return RCF_ERR_NULL;
    rc = rcf_stem_forward_bf16(in->flow, desc->flow_bstride, ndir, B, H, W, ks, cw1, cb1, desc->clamp_t, stem_slope, hb->a_hi,
                               m.a_lo ? hb->a_lo : nullptr, hb->sign, m.stem, stream);
    if (rc != RCF_OK) return rc;
    rc = rcf_conv64_pack_weights(cw2, hb->wpack, m.pack, stream);
    if (rc != RCF_OK) return rc;
    rc = rcf_conv64_forward(hb->a_hi, m.a_lo ? hb->a_lo : nullptr, hb->wpack, hb->feat, ndir * B, H, W, m.fwd, stream);
    if (rc != RCF_OK) return rc;
    RcfInputs in2 = *in;
    RcfDesc d2 = *desc;
    for (int i = 0; i < ndir; ++i) in2.feat[i] = hb->feat + (long long)i * B * img;
    if (resid_h > 0 && resid_w > 0) {                      // reference :271-273, :294-296 (both directions in one launch)
        if (!hb->resid_up) return RCF_ERR_NULL;
        const long long per = (long long)B * 2 * desc->K * H * W;
        const float* rin[2] = {in->resid[0], ndir > 1 ? in->resid[1] : nullptr};
        float* rout[2] = {hb->resid_up, ndir > 1 ? hb->resid_up + per : nullptr};
        rc = rcf_resize_bilinear_forward(rin, rout, ndir, B * 2 * desc->K, resid_h, resid_w, H, W, 0, stream);
        if (rc != RCF_OK) return rc;
        for (int i = 0; i < ndir; ++i) {
            in2.resid[i] = rout[i];
            d2.resid_bstride[i] = (long long)2 * desc->K * H * W;
        }
    }
    return rcf_forward(&d2, &in2, loss, ctx, ws, vis, stream);
}

extern "C" int rcf_head_backward(const RcfDesc* desc, const RcfInputs* in, const float* grad_loss, const void* ctx, void* ws,
                                 RcfGrads* grads, int ks, float stem_slope, int nprod, int need_conv_grads, int resid_h, int resid_w,
                                 const RcfHeadBuffers* hb, void* stream) {
    int rc = check_head(desc, in, hb);
    if (rc != RCF_OK) return rc;
    if (!grads || !hb->feat) return RCF_ERR_NULL;
    const int ndir = desc->ndir, B = desc->B, H = desc->H, W = desc->W;
    const long long img = (long long)H * W * HEAD_CF;
    const HeadModes m = head_modes(nprod);
    RcfInputs in2 = *in;
    RcfGrads g2 = *grads;
    for (int i = 0; i < ndir; ++i) {
        in2.feat[i] = hb->feat + (long long)i * B * img;
        g2.dfeat[i] = nullptr;
        g2.dfeat_hi[i] = g2.dfeat_lo[i] = nullptr;
        if (need_conv_grads) {
            if (!hb->g_hi || (m.g_lo && !hb->g_lo)) return RCF_ERR_NULL;
            g2.dfeat_hi[i] = static_cast<uint16_t*>(hb->g_hi) + (long long)i * B * img;
            if (m.g_lo) g2.dfeat_lo[i] = static_cast<uint16_t*>(hb->g_lo) + (long long)i * B * img;
        }
    }
    if (!need_conv_grads) g2.dfeat_bias = nullptr;
    RcfDesc d2 = *desc;
    const bool lowres = resid_h > 0 && resid_w > 0;
    const long long per = (long long)B * 2 * desc->K * H * W;
    if (lowres) {
        if (!hb->resid_up) return RCF_ERR_NULL;
        for (int i = 0; i < ndir; ++i) {
            in2.resid[i] = hb->resid_up + (long long)i * per;
            d2.resid_bstride[i] = (long long)2 * desc->K * H * W;
            if (grads->dresid[i]) {
                if (!hb->dresid_up) return RCF_ERR_NULL;
                g2.dresid[i] = hb->dresid_up + (long long)i * per;
                d2.dresid_bstride[i] = (long long)2 * desc->K * H * W;
            }
        }
    }
    d2.dfeat_f16 = (need_conv_grads && m.g_f16) ? 1 : 0;
    rc = rcf_backward(&d2, &in2, grad_loss, ctx, ws, &g2, stream);
    if (rc != RCF_OK) return rc;
    if (lowres) {                                          // gradient of the up-sampling: a deterministic gather
        const float* gout[2] = {nullptr, nullptr};
        float* gin[2] = {nullptr, nullptr};
        int n = 0;
        for (int i = 0; i < ndir; ++i)
            if (grads->dresid[i]) { gout[n] = g2.dresid[i]; gin[n] = grads->dresid[i]; ++n; }
        if (n > 0) {
            rc = rcf_resize_bilinear_backward(gout, gin, n, B * 2 * desc->K, resid_h, resid_w, H, W, 0, stream);
            if (rc != RCF_OK) return rc;
        }
    }
    if (!need_conv_grads) return RCF_OK;
    if (!hb->a_hi || !hb->sign || !hb->wpack || !hb->d_a1 || !hb->wgrad_ws || !hb->stem_ws || !hb->d_cw1 || !hb->d_cb1 || !hb->d_cw2)
        return RCF_ERR_NULL;
    const void* wp_bwd = static_cast<const uint8_t*>(hb->wpack) + RCF_CONV64_WPACK_BYTES;
    rc = rcf_conv64_forward(hb->g_hi, m.g_lo ? hb->g_lo : nullptr, wp_bwd, hb->d_a1, ndir * B, H, W, m.dgrad, stream);
    if (rc != RCF_OK) return rc;
    // fp16 mode: g_hi = dG * s, hence d_a1 = dA1 * s; the two reductions below divide s out (gmax: where s comes from)
    const float* gmax = m.g_f16 ? rcf_ws_gmax(&d2, ws) : nullptr;
    rc = rcf_conv64_wgrad_ex(hb->a_hi, m.a_lo ? hb->a_lo : nullptr, hb->g_hi, m.g_lo ? hb->g_lo : nullptr, hb->d_cw2, hb->wgrad_ws,
                             ndir * B, H, W, m.wgrad, gmax, ndir * B, stream);
    if (rc != RCF_OK) return rc;
    return rcf_stem_backward_ex(in->flow, desc->flow_bstride, ndir, B, H, W, HEAD_CF, ks, desc->clamp_t, stem_slope, nullptr, hb->sign,
                                hb->d_a1, hb->d_cw1, hb->d_cb1, hb->stem_ws, nprod, gmax, ndir * B, stream);
}
