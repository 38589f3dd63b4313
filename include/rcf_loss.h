/*
 * rcf_loss.h -- C ABI of librcf_loss.so: the RCF relaxed-common-fate motion loss, forward and
 * backward, as hand-written sm_100a CUDA kernels.
 *
 * The reference (TonyLianLong/RCF-UnsupVideoSeg) has no FFI on this path: the whole loss is the
 * Python body of models/flow_aggregation_head_with_residual.py:33-399 executed op by op through
 * ATen.  This header is therefore NEW surface; every entry point names the reference lines whose
 * arithmetic it replaces.  The Python drop-in (rcf_unsupvideoseg_b200/head.py) binds it with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers to fp32 unless said
 *     otherwise; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no allocation, no host synchronisation, no default-stream use inside rcf_forward/backward:
 *     the caller supplies `ctx` (state kept from forward to backward) and `ws` (scratch) sized by
 *     rcf_query_sizes();
 *   - return value: 0 ok, <0 invalid argument (RCF_ERR_*), >0 a cudaError_t.  Nothing throws or
 *     exits across the boundary (unlike tools/torchCRF, permutohedral_gpu.cu:44-53).
 *
 * Tensor layouts (identical to the reference's, NCHW fp32):
 *   mask  [B, K, H, W]   one frame's soft masks; batch stride free (views masks[:, i] of
 *                        [B,2,K,H,W] are taken as they are, reference :331-332), inner [K,H,W] dense
 *   flow  [B, 2, H, W]   RAFT flow of that frame (un-clamped; the clamp of :155-156 is fused)
 *   resid [B, 2K, H, W]  residual map, channel c*K+k  (unflatten(1,(2,K)), reference :274-275)
 *   feat  [B, Cf, H, W]  output of flow_feat_before_agg (reference :246); only when Cf > 0
 *   theta [B, 2, K]      per-segment constant flow (output of flow_feat_after_agg, :258) when supplied
 * A "direction" is one (mask, flow, resid[, feat]) quadruple: forward = (masks[:,0], gt_fw_flows[:,0],
 * all_pred_residual_fw), backward = (masks[:,1], gt_bw_flows[:,0], all_pred_residual_bw)
 * (reference :331-347).  One call processes ndir (1 or 2) directions.
 */
#ifndef RCF_LOSS_H_
#define RCF_LOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RCF_API __attribute__((visibility("default")))
#else
#define RCF_API
#endif

#define RCF_ABI_VERSION 4
#define RCF_MAX_K 8          /* mask_layer supported by the compiled kernels */
#define RCF_MAX_CF 256       /* num_flow_feat_channels supported by the segment kernels */

#define RCF_OK 0
#define RCF_ERR_NULL (-1)        /* required pointer missing */
#define RCF_ERR_SHAPE (-2)       /* B/K/H/W/Cf/D/ndir out of range */
#define RCF_ERR_UNSUPPORTED (-3) /* valid for the reference, not compiled here (e.g. K > RCF_MAX_K) */
#define RCF_ERR_ALIGN (-4)       /* pointer not 4-byte aligned / stride inconsistent */
#define RCF_ERR_MODE (-5)        /* inconsistent flags (e.g. theta_mode 1 without feat / weights) */

/* Problem descriptor (host memory, read during the call only). */
typedef struct RcfDesc {
    int32_t B;                 /* samples in this call (per rank)                                   */
    int32_t K;                 /* mask_layer, 1..RCF_MAX_K                         (ref :53,:75)    */
    int32_t H, W;              /* mask_size                                        (ref :60,:111)   */
    int32_t Cf;                /* num_flow_feat_channels when pooling feat, else 0 (ref :56)        */
    int32_t D;                 /* 0 free_residual | 2 free_residual_with_affine | 5 ..._quadratic
                                                                                   (ref :123-127)   */
    int32_t ndir;              /* 1 or 2                                                            */
    int32_t theta_mode;        /* 0: theta supplied; 1: pool feat + segment MLP    (ref :246-258)   */
    int32_t robust;            /* outlier_robust_loss                              (ref :359-368)   */
    int32_t unbounded_residual;/* residual_adjustment_scale == -1 && free_residual (ref :282-286)   */
    float eps, q;              /* robust loss (|d|+eps)^q                          (ref :58-59)     */
    float resid_scale;         /* residual_adjustment_scale                        (ref :61)        */
    float pred_div;            /* pred_div_coeff                                   (ref :73)        */
    float clamp_t;             /* clamp_flow_t; < 0 means "no clamp"               (ref :155-156)   */
    float inv_n;               /* 1 / (number of elements in the mean); 0 => 1/(B*2*H*W) (ref :361) */
    int64_t mask_bstride[2];   /* batch strides in ELEMENTS, per direction                          */
    int64_t flow_bstride[2];
    int64_t resid_bstride[2];
    int64_t feat_bstride[2];
    int64_t dmask_bstride[2];  /* strides of the gradient buffers (rcf_backward)                    */
    int64_t dresid_bstride[2];
    int64_t dfeat_bstride[2];
    /* visualisation flows (reference :18-30, :370-395): out[b*vis_bstride + dir*vis_dstride + c*H*W + p]
       = value * vis_scale[c].  The reference uses [B,4,H,W] with vis_scale = {2/H, 2/W}.            */
    int64_t vis_bstride;
    int64_t vis_dstride;
    float vis_scale[2];
    /* LeakyReLU fused into the feature loads: when feat holds the PRE-activation of the last conv of
       flow_feat_before_agg (reference :89-92), set the negative slope here (0.1) and the library applies the
       activation on load and its derivative to dfeat; 0 or 1 means feat is used as it is. */
    float feat_lrelu_slope;
    /* 0: feat / dfeat are [B,Cf,H,W] planes (NCHW);  1: channels-last, element (b,f,p) at b*bstride + p*Cf + f (what
       cuDNN's tensor-core convolutions produce natively; avoids its NCHW<->NHWC transposes).  Requires Cf % 4 == 0,
       Cf <= 128 and 256 % (Cf/4) == 0. */
    int32_t feat_nhwc;
    /* rcf_backward: 0 = grad_loss holds ndir floats (one upstream gradient per direction); 1 = grad_loss holds ONE float,
       the gradient of the total loss[ndir] = flow_loss['seg'] (reference :397), applied to every direction.  The
       reference's caller only ever differentiates 'seg' (rcf_model.py:464-470). */
    int32_t grad_loss_total;
    int32_t dfeat_f16;         /* channels-last head only: RcfGrads.dfeat_hi receives IEEE fp16 words of dfeat * 2^e (dfeat_lo unused),
                                  e chosen on the device from the gradient's magnitude (see rcf_head_backward); the library's own
                                  tcgen05 kernels consume it and divide the scale out of their results */
} RcfDesc;

typedef struct RcfInputs {
    const float* mask[2];
    const float* flow[2];
    const float* resid[2];
    const float* feat[2];      /* NULL when Cf == 0 */
    const float* theta[2];     /* theta_mode 0 */
    const float* w1;           /* flow_feat_after_agg.0.weight [Cf,Cf]   theta_mode 1 (ref :96) */
    const float* b1;           /* flow_feat_after_agg.0.bias   [Cf]                             */
    const float* w2;           /* flow_feat_after_agg.2.weight [2,Cf]                 (ref :99) */
    const float* b2;           /* flow_feat_after_agg.2.bias   [2]                              */
    const float* feat_bias;    /* optional [Cf] (channels-last feat only): bias of the last conv of flow_feat_before_agg
                                  (ref :89-91) added on load, so the conv runs bias-free and its separate bias-add and
                                  bias-gradient reduction kernels disappear; NULL = feat already contains the bias */
} RcfInputs;

/* Optional per-pixel outputs of the forward pass; any pointer may be NULL. */
typedef struct RcfVisOut {
    float* gt;     /* prepared (clamped) flow             -> flows['gt_flow']      */
    float* pred;   /* reconstructed flow                  -> flows['pred_flow']    */
    float* agg;    /* piece-wise constant part            -> flows['agg_flow']     */
    float* res;    /* residual part                       -> flows['residual_adj'] */
    float* aff;    /* affine / quadratic part (D > 0)     -> flows['affine_flow']  */
} RcfVisOut;

/* Gradient outputs; any pointer may be NULL (that gradient is skipped). */
typedef struct RcfGrads {
    float* dmask[2];   /* [B,K,H,W]  with dmask_bstride  */
    float* dresid[2];  /* [B,2K,H,W] with dresid_bstride */
    float* dfeat[2];   /* [B,Cf,H,W] with dfeat_bstride  (theta_mode 1) */
    float* dtheta[2];  /* [B,2,K] dense                  (theta_mode 0) */
    float* dw1;        /* [Cf,Cf]  summed over directions (theta_mode 1) */
    float* db1;        /* [Cf] */
    float* dw2;        /* [2,Cf] */
    float* db2;        /* [2] */
    float* dfeat_bias; /* [Cf] gradient of RcfInputs.feat_bias (sum over pixels of dfeat), or NULL */
    /* ABI 3, channels-last feat only: dfeat written as the bf16 pair dfeat ~ hi + lo (same element strides as dfeat,
       dfeat_bstride % 8 == 0) -- the operand format of rcf_conv64_forward / rcf_conv64_wgrad; dfeat_lo may be NULL.
       Independent of `dfeat` (either, both or none may be requested). */
    void* dfeat_hi[2];
    void* dfeat_lo[2];
} RcfGrads;

/* ABI version of the loaded library (== RCF_ABI_VERSION of the header it was built from). */
RCF_API int rcf_abi_version(void);

/* Human-readable text for a return code of this library (RCF_ERR_* or a cudaError_t). */
RCF_API const char* rcf_error_string(int code);

/* Bytes of `ctx` (forward -> backward state: per-segment statistics, replaces autograd's saved
 * per-pixel intermediates of reference :242-304) and of `ws` (scratch partial sums). */
RCF_API int rcf_query_sizes(const RcfDesc* desc, size_t* ctx_bytes, size_t* ws_bytes);

/* Forward: replaces norm_and_clamp_flow's clamp (:155-156), aggregate_flow_with_residual (:235-310,
 * incl. get_demean_affine_flow :164-233) and the loss terms (:359-368) for ndir directions.
 * loss (device, fp32, ndir + 1 floats): loss[dir] receives flow_loss['seg_fw'] / ['seg_bw'] and loss[ndir] their
 * sum flow_loss['seg'] (:397).  `vis` may be NULL. */
RCF_API int rcf_forward(const RcfDesc* desc, const RcfInputs* in, float* loss, void* ctx, void* ws,
                const RcfVisOut* vis, void* stream);

/* Backward of sum_dir grad_loss[dir] * loss[dir] (grad_loss: device, fp32, [ndir]; or one float, the gradient of
 * loss[ndir], when desc->grad_loss_total); replaces the autograd replay of every op of :242-368.  `ctx` must come from rcf_forward on the same inputs. */
RCF_API int rcf_backward(const RcfDesc* desc, const RcfInputs* in, const float* grad_loss, const void* ctx,
                 void* ws, const RcfGrads* grads, void* stream);

/* ---- utils/warp_utils.py sampling ops (reference :27-94; used by the AMD baseline, not by the RCF head) ----
 * All tensors dense NCHW fp32 device pointers.  pad_border: 0 = padding_mode 'zeros', 1 = 'border'. */

/* flow_warp (utils/warp_utils.py:84-94): out[b,c,y,x] = bilinear sample of x[b,c] at (x + flow[b,0], y + flow[b,1]),
 * i.e. grid_sample(..., mode='bilinear', align_corners=True).  x,out [B,C,H,W]; flow [B,2,H,W]. */
RCF_API int rcf_flow_warp_forward(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                                  int pad_border, void* stream);

/* Backward of flow_warp.  grad_x (zero-filled here, then bilinear scatter with atomics) and/or grad_flow may be NULL. */
RCF_API int rcf_flow_warp_backward(const float* x, const float* flow, const float* grad_out, float* grad_x,
                                   float* grad_flow, int B, int C, int H, int W, int pad_border, void* stream);

/* get_corresponding_map (utils/warp_utils.py:27-81): bilinear forward splat of ones at `coords` [B,2,H,W]
 * (absolute x,y positions) into out [B,1,H,W]; scratch_u64 = B*H*W 64-bit words (order-independent fixed-point sums). */
RCF_API int rcf_corresponding_map(const float* coords, float* out, void* scratch_u64, int B, int H, int W, void* stream);

/* ---- first layer of flow_feat_before_agg (reference :84-88): act = LeakyReLU(conv_ks(clamp(flow)) + bias) ----------
 * flow[dir]: [B,2,H,W] planes with batch stride flow_bstride[dir] (elements); w [Cf,2,ks,ks]; b [Cf];
 * act / dact: channels-last [ndir*B, H, W, Cf] dense (what the second, tensor-core convolution consumes natively).
 * ks in {1,3,5}; Cf % 4 == 0, Cf <= 128, 256 % (Cf/4) == 0, else RCF_ERR_UNSUPPORTED (use the framework's conv).
 * clamp_t < 0: no clamp.  No gradient w.r.t. the flow is produced (the RAFT flow carries none). */
RCF_API int rcf_stem_forward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W, int Cf,
                             int ks, const float* w, const float* b, float clamp_t, float slope, float* act,
                             uint32_t* sign, void* stream);
/* The same layer with the activation written as the bf16 pair act ~ act_hi + act_lo (channels-last [ndir*B, H, W, 64]
 * bf16 each; act_lo may be NULL) instead of fp32: the operand format of rcf_conv64_forward.  Cf = 64 only.
 * nprod (here and in rcf_stem_backward): TF32 products per fp32 product on the tensor-core path -- >= 3: hi*hi + hi*lo +
 * lo*hi (fp32-grade, what torch.backends.cudnn.allow_tf32 = False asks for); 1 or 2: one product of round-to-nearest TF32
 * operands (what cuDNN runs for this layer under allow_tf32 = True / autocast; a third of the MMAs). */
RCF_API int rcf_stem_forward_bf16(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W,
                                  int ks, const float* w, const float* b, float clamp_t, float slope, void* act_hi,
                                  void* act_lo, uint32_t* sign, int nprod, void* stream);
RCF_API int rcf_stem_workspace_bytes(int ndir, int B, int H, int W, int Cf, int ks, size_t* bytes);
/* dw [Cf,2,ks,ks], db [Cf] from dact (gradient w.r.t. act) and EITHER the forward output act OR the sign bits the forward
 * wrote; ws from rcf_stem_workspace_bytes.
 * sign (optional, Cf == 64 only): [ndir*B, H, W, Cf/32] 32-bit words, bit f%32 of word f/32 set <=> the pre-activation of
 * channel f is <= 0.  With Cf == 64 both passes run on tensor cores (mma.sync TF32, 3xTF32 split = fp32-grade accuracy)
 * and the backward reads these 8 B/px instead of the 256 B/px activation map. */
RCF_API int rcf_stem_backward(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W, int Cf,
                              int ks, float clamp_t, float slope, const float* act, const uint32_t* sign, const float* dact,
                              float* dw, float* db, void* ws, int nprod, void* stream);

/* ---- second layer of flow_feat_before_agg (reference :89-91): Conv2d(64 -> 64, 3x3, padding 1, no bias here) on the
 * 5th-generation tensor cores (TMA tiled loads -> tcgen05.mma with bf16 operands -> fp32 accumulation in tensor memory;
 * csrc/rcf_conv64.cu).  Replaces the cuDNN kernels ATen picks for nn.Conv2d.forward / its data gradient.
 * Activations are passed as TWO channels-last bf16 tensors [nimg, H, W, 64], x ~ in_hi + in_lo (what rcf_split_bf16 and
 * the library's own producers write); out: channels-last fp32 [nimg, H, W, 64]; all dense, 16-byte aligned.
 * wpack: RCF_CONV64_WPACK_BYTES device bytes filled by rcf_conv64_pack_weights from the [64,64,3,3] fp32 weight;
 * transpose_flip = 1 packs the operator of the DATA GRADIENT (din = conv(dout, W^T flipped)), so rcf_conv64_forward
 * computes it with the same kernel; transpose_flip = 2 packs BOTH in one launch (wpack then holds 2 x
 * RCF_CONV64_WPACK_BYTES: the forward image followed by the data-gradient image).
 * nprod (kernel mode): 3 = three bf16 products per fp32 product (hi + lo of both operands, fp32-grade ~1e-5), 2 = in_hi x
 * [w_hi | w_lo] stacked in the MMA's N (two products), 1 = in_hi x w_hi (one product; bf16 operands, or IEEE fp16 with the
 * format flags below = TF32-class, what rcf_head_* use at torch's default precision).  in_lo may be NULL unless nprod == 3. */
#define RCF_CONV64_WPACK_BYTES (9 * 16384 + 2 * (9 * 8192 + 9 * 4096))   /* one-CTA image + the two CTA-pair images */
/* Operand-format flags, OR-ed into `nprod` of rcf_conv64_forward / rcf_conv64_wgrad (nprod must then be 1) and into
 * `transpose_flip` of rcf_conv64_pack_weights: the 16-bit words are IEEE fp16 (11-bit significand) instead of bf16.
 *   RCF_CONV64_A_F16: the activation operand (in_hi of rcf_conv64_forward, x_hi of rcf_conv64_wgrad) is fp16;
 *   RCF_CONV64_W_F16: the packed weights are fp16 (pass it to the pack call AND to the conv call that uses that image).
 * One product of fp16 activations and fp16 weights has TF32-class accuracy (2^-11 per operand) at the cost of the plain
 * bf16 mode; gradients stay bf16 (their range does not fit fp16 without scaling).  rcf_head_* select this for nprod = 2. */
#define RCF_CONV64_A_F16 0x100
#define RCF_CONV64_W_F16 0x200
#define RCF_STEM_OUT_F16 0x100   /* OR-ed into `nprod` of rcf_stem_forward_bf16: act_hi is written as fp16 (saturating) */
RCF_API int rcf_conv64_pack_weights(const float* w, void* wpack, int transpose_flip, void* stream);
RCF_API int rcf_conv64_forward(const void* in_hi, const void* in_lo, const void* wpack, float* out, int nimg, int H, int W,
                               int nprod, void* stream);
/* Weight gradient of the same convolution (replaces cuDNN's wgrad kernel): dw [64,64,3,3] fp32 from the layer input
 * x ~ x_hi + x_lo and the output gradient g ~ g_hi + g_lo, all channels-last bf16 [nimg, H, W, 64]
 * (csrc/rcf_conv64_wgrad.cu: TMA row-segment tiles, tcgen05.mma with the pixel axis as K, six taps per MMA).
 * nprod 3: x_hi*g_hi + x_lo*g_hi + x_hi*g_lo (fp32-grade); 2: x_hi*(g_hi + g_lo) (x_lo may be NULL); 1: x_hi*g_hi.
 * ws: rcf_conv64_wgrad_workspace_bytes() device bytes (per-CTA partials, summed in a fixed order: bit-reproducible). */
RCF_API int rcf_conv64_wgrad_workspace_bytes(int nimg, int H, int W, size_t* bytes);
RCF_API int rcf_conv64_wgrad(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* ws,
                             int nimg, int H, int W, int nprod, void* stream);

/* ---- the default head as two calls (what fused_head.py binds): stem -> weight pack -> tcgen05 conv -> rcf_forward, and
 * rcf_backward -> data gradient -> weight gradient -> stem backward.  Same kernels and results as the individual entry
 * points above; one call each way keeps the host cost per step (ctypes marshalling, Python glue) off the launch-bound
 * 96x96 / 48x48 training shapes.  All buffers are caller-owned device memory; feat / dfeat pointers inside `in` / `grads`
 * are filled by the library from RcfHeadBuffers (desc->feat_nhwc = 1, Cf = 64, feat_bstride = dfeat_bstride = H*W*64).
 * nprod here is the head's conv-precision LEVEL: 3 = fp32-grade (bf16 (hi, lo) pairs, three products per fp32 product,
 * 3xTF32 stem; a_lo / g_lo required), 2 = TF32-class (ONE product of IEEE fp16 operands: a_hi and the packed weights hold
 * fp16, g_hi holds fp16(dfeat * 2^e) with e chosen on the device -- RcfDesc.dfeat_f16 --, the 2^e is divided out of dW2, dW1
 * and db1 inside; one-product TF32 stem), 1 = plain bf16 (autocast class). */
typedef struct RcfHeadBuffers {
    void* a_hi;            /* bf16 [ndir*B,H,W,64]: LeakyReLU(conv1(clamp(flow))), hi words                         */
    void* a_lo;            /* same shape, lo words; NULL unless nprod == 3                                          */
    uint32_t* sign;        /* [ndir*B*H*W*2] sign bits of the stem's pre-activation                                 */
    void* wpack;           /* forward: 2 x RCF_CONV64_WPACK_BYTES (both orientations are packed); backward: same    */
    float* feat;           /* fp32 [ndir*B,H,W,64]: conv2 output without bias                                       */
    /* backward only */
    void* g_hi;            /* bf16 [ndir*B,H,W,64]: gradient w.r.t. feat, hi words                                  */
    void* g_lo;            /* lo words; NULL when nprod == 1                                                        */
    float* d_a1;           /* fp32 [ndir*B,H,W,64]: gradient w.r.t. the stem's output                               */
    void* wgrad_ws;        /* rcf_conv64_wgrad_workspace_bytes(ndir*B, H, W)                                        */
    void* stem_ws;         /* rcf_stem_workspace_bytes(ndir, B, H, W, 64, ks)                                       */
    float* d_cw1;          /* [64,2,ks,ks] */
    float* d_cb1;          /* [64]         */
    float* d_cw2;          /* [64,64,3,3]  */
    /* residual predicted at a lower resolution (allow_residual_resize, reference :271-273, :294-296); resid_h > 0 only */
    float* resid_up;       /* fp32 [ndir][B,2K,H,W]: the bilinearly up-sampled residuals (forward writes, backward reads) */
    float* dresid_up;      /* fp32 [ndir][B,2K,H,W]: backward scratch for the full-resolution residual gradient        */
} RcfHeadBuffers;
/* resid_h, resid_w > 0: in->resid[i] are dense [B,2K,resid_h,resid_w] tensors that the library up-samples to [H,W]
 * (align_corners = 0) into hb->resid_up first; the backward then returns grads->dresid[i] at [resid_h,resid_w] too
 * (dense).  0: residuals are given at [H,W] with desc->resid_bstride as usual. */
RCF_API int rcf_head_forward(const RcfDesc* desc, const RcfInputs* in, const float* cw1, const float* cb1, const float* cw2,
                             int ks, float stem_slope, int nprod, int resid_h, int resid_w, const RcfHeadBuffers* hb, float* loss,
                             void* ctx, void* ws, const RcfVisOut* vis, void* stream);
/* need_conv_grads = 0: only rcf_backward runs (no gradient reaches the conv parameters). */
RCF_API int rcf_head_backward(const RcfDesc* desc, const RcfInputs* in, const float* grad_loss, const void* ctx, void* ws,
                              RcfGrads* grads, int ks, float stem_slope, int nprod, int need_conv_grads, int resid_h, int resid_w,
                              const RcfHeadBuffers* hb, void* stream);
/* x (n fp32 values, n % 4 == 0) -> hi = bf16(x), lo = bf16(x - hi) (lo may be NULL). */
RCF_API int rcf_split_bf16(const float* x, void* hi, void* lo, size_t n, void* stream);
/* Measurement hook: device buffer of 64 x 8 int64 filled by CTA 0 of the following conv launches with clock64 stamps
 * (NULL switches it off). */
RCF_API int rcf_debug_conv64_trace(void* buf);
/* Test hook (synchronises the device): 1 if a tcgen05 kernel of this process reported a barrier time-out since the
 * previous call (a protocol bug: results of that launch are invalid), 0 otherwise. */
RCF_API int rcf_debug_conv64_status(void);

/* ---- input staging: bilinear resize of dense NCHW fp32 planes ------------------------------------
 * Replaces F.interpolate(all_pred_residual, mask_size, mode='bilinear') of the head (reference :271-273, :294-296;
 * align_corners = 0) and mmseg's resize() of the RAFT flows in the caller (models/rcf_model.py:438-442; align_corners
 * from the config).  nten (1 or 2) tensors of `planes` = B*C planes each are resized [h,w] -> [H,W] in one launch.
 * The backward is a deterministic gather (no atomics, no zero-fill): grad_in[h,w] from grad_out[H,W]. */
RCF_API int rcf_resize_bilinear_forward(const float* const* in, float* const* out, int nten, int planes, int h, int w,
                                        int H, int W, int align_corners, void* stream);
RCF_API int rcf_resize_bilinear_backward(const float* const* grad_out, float* const* grad_in, int nten, int planes, int h,
                                         int w, int H, int W, int align_corners, void* stream);

/* RAFT flow staging: in [N, h, w, C] (HWC as np.load returns it, dataset/data.py:114-133; C <= 4) -> out [N, C, H, W]:
 * per-channel scale (HOST float[C] or NULL; FlowTransform.scale_flow, dataset/transforms.py:842-849), HWC -> CHW
 * (np.transpose(flow, (2,0,1)), :850) and the bilinear resize to mask_size (models/rcf_model.py:438-442) in one pass.
 * No gradient: the RAFT flow is ground truth. */
RCF_API int rcf_flow_stage_hwc(const float* in, float* out, int N, int C, int h, int w, int H, int W, int align_corners,
                               const float* channel_scale_host, void* stream);

/* ---- caller-side mask preparation and mask losses ------------------------------------------------
 * Reference: models/rcf_model.py:433-434 (softmax over K, log-softmax OF the result), :376-378 (entropy loss),
 * :380-408 (PL / CRF positive/negative weighted MSE on the object channel), models/compactness_head.py:33-56
 * (compactness of one channel).  logits / masks / grad_masks / dlogits: dense [nframes, K, H*W] fp32 (the reference's
 * [B, I, K, H, W] with nframes = B*I); target: [nframes, H*W] (pl_masks / crf_masks).
 * Forward, one pass: masks = softmax over K; losses[0] = -(masks * log_softmax(masks)).sum(K).mean();
 * losses[1] = compactness of channel compact_channel (0 when off); losses[2] = pl_pos_weight * mean(max(t-m,0)^2) +
 * pl_neg_weight * mean(min(t-m,0)^2) with m = masks[:, pl_channel] and t = target (> pl_threshold when that is != -1);
 * losses[3] = get_sharpen_loss (:350-374): KL(log_softmax(masks) || sharpen(masks, T)) or the object-aware hinge.
 * frame_stats [nframes, 2] receives the per-frame centroid (needed by the backward when compactness is on).
 * Backward, one pass: dlogits = softmax-backward of (grad_masks + sum_i grad_losses[i] * dlosses[i]/dmasks); grad_masks
 * (the motion loss's mask gradient) and grad_losses (device float[4]) may each be NULL.  losses is float[4]. */
typedef struct RcfMaskCfg {
    int32_t nframes, K, H, W;
    int32_t compact_channel;   /* -1: off                                  (compactness_head.py:19-27)  */
    int32_t pl_channel;        /* object channel of the PL/CRF loss, -1: off (rcf_model.py:388, :403)    */
    float pl_threshold;        /* pl_mask_pos_th / crf_mask_pos_th; -1: target used as it is (:384, :399) */
    float pl_pos_weight;       /* pl_pos_weight / crf_pos_weight             (:393, :408)                 */
    float pl_neg_weight;
    int32_t sharpen_mode;      /* 0 off | 1 KL to sharpen(masks, T) (:370-373, utils/loss_utils.py:105-108) | 2 object-aware hinge (:362-369) */
    int32_t sharpen_channel;   /* object channel of mode 2 */
    float t_sharpen;           /* t_sharpen (:40, default 0.25) */
} RcfMaskCfg;
RCF_API int rcf_mask_prep_workspace_floats(int nframes, int P, size_t* nfloats);
RCF_API int rcf_mask_losses_forward(const RcfMaskCfg* cfg, const float* logits, const float* target, float* masks,
                                    float* losses, float* frame_stats, float* ws, void* stream);
RCF_API int rcf_mask_losses_backward(const RcfMaskCfg* cfg, const float* masks, const float* target, const float* grad_masks,
                                     const float* grad_losses, const float* frame_stats, float* dlogits, void* stream);

/* Measurement hook (bench.py): record the two caller-owned cudaEvent_t handles immediately before and
 * after the launch of streaming kernel `which` in the following rcf_forward / rcf_backward calls of this
 * process (on the stream those calls are given).  which: 0 off, 1 k_moments, 2 k_loss, 3 k_bwd,
 * 4 k_pool, 5 k_pool_bwd.  Has no effect on results. */
#define RCF_TIME_OFF 0
#define RCF_TIME_MOMENTS 1
#define RCF_TIME_LOSS 2
#define RCF_TIME_BWD 3
#define RCF_TIME_POOL 4
#define RCF_TIME_POOL_BWD 5
RCF_API int rcf_debug_time_kernel(int which, void* start_event, void* stop_event);

/* Implementation switches for A/B measurements (process-wide; results are bit-identical either way).
 * (options 1 and 2 belonged to a single-launch forward experiment that was measured slower and removed) */
/* These switches are process-wide measurement aids.  Do not change RCF_OPT_SINGLE_PASS between an rcf_forward and the
 * rcf_backward that consumes its ctx: the backward would look for coefficients the forward did not write. */
#define RCF_OPT_L2_HINTS 3       /* evict-first streaming loads of flow/residual in pass 2 (default 1) */
#define RCF_OPT_SINGLE_PASS 4    /* theta_mode 0 with D == 0: skip pass 1, S_k is accumulated inside pass 2 (default 1) */
#define RCF_OPT_PDL 5            /* programmatic dependent launch between the library's consecutive kernels (default 1) */
#define RCF_OPT_CONV64_PAIR 7    /* 1 (default): conv fprop / data gradient on CTA pairs (cta_group::2); 0: one-CTA kernel */
#define RCF_OPT_CONV64_DEBUG 6   /* measurement only, INVALID results: bit 0 no epilogue stores, bit 1 no producer loads, bit 2 no MMAs */
RCF_API int rcf_debug_set_option(int option, int value);

#ifdef __cplusplus
}
#endif
#endif /* RCF_LOSS_H_ */
