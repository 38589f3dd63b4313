// rcf_internal.h -- entry points shared between the library's own translation units (not exported, not part of the C ABI).
#pragma once
#include "rcf_loss.h"

// Address of the per-frame-direction gradient maxima inside the backward workspace `ws` of rcf_backward(desc, ...):
// nfd = ndir * B floats written by k_segment_bwd; rcf_grad_scale (rcf_common.cuh) turns them into the fp16 gradient scale.
const float* rcf_ws_gmax(const RcfDesc* desc, const void* ws);

// rcf_conv64_wgrad / rcf_stem_backward whose results are divided by rcf_grad_scale(gmax, nfd) in their final reduction
// (gmax == nullptr: no scaling, identical to the public entry points).
int rcf_conv64_wgrad_ex(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* ws, int nimg,
                        int H, int W, int nprod, const float* gmax, int nfd, void* stream);
int rcf_stem_backward_ex(const float* const* flow, const int64_t* flow_bstride, int ndir, int B, int H, int W, int Cf, int ks,
                         float clamp_t, float slope, const float* act, const uint32_t* sign, const float* dact, float* dw,
                         float* db, void* ws, int nprod, const float* gmax, int nfd, void* stream);
