// rcf_capi.cu -- the extern "C" boundary declared in include/rcf_loss.h.
// Validates the descriptor, lays out ctx / ws, and enqueues the kernels on the caller's stream.
// No allocation, no synchronisation, no default-stream use.
#include <math.h>
#include <string.h>

#include "rcf_common.cuh"
#include "rcf_internal.h"

namespace {

int g_l2_hints = 1;        // RCF_OPT_L2_HINTS
int g_single_pass = 1;     // RCF_OPT_SINGLE_PASS
int g_pdl = 1;             // RCF_OPT_PDL

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

int validate_desc(const RcfDesc* d) {
    if (!d) return RCF_ERR_NULL;
    if (d->B < 1 || d->K < 1 || d->H < 1 || d->W < 1 || d->Cf < 0) return RCF_ERR_SHAPE;
    if ((long long)d->H * d->W > 0x7fffffffLL / 4) return RCF_ERR_SHAPE;
    if (d->ndir != 1 && d->ndir != 2) return RCF_ERR_SHAPE;
    if ((long long)d->ndir * d->B > 65535) return RCF_ERR_SHAPE;   // gridDim.y
    if (d->D != 0 && d->D != 2 && d->D != 5) return RCF_ERR_SHAPE;
    if (d->K > RCF_MAX_K || d->Cf > RCF_MAX_CF) return RCF_ERR_UNSUPPORTED;
    if (d->theta_mode != 0 && d->theta_mode != 1) return RCF_ERR_MODE;
    if (d->theta_mode == 1 && d->Cf < 1) return RCF_ERR_MODE;
    if (d->theta_mode == 0 && d->Cf != 0) return RCF_ERR_MODE;
    if (d->unbounded_residual && d->D != 0) return RCF_ERR_MODE;   // reference :279-286 only in free_residual
    if (!(d->pred_div != 0.0f)) return RCF_ERR_MODE;
    if (d->feat_nhwc && (d->theta_mode != 1 || d->Cf % 4 || d->Cf > 128 || 256 % (d->Cf / 4))) return RCF_ERR_MODE;
    return RCF_OK;
}

void fill_common(RcfK& a, const RcfDesc& d, const RcfInputs& in, const RcfLayout& L, void* ctx, void* ws) {
    memset(&a, 0, sizeof(a));
    a.B = d.B; a.K = d.K; a.H = d.H; a.W = d.W; a.P = d.H * d.W; a.Cf = d.Cf; a.D = d.D;
    a.ndir = d.ndir; a.nfd = L.nfd;
    a.robust = d.robust; a.unbounded = d.unbounded_residual; a.theta_mode = d.theta_mode;
    a.eps = d.eps; a.q = d.q;
    a.scale = d.unbounded_residual ? 1.0f : d.resid_scale;
    a.ex2_scale = (float)(2.0 * 1.4426950408889634 / (double)d.pred_div);
    a.dres_scale = d.unbounded_residual ? 1.0f : d.resid_scale / d.pred_div;
    a.clamp_t = d.clamp_t;
    a.inv_n = d.inv_n > 0.0f ? d.inv_n : (float)(1.0 / ((double)d.B * 2.0 * d.H * d.W));
    a.cy = 0.5f * (float)(d.H - 1); a.cx = 0.5f * (float)(d.W - 1);
    a.sy = 1.0f / fmaxf(a.cy, 1.0f); a.sx = 1.0f / fmaxf(a.cx, 1.0f);
    a.feat_slope = (d.feat_lrelu_slope > 0.0f) ? d.feat_lrelu_slope : 1.0f;
    a.feat_nhwc = d.feat_nhwc ? 1 : 0;
    for (int i = 0; i < 2; ++i) {
        a.mask[i] = in.mask[i]; a.flow[i] = in.flow[i]; a.resid[i] = in.resid[i];
        a.feat[i] = in.feat[i]; a.theta[i] = in.theta[i];
        a.mask_bs[i] = d.mask_bstride[i]; a.flow_bs[i] = d.flow_bstride[i];
        a.resid_bs[i] = d.resid_bstride[i]; a.feat_bs[i] = d.feat_bstride[i];
    }
    a.w1 = in.w1; a.b1 = in.b1; a.w2 = in.w2; a.b2 = in.b2;
    a.feat_bias = d.feat_nhwc ? in.feat_bias : nullptr;
    char* c = static_cast<char*>(ctx);
    a.segd = reinterpret_cast<double*>(c + L.c_segd);
    a.coef = reinterpret_cast<float*>(c + L.c_coef);
    a.mlp = reinterpret_cast<double*>(c + L.c_mlp);
    a.gm = reinterpret_cast<double*>(c + L.c_gm);
    char* w = static_cast<char*>(ws);
    a.part1 = reinterpret_cast<float*>(w + L.w_part1);
    a.partp = reinterpret_cast<float*>(w + L.w_partp);
    a.part2 = reinterpret_cast<float*>(w + L.w_part2);
    a.coefb = reinterpret_cast<float*>(w + L.w_coefb);
    a.gscale = reinterpret_cast<float*>(w + L.w_gscale);
    a.poolbar = reinterpret_cast<float*>(w + L.w_poolbar);
    a.dh = reinterpret_cast<double*>(w + L.w_dh);
    a.thbar = reinterpret_cast<double*>(w + L.w_thbar);
    a.dbpart = reinterpret_cast<float*>(w + L.w_dbpart);
    a.dbfd = reinterpret_cast<double*>(w + L.w_dbfd);
    a.poolsum = reinterpret_cast<double*>(w + L.w_poolsum);
    a.cnt = reinterpret_cast<unsigned int*>(w + L.w_cnt);
    a.gmax = reinterpret_cast<float*>(w + L.w_gmax);
    a.dfeat_f16 = 0;
    a.nblkpb = L.nblkpb; a.poolchunk = L.poolchunk; a.pooltp = L.pooltp;
    a.mlp_smem = (d.theta_mode == 1 && d.Cf % 4 == 0 && d.Cf <= 128 && aligned16(in.w1)) ? 1 : 0;
    a.pdl = g_pdl;
    a.l2_hints = g_l2_hints;   // pass 2 streams flow/residual evict-first so the masks of pass 1 survive in L2
    a.nchunk1 = L.nchunk1; a.nchunk2 = L.nchunk2; a.nchunkb = L.nchunkb; a.nchunkp = L.nchunkp;
    a.chunk2 = L.chunk2;
    a.pool_sums = L.pool_sums;
}

int validate_inputs(const RcfDesc& d, const RcfInputs& in) {
    for (int i = 0; i < d.ndir; ++i) {
        if (!in.mask[i] || !in.flow[i] || !in.resid[i]) return RCF_ERR_NULL;
        if (!aligned4(in.mask[i]) || !aligned4(in.flow[i]) || !aligned4(in.resid[i])) return RCF_ERR_ALIGN;
        if (d.theta_mode == 1) { if (!in.feat[i]) return RCF_ERR_NULL; if (!aligned4(in.feat[i])) return RCF_ERR_ALIGN; }
        else if (!in.theta[i]) return RCF_ERR_NULL;
    }
    if (d.theta_mode == 1 && (!in.w1 || !in.b1 || !in.w2 || !in.b2)) return RCF_ERR_NULL;
    if (in.feat_bias && !d.feat_nhwc) return RCF_ERR_MODE;
    if (d.theta_mode == 1 && d.feat_nhwc)      // the channels-last pooling kernels use 128-bit accesses unconditionally
        for (int i = 0; i < d.ndir; ++i)
            if (!aligned16(in.feat[i]) || d.feat_bstride[i] % 4) return RCF_ERR_ALIGN;
    return RCF_OK;
}

// 128-bit path: every plane start must be 16-byte aligned
bool vec_ok_inputs(const RcfDesc& d, const RcfInputs& in, bool with_feat) {
    if (((long long)d.H * d.W) % 4 != 0 || d.W < 4) return false;     // a pack crosses at most one line end (px_coords_rc)
    for (int i = 0; i < d.ndir; ++i) {
        if (!aligned16(in.mask[i]) || !aligned16(in.flow[i]) || !aligned16(in.resid[i])) return false;
        if (d.mask_bstride[i] % 4 || d.flow_bstride[i] % 4 || d.resid_bstride[i] % 4) return false;
        if (with_feat && (!aligned16(in.feat[i]) || d.feat_bstride[i] % 4)) return false;
    }
    return true;
}

struct TimeHook { int which = 0; cudaEvent_t start = nullptr, stop = nullptr; };
TimeHook g_hook;   // process-wide: autograd runs rcf_backward on its own device thread

struct ScopedTime {
    cudaStream_t s; bool on;
    ScopedTime(int which, cudaStream_t st) : s(st), on(g_hook.which == which && g_hook.start && g_hook.stop) {
        if (on) cudaEventRecord(g_hook.start, s);
    }
    ~ScopedTime() { if (on) cudaEventRecord(g_hook.stop, s); }
};

}  // namespace

int rcf_pdl_enabled() { return g_pdl; }

extern "C" int rcf_debug_time_kernel(int which, void* start_event, void* stop_event) {
    if (which < 0 || which > 5) return RCF_ERR_MODE;
    g_hook.which = which;
    g_hook.start = static_cast<cudaEvent_t>(start_event);
    g_hook.stop = static_cast<cudaEvent_t>(stop_event);
    return RCF_OK;
}

extern "C" void rcf_conv64_set_debug(int v);
extern "C" void rcf_conv64_set_pair(int v);
extern "C" int rcf_debug_set_option(int option, int value) {
    if (option == RCF_OPT_CONV64_DEBUG) { rcf_conv64_set_debug(value); return RCF_OK; }
    if (option == RCF_OPT_CONV64_PAIR) { rcf_conv64_set_pair(value); return RCF_OK; }
    if (option == RCF_OPT_L2_HINTS) { g_l2_hints = value ? 1 : 0; return RCF_OK; }
    if (option == RCF_OPT_SINGLE_PASS) { g_single_pass = value ? 1 : 0; return RCF_OK; }
    if (option == RCF_OPT_PDL) { g_pdl = value ? 1 : 0; return RCF_OK; }
    return RCF_ERR_MODE;
}

const float* rcf_ws_gmax(const RcfDesc* desc, const void* ws) {
    const RcfLayout L = rcf_make_layout(*desc);
    return reinterpret_cast<const float*>(static_cast<const unsigned char*>(ws) + L.w_gmax);
}

extern "C" int rcf_abi_version(void) { return RCF_ABI_VERSION; }

extern "C" const char* rcf_error_string(int code) {
    switch (code) {
        case RCF_OK: return "ok";
        case RCF_ERR_NULL: return "rcf: required pointer is NULL";
        case RCF_ERR_SHAPE: return "rcf: B/K/H/W/Cf/D/ndir out of range";
        case RCF_ERR_UNSUPPORTED: return "rcf: configuration not compiled into this library (K > RCF_MAX_K or Cf > RCF_MAX_CF)";
        case RCF_ERR_ALIGN: return "rcf: pointer misaligned or stride inconsistent";
        case RCF_ERR_MODE: return "rcf: inconsistent mode flags";
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "rcf: unknown error";
}

extern "C" int rcf_query_sizes(const RcfDesc* desc, size_t* ctx_bytes, size_t* ws_bytes) {
    const int v = validate_desc(desc);
    if (v != RCF_OK) return v;
    const RcfLayout L = rcf_make_layout(*desc);
    if (ctx_bytes) *ctx_bytes = L.c_bytes;
    if (ws_bytes) *ws_bytes = L.w_bytes;
    return RCF_OK;
}

#define RCF_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

extern "C" int rcf_forward(const RcfDesc* desc, const RcfInputs* in, float* loss, void* ctx, void* ws,
                           const RcfVisOut* vis, void* stream) {
    int v = validate_desc(desc);
    if (v != RCF_OK) return v;
    if (!in || !loss || !ctx || !ws) return RCF_ERR_NULL;
    v = validate_inputs(*desc, *in);
    if (v != RCF_OK) return v;
    const RcfLayout L = rcf_make_layout(*desc);
    RcfK a;
    fill_common(a, *desc, *in, L, ctx, ws);
    a.loss = loss;
    bool vec = vec_ok_inputs(*desc, *in, false);
    if (vis) {
        a.vis_gt = vis->gt; a.vis_pred = vis->pred; a.vis_agg = vis->agg; a.vis_res = vis->res;
        a.vis_aff = desc->D > 0 ? vis->aff : nullptr;
        a.vis_bs = desc->vis_bstride; a.vis_ds = desc->vis_dstride;
        a.vis_scale[0] = desc->vis_scale[0]; a.vis_scale[1] = desc->vis_scale[1];
        float* ptrs[5] = {a.vis_gt, a.vis_pred, a.vis_agg, a.vis_res, a.vis_aff};
        for (float* p : ptrs)
            if (p && (!aligned16(p) || desc->vis_bstride % 4 || desc->vis_dstride % 4)) vec = false;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    a.single_pass = (desc->theta_mode == 0 && desc->D == 0 && g_single_pass) ? 1 : 0;
    if (a.single_pass) {
        // theta supplied, no affine fit: nothing in pass 2 depends on pass 1, so the forward reads every input once; the
        // coefficient pack is just theta, which k_loss / k_bwd read directly (no k_segment_fwd launch either)
        { ScopedTime t(RCF_TIME_LOSS, s); RCF_CUDA(rcf_launch_loss(a, vec, s)); }
    } else {
        if (!a.pool_sums) { ScopedTime t(RCF_TIME_MOMENTS, s); RCF_CUDA(rcf_launch_moments(a, vec, s)); }
        if (desc->theta_mode == 1) { ScopedTime t(RCF_TIME_POOL, s); RCF_CUDA(rcf_launch_pool(a, vec_ok_inputs(*desc, *in, true), s)); }
        RCF_CUDA(rcf_launch_segment_fwd(a, s));
        { ScopedTime t(RCF_TIME_LOSS, s); RCF_CUDA(rcf_launch_loss(a, vec, s)); }
    }
    RCF_CUDA(rcf_launch_finalize(a, s));
    return RCF_OK;
}

extern "C" int rcf_backward(const RcfDesc* desc, const RcfInputs* in, const float* grad_loss, const void* ctx,
                            void* ws, const RcfGrads* grads, void* stream) {
    int v = validate_desc(desc);
    if (v != RCF_OK) return v;
    if (!in || !grad_loss || !ctx || !ws || !grads) return RCF_ERR_NULL;
    v = validate_inputs(*desc, *in);
    if (v != RCF_OK) return v;
    const RcfLayout L = rcf_make_layout(*desc);
    RcfK a;
    fill_common(a, *desc, *in, L, const_cast<void*>(ctx), ws);
    a.grad_loss = grad_loss;
    a.single_pass = (desc->theta_mode == 0 && desc->D == 0 && g_single_pass) ? 1 : 0;   // as in rcf_forward: coef = theta
    a.grad_total = desc->grad_loss_total ? 1 : 0;
    bool vec = vec_ok_inputs(*desc, *in, false);
    bool vec_pool = vec_ok_inputs(*desc, *in, true);
    bool any_dfeat = false;
    for (int i = 0; i < desc->ndir; ++i) {
        a.dmask[i] = grads->dmask[i]; a.dresid[i] = grads->dresid[i];
        a.dfeat[i] = desc->theta_mode == 1 ? grads->dfeat[i] : nullptr;
        a.dfeat_hi[i] = (desc->theta_mode == 1 && desc->feat_nhwc) ? static_cast<uint32_t*>(grads->dfeat_hi[i]) : nullptr;
        a.dfeat_lo[i] = (a.dfeat_hi[i] && !desc->dfeat_f16) ? static_cast<uint32_t*>(grads->dfeat_lo[i]) : nullptr;
        if (a.dfeat_hi[i] && desc->dfeat_f16) a.dfeat_f16 = 1;
        if (a.dfeat_hi[i] && (!aligned16(a.dfeat_hi[i]) || (a.dfeat_lo[i] && !aligned16(a.dfeat_lo[i])) || desc->dfeat_bstride[i] % 8))
            return RCF_ERR_ALIGN;
        if (a.dfeat_hi[i]) any_dfeat = true;
        a.dtheta[i] = desc->theta_mode == 0 ? grads->dtheta[i] : nullptr;
        a.dmask_bs[i] = desc->dmask_bstride[i]; a.dresid_bs[i] = desc->dresid_bstride[i];
        a.dfeat_bs[i] = desc->dfeat_bstride[i];
        if (a.dmask[i] && (!aligned16(a.dmask[i]) || a.dmask_bs[i] % 4)) { vec = false; vec_pool = false; }
        if (a.dresid[i] && (!aligned16(a.dresid[i]) || a.dresid_bs[i] % 4)) vec = false;
        if (a.dfeat[i] && (!aligned16(a.dfeat[i]) || a.dfeat_bs[i] % 4)) {
            if (desc->feat_nhwc) return RCF_ERR_ALIGN;      // channels-last dfeat is always written with 128-bit stores
            vec_pool = false;
        }
        if (a.dmask[i] && !aligned4(a.dmask[i])) return RCF_ERR_ALIGN;
        if (a.dresid[i] && !aligned4(a.dresid[i])) return RCF_ERR_ALIGN;
        if (a.dfeat[i]) any_dfeat = true;
    }
    a.dw1 = grads->dw1; a.db1 = grads->db1; a.dw2 = grads->dw2; a.db2 = grads->db2;
    a.dfeat_bias = (desc->feat_nhwc && in->feat_bias) ? grads->dfeat_bias : nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    RCF_CUDA(rcf_launch_segment_bwd(a, s));
    { ScopedTime t(RCF_TIME_BWD, s); RCF_CUDA(rcf_launch_bwd(a, vec, s)); }
    bool any_dmask = false;
    for (int i = 0; i < desc->ndir; ++i) any_dmask |= (a.dmask[i] != nullptr);
    if (desc->theta_mode == 1 && (any_dmask || any_dfeat || a.dfeat_bias)) {
        // adds the pooled-feature term onto dmask (read-modify-write) and writes dfeat
        ScopedTime t(RCF_TIME_POOL_BWD, s);
        RCF_CUDA(rcf_launch_pool_bwd(a, vec_pool, s));
    }
    return RCF_OK;
}
