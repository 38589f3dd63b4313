"""ctypes binding of librcf_loss.so (include/rcf_loss.h).  No torch types cross this boundary.

The library is the product: if it cannot be loaded this module raises -- there is no CPU or
PyTorch fallback for the loss.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librcf_loss.so")

RCF_ABI_VERSION = 4
RCF_MAX_K = 8
RCF_MAX_CF = 256

_f32p = C.POINTER(C.c_float)
_i64x2 = C.c_int64 * 2
_ptr2 = C.c_void_p * 2


class RcfDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cf", C.c_int32), ("D", C.c_int32), ("ndir", C.c_int32), ("theta_mode", C.c_int32),
        ("robust", C.c_int32), ("unbounded_residual", C.c_int32),
        ("eps", C.c_float), ("q", C.c_float), ("resid_scale", C.c_float), ("pred_div", C.c_float),
        ("clamp_t", C.c_float), ("inv_n", C.c_float),
        ("mask_bstride", _i64x2), ("flow_bstride", _i64x2), ("resid_bstride", _i64x2), ("feat_bstride", _i64x2),
        ("dmask_bstride", _i64x2), ("dresid_bstride", _i64x2), ("dfeat_bstride", _i64x2),
        ("vis_bstride", C.c_int64), ("vis_dstride", C.c_int64), ("vis_scale", C.c_float * 2),
        ("feat_lrelu_slope", C.c_float), ("feat_nhwc", C.c_int32), ("grad_loss_total", C.c_int32),
        ("dfeat_f16", C.c_int32),
    ]


class RcfMaskCfg(C.Structure):
    _fields_ = [("nframes", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("compact_channel", C.c_int32), ("pl_channel", C.c_int32),
                ("pl_threshold", C.c_float), ("pl_pos_weight", C.c_float), ("pl_neg_weight", C.c_float),
                ("sharpen_mode", C.c_int32), ("sharpen_channel", C.c_int32), ("t_sharpen", C.c_float)]


class RcfInputs(C.Structure):
    _fields_ = [
        ("mask", _ptr2), ("flow", _ptr2), ("resid", _ptr2), ("feat", _ptr2), ("theta", _ptr2),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("feat_bias", C.c_void_p),
    ]


class RcfVisOut(C.Structure):
    _fields_ = [("gt", C.c_void_p), ("pred", C.c_void_p), ("agg", C.c_void_p), ("res", C.c_void_p),
                ("aff", C.c_void_p)]


class RcfGrads(C.Structure):
    _fields_ = [
        ("dmask", _ptr2), ("dresid", _ptr2), ("dfeat", _ptr2), ("dtheta", _ptr2),
        ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p), ("dfeat_bias", C.c_void_p),
        ("dfeat_hi", _ptr2), ("dfeat_lo", _ptr2),
    ]


class RcfHeadBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("a_hi", "a_lo", "sign", "wpack", "feat", "g_hi", "g_lo", "d_a1", "wgrad_ws", "stem_ws",
                                          "d_cw1", "d_cb1", "d_cw2", "resid_up", "dresid_up")]


EXPORTED_SYMBOLS = ("rcf_abi_version", "rcf_error_string", "rcf_query_sizes", "rcf_forward", "rcf_backward",
                    "rcf_debug_time_kernel", "rcf_flow_warp_forward", "rcf_flow_warp_backward", "rcf_corresponding_map",
                    "rcf_debug_set_option", "rcf_stem_forward", "rcf_stem_workspace_bytes", "rcf_stem_backward",
                    "rcf_resize_bilinear_forward", "rcf_resize_bilinear_backward", "rcf_flow_stage_hwc",
                    "rcf_mask_prep_workspace_floats", "rcf_mask_losses_forward", "rcf_mask_losses_backward",
                    "rcf_conv64_pack_weights", "rcf_conv64_forward", "rcf_split_bf16", "rcf_debug_conv64_status", "rcf_stem_forward_bf16", "rcf_conv64_wgrad_workspace_bytes", "rcf_conv64_wgrad",
                    "rcf_debug_conv64_trace", "rcf_head_forward", "rcf_head_backward")

_lib = None
_lock = threading.Lock()


class RcfLibraryError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True):
    """dlopen librcf_loss.so; optionally build it with nvcc first.  Raises RcfLibraryError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RcfLibraryError(f"{LIB_PATH} not built; run `python -m rcf_unsupvideoseg_b200.build`")
            from . import build as _build
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RcfLibraryError(f"cannot build {LIB_PATH}: {e}. The RCF loss has no CPU fallback.") from e
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:
            raise RcfLibraryError(f"cannot load {LIB_PATH}: {e}. The RCF loss has no CPU fallback.") from e
        lib.rcf_abi_version.restype = C.c_int
        lib.rcf_abi_version.argtypes = []
        lib.rcf_error_string.restype = C.c_char_p
        lib.rcf_error_string.argtypes = [C.c_int]
        lib.rcf_query_sizes.restype = C.c_int
        lib.rcf_query_sizes.argtypes = [C.POINTER(RcfDesc), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        lib.rcf_forward.restype = C.c_int
        lib.rcf_forward.argtypes = [C.POINTER(RcfDesc), C.POINTER(RcfInputs), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(RcfVisOut), C.c_void_p]
        lib.rcf_backward.restype = C.c_int
        lib.rcf_backward.argtypes = [C.POINTER(RcfDesc), C.POINTER(RcfInputs), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(RcfGrads), C.c_void_p]
        lib.rcf_debug_time_kernel.restype = C.c_int
        lib.rcf_debug_time_kernel.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        lib.rcf_flow_warp_forward.restype = C.c_int
        lib.rcf_flow_warp_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p]
        lib.rcf_flow_warp_backward.restype = C.c_int
        lib.rcf_flow_warp_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.rcf_corresponding_map.restype = C.c_int
        lib.rcf_corresponding_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.rcf_stem_forward.restype = C.c_int
        lib.rcf_stem_forward.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        lib.rcf_stem_forward_bf16.restype = C.c_int
        lib.rcf_stem_forward_bf16.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_void_p]
        lib.rcf_stem_workspace_bytes.restype = C.c_int
        lib.rcf_stem_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        lib.rcf_stem_backward.restype = C.c_int
        lib.rcf_stem_backward.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        for fn in (lib.rcf_resize_bilinear_forward, lib.rcf_resize_bilinear_backward):
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_int, C.c_int, C.c_void_p]
        lib.rcf_flow_stage_hwc.restype = C.c_int
        lib.rcf_flow_stage_hwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(C.c_float), C.c_void_p]
        lib.rcf_mask_prep_workspace_floats.restype = C.c_int
        lib.rcf_mask_prep_workspace_floats.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        lib.rcf_mask_losses_forward.restype = C.c_int
        lib.rcf_mask_losses_forward.argtypes = [C.POINTER(RcfMaskCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rcf_mask_losses_backward.restype = C.c_int
        lib.rcf_mask_losses_backward.argtypes = [C.POINTER(RcfMaskCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rcf_conv64_pack_weights.restype = C.c_int
        lib.rcf_conv64_pack_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.rcf_conv64_forward.restype = C.c_int
        lib.rcf_conv64_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p]
        lib.rcf_conv64_wgrad_workspace_bytes.restype = C.c_int
        lib.rcf_conv64_wgrad_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        lib.rcf_conv64_wgrad.restype = C.c_int
        lib.rcf_conv64_wgrad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p]
        lib.rcf_head_forward.restype = C.c_int
        lib.rcf_head_forward.argtypes = [C.POINTER(RcfDesc), C.POINTER(RcfInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(RcfHeadBuffers), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rcf_head_backward.restype = C.c_int
        lib.rcf_head_backward.argtypes = [C.POINTER(RcfDesc), C.POINTER(RcfInputs), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(RcfGrads), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(RcfHeadBuffers), C.c_void_p]
        lib.rcf_split_bf16.restype = C.c_int
        lib.rcf_split_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.rcf_debug_conv64_trace.restype = C.c_int
        lib.rcf_debug_conv64_trace.argtypes = [C.c_void_p]
        lib.rcf_debug_conv64_status.restype = C.c_int
        lib.rcf_debug_conv64_status.argtypes = []
        lib.rcf_debug_set_option.restype = C.c_int
        lib.rcf_debug_set_option.argtypes = [C.c_int, C.c_int]
        if lib.rcf_abi_version() != RCF_ABI_VERSION:
            raise RcfLibraryError(f"ABI mismatch: library {lib.rcf_abi_version()} vs binding {RCF_ABI_VERSION}")
        if os.environ.get("RCF_PDL", "1") == "0":      # A/B switch for measurements (results are identical either way)
            lib.rcf_debug_set_option(5, 0)
        _lib = lib
    return _lib


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(dev):
    """`with torch.cuda.device(dev)` only when dev is not already current: the context manager costs ~10 us of host time
    per use, and a step of the drop-in head enters it 8 times (launch-bound at the 96x96 / 48x48 training shapes)."""
    import torch
    cur = torch._C._cuda_getDevice()
    idx = dev.index if dev.index is not None else cur
    return _NO_GUARD if cur == idx else torch.cuda.device(idx)


def raw_stream(dev) -> int:
    """cudaStream_t of torch's current stream on `dev` as an integer.  `_lib.raw_stream(dev)` builds a
    Stream object per call (~9 us; ten of them per step of the head at the training shapes); this is one C call."""
    import torch
    idx = dev.index if dev.index is not None else torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(idx)


def check(code: int, what: str):
    if code != 0:
        msg = load_library().rcf_error_string(code).decode()
        raise RuntimeError(f"{what} failed with code {code}: {msg}")
