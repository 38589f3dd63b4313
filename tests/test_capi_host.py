"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, the ctypes mirrors match the C structs byte for byte, and argument validation works
without touching a GPU (no kernel is launched here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rcf_loss.h")


@pytest.fixture(scope="module")
def lib():
    import rcf_unsupvideoseg_b200 as pkg
    return pkg.load_library()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"RCF_API\s+[\w\s\*]+?\b(rcf_\w+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from rcf_unsupvideoseg_b200 import _lib
    syms = declared_symbols()
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(syms) <= exported
    assert all(e.startswith("rcf_") for e in exported), exported   # nothing else leaks out of the .so


def test_ctypes_structs_match_c_layout():
    from rcf_unsupvideoseg_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "rcf_loss.h"
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("RcfDesc %zu\nRcfInputs %zu\nRcfVisOut %zu\nRcfGrads %zu\nRcfMaskCfg %zu\n", sizeof(RcfDesc), sizeof(RcfInputs), sizeof(RcfVisOut), sizeof(RcfGrads), sizeof(RcfMaskCfg));
  F(RcfDesc,grad_loss_total);
  F(RcfMaskCfg,nframes); F(RcfMaskCfg,K); F(RcfMaskCfg,H); F(RcfMaskCfg,W); F(RcfMaskCfg,compact_channel); F(RcfMaskCfg,pl_channel);
  F(RcfMaskCfg,pl_threshold); F(RcfMaskCfg,pl_pos_weight); F(RcfMaskCfg,pl_neg_weight);
  F(RcfMaskCfg,sharpen_mode); F(RcfMaskCfg,sharpen_channel); F(RcfMaskCfg,t_sharpen);
  F(RcfDesc,B); F(RcfDesc,K); F(RcfDesc,H); F(RcfDesc,W); F(RcfDesc,Cf); F(RcfDesc,D); F(RcfDesc,ndir); F(RcfDesc,theta_mode);
  F(RcfDesc,robust); F(RcfDesc,unbounded_residual); F(RcfDesc,eps); F(RcfDesc,q); F(RcfDesc,resid_scale); F(RcfDesc,pred_div);
  F(RcfDesc,clamp_t); F(RcfDesc,inv_n); F(RcfDesc,mask_bstride); F(RcfDesc,flow_bstride); F(RcfDesc,resid_bstride);
  F(RcfDesc,feat_bstride); F(RcfDesc,dmask_bstride); F(RcfDesc,dresid_bstride); F(RcfDesc,dfeat_bstride);
  F(RcfDesc,vis_bstride); F(RcfDesc,vis_dstride); F(RcfDesc,vis_scale); F(RcfDesc,feat_lrelu_slope); F(RcfDesc,feat_nhwc); F(RcfDesc,dfeat_f16);
  F(RcfInputs,mask); F(RcfInputs,flow); F(RcfInputs,resid); F(RcfInputs,feat); F(RcfInputs,theta); F(RcfInputs,w1); F(RcfInputs,b1); F(RcfInputs,w2); F(RcfInputs,b2); F(RcfInputs,feat_bias);
  F(RcfVisOut,gt); F(RcfVisOut,pred); F(RcfVisOut,agg); F(RcfVisOut,res); F(RcfVisOut,aff);
  printf("RcfHeadBuffers %zu\n", sizeof(RcfHeadBuffers)); F(RcfHeadBuffers,a_hi); F(RcfHeadBuffers,feat); F(RcfHeadBuffers,g_hi); F(RcfHeadBuffers,d_cw2); F(RcfHeadBuffers,resid_up); F(RcfHeadBuffers,dresid_up);
  F(RcfGrads,dmask); F(RcfGrads,dresid); F(RcfGrads,dfeat); F(RcfGrads,dtheta); F(RcfGrads,dw1); F(RcfGrads,db1); F(RcfGrads,dw2); F(RcfGrads,db2); F(RcfGrads,dfeat_bias); F(RcfGrads,dfeat_hi); F(RcfGrads,dfeat_lo);
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "layout.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "layout")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    structs = {"RcfDesc": _lib.RcfDesc, "RcfInputs": _lib.RcfInputs, "RcfVisOut": _lib.RcfVisOut, "RcfGrads": _lib.RcfGrads,
               "RcfMaskCfg": _lib.RcfMaskCfg, "RcfHeadBuffers": _lib.RcfHeadBuffers}
    for line in out.splitlines():
        name, val = line.split()
        if "." in name:
            s, f = name.split(".")
            assert getattr(structs[s], f).offset == int(val), name
        else:
            assert C.sizeof(structs[name]) == int(val), name


def test_query_sizes_and_validation(lib):
    from rcf_unsupvideoseg_b200 import _lib
    d = _lib.RcfDesc()
    d.B, d.K, d.H, d.W, d.Cf, d.D, d.ndir, d.theta_mode = 16, 4, 480, 854, 0, 0, 2, 0
    d.pred_div = 10.0
    ctx, ws = C.c_size_t(), C.c_size_t()
    assert lib.rcf_query_sizes(C.byref(d), C.byref(ctx), C.byref(ws)) == 0
    assert 0 < ctx.value < 1 << 20 and 0 < ws.value < 64 << 20      # O(B*K) state, never per-pixel
    d.K = 9
    assert lib.rcf_query_sizes(C.byref(d), C.byref(ctx), C.byref(ws)) == -3
    assert b"not compiled" in lib.rcf_error_string(-3)
    d.K, d.D = 4, 3
    assert lib.rcf_query_sizes(C.byref(d), C.byref(ctx), C.byref(ws)) == -2
    d.D, d.theta_mode = 0, 1                                          # MLP mode without features
    assert lib.rcf_query_sizes(C.byref(d), C.byref(ctx), C.byref(ws)) == -5
    d.theta_mode = 0
    inp = _lib.RcfInputs()                                            # all NULL -> rejected before any launch
    assert lib.rcf_forward(C.byref(d), C.byref(inp), None, None, None, None, None) == -1
    assert lib.rcf_backward(C.byref(d), C.byref(inp), None, None, None, None, None) == -1
    assert lib.rcf_abi_version() == _lib.RCF_ABI_VERSION


def test_head_entry_points_validate_before_any_launch(lib):
    """rcf_head_forward / rcf_head_backward (the default head as one call each way) reject NULL / wrong-mode arguments with
    status codes on the host, without touching a device."""
    from rcf_unsupvideoseg_b200 import _lib
    d = _lib.RcfDesc()
    d.B, d.K, d.H, d.W, d.Cf, d.D, d.ndir, d.theta_mode, d.feat_nhwc = 2, 4, 24, 32, 64, 0, 2, 1, 1
    inp, hb, g = _lib.RcfInputs(), _lib.RcfHeadBuffers(), _lib.RcfGrads()
    assert lib.rcf_head_forward(None, C.byref(inp), None, None, None, 3, 0.1, 2, 0, 0, C.byref(hb), None, None, None, None, None) == -1
    assert lib.rcf_head_forward(C.byref(d), C.byref(inp), None, None, None, 3, 0.1, 2, 0, 0, C.byref(hb), None, None, None, None, None) == -2   # RCF_ERR_SHAPE: feat_bstride != H*W*64
    d.feat_bstride[0] = d.feat_bstride[1] = 24 * 32 * 64
    assert lib.rcf_head_forward(C.byref(d), C.byref(inp), None, None, None, 3, 0.1, 2, 0, 0, C.byref(hb), None, None, None, None, None) == -1   # NULL weights / buffers
    d.Cf = 32
    assert lib.rcf_head_forward(C.byref(d), C.byref(inp), None, None, None, 3, 0.1, 2, 0, 0, C.byref(hb), None, None, None, None, None) == -5   # not the default head
    d.Cf = 64
    assert lib.rcf_head_backward(C.byref(d), C.byref(inp), None, None, None, None, 3, 0.1, 2, 1, 0, 0, C.byref(hb), None) == -1


def test_no_oracle_import_in_product():
    """The shipped package must never reach into oracle/ (or any CPU fallback)."""
    pkg_dir = os.path.join(ROOT, "rcf_unsupvideoseg_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "torch_port" not in txt and "rcf_oracle" not in txt, f
