"""Second convolution of flow_feat_before_agg (reference :89-91) with its two backward GEMMs on two CUDA streams.

The convolution itself stays on cuDNN's sm_100 tensor-core kernels.  What this Function changes is the
SCHEDULE of its backward: autograd runs data-gradient and weight-gradient back to back on one stream, although they are
independent (both only need dG) and, at the 96x96 / 48x48 training shapes, neither fills the GPU (wgrad: 288 CTAs for 52 us,
dgrad + the stem's weight gradient that depends on it: ~57 us).  Here the weight gradient is issued on a side stream that
forks after dG is ready and joins before the gradients are handed back, so the two chains overlap -- in eager mode and,
because the fork/join is expressed with stream waits, inside captured CUDA graphs as parallel branches.
"""
from __future__ import annotations

import torch

_SIDE = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _SIDE.get(idx)
    if s is None:
        s = _SIDE[idx] = torch.cuda.Stream(device=idx)
    return s


class _ConvDualStreamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, padding, dilation, groups):
        ctx.conf = (tuple(stride), tuple(padding), tuple(dilation), int(groups))
        ctx.save_for_backward(x, weight)
        return torch.ops.aten.convolution(x, weight, None, ctx.conf[0], ctx.conf[1], ctx.conf[2], False, [0, 0], ctx.conf[3])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = gw = None
        if need_x and need_w and g.is_cuda:
            main = torch.cuda.current_stream(g.device)
            side = _side_stream(g.device)
            side.wait_stream(main)                     # dG (and everything before it) is ready
            with torch.cuda.stream(side):
                gw = torch.ops.aten.convolution_backward(g, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                         [False, True, False])[1]
            gx = torch.ops.aten.convolution_backward(g, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                     [True, False, False])[0]
            # the consumers of gx (LeakyReLU / stem backward) run on `main` while wgrad is still busy on `side`;
            # the join is deferred to the point where gw is handed to autograd
            main.wait_stream(side)
            gw.record_stream(main)
            g.record_stream(side)
        else:
            res = torch.ops.aten.convolution_backward(g, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                      [bool(need_x), bool(need_w), False])
            gx, gw = res[0], res[1]
        return gx, gw, None, None, None, None


def conv2d_dual_stream(x, weight, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    """F.conv2d(x, weight, None, ...) whose backward overlaps the data- and weight-gradient kernels on two streams."""
    return _ConvDualStreamFn.apply(x, weight, stride, padding, dilation, groups)
