"""Times the tcgen05 conv (fprop / data gradient) against cuDNN on the same box.  python tools/time_conv64.py [N H W]"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from rcf_unsupvideoseg_b200 import conv64 as c64  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    shapes = [(4, 480, 854), (16, 96, 96), (16, 48, 48)] if len([a for a in sys.argv[1:] if a.isdigit()]) < 3 else [tuple(int(a) for a in sys.argv[1:4])]
    for N, H, W in shapes:
        x = torch.randn(N, 64, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
        w = torch.randn(64, 64, 3, 3, device="cuda") / 24
        wp = c64.pack_weights(w, False)
        flops = 2 * N * H * W * 64 * 64 * 9
        row = [f"{N}x64x{H}x{W}"]
        from rcf_unsupvideoseg_b200 import _lib as _l
        for pair in (1, 0):
            _l.load_library().rcf_debug_set_option(7, pair)
            for nprod in (1, 2, 3):
                hi, lo = c64.split_bf16(x)
                t = timeit(lambda: c64.conv64_pair(hi, lo, wp, nprod))
                row.append(f"{'pair' if pair else 'one-CTA'} nprod{nprod} {t:8.1f} us ({flops / t / 1e6:6.1f} TFLOP/s)")
        _l.load_library().rcf_debug_set_option(7, 1)
        if "--debug" in sys.argv:
            from rcf_unsupvideoseg_b200 import _lib
            lib = _lib.load_library()
            for dbg in (1, 4, 5):
                lib.rcf_debug_set_option(6, dbg)
                row.append(f"dbg{dbg} nprod2 {timeit(lambda: c64.conv64_pair(hi, lo, wp, 2)):8.1f} us")
            lib.rcf_debug_set_option(6, 0)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            t = timeit(lambda: F.conv2d(x, w, None, 1, 1))
            row.append(f"cudnn {'tf32' if tf32 else 'fp32'} {t:8.1f} us")
        gy = torch.randn_like(x)
        g_hi, g_lo = c64.split_bf16(gy)
        for nprod in (1, 2, 3):
            t = timeit(lambda: c64.conv64_wgrad_pair(hi, lo, g_hi, g_lo, nprod))
            row.append(f"wgrad nprod{nprod} {t:8.1f} us")
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            t = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (False, True, False)))
            row.append(f"cudnn wgrad {'tf32' if tf32 else 'fp32'} {t:8.1f} us")
            t = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (True, False, False)))
            row.append(f"cudnn dgrad {'tf32' if tf32 else 'fp32'} {t:8.1f} us")
        t = timeit(lambda: c64.pack_weights(w, False))
        row.append(f"pack {t:5.1f} us")
        print(" | ".join(row), flush=True)


if __name__ == "__main__":
    main()
