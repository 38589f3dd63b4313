"""Pipeline timeline of CTA 0 of the tcgen05 conv (clock64 stamps per tile).  python tools/trace_conv64.py [nprod] [N H W]"""
import sys
import torch
sys.path.insert(0, ".")
from rcf_unsupvideoseg_b200 import _lib, conv64 as c64  # noqa: E402

nprod = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 480, 854)
lib = _lib.load_library()
x = torch.randn(N, 64, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn(64, 64, 3, 3, device="cuda") / 24
wp = c64.pack_weights(w, False)
hi, lo = c64.split_bf16(x)
for _ in range(3):
    c64.conv64_pair(hi, lo, wp, nprod)
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib.rcf_debug_conv64_trace(buf.data_ptr())
c64.conv64_pair(hi, lo, wp, nprod)
torch.cuda.synchronize()
lib.rcf_debug_conv64_trace(None)
t = buf.cpu().view(64, 8)
t0 = int(t[0, 3])
names = ["tile landed", "tmem free", "mma issued", "A buf free", "acc complete", "epi done", "(ns)", "ld done"]
n_ok = max(i for i in range(64) if int(t[i, 3]) != 0)
dclk, dns = int(t[n_ok, 3]) - int(t[0, 3]), int(t[n_ok, 6]) - int(t[0, 6])
print(f"clock64 rate during the kernel: {dclk / dns:.3f} GHz ({dclk} clk in {dns} ns over {n_ok} tiles)")
print(f"nprod {nprod}  {N}x64x{H}x{W}; clocks relative to the first TMA issue of CTA 0")
print("tile " + " ".join(f"{names[k]:>13}" for k in (0, 1, 2, 3, 4, 7, 5)))
for i in range(24):
    if int(t[i, 0]) == 0:
        break
    print(f"{i:4d} " + " ".join(f"{int(t[i, k]) - t0:13d}" for k in (0, 1, 2, 3, 4, 7, 5)))
