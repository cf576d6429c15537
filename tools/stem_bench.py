"""Times the uint8 stem alone (batch 64, 640x640, cout 32) through the C ABI.  usage: stem_bench.py [--reps 10]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import _lib
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=10); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=640); ap.add_argument("--cout", type=int, default=32)
a = ap.parse_args()
L = _lib.lib()
B, H, W, C = a.batch, a.size, a.size, a.cout
g = torch.Generator().manual_seed(0)
u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).cuda()
w = (torch.randn((C, 27), generator=g) * 0.27).contiguous(); b = torch.randn((C,), generator=g) * 0.3
out = torch.zeros((B, H // 2, W // 2, C), dtype=torch.bfloat16).cuda()
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(a.reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(L.tod_stem_conv_nhwc_u8(u8.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, C, C, st), "stem"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[2:]))
mb = (u8.numel() + out.numel() * 2) / 1e6
pass
print(f"stem u8 {B}x{H}x{W} -> {C}: {ms * 1e3:.1f} us, {mb:.0f} MB algorithmic, {mb / ms / 1e3:.2f} TB/s")
