"""Times the SPPF pooling kernel alone (batch 64, 20x20, c_ 256 in a 1024-wide concat buffer).  usage: pool_bench.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import _lib
L = _lib.lib()
B, h, w, c = 64, 20, 20, 256
buf = torch.randn((B, h, w, 4 * c), dtype=torch.float32).to(torch.bfloat16).cuda()
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(L.tod_sppf_pool_nhwc_bf16(buf.data_ptr(), B, h, w, c, 4 * c, st), "pool"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
mb = B * h * w * c * 2 * 4 / 1e6
print(f"sppf pool {B}x{h}x{w}x{c}: {np.median(ts[2:]):.1f} us, {mb:.1f} MB algorithmic, {mb / np.median(ts[2:]) / 1e3 * 1e3:.0f} GB/s")
