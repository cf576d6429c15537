"""Times the SPPF pooling kernel alone: batch 64, 20x20 (640^2) and batch 16, 40x40 (1280^2), c_ 256 in a 1024-wide concat
buffer; single launches between events and chained launches (launch overhead amortised).  TOD_POOL_TMA=0 selects the
LDG / STG kernels.  usage: pool_bench.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for B, h, w, c in ((64, 20, 20, 256), (16, 40, 40, 256)):
    buf = torch.randn((B, h, w, 4 * c), dtype=torch.float32).to(torch.bfloat16).cuda()
    def once():
        _lib.check(L.tod_sppf_pool_nhwc_bf16(buf.data_ptr(), B, h, w, c, 4 * c, st), "pool")
    ts = []
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); once(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        once()
    e1.record()
    torch.cuda.synchronize()
    chain = e0.elapsed_time(e1) * 1e3 / 20
    mb = B * h * w * c * 2 * 4 / 1e6
    t = float(np.median(ts[2:]))
    print(f"sppf pool {B}x{h}x{w}x{c} (TOD_POOL_TMA={os.environ.get('TOD_POOL_TMA', '1')}): single {t:.1f} us = {mb / t * 1e3:.0f} GB/s, "
          f"chained {chain:.1f} us = {mb / chain * 1e3:.0f} GB/s ({mb:.1f} MB algorithmic)")
