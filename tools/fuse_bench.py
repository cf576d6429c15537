"""Times the last conv of every head tower unfused (f32 raw map out) and fused with its decode share.  usage: fuse_bench.py"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import synth
from transparent_object_detection_b200 import BaseModel
from transparent_object_detection_b200._lib import check
C_, d, m = synth.SCALES["s"]
model = BaseModel(80, C_, d, m).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
eng = model.engine(64, 640, 640)
x = torch.from_numpy(synth.make_images_u8(64, 640, 640, seed=3)).cuda()
eng.run_network(x); torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=7):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))
tot_u = tot_f = 0
for kind, name, payload in eng.ops:
    if name in eng.head_fuse:
        u = t(lambda: check(eng.L.tod_conv2d_nhwc_bf16(C.byref(payload), st), name))
        f = t(lambda: check(eng.L.tod_conv2d_head_decode(C.byref(payload), C.byref(eng.head_fuse[name]), st), name))
        tot_u += u; tot_f += f
        print(f"{name:16s} unfused {u:6.1f} us   fused {f:6.1f} us")
dec = t(lambda: eng.run_decode(False, False, True))
print(f"sum unfused {tot_u:.1f} + decode {dec:.1f} = {tot_u + dec:.1f} us   fused {tot_f:.1f} us")
