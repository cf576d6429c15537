"""Run one warm pass of the engine, then the selected ops again (for ncu -k/-s selection).
usage: profile_ops.py [--batch 64] [--size 640] [--scale s] [--ops name1,name2,...|all] [--nms] [--decode]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import synth                                        # noqa: E402
from transparent_object_detection_b200 import BaseModel         # noqa: E402
from transparent_object_detection_b200._lib import check        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--scale", default="s")
    ap.add_argument("--ops", default="all")
    ap.add_argument("--nms", action="store_true")
    ap.add_argument("--decode", action="store_true")
    ap.add_argument("--fused", action="store_true", help="run the head's last convs with the fused decode epilogue")
    a = ap.parse_args()
    C_, d, m = synth.SCALES[a.scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    eng = model.engine(a.batch, a.size, a.size)
    eng.x_static.copy_(torch.from_numpy(synth.make_images(a.batch, a.size, a.size, seed=3)))
    eng.run_network(); eng.run_decode(False, False, True); eng.run_nms(0.05, 0.5)      # warm pass: fills every buffer
    torch.cuda.synchronize()
    print("warm pass launches:", eng.launches_per_pass, "convs:", sum(1 for k, _, _ in eng.ops if k == "conv"), flush=True)
    st = torch.cuda.current_stream().cuda_stream
    want = None if a.ops == "all" else set(a.ops.split(","))
    for kind, name, payload in eng.ops:
        if want is not None and name not in want:
            continue
        if kind == "conv":
            if a.fused and name in eng.head_fuse:
                check(eng.L.tod_conv2d_head_decode(C.byref(payload), C.byref(eng.head_fuse[name]), st), name)
            else:
                check(eng.L.tod_conv2d_nhwc_bf16(C.byref(payload), st), name)
        elif kind == "stem":
            w, b, out = payload
            check(eng.L.tod_stem_conv_nchw_f32(eng.x_static.data_ptr(), w.data_ptr(), b.data_ptr(), out.ptr, a.batch, a.size, a.size,
                                               C_, out.pitch, st), name)
        else:
            buf, c_ = payload
            check(eng.L.tod_sppf_pool_nhwc_bf16(buf.ptr, a.batch, buf.h, buf.w, c_, buf.pitch, st), name)
        print("ran", name, flush=True)
    if a.decode:
        eng.run_decode(False, False, True)
    if a.nms:
        eng.run_nms(0.05, 0.5)
    torch.cuda.synchronize()
    print("candidates/img (mean):", float((eng.cand_conf >= 0.05).sum(1).float().mean()), "kept/img:", float(eng.keep_count.float().mean()))


if __name__ == "__main__":
    main()
