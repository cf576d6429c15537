"""Run the engine op by op and check every conv against an fp32 torch evaluation of the same op on the actual
contents of its input view -- localises wiring bugs to the first bad layer.  usage: gpu_netcheck.py scale B H W"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import synth                                        # noqa: E402
from transparent_object_detection_b200 import BaseModel         # noqa: E402
from transparent_object_detection_b200._lib import check        # noqa: E402


def main():
    scale, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    C_, d, m = synth.SCALES[scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    eng = model.engine(B, H, W)
    x = torch.from_numpy(synth.make_images(B, H, W, seed=2)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    nbad = 0
    for kind, name, payload in eng.ops:
        if kind == "stem":
            w, b, out = payload
            check(eng.L.tod_stem_conv_nchw_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.ptr, B, H, W, C_, out.pitch, st), name)
            torch.cuda.synchronize()
            want = F.silu(F.conv2d(x, w.view(C_, 3, 3, 3).cuda(), b.cuda(), stride=2, padding=1)).permute(0, 2, 3, 1)
            got = out.tensor().float()
        elif kind == "pool":
            buf, c_ = payload
            check(eng.L.tod_sppf_pool_nhwc_bf16(buf.ptr, B, buf.h, buf.w, c_, buf.pitch, st), name)
            torch.cuda.synchronize()
            t = buf.tensor().float().permute(0, 3, 1, 2)
            ys = [t[:, :c_]]
            for _ in range(3):
                ys.append(F.max_pool2d(ys[-1], 5, 1, 2))
            want = torch.cat(ys, 1).permute(0, 2, 3, 1)
            got = buf.tensor().float()
        else:
            meta = eng.conv_meta[name]
            src, dst = meta["src"], meta["dst"]
            xin = src.tensor().float().permute(0, 3, 1, 2)
            wq = meta["w"].cuda().to(torch.bfloat16).float()
            y = F.conv2d(xin, wq, None, stride=meta["stride"], padding=wq.shape[-1] // 2)
            if meta["b"] is not None:
                y = y + meta["b"].cuda().float().view(1, -1, 1, 1)
            if meta["upadd"] is not None:
                y = y + F.interpolate(meta["upadd"].permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
            if meta["act"] == 1:
                y = F.silu(y)
            if meta["residual"] is not None:
                y = y + meta["residual"].tensor().float().permute(0, 3, 1, 2)
            want = y.permute(0, 2, 3, 1)
            check(eng.L.tod_conv2d_nhwc_bf16(C.byref(payload), st), name)
            torch.cuda.synchronize()
            got = dst.tensor().float()
        err = (got - want).abs()
        tol = 3e-3 if (kind == "conv" and eng.conv_meta[name]["out_f32"]) else 3e-2
        bad = float((err > tol + tol * want.abs()).float().mean())
        flag = "OK  " if bad == 0 and not torch.isnan(got).any() else "FAIL"
        nbad += flag == "FAIL"
        extra = ""
        if kind == "conv":
            dd = payload
            extra = f" cin {dd.cin} cout {dd.cout} k{dd.ksize} s{dd.stride} {dd.hin}x{dd.win} xp {dd.x_pitch} op {dd.out_pitch}"
        print(f"{flag} {name:28s} max_err {float(err.max()):.4f} ref_absmax {float(want.abs().max()):.3f} bad {bad:.4f}{extra}", flush=True)
    print("bad ops:", nbad)


if __name__ == "__main__":
    main()
