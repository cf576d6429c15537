"""SM clock under a sustained stream of ONE conv layer (tod_debug_set_timeline records the SM cycle counter and %globaltimer
at every CTA exit): a tensor-bound layer against an HBM-bound one tells whether the board's power cap, not the kernel,
sets the clock -- and with it how much a fuller tensor pipe can buy.
usage: clock_probe.py [--ops head.cls.0.2,backbone.dark2.1.cv2,backbone.dark4.1.m.0.cv1] [--seconds 0.6]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import BaseModel, synth         # noqa: E402
from transparent_object_detection_b200._lib import check               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="head.cls.0.2,backbone.dark2.1.cv2,backbone.dark4.1.m.0.cv1,backbone.dark3.1.m.0.cv1")
    ap.add_argument("--seconds", type=float, default=1.5)
    a = ap.parse_args()
    C_, d, m = synth.SCALES["s"]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    eng = model.engine(64, 640, 640)
    eng.x_static.copy_(torch.from_numpy(synth.make_images(64, 640, 640, seed=3)))
    eng.run_network()
    torch.cuda.synchronize()
    L, st = eng.L, torch.cuda.current_stream().cuda_stream
    cap = 1 << 21
    buf = torch.zeros(2 + 12 * cap, dtype=torch.int64, device="cuda")
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    for name in a.ops.split(","):
        dd = next(p for k, n, p in eng.ops if n == name)
        check(L.tod_conv2d_nhwc_bf16(C.byref(dd), st), name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); check(L.tod_conv2d_nhwc_bf16(C.byref(dd), st), name); e1.record(); torch.cuda.synchronize()
        n = max(50, int(a.seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
        buf.zero_(); buf[1] = cap
        check(L.tod_debug_set_timeline(buf.data_ptr()), "timeline")
        pw, stop = [], [False]

        def sample():                      # board power while the GPU works (the host enqueues far ahead of the device)
            while not stop[0]:
                pw.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                time.sleep(0.05)
        import threading
        th = threading.Thread(target=sample, daemon=True)
        e0.record()
        for i in range(n):
            check(L.tod_conv2d_nhwc_bf16(C.byref(dd), st), name)
        e1.record()
        th.start()
        torch.cuda.synchronize()
        stop[0] = True
        th.join()
        check(L.tod_debug_set_timeline(None), "timeline off")
        cnt = min(int(buf[0]), cap)
        rec = buf[2:2 + 12 * cnt].view(-1, 12).cpu().numpy()
        sm = (rec[:, 1] & 0xffff).astype(np.int64)
        cyc = (rec[:, 1].astype(np.uint64) >> np.uint64(16)).astype(np.int64)
        t1 = rec[:, 3].astype(np.int64)
        ghz_first, ghz_last = [], []
        for s_ in range(int(sm.max()) + 1):
            sel = np.nonzero(sm == s_)[0]
            o = sel[np.argsort(t1[sel])]
            if len(o) < 20:
                continue
            q = len(o) // 4
            ghz_first.append((cyc[o[q]] - cyc[o[0]]) / max(t1[o[q]] - t1[o[0]], 1))
            ghz_last.append((cyc[o[-1]] - cyc[o[-q]]) / max(t1[o[-1]] - t1[o[-q]], 1))
        ms = e0.elapsed_time(e1) / n
        ho, wo = dd.hin // dd.stride, dd.win // dd.stride
        gflop = 2.0 * 64 * ho * wo * dd.cout * dd.cin * dd.ksize ** 2 / 1e9
        print(f"{name:28s} {dd.cin}->{dd.cout} k{dd.ksize} @{ho}x{wo}: {n} launches, {ms * 1e3:.1f} us each = {gflop / ms:.0f} TFLOP/s; SM clock "
              f"first quarter {np.median(ghz_first):.3f} GHz, last quarter {np.median(ghz_last):.3f} GHz; board power "
              f"median {np.median(pw) if pw else float('nan'):.0f} W, max {np.max(pw) if pw else float('nan'):.0f} W ({len(pw)} samples)", flush=True)


if __name__ == "__main__":
    main()
