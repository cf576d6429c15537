"""Per-role wait breakdown of the halo conv kernel (tod_debug_set_conv_profile): for each selected layer prints what
fraction of the kernel each role's issuing thread spent waiting, averaged over CTAs.

usage: conv_profile.py [--filter name-substring[,substr2]] [--batch 64] [--size 640] [--scale s] [--m M] [--no-station]
"""
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
from tools.conv_bench import clone_desc                         # noqa: E402

NAMES = ["prod.total", "prod.wait_Aempty", "prod.wait_Bempty", "mma.total", "mma.wait_tmem_empty", "mma.wait_Afull",
         "mma.wait_Bfull", "-", "epi0.total", "epi0.wait_tmem_full", "epi0.wait_stage_free", "epi0.wait_panel_written",
         "epi1.total", "epi1.wait_tmem_full", "epi1.wait_stage_free", "epi1.wait_panel_written"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--scale", default="s")
    ap.add_argument("--filter", default="")
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--no-station", action="store_true")
    ap.add_argument("--trace", type=int, default=0, help="print the first N entries of CTA 0's MMA-group timeline")
    ap.add_argument("--pair", type=int, default=0, help="1: CTA-pair kernel (TOD_CONV_PAIR_ON), -1: forced off")
    a = ap.parse_args()
    C_, d, m = synth.SCALES[a.scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    eng = model.engine(a.batch, a.size, a.size)
    eng.x_static.copy_(torch.from_numpy(synth.make_images(a.batch, a.size, a.size, seed=3)))
    eng.run_network()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    L = eng.L
    prof = torch.zeros((148 * 16 + 3 * 1024 + 8,), dtype=torch.int64, device="cuda")
    filt = [f for f in a.filter.split(",") if f]
    for kind, name, payload in eng.ops:
        if kind != "conv" or (filt and not any(f in name for f in filt)):
            continue
        dv = clone_desc(payload, variant=2, m=a.m, no_station=1 if a.no_station else 0, pair=a.pair)
        check(L.tod_conv2d_nhwc_bf16(C.byref(dv), st), name)      # warm
        torch.cuda.synchronize()
        prof.zero_()
        check(L.tod_debug_set_conv_profile(prof.data_ptr()), "set profile")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.tod_conv2d_nhwc_bf16(C.byref(dv), st), name)
        e1.record()
        torch.cuda.synchronize()
        check(L.tod_debug_set_conv_profile(None), "clear profile")
        ms = e0.elapsed_time(e1)
        raw = prof.cpu().numpy()
        pr = raw[:148 * 16].reshape(148, 16).astype(np.float64)
        used = pr[:, 3] > 0
        avg = pr[used].mean(0)
        dd = payload
        print(f"== {name}  {dd.cin}->{dd.cout} k{dd.ksize} s{dd.stride} @{dd.hin // dd.stride}x{dd.win // dd.stride}  {ms * 1e3:.1f} us,"
              f" {int(used.sum())} CTAs, mma role {avg[3]:.0f} cycles")
        tot = max(avg[3], 1.0)
        for i, nm in enumerate(NAMES):
            if nm == "-" or nm.endswith(".total"):
                continue
            print(f"   {nm:26s} {avg[i]:10.0f} cyc  {100 * avg[i] / tot:5.1f}% of mma-role time")
        if a.trace:
            n = int(raw[148 * 16 + 3 * 1024])
            tr = raw[148 * 16:148 * 16 + 3 * n].reshape(n, 3)
            print(f"   CTA 0 timeline, {n} groups: [wait start, ready, issue end] -> wait, issue, gap to next group")
            for i in range(min(n, a.trace)):
                nxt = tr[i + 1, 0] - tr[i, 2] if i + 1 < n else 0
                print(f"     g{i:3d} t={tr[i, 0]:8d} wait {tr[i, 1] - tr[i, 0]:6d} issue {tr[i, 2] - tr[i, 1]:6d} gap {nxt:6d}")


if __name__ == "__main__":
    main()
