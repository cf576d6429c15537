"""Fixed cost of a conv launch: every distinct conv of the detector at several batch sizes, timed (a) as one launch
between two events and (b) as `--chain` back-to-back launches of the same descriptor (programmatic dependent launch
overlaps prologue and tail).  t(batch) = fixed + slope * batch; the intercept is what a launch costs before any work.

usage: fixed_cost.py [--batches 1,4,16,64] [--filter substr,...] [--chain 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import BaseModel, synth         # noqa: E402
from transparent_object_detection_b200._lib import check               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--scale", default="s")
    ap.add_argument("--filter", default="")
    ap.add_argument("--chain", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/fixed_cost.json")
    a = ap.parse_args()
    C_, d, m = synth.SCALES[a.scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    res = {}
    batches = [int(b) for b in a.batches.split(",")]
    for B in batches:
        eng = model.engine(B, a.size, a.size)
        eng.x_static.copy_(torch.from_numpy(synth.make_images(B, a.size, a.size, seed=3)))
        eng.run_network()
        torch.cuda.synchronize()
        st = torch.cuda.current_stream().cuda_stream
        L = eng.L
        for kind, name, dd in eng.ops:
            if kind != "conv" or (a.filter and not any(f in name for f in a.filter.split(","))):
                continue
            def once():
                check(L.tod_conv2d_nhwc_bf16(C.byref(dd), st), name)
            for _ in range(3):
                once()
            torch.cuda.synchronize()
            single = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); once(); e1.record()
                torch.cuda.synchronize()
                single.append(e0.elapsed_time(e1) * 1e3)
            chain = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.chain):
                    once()
                e1.record()
                torch.cuda.synchronize()
                chain.append(e0.elapsed_time(e1) * 1e3 / a.chain)
            ho, wo = dd.hin // dd.stride, dd.win // dd.stride
            r = res.setdefault(name, {"shape": f"{dd.cin}->{dd.cout} k{dd.ksize} s{dd.stride} @{ho}x{wo}", "single": {}, "chain": {}})
            r["single"][B], r["chain"][B] = float(np.median(single)), float(np.median(chain))
        del eng
        torch.cuda.empty_cache()
    print(f"{'op':28s} {'shape':28s} " + " ".join(f"B={b:<3d} single/chain" for b in batches))
    for name, r in res.items():
        print(f"{name:28s} {r['shape']:28s} " + " ".join(f"{r['single'][b]:7.1f}/{r['chain'][b]:6.1f}  " for b in batches), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
