"""Per-layer conv micro-benchmark: every distinct conv shape of the detector (scale s, batch 64, 640x640 by default) through
the C ABI, for several kernel variants, timed with CUDA events and checked against the CUDA-core check kernel.

usage: conv_bench.py [--batch 64] [--size 640] [--scale s] [--variants v1,halo,halo_m1,halo_ring] [--reps 5]
                     [--filter substr] [--check] [--out gpurun_out/conv_bench.json]
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
from transparent_object_detection_b200 import synth                                        # noqa: E402
from transparent_object_detection_b200 import BaseModel         # noqa: E402
from transparent_object_detection_b200._lib import ConvDesc, check   # noqa: E402

VARIANTS = {
    "auto": dict(variant=0),
    "v1": dict(variant=1),
    "halo": dict(variant=2),
    "halo_m1": dict(variant=2, m=1),
    "halo_m2": dict(variant=2, m=2),
    "halo_ring": dict(variant=2, no_station=1),
    "halo_sa2": dict(variant=2, stages=2),
    "halo_n128": dict(variant=2, ncap=128),
    "halo_n128_m1": dict(variant=2, ncap=128, m=1),
    "halo_n64": dict(variant=2, ncap=64),
    "halo_bk32": dict(variant=2, bk=32),
    "halo_bk32_m1": dict(variant=2, bk=32, m=1),
    "halo_bk16": dict(variant=2, bk=16),
    "halo_n64_bk32": dict(variant=2, ncap=64, bk=32),
    "halo_n64_bk32_m1": dict(variant=2, ncap=64, bk=32, m=1),
    "halo_n128_bk32": dict(variant=2, ncap=128, bk=32),
    "halo_n128_bk32_m1": dict(variant=2, ncap=128, bk=32, m=1),
    "halo_n64_m1": dict(variant=2, ncap=64, m=1),
    "halo_ring_m1": dict(variant=2, no_station=1, m=1),
    "halo_ring_m2": dict(variant=2, no_station=1, m=2),
    "pair": dict(variant=2, pair=1),             # tcgen05 cta_group::2 (TOD_CONV_PAIR_ON)
    "pair_m1": dict(variant=2, pair=1, m=1),
    "pair_n128": dict(variant=2, pair=1, ncap=128),
    "pair_ring": dict(variant=2, pair=1, no_station=1),
    "nopair": dict(variant=2, pair=-1),
    "halo_noact": dict(variant=2, act=0),        # timing experiments only (results differ by construction)
    "halo_nores": dict(variant=2, nores=1),
}


def clone_desc(d: ConvDesc, **kw) -> ConvDesc:
    n = ConvDesc()
    C.memmove(C.byref(n), C.byref(d), C.sizeof(ConvDesc))
    n.reserved[0] = kw.get("variant", 0)
    n.reserved[1] = kw.get("m", 0)
    n.reserved[2] = kw.get("no_station", 0)
    n.reserved[3] = kw.get("ncap", 0)
    n.num_stages = kw.get("stages", 0)
    if kw.get("pair"):
        n.flags = (n.flags & ~24) | (8 if kw["pair"] > 0 else 16)
    if "bk" in kw:
        n.block_k = kw["bk"]
    if "act" in kw:
        n.act = kw["act"]
    if kw.get("nores"):
        n.d_residual = None
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--scale", default="s")
    ap.add_argument("--variants", default="v1,halo")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--filter", default="")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default="gpurun_out/conv_bench.json")
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
    names = a.variants.split(",")
    rows = []
    totals = {n: 0.0 for n in names}
    for kind, name, payload in eng.ops:
        if kind != "conv" or (a.filter and not any(f in name for f in a.filter.split(","))):
            continue
        dd = payload
        ho, wo = dd.hin // dd.stride, dd.win // dd.stride
        gflop = 2.0 * a.batch * ho * wo * dd.cout * dd.cin * dd.ksize ** 2 / 1e9
        esize = 4 if dd.out_dtype == 1 else 2
        mbytes = (a.batch * dd.hin * dd.win * dd.cin * 2 + a.batch * ho * wo * dd.cout * esize) / 1e6
        row = {"op": name, "shape": f"{dd.cin}->{dd.cout} k{dd.ksize} s{dd.stride} @{ho}x{wo}", "gflop": gflop, "mbytes": mbytes}
        ref = None
        out_t = eng.conv_meta[name]["dst"].buf
        for vn in names:
            dv = clone_desc(dd, **VARIANTS[vn])
            try:
                for _ in range(2):
                    check(L.tod_conv2d_nhwc_bf16(C.byref(dv), st), name)
                torch.cuda.synchronize()
            except Exception as e:  # keep going: a failing variant is a result too
                row[vn] = None
                row[vn + "_err"] = str(e)[:200]
                print(name, vn, "FAILED", str(e)[:200], flush=True)
                continue
            ts = []
            for _ in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                check(L.tod_conv2d_nhwc_bf16(C.byref(dv), st), name)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            row[vn] = ms * 1e3
            totals[vn] += ms
            if a.check:
                got = out_t.clone()
                if ref is None:
                    check(L.tod_conv2d_nhwc_bf16_simt_check(C.byref(dd), st), name)
                    torch.cuda.synchronize()
                    ref = out_t.clone()
                    check(L.tod_conv2d_nhwc_bf16(C.byref(dv), st), name)
                    torch.cuda.synchronize()
                    got = out_t.clone()
                err = (got.float() - ref.float()).abs()
                tol = 2e-2 + 2e-2 * ref.float().abs()
                row[vn + "_bad"] = float((err > tol).float().mean())
                row[vn + "_maxerr"] = float(err.max())
        rows.append(row)
        msg = f"{name:28s} {row['shape']:30s} {gflop:6.1f} GF {mbytes:6.0f} MB"
        for vn in names:
            if row.get(vn) is not None:
                msg += f" | {vn} {row[vn]:7.1f}us {gflop / row[vn] * 1e3:6.0f}TF {mbytes / row[vn]:5.2f}TB/s"
                if a.check:
                    msg += f" bad={row[vn + '_bad']:.1e}"
        print(msg, flush=True)
    print("TOTAL ms:", {k: round(v, 3) for k, v in totals.items()}, flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump({"args": vars(a), "rows": rows, "totals_ms": totals}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
