"""Device letterbox throughput (CUDA events) next to Pillow on the host cores.
usage: letterbox_bench.py [--n 64] [--src 640x480] [--dst 640x640] [--reps 20]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200.preprocess import Letterbox, letterbox_geometry   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--src", default="640x480")
    ap.add_argument("--dst", default="640x640")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    iw, ih = map(int, a.src.split("x"))
    w, h = map(int, a.dst.split("x"))
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (a.n, ih, iw, 3), dtype=np.uint8)
    src = torch.from_numpy(imgs).cuda()
    out = torch.zeros((a.n, h, w, 3), dtype=torch.uint8, device="cuda")
    lb = Letterbox((h, w), True)
    for _ in range(3):
        lb(src, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        lb(src, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nw, nh, _, _ = letterbox_geometry(iw, ih, w, h, True)
    alg = a.n * (ih * iw * 3 + h * w * 3)                    # source read once + canvas written once
    print(f"device letterbox {a.src}->{a.dst} x{a.n}: {ms * 1e3:.1f} us per batch, {a.n / ms * 1e3:.0f} images/s, "
          f"{alg / ms / 1e6:.1f} GB/s algorithmic (source + canvas; intermediate {a.n * ih * nw * 3 / 1e6:.1f} MB extra)")
    try:
        from PIL import Image
        pil = [Image.fromarray(im) for im in imgs[:8]]
        t0 = time.perf_counter()
        for im in pil:
            canvas = Image.new("RGB", (w, h), (128, 128, 128))
            canvas.paste(im.resize((nw, nh), Image.BICUBIC), ((w - nw) // 2, (h - nh) // 2))
        dt = (time.perf_counter() - t0) / len(pil)
        print(f"Pillow on one host core: {dt * 1e3:.2f} ms per image = {1 / dt:.0f} images/s")
        got = out[:8].cpu().numpy()
        print("bit-exact vs Pillow (last image):", bool(np.array_equal(got[7], np.asarray(canvas))))
    except ImportError:
        print("Pillow not installed")


if __name__ == "__main__":
    main()
