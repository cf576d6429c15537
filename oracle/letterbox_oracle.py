"""CPU oracle of the reference's image preprocessing: resize_image (utils/utils.py:16-30) + the uint8 canvas that
callbacks.py:142-144 / get_map.py:57-60 turn into the network input.

TEST INFRASTRUCTURE -- NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import it.

The arithmetic lives in a third-party dependency that is not under /root/reference: Pillow's `Image.resize(size,
Image.BICUBIC)` (unpinned by the reference; 12.2.0 in the authoring container).  This file restates Pillow's published
algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
ImagingResampleVertical_8bpc) in numpy integer arithmetic:
  * per output index: centre = (i + 0.5) * scale, support = 2 * max(scale, 1), window [xmin, xmin + n) with
    xmin = max(0, int(centre - support + 0.5)), xmax = min(in, int(centre + support + 0.5)); bicubic weights (a = -0.5) at
    (x + xmin - centre + 0.5) / max(scale, 1), normalised to sum 1 in double;
  * weights -> 22-bit fixed point, rounded half away from zero;
  * horizontal pass over every input row, then vertical pass, each `clip8((2^21 + sum(pixel * k)) >> 22)` to uint8.
Parity pin: tests/test_letterbox_cpu.py compares this restatement with Pillow itself (bit-exact, random sizes) and with
tests/golden/letterbox.npz written by oracle/make_golden_letterbox.py from Pillow's outputs.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2   # Resample.c


def _bicubic(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (in0 = 0, in1 = in_size).
    -> bounds int32 (out, 2) [xmin, count], coefficients int32 (out, ksize), ksize."""
    scale = float(in_size) / out_size          # (double)(in1 - in0) / outSize with float in0, in1
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass_rows(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One separable pass along axis 1 of a (rows, in, ch) uint8 array -> (rows, out, ch) uint8."""
    rows, _, ch = img.shape
    out = np.empty((rows, bounds.shape[0], ch), np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (src[:, xmin:xmin + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), Image.BICUBIC) on an (h, w, ch) uint8 array (ImagingResample: horizontal
    pass first, then vertical; a pass whose size does not change is skipped)."""
    h, w, _ = img.shape
    cur = img
    if out_w != w:
        b, k, _ = precompute_coeffs(w, out_w)
        cur = _pass_rows(cur, b, k)
    if out_h != h:
        b, k, _ = precompute_coeffs(h, out_h)
        cur = _pass_rows(cur.transpose(1, 0, 2), b, k).transpose(1, 0, 2)
    return np.ascontiguousarray(cur)


def letterbox_geometry(iw: int, ih: int, w: int, h: int, letterbox_image: bool) -> Tuple[int, int, int, int]:
    """resize_image utils/utils.py:18-27: -> (nw, nh, x0, y0) of the resized image inside the (w, h) canvas."""
    if not letterbox_image:
        return w, h, 0, 0
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


def resize_image_u8(img: np.ndarray, size: Tuple[int, int], letterbox_image: bool) -> np.ndarray:
    """reference resize_image (utils/utils.py:16-30) on an (h, w, 3) uint8 RGB array; size = (w, h) -> (h, w, 3) uint8
    (grey 128 canvas + pasted bicubic resize when letterbox_image)."""
    ih, iw, _ = img.shape
    w, h = size
    nw, nh, x0, y0 = letterbox_geometry(iw, ih, w, h, letterbox_image)
    small = resize_bicubic_u8(img, nw, nh)
    if not letterbox_image:
        return small
    canvas = np.full((h, w, 3), 128, np.uint8)
    canvas[y0:y0 + nh, x0:x0 + nw] = small
    return canvas
