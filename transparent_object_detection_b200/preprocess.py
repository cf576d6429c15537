"""Device-side letterbox: the reference's resize_image (utils/utils.py:16-30) as called by the detect pipelines
(utils/callbacks.py:142-143, dataset/coco/get_map.py:57-59), bit-exact with Pillow's BICUBIC resize, writing straight
into the uint8 NHWC input batch of the network.  The kernels are in csrc/letterbox.cu
behind tod_letterbox_bicubic_u8; the per-axis windows and fixed-point weights come from the library's host function
tod_resample_coeffs_bicubic and are cached per (source size, resized size, device).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from ._lib import LetterboxDesc, check, lib


def letterbox_geometry(iw: int, ih: int, w: int, h: int, letterbox_image: bool) -> Tuple[int, int, int, int]:
    """utils/utils.py:18-27 -> (nw, nh, x0, y0): size and position of the resized image on the (w, h) canvas."""
    if not letterbox_image:
        return w, h, 0, 0
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's bicubic windows / 22-bit weights of one axis (host): bounds i32 (out, 2), coefficients i32 (out, ksize)."""
    L = lib()
    ks = C.c_int32()
    check(L.tod_resample_coeffs_bicubic(in_size, out_size, None, None, C.byref(ks)), "tod_resample_coeffs_bicubic")
    bounds = np.zeros((out_size, 2), np.int32)
    coef = np.zeros((out_size, ks.value), np.int32)
    check(L.tod_resample_coeffs_bicubic(in_size, out_size, bounds.ctypes.data, coef.ctypes.data, C.byref(ks)),
          "tod_resample_coeffs_bicubic")
    return bounds, coef, ks.value


class Letterbox:
    """letterbox = Letterbox((H, W)); letterbox(images, out) with images uint8 (n, ih, iw, 3) (device, or host -> copied)
    and out a uint8 (n, H, W, 3) device tensor (e.g. a slice of the engine's static input batch)."""

    def __init__(self, input_shape: Sequence[int], letterbox_image: bool = True, pad_value: int = 128):
        self.input_shape = (int(input_shape[0]), int(input_shape[1]))
        self.letterbox_image, self.pad_value = bool(letterbox_image), int(pad_value)
        self._plans: Dict[Tuple, dict] = {}

    def _plan(self, ih: int, iw: int, device: torch.device) -> dict:
        key = (ih, iw, str(device))
        pl = self._plans.get(key)
        if pl is None:
            H, W = self.input_shape
            nw, nh, x0, y0 = letterbox_geometry(iw, ih, W, H, self.letterbox_image)
            if nw <= 0 or nh <= 0:
                raise ValueError(f"image {iw}x{ih} collapses to {nw}x{nh} on a {W}x{H} canvas")
            pl = dict(nw=nw, nh=nh, x0=x0, y0=y0, xb=None, xk=None, xks=0, yb=None, yk=None, yks=0)
            if nw != iw:
                b, k, ks = resample_coeffs(iw, nw)
                pl.update(xb=torch.from_numpy(b).to(device), xk=torch.from_numpy(k).to(device), xks=ks)
            if nh != ih:
                b, k, ks = resample_coeffs(ih, nh)
                pl.update(yb=torch.from_numpy(b).to(device), yk=torch.from_numpy(k).to(device), yks=ks)
            self._plans[key] = pl
        return pl

    def __call__(self, images: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if images.dim() == 3:
            images = images.unsqueeze(0)
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
            raise ValueError(f"images must be uint8 (n, h, w, 3), got {images.dtype} {tuple(images.shape)}")
        H, W = self.input_shape
        if not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (images.shape[0], H, W, 3)):
            raise ValueError(f"out must be a contiguous uint8 CUDA tensor of shape {(images.shape[0], H, W, 3)}")
        dev = out.device
        src = images.to(dev, non_blocking=True).contiguous()
        n, ih, iw, _ = src.shape
        pl = self._plan(ih, iw, dev)
        need_h = pl["nw"] != iw
        tmp = torch.empty((n, ih, W, 3), dtype=torch.uint8, device=dev) if need_h else None   # canvas-width rows (tod.h)
        d = LetterboxDesc()
        d.d_src, d.d_dst = src.data_ptr(), out.data_ptr()
        d.d_tmp = tmp.data_ptr() if tmp is not None else None
        d.src_image_stride, d.dst_image_stride = ih * iw * 3, H * W * 3
        d.n, d.src_h, d.src_w, d.dst_h, d.dst_w = n, ih, iw, H, W
        d.new_h, d.new_w, d.off_y, d.off_x, d.pad_value = pl["nh"], pl["nw"], pl["y0"], pl["x0"], self.pad_value
        d.d_xbounds = pl["xb"].data_ptr() if pl["xb"] is not None else None
        d.d_xcoef = pl["xk"].data_ptr() if pl["xk"] is not None else None
        d.d_ybounds = pl["yb"].data_ptr() if pl["yb"] is not None else None
        d.d_ycoef = pl["yk"].data_ptr() if pl["yk"] is not None else None
        d.xksize, d.yksize = pl["xks"], pl["yks"]
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev)
            check(lib().tod_letterbox_bicubic_u8(C.byref(d), st.cuda_stream), "tod_letterbox_bicubic_u8")
            # src / tmp are consumed by the kernels just enqueued on this stream; the caching allocator keeps
            # stream-ordered reuse safe as long as they are recorded on it
            src.record_stream(st)
            if tmp is not None:
                tmp.record_stream(st)
        return out
