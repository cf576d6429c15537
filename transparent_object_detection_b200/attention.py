"""Attention blocks of the reference's current-source backbone / head, block level: the
same constructors, parameter names and call signature as model/blocks.py, evaluated by libtod.so on NHWC bf16.  They are
not yet part of the captured network plan (DetectorEngine builds the plain topology).  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ._lib import CbamDesc, check, lib


def cbam_nhwc(x: torch.Tensor, fc1: torch.Tensor, fc2: torch.Tensor, conv: torch.Tensor, out: torch.Tensor = None,
              channels: int = None) -> torch.Tensor:
    """x: bf16 CUDA (B, H, W, pitch) NHWC (the first `channels` of every pixel are used); fc1 (hidden, C), fc2 (C, hidden),
    conv (2, k, k) float32 CUDA.  Returns `out` (default: a new (B, H, W, C) bf16 tensor; may be x itself)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()):
        raise ValueError("x must be a contiguous bf16 CUDA tensor (B, H, W, pitch)")
    B, H, W, pitch = x.shape
    c = pitch if channels is None else channels
    if out is None:
        out = torch.empty((B, H, W, c), dtype=torch.bfloat16, device=x.device)
    L = lib()
    work = torch.empty(int(L.tod_cbam_workspace_floats(B, H, W, c)), dtype=torch.float32, device=x.device)
    d = CbamDesc()
    d.d_x, d.d_out, d.d_work = x.data_ptr(), out.data_ptr(), work.data_ptr()
    d.d_fc1, d.d_fc2, d.d_conv = fc1.data_ptr(), fc2.data_ptr(), conv.data_ptr()
    d.batch, d.h, d.w, d.c, d.hidden, d.ksize = B, H, W, c, fc1.shape[0], conv.shape[-1]
    d.x_pitch, d.out_pitch = pitch, out.shape[3]
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device)
        check(L.tod_cbam_nhwc_bf16(C.byref(d), st.cuda_stream), "tod_cbam_nhwc_bf16")
        work.record_stream(st)
    return out


class CBAM(nn.Module):
    """reference CBAM (model/blocks.py:190-223): same constructor, same parameter names (fc1, fc2, conv), forward on the
    reference's NCHW float tensor; the arithmetic runs in csrc/cbam.cu on the NHWC bf16 copy."""

    def __init__(self, channels: int, reduction: int = 16, kernel_size: int = 7):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, channels // reduction, 1, bias=False)
        self.fc2 = nn.Conv2d(channels // reduction, channels, 1, bias=False)
        self.conv = nn.Conv2d(2, 1, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        xn = x.to(dev).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        y = cbam_nhwc(xn, f32(self.fc1.weight.flatten(1)), f32(self.fc2.weight.flatten(1)), f32(self.conv.weight[0]))
        return y.permute(0, 3, 1, 2).to(x.dtype).to(x.device)
