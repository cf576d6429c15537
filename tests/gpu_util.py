"""Helpers shared by the `-m gpu` tests and tools/gpu_diag.py: drive single C-ABI calls from torch tensors."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from transparent_object_detection_b200 import _lib
from transparent_object_detection_b200._lib import ConvDesc, check
from transparent_object_detection_b200.engine import pack_conv_weight


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def run_conv(x_buf: torch.Tensor, x_off: int, cin: int, w: torch.Tensor, bias: Optional[torch.Tensor],
             out_buf: torch.Tensor, out_off: int, stride: int = 1, act: int = 1,
             residual: Optional[torch.Tensor] = None, res_off: int = 0, upadd: Optional[torch.Tensor] = None,
             block_k: int = 0, num_stages: int = 0, simt: bool = False, variant: int = 0, m: int = 0,
             no_station: int = 0, flags: int = 0):
    """x_buf bf16 (B, H, W, pitch) cuda; w f32 (cout, cin, k, k); out_buf bf16|f32 (B, Ho, Wo, pitch)."""
    L = _lib.lib()
    cout, _, k, _ = w.shape
    wp = pack_conv_weight(w.cpu().float(), block_k).cuda()
    d = ConvDesc()
    d.d_x = x_buf.data_ptr() + x_off * 2
    d.d_w = wp.data_ptr()
    b = bias.float().cuda().contiguous() if bias is not None else None
    d.d_bias = b.data_ptr() if b is not None else None
    d.d_residual = residual.data_ptr() + res_off * 2 if residual is not None else None
    d.d_upadd = upadd.data_ptr() if upadd is not None else None
    d.d_out = out_buf.data_ptr() + out_off * out_buf.element_size()
    d.batch, d.hin, d.win = x_buf.shape[0], x_buf.shape[1], x_buf.shape[2]
    d.cin, d.cout, d.ksize, d.stride = cin, cout, k, stride
    d.x_pitch, d.out_pitch = x_buf.shape[3], out_buf.shape[3]
    d.res_pitch = residual.shape[3] if residual is not None else 0
    d.act = act
    d.out_dtype = 1 if out_buf.dtype == torch.float32 else 0
    d.block_k, d.num_stages = block_k, num_stages
    d.reserved[0], d.reserved[1], d.reserved[2] = variant, m, no_station   # kernel variant knobs (conv_tcgen05.cu)
    d.flags = flags
    fn = L.tod_conv2d_nhwc_bf16_simt_check if simt else L.tod_conv2d_nhwc_bf16
    check(fn(C.byref(d), stream()), "conv")
    torch.cuda.synchronize()
    return wp  # keep alive until sync


def conv_reference(x_buf, x_off, cin, w, bias, stride, act, residual=None, res_off=0, upadd=None):
    """fp32 torch evaluation of the same op on the bf16-rounded operands -> (B, Ho, Wo, cout) f32."""
    cout, _, k, _ = w.shape
    x = x_buf[..., x_off:x_off + cin].float().permute(0, 3, 1, 2)
    wq = w.cuda().to(torch.bfloat16).float()
    y = F.conv2d(x, wq, None, stride=stride, padding=k // 2)
    if bias is not None:
        y = y + bias.cuda().float().view(1, -1, 1, 1)
    if upadd is not None:
        y = y + F.interpolate(upadd.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    if act == 1:
        y = F.silu(y)
    if residual is not None:
        y = y + residual[..., res_off:res_off + cout].float().permute(0, 3, 1, 2)
    return y.permute(0, 2, 3, 1).contiguous()


def error_report(got: torch.Tensor, want: torch.Tensor, tag: str, tol_abs: float, tol_rel: float) -> dict:
    """got/want (B, H, W, C) f32.  Summarises where (rows = pixels, columns = channels) mismatches sit."""
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bad = err > (tol_abs + tol_rel * want.abs())
    rep = {"tag": tag, "max_abs": float(err.max()), "ref_absmax": float(want.abs().max()),
           "bad_frac": float(bad.float().mean()), "nan": int(torch.isnan(got).sum())}
    if bad.any():
        flat = bad.reshape(-1, bad.shape[-1])
        rows = flat.any(1).nonzero().flatten()
        cols = flat.any(0).nonzero().flatten()
        rep["bad_rows"] = f"{rows.numel()}/{flat.shape[0]} first {rows[:12].tolist()} mod8 hist {torch.bincount(rows % 8, minlength=8).tolist()}"
        rep["bad_cols"] = f"{cols.numel()}/{flat.shape[1]} first {cols[:16].tolist()}"
        i = int(err.reshape(-1).argmax())
        rep["worst"] = f"idx {i} got {float(got.reshape(-1)[i]):.5f} want {float(want.reshape(-1)[i]):.5f}"
    return rep
