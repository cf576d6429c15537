"""Attention blocks of the reference's current-source backbone / head: the same constructors,
parameter names and call signature as model/blocks.py:190-254, evaluated by libtod.so on NHWC bf16.  These classes are the
block-level drop-ins; inside the captured network plan (`BaseModel(..., attention=True)`) the engine issues the same C-ABI
calls on its own pre-allocated buffers (engine.py:_cbam, _self_attention) and shares `_gemm` / `unfused_attention_image`
with this module.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ._lib import CbamDesc, check, lib

LOG2E = 1.4426950408889634


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


# ----------------------------------------------------------------------------------------------- SelfAttention
def _gemm(L, st, x_ptr, h, w, cin, x_pitch, w_ptr, cout, out_ptr, out_pitch, out_f32=False, bias_ptr=None,
          res_ptr=None, res_pitch=0, what="gemm", dynamic_w=False):
    """D[h*w, cout] = A[h*w, cin] . W[cout, cin]^T (+ bias) (+ residual) through tod_conv2d_nhwc_bf16 as a 1x1 conv.
    dynamic_w: the "weight" operand is an activation written by an earlier kernel of the stream (k in q . k^T, x in
    (gamma Wv) . x^T, v^T in P . v^T): the conv kernel must not prefetch it ahead of its programmatic-launch wait."""
    from ._lib import ConvDesc, TOD_ACT_NONE, TOD_CONV_DYNAMIC_W, TOD_OUT_BF16, TOD_OUT_F32
    d = ConvDesc()
    d.flags = TOD_CONV_DYNAMIC_W if dynamic_w else 0
    d.d_x, d.d_w, d.d_out = x_ptr, w_ptr, out_ptr
    d.d_bias, d.d_residual = bias_ptr, res_ptr
    d.batch, d.hin, d.win, d.cin, d.cout, d.ksize, d.stride = 1, h, w, cin, cout, 1, 1
    d.x_pitch, d.out_pitch, d.res_pitch = x_pitch, out_pitch, res_pitch
    d.act, d.out_dtype = TOD_ACT_NONE, (TOD_OUT_F32 if out_f32 else TOD_OUT_BF16)
    check(L.tod_conv2d_nhwc_bf16(C.byref(d), st), what)


def unfused_attention_image(L, st, xi, h, w, Cc, d16, wq, bq, wk, bk, wv, bv, qi, ki, S, P, vT, out_i):
    """One image of the unfused SelfAttention chain (model/blocks.py:239-253) on raw device pointers: q / k projections,
    S = q k^T (f32, N x N), row softmax -> P (bf16), vT = (gamma Wv) x^T, out = P vT^T + gamma bv + x."""
    N = h * w
    _gemm(L, st, xi, h, w, Cc, Cc, wq, d16, qi, d16, bias_ptr=bq, what="query")
    _gemm(L, st, xi, h, w, Cc, Cc, wk, d16, ki, d16, bias_ptr=bk, what="key")
    _gemm(L, st, qi, h, w, d16, d16, ki, N, S, N, out_f32=True, what="scores", dynamic_w=True)
    check(L.tod_softmax_rows_f32_bf16(S, P, N, N, N, N, st), "softmax")
    _gemm(L, st, wv, 1, Cc, Cc, Cc, xi, N, vT, N, what="value^T", dynamic_w=True)
    _gemm(L, st, P, h, w, N, N, vT, Cc, out_i, Cc, bias_ptr=bv, res_ptr=xi, res_pitch=Cc, what="attention output",
          dynamic_w=True)


def self_attention_fused_nhwc(x: torch.Tensor, wq, bq, wk, bk, wv, bv, gamma: float) -> torch.Tensor:
    """The same block with tod_attention_fused (csrc/attention_tcgen05.cu): q / k projections over the whole batch, V^T per
    image, then ONE kernel that keeps the scores and the attention weights on chip."""
    from ._lib import AttentionDesc
    from .engine import pack_conv_weight
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()):
        raise ValueError("x must be a contiguous bf16 CUDA tensor (B, H, W, C)")
    B, H, W, Cc = x.shape
    N = H * W
    if N % 16 or Cc % 32 or Cc > 256:
        raise ValueError(f"fused SelfAttention needs H * W % 16 == 0 and C % 32 == 0, C <= 256 (got {H}x{W}, C = {Cc})")
    dev, L = x.device, lib()
    d = wq.shape[0]
    d16 = 16 if d <= 16 else (32 if d <= 32 else 64)
    if d > 64:
        raise ValueError("fused SelfAttention supports q/k widths up to 64")

    def padded(wt, bs, scale=1.0):
        wp, bp = torch.zeros((d16, Cc, 1, 1)), torch.zeros((d16,))
        wp[:d], bp[:d] = scale * wt.detach().float().cpu().reshape(d, Cc, 1, 1), scale * bs.detach().float().cpu()
        return pack_conv_weight(wp).to(dev), bp.to(dev)

    wq_p, bq_p = padded(wq, bq, LOG2E)          # base-2 logits: the kernel's softmax is one MUFU.EX2 per score
    wk_p, bk_p = padded(wk, bk)
    bv_g = (float(gamma) * bv.detach().float().cpu()).to(dev).contiguous()
    q = torch.empty((B, N, d16), dtype=torch.bfloat16, device=dev)
    k = torch.empty_like(q)
    vT = torch.empty((B, Cc, N), dtype=torch.bfloat16, device=dev)
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _gemm(L, st, x.data_ptr(), B * H, W, Cc, Cc, wq_p.data_ptr(), d16, q.data_ptr(), d16, bias_ptr=bq_p.data_ptr(), what="query")
        _gemm(L, st, x.data_ptr(), B * H, W, Cc, Cc, wk_p.data_ptr(), d16, k.data_ptr(), d16, bias_ptr=bk_p.data_ptr(), what="key")
        wv_p = pack_conv_weight((float(gamma) * wv.detach().float().cpu()).reshape(Cc, Cc, 1, 1)).to(dev)
        vn = torch.empty((B, N, Cc), dtype=torch.bfloat16, device=dev)
        _gemm(L, st, x.data_ptr(), B * H, W, Cc, Cc, wv_p.data_ptr(), Cc, vn.data_ptr(), Cc, what="value")
        check(L.tod_transpose_bf16(vn.data_ptr(), vT.data_ptr(), B, N, Cc, Cc, N, st), "tod_transpose_bf16")
        a = AttentionDesc()
        a.d_q, a.d_k, a.d_vt, a.d_bias = q.data_ptr(), k.data_ptr(), vT.data_ptr(), bv_g.data_ptr()
        a.d_x, a.d_out = x.data_ptr(), out.data_ptr()
        a.batch, a.n, a.c, a.d16, a.x_pitch, a.out_pitch = B, N, Cc, d16, Cc, Cc
        check(L.tod_attention_fused(C.byref(a), st), "tod_attention_fused")
        torch.cuda.current_stream(dev).synchronize()
    return out


def self_attention_nhwc(x: torch.Tensor, wq, bq, wk, bk, wv, bv, gamma: float) -> torch.Tensor:
    """reference SelfAttention.forward (model/blocks.py:236-254) on a dense NHWC bf16 CUDA tensor (B, H, W, C), H * W % 16
    == 0.  UNFUSED baseline: per image the N x N scores are materialised (f32), soft-maxed into bf16 weights and applied
    with a second GEMM; every GEMM is the tcgen05 conv kernel used as a 1x1 conv:
        q, k = x Wq^T + bq, x Wk^T + bk                (output channels zero-padded to a multiple of 16)
        S    = q k^T                                   (A = q_i, "weights" = k_i: its NHWC rows are already [N, d] K-major)
        P    = softmax_rows(S)                         (tod_softmax_rows_f32_bf16)
        vT   = (gamma Wv) x_i^T                        (A = gamma Wv, "weights" = x_i) -> [C, N], K-major for the last GEMM
        out  = P vT^T + gamma bv + x_i                 (softmax rows sum to one, so the value bias moves out of the sum)"""
    from .engine import pack_conv_weight
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()):
        raise ValueError("x must be a contiguous bf16 CUDA tensor (B, H, W, C)")
    B, H, W, Cc = x.shape
    N = H * W
    if N % 16 or Cc % 16:
        raise ValueError(f"SelfAttention needs H * W and C to be multiples of 16 (got {H}x{W}, C = {Cc})")
    dev = x.device
    L = lib()
    d = wq.shape[0]
    d16 = (d + 15) // 16 * 16

    def padded(wt, bs):
        wp = torch.zeros((d16, Cc, 1, 1), dtype=torch.float32)
        bp = torch.zeros((d16,), dtype=torch.float32)
        wp[:d] = wt.detach().float().cpu().reshape(d, Cc, 1, 1)
        bp[:d] = bs.detach().float().cpu()
        return pack_conv_weight(wp).to(dev), bp.to(dev)

    wq_p, bq_p = padded(wq, bq)
    wk_p, bk_p = padded(wk, bk)
    wv_g = (float(gamma) * wv.detach().float().cpu().reshape(Cc, Cc)).to(torch.bfloat16).to(dev).contiguous()   # A operand [C, C_in]
    bv_g = (float(gamma) * bv.detach().float().cpu()).to(dev).contiguous()
    q = torch.empty((B, H, W, d16), dtype=torch.bfloat16, device=dev)
    k = torch.empty_like(q)
    out = torch.empty_like(x)
    S = torch.empty((N, N), dtype=torch.float32, device=dev)
    P = torch.empty((N, N), dtype=torch.bfloat16, device=dev)
    vT = torch.empty((Cc, N), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for i in range(B):
            xi = x.data_ptr() + i * N * Cc * 2
            unfused_attention_image(L, st, xi, H, W, Cc, d16, wq_p.data_ptr(), bq_p.data_ptr(), wk_p.data_ptr(), bk_p.data_ptr(),
                                    wv_g.data_ptr(), bv_g.data_ptr(), q.data_ptr() + i * N * d16 * 2, k.data_ptr() + i * N * d16 * 2,
                                    S.data_ptr(), P.data_ptr(), vT.data_ptr(), out.data_ptr() + i * N * Cc * 2)
        torch.cuda.current_stream(dev).synchronize()      # the temporaries above are freed on return
    return out


class SelfAttention(nn.Module):
    """reference SelfAttention (model/blocks.py:226-254): same constructor and parameter names (query, key, value, gamma),
    forward on the reference's NCHW float tensor."""

    def __init__(self, channels: int):
        super().__init__()
        self.query = nn.Conv2d(channels, channels // 8, kernel_size=1)
        self.key = nn.Conv2d(channels, channels // 8, kernel_size=1)
        self.value = nn.Conv2d(channels, channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))

    fused = True      # tod_attention_fused when the shape allows it (C % 32 == 0, C <= 256); else the unfused GEMM chain

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        xn = x.to(dev).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        fn = self_attention_fused_nhwc if (self.fused and xn.shape[3] % 32 == 0 and xn.shape[3] <= 256) else self_attention_nhwc
        y = fn(xn, self.query.weight, self.query.bias, self.key.weight, self.key.bias, self.value.weight,
                                self.value.bias, float(self.gamma.detach()))
        return y.permute(0, 3, 1, 2).to(x.dtype).to(x.device)
