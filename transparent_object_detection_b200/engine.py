"""Execution plan for the detector hot path on one B200.

Host-side only: folds BatchNorm into the conv weights (reference fuse_conv, model/blocks.py:160-187),
packs them to the K-major bf16 layout the tcgen05 kernel consumes, lays the activations out in HBM as
NHWC bf16 buffers whose concat inputs are pre-allocated (producers write at channel offsets -- no
torch.cat, no chunk copies), and records the kernel sequence of

    Backbone.forward (model/backbone.py:50-59)  ->  Neck.forward (model/neck.py:55-61, C2f stages per
    SURVEY F4)  ->  Head.forward (model/head.py:46-61)  ->  decode_box + non_max_suppression
    (utils/bbox_utils.py:66-82, 119-175)

as a list of C-ABI calls (include/tod.h).  Every arithmetic step runs in libtod.so; PyTorch supplies
device memory, streams and CUDA-graph capture only.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (CbamDesc, ConvDesc, ConvTailDesc, DecodeDesc, HeadFuseDesc, TOD_ACT_NONE, TOD_ACT_SILU, TOD_CONV_REVERSE,
                   TOD_FUSE_BOX, TOD_FUSE_CLS, TOD_OUT_BF16, TOD_OUT_F32, check)

BN_EPS = 1e-5


@dataclass
class View:
    """A channel window [c_off, c_off + c) of an NHWC buffer (B, h, w, pitch)."""
    buf: torch.Tensor
    c_off: int
    c: int

    @property
    def h(self) -> int:
        return self.buf.shape[1]

    @property
    def w(self) -> int:
        return self.buf.shape[2]

    @property
    def pitch(self) -> int:
        return self.buf.shape[3]

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.c_off * self.buf.element_size()

    def sub(self, off: int, c: int) -> "View":
        assert 0 <= off and off + c <= self.c
        return View(self.buf, self.c_off + off, c)

    def tensor(self) -> torch.Tensor:
        return self.buf[..., self.c_off:self.c_off + self.c]


def _t(sd, key) -> torch.Tensor:
    v = sd[key]
    if not isinstance(v, torch.Tensor):
        v = torch.from_numpy(np.asarray(v))
    return v.detach().to("cpu", torch.float32)


def fold_conv_bn(sd, prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """W' = W * g/sqrt(var+eps) per output channel, b' = beta - g*mean/sqrt(var+eps)  (model/blocks.py:179-185)."""
    w = _t(sd, prefix + ".conv.weight")
    g, beta = _t(sd, prefix + ".norm.weight"), _t(sd, prefix + ".norm.bias")
    mean, var = _t(sd, prefix + ".norm.running_mean"), _t(sd, prefix + ".norm.running_var")
    scale = g / torch.sqrt(var + BN_EPS)
    return w * scale.view(-1, 1, 1, 1), beta - mean * scale


def pack_conv_weight(w: torch.Tensor, block_k_hint: int = 0) -> torch.Tensor:
    """[cout, cin, k, k] f32 -> [cout, k*k*cin_pad] bf16 with K index = tap*cin_pad + c (tod_conv_weight_layout)."""
    cout, cin, k, _ = w.shape
    _, cin_pad, k_total = _lib.weight_layout(cin, k, block_k_hint)
    p = torch.zeros(cout, k * k, cin_pad, dtype=torch.float32)
    p[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    return p.reshape(cout, k_total).to(torch.bfloat16).contiguous()


class DetectorEngine:
    """Fixed-shape plan: (batch, 3, in_h, in_w) float32 NCHW in -> raw maps / head tensor / detections."""

    def __init__(self, state_dict, num_classes: int, base_channels: int, base_depth: int, deep_mul: float,
                 batch: int, in_h: int, in_w: int, device: Optional[torch.device] = None, attention: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("transparent_object_detection_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.L = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda")
        with torch.cuda.device(self.device):
            ok = self.L.tod_device_ok()
        if ok != 1:
            raise RuntimeError("libtod.so targets sm_100a (B200); current device is not compute capability 10.x")
        if in_h % 32 or in_w % 32:
            raise ValueError("input height/width must be multiples of 32")
        self.nc, self.C, self.d, self.deep_mul = num_classes, base_channels, base_depth, deep_mul
        self.C5 = int(base_channels * 16 * deep_mul)
        self.batch, self.in_h, self.in_w = batch, in_h, in_w
        self.attention = bool(attention)   # current-source backbone / head: CBAM and SelfAttention ops in the plan
        self.ops: List[Tuple[str, str, object]] = []   # (kind, name, payload)
        self._keep: List[torch.Tensor] = []            # weights / biases / activation arena kept alive
        self.conv_flops = 0
        self.conv_meta: Dict[str, dict] = {}          # fp32 weights / views per conv (tools/gpu_netcheck.py)
        self.launches_forward = 0
        self.fork_head = True          # head towers as parallel graph branches (graph_for)
        self.prioritise_critical_path = os.environ.get("TOD_GRAPH_PRIO", "1") != "0"
        self._box_params = None
        self._graphs: Dict[Tuple, "torch.cuda.CUDAGraph"] = {}   # (input kind, slot, conf, iou, head_out, decoded)
        self._inputs: Dict[Tuple[str, int], torch.Tensor] = {}   # static input buffers per (kind, slot)
        self._packed: Optional[dict] = None
        self._side_streams: Dict[Tuple[str, int], "torch.cuda.Stream"] = {}
        self._build(state_dict)

    # ------------------------------------------------------------------ memory
    def _buf(self, h: int, w: int, c: int, dtype=torch.bfloat16) -> View:
        t = torch.zeros((self.batch, h, w, c), dtype=dtype, device=self.device)
        self._keep.append(t)   # descriptors hold raw pointers: the arena must outlive them
        return View(t, 0, c)

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous().to(self.device)
        self._keep.append(t)
        return t

    # ------------------------------------------------------------------ op builders
    def _conv(self, name: str, w: torch.Tensor, b: Optional[torch.Tensor], src: View, dst: View, stride: int = 1,
              act: int = TOD_ACT_SILU, residual: Optional[View] = None, upadd: Optional[torch.Tensor] = None,
              out_f32: bool = False) -> None:
        cout, cin, k, _ = w.shape
        assert cin == src.c and cout == dst.c, (name, cin, src.c, cout, dst.c)
        wp = self._dev(pack_conv_weight(w))
        d = ConvDesc()
        d.d_x, d.d_w, d.d_out = src.ptr, wp.data_ptr(), dst.ptr
        d.d_bias = self._dev(b.to(torch.float32)).data_ptr() if b is not None else None
        d.d_residual = residual.ptr if residual is not None else None
        d.d_upadd = upadd.data_ptr() if upadd is not None else None
        d.batch, d.hin, d.win, d.cin, d.cout = self.batch, src.h, src.w, cin, cout
        d.ksize, d.stride = k, stride
        d.x_pitch, d.out_pitch = src.pitch, dst.pitch
        d.res_pitch = residual.pitch if residual is not None else 0
        d.act, d.out_dtype = act, (TOD_OUT_F32 if out_f32 else TOD_OUT_BF16)
        assert dst.h == src.h // stride and dst.w == src.w // stride, name
        self.ops.append(("conv", name, d))
        self.conv_meta[name] = dict(w=w, b=b, src=src, dst=dst, stride=stride, act=act, residual=residual, upadd=upadd,
                                    out_f32=out_f32)
        self.conv_flops += 2 * self.batch * dst.h * dst.w * cout * cin * k * k

    def _cbam(self, sd, prefix: str, v: View) -> None:
        """reference CBAM (model/blocks.py:206-223) in place on a channel view (tod_cbam_nhwc_bf16)."""
        fc1 = self._dev(_t(sd, prefix + ".fc1.weight").float().flatten(1))
        fc2 = self._dev(_t(sd, prefix + ".fc2.weight").float().flatten(1))
        cw = self._dev(_t(sd, prefix + ".conv.weight").float()[0])
        work = torch.zeros(int(self.L.tod_cbam_workspace_floats(self.batch, v.h, v.w, v.c)), dtype=torch.float32, device=self.device)
        self._keep.append(work)
        d = CbamDesc()
        d.d_x, d.d_out, d.d_work = v.ptr, v.ptr, work.data_ptr()
        d.d_fc1, d.d_fc2, d.d_conv = fc1.data_ptr(), fc2.data_ptr(), cw.data_ptr()
        d.batch, d.h, d.w, d.c, d.hidden, d.ksize = self.batch, v.h, v.w, v.c, fc1.shape[0], cw.shape[-1]
        d.x_pitch, d.out_pitch = v.pitch, v.pitch
        self.ops.append(("cbam", prefix, d))

    def _self_attention(self, sd, prefix: str, v: View) -> None:
        """reference SelfAttention (model/blocks.py:236-254) in place on a dense map: the unfused GEMM chain of
        attention.self_attention_nhwc with every operand and temporary allocated once (graph-capturable)."""
        assert v.pitch == v.c and v.c_off == 0, "SelfAttention needs a dense buffer"
        Cc, N = v.c, v.h * v.w
        assert N % 16 == 0 and Cc % 16 == 0, (N, Cc)
        wq, wk = _t(sd, prefix + ".query.weight").float(), _t(sd, prefix + ".key.weight").float()
        dq = wq.shape[0]
        d16 = (dq + 15) // 16 * 16
        fused = Cc % 32 == 0 and Cc <= 256 and dq <= 64 and os.environ.get("TOD_ATTN_FUSED", "1") != "0"
        if fused:
            d16 = 16 if dq <= 16 else (32 if dq <= 32 else 64)
        gamma = float(_t(sd, prefix + ".gamma").reshape(-1)[0])

        def padded(wt, bs, scale=1.0):
            wp, bp = torch.zeros((d16, Cc, 1, 1)), torch.zeros((d16,))
            wp[:dq], bp[:dq] = scale * wt.reshape(dq, Cc, 1, 1), scale * bs.float()
            return self._dev(pack_conv_weight(wp)), self._dev(bp)

        plan = dict(v=v, N=N, C=Cc, d16=d16, fused=fused)
        plan["wq"], plan["bq"] = padded(wq, _t(sd, prefix + ".query.bias"), 1.4426950408889634 if fused else 1.0)   # base-2 logits
        plan["wk"], plan["bk"] = padded(wk, _t(sd, prefix + ".key.bias"))
        plan["wv"] = self._dev((gamma * _t(sd, prefix + ".value.weight").float().reshape(Cc, Cc)).to(torch.bfloat16))
        plan["bv"] = self._dev(gamma * _t(sd, prefix + ".value.bias").float())
        temps = [("q", (self.batch, N, d16), torch.bfloat16), ("k", (self.batch, N, d16), torch.bfloat16)]
        if fused:      # tod_attention_fused: the scores never leave the SM; v = x (gamma Wv)^T batched, then transposed
            temps += [("vT", (self.batch, Cc, N), torch.bfloat16), ("vn", (self.batch, N, Cc), torch.bfloat16)]
            plan["wvp"] = self._dev(pack_conv_weight((gamma * _t(sd, prefix + ".value.weight").float()).reshape(Cc, Cc, 1, 1)))
        else:
            temps += [("S", (N, N), torch.float32), ("P", (N, N), torch.bfloat16), ("vT", (Cc, N), torch.bfloat16)]
        for name, shape, dt in temps:
            plan[name] = torch.zeros(shape, dtype=dt, device=self.device)
            self._keep.append(plan[name])
        self.ops.append(("attn", prefix, plan))
        self.conv_flops += 2 * self.batch * (2 * N * d16 * Cc + N * N * d16 + N * Cc * Cc + N * N * Cc)

    def _run_attention(self, plan: dict, st: int) -> None:
        from .attention import _gemm, unfused_attention_image
        L, v, N, Cc, d16 = self.L, plan["v"], plan["N"], plan["C"], plan["d16"]
        if plan["fused"]:
            from ._lib import AttentionDesc
            B = self.batch
            _gemm(L, st, v.ptr, B * v.h, v.w, Cc, Cc, plan["wq"].data_ptr(), d16, plan["q"].data_ptr(), d16,
                  bias_ptr=plan["bq"].data_ptr(), what="query")
            _gemm(L, st, v.ptr, B * v.h, v.w, Cc, Cc, plan["wk"].data_ptr(), d16, plan["k"].data_ptr(), d16,
                  bias_ptr=plan["bk"].data_ptr(), what="key")
            _gemm(L, st, v.ptr, B * v.h, v.w, Cc, Cc, plan["wvp"].data_ptr(), Cc, plan["vn"].data_ptr(), Cc, what="value")
            check(L.tod_transpose_bf16(plan["vn"].data_ptr(), plan["vT"].data_ptr(), B, N, Cc, Cc, N, st), "tod_transpose_bf16")
            a = AttentionDesc()
            a.d_q, a.d_k, a.d_vt, a.d_bias = plan["q"].data_ptr(), plan["k"].data_ptr(), plan["vT"].data_ptr(), plan["bv"].data_ptr()
            a.d_x, a.d_out = v.ptr, v.ptr
            a.batch, a.n, a.c, a.d16, a.x_pitch, a.out_pitch = B, N, Cc, d16, Cc, Cc
            check(L.tod_attention_fused(C.byref(a), st), "tod_attention_fused")
            return
        for i in range(self.batch):
            xi = v.ptr + i * N * Cc * 2
            unfused_attention_image(L, st, xi, v.h, v.w, Cc, d16, plan["wq"].data_ptr(), plan["bq"].data_ptr(),
                                    plan["wk"].data_ptr(), plan["bk"].data_ptr(), plan["wv"].data_ptr(), plan["bv"].data_ptr(),
                                    plan["q"].data_ptr() + i * N * d16 * 2, plan["k"].data_ptr() + i * N * d16 * 2,
                                    plan["S"].data_ptr(), plan["P"].data_ptr(), plan["vT"].data_ptr(), xi)

    def _conv_bn(self, sd, prefix: str, src: View, dst: View, stride: int = 1, residual: Optional[View] = None) -> None:
        w, b = fold_conv_bn(sd, prefix)
        self._conv(prefix, w, b, src, dst, stride, TOD_ACT_SILU, residual)

    def _c2f(self, sd, prefix: str, src: View, dst: View, n: int, shortcut: bool,
             up_src: Optional[View] = None) -> None:
        """reference C2f.forward (model/blocks.py:104-108).  `up_src`: the stage input is
        cat([upsample2x(up_src), src]) (model/neck.py:57-58): cv1's weight is split column-wise; the up_src part
        is evaluated at LOW resolution (f32, no bias/act) and added pre-activation at [h>>1][w>>1]."""
        c = dst.c // 2
        cat = self._buf(src.h, src.w, (2 + n) * c)
        w, b = fold_conv_bn(sd, prefix + ".cv1")
        if up_src is None:
            self._conv(prefix + ".cv1", w, b, src, cat.sub(0, 2 * c))
        else:
            z = torch.zeros((self.batch, up_src.h, up_src.w, 2 * c), dtype=torch.float32, device=self.device)
            self._keep.append(z)
            zv = View(z, 0, 2 * c)
            self._conv(prefix + ".cv1[up]", w[:, :up_src.c].contiguous(), None, up_src, zv, act=TOD_ACT_NONE, out_f32=True)
            self._conv(prefix + ".cv1", w[:, up_src.c:].contiguous(), b, src, cat.sub(0, 2 * c), upadd=z)
        tmp = self._buf(src.h, src.w, c)
        for j in range(n):
            inp = cat.sub((1 + j) * c, c)
            self._conv_bn(sd, f"{prefix}.m.{j}.cv1", inp, tmp)
            self._conv_bn(sd, f"{prefix}.m.{j}.cv2", tmp, cat.sub((2 + j) * c, c), residual=inp if shortcut else None)
        self._conv_bn(sd, prefix + ".cv2", cat, dst)

    # ------------------------------------------------------------------ the network
    def _build(self, sd) -> None:
        C, d, C5, nc, B = self.C, self.d, self.C5, self.nc, self.batch
        H, W = self.in_h, self.in_w
        dev = self.device
        self.x_static = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        # ---- backbone (model/backbone.py:20-48)
        w, b = fold_conv_bn(sd, "backbone.stem")
        stem = self._buf(H // 2, W // 2, C)
        wh, bh = w.reshape(C, 27).contiguous().float(), b.contiguous().float()      # HOST tensors (kernel parameter)
        self._keep += [wh, bh]
        self.ops.append(("stem", "backbone.stem", (wh, bh, stem)))
        self.conv_flops += 2 * B * (H // 2) * (W // 2) * C * 27
        d2 = self._buf(H // 4, W // 4, 2 * C)
        self._conv_bn(sd, "backbone.dark2.0", stem, d2, 2)
        d2o = self._buf(H // 4, W // 4, 2 * C)
        self._c2f(sd, "backbone.dark2.1", d2, d2o, d, True)
        if self.attention:
            self._cbam(sd, "backbone.dark2.2", d2o)
        d3 = self._buf(H // 8, W // 8, 4 * C)
        self._conv_bn(sd, "backbone.dark3.0", d2o, d3, 2)
        p3 = self._buf(H // 8, W // 8, 4 * C)
        self._c2f(sd, "backbone.dark3.1", d3, p3, 2 * d, True)
        if self.attention:
            self._self_attention(sd, "backbone.dark3.2", p3)
        d4 = self._buf(H // 16, W // 16, 8 * C)
        self._conv_bn(sd, "backbone.dark4.0", p3, d4, 2)
        p4 = self._buf(H // 16, W // 16, 8 * C)
        self._c2f(sd, "backbone.dark4.1", d4, p4, 2 * d, True)
        if self.attention:
            self._cbam(sd, "backbone.dark4.2", p4)
        d5 = self._buf(H // 32, W // 32, C5)
        self._conv_bn(sd, "backbone.dark5.0", p4, d5, 2)
        d5o = self._buf(H // 32, W // 32, C5)
        self._c2f(sd, "backbone.dark5.1", d5, d5o, d, True)
        # SPPF (model/blocks.py:138-142): cv1 -> slot 0 of the 4*c_ buffer, pools -> slots 1..3, cv2 reads all
        c_ = C5 // 2
        sp = self._buf(H // 32, W // 32, 4 * c_)
        self._conv_bn(sd, "backbone.dark5.2.cv1", d5o, sp.sub(0, c_))
        self.ops.append(("pool", "backbone.dark5.2.m", (sp, c_)))
        # neck concat buffers (model/neck.py:57-60): cat6 = [h5(h4) | p5], cat4 = [h3(h2) | h1]
        cat6 = self._buf(H // 32, W // 32, 8 * C + C5)
        p5 = cat6.sub(8 * C, C5)
        self._conv_bn(sd, "backbone.dark5.2.cv2", sp, p5)
        cat4 = self._buf(H // 16, W // 16, 4 * C + 8 * C)
        h1 = cat4.sub(4 * C, 8 * C)
        # ---- neck
        self._c2f(sd, "neck.h1", p4, h1, d, False, up_src=p5)          # cat([up(p5), p4])
        h2 = self._buf(H // 8, W // 8, 4 * C)
        self._c2f(sd, "neck.h2", p3, h2, d, False, up_src=h1)          # cat([up(h1), p3])
        self._conv_bn(sd, "neck.h3", h2, cat4.sub(0, 4 * C), 2)
        h4 = self._buf(H // 16, W // 16, 8 * C)
        self._c2f(sd, "neck.h4", cat4, h4, d, False)                   # cat([h3(h2), h1])
        self._conv_bn(sd, "neck.h5", h4, cat6.sub(0, 8 * C), 2)
        h6 = self._buf(H // 32, W // 32, C5)
        self._c2f(sd, "neck.h6", cat6, h6, d, False)                   # cat([h5(h4), p5])
        self.features = {"p3": p3, "p4": p4, "p5": p5, "h2": h2, "h4": h4, "h6": h6}
        # ---- head (model/head.py:19-51)
        feats = (h2, h4, h6)
        c1 = max(feats[0].c, nc)
        c2 = max(feats[0].c // 4, 64)
        ncp = (nc + 15) // 16 * 16           # class conv output padded to the MMA N granularity
        self.raw_pitch = 64 + ncp
        self.raw: List[torch.Tensor] = []
        fuse_head0 = os.environ.get("TOD_FUSE_HEAD0", "1") != "0" and not self.attention   # (a CBAM follows each .0 otherwise)
        for i, f in enumerate(feats):
            raw = torch.zeros((B, f.h, f.w, self.raw_pitch), dtype=torch.float32, device=dev)
            self.raw.append(raw)
            rv = View(raw, 0, self.raw_pitch)
            tb2 = self._buf(f.h, f.w, c2)
            if fuse_head0:
                # box.i.0 and cls.i.0 read the same feature map (model/head.py:26,37): ONE conv with the two weight sets
                # stacked along cout writes [box c2 | cls c1] and the towers' second convs read channel windows of it
                t01 = self._buf(f.h, f.w, c2 + c1)
                tb1, tc1 = t01.sub(0, c2), t01.sub(c2, c1)
                (wb_, bb_), (wc_, bc_) = fold_conv_bn(sd, f"head.box.{i}.0"), fold_conv_bn(sd, f"head.cls.{i}.0")
                self._conv(f"head.boxcls.{i}.0", torch.cat([wb_, wc_]), torch.cat([bb_, bc_]), f, t01)
            else:
                tb1, tc1 = self._buf(f.h, f.w, c2), self._buf(f.h, f.w, c1)
                self._conv_bn(sd, f"head.box.{i}.0", f, tb1)
            if self.attention:
                self._cbam(sd, f"head.box.{i}.1", tb1)
            self._conv_bn(sd, f"head.box.{i}.2", tb1, tb2)
            if self.attention:
                self._cbam(sd, f"head.box.{i}.3", tb2)
            self._conv(f"head.box.{i}.4", _t(sd, f"head.box.{i}.4.weight"), _t(sd, f"head.box.{i}.4.bias"), tb2,
                       rv.sub(0, 64), act=TOD_ACT_NONE, out_f32=True)
            tc2 = self._buf(f.h, f.w, c1)
            if not fuse_head0:
                self._conv_bn(sd, f"head.cls.{i}.0", f, tc1)
            if self.attention:
                self._cbam(sd, f"head.cls.{i}.1", tc1)
            self._conv_bn(sd, f"head.cls.{i}.2", tc1, tc2)
            if self.attention:
                self._cbam(sd, f"head.cls.{i}.3", tc2)
            wc = torch.zeros(ncp, c1, 1, 1)
            bc = torch.zeros(ncp)
            wc[:nc] = _t(sd, f"head.cls.{i}.4.weight")
            bc[:nc] = _t(sd, f"head.cls.{i}.4.bias")
            self._conv(f"head.cls.{i}.4", wc, bc, tc2, rv.sub(64, ncp), act=TOD_ACT_NONE, out_f32=True)
        self.level_shapes = [(f.h, f.w) for f in feats]
        self.anchors = sum(h * w for h, w in self.level_shapes)
        A = self.anchors
        # ---- decode + NMS buffers
        self.head_out = torch.zeros((B, 4 + nc, A), dtype=torch.float32, device=dev)
        self.decoded = torch.zeros((B, A, 4 + nc), dtype=torch.float32, device=dev)
        self.cand_box = torch.zeros((B, A, 4), dtype=torch.float32, device=dev)
        self.cand_conf = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self.cand_cls = torch.zeros((B, A), dtype=torch.int32, device=dev)
        self.nms_work = torch.zeros(int(self.L.tod_nms_workspace_bytes(B, A)), dtype=torch.uint8, device=dev)
        self.keep_idx = torch.zeros((B, A), dtype=torch.int32, device=dev)
        self.keep_count = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.dets = torch.zeros((B, A, 6), dtype=torch.float32, device=dev)
        self.launches_forward = len(self.ops)
        # fused head outputs (detect path): the last conv of each tower decodes its own accumulator rows into the NMS
        # candidates (tod_conv2d_head_decode); the raw maps and the decode kernel are then not needed
        self.head_fuse: Dict[str, HeadFuseDesc] = {}
        off = 0
        for i, (h, w) in enumerate(self.level_shapes):
            for tower, mode in (("box", TOD_FUSE_BOX), ("cls", TOD_FUSE_CLS)):
                f = HeadFuseDesc()
                f.mode, f.nc, f.level_off, f.anchors = mode, nc, off, A
                f.in_h, f.in_w, f.stride = H, W, float(H // h)
                f.d_cand_box, f.d_cand_conf, f.d_cand_cls = (self.cand_box.data_ptr(), self.cand_conf.data_ptr(),
                                                             self.cand_cls.data_ptr())
                self.head_fuse[f"head.{tower}.{i}.4"] = f
            off += h * w
        self.fuse_head_decode = True    # graph_for: candidates straight from the head convs
        # fused 1x1 tails (tod_conv2d_tail1x1): a conv with 64 output channels whose ONLY consumer is a 1x1 64 -> 64 conv
        # (backbone.dark2[0] -> dark2[1].cv1 at scale s: 210 MB written and read back per batch-64 pass otherwise)
        self.tail_fuse: Dict[str, Tuple[ConvTailDesc, str]] = {}
        self.tail_skip = set()
        if os.environ.get("TOD_FUSE_TAIL", "1") != "0":
            a, b_ = self.conv_meta.get("backbone.dark2.0"), self.conv_meta.get("backbone.dark2.1.cv1")
            if (a is not None and b_ is not None and a["dst"].c == 64 and a["dst"].pitch == 64 and b_["src"].ptr == a["dst"].ptr
                    and b_["src"].c == 64 and b_["dst"].c == 64 and b_["w"].shape[-1] == 1 and b_["upadd"] is None
                    and b_["residual"] is None and not b_["out_f32"] and b_["act"] == TOD_ACT_SILU and a["act"] == TOD_ACT_SILU):
                db = next(p for k, n, p in self.ops if n == "backbone.dark2.1.cv1")
                t = ConvTailDesc()
                t.d_w2, t.d_bias2, t.d_out2 = db.d_w, db.d_bias, db.d_out
                t.cout2, t.out2_pitch, t.act2 = 64, db.out_pitch, TOD_ACT_SILU
                self.tail_fuse["backbone.dark2.0"] = (t, "backbone.dark2.1.cv1")
                self.tail_skip.add("backbone.dark2.1.cv1")
        # box towers on the fused-decode path (tod_conv2d_tail1x1_box_decode): Conv3x3 64 -> 64 + the bare 64 -> 64 logit
        # conv + DFL / dist2bbox in one kernel per level (model/head.py:36-42,53-61)
        self.tail_box: Dict[str, Tuple[ConvTailDesc, HeadFuseDesc]] = {}
        self.tail_box_skip = set()
        if os.environ.get("TOD_FUSE_TAIL", "1") != "0" and not self.attention:   # (a CBAM sits between .2 and .4 otherwise)
            for i in range(len(self.level_shapes)):
                na, nb = f"head.box.{i}.2", f"head.box.{i}.4"
                a, b_ = self.conv_meta[na], self.conv_meta[nb]
                if a["dst"].c == 64 and a["src"].c % 16 == 0 and b_["src"].ptr == a["dst"].ptr and b_["w"].shape[0] == 64:
                    db = next(p for k, n, p in self.ops if n == nb)
                    t = ConvTailDesc()
                    t.d_w2, t.d_bias2, t.d_out2 = db.d_w, db.d_bias, None
                    t.cout2, t.out2_pitch, t.act2 = 64, 64, TOD_ACT_NONE
                    self.tail_box[na] = (t, self.head_fuse[nb])
                    self.tail_box_skip.add(nb)

        if os.environ.get("TOD_SNAKE", "1") != "0":
            self._assign_tile_order()

    def _assign_tile_order(self) -> None:
        """Consecutive convs walk the batch in opposite directions (TOD_CONV_REVERSE on every other one): a layer then starts
        on the images its producer wrote LAST, which are the ones still in the 126 MB L2 (a batch-64 activation at 160^2 or
        80^2 is 105-420 MB, so a same-direction consumer always starts on evicted lines).  The stem writes first-to-last."""
        rev, direction = False, {}
        for kind, name, payload in self.ops:
            if kind != "conv" or name.startswith("head.") or name in self.tail_skip:
                continue
            rev = not rev
            payload.flags = (payload.flags | TOD_CONV_REVERSE) if rev else (payload.flags & ~TOD_CONV_REVERSE)
            direction[name] = rev
        for lvl, feat in enumerate(("neck.h2.cv2", "neck.h4.cv2", "neck.h6.cv2")):
            base = direction.get(feat, False)
            for kind, name, payload in self.ops:
                if kind == "conv" and name.startswith(f"head.boxcls.{lvl}."):
                    base = not base
                    payload.flags = (payload.flags | TOD_CONV_REVERSE) if base else (payload.flags & ~TOD_CONV_REVERSE)
            for tower in ("box", "cls"):
                r = base
                for kind, name, payload in self.ops:
                    if kind == "conv" and name.startswith(f"head.{tower}.{lvl}."):
                        r = not r
                        payload.flags = (payload.flags | TOD_CONV_REVERSE) if r else (payload.flags & ~TOD_CONV_REVERSE)

    # ------------------------------------------------------------------ execution
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def input_buffer(self, kind: str = "f32", slot: int = 0) -> torch.Tensor:
        """Static device input of the captured graphs: kind "f32" = float32 NCHW in [0, 1] (the reference's tensor),
        kind "u8" = uint8 NHWC letterboxed RGB (the /255 is fused into the stem).  Two slots allow the upload of batch
        i+1 to overlap the graph replay of batch i."""
        if kind == "f32" and slot == 0:
            return self.x_static
        key = (kind, slot)
        if key not in self._inputs:
            shape = (self.batch, 3, self.in_h, self.in_w) if kind == "f32" else (self.batch, self.in_h, self.in_w, 3)
            self._inputs[key] = torch.zeros(shape, dtype=torch.float32 if kind == "f32" else torch.uint8, device=self.device)
        return self._inputs[key]

    def packed_buffers(self) -> dict:
        """Result buffers of the captured graphs (tod_pack_detections): ONE device buffer = int32 offsets[B + 1] (padded to
        `hdr` bytes) followed by the compacted rows, and its pinned host mirror, which the graph itself fills with a single
        fixed-size device-to-host copy of the header and the first `cap_rows` rows (collect() fetches the rest only when a
        batch holds more)."""
        if self._packed is None:
            B, A = self.batch, self.anchors
            hdr = (4 * (B + 1) + 31) // 32 * 32
            cap_rows = min(B * A, max(B * 256, 4096))
            dev = torch.zeros(hdr + B * A * 24, dtype=torch.uint8, device=self.device)
            host = torch.zeros(hdr + cap_rows * 24, dtype=torch.uint8).pin_memory()
            work = torch.zeros(max(int(self.L.tod_pack_workspace_bytes(B, A)), 8), dtype=torch.uint8, device=self.device)
            self._packed = dict(dev=dev, host=host, hdr=hdr, cap_rows=cap_rows, work=work,
                                rows_corrected=torch.zeros((B, A, 6), dtype=torch.float32, device=self.device),
                                host_offsets=host[:4 * (B + 1)].view(torch.int32).numpy(),
                                host_rows=host[hdr:].view(torch.float32).view(cap_rows, 6).numpy())
        return self._packed

    def run_network(self, x: Optional[torch.Tensor] = None, fork: bool = False, fused_decode: bool = False) -> None:
        """Enqueue stem + every conv + SPPF pooling (the raw head maps land in self.raw).
        x: float32 (B, 3, H, W) in [0, 1]  or  uint8 (B, H, W, 3).
        fork: issue the six head towers (box / cls x three levels, mutually independent: model/head.py:24-44) on side
        streams as soon as their input feature exists, so that they overlap the rest of the neck and one another
        (used inside graph capture, where the forks become parallel graph branches).
        fused_decode: the last conv of every head tower writes the NMS candidates itself (no raw maps, no run_decode)."""
        main = torch.cuda.current_stream(self.device)
        L = self.L
        if x is None:
            x = self.x_static
        u8 = x.dtype == torch.uint8
        want = (self.batch, self.in_h, self.in_w, 3) if u8 else tuple(self.x_static.shape)
        assert x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.is_contiguous() and tuple(x.shape) == want, \
            (x.dtype, tuple(x.shape), want)
        def issue(kind, name, payload, stream):
            st = stream.cuda_stream
            if kind == "conv":
                if name in self.tail_skip or (fused_decode and name in self.tail_box_skip):
                    return
                if fused_decode and name in self.tail_box:
                    t, f = self.tail_box[name]
                    check(L.tod_conv2d_tail1x1_box_decode(C.byref(payload), C.byref(t), C.byref(f), st), name)
                elif name in self.tail_fuse:
                    check(L.tod_conv2d_tail1x1(C.byref(payload), C.byref(self.tail_fuse[name][0]), st), name)
                elif fused_decode and name in self.head_fuse:
                    check(L.tod_conv2d_head_decode(C.byref(payload), C.byref(self.head_fuse[name]), st), name)
                else:
                    check(L.tod_conv2d_nhwc_bf16(C.byref(payload), st), name)
            elif kind == "stem":
                w, b, out = payload
                fn = L.tod_stem_conv_nhwc_u8 if u8 else L.tod_stem_conv_nchw_f32
                check(fn(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.ptr, self.batch, self.in_h, self.in_w, self.C,
                         out.pitch, st), name)
            elif kind == "pool":
                buf, c_ = payload
                check(L.tod_sppf_pool_nhwc_bf16(buf.ptr, self.batch, buf.h, buf.w, c_, buf.pitch, st), name)
            elif kind == "cbam":
                check(L.tod_cbam_nhwc_bf16(C.byref(payload), st), name)
            elif kind == "attn":
                self._run_attention(payload, st)
            else:  # pragma: no cover
                raise AssertionError(kind)

        if not fork:
            for kind, name, payload in self.ops:
                issue(kind, name, payload, main)
            return
        feature_done = {"neck.h2.cv2": 0, "neck.h4.cv2": 1, "neck.h6.cv2": 2}     # last op of each head input
        joins = []
        for kind, name, payload in self.ops:
            if name.startswith("head."):
                continue
            issue(kind, name, payload, main)
            lvl = feature_done.get(name)
            if lvl is None:
                continue
            ready = torch.cuda.Event()
            ready.record(main)
            shared = [(k2, n2, p2) for k2, n2, p2 in self.ops if n2.startswith(f"head.boxcls.{lvl}.")]   # TOD_FUSE_HEAD0
            for tower in ("cls", "box") if shared else ("box", "cls"):
                # stream priorities become kernel-node priorities in the captured graph: the big level-0 towers are
                # throughput work that fills gaps, the small deep-level towers are the tail of the critical path
                if (tower, lvl) not in self._side_streams:
                    prio = 0 if (lvl == 0 or not self.prioritise_critical_path) else -1
                    self._side_streams[(tower, lvl)] = torch.cuda.Stream(self.device, priority=prio)
                side = self._side_streams[(tower, lvl)]
                side.wait_event(ready)
                if shared and tower == "cls":          # the stacked first conv runs on the class tower's stream ...
                    for k2, n2, p2 in shared:
                        issue(k2, n2, p2, side)
                    ready = torch.cuda.Event()         # ... and the box tower starts after it
                    ready.record(side)
                prefix = f"head.{tower}.{lvl}."
                for k2, n2, p2 in self.ops:
                    if n2.startswith(prefix):
                        issue(k2, n2, p2, side)
                done = torch.cuda.Event()
                done.record(side)
                joins.append(done)
        for done in joins:
            main.wait_event(done)

    def run_decode(self, head_out: bool = True, decoded: bool = False, candidates: bool = True) -> None:
        d = DecodeDesc()
        for i, (h, w) in enumerate(self.level_shapes):
            d.d_raw[i] = self.raw[i].data_ptr()
            d.h[i], d.w[i] = h, w
            d.stride[i] = float(self.in_h // h)
        d.raw_pitch, d.batch, d.nc, d.in_h, d.in_w = self.raw_pitch, self.batch, self.nc, self.in_h, self.in_w
        d.d_head_out = self.head_out.data_ptr() if head_out else None
        d.d_decoded = self.decoded.data_ptr() if decoded else None
        if candidates:
            d.d_cand_box, d.d_cand_conf, d.d_cand_cls = (self.cand_box.data_ptr(), self.cand_conf.data_ptr(),
                                                         self.cand_cls.data_ptr())
        check(self.L.tod_head_decode(C.byref(d), self._stream()), "tod_head_decode")

    def run_nms(self, conf_thres: float, nms_thres: float) -> None:
        check(self.L.tod_nms(self.cand_box.data_ptr(), self.cand_conf.data_ptr(), self.cand_cls.data_ptr(), self.batch,
                             self.anchors, float(np.float32(conf_thres)), float(nms_thres), self.nms_work.data_ptr(),
                             self.nms_work.numel(), self.keep_idx.data_ptr(), self.keep_count.data_ptr(),
                             self.dets.data_ptr(), self._stream()), "tod_nms")

    # number of kernels one full pass enqueues (forward ops + decode + 3 NMS kernels)
    @property
    def launches_per_pass(self) -> int:
        extra = sum(4 if k == "cbam" else ((4 if p["fused"] else 6 * self.batch - 1) if k == "attn" else 0)
                    for k, _, p in self.ops)
        return (len(self.ops) + extra - len(self.tail_skip) - (len(self.tail_box_skip) if self.fuse_head_decode else 0)
                + (0 if self.fuse_head_decode else 1) + 3 + 1)      # ... + 3 NMS kernels + result packing

    def box_params(self) -> torch.Tensor:
        """(B, 6) float64 device buffer read by tod_correct_boxes in the `corrected` graphs: per image
        offset_y, offset_x, scale_y, scale_x, image_h, image_w (model.py:box_correction_params)."""
        if self._box_params is None:
            self._box_params = torch.zeros((self.batch, 6), dtype=torch.float64, device=self.device)
            self._box_params[:, 2:4] = 1.0
        return self._box_params

    def run_pack(self, corrected: int = -1, max_boxes: int = 0) -> None:
        """NMS rows -> (un-letterboxed rows) -> packed result buffer -> pinned host mirror (one fixed-size D2H copy).
        corrected = 0 / 1: tod_correct_boxes with letterbox off / on (parameters in box_params()); -1: raw NMS rows."""
        pk = self.packed_buffers()
        rows = self.dets
        if corrected >= 0:
            rows = pk["rows_corrected"]
            check(self.L.tod_correct_boxes(self.dets.data_ptr(), self.keep_count.data_ptr(), self.batch, self.anchors,
                                           self.box_params().data_ptr(), int(corrected), rows.data_ptr(), self._stream()),
                  "tod_correct_boxes")
        dev = pk["dev"]
        check(self.L.tod_pack_detections(rows.data_ptr(), self.keep_count.data_ptr(), self.batch, self.anchors, int(max_boxes),
                                         dev.data_ptr(), dev.data_ptr() + pk["hdr"], pk["work"].data_ptr(), pk["work"].numel(),
                                         self._stream()), "tod_pack_detections")
        pk["host"].copy_(dev[:pk["host"].numel()], non_blocking=True)

    def graph_for(self, kind: str, slot: int, conf_thres: float, nms_thres: float, corrected: int = -1, max_boxes: int = 0):
        """CUDA graph of network + decode + NMS + result packing (+ the D2H copy of the packed rows into the pinned host
        mirror) on the static input (kind, slot); captured on first use.  corrected = 0 / 1: rows un-letterboxed on the
        device; max_boxes > 0: the reference detect loop's top-k (utils/callbacks.py:159-166) on the device."""
        key = (kind, slot, float(conf_thres), float(nms_thres), int(corrected), int(max_boxes))
        g = self._graphs.get(key)
        if g is not None:
            return g
        x = self.input_buffer(kind, slot)
        self.packed_buffers()

        def body():
            self.run_network(x, fork=self.fork_head, fused_decode=self.fuse_head_decode)
            if not self.fuse_head_decode:
                self.run_decode(False, False, True)
            self.run_nms(conf_thres, nms_thres)
            self.run_pack(corrected, max_boxes)

        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            body()           # warm-up outside capture (function attributes, module loading)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(self.device, priority=-1 if self.prioritise_critical_path else 0)   # backbone + neck chain
        with torch.cuda.graph(g, stream=cap):
            body()
        self._graphs[key] = g
        return g

    def raw_maps_nchw(self) -> List[torch.Tensor]:
        """Training-mode Head output layout (B, 64+nc, h, w) (model/head.py:50-51) as views of the raw maps."""
        return [r[..., :64 + self.nc].permute(0, 3, 1, 2) for r in self.raw]

    def feature_nchw(self, name: str) -> torch.Tensor:
        return self.features[name].tensor().permute(0, 3, 1, 2).float()
