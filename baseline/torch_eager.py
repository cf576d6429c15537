"""Same-box LIBRARY baseline: what the reference's detect loop costs on this GPU when every op goes to the vendor
libraries -- eager PyTorch, bf16, channels_last (cuDNN convolutions with autotune, ATen BatchNorm / SiLU / max-pool /
upsample / cat / softmax) followed by the reference's own post-processing loop with torchvision's CUDA NMS
("the kernel to beat on the same box").

NOT the product and NOT the checker: bench.py times it next to the product arm (key `library_baseline`) so that the
hand-written kernels are compared with cuDNN + torchvision and not only with a CPU.  It mirrors the reference's module
structure so that a reference state_dict loads:

    Conv        act(norm(conv(x)))                                    model/blocks.py:22-58
    Bottleneck  x + cv2(cv1(x))                                       model/blocks.py:61-82
    C2f         cv1 -> chunk -> n Bottlenecks -> cat -> cv2           model/blocks.py:85-116
    SPPF        cv1 -> 3 chained MaxPool2d(5, 1, 2) -> cat -> cv2     model/blocks.py:119-142
    backbone / neck (C2f stages) / head                    model/backbone.py:17-59, neck.py:17-61, head.py:11-61
    detect loop decode_box -> non_max_suppression                     utils/bbox_utils.py:66-82, 119-182
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.norm = nn.BatchNorm2d(c2)

    def forward(self, x):
        return F.silu(self.norm(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut):
        super().__init__()
        self.cv1, self.cv2, self.add = Conv(c, c, 3), Conv(c, c, 3), shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n, shortcut):
        super().__init__()
        self.c = c2 // 2
        self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1), Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1, self.cv2 = Conv(c1, c1 // 2, 1), Conv(c1 * 2, c2, 1)

    def forward(self, x):
        x = self.cv1(x)
        y1 = F.max_pool2d(x, 5, 1, 2)
        y2 = F.max_pool2d(y1, 5, 1, 2)
        return self.cv2(torch.cat((x, y1, y2, F.max_pool2d(y2, 5, 1, 2)), 1))


class EagerDetector(nn.Module):
    """Plain-topology detector with the reference's parameter names (SURVEY 8b state-dict layout)."""

    def __init__(self, nc: int, C: int, d: int, deep_mul: float):
        super().__init__()
        C5 = int(C * 16 * deep_mul)
        self.nc = nc
        bb = nn.Module()
        bb.stem = Conv(3, C, 3, 2)
        bb.dark2 = nn.Sequential(Conv(C, 2 * C, 3, 2), C2f(2 * C, 2 * C, d, True))
        bb.dark3 = nn.Sequential(Conv(2 * C, 4 * C, 3, 2), C2f(4 * C, 4 * C, 2 * d, True))
        bb.dark4 = nn.Sequential(Conv(4 * C, 8 * C, 3, 2), C2f(8 * C, 8 * C, 2 * d, True))
        bb.dark5 = nn.Sequential(Conv(8 * C, C5, 3, 2), C2f(C5, C5, d, True), SPPF(C5, C5))
        self.backbone = bb
        nk = nn.Module()
        nk.h1, nk.h2 = C2f(C5 + 8 * C, 8 * C, d, False), C2f(12 * C, 4 * C, d, False)
        nk.h3, nk.h4 = Conv(4 * C, 4 * C, 3, 2), C2f(12 * C, 8 * C, d, False)
        nk.h5, nk.h6 = Conv(8 * C, 8 * C, 3, 2), C2f(C5 + 8 * C, C5, d, False)
        self.neck = nk
        hd = nn.Module()
        filters = (4 * C, 8 * C, C5)
        c1, c2 = max(filters[0], nc), max(filters[0] // 4, 64)

        def tower(f, cm, co):     # Sequential indices 0 / 2 / 4 like the current source (1 and 3 are the attention slots)
            return nn.Sequential(Conv(f, cm, 3), nn.Identity(), Conv(cm, cm, 3), nn.Identity(), nn.Conv2d(cm, co, 1))

        hd.dfl = nn.Module()
        hd.dfl.conv = nn.Conv2d(16, 1, 1, bias=False)
        hd.cls = nn.ModuleList(tower(f, c1, nc) for f in filters)
        hd.box = nn.ModuleList(tower(f, c2, 64) for f in filters)
        self.head = hd
        self.strides = (8.0, 16.0, 32.0)

    def forward(self, x):
        b = self.backbone
        x = b.dark2(b.stem(x))
        p3 = b.dark3(x)
        p4 = b.dark4(p3)
        p5 = b.dark5(p4)
        n = self.neck
        h1 = n.h1(torch.cat((F.interpolate(p5, scale_factor=2.0, mode="nearest"), p4), 1))
        h2 = n.h2(torch.cat((F.interpolate(h1, scale_factor=2.0, mode="nearest"), p3), 1))
        h4 = n.h4(torch.cat((n.h3(h2), h1), 1))
        h6 = n.h6(torch.cat((n.h5(h4), p5), 1))
        feats = [torch.cat((self.head.box[i](f), self.head.cls[i](f)), 1) for i, f in enumerate((h2, h4, h6))]
        # eval branch of Head.forward (model/head.py:53-61) -- in float32 like the reference's decode arithmetic
        pts, st = [], []
        for f, s in zip(feats, self.strides):
            h, w = f.shape[2:]
            sy, sx = torch.meshgrid(torch.arange(h, device=f.device, dtype=torch.float32) + 0.5,
                                    torch.arange(w, device=f.device, dtype=torch.float32) + 0.5, indexing="ij")
            pts.append(torch.stack((sx, sy), -1).view(-1, 2))
            st.append(torch.full((h * w, 1), s, device=f.device, dtype=torch.float32))
        anchors, strides = torch.cat(pts).t(), torch.cat(st).t()
        y = torch.cat([f.reshape(f.shape[0], 64 + self.nc, -1) for f in feats], 2).float()
        box, cls = y.split((64, self.nc), 1)
        bsz, _, a = box.shape
        dist = (box.view(bsz, 4, 16, a).softmax(2) * torch.arange(16, device=y.device, dtype=torch.float32).view(1, 1, 16, 1)).sum(2)
        lt, rb = dist.split(2, 1)
        x1y1, x2y2 = anchors.unsqueeze(0) - lt, anchors.unsqueeze(0) + rb
        return torch.cat((torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides, cls.sigmoid()), 1)


def build(state_dict, nc: int, C: int, d: int, deep_mul: float, device, dtype=torch.bfloat16) -> EagerDetector:
    m = EagerDetector(nc, C, d, deep_mul)
    sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))) for k, v in state_dict.items()}
    m.load_state_dict(sd, strict=True)
    return m.eval().to(device=device, dtype=dtype).to(memory_format=torch.channels_last)


def detect(model: EagerDetector, images: torch.Tensor, input_shape, conf_thres: float, nms_thres: float) -> List[Optional[np.ndarray]]:
    """One batch through the reference's detect loop with library kernels only.  images: (B, 3, H, W) on the model's device
    and dtype, channels_last.  Rows [y1, x1, y2, x2, conf, cls] in input pixels (image_shape = input_shape)."""
    from torchvision.ops import nms
    H, W = input_shape
    with torch.no_grad():
        pred = model(images).permute(0, 2, 1).contiguous()                       # decode_box on the head tensor
        pred[:, :, :4] /= torch.tensor([W, H, W, H], dtype=pred.dtype, device=pred.device)
        xy, wh = pred[:, :, 0:2].clone(), pred[:, :, 2:4].clone()
        pred[:, :, 0:2], pred[:, :, 2:4] = xy - wh / 2, xy + wh / 2               # utils/bbox_utils.py:144-149
        out: List[Optional[np.ndarray]] = [None] * pred.shape[0]
        for i, ip in enumerate(pred):                                             # :151-180, per image / per class
            conf, cls = torch.max(ip[:, 4:], 1, keepdim=True)
            mask = conf[:, 0] >= conf_thres
            ip, conf, cls = ip[mask], conf[mask], cls[mask]
            if not ip.size(0):
                continue
            det = torch.cat((ip[:, :4], conf.float(), cls.float()), 1)
            kept = []
            for c in det[:, -1].unique():
                dc = det[det[:, -1] == c]
                kept.append(dc[nms(dc[:, :4], dc[:, 4], nms_thres)])
            o = torch.cat(kept).cpu().numpy()
            yx, hw = ((o[:, 0:2] + o[:, 2:4]) / 2)[:, ::-1], (o[:, 2:4] - o[:, 0:2])[:, ::-1]
            o[:, :4] = np.concatenate((yx - hw / 2, yx + hw / 2), 1) * np.array([H, W, H, W], dtype=np.float32)
            out[i] = o
    return out
