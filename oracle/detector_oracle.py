"""CPU oracle: fp32 restatement of the reference detector inference hot path.

TEST INFRASTRUCTURE -- NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.  The product path
(transparent_object_detection_b200) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so this
restatement is pinned against the reference's OWN code executed in the authoring container
(oracle/make_golden.py imports /root/reference and writes tests/golden/*.npz; the
`-m "not gpu"` tests compare this file with those fixtures).  Third-party arithmetic on the
path: torchvision.ops.nms (unpinned by the reference; fixtures generated with
torchvision 0.26.0+cu128 CPU kernel) -- restated in `nms_greedy` below and, in plain C, in
oracle/nms_oracle.c (built by oracle/Makefile, wrapped by oracle/nms_c.py).

Every function cites the reference file:line it follows.  Weights come in as a dict with
the reference's state_dict key layout (oracle/synth.py:state_dict_shapes).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, model/blocks.py:49


def _t(sd, key) -> torch.Tensor:
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


# ----------------------------------------------------------------------------- blocks
def conv_bn_silu(sd, prefix: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """reference Conv.forward = act(norm(conv(x))), model/blocks.py:52-54; pad = k//2 (autopad :6-19)."""
    w = _t(sd, prefix + ".conv.weight")
    k = w.shape[-1]
    y = F.conv2d(x, w, None, stride=stride, padding=k // 2)
    y = F.batch_norm(y, _t(sd, prefix + ".norm.running_mean"), _t(sd, prefix + ".norm.running_var"),
                     _t(sd, prefix + ".norm.weight"), _t(sd, prefix + ".norm.bias"), False, 0.0, BN_EPS)
    return F.silu(y)


def conv_bn_silu_bf16(sd, prefix: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """The same op with the build's rounding points: BN folded into the weights (fold_bn), weights
    and input activations rounded to bf16, fp32 accumulation + fp32 bias + SiLU, output rounded to bf16.  Used by the
    tests to separate the deviation that bf16 storage itself causes on a given random-init network from kernel error."""
    bf = lambda t: t.to(torch.bfloat16).float()
    w, b = fold_bn(sd, prefix)
    y = F.conv2d(bf(x), bf(w), None, stride=stride, padding=w.shape[-1] // 2) + b.view(1, -1, 1, 1)
    return bf(F.silu(y))


class bf16_emulation:
    """with bf16_emulation(): every Conv of backbone / neck / head towers evaluates through conv_bn_silu_bf16."""

    def __enter__(self):
        global conv_bn_silu
        self._saved = conv_bn_silu
        conv_bn_silu = conv_bn_silu_bf16
        return self

    def __exit__(self, *exc):
        global conv_bn_silu
        conv_bn_silu = self._saved
        return False


def fold_bn(sd, prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference fuse_conv, model/blocks.py:179-185: W' = diag(g/sqrt(eps+var)) W, b' = beta - g*mean/sqrt(var+eps)."""
    w = _t(sd, prefix + ".conv.weight")
    g, beta = _t(sd, prefix + ".norm.weight"), _t(sd, prefix + ".norm.bias")
    mean, var = _t(sd, prefix + ".norm.running_mean"), _t(sd, prefix + ".norm.running_var")
    w_norm = torch.diag(g.div(torch.sqrt(BN_EPS + var)))
    wf = torch.mm(w_norm, w.reshape(w.shape[0], -1)).view(w.shape)
    bf = beta - g.mul(mean).div(torch.sqrt(var + BN_EPS))
    return wf, bf


def bottleneck(sd, prefix: str, x: torch.Tensor, shortcut: bool) -> torch.Tensor:
    """reference Bottleneck.forward, model/blocks.py:80-82 (inside C2f: k=(3,3),(3,3), e=1.0, c1==c2)."""
    y = conv_bn_silu(sd, prefix + ".cv2", conv_bn_silu(sd, prefix + ".cv1", x))
    return x + y if shortcut else y


def c2f(sd, prefix: str, x: torch.Tensor, n: int, shortcut: bool) -> torch.Tensor:
    """reference C2f.forward, model/blocks.py:104-108."""
    y = list(conv_bn_silu(sd, prefix + ".cv1", x).chunk(2, 1))
    for j in range(n):
        y.append(bottleneck(sd, f"{prefix}.m.{j}", y[-1], shortcut))
    return conv_bn_silu(sd, prefix + ".cv2", torch.cat(y, 1))


def sppf(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """reference SPPF.forward, model/blocks.py:138-142: three chained MaxPool2d(5, 1, 2)."""
    y = [conv_bn_silu(sd, prefix + ".cv1", x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    return conv_bn_silu(sd, prefix + ".cv2", torch.cat(y, 1))


def sppf_pools(x: torch.Tensor) -> torch.Tensor:
    """The pooling part alone: cat(x, p5, p9, p13) along channels (model/blocks.py:139-141)."""
    y = [x]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    return torch.cat(y, 1)


# ----------------------------------------------------------------------------- attention blocks
def cbam(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """reference CBAM.forward, model/blocks.py:206-223: channel attention sigmoid(fc2(relu(fc1(avgpool))) +
    fc2(relu(fc1(maxpool)))) (1x1 convs without bias, reduction 16), then spatial attention
    sigmoid(conv7x7(cat[mean_c, max_c])) (padding 3, no bias) on the channel-scaled tensor."""
    w1, w2, w7 = _t(sd, prefix + ".fc1.weight"), _t(sd, prefix + ".fc2.weight"), _t(sd, prefix + ".conv.weight")
    mlp = lambda v: F.conv2d(F.relu(F.conv2d(v, w1)), w2)
    ca = torch.sigmoid(mlp(F.adaptive_avg_pool2d(x, 1)) + mlp(F.adaptive_max_pool2d(x, 1)))
    x = x * ca
    st = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
    return x * torch.sigmoid(F.conv2d(st, w7, None, padding=w7.shape[-1] // 2))


def self_attention(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """reference SelfAttention.forward, model/blocks.py:236-254: q, k = 1x1 convs to C/8 (+bias), v = 1x1 conv to C (+bias);
    out[c, i] = sum_j v[c, j] * softmax_j(q_i . k_j); gamma * out + x."""
    b, c, h, w = x.shape
    q = F.conv2d(x, _t(sd, prefix + ".query.weight"), _t(sd, prefix + ".query.bias")).view(b, -1, h * w).permute(0, 2, 1)
    k = F.conv2d(x, _t(sd, prefix + ".key.weight"), _t(sd, prefix + ".key.bias")).view(b, -1, h * w)
    v = F.conv2d(x, _t(sd, prefix + ".value.weight"), _t(sd, prefix + ".value.bias")).view(b, -1, h * w)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    out = torch.bmm(v, att.permute(0, 2, 1)).view(b, c, h, w)
    return _t(sd, prefix + ".gamma") * out + x


# ----------------------------------------------------------------------------- network
def backbone(sd, x: torch.Tensor, d: int, attention: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """reference Backbone.forward, model/backbone.py:50-59.  attention=False: the plain topology (attention modules =
    Identity); attention=True: the current source with CBAM after dark2 / dark4's C2f and SelfAttention after
    dark3's (:26,33,40)."""
    x = conv_bn_silu(sd, "backbone.stem", x, 2)
    x = c2f(sd, "backbone.dark2.1", conv_bn_silu(sd, "backbone.dark2.0", x, 2), d, True)
    if attention:
        x = cbam(sd, "backbone.dark2.2", x)
    x = c2f(sd, "backbone.dark3.1", conv_bn_silu(sd, "backbone.dark3.0", x, 2), 2 * d, True)
    if attention:
        x = self_attention(sd, "backbone.dark3.2", x)
    feat1 = x
    x = c2f(sd, "backbone.dark4.1", conv_bn_silu(sd, "backbone.dark4.0", x, 2), 2 * d, True)
    if attention:
        x = cbam(sd, "backbone.dark4.2", x)
    feat2 = x
    x = c2f(sd, "backbone.dark5.1", conv_bn_silu(sd, "backbone.dark5.0", x, 2), d, True)
    feat3 = sppf(sd, "backbone.dark5.2", x)
    return feat1, feat2, feat3


def neck(sd, feats: Sequence[torch.Tensor], d: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """reference Neck.forward, model/neck.py:55-61 with C2f(shortcut=False) stages."""
    p3, p4, p5 = feats
    up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")  # nn.Upsample, neck.py:17
    h1 = c2f(sd, "neck.h1", torch.cat([up(p5), p4], 1), d, False)
    h2 = c2f(sd, "neck.h2", torch.cat([up(h1), p3], 1), d, False)
    h4 = c2f(sd, "neck.h4", torch.cat([conv_bn_silu(sd, "neck.h3", h2, 2), h1], 1), d, False)
    h6 = c2f(sd, "neck.h6", torch.cat([conv_bn_silu(sd, "neck.h5", h4, 2), p5], 1), d, False)
    return h2, h4, h6


def head_raw(sd, feats: Sequence[torch.Tensor], attention: bool = False) -> List[torch.Tensor]:
    """reference Head.forward up to the training-mode return, model/head.py:46-51: cat(box, cls) per level; attention=True:
    the current source's CBAM after each of the two Convs of a tower (:28,30,39,41)."""
    out = []
    for i, x in enumerate(feats):
        t = []
        for name in ("box", "cls"):
            p = f"head.{name}.{i}"
            y = conv_bn_silu(sd, p + ".0", x)
            if attention:
                y = cbam(sd, p + ".1", y)
            y = conv_bn_silu(sd, p + ".2", y)
            if attention:
                y = cbam(sd, p + ".3", y)
            t.append(F.conv2d(y, _t(sd, p + ".4.weight"), _t(sd, p + ".4.bias")))
        out.append(torch.cat(t, 1))
    return out


def make_anchors(shapes: Sequence[Tuple[int, int]], strides: Sequence[float],
                 offset: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference make_anchors, utils/bbox_utils.py:14-37 -> (A, 2) points, (A, 1) strides."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(st)


def dfl(box: torch.Tensor) -> torch.Tensor:
    """reference DFL.forward, model/blocks.py:154-157: softmax over 16 bins, expectation with arange(16)."""
    b, _, a = box.shape
    p = box.view(b, 4, 16, a).transpose(2, 1).softmax(1)
    w = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return F.conv2d(p, w).view(b, 4, a)


def head_decode(raw: Sequence[torch.Tensor], nc: int, strides: Sequence[float] = (8.0, 16.0, 32.0)) -> torch.Tensor:
    """reference Head.forward eval branch, model/head.py:53-61 -> (B, 4+nc, A), xywh in input pixels."""
    anchors, st = (t.transpose(0, 1) for t in make_anchors([r.shape[2:] for r in raw], strides))
    b = raw[0].shape[0]
    x = torch.cat([r.reshape(b, 64 + nc, -1) for r in raw], 2)
    box, cls = x.split((64, nc), 1)
    lt, rb = torch.split(dfl(box), 2, 1)
    a = anchors.unsqueeze(0) - lt
    bb = anchors.unsqueeze(0) + rb
    box = torch.cat(((a + bb) / 2, bb - a), 1)
    return torch.cat((box * st, cls.sigmoid()), 1)


def forward(sd, x: torch.Tensor, nc: int, d: int, training: bool = False, attention: bool = False):
    """reference BaseModel.forward, model/base.py:18-24 (head.stride = 8,16,32 per SURVEY F6).  attention=True: the
    current-source backbone and head (CBAM / SelfAttention) around the plain neck (the current-source neck does not run:
    SURVEY F3)."""
    raw = head_raw(sd, neck(sd, backbone(sd, x, d, attention), d), attention)
    return raw if training else head_decode(raw, nc)


def decode_box(head_out: torch.Tensor, input_shape: Tuple[int, int]) -> torch.Tensor:
    """reference DecodeBox.decode_box, utils/bbox_utils.py:66-82, applied to the eval head tensor.

    The reference method expects the upstream 5-tuple; on the head tensor the same arithmetic is
    permute(0,2,1) then xywh / (W,H,W,H) (verified bit-equal)."""
    y = head_out.permute(0, 2, 1).clone()
    y[:, :, :4] = y[:, :, :4] / torch.tensor(
        [input_shape[1], input_shape[0], input_shape[1], input_shape[0]], dtype=y.dtype)
    return y


def decode_box_tuple(dbox: torch.Tensor, cls: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor,
                     input_shape: Tuple[int, int]) -> torch.Tensor:
    """reference DecodeBox.decode_box on the upstream 5-tuple (dbox, cls, origin_cls, anchors, strides),
    utils/bbox_utils.py:75-82 with dist2bbox :51-57: dbox (B, 4, A) DFL distances, cls (B, nc, A) logits,
    anchors (2, A), strides (1, A) -> (B, A, 4+nc)."""
    lt, rb = torch.split(dbox, 2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    box = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
    y = torch.cat((box, cls.sigmoid()), 1).permute(0, 2, 1).clone()
    y[:, :, :4] = y[:, :, :4] / torch.tensor(
        [input_shape[1], input_shape[0], input_shape[1], input_shape[0]], dtype=y.dtype)
    return y


def loss_bbox_decode(anchor_points: torch.Tensor, pred_dist: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """reference Loss.bbox_decode, model/loss.py:333-337: pred_dist (B, A, 4 * reg_max) logits -> softmax over the bins of each
    side, `.matmul(proj)` with proj = arange(reg_max), then dist2bbox(xywh=False) (utils/bbox_utils.py:51-55 with the default
    dim -1) about anchor_points (A, 2) -> (B, A, 4) corners in grid units.  reg_max == 1 (use_dfl False): no softmax."""
    if reg_max > 1:
        b, a, c = pred_dist.shape
        pred_dist = pred_dist.view(b, a, 4, c // 4).softmax(3).matmul(torch.arange(reg_max, dtype=pred_dist.dtype))
    lt, rb = torch.split(pred_dist, 2, -1)
    return torch.cat((anchor_points - lt, anchor_points + rb), -1)


# ----------------------------------------------------------------------------- NMS
def nms_greedy(boxes: np.ndarray, scores: np.ndarray, iou_thr: float) -> np.ndarray:
    """torchvision.ops.nms CPU semantics (called at utils/bbox_utils.py:172), restated.

    float32 arithmetic in torchvision's operation order; stable descending sort; a later box is
    suppressed when float32 IoU, widened to double, is > the double threshold (verified against
    torchvision 0.26.0: IoU == f32(0.4) at thr 0.4 IS suppressed, IoU == f32(0.65) at thr 0.65 is
    not); zero-area pairs give NaN and survive."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    b = boxes.astype(np.float32, copy=False)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(i)
            rest = order[_i + 1:]
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(zero, xx2 - xx1)
            h = np.maximum(zero, yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            suppressed[rest[ovr.astype(np.float64) > float(iou_thr)]] = True
    return np.asarray(keep, dtype=np.int64)


def correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """reference DecodeBox.correct_boxes, utils/bbox_utils.py:84-117 (numpy, same dtype flow:
    box_hw scaled in place in float32, box_yx promoted to float64)."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.array(input_shape)
    image_shape = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(image_shape * np.min(input_shape / image_shape))
        offset = (input_shape - new_shape) / 2.0 / input_shape
        scale = input_shape / new_shape
        box_yx = (box_yx - offset) * scale
        box_hw *= scale
    box_mins = box_yx - (box_hw / 2.0)
    box_maxes = box_yx + (box_hw / 2.0)
    boxes = np.concatenate([box_mins[..., 0:1], box_mins[..., 1:2], box_maxes[..., 0:1], box_maxes[..., 1:2]], axis=-1)
    boxes *= np.concatenate([image_shape, image_shape], axis=-1)
    return boxes


def nms_keep_indices(prediction: np.ndarray, num_classes: int, conf_thres: float, nms_thres: float
                     ) -> List[np.ndarray]:
    """Anchor indices kept per image, in the reference's output order (class ascending, score
    descending within class) -- utils/bbox_utils.py:144-175.  `prediction` is (B, A, 4+nc) with
    normalised xywh; it is NOT modified."""
    pred = np.array(prediction, dtype=np.float32, copy=True)
    half = np.float32(2)
    corner = np.empty_like(pred[:, :, :4])
    corner[:, :, 0] = pred[:, :, 0] - pred[:, :, 2] / half
    corner[:, :, 1] = pred[:, :, 1] - pred[:, :, 3] / half
    corner[:, :, 2] = pred[:, :, 0] + pred[:, :, 2] / half
    corner[:, :, 3] = pred[:, :, 1] + pred[:, :, 3] / half
    out = []
    thr = np.float32(conf_thres)  # torch compares a float32 tensor with a python scalar in float32
    for i in range(pred.shape[0]):
        sc = pred[i, :, 4:4 + num_classes]
        cls = np.argmax(sc, axis=1)  # first max -> lowest class id on ties (torch.max CPU)
        conf = sc[np.arange(sc.shape[0]), cls]
        idx = np.nonzero(conf >= thr)[0]
        kept = []
        for c in np.unique(cls[idx]):
            seg = idx[cls[idx] == c]
            k = nms_greedy(corner[i, seg], conf[seg], nms_thres)
            kept.append(seg[k])
        out.append(np.concatenate(kept) if kept else np.zeros((0,), dtype=np.int64))
    return out


def non_max_suppression(prediction: np.ndarray, num_classes: int, input_shape, image_shape,
                        letterbox_image: bool, conf_thres: float = 0.5, nms_thres: float = 0.4,
                        keep_fn=None) -> List[Optional[np.ndarray]]:
    """reference DecodeBox.non_max_suppression, utils/bbox_utils.py:119-182.

    Like the reference it rewrites prediction[:, :, :4] to corner form IN PLACE (:144-149) and
    returns, per image, None or float32 (n, 6) rows [y1, x1, y2, x2, conf, cls] in image pixels.
    `keep_fn` swaps the selection step for another restatement with the same contract (oracle/nms_c.py)."""
    keep = (keep_fn or nms_keep_indices)(prediction, num_classes, conf_thres, nms_thres)
    half = np.float32(2)
    xywh = prediction[:, :, :4].copy()
    prediction[:, :, 0] = xywh[:, :, 0] - xywh[:, :, 2] / half
    prediction[:, :, 1] = xywh[:, :, 1] - xywh[:, :, 3] / half
    prediction[:, :, 2] = xywh[:, :, 0] + xywh[:, :, 2] / half
    prediction[:, :, 3] = xywh[:, :, 1] + xywh[:, :, 3] / half
    output: List[Optional[np.ndarray]] = [None] * len(prediction)
    for i, k in enumerate(keep):
        if k.size == 0:
            continue
        sc = prediction[i, k, 4:4 + num_classes]
        cls = np.argmax(sc, axis=1)
        det = np.empty((k.size, 6), dtype=np.float32)
        det[:, :4] = prediction[i, k, :4]
        det[:, 4] = sc[np.arange(k.size), cls]
        det[:, 5] = cls.astype(np.float32)
        box_xy, box_wh = (det[:, 0:2] + det[:, 2:4]) / 2, det[:, 2:4] - det[:, 0:2]
        det[:, :4] = correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image)
        output[i] = det
    return output


def detect(sd, images: torch.Tensor, nc: int, d: int, image_shape, letterbox_image: bool,
           conf_thres: float, nms_thres: float) -> List[Optional[np.ndarray]]:
    """The in-repo detect pipeline, utils/callbacks.py:147-154: net -> decode_box -> non_max_suppression."""
    h, w = images.shape[2:]
    with torch.no_grad():
        y = decode_box(forward(sd, images, nc, d), (h, w)).numpy()
    return non_max_suppression(y, nc, (h, w), image_shape, letterbox_image, conf_thres, nms_thres)
