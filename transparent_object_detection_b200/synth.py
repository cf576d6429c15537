"""Deterministic synthetic weights / images / predictions for the detector hot path.

Input generation only (no checker code): bench.py, the tools and the tests use it; the oracle package re-exports it.

Everything here is numpy-seeded so that the same tensors can be regenerated on any
machine (this container, the GPU box) without the reference being importable.

The key -> shape table restates the *plain* topology of the reference detector
(SURVEY.md F4/F5): reference `BaseModel` (model/base.py:8-16) wired from
`Backbone` (model/backbone.py:17-48), the C2f-stage neck (model/neck.py:19-53 channel
comments) and `Head` (model/head.py:11-44) with the attention modules replaced by
Identity.  `oracle/make_golden.py` asserts that this table equals the state_dict of the
patched reference model key by key.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

SCALES = {  # (base_channels, base_depth, deep_mul)  -- config.yaml:3-8 via SURVEY.md section 8
    "n": (16, 1, 1.0),
    "s": (32, 1, 1.0),
    "m": (48, 2, 0.75),
    "l": (64, 3, 0.5),
    "x": (96, 3, 0.5),
}


def _conv_keys(table, prefix, c1, c2, k):
    """reference Conv = Conv2d(no bias) + BatchNorm2d (model/blocks.py:40-50)."""
    table[prefix + ".conv.weight"] = (c2, c1, k, k)
    table[prefix + ".norm.weight"] = (c2,)
    table[prefix + ".norm.bias"] = (c2,)
    table[prefix + ".norm.running_mean"] = (c2,)
    table[prefix + ".norm.running_var"] = (c2,)
    table[prefix + ".norm.num_batches_tracked"] = ()


def _c2f_keys(table, prefix, c1, c2, n):
    """reference C2f (model/blocks.py:98-102): cv1, cv2, then the m list."""
    c = int(c2 * 0.5)
    _conv_keys(table, prefix + ".cv1", c1, 2 * c, 1)
    _conv_keys(table, prefix + ".cv2", (2 + n) * c, c2, 1)
    for j in range(n):
        _conv_keys(table, f"{prefix}.m.{j}.cv1", c, c, 3)
        _conv_keys(table, f"{prefix}.m.{j}.cv2", c, c, 3)


def _cbam_keys(table, prefix, c):
    """reference CBAM (model/blocks.py:192-204): fc1, fc2 (1x1, no bias, reduction 16), conv (2 -> 1, 7x7, no bias)."""
    table[prefix + ".fc1.weight"] = (c // 16, c, 1, 1)
    table[prefix + ".fc2.weight"] = (c, c // 16, 1, 1)
    table[prefix + ".conv.weight"] = (1, 2, 7, 7)


def attention_shapes(nc: int, C: int, d: int, deep_mul: float) -> "OrderedDict[str, tuple]":
    """Extra keys of the CURRENT-SOURCE backbone / head (model/backbone.py:26,33,40, model/head.py:28,30,39,41)."""
    t: "OrderedDict[str, tuple]" = OrderedDict()
    _cbam_keys(t, "backbone.dark2.2", 2 * C)
    c = 4 * C
    for name, co in (("query", c // 8), ("key", c // 8), ("value", c)):
        t[f"backbone.dark3.2.{name}.weight"] = (co, c, 1, 1)
        t[f"backbone.dark3.2.{name}.bias"] = (co,)
    t["backbone.dark3.2.gamma"] = (1,)
    _cbam_keys(t, "backbone.dark4.2", 8 * C)
    c1, c2 = max(4 * C, nc), max(4 * C // 4, 64)
    for name, cm in (("cls", c1), ("box", c2)):
        for i in range(3):
            _cbam_keys(t, f"head.{name}.{i}.1", cm)
            _cbam_keys(t, f"head.{name}.{i}.3", cm)
    return t


def make_attention_state_dict(nc: int, C: int, d: int, deep_mul: float, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Weights of the attention blocks: N(0, sqrt(2 / fan_in)) convs, small biases, gamma = 0.5 (the reference's zero
    initialisation would switch the SelfAttention term off: SURVEY F10)."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, shape in attention_shapes(nc, C, d, deep_mul).items():
        r = _rng(seed, key)
        if key.endswith("gamma"):
            v = np.full(shape, 0.5, np.float32)
        elif key.endswith("bias"):
            v = (0.1 * r.standard_normal(shape)).astype(np.float32)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            scale = np.sqrt(2.0 / fan_in) * (0.5 if ".query." in key or ".key." in key else 1.0)
            v = (r.standard_normal(shape) * scale).astype(np.float32)
        out[key] = v
    return out


def state_dict_shapes(nc: int, C: int, d: int, deep_mul: float) -> "OrderedDict[str, tuple]":
    """Key -> shape for the plain-topology detector, in reference module order."""
    C5 = int(C * 16 * deep_mul)
    t: "OrderedDict[str, tuple]" = OrderedDict()
    # backbone (model/backbone.py:20-48)
    _conv_keys(t, "backbone.stem", 3, C, 3)
    _conv_keys(t, "backbone.dark2.0", C, 2 * C, 3)
    _c2f_keys(t, "backbone.dark2.1", 2 * C, 2 * C, d)
    _conv_keys(t, "backbone.dark3.0", 2 * C, 4 * C, 3)
    _c2f_keys(t, "backbone.dark3.1", 4 * C, 4 * C, 2 * d)
    _conv_keys(t, "backbone.dark4.0", 4 * C, 8 * C, 3)
    _c2f_keys(t, "backbone.dark4.1", 8 * C, 8 * C, 2 * d)
    _conv_keys(t, "backbone.dark5.0", 8 * C, C5, 3)
    _c2f_keys(t, "backbone.dark5.1", C5, C5, d)
    _conv_keys(t, "backbone.dark5.2.cv1", C5, C5 // 2, 1)       # SPPF (model/blocks.py:132-135)
    _conv_keys(t, "backbone.dark5.2.cv2", (C5 // 2) * 4, C5, 1)
    # neck (model/neck.py:19-53, stages = C2f(..., shortcut=False) per SURVEY F4)
    _c2f_keys(t, "neck.h1", C5 + 8 * C, 8 * C, d)
    _c2f_keys(t, "neck.h2", 8 * C + 4 * C, 4 * C, d)
    _conv_keys(t, "neck.h3", 4 * C, 4 * C, 3)
    _c2f_keys(t, "neck.h4", 8 * C + 4 * C, 8 * C, d)
    _conv_keys(t, "neck.h5", 8 * C, 8 * C, 3)
    _c2f_keys(t, "neck.h6", C5 + 8 * C, C5, d)
    # head (model/head.py:19-44); module registration order: dfl, cls, box
    filters = (4 * C, 8 * C, C5)
    c1 = max(filters[0], nc)
    c2 = max(filters[0] // 4, 64)
    t["head.dfl.conv.weight"] = (1, 16, 1, 1)
    for name, cm, co in (("cls", c1, nc), ("box", c2, 64)):
        for i, f in enumerate(filters):
            _conv_keys(t, f"head.{name}.{i}.0", f, cm, 3)
            _conv_keys(t, f"head.{name}.{i}.2", cm, cm, 3)
            t[f"head.{name}.{i}.4.weight"] = (co, cm, 1, 1)
            t[f"head.{name}.{i}.4.bias"] = (co,)
    return t


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def make_state_dict(nc: int, C: int, d: int, deep_mul: float, seed: int = 0,
                    cls_bias_mean: float = -5.0, cls_bias_std: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """Synthetic but well-conditioned weights.

    conv weights: N(0, sqrt(2/fan_in)) -- the distribution of the reference's own
    weights_init(net, 'kaiming') (model/train_utils.py:108-120); BN weight N(1, 0.02)
    (train_utils.py:123-125); BN bias/mean N(0, 0.1), var U(0.5, 1.5) so that BN folding is
    exercised; DFL projection arange(16) (model/blocks.py:150-152); box bias 1.0
    (model/head.py:67); class bias N(cls_bias_mean, cls_bias_std) so that some anchors clear
    the confidence thresholds.
    """
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, shape in state_dict_shapes(nc, C, d, deep_mul).items():
        r = _rng(seed, key)
        if key == "head.dfl.conv.weight":
            v = np.arange(16, dtype=np.float32).reshape(shape)
        elif key.endswith("num_batches_tracked"):
            v = np.zeros((), dtype=np.int64)
        elif key.endswith("conv.weight") or key.endswith(".4.weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            v = (r.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif key.endswith("norm.weight"):
            v = (1.0 + 0.02 * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("norm.bias") or key.endswith("running_mean"):
            v = (0.1 * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("running_var"):
            v = r.uniform(0.5, 1.5, shape).astype(np.float32)
        elif key.endswith(".4.bias"):
            if ".box." in key:
                v = np.ones(shape, dtype=np.float32)
            else:
                v = (cls_bias_mean + cls_bias_std * r.standard_normal(shape)).astype(np.float32)
        else:  # pragma: no cover
            raise KeyError(key)
        out[key] = v
    return out


def make_images(batch: int, h: int, w: int, seed: int) -> np.ndarray:
    """float32 NCHW images in [0, 1) (what preprocess_input produces, utils/utils.py:65-67)."""
    r = np.random.Generator(np.random.PCG64([seed, 0x1A6E5]))
    return r.random((batch, 3, h, w), dtype=np.float32)


def make_images_u8(batch: int, h: int, w: int, seed: int) -> np.ndarray:
    """uint8 NHWC (B, H, W, 3) letterboxed-image stand-ins; the reference's tensor for the same pixels is
    `images_u8_to_f32` of it (np.array(image, float32) / 255.0, HWC -> CHW, utils/callbacks.py:142-144)."""
    r = np.random.Generator(np.random.PCG64([seed, 0x08E5]))
    return r.integers(0, 256, (batch, h, w, 3), dtype=np.uint8)


def images_u8_to_f32(u8: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.transpose(u8.astype(np.float32) / np.float32(255.0), (0, 3, 1, 2)))


def make_dense_predictions(batch: int, anchors: int = 8400, nc: int = 80, objects: int = 120,
                           seed: int = 1234) -> np.ndarray:
    """clustered boxes, (B, A, 4+nc) float32 normalised xywh + scores."""
    r = np.random.Generator(np.random.PCG64([seed, 0xD5E5]))
    pred = np.empty((batch, anchors, 4 + nc), dtype=np.float32)
    for b in range(batch):
        cxy = r.uniform(0.05, 0.95, (objects, 2))
        wh = np.exp(r.uniform(np.log(0.03), np.log(0.5), (objects, 2)))
        cls = r.integers(0, nc, objects)
        pick = r.integers(0, objects, anchors)
        box_c = cxy[pick] + r.standard_normal((anchors, 2)) * 0.02 * wh[pick]
        box_s = wh[pick] * np.exp(r.standard_normal((anchors, 2)) * 0.1)
        scores = r.uniform(0.0, 8e-4, (anchors, nc))
        scores[np.arange(anchors), cls[pick]] = 0.001 + 0.999 * r.random(anchors) ** 2
        pred[b, :, 0:2] = box_c
        pred[b, :, 2:4] = box_s
        pred[b, :, 4:] = scores
    return pred
