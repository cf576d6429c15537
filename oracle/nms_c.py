"""ctypes wrapper of oracle/nms_oracle.c (TEST INFRASTRUCTURE, see the header of that file)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnms_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "nms_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s", "-B", "libnms_oracle.so"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.tod_oracle_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
        L.tod_oracle_nms.restype = C.c_int32
        L.tod_oracle_nms_image.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_double, C.c_void_p]
        L.tod_oracle_nms_image.restype = C.c_int32
        _lib = L
    return _lib


def nms(boxes: np.ndarray, scores: np.ndarray, iou_thr: float) -> np.ndarray:
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    s = np.ascontiguousarray(scores, dtype=np.float32)
    keep = np.empty((max(len(s), 1),), dtype=np.int32)
    n = lib().tod_oracle_nms(b.ctypes.data, s.ctypes.data, len(s), float(iou_thr), keep.ctypes.data)
    return keep[:n].astype(np.int64)


def nms_keep_indices(prediction: np.ndarray, num_classes: int, conf_thres: float, nms_thres: float):
    """Same contract as detector_oracle.nms_keep_indices: list of kept anchor indices per image."""
    pred = np.ascontiguousarray(prediction, dtype=np.float32)
    out = []
    for i in range(pred.shape[0]):
        keep = np.empty((pred.shape[1],), dtype=np.int32)
        n = lib().tod_oracle_nms_image(pred[i].ctypes.data, pred.shape[1], num_classes, float(np.float32(conf_thres)),
                                       float(nms_thres), keep.ctypes.data)
        out.append(keep[:n].astype(np.int64))
    return out
