"""B200-native (sm_100a) implementation of the detector inference hot path of
mohamed22311/Transparent-Object-Detection: forward -> decode -> NMS behind the reference's call surface.

    from transparent_object_detection_b200 import BaseModel, DecodeBox, Detector

All arithmetic runs in libtod.so (hand-written CUDA, C ABI in include/tod.h); there is no CPU or
alternative-backend fallback.
"""
from .model import BaseModel, DecodeBox, Detector, LossDecode, parameter_table  # noqa: F401
from .engine import DetectorEngine, fold_conv_bn, pack_conv_weight  # noqa: F401
from ._lib import lib, LIB_PATH, SYMBOLS, TodError  # noqa: F401
from .sharding import shard_range, weighted_shard_ranges, gather_detections  # noqa: F401

__all__ = ["BaseModel", "DecodeBox", "Detector", "LossDecode", "DetectorEngine", "lib", "LIB_PATH", "SYMBOLS", "TodError",
           "shard_range", "weighted_shard_ranges", "gather_detections"]
