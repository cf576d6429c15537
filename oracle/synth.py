"""Synthetic weights / images / predictions: moved to transparent_object_detection_b200/synth.py (input generation is not
checker code, and bench.py's product arm must not import oracle/).  This shim keeps `from oracle import synth` working for
the tests and the golden-vector scripts."""
from transparent_object_detection_b200.synth import *          # noqa: F401,F403
from transparent_object_detection_b200.synth import SCALES, _rng, _conv_keys, _c2f_keys, _cbam_keys  # noqa: F401
